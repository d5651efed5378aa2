// gather_peak.cu -- what can random whole-row gathers reach on this HBM?  (round 2)
// The traversal kernel (K3) moves 3 KB rows picked by a dependent chain; this probe issues the SAME copies with no chain at all
// (ids from a hash), so its GB/s is the ceiling K3's 0.86 of the STREAM-copy peak has to be read against -- and it compares the
// two ways of issuing a row gather on sm_100a:
//   mode 0  cp.async.bulk (1-D, one copy per row, SASS UBLKCP)            -- what K3 uses
//   mode 1  cp.async.bulk.tensor.2d.tile::gather4, tensor-map box {256,1}  -- 4 rows x 256 columns per instruction (UTMALDG)
//   mode 2  the same with box {256,4}
// usage: gather_peak <mode> [warps per CTA = 8] [rings per warp = 1] [rows = 1000000] [dim = 768] [iters = 400]
// prints one JSON line.  Every landed row is checked against its id (first float of the row), so a mode that does not do what
// this file thinks it does reports "ok": false instead of a bandwidth.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s: %s\"}\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t par)
{
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(par) : "memory");
	return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, uint32_t par)
{
	for (uint32_t i = 0; i < (1u << 22); i++)
		if (mbar_try(bar, par)) return true;
	return false;
}
__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }

struct P
{
	const float *table;
	uint32_t	 n_rows, row_f, iters, rings, mode;
	unsigned long long *bytes;
	int			*bad;
};

__global__ void __launch_bounds__(1024) gather_kernel(const __grid_constant__ CUtensorMap tm, const P p)
{
	extern __shared__ __align__(1024) unsigned char smem[];
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
	const uint32_t row_b = p.row_f * 4u;
	const uint32_t ring_b = 8u * row_b;	 // 8 rows per ring; gather4 layout: [2 groups of 4 rows][3 boxes][4][256] = same size at dim 768
	uint64_t	  *bars = reinterpret_cast<uint64_t *>(smem);
	unsigned char *rings = smem + 1024 + (size_t) warp * p.rings * ring_b;
	if (lane == 0)
		for (uint32_t r = 0; r < p.rings; r++) mbar_init(&bars[warp * 4 + r], 1);
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncthreads();
	const uint64_t pol = pol_first();
	const uint32_t gw = blockIdx.x * nw + warp;
	uint32_t	   par[4] = {0, 0, 0, 0};
	uint32_t	   ids[4][8];  // per ring, this lane's copy of the 8 ids (all lanes compute the same)
	unsigned long long moved = 0;
	auto id_of = [&](uint32_t it, uint32_t j) { return (uint32_t) (((uint64_t) (gw * 2654435761u + it * 40503u + j * 2246822519u + (it * 7u + j) * (it + 13u)) * 2654435761ull >> 16) % p.n_rows); };
	auto issue = [&](uint32_t r, uint32_t it) {
		unsigned char *ring = rings + (size_t) r * ring_b;
		uint64_t	  *bar = &bars[warp * 4 + r];
#pragma unroll
		for (uint32_t j = 0; j < 8; j++) ids[r][j] = id_of(it, j);
		if (lane == 0) mbar_expect(bar, ring_b);
		__syncwarp();
		if (p.mode == 0)
		{
			if (lane < 8)
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(ring + lane * row_b)),
							 "l"(p.table + (size_t) ids[r][lane] * p.row_f), "r"(row_b), "r"(smem_u32(bar)), "l"(pol)
							 : "memory");
		}
		else
		{
			// 2 groups of 4 rows x (row_f / 256) column boxes: one instruction per (group, box)
			const uint32_t nbox = p.row_f / 256u;
			if (lane < 2u * nbox)
			{
				const uint32_t g = lane / nbox, c = lane % nbox;
				asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;" ::"r"(
								 smem_u32(ring + (size_t) (g * nbox + c) * 4096u)),
							 "l"(&tm), "r"(smem_u32(bar)), "r"((int32_t) (c * 256u)), "r"((int32_t) ids[r][g * 4 + 0]), "r"((int32_t) ids[r][g * 4 + 1]),
							 "r"((int32_t) ids[r][g * 4 + 2]), "r"((int32_t) ids[r][g * 4 + 3]), "l"(pol)
							 : "memory");
			}
		}
	};
	for (uint32_t r = 0; r < p.rings && r < p.iters; r++) issue(r, r);
	for (uint32_t it = 0; it < p.iters; it++)
	{
		const uint32_t r = it % p.rings;
		if (!mbar_wait_bounded(&bars[warp * 4 + r], par[r]))
		{
			if (lane == 0) atomicExch(p.bad, 2);
			break;
		}
		par[r] ^= 1u;
		// check: the first float of every landed row is its id
		if (lane < 8)
		{
			const unsigned char *ring = rings + (size_t) r * ring_b;
			const uint32_t		 nbox = p.row_f / 256u;
			const float			*rowp = p.mode == 0 ? reinterpret_cast<const float *>(ring + lane * row_b)
												   : reinterpret_cast<const float *>(ring + (size_t) ((lane / 4) * nbox) * 4096u + (lane % 4) * 1024u);
			if (rowp[0] != (float) ids[r][lane] || rowp[255] != (float) ids[r][lane] + 255.0f) atomicExch(p.bad, 1);
		}
		moved += ring_b;
		__syncwarp();
		if (it + p.rings < p.iters) issue(r, it + p.rings);
	}
	if (lane == 0) atomicAdd(p.bytes, moved);
}

__global__ void fill_kernel(float *t, uint32_t n_rows, uint32_t row_f)
{
	const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < (size_t) n_rows * row_f) t[i] = (float) (i / row_f) + (float) (i % row_f);
}

int main(int argc, char **argv)
{
	const uint32_t mode = argc > 1 ? atoi(argv[1]) : 0, warps = argc > 2 ? atoi(argv[2]) : 8, rings = argc > 3 ? atoi(argv[3]) : 1;
	const uint32_t n_rows = argc > 4 ? atoi(argv[4]) : 1000000, dim = argc > 5 ? atoi(argv[5]) : 768, iters = argc > 6 ? atoi(argv[6]) : 400;
	if (rings < 1 || rings > 4 || warps < 1 || warps > 32 || dim % 256) { printf("{\"error\": \"bad arguments\"}\n"); return 2; }
	float *table;
	CK(cudaMalloc(&table, (size_t) n_rows * dim * 4));
	fill_kernel<<<(unsigned) (((size_t) n_rows * dim + 255) / 256), 256>>>(table, n_rows, dim);
	CK(cudaDeviceSynchronize());
	CUtensorMap tm;
	memset(&tm, 0, sizeof(tm));
	if (mode != 0)
	{
		void						   *fn = nullptr;
		cudaDriverEntryPointQueryResult qr;
		CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
		typedef CUresult (*enc_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
								  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
		const cuuint64_t gdim[2] = {dim, n_rows}, gstr[1] = {(cuuint64_t) dim * 4};
		const cuuint32_t box[2] = {256, mode == 1 ? 1u : 4u}, estr[2] = {1, 1};
		const CUresult	 r = ((enc_t) fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, table, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
										  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r != CUDA_SUCCESS) { printf("{\"mode\": %u, \"error\": \"cuTensorMapEncodeTiled %d\"}\n", mode, (int) r); return 3; }
	}
	unsigned long long *d_bytes;
	int				   *d_bad;
	CK(cudaMalloc(&d_bytes, 8));
	CK(cudaMalloc(&d_bad, 4));
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	const size_t smem = 1024 + (size_t) warps * rings * 8 * dim * 4;
	if (smem > 232448) { printf("{\"error\": \"shared memory %zu\"}\n", smem); return 2; }
	CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	P p = {table, n_rows, dim, iters, rings, mode, d_bytes, d_bad};
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	float best = 1e30f;
	unsigned long long bytes = 0;
	int				   bad = 0;
	for (int rep = 0; rep < 4; rep++)
	{
		CK(cudaMemset(d_bytes, 0, 8));
		CK(cudaMemset(d_bad, 0, 4));
		CK(cudaEventRecord(e0));
		gather_kernel<<<prop.multiProcessorCount, warps * 32, smem>>>(tm, p);
		CK(cudaEventRecord(e1));
		CK(cudaDeviceSynchronize());
		float ms;
		CK(cudaEventElapsedTime(&ms, e0, e1));
		CK(cudaMemcpy(&bytes, d_bytes, 8, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
		if (rep > 0 && ms < best) best = ms;
		if (bad) break;
	}
	printf("{\"mode\": %u, \"warps\": %u, \"rings\": %u, \"rows\": %u, \"dim\": %u, \"iters\": %u, \"ok\": %s, \"bad\": %d, \"ms\": %.3f, \"gbs\": %.1f, \"bytes_in_flight_per_sm\": %zu}\n", mode,
		   warps, rings, n_rows, dim, iters, bad ? "false" : "true", bad, best, bad ? 0.0 : (double) bytes / (best * 1e-3) / 1e9, (size_t) warps * rings * 8 * dim * 4);
	return bad ? 1 : 0;
}
