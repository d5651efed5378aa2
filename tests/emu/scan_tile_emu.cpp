// Runs pgemb::scan_tile_kernel on the host, one pthread per CUDA thread, one CTA at a time (tests/emu).
#include <cuda_runtime.h>  // the stand-in in tests/emu/fake_cuda
#include <thread>
#include <vector>

thread_local dim3 threadIdx, blockIdx;
dim3			  blockDim, gridDim;
pthread_barrier_t g_cta_barrier;

#include "../../pg_embedding_b200/csrc/scan_tile_kernel.cuh"

using namespace pgemb;

template <int METRIC>
static void run(const float *vectors, const float *norms, uint32_t row_f, uint32_t dim, const float *queries, uint32_t q_stride,
				const float *qnorms, uint32_t nq, uint32_t r0, uint32_t nr, float *out)
{
	gridDim.x = (nq + ScanTile<METRIC>::TQ - 1) / ScanTile<METRIC>::TQ;
	gridDim.y = (nr + kScanTileRows - 1) / kScanTileRows;
	blockDim.x = kScanThreads;
	pthread_barrier_init(&g_cta_barrier, nullptr, kScanThreads);
	for (unsigned by = 0; by < gridDim.y; by++)
		for (unsigned bx = 0; bx < gridDim.x; bx++)
		{
			std::vector<std::thread> th;
			for (unsigned t = 0; t < (unsigned) kScanThreads; t++)
				th.emplace_back([=]() {
					threadIdx.x = t;
					blockIdx.x = bx;
					blockIdx.y = by;
					scan_tile_kernel<METRIC>(vectors, norms, row_f, dim, queries, q_stride, qnorms, nq, r0, nr, out);
				});
			for (auto &x : th) x.join();
		}
	pthread_barrier_destroy(&g_cta_barrier);
}

extern "C" void emu_scan_tile(int metric, const float *vectors, const float *norms, uint32_t row_f, uint32_t dim, const float *queries,
							  uint32_t q_stride, const float *qnorms, uint32_t nq, uint32_t r0, uint32_t nr, float *out)
{
	if (metric == 0) run<M_L2>(vectors, norms, row_f, dim, queries, q_stride, qnorms, nq, r0, nr, out);
	else if (metric == 1) run<M_COS>(vectors, norms, row_f, dim, queries, q_stride, qnorms, nq, r0, nr, out);
	else run<M_MAN>(vectors, norms, row_f, dim, queries, q_stride, qnorms, nq, r0, nr, out);
}
