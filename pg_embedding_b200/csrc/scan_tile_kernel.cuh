// scan_tile_kernel.cuh -- tiled form of the exact scan's distance step (`val <op> q` under a
// sequential scan, embedding.c:1022-1062), same output as scan_dist_kernel: out[q * nr + j] = dist(query q, node r0 + j).
//
// scan_dist_kernel gives every (query,row) pair its own LANES threads, so each row is re-read from L2/HBM once per
// query and every pair pays a shuffle reduction.  Here a CTA keeps a tile of 64 rows x 64 dimensions and a tile of
// queries in shared memory and every thread owns a small block of pairs in registers:
//     cosine / manhattan: 4 queries x 2 rows x 4 accumulator chains   (CTA: 32 queries x 64 rows)
//     L2:                 2 queries x 2 rows x 8 accumulator chains   (CTA: 16 queries x 64 rows)
// A thread owns ALL accumulator chains of its pairs, so the reference's lane order needs no shuffles: chain j of a
// pair sums the elements i == j (mod 4) -- for L2 the 16-float blocks S[j] += d[j]^2 + d[8+j]^2 -- in increasing
// i across the dimension chunks, then the same horizontal sum, tail and epilogue as dist_exact.cuh.  Same rounding
// sequence per pair => bit-identical distances; only the schedule differs.
//
// Shared-memory traffic decides the tile shape: the queries of a warp are warp-uniform (broadcast loads, one
// wavefront), the rows differ per lane (pitch == 16 mod 128 bytes: conflict-free LDS.128).  Per 4 dimensions a
// cosine thread issues 4 + 2 vector loads for 64 mul/add instructions, so the ALU, not the LSU, is the limit.
#pragma once
#include "common.cuh"
#include "dist_exact.cuh"

namespace pgemb {

constexpr int kScanTileRows = 64;	 // rows per CTA tile (lane, lane + 32)
constexpr int kScanTileK = 64;		 // dimensions per chunk
constexpr int kScanRowPitch = kScanTileK + 4;  // floats; 272 B == 16 (mod 128)
constexpr int kScanThreads = 256;

template <int METRIC> struct ScanTile
{
	static constexpr int QPT = (METRIC == M_L2) ? 2 : 4;   // queries per thread (= per warp)
	static constexpr int TQ = QPT * (kScanThreads / 32);  // queries per CTA
};

template <int METRIC>
__global__ void __launch_bounds__(kScanThreads) scan_tile_kernel(const float *__restrict__ vectors, const float *__restrict__ norms,
																 uint32_t row_f, uint32_t dim, const float *__restrict__ queries,
																 uint32_t q_stride, const float *__restrict__ qnorms, uint32_t nq,
																 uint32_t r0, uint32_t nr, float *__restrict__ out)
{
	constexpr int QPT = ScanTile<METRIC>::QPT;
	constexpr int TQ = ScanTile<METRIC>::TQ;
	constexpr int NCH = (METRIC == M_L2) ? 8 : 4;			// accumulator chains per pair
	constexpr int RLD = kScanTileRows * (kScanTileK / 4) / kScanThreads;  // float4 row loads per thread per chunk (4)
	constexpr int QLD = TQ * kScanTileK / kScanThreads;					  // scalar query loads per thread per chunk
	__shared__ __align__(16) float rows_s[kScanTileRows * kScanRowPitch];
	__shared__ __align__(16) float qs_s[TQ * kScanTileK];

	const int	   lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	// x = query tile (fastest: CTAs that share a row tile run together and find it in L2), y = row tile (<= 256 per chunk)
	const uint32_t q_base = blockIdx.x * TQ;
	const uint32_t row_base = blockIdx.y * kScanTileRows;  // tile's first row, relative to r0
	const int	   main_n = main_len<METRIC>((int) dim);
	if (nr == 0 || nq == 0) return;

	float acc[QPT][2][NCH];
#pragma unroll
	for (int a = 0; a < QPT; a++)
#pragma unroll
		for (int b = 0; b < 2; b++)
#pragma unroll
			for (int c = 0; c < NCH; c++) acc[a][b][c] = 0.0f;

	// register staging of the next chunk (rows: float4, rows are 16-byte aligned; queries: scalars, any stride).
	// Everything that does not depend on the chunk -- source pointers, shared-memory slots -- is computed once here: the
	// per-chunk staging is then a handful of loads and stores next to the ~1000 arithmetic instructions of a chunk.
	float4		 rreg[RLD];
	float		 qreg[QLD];
	const float *rsrc[RLD];
	float4		*rdst[RLD];
	int			 rcol[RLD];
#pragma unroll
	for (int u = 0; u < RLD; u++)
	{
		const int f = (int) threadIdx.x + kScanThreads * u;
		const int r = f / (kScanTileK / 4), c4 = f % (kScanTileK / 4);
		uint32_t  rr = row_base + (uint32_t) r;
		if (rr >= nr) rr = nr - 1;	// clamp: padding rows are computed and dropped
		rcol[u] = c4 * 4;
		rsrc[u] = vectors + (size_t) (r0 + rr) * row_f + c4 * 4;
		rdst[u] = reinterpret_cast<float4 *>(&rows_s[r * kScanRowPitch + c4 * 4]);
	}
	const float *qsrc[QLD];
	int			 qcol[QLD];
#pragma unroll
	for (int u = 0; u < QLD; u++)
	{
		const int e = (int) threadIdx.x + kScanThreads * u;
		const int q = e / kScanTileK, c = e % kScanTileK;
		uint32_t  qq = q_base + (uint32_t) q;
		if (qq >= nq) qq = nq - 1;
		qcol[u] = c;
		qsrc[u] = queries + (size_t) qq * q_stride + c;
	}
	auto fetch = [&](int kc) {
#pragma unroll
		for (int u = 0; u < RLD; u++)
			rreg[u] = (kc + rcol[u] < main_n) ? __ldg(reinterpret_cast<const float4 *>(rsrc[u] + kc)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
		for (int u = 0; u < QLD; u++) qreg[u] = (kc + qcol[u] < main_n) ? __ldg(qsrc[u] + kc) : 0.0f;
	};
	auto stash = [&]() {
#pragma unroll
		for (int u = 0; u < RLD; u++) *rdst[u] = rreg[u];
#pragma unroll
		for (int u = 0; u < QLD; u++) qs_s[(int) threadIdx.x + kScanThreads * u] = qreg[u];  // e = q * kScanTileK + c
	};

	const float *rp0 = &rows_s[lane * kScanRowPitch];
	const float *rp1 = &rows_s[(lane + 32) * kScanRowPitch];
	const float *qp = &qs_s[warp * QPT * kScanTileK];

	if (main_n > 0) fetch(0);
	for (int kc = 0; kc < main_n; kc += kScanTileK)
	{
		__syncthreads();  // everyone is done with the previous chunk in shared memory
		stash();
		__syncthreads();
		if (kc + kScanTileK < main_n) fetch(kc + kScanTileK);  // in flight while this chunk is scored
		const int kn = min(kScanTileK, main_n - kc);
		if (METRIC == M_L2)
		{
			// 16-float blocks: S[j] += (x[j]-y[j])^2 + (x[8+j]-y[8+j])^2, j = 0..7 (distfunc.c:44-56 as vectorised)
			for (int i = 0; i < kn; i += 16)
			{
#pragma unroll
				for (int h = 0; h < 2; h++)
				{
					float4 x0[QPT], x1[QPT], y0[2], y1[2];
#pragma unroll
					for (int a = 0; a < QPT; a++)
					{
						x0[a] = *reinterpret_cast<const float4 *>(qp + a * kScanTileK + i + h * 4);
						x1[a] = *reinterpret_cast<const float4 *>(qp + a * kScanTileK + i + 8 + h * 4);
					}
					y0[0] = *reinterpret_cast<const float4 *>(rp0 + i + h * 4);
					y1[0] = *reinterpret_cast<const float4 *>(rp0 + i + 8 + h * 4);
					y0[1] = *reinterpret_cast<const float4 *>(rp1 + i + h * 4);
					y1[1] = *reinterpret_cast<const float4 *>(rp1 + i + 8 + h * 4);
#pragma unroll
					for (int a = 0; a < QPT; a++)
#pragma unroll
						for (int b = 0; b < 2; b++)
						{
#define PGEMB_L2_LANE(C, F)                                                                            \
	{                                                                                                  \
		const float d0 = __fsub_rn(x0[a].F, y0[b].F), d1 = __fsub_rn(x1[a].F, y1[b].F);                \
		acc[a][b][h * 4 + C] = __fadd_rn(acc[a][b][h * 4 + C], __fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1))); \
	}
							PGEMB_L2_LANE(0, x)
							PGEMB_L2_LANE(1, y)
							PGEMB_L2_LANE(2, z)
							PGEMB_L2_LANE(3, w)
#undef PGEMB_L2_LANE
						}
				}
			}
		}
		else
		{
			constexpr int TERM = (METRIC == M_COS) ? 0 : 1;
#pragma unroll 4
			for (int i = 0; i < kn; i += 4)
			{
				float4 qv[QPT], rv[2];
#pragma unroll
				for (int a = 0; a < QPT; a++) qv[a] = *reinterpret_cast<const float4 *>(qp + a * kScanTileK + i);
				rv[0] = *reinterpret_cast<const float4 *>(rp0 + i);
				rv[1] = *reinterpret_cast<const float4 *>(rp1 + i);
#pragma unroll
				for (int a = 0; a < QPT; a++)
#pragma unroll
					for (int b = 0; b < 2; b++)
					{
						acc[a][b][0] = __fadd_rn(acc[a][b][0], term4<TERM>(qv[a].x, rv[b].x));
						acc[a][b][1] = __fadd_rn(acc[a][b][1], term4<TERM>(qv[a].y, rv[b].y));
						acc[a][b][2] = __fadd_rn(acc[a][b][2], term4<TERM>(qv[a].z, rv[b].z));
						acc[a][b][3] = __fadd_rn(acc[a][b][3], term4<TERM>(qv[a].w, rv[b].w));
					}
			}
		}
	}

	// ---- horizontal sum, tail, epilogue: exactly dist_exact.cuh's, one pair at a time ------------------
#pragma unroll
	for (int a = 0; a < QPT; a++)
	{
		const uint32_t q = q_base + (uint32_t) (warp * QPT + a);
		if (q >= nq) continue;
		const float *qg = queries + (size_t) q * q_stride;
#pragma unroll
		for (int b = 0; b < 2; b++)
		{
			const uint32_t j = row_base + (uint32_t) (lane + 32 * b);
			if (j >= nr) continue;
			const float *vg = vectors + (size_t) (r0 + j) * row_f;
			float		 d;
			if (METRIC == M_L2)
			{
				float res = hsum4(__fadd_rn(acc[a][b][0], acc[a][b][4 % NCH]), __fadd_rn(acc[a][b][1], acc[a][b][5 % NCH]),
								  __fadd_rn(acc[a][b][2], acc[a][b][6 % NCH]), __fadd_rn(acc[a][b][3], acc[a][b][7 % NCH]));
				res = l2_tail_exact(res, qg + main_n, vg + main_n, (int) dim - main_n);
				d = __fsqrt_rn(res);
			}
			else
			{
				constexpr int TERM = (METRIC == M_COS) ? 0 : 1;
				float		  res = hsum4(acc[a][b][0], acc[a][b][1], acc[a][b][2], acc[a][b][3]);
				for (int e = main_n; e < (int) dim; e++) res = __fadd_rn(res, term4<TERM>(qg[e], vg[e]));
				d = (METRIC == M_COS) ? cosine_finish(res, qnorms[q], norms[r0 + j]) : res;
			}
			out[(size_t) q * nr + j] = d;
		}
	}
}

}  // namespace pgemb
