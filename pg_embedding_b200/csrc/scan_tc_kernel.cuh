// scan_tc_kernel.cuh -- exact scan with a tensor-core FILTER (prototype, -DPGEMB_PROTO, opt-in PGEMB_SCAN_TC=1).
//
// the brute-force operator path (`ORDER BY val <op> q LIMIT k` without the index,
// embedding.c:1022-1062) is the one place of this extension where a dense  queries x rows  contraction exists.  A
// tensor-core GEMM cannot give the reference's bits (TF32 drops 13 mantissa bits, and the reference's summation order is
// fixed), so it is used only to DISCARD rows: per chunk of rows
//     1. S = Q . V^T                     one TF32 GEMM (cuBLAS, fp32 accumulate) -- approximate dot products;
//     2. scan_select_tc_kernel (below)   per query, a row is a candidate unless a rigorous lower bound of its distance,
//                                        computed from S and the exact squared norms, already exceeds the current k-th best
//                                        exact distance; candidates are re-scored with the reference-exact arithmetic
//                                        (dist_exact.cuh) inside the same warp and folded into the running top-k by
//                                        (dist,label) exactly like scan_select_kernel does.
// The result is identical to pgemb_scan_topk's exact path -- same labels, same order, bit-identical distances -- as long
// as the error bound holds:  |S - q.v| <= rel_err * |q| |v|  (Cauchy-Schwarz bounds sum |q_i v_i|).  rel_err is chosen by the
// host for TF32 (both operands cut to 10 mantissa bits: 2 * 2^-10, plus fp32 accumulation over `dim` terms, plus slack).
// Every re-scored candidate also CHECKS the bound against its exact distance and raises *viol if it is ever exceeded;
// the host then repeats the scan on the exact path and reports it -- a wrong assumption about the library's
// arithmetic cannot silently change results.
//
// cosine:    dist = 1 - q.v / sqrt(|q|^2 |v|^2)          lower bound: (1 - S/scale) - rel_err - 1e-4
// L2:        dist^2 = |q|^2 + |v|^2 - 2 q.v               lower bound: sqrt(max(0, d2 - 2 rel_err scale - 1e-4 (|q|^2+|v|^2)))
// manhattan: no bilinear form -- not supported here (the host keeps the exact path).
#pragma once
#include "common.cuh"
#include "dist_exact.cuh"

namespace pgemb {

template <int METRIC>
__global__ void scan_select_tc_kernel(const float *__restrict__ S, const float *__restrict__ vectors, uint32_t row_f, uint32_t dim,
									  const float *__restrict__ vnorm2 /* [nr] squared norms of rows r0.. (cosine: the cached ones) */,
									  const float *__restrict__ queries, uint32_t q_stride, const float *__restrict__ qnorm2,
									  const uint64_t *__restrict__ labels, uint32_t nq, uint32_t r0, uint32_t nr, uint32_t k, float rel_err,
									  uint32_t *__restrict__ top_d, uint64_t *__restrict__ top_l, uint32_t *__restrict__ top_n,
									  uint32_t *__restrict__ tmp_d, uint64_t *__restrict__ tmp_l, uint32_t *__restrict__ rescored,
									  int *__restrict__ viol)
{
	static_assert(METRIC == M_L2 || METRIC == M_COS, "the filter needs a bilinear form");
	constexpr int TPR = MetricLanes<METRIC>::LANES;	 // lanes per exact pair: 8 (L2) or 4 (cosine)
	constexpr int G = 32 / TPR;						 // candidates re-scored concurrently by one warp
	__shared__ uint32_t cd[4][kScanCand];
	__shared__ uint64_t cl[4][kScanCand];
	const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t q = blockIdx.x * 4 + w;
	if (q >= nq) return;
	uint32_t *td = top_d + (size_t) q * k, *sd = tmp_d + (size_t) q * k;
	uint64_t *tl = top_l + (size_t) q * k, *sl = tmp_l + (size_t) q * k;
	uint32_t  n = top_n[q];
	uint32_t  nc = 0, n_resc = 0;
	const uint32_t lt = (1u << lane) - 1u;
	const int	   grp = (int) lane / TPR, sub = (int) lane % TPR;
	const float	  *qp = queries + (size_t) q * q_stride;
	const float	   qn = qnorm2[q];
	auto less = [](uint32_t d1, uint64_t l1, uint32_t d2, uint64_t l2) { return d1 < d2 || (d1 == d2 && l1 < l2); };
	auto merge = [&]() {
		// rank every element of top (n) and cand (nc) in their union; keep ranks < k  (as scan_select_kernel)
		__syncwarp();
		const uint32_t total = n + nc;
		for (uint32_t i = lane; i < total; i += 32)
		{
			const bool	   from_top = i < n;
			const uint32_t d = from_top ? td[i] : cd[w][i - n];
			const uint64_t l = from_top ? tl[i] : cl[w][i - n];
			uint32_t	   rank = 0;
			for (uint32_t j = 0; j < n; j++) rank += (j != i && (less(td[j], tl[j], d, l) || (!less(d, l, td[j], tl[j]) && j < i))) ? 1u : 0u;
			for (uint32_t j = 0; j < nc; j++)
			{
				const uint32_t jj = n + j;
				rank += (jj != i && (less(cd[w][j], cl[w][j], d, l) || (!less(d, l, cd[w][j], cl[w][j]) && jj < i))) ? 1u : 0u;
			}
			if (rank < k) { sd[rank] = d; sl[rank] = l; }
		}
		__syncwarp();
		n = total < k ? total : k;
		for (uint32_t i = lane; i < n; i += 32) { td[i] = sd[i]; tl[i] = sl[i]; }
		nc = 0;
		__syncwarp();
	};
	for (uint32_t base = 0; base < nr; base += 32)
	{
		const uint32_t j = base + lane;
		bool		   maybe = false;
		uint64_t	   l = 0;
		float		   approx = 0.f, slack = 0.f;  // approx: distance (cosine) or squared distance (L2) from the GEMM; slack: its error bound
		if (j < nr)
		{
			l = labels[r0 + j];
			if (((l >> 48) & 1ull) == 0)
			{
				const float s = S[(size_t) q * nr + j];
				const float vn = vnorm2[j];
				const float scale = sqrtf(qn * vn);	 // >= sum |q_i v_i| (Cauchy-Schwarz), up to rounding covered by the slack
				float		lb;
				if (METRIC == M_COS)
				{
					approx = 1.0f - s / scale;
					slack = rel_err + 1e-4f;
					lb = approx - slack;
				}
				else
				{
					approx = qn + vn - 2.0f * s;
					slack = 2.0f * rel_err * scale + 1e-4f * (qn + vn);
					const float lb2 = approx - slack;
					lb = lb2 > 0.f ? sqrtf(lb2) * (1.0f - 1e-6f) : 0.f;
				}
				// not (lb > worst): NaN (zero vectors under cosine) and ties stay candidates
				maybe = (n < k) || !(lb > o2f(td[k - 1]));
			}
		}
		uint32_t mm = __ballot_sync(0xffffffffu, maybe);
		if (mm == 0u) continue;
		// ---- exact re-scoring of the candidates, G at a time, by groups of TPR lanes ---------------------------------
		const uint32_t my_rank = __popc(mm & lt);  // rank of this lane's row among the candidates of this step
		float		   dex = 0.f;
		uint32_t	   done = 0;
		n_resc += (uint32_t) __popc(mm);
		while (mm)
		{
			uint32_t t = mm;
			for (int i = 0; i < grp; i++) t &= t - 1u;	// group g takes the g-th remaining candidate
			const bool	   have = t != 0u;
			const uint32_t pj = base + (have ? (uint32_t) __ffs(t) - 1u : (uint32_t) __ffs(mm) - 1u);  // idle groups shadow the first one
			const float	  *vp = vectors + (size_t) (r0 + pj) * row_f;
			const float	   d = distance_exact<METRIC, TPR>(qp, vp, (int) dim, qn, vnorm2[pj], sub);
			// deliver: the lane whose row has rank r among the candidates reads group (r - done)'s result
			const int	src = ((int) my_rank - (int) done) * TPR;
			const float got = __shfl_sync(0xffffffffu, d, (src >= 0 && src < 32) ? src : 0);
			if (maybe && my_rank >= done && my_rank < done + (uint32_t) G) dex = got;
			done += (uint32_t) G;
			for (int i = 0; i < G && mm; i++) mm &= mm - 1u;
		}
		bool	 take = false;
		uint32_t d = 0;
		if (maybe)
		{
			// tripwire: the exact value must lie within the assumed error bound of the approximation
			const float ex = (METRIC == M_COS) ? dex : dex * dex;
			if (fabsf(ex - approx) > slack * 1.5f + 1e-6f * fabsf(ex)) *viol = 1;
			d = f2o(dex);
			take = (n < k) || less(d, l, td[k - 1], tl[k - 1]);
		}
		const uint32_t m = __ballot_sync(0xffffffffu, take);
		if (m)
		{
			if (nc + (uint32_t) __popc(m) > kScanCand) merge();
			if (take)
			{
				const uint32_t at = nc + __popc(m & lt);
				cd[w][at] = d;
				cl[w][at] = l;
			}
			nc += __popc(m);
			// the first k candidates establish the threshold: merge them at once so that the filter starts to discard
			if (n < k && nc >= k) merge();
		}
	}
	if (nc) merge();
	if (lane == 0)
	{
		top_n[q] = n;
		if (rescored) atomicAdd(rescored, n_resc);
	}
}

}  // namespace pgemb
