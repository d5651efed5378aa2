// dist_exact.cuh -- the three pg_embedding distance functions with the reference binary's exact
// fp32 summation order, written for a GROUP of TPR cooperating threads per (query,row) pair.
//
// Reference: distfunc.c:28-65 (l2, AVX2 path), :133-145 (cosine), :147-155 (manhattan), built with
// -Ofast (reference Makefile:14).  Because of -Ofast the executed order is the vectoriser's, not the
// source's; it is documented in DESIGN.md section 4 / oracle/hnsw_oracle.c and reproduced here with
// explicitly rounded __fmul_rn/__fadd_rn/__fsub_rn (never contracted to FMA), so every distance is
// bit-identical to the reference's and traversal decisions cannot diverge.
//
//   L2:        8 accumulator lanes; per 16-float block  S[j] += (d[j]^2 + d[8+j]^2);
//              t[j]=S[j]+S[j+4]; res=(t0+t2)+(t1+t3); gcc's vectorised tail for dim%16; sqrtf.
//   cosine:    4 lanes for dot (and for the two squared norms); hsum4 = (s0+s2)+(s1+s3);
//              scalar tail for dim%4; 1 - dot/sqrt(nb*na) with the product in fp32 and
//              sqrt / divide / subtract in fp64.
//   manhattan: 4 lanes of |a-b|; hsum4; scalar tail.
//
// Thread mapping: a row is scored by TPR consecutive lanes of a warp (TPR in {1,2,4,8}, TPR <= LANES);
// lane `sub` owns accumulator lanes [sub*LPT, (sub+1)*LPT), LPT = LANES/TPR, and reads them with one
// LPT-wide vector load per block.  The per-lane chains are sequential in the dimension index, exactly
// like the SIMD lanes of the reference; the horizontal sum is done after exchanging the lane sums with
// width-TPR shuffles.  All TPR lanes return the same value.
#pragma once
#include "common.cuh"

namespace pgemb {

enum : int { M_L2 = 0, M_COS = 1, M_MAN = 2 };

template <int METRIC> struct MetricLanes { static constexpr int LANES = (METRIC == M_L2) ? 8 : 4; };

template <int N> struct VecLd;
template <> struct VecLd<1> {
	float v[1];
	__device__ __forceinline__ static VecLd ld(const float *p) { VecLd r; r.v[0] = *p; return r; }
};
template <> struct VecLd<2> {
	float v[2];
	__device__ __forceinline__ static VecLd ld(const float *p)
	{
		VecLd r; float2 t = *reinterpret_cast<const float2 *>(p); r.v[0] = t.x; r.v[1] = t.y; return r;
	}
};
template <> struct VecLd<4> {
	float v[4];
	__device__ __forceinline__ static VecLd ld(const float *p)
	{
		VecLd r; float4 t = *reinterpret_cast<const float4 *>(p);
		r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
	}
};

__device__ __forceinline__ float hsum4(float s0, float s1, float s2, float s3)
{
	return __fadd_rn(__fadd_rn(s0, s2), __fadd_rn(s1, s3));
}

// ---- 4-lane metrics: generic "sum of term(q[i], v[i])" -------------------------------------
// TERM: 0 = q*v (dot), 1 = |q-v| (manhattan), 2 = v*v (squared norm of v; q unused)
template <int TERM>
__device__ __forceinline__ float term4(float qv, float vv)
{
	if (TERM == 0) return __fmul_rn(qv, vv);
	if (TERM == 1) return fabsf(__fsub_rn(qv, vv));
	return __fmul_rn(vv, vv);
}

// Returns hsum + scalar tail, identical in all TPR lanes of the group.
// q, v must be aligned to LPT*4 bytes (LPT = 4/TPR); `sub` = lane index inside the group.
template <int TERM, int TPR>
__device__ __forceinline__ float sum4_exact(const float *__restrict__ q, const float *__restrict__ v, int dim, int sub)
{
	static_assert(TPR == 1 || TPR == 2 || TPR == 4, "4-lane metrics: TPR in {1,2,4}");
	constexpr int LPT = 4 / TPR;
	float		  s[LPT];
#pragma unroll
	for (int j = 0; j < LPT; j++) s[j] = 0.0f;
	const int	 main_n = dim & ~3;
	const float *qp = q + sub * LPT;
	const float *vp = v + sub * LPT;
#pragma unroll 8
	for (int i = 0; i < main_n; i += 4)
	{
		VecLd<LPT> a = VecLd<LPT>::ld(qp + i);
		VecLd<LPT> b = VecLd<LPT>::ld(vp + i);
#pragma unroll
		for (int j = 0; j < LPT; j++) s[j] = __fadd_rn(s[j], term4<TERM>(a.v[j], b.v[j]));
	}
	float full[4];
#pragma unroll
	for (int j = 0; j < 4; j++) full[j] = __shfl_sync(kFull, s[j % LPT], j / LPT, TPR);
	float res = hsum4(full[0], full[1], full[2], full[3]);
	for (int k = main_n; k < dim; k++) res = __fadd_rn(res, term4<TERM>(q[k], v[k]));
	return res;
}

// distfunc.c:144: 1 - (distance / sqrt(norma * normb)), product in fp32, the rest in fp64.
__device__ __forceinline__ float cosine_finish(float dot, float na, float nb)
{
	float  prod = __fmul_rn(nb, na);
	double r = 1.0 - __ddiv_rn((double) dot, __dsqrt_rn((double) prod));
	return __double2float_rn(r);
}

// gcc's vectorised epilogue of the `while (x < pEnd2)` loop (distfunc.c:58-62) for r = dim%16 leftover
// floats starting at x/y; see oracle/hnsw_oracle.c:l2_dist_avx2_order for the derivation.
__device__ __forceinline__ float l2_tail_exact(float res, const float *__restrict__ x, const float *__restrict__ y, int r)
{
	if (r == 0) return res;
	int	  pos = 0;
	float u0 = 0.f, u1 = 0.f, u2 = 0.f, u3 = 0.f;
	bool  have8 = r >= 8;
	if (have8)
	{
		float E[8];
#pragma unroll
		for (int k = 0; k < 8; k++)
		{
			float d = __fsub_rn(x[pos + k], y[pos + k]);
			E[k] = __fmul_rn(d, d);
		}
		u0 = __fadd_rn(E[0], E[4]); u1 = __fadd_rn(E[1], E[5]);
		u2 = __fadd_rn(E[2], E[6]); u3 = __fadd_rn(E[3], E[7]);
		pos += 8;
		r -= 8;
	}
	if (r >= 4)
	{
		float w[4];
#pragma unroll
		for (int j = 0; j < 4; j++)
		{
			float d = __fsub_rn(x[pos + j], y[pos + j]);
			w[j] = __fmul_rn(d, d);
		}
		w[0] = __fadd_rn(w[0], u0); w[1] = __fadd_rn(w[1], u1);
		w[2] = __fadd_rn(w[2], u2); w[3] = __fadd_rn(w[3], u3);
		res = __fadd_rn(res, hsum4(w[0], w[1], w[2], w[3]));
		pos += 4;
		r -= 4;
	}
	else if (have8)
		res = __fadd_rn(hsum4(u0, u1, u2, u3), res);
	for (int k = 0; k < r; k++)
	{
		float d = __fsub_rn(x[pos + k], y[pos + k]);
		res = __fadd_rn(res, __fmul_rn(d, d));
	}
	return res;
}

// ---- L2 (8 lanes, 16-float blocks) ----------------------------------------------------------
template <int TPR>
__device__ __forceinline__ float l2_exact(const float *__restrict__ x, const float *__restrict__ y, int dim, int sub)
{
	static_assert(TPR == 1 || TPR == 2 || TPR == 4 || TPR == 8, "L2: TPR in {1,2,4,8}");
	constexpr int LPT = 8 / TPR;				   // accumulator lanes per thread
	constexpr int VW = (LPT >= 4) ? 4 : LPT;	   // vector width of one load
	constexpr int NV = LPT / VW;				   // loads per half block
	float		  S[LPT];
#pragma unroll
	for (int j = 0; j < LPT; j++) S[j] = 0.0f;
	const int	 main_n = dim & ~15;
	const float *xp = x + sub * LPT;
	const float *yp = y + sub * LPT;
#pragma unroll 4
	for (int i = 0; i < main_n; i += 16)
	{
#pragma unroll
		for (int h = 0; h < NV; h++)
		{
			VecLd<VW> x0 = VecLd<VW>::ld(xp + i + h * VW);
			VecLd<VW> y0 = VecLd<VW>::ld(yp + i + h * VW);
			VecLd<VW> x1 = VecLd<VW>::ld(xp + i + 8 + h * VW);
			VecLd<VW> y1 = VecLd<VW>::ld(yp + i + 8 + h * VW);
#pragma unroll
			for (int j = 0; j < VW; j++)
			{
				float d0 = __fsub_rn(x0.v[j], y0.v[j]);
				float d1 = __fsub_rn(x1.v[j], y1.v[j]);
				float pp = __fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1));
				S[h * VW + j] = __fadd_rn(S[h * VW + j], pp);
			}
		}
	}
	float full[8];
#pragma unroll
	for (int j = 0; j < 8; j++) full[j] = __shfl_sync(kFull, S[j % LPT], j / LPT, TPR);
	float res = hsum4(__fadd_rn(full[0], full[4]), __fadd_rn(full[1], full[5]),
					  __fadd_rn(full[2], full[6]), __fadd_rn(full[3], full[7]));
	res = l2_tail_exact(res, x + main_n, y + main_n, dim - main_n);
	return __fsqrt_rn(res);
}

// length of the vectorised main loop of the reference: dim rounded down to its SIMD block (16 floats for the AVX2 L2 loop, 4 otherwise)
template <int METRIC>
__device__ __forceinline__ int main_len(int dim) { return (METRIC == M_L2) ? (dim & ~15) : (dim & ~3); }

// ---- one entry point ------------------------------------------------------------------------
// q = the query / new point (reference argument `ax`), v = the stored node (`bx`).
// qn, vn: cached squared norms (cosine only; same lane order -> same bits as distfunc.c:141-142).
template <int METRIC, int TPR>
__device__ __forceinline__ float distance_exact(const float *__restrict__ q, const float *__restrict__ v, int dim,
												float qn, float vn, int sub)
{
	if (METRIC == M_L2)
		return l2_exact<TPR>(q, v, dim, sub);
	else if (METRIC == M_COS)
		return cosine_finish(sum4_exact<0, (TPR > 4 ? 4 : TPR)>(q, v, dim, sub), qn, vn);
	else
		return sum4_exact<1, (TPR > 4 ? 4 : TPR)>(q, v, dim, sub);
}

// Squared norm in the cosine lane order (what pgemb_index_append caches per node).
template <int TPR>
__device__ __forceinline__ float sqnorm_exact(const float *__restrict__ v, int dim, int sub)
{
	return sum4_exact<2, TPR>(v, v, dim, sub);
}

}  // namespace pgemb
