// Host stand-in for <cublas_v2.h> (tests/emu): the one cuBLAS call the PGEMB_SCAN_TC prototype makes.
// cublasGemmEx computes the products in double and then PERTURBS every entry by
//     +/- PGEMB_EMU_GEMM_ERR_PPM * 1e-6 * sum_k |a_k b_k|        (sign from a hash of the entry's position)
// i.e. it plays an adversarial reduced-precision GEMM whose error is as large as the filter's assumed bound allows --
// the filter must still return exactly the exact scan's results.  With an error larger than the bound the filter's
// tripwire has to fire (tests/test_capi_emulated.py).  Test infrastructure only.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>

typedef struct EmuCublas { int dummy; } *cublasHandle_t;
typedef enum { CUBLAS_STATUS_SUCCESS = 0, CUBLAS_STATUS_INVALID_VALUE = 7 } cublasStatus_t;
typedef enum { CUBLAS_OP_N = 0, CUBLAS_OP_T = 1 } cublasOperation_t;
typedef enum { CUDA_R_32F = 0 } cudaDataType;
typedef enum { CUBLAS_COMPUTE_32F = 68, CUBLAS_COMPUTE_32F_FAST_TF32 = 77 } cublasComputeType_t;
typedef enum { CUBLAS_GEMM_DEFAULT = -1 } cublasGemmAlgo_t;

static inline cublasStatus_t cublasCreate_v2(cublasHandle_t *h) { *h = new EmuCublas{0}; return CUBLAS_STATUS_SUCCESS; }
static inline cublasStatus_t cublasDestroy_v2(cublasHandle_t h) { delete h; return CUBLAS_STATUS_SUCCESS; }
static inline cublasStatus_t cublasSetStream_v2(cublasHandle_t, cudaStream_t) { return CUBLAS_STATUS_SUCCESS; }

// column-major C(m x n) = alpha * op(A)(m x k) * op(B)(k x n) + beta * C, fp32 in/out
static inline cublasStatus_t cublasGemmEx(cublasHandle_t, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const void *alpha,
										  const void *A, cudaDataType, int lda, const void *B, cudaDataType, int ldb, const void *beta, void *C,
										  cudaDataType, int ldc, cublasComputeType_t, cublasGemmAlgo_t)
{
	const float *a = (const float *) A, *b = (const float *) B;
	float		*c = (float *) C;
	const float	 al = *(const float *) alpha, be = *(const float *) beta;
	const char	*e = getenv("PGEMB_EMU_GEMM_ERR_PPM");
	const double rel = (e && *e) ? atof(e) * 1e-6 : 0.0;
	for (int j = 0; j < n; j++)
		for (int i = 0; i < m; i++)
		{
			double acc = 0.0, mag = 0.0;
			for (int kk = 0; kk < k; kk++)
			{
				const double av = (ta == CUBLAS_OP_N) ? a[(size_t) i + (size_t) kk * lda] : a[(size_t) kk + (size_t) i * lda];
				const double bv = (tb == CUBLAS_OP_N) ? b[(size_t) kk + (size_t) j * ldb] : b[(size_t) j + (size_t) kk * ldb];
				acc += av * bv;
				mag += fabs(av * bv);
			}
			uint32_t h = (uint32_t) i * 2654435761u ^ (uint32_t) j * 40503u;
			h ^= h >> 15;
			const double sign = (h & 1u) ? 1.0 : -1.0;
			const double v = acc + sign * rel * mag;
			c[(size_t) i + (size_t) j * ldc] = (float) (al * v + (be != 0.f ? be * c[(size_t) i + (size_t) j * ldc] : 0.0));
		}
	return CUBLAS_STATUS_SUCCESS;
}
