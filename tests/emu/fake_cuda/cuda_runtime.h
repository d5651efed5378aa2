// Host stand-in for <cuda_runtime.h>: lets g++ compile the .cuh kernels that use no warp-level primitives
// (barriers + shared memory + plain arithmetic only) and run them with one host thread per CUDA thread.
// Test infrastructure (tests/emu); never part of the product build.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <pthread.h>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 { unsigned x = 1, y = 1, z = 1; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

extern thread_local dim3 threadIdx, blockIdx;
extern dim3				 blockDim, gridDim;
extern pthread_barrier_t g_cta_barrier;
static inline void		 __syncthreads() { pthread_barrier_wait(&g_cta_barrier); }

// correctly rounded single operations; the TU is compiled with -ffp-contract=off and without fast-math
static inline float	 __fadd_rn(float a, float b) { return a + b; }
static inline float	 __fsub_rn(float a, float b) { return a - b; }
static inline float	 __fmul_rn(float a, float b) { return a * b; }
static inline float	 __fsqrt_rn(float a) { return sqrtf(a); }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float	 __double2float_rn(double a) { return (float) a; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t) p; }
// warp primitives are not emulated: kernels that need them cannot run here
static inline float __shfl_sync(unsigned, float v, int, int = 32) { __builtin_trap(); return v; }
using std::max;
using std::min;
