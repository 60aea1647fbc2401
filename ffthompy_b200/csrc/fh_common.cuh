// fh_common.cuh — shared host/device helpers for libffthom_b200 (sm_100a only).
//
// All arithmetic on the Fourier–Galerkin path is fp64 / complex128 (reference:
// ffthompy/tensors/objects.py:119-121).  Complex numbers travel as interleaved
// (re, im) doubles, i.e. numpy complex128 layout == double2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>

typedef double2 cplx;

// ---------------------------------------------------------------- error state
#define FH_OK 0
#define FH_ERR_CUDA -1
#define FH_ERR_ARG -2
#define FH_ERR_UNSUPPORTED -3
#define FH_ERR_ALLOC -4

int fh_set_error(int code, const char* fmt, ...);
cudaStream_t fh_stream();
int fh_num_sms();
int fh_max_smem_optin();
void fh_count_launch();

#define FH_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t _e = (call);                                                   \
        if (_e != cudaSuccess)                                                     \
            return fh_set_error(FH_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, \
                                #call, cudaGetErrorString(_e));                    \
    } while (0)

#define FH_LAUNCH_CHECK()                                                               \
    do {                                                                                \
        fh_count_launch();                                                              \
        cudaError_t _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess)                                                          \
            return fh_set_error(FH_ERR_CUDA, "%s:%d: kernel launch -> %s", __FILE__,    \
                                __LINE__, cudaGetErrorString(_e));                      \
    } while (0)

#define FH_REQUIRE(cond, ...)                                   \
    do {                                                        \
        if (!(cond)) return fh_set_error(FH_ERR_ARG, __VA_ARGS__); \
    } while (0)

// ---------------------------------------------------------------- complex math
__host__ __device__ __forceinline__ cplx cmake(double r, double i) { return make_double2(r, i); }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }

// ---------------------------------------------------------------- grid helpers
// Signed integer frequency of storage index i on an axis of length n in FFT
// order (reference: trigpol.py:17-23 — arange(fix(-n/2), fix(n/2+0.5)) after
// ifftshift; for even n the Nyquist bin n/2 carries k = -n/2).
__host__ __device__ __forceinline__ int fh_freq(int i, int n) { return (i < (n + 1) / 2) ? i : i - n; }
// Same for centred ('c') storage: index i -> k = i - fix(n/2).
__host__ __device__ __forceinline__ int fh_freq_c(int i, int n) { return i - n / 2; }

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum; result valid in thread 0.  `red` is >= 32 doubles of smem.
// Fixed reduction tree => deterministic for a fixed launch shape.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect `red` against a previous use
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = (lane < nw) ? red[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}
__device__ __forceinline__ double block_max(double v, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = (lane < nw) ? red[lane] : 0.0;
        r = warp_max(r);
    }
    return r;
}

static inline int64_t fh_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
