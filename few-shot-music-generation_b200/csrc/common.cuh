// common.cuh — shared helpers for libfsmg (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace fsmg {

// ---- error plumbing (nothing throws across the C ABI) -------------------------------------
std::string& last_error();
int set_error(int code, const char* fmt, ...);

#define FSMG_CUDA_OK(expr)                                                                   \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return ::fsmg::set_error(-2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                     __FILE__, __LINE__);                                    \
    } while (0)

#define FSMG_LAUNCH_OK()                                                                     \
    do {                                                                                     \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess)                                                               \
            return ::fsmg::set_error(-2, "kernel launch failed: %s (%s:%d)",                 \
                                     cudaGetErrorString(_e), __FILE__, __LINE__);            \
    } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- device helpers --------------------------------------------------------------------------
// libdevice-accurate transcendentals (<= 2 ulp): parity with the fp32 reference comes first;
// MUFU.TANH / ex2.approx shortcuts would cost ~1e-3 relative on small pre-activations.
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return tanhf(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum; `red` must hold >= 32 floats of shared memory; result valid in every thread
__device__ __forceinline__ float block_sum(float v, float* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0f;
    if (w == 0) r = warp_sum(r);
    if (threadIdx.x == 0) red[0] = r;
    __syncthreads();
    return red[0];
}
__device__ __forceinline__ float block_max(float v, float* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    float r = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
    if (w == 0) r = warp_max(r);
    if (threadIdx.x == 0) red[0] = r;
    __syncthreads();
    return red[0];
}

}  // namespace fsmg
