// probe.cuh — hardware probe (test infrastructure): where does tcgen05.mma.cta_group::2 put the rows of D in each CTA's TMEM for a
// given (M, N)?  One cluster of two CTAs computes D[r, n] = 256 r + n exactly (A[r, 0] = r, A[r, 1] = 1, B[n, 0] = 256, B[n, 1] = n,
// fp16 operands in the canonical SWIZZLE_128B K-major layout, K = 64) and every CTA dumps its whole 128-lane x N-column window.
#pragma once
#include "tc_gemm.cuh"
#include "tc_lstm.cuh"

namespace fsmg {
namespace tc {

__device__ __forceinline__ uint32_t sw128_offset(int r, int k) {   // byte offset of element (row r, k) in a [rows x 64] fp16 tile
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2);
}

__global__ void __launch_bounds__(192, 1) mma_probe_2sm_kernel(int M, int N, float* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                 // [128 rows x 64] (M/2 used)
    uint8_t* sB = smem + 16384;         // [128 rows x 64] (N/2 used)
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 32768 + 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int prank = (int)cluster_ctarank();
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    __syncthreads();
    const int mh = M / 2, nh = N / 2;
    for (int r = threadIdx.x; r < mh; r += blockDim.x) {
        *reinterpret_cast<__half*>(sA + sw128_offset(r, 0)) = __float2half_rn((float)(prank * mh + r));
        *reinterpret_cast<__half*>(sA + sw128_offset(r, 1)) = __float2half_rn(1.0f);
    }
    for (int n = threadIdx.x; n < nh; n += blockDim.x) {
        *reinterpret_cast<__half*>(sB + sw128_offset(n, 0)) = __float2half_rn(256.0f);
        *reinterpret_cast<__half*>(sB + sw128_offset(n, 1)) = __float2half_rn((float)(prank * nh + n));
    }
    if (warp == 0 && lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc_2sm(slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *slot;
    // poison the accumulator window so untouched cells are recognisable
    if (warp >= 2) {
        uint32_t w[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) w[e] = __float_as_uint(-1.0f);
        for (int c0 = 0; c0 < N; c0 += 32) tmem_st32(tmem_base + c0 + ((uint32_t)((warp & 3) * 32) << 16), w);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 1 && lane == 0 && prank == 0) {
        const uint32_t idesc = make_idesc_m(M, N);
        const uint64_t a0 = make_smem_desc(smem_u32(sA), 16, 1024), b0 = make_smem_desc(smem_u32(sB), 16, 1024);
        for (int k = 0; k < 4; ++k) umma_f16_2sm(tmem_base, a0 + 2 * k, b0 + 2 * k, idesc, k > 0 ? 1u : 0u);
        umma_commit_2sm_mc(bar, (uint16_t)0x3);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    if (warp >= 2) {
        const int quad = warp & 3;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + c0 + ((uint32_t)(quad * 32) << 16), r);
            tmem_ld_wait();
            for (int e = 0; e < 32 && c0 + e < N; ++e)
                out[((int64_t)prank * 128 + quad * 32 + lane) * N + c0 + e] = __uint_as_float(r[e]);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

}  // namespace tc

static inline int mma_probe_2sm(int M, int N, float* d_out, cudaStream_t s) {
    if ((M != 128 && M != 256) || N < 16 || N > 256 || (N % 32) != 0) return set_error(-1, "mma probe: M in {128, 256}, N multiple of 32 up to 256");
    const int smem = 34 * 1024 + 1024;
    FSMG_CUDA_OK(cudaFuncSetAttribute(tc::mma_probe_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(2);
    cfg.blockDim = dim3(192);        // warp 0: barrier init, warp 1: TMEM alloc + MMA, warps 2-5: one per TMEM lane quadrant
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FSMG_CUDA_OK(cudaLaunchKernelEx(&cfg, tc::mma_probe_2sm_kernel, M, N, d_out));
    return 0;
}

}  // namespace fsmg
