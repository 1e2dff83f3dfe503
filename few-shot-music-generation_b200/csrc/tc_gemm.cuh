// tc_gemm.cuh — the tcgen05 / TMA / TMEM route of libfsmg (sm_100a only).
//
// One persistent, warp-specialised GEMM core (fp16 x fp16 -> fp32 in TMEM) with pluggable epilogues:
//   warp 0      : TMA producer  (cp.async.bulk.tensor 2D, SWIZZLE_128B, mbarrier complete_tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (default CL = 2: the leader CTA of a cluster of two issues ONE
//                 tcgen05.mma.cta_group::2 of 256 x BN x 16 for the pair; CL = 1: UMMA 128 x BN x 16, cta_group::1)
//   warps 2..17 : sixteen epilogue warps (TMEM lane quadrant x column quarter, tcgen05.ld 32x32b in 16-column chunks),
//                 double-buffered TMEM accumulators (single-buffered for the 256 x 512 pair tiles); in the XF instantiations the
//                 same warps first transform the A operand of every k block in shared memory (fused softmax gradient, see below)
// C[M,N] (op)= alpha * A * B^T with A, B each either K-major (row = m/n, contiguous k) or MN-major
// (row = k, contiguous m/n) — the latter serves the weight-gradient contractions over tokens.
//
// Contractions of the hot path that run here: K3 input GEMM, K4 per-step recurrent GEMM
// (until the persistent kernel takes over), K5+K7 projection with fused online log-sum-exp / NLL,
// K8 dgrad / wgrad GEMMs — for the projection's two, K7's softmax - onehot is rebuilt on their A operand (XF).
#pragma once
#include <type_traits>
#include <cuda.h>

#include "common.cuh"
#include "simt_kernels.cuh"

namespace fsmg {

namespace tc {

constexpr int BM = 128;      // UMMA M (cta_group::1)
constexpr int BK = 64;       // 64 fp16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 16;                // four epilogue warpgroups (4 warps per SM sub-partition): each owns a quarter of the tile's columns
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int EPI_STAGE_BYTES = 2048;            // per epilogue warp: 32 rows x 16 fp32 staging (swizzled) / 1 KB fp16 staging + bias

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the thread is parked by the hardware until the phase flips (or the hint expires) instead of
// spinning — the single-lane producer / MMA waiters were costing ~17 % of all issued instructions on the epilogue warps' SMSPs
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(addr), "r"(parity), "r"(0x989680u) : "memory");
}
// bare MUFU.EX2 (exp2f() without -use_fast_math wraps it in two FMULs for denormal inputs, which the epilogues never produce)
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// multicast variant: the box lands at the same smem offset of every CTA in `mask`, and each destination CTA's
// mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1)
        : "memory");
}
// cta_group::2 flavour: data lands in THIS CTA's smem, the complete_tx goes to the mbarrier at the same offset in CTA 0 of the pair
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar_local) {
    asm volatile(
        "{\n"
        ".reg .b32 rb;\n"
        "mapa.shared::cluster.u32 rb, %2, 0;\n"
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [rb];\n"
        "}\n" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar_local)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// same, arriving on the barrier at this offset in every CTA of `mask` (a smem stage fed by multicast is free only
// when all CTAs that received the data have consumed it)
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// ---- cta_group::2 (two SMs cooperate on one 256-row MMA; issued by the leader CTA of the pair only) ---------------------
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `cta` of the cluster.  Default semantics (.release.cta): the
// .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR, which makes the arriving warp drain all of its outstanding global
// stores first (measured: 17 % of the 2-SM LSE kernel's stall samples); what the waiter consumes here is TMEM / tensor-core
// state, ordered by tcgen05.fence::before_thread_sync / after_thread_sync around the barrier, not generic memory.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}
// cluster-scope release flavour for one-shot hand-offs of generic/async-proxy DATA to the peer (cost irrelevant there)
__device__ __forceinline__ void mbar_arrive_remote_release(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <-> TMEM lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_dep16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait for outstanding tcgen05.ld and tie the destination registers to the wait, so that no use of them can be
// scheduled ahead of it (needed when loads are software-pipelined across chunks)
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// ---- descriptors ------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (sm_100 "SmemDescriptor": cute/arch/mma_sm100_desc.hpp):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (kind::f16): c_format F32 (bits 4-5 = 1), a/b format F16 (0), a_major bit 15,
// b_major bit 16 (0 = K-major, 1 = MN-major), N>>3 at bits 17-22, M>>4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc_m(int m, int n) {   // K-major operands, explicit M (256 for cta_group::2)
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_full(int m, int n, bool a_mn, bool b_mn) {
    return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
    return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ---- epilogue parameter block ----------------------------------------------------------------------
enum EpiMode { EPI_STORE = 0, EPI_LSE = 1, EPI_SCATTER = 2 };

struct EpiParams {
    // EPI_STORE
    void* C; int64_t ldc;
    const float* bias; float alpha;
    int c_half, accumulate, atomic;
    int vec_ok;                         // host-checked: C base and ldc allow 16-byte (fp32) / 8-byte (fp16) row-chunk accesses
    // EPI_SCATTER (embedding gradient): row r of the result is atomically added to C[y[row0 + r], :] (C = dense gradient table,
    // ldc = its row length) and the squared Frobenius norm of the scattered rows is accumulated into *tgt (TF clip quirk, A.6)
    // EPI_LSE: logits = acc + bias; per (row, n-tile) online (max, sumexp); target-logit pick; optional fp16 logits
    float2* part; int n_tiles_total;    // part[row * n_tiles_total + n_blk]
    const int32_t* y; int64_t row0;     // y[row0 + row]
    float* tgt;                         // tgt[row]
    __half* logits16; int64_t ld16;
    // fused softmax gradient (see "XF" below).  EPI_LSE with exp_store: the fp16 chunk receives e = exp(logit - cmax) of every
    // 16-column chunk instead of the logit, and cmax goes to cmaxT[(col / 16) * ld_cmax + row].  EPI_STORE with XF: the consumer
    // GEMMs rebuild dlogits = e * exp(cmax - lse) - onehot(y) in shared memory (lse / y indexed by row0 + chunk row).
    int exp_store;
    float* cmaxT; int64_t ld_cmax; int n_c16;
    const float* xf_lse;
    float* xf_db;                       // XF = 2: db[col] += alpha * column sums of dlogits (pieces of N tile 0 only)
    int xf_diag;                        // timing diagnostics only (FSMG_XF_DIAG, wrong results): bit 0 no column sums, bit 1 no transform math
};

struct GemmShape {
    int M, N, K;
    int n_m, n_n, n_s;      // tiles along M, N and K-splits
    int streamk;            // 1: each cluster owns a contiguous range of (cluster-tile, k-block) units (balanced; RED epilogue)
    int units_per_cluster;  // stream-K: k-block units per cluster
    int n_mp;               // M tiles per cluster-tile row = ceil(n_m / CL): a cluster of CL CTAs owns CL consecutive M tiles of one N tile
    int kb_per_split;       // k-blocks per split
    int kb_total;
    int astat_s;            // A-stationary schedule: N ranges per M pair-block (items = n_mp * astat_s)
};

// Work iterator shared by the three warp roles.  mode 0: cluster-tiles (optionally K-split) strided over the clusters;
// mode 1 (stream-K): a contiguous range of k-block units, cut into (tile, kb0, kb1) pieces at tile boundaries.
struct WorkIter {
    int cur, end, stride;
    __device__ __forceinline__ WorkIter(const GemmShape& sh, int cluster_id, int n_clusters) {
        if (sh.streamk) {
            const int total = sh.n_mp * sh.n_n * sh.kb_total;
            cur = min(total, cluster_id * sh.units_per_cluster);
            end = min(total, cur + sh.units_per_cluster);
            stride = 0;
        } else {
            cur = cluster_id;
            end = sh.n_mp * sh.n_n * sh.n_s;
            stride = n_clusters;
        }
    }
    __device__ __forceinline__ bool next(const GemmShape& sh, int& tile_mn, int& kb0, int& kb1) {
        if (cur >= end) return false;
        if (sh.streamk) {
            tile_mn = cur / sh.kb_total;
            kb0 = cur - tile_mn * sh.kb_total;
            kb1 = min(sh.kb_total, kb0 + (end - cur));
            cur += kb1 - kb0;
        } else {
            const int tiles_mn = sh.n_mp * sh.n_n;
            tile_mn = cur % tiles_mn;
            const int s_blk = cur / tiles_mn;
            kb0 = s_blk * sh.kb_per_split;
            kb1 = min(sh.kb_total, kb0 + sh.kb_per_split);
            cur += stride;
        }
        return true;
    }
};

// A-stationary schedule (logits GEMM, K <= 512): an item is (M pair-block, contiguous range of N tiles); the pair keeps its 256 x K
// block of A resident in smem for the whole item and streams only B.  Items are strided over the pairs.
struct AstatIter {
    int item, total, stride, m_pb, n_cur, n_hi;
    bool first, last;      // first / last N tile of the current item
    __device__ __forceinline__ AstatIter(const GemmShape& sh, int cluster_id, int n_clusters) {
        stride = n_clusters;
        item = cluster_id - n_clusters;
        total = sh.n_mp * sh.astat_s;
        m_pb = 0; n_cur = 0; n_hi = 0; first = false; last = false;
    }
    __device__ __forceinline__ bool next(const GemmShape& sh, int& tile_mn, int& kb0, int& kb1) {
        if (n_cur >= n_hi) {
            item += stride;
            if (item >= total) return false;
            m_pb = item / sh.astat_s;
            const int sidx = item - m_pb * sh.astat_s;
            n_cur = (sidx * sh.n_n) / sh.astat_s;
            n_hi = ((sidx + 1) * sh.n_n) / sh.astat_s;
            first = true;
        } else {
            first = false;
        }
        last = (n_cur + 1 == n_hi);
        tile_mn = n_cur * sh.n_mp + m_pb;
        kb0 = 0; kb1 = sh.kb_total;
        ++n_cur;
        return true;
    }
};
struct WorkIterFlags : WorkIter {   // same interface for the streaming schedule (flags unused)
    bool first = true, last = true;
    __device__ __forceinline__ WorkIterFlags(const GemmShape& sh, int cluster_id, int n_clusters) : WorkIter(sh, cluster_id, n_clusters) {}
};

constexpr int ASTAT_KB = 8;   // resident A panels (64 k each): K <= 512

template <int BN, int CL, bool ASTAT = false>
struct SmemLayout {
    static constexpr int A_BYTES = BM * BK * 2;              // 16 KB
    static constexpr int B_BYTES = (BN / CL) * BK * 2;       // CL = 2 (cta_group::2): each CTA of the pair holds half of the B tile
    static constexpr int RES_BYTES = ASTAT ? ASTAT_KB * A_BYTES : 0;       // A-stationary: this CTA's 128 x 512 block of A stays resident
    static constexpr int STAGE_BYTES = ASTAT ? B_BYTES : A_BYTES + B_BYTES;    // 48 / 32 / 24 KB (16 KB when only B streams)
    static constexpr int RING = 192 * 1024 - RES_BYTES;
    static constexpr int STAGES = RING / STAGE_BYTES > 8 ? 8 : RING / STAGE_BYTES;
    static constexpr int BAR_BYTES = 512;
    static constexpr int EPI_BYTES = NUM_EPI_WARPS * EPI_STAGE_BYTES;
    static constexpr int TOTAL = RES_BYTES + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;  // + alignment slack
};

// CL = 2: the two CTAs of a cluster issue ONE tcgen05.mma.cta_group::2 of 256 x BN x 16 per K step: each CTA streams its own 128
// rows of A and only HALF of the B tile, the tensor cores read both CTAs' shared memory.  The K-streaming mainloop is bound by the
// bytes that must LAND in each SM's smem (measured ~40 B/clk/SM; a 1-SM 128 x 256 tile needs 96 B/clk at full tensor rate) —
// TMA multicast (the previous CL = 2 scheme) cut L2 reads but not that ingest; the 2-SM MMA cuts it by a third.
//
// XF (fused softmax gradient; cta_group::2 only).  The A operand of the two projection-backward GEMMs is dlogits = softmax - onehot.
// Instead of a separate HBM pass that rewrites the fp16 logits chunk in place, the logits GEMM stores e = exp(logit - cmax) per
// 16-column chunk (its epilogue computes those exponentials anyway) and the consumers finish the job on the tile that TMA has just
// landed: A loads complete on a CTA-LOCAL barrier (raw_full), the sixteen epilogue warps — idle during the K loop of the
// single-buffered 256 x 512 tiles — multiply every 32-byte piece by its row's exp(cmax - lse) (one MUFU per 16 elements), subtract
// the one-hot target, fence the generic writes for the async proxy and arrive on the leader's full barrier next to the TMA bytes of
// B.  XF = 1: A K-major (dH = dlogits * Ws^T: smem row = token, 128-byte row = one 64-column k block).  XF = 2: A MN-major
// (dWs^T = dlogits^T * hs: smem row = token of the k block, two 64-column boxes); each thread owns fixed vocabulary columns there,
// so the bias gradient (column sums of dlogits) accumulates in registers and leaves with a few REDs per work item.
template <int BN, int EPI, bool A_MN, bool B_MN, int CL, bool ASTAT = false, int XF = 0>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmShape sh, EpiParams ep) {
    using L = SmemLayout<BN, CL, ASTAT>;
    using Iter = std::conditional_t<ASTAT, AstatIter, WorkIterFlags>;
    static_assert(!ASTAT || (CL == 2 && !A_MN && !B_MN && BN <= 256), "A-stationary schedule: K-major operands, cta_group::2");
    static_assert(XF == 0 || (CL == 2 && !ASTAT && EPI == EPI_STORE && (XF == 1 ? !A_MN : A_MN)), "operand transform: pair tiles, plain-store epilogue");
    constexpr int STAGES = L::STAGES;
    // BN = 512 (cta_group::2 only): the pair's tile is 256 x 512, issued as two N = 256 MMAs per K step.  A is fetched once for the
    // whole 512-wide N extent (the L2 -> SM operand traffic, not the tensor pipe, bounds the large-K GEMMs) and the accumulator
    // fills all 512 TMEM columns, so it is single-buffered: only used when the K loop dwarfs the epilogue.
    static_assert(BN <= 256 || CL == 2, "BN = 512 needs cta_group::2");
    constexpr int NMMA = BN > 256 ? BN / 256 : 1;   // MMAs per K step
    constexpr int MMA_N = BN / NMMA;                // N of one MMA
    constexpr int ACC = BN > 256 ? 1 : 2;           // accumulator stages in TMEM (ACC * BN = 512 columns)
    constexpr uint32_t B_PART = (MMA_N / CL) * BK * 2;   // bytes of one MMA's B operand in this CTA's stage
    extern __shared__ uint8_t smem_raw[];
    // keep the pointer in the shared address space (pointer arithmetic on the extern array, no integer round trip):
    // otherwise the staging accesses compile to generic LD/ST instead of LDS/STS
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* res_a = smem;                       // ASTAT: ASTAT_KB resident A panels (128 rows x 64 k each)
    uint8_t* ring = smem + L::RES_BYTES;
    uint8_t* epi_smem = ring + STAGES * L::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + L::EPI_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;       // [2]
    uint64_t* a_full = tmem_empty + 2;          // [ASTAT_KB] ASTAT: panel kb of the current item has landed (leader: both CTAs' bytes)
    uint64_t* a_empty = a_full + ASTAT_KB;      // [ASTAT_KB] ASTAT: the item's last MMAs on panel kb have retired
    uint64_t* raw_full = a_empty + ASTAT_KB;    // [STAGES] XF: this CTA's A tile of the stage has landed (CTA-local)
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(raw_full + 8);
    static_assert((3 * 8 + 4 + 2 * ASTAT_KB) * 8 + 8 <= L::BAR_BYTES, "barrier block");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta_rank = (CL == 2) ? (int)cluster_ctarank() : 0;
    const int cluster_id = (CL == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_clusters = (CL == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        // CL = 2 (cta_group::2): the leader's full barrier collects both CTAs' TMA bytes, its tmem_empty both CTAs' epilogue warps;
        // empty / tmem_full are signalled in both CTAs by the leader's multicast tcgen05.commit
        // XF: the leader's full barrier additionally collects one arrival per transform warp of both CTAs
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], XF ? 1 + CL * NUM_EPI_WARPS : 1); mbar_init(&empty_bar[i], 1); }
        if (XF) for (int i = 0; i < STAGES; ++i) mbar_init(&raw_full[i], 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], CL * NUM_EPI_WARPS); }
        if (ASTAT) for (int i = 0; i < ASTAT_KB; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        fence_barrier_init();
    }
    if (warp == 1) { if (CL == 2) tmem_alloc_2sm(tmem_base_slot, ACC * BN); else tmem_alloc(tmem_base_slot, ACC * BN); }   // accumulator stages
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();      // peer barriers are initialised before any multicast / remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    // programmatic dependent launch (decode loop): everything above touched only this CTA's own state; operands, accumulation
    // targets and epilogue inputs are read / written below, after the predecessor grid has completed (no-ops otherwise)
    griddep_wait();
    griddep_launch();

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            uint32_t a_phase = 0;
            Iter it(sh, cluster_id, n_clusters);
            int tile_mn, kb0, kb1;
            while (it.next(sh, tile_mn, kb0, kb1)) {
                const int m_blk = (tile_mn % sh.n_mp) * CL + cta_rank, n_blk = tile_mn / sh.n_mp;
                if (ASTAT && it.first) {
                    // new item: refill the resident A panels.  Panel kb is released by the previous item's last tile as soon as its
                    // MMAs on that panel retire, so the refill overlaps that tile's remaining K steps and its epilogue.
                    for (int kb = 0; kb < sh.kb_total; ++kb) {
                        mbar_wait(&a_empty[kb], a_phase ^ 1);
                        if (cta_rank == 0) mbar_expect_tx(&a_full[kb], 2 * L::A_BYTES);
                        tma_load_2d_2sm(res_a + kb * L::A_BYTES, &map_a, kb * BK, m_blk * BM, &a_full[kb]);
                    }
                    a_phase ^= 1;
                }
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = ring + stage * L::STAGE_BYTES;
                    uint8_t* sb = ASTAT ? sa : sa + L::A_BYTES;
                    if (ASTAT) {   // only this CTA's half of the B tile streams
                        if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
                        tma_load_2d_2sm(sb, &map_b, kb * BK, n_blk * BN + cta_rank * (BN / 2), &full_bar[stage]);
                    } else if (CL == 1) {
                        mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                        if (A_MN) {
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, &map_a, m_blk * BM + j * 64, kb * BK, &full_bar[stage]);
                        } else {
                            tma_load_2d(sa, &map_a, kb * BK, m_blk * BM, &full_bar[stage]);
                        }
                        if (B_MN) {
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &map_b, n_blk * BN + j * 64, kb * BK, &full_bar[stage]);
                        } else {
                            tma_load_2d(sb, &map_b, kb * BK, n_blk * BN, &full_bar[stage]);
                        }
                    } else {
                        // cta_group::2: this CTA streams ITS 128 rows of A and ITS half of the B tile (the tensor core reads both CTAs'
                        // smem); every load completes on the leader's barrier, which therefore expects both CTAs' bytes
                        if (XF) {
                            // A goes through this CTA's transform warps first: its bytes complete on the local raw_full barrier
                            mbar_expect_tx(&raw_full[stage], L::A_BYTES);
                            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (L::STAGE_BYTES - L::A_BYTES));
                            if (A_MN) {
#pragma unroll
                                for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, &map_a, m_blk * BM + j * 64, kb * BK, &raw_full[stage]);
                            } else {
                                tma_load_2d(sa, &map_a, kb * BK, m_blk * BM, &raw_full[stage]);
                            }
                        } else {
                            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
                            if (A_MN) {
#pragma unroll
                                for (int j = 0; j < BM / 64; ++j) tma_load_2d_2sm(sa + j * 8192, &map_a, m_blk * BM + j * 64, kb * BK, &full_bar[stage]);
                            } else {
                                tma_load_2d_2sm(sa, &map_a, kb * BK, m_blk * BM, &full_bar[stage]);
                            }
                        }
                        // MMA j covers N columns [j * MMA_N, (j + 1) * MMA_N) of the tile; this CTA supplies its half of each
#pragma unroll
                        for (int j = 0; j < NMMA; ++j) {
                            const int n0 = n_blk * BN + j * MMA_N + cta_rank * (MMA_N / 2);
                            if (B_MN) {
#pragma unroll
                                for (int jj = 0; jj < MMA_N / 128; ++jj)
                                    tma_load_2d_2sm(sb + j * B_PART + jj * 8192, &map_b, n0 + jj * 64, kb * BK, &full_bar[stage]);
                            } else {
                                tma_load_2d_2sm(sb + j * B_PART, &map_b, kb * BK, n0, &full_bar[stage]);
                            }
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread; with cta_group::2 only the leader CTA of the pair issues, for both) =====
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = (CL == 2) ? make_idesc_full(256, MMA_N, A_MN, B_MN) : make_idesc(BN, A_MN, B_MN);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            uint32_t a_phase = 0;
            Iter it(sh, cluster_id, n_clusters);
            int tile_mn, kb0, kb1;
            while (it.next(sh, tile_mn, kb0, kb1)) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (ASTAT && it.first) mbar_wait(&a_full[kb], a_phase);   // the item's A panel kb is resident (both CTAs)
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = ASTAT ? smem_u32(res_a + kb * L::A_BYTES) : smem_u32(ring + stage * L::STAGE_BYTES);
                    const uint32_t sb = ASTAT ? smem_u32(ring + stage * L::STAGE_BYTES) : sa + L::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // K-major SW128: atom = 8 rows x 128 B, SBO = 1024 B, +32 B per UMMA_K inside the swizzle row.
                        // MN-major SW128: atom = 8 k-rows x 64 mn, SBO = 1024 B (next 8 k), LBO = 8192 B (next 64 mn), +2048 B per UMMA_K.
                        const uint64_t a_desc = A_MN ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
                        const uint64_t b_desc = B_MN ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
                        if (CL == 2) {
#pragma unroll
                            for (int j = 0; j < NMMA; ++j)   // 256 x MMA_N x 16 over both SMs; (B_PART >> 4) = descriptor address units
                                umma_f16_2sm(d_tmem + j * MMA_N, a_desc, b_desc + (uint64_t)(j * (B_PART >> 4)), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        } else {
                            umma_f16(d_tmem, a_desc, b_desc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                    }
                    if (CL == 2) umma_commit_2sm_mc(&empty_bar[stage], (uint16_t)0x3);   // frees the stage in both CTAs
                    else umma_commit(&empty_bar[stage]);        // smem slot reusable once these MMAs retire
                    if (ASTAT && it.last) umma_commit_2sm_mc(&a_empty[kb], (uint16_t)0x3);   // panel kb may be refilled for the next item
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (ASTAT && it.last) a_phase ^= 1;
                if (CL == 2) umma_commit_2sm_mc(&tmem_full[acc], (uint16_t)0x3);   // accumulators complete in both CTAs -> both epilogues
                else umma_commit(&tmem_full[acc]);
                if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps 2..17: TMEM lane quadrant = warp % 4; warpgroup (warp-2)/4 owns a quarter of the tile's columns,
        // processed in 16-column chunks (registers: 576 threads have to fit, and 4 warps per sub-partition hide each other's latencies)
        const int quad = warp & 3;
        const int cg = (warp - 2) >> 2;                          // column group 0..3
        constexpr int CW = 16;
        constexpr int GCOLS = BN / 4;                            // columns per warp
        constexpr int CHUNKS = GCOLS / CW;                       // 4 (BN = 256) or 2 (BN = 128)
        uint8_t* stg = epi_smem + (warp - 2) * EPI_STAGE_BYTES;
        float4* st4 = reinterpret_cast<float4*>(stg);            // fp32 staging: 32 rows x 4 float4, slot c4 ^ ((row >> 1) & 3)
        uint4* sh4 = reinterpret_cast<uint4*>(stg);              // fp16 staging: 32 rows x 2 uint4, slot c2 ^ ((row >> 2) & 1)
        float* bias_s = reinterpret_cast<float*>(stg + 1024);    // EPI_LSE: bias of this warp's GCOLS columns
        int acc = 0; uint32_t acc_phase = 0;
        int xstage = 0; uint32_t xphase = 0;                     // XF: position in the operand ring (same walk as producer / MMA)
        Iter it(sh, cluster_id, n_clusters);
        int tile_mn, kb0, kb1;
        while (it.next(sh, tile_mn, kb0, kb1)) {
            const int m_blk = (tile_mn % sh.n_mp) * CL + cta_rank, n_blk = tile_mn / sh.n_mp;
            if constexpr (XF != 0) {
                // ===== operand transform: this CTA's A tile of every k block, in place, before the MMA may read it =====
                // 512 threads x 32 bytes = the 128 smem rows x 128 bytes of the tile.  Thread -> smem row R, 16-column chunk c4 of
                // the row (two 16-byte units; SWIZZLE_128B keeps a unit intact and moves it to slot unit ^ (R & 7)).
                const int te = (int)threadIdx.x - 64;
                const int R = te >> 2, c4 = te & 3;
                const uint32_t off0 = (uint32_t)R * 128u + ((uint32_t)((2 * c4) ^ (R & 7)) << 4);
                const uint32_t off1 = (uint32_t)R * 128u + ((uint32_t)((2 * c4 + 1) ^ (R & 7)) << 4);
                constexpr float L2E = 1.4426950408889634f;
                // XF = 1: R = token row of the tile (fixed for the item), chunk = kb * 4 + c4 walks with kb
                // XF = 2: R = (box j = R >> 6, token kk = R & 63 of the k block), chunk = (m_blk * 128 + j * 64) / 16 + c4 fixed
                const int xrow = m_blk * BM + R;                              // XF = 1: chunk-local token row
                const int xcol0 = m_blk * BM + (R >> 6) * 64 + c4 * 16;       // XF = 2: first vocabulary column of this thread
                float nlse = 0.0f; int ytok = -1; bool xok = false;           // XF = 1: per item
                if (XF == 1) {
                    xok = xrow < sh.M;
                    if (xok) { nlse = -__ldg(ep.xf_lse + ep.row0 + xrow) * L2E; ytok = __ldg(ep.y + ep.row0 + xrow); }
                }
                float csum[16];
                if (XF == 2) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) csum[e] = 0.0f;
                }
                // Per-(row, chunk) scalars of a k block (chunk maximum; XF = 2 also the token's lse and target) come from global
                // memory.  They are fetched FOUR k blocks ahead, right after the arrive of the current one: a load still in flight
                // would stall both its first use and the proxy fence (measured: 30 % of this loop's stall samples with a fetch
                // one block ahead issued before the wait; 16 % at the first use in the dWs kernel with two blocks — the chunk-maximum
                // table comes from DRAM behind the 369 MB operand stream) — this way every load has three k-block periods to land.
                struct XfOp { float cm, ls; int y; bool ok; };
                auto fetch = [&](XfOp& o, int kb) {
                    o.cm = 0.0f; o.ls = 0.0f; o.y = -1; o.ok = false;
                    if (kb >= kb1) return;
                    if (XF == 1) {
                        const int c16 = kb * 4 + c4;
                        o.ok = xok && c16 < ep.n_c16;
                        if (o.ok) o.cm = __ldg(ep.cmaxT + (int64_t)c16 * ep.ld_cmax + xrow);
                    } else {
                        const int tok = kb * BK + (R & 63);
                        const int c16 = xcol0 >> 4;
                        o.ok = tok < sh.K && c16 < ep.n_c16;
                        if (o.ok) {
                            o.cm = __ldg(ep.cmaxT + (int64_t)c16 * ep.ld_cmax + tok);
                            o.ls = __ldg(ep.xf_lse + ep.row0 + tok);
                            o.y = __ldg(ep.y + ep.row0 + tok);
                        }
                    }
                };
                auto process = [&](const XfOp& o, int kb) {
                    mbar_wait(&raw_full[xstage], xphase);
                    uint8_t* sa = ring + xstage * L::STAGE_BYTES;
                    if (!(ep.xf_diag & 2)) {      // (diagnostics: bit 1 = barrier hand-over only)
                        uint4 u0 = *reinterpret_cast<uint4*>(sa + off0);
                        uint4 u1 = *reinterpret_cast<uint4*>(sa + off1);
                        // scale of this (row, 16-column chunk): exp(cmax - lse), rounded to fp16 like the operand it multiplies
                        const float sc = o.ok ? fast_ex2(XF == 1 ? fmaf(o.cm, L2E, nlse) : (o.cm - o.ls) * L2E) : 0.0f;
                        const __half2 s2 = __float2half2_rn(sc);
                        __half2* h0 = reinterpret_cast<__half2*>(&u0);
                        __half2* h1 = reinterpret_cast<__half2*>(&u1);
#pragma unroll
                        for (int q = 0; q < 4; ++q) { h0[q] = __hmul2(h0[q], s2); h1[q] = __hmul2(h1[q], s2); }
                        // one-hot target: column (XF = 1: ytok - kb * 64, XF = 2: y - first column of the box) relative to this chunk
                        const int tc16 = (XF == 1 ? ytok - kb * BK : o.y - (m_blk * BM + (R >> 6) * 64)) - c4 * 16;
                        if (o.ok && (unsigned)tc16 < 16u) {
                            const __half one = __float2half_rn(1.0f);
                            __half* a0 = reinterpret_cast<__half*>(&u0);
                            __half* a1 = reinterpret_cast<__half*>(&u1);
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                if (tc16 == e) a0[e] = __hsub(a0[e], one);
                                if (tc16 == 8 + e) a1[e] = __hsub(a1[e], one);
                            }
                        }
                        *reinterpret_cast<uint4*>(sa + off0) = u0;
                        *reinterpret_cast<uint4*>(sa + off1) = u1;
                        if (XF == 2 && !(ep.xf_diag & 1)) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float2 f0 = __half22float2(h0[q]), f1 = __half22float2(h1[q]);
                                csum[2 * q] += f0.x; csum[2 * q + 1] += f0.y;
                                csum[8 + 2 * q] += f1.x; csum[8 + 2 * q + 1] += f1.y;
                            }
                        }
                        fence_proxy_async();      // generic-proxy writes -> visible to the tensor core's async-proxy reads
                    }
                    __syncwarp();
                    if (lane == 0) {
                        if (cta_rank == 1) mbar_arrive_remote(&full_bar[xstage], 0); else mbar_arrive(&full_bar[xstage]);
                    }
                    if (++xstage == STAGES) { xstage = 0; xphase ^= 1; }
                };
                XfOp opA, opB, opC, opD;
                fetch(opA, kb0);
                fetch(opB, kb0 + 1);
                fetch(opC, kb0 + 2);
                fetch(opD, kb0 + 3);
                for (int kb = kb0; kb < kb1; kb += 4) {
                    process(opA, kb);
                    fetch(opA, kb + 4);
                    if (kb + 1 < kb1) { process(opB, kb + 1); fetch(opB, kb + 5); }
                    if (kb + 2 < kb1) { process(opC, kb + 2); fetch(opC, kb + 6); }
                    if (kb + 3 < kb1) { process(opD, kb + 3); fetch(opD, kb + 7); }
                }
                if (XF == 2 && ep.xf_db != nullptr && n_blk == 0) {
                    // column sums over this piece's tokens: lanes of a warp that share (te & 3) hold the same columns for 8 tokens
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        float v = csum[e];
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        csum[e] = v;
                    }
                    if (lane < 4) {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            if (xcol0 + e < sh.M) atomicAdd(ep.xf_db + xcol0 + e, ep.alpha * csum[e]);
                    }
                }
            }
            const int row_w0 = m_blk * BM + quad * 32;            // first row of this warp
            const int row = row_w0 + lane;                         // row held by this thread in TMEM
            const bool row_ok = row < sh.M;
            const int cb = n_blk * BN + cg * GCOLS;                // first column of this warp
            const uint32_t t_row = tmem_base + acc * BN + cg * GCOLS + ((uint32_t)(quad * 32) << 16);
            // EPI_LSE: every 16-column chunk keeps its OWN (max, sum exp(v - max)); the pairs are merged after the loop.  A running
            // (max, sum) would chain the chunks (chunk c+1's exponentials wait for chunk c's max): with only four warps per SM
            // sub-partition that dependency, not the issue or MUFU rate, bounded the epilogue.
            float cmx[CHUNKS], csm[CHUNKS];
#pragma unroll
            for (int cc = 0; cc < CHUNKS; ++cc) { cmx[cc] = -INFINITY; csm[cc] = 0.0f; }
            float sq_acc = 0.0f;
            int tgt_col = -1;
            if (EPI == EPI_LSE && row_ok) tgt_col = ep.y[ep.row0 + row] - cb;
            float bias_r[GCOLS / 32];
            if (EPI == EPI_LSE) {     // bias of this warp's columns: loads issued now, parked in smem after the accumulator wait below
#pragma unroll
                for (int i = 0; i < GCOLS / 32; ++i) {
                    const int colb = cb + i * 32 + lane;
                    bias_r[i] = colb < sh.N ? ep.bias[colb] : 0.0f;
                }
            }
            constexpr int NB4 = BN <= 256 ? CHUNKS : 1;   // BN = 512: fetched per chunk instead (register budget; its K loops are long)
            float4 bias4[NB4];
            const bool add_bias = ep.bias != nullptr && kb0 == 0;   // the piece that starts the K range carries the bias
            auto fetch_bias = [&](int cc) -> float4 {               // bias of the 4 columns this lane writes out for chunk cc
                const int colv = cb + cc * CW + 4 * (lane & 3);
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (add_bias) {
                    if (colv + 3 < sh.N) b = *reinterpret_cast<const float4*>(ep.bias + colv);   // bias tensors are 256-B aligned
                    else {
                        b.x = colv < sh.N ? ep.bias[colv] : 0.f; b.y = colv + 1 < sh.N ? ep.bias[colv + 1] : 0.f;
                        b.z = colv + 2 < sh.N ? ep.bias[colv + 2] : 0.f; b.w = colv + 3 < sh.N ? ep.bias[colv + 3] : 0.f;
                    }
                }
                return b;
            };
            if constexpr (EPI == EPI_STORE && BN <= 256) {   // fetched before the accumulator is awaited
#pragma unroll
                for (int cc = 0; cc < CHUNKS; ++cc) bias4[cc] = fetch_bias(cc);
            }
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            uint32_t rbuf[2][16];
            tmem_ld16(t_row, rbuf[0]);
            if (EPI == EPI_LSE) {     // read back per chunk as broadcast float4s
#pragma unroll
                for (int i = 0; i < GCOLS / 32; ++i) bias_s[i * 32 + lane] = bias_r[i];
                __syncwarp();
            }
#pragma unroll
            for (int cc = 0; cc < CHUNKS; ++cc) {
                uint32_t (&r)[16] = rbuf[cc & 1];
                tmem_ld_wait_dep16(r);
                if (cc + 1 < CHUNKS) tmem_ld16(t_row + (cc + 1) * CW, rbuf[(cc + 1) & 1]);   // next chunk streams in meanwhile
                const int col0 = cb + cc * CW;
                if (col0 >= sh.N) continue;   // warp-uniform
                const bool full = col0 + CW <= sh.N;
                if (EPI == EPI_STORE || EPI == EPI_SCATTER) {
                    // phase 1: registers (one row per lane) -> swizzled smem
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4)
                        st4[lane * 4 + (c4 ^ ((lane >> 1) & 3))] =
                            make_float4(ep.alpha * __uint_as_float(r[4 * c4]), ep.alpha * __uint_as_float(r[4 * c4 + 1]),
                                        ep.alpha * __uint_as_float(r[4 * c4 + 2]), ep.alpha * __uint_as_float(r[4 * c4 + 3]));
                    __syncwarp();
                    // phase 2: 4 lanes cover one 64-byte row segment, 8 rows per instruction -> coalesced sectors
                    const int c4 = lane & 3;
                    const int colv = col0 + 4 * c4;
                    const bool vec = full && ep.vec_ok;
                    float4 v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int rr = 8 * k + (lane >> 2);
                        v[k] = st4[rr * 4 + (c4 ^ ((rr >> 1) & 3))];
                    }
                    if (EPI == EPI_SCATTER) {
                        float* C = reinterpret_cast<float*>(ep.C);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int grow = row_w0 + 8 * k + (lane >> 2);
                            if (grow >= sh.M) continue;
                            if (C == nullptr) {     // norm only: the dense gradient comes from the token-sorted segment sums
                                if (full) sq_acc += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
                                else {
                                    const float e4[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e)
                                        if (colv + e < sh.N) sq_acc += e4[e] * e4[e];
                                }
                                continue;
                            }
                            float* dst = C + (int64_t)ep.y[ep.row0 + grow] * ep.ldc + colv;
                            if (vec) {
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[k].x), "f"(v[k].y), "f"(v[k].z), "f"(v[k].w) : "memory");
                                sq_acc += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
                            } else {
                                const float e4[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (colv + e < sh.N) { atomicAdd(dst + e, e4[e]); sq_acc += e4[e] * e4[e]; }
                            }
                        }
                    } else {
                        float4 b4;
                        if constexpr (BN <= 256) b4 = bias4[cc]; else b4 = fetch_bias(cc);
#pragma unroll
                        for (int k = 0; k < 4; ++k) { v[k].x += b4.x; v[k].y += b4.y; v[k].z += b4.z; v[k].w += b4.w; }
                        if (ep.c_half) {
                            __half* C = reinterpret_cast<__half*>(ep.C);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int grow = row_w0 + 8 * k + (lane >> 2);
                                if (grow >= sh.M) continue;
                                __half* dst = C + (int64_t)grow * ep.ldc + colv;
                                if (vec) {
                                    __align__(8) __half2 hh[2] = {__floats2half2_rn(v[k].x, v[k].y), __floats2half2_rn(v[k].z, v[k].w)};
                                    *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<uint2*>(hh);
                                } else {
                                    if (colv < sh.N) dst[0] = __float2half_rn(v[k].x);
                                    if (colv + 1 < sh.N) dst[1] = __float2half_rn(v[k].y);
                                    if (colv + 2 < sh.N) dst[2] = __float2half_rn(v[k].z);
                                    if (colv + 3 < sh.N) dst[3] = __float2half_rn(v[k].w);
                                }
                            }
                        } else {
                            float* C = reinterpret_cast<float*>(ep.C);
                            if (vec && ep.accumulate && !ep.atomic) {
                                float4 o[4];   // all (coalesced) loads in flight before the first dependent add/store
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const int grow = row_w0 + 8 * k + (lane >> 2);
                                    o[k] = grow < sh.M ? __ldcg(reinterpret_cast<const float4*>(C + (int64_t)grow * ep.ldc + colv)) : make_float4(0.f, 0.f, 0.f, 0.f);
                                }
#pragma unroll
                                for (int k = 0; k < 4; ++k) { v[k].x += o[k].x; v[k].y += o[k].y; v[k].z += o[k].z; v[k].w += o[k].w; }
                            }
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int grow = row_w0 + 8 * k + (lane >> 2);
                                if (grow >= sh.M) continue;
                                float* dst = C + (int64_t)grow * ep.ldc + colv;
                                if (vec) {
                                    if (ep.atomic)
                                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[k].x), "f"(v[k].y), "f"(v[k].z), "f"(v[k].w) : "memory");
                                    else
                                        *reinterpret_cast<float4*>(dst) = v[k];
                                } else {
                                    const float e4[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        if (colv + e < sh.N) {
                                            if (ep.atomic) atomicAdd(dst + e, e4[e]);
                                            else if (ep.accumulate) dst[e] += e4[e];
                                            else dst[e] = e4[e];
                                        }
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();   // staging buffer is reused by the next chunk
                } else {  // EPI_LSE
                    float v[CW];
                    float cmax = -INFINITY;
                    if (full) {   // warp-uniform fast path: no per-element column guards
#pragma unroll
                        for (int j = 0; j < CW; j += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + cc * CW + j);
                            v[j] = __uint_as_float(r[j]) + b4.x;
                            v[j + 1] = __uint_as_float(r[j + 1]) + b4.y;
                            v[j + 2] = __uint_as_float(r[j + 2]) + b4.z;
                            v[j + 3] = __uint_as_float(r[j + 3]) + b4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < CW; j += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + cc * CW + j);
                            v[j] = (col0 + j < sh.N) ? __uint_as_float(r[j]) + b4.x : -INFINITY;
                            v[j + 1] = (col0 + j + 1 < sh.N) ? __uint_as_float(r[j + 1]) + b4.y : -INFINITY;
                            v[j + 2] = (col0 + j + 2 < sh.N) ? __uint_as_float(r[j + 2]) + b4.z : -INFINITY;
                            v[j + 3] = (col0 + j + 3 < sh.N) ? __uint_as_float(r[j + 3]) + b4.w : -INFINITY;
                        }
                    }
                    {
                        const float m0 = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), m1 = fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7]));
                        const float m2 = fmaxf(fmaxf(v[8], v[9]), fmaxf(v[10], v[11])), m3 = fmaxf(fmaxf(v[12], v[13]), fmaxf(v[14], v[15]));
                        cmax = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    }
                    const int tj = tgt_col - cc * CW;
                    if (row_ok && tj >= 0 && tj < CW) {
                        float tv = 0.0f;
#pragma unroll
                        for (int j = 0; j < CW; ++j) tv = (j == tj) ? v[j] : tv;
                        ep.tgt[row] = tv;
                    }
                    const float nmax_l2 = cmax * 1.4426950408889634f;
                    float sa0 = 0.0f, sa1 = 0.0f, sa2 = 0.0f, sa3 = 0.0f;   // independent chains
                    if (ep.exp_store) {
                        // fused softmax gradient: the chunk buffer receives the exponentials (in (0, 1], the chunk's largest = 1) and
                        // the consumers rescale them by exp(cmax - lse); v[] is dead after the target pick above
#pragma unroll
                        for (int j = 0; j < CW; j += 4) {
                            v[j] = fast_ex2(fmaf(v[j], 1.4426950408889634f, -nmax_l2));
                            v[j + 1] = fast_ex2(fmaf(v[j + 1], 1.4426950408889634f, -nmax_l2));
                            v[j + 2] = fast_ex2(fmaf(v[j + 2], 1.4426950408889634f, -nmax_l2));
                            v[j + 3] = fast_ex2(fmaf(v[j + 3], 1.4426950408889634f, -nmax_l2));
                            sa0 += v[j]; sa1 += v[j + 1]; sa2 += v[j + 2]; sa3 += v[j + 3];
                        }
                        if (row_ok) ep.cmaxT[(int64_t)(col0 >> 4) * ep.ld_cmax + row] = cmax;   // 32 rows of a warp: one 128-byte line
                    } else {
#pragma unroll
                        for (int j = 0; j < CW; j += 4) {
                            sa0 += fast_ex2(fmaf(v[j], 1.4426950408889634f, -nmax_l2));          // one FFMA + MUFU.EX2 per logit
                            sa1 += fast_ex2(fmaf(v[j + 1], 1.4426950408889634f, -nmax_l2));
                            sa2 += fast_ex2(fmaf(v[j + 2], 1.4426950408889634f, -nmax_l2));
                            sa3 += fast_ex2(fmaf(v[j + 3], 1.4426950408889634f, -nmax_l2));
                        }
                    }
                    cmx[cc] = cmax;
                    csm[cc] = (sa0 + sa1) + (sa2 + sa3);
                    if (ep.logits16 && row_ok) {
                        // fp16 logits for the backward: this lane's 16 columns are 32 contiguous bytes of its row = one full sector:
                        // ONE 256-bit store, no smem transpose (the staged variant spent a third of the chunk's stall samples on
                        // STS -> syncwarp -> LDS -> STG)
                        __half* dst = ep.logits16 + (int64_t)row * ep.ld16 + col0;
                        uint32_t pk[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const __half2 h2 = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
                            pk[q] = *reinterpret_cast<const uint32_t*>(&h2);
                        }
                        if (full && ep.vec_ok) {      // host-checked: base 32-B aligned, ld16 % 16 == 0
                            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                                         "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                                         : "memory");
                        } else {
                            const __half* ph = reinterpret_cast<const __half*>(pk);
#pragma unroll
                            for (int e = 0; e < CW; ++e)
                                if (col0 + e < sh.N) dst[e] = ph[e];
                        }
                    }
                }
            }
            if (EPI == EPI_LSE && row_ok) {
                float gmax = cmx[0];
#pragma unroll
                for (int cc = 1; cc < CHUNKS; ++cc) gmax = fmaxf(gmax, cmx[cc]);
                float gsum = 0.0f;
                if (gmax > -INFINITY) {
#pragma unroll
                    for (int cc = 0; cc < CHUNKS; ++cc)    // chunks past the last column carry (-inf, 0): ex2(-inf) = 0
                        gsum += csm[cc] * fast_ex2((cmx[cc] - gmax) * 1.4426950408889634f);
                }
                ep.part[(int64_t)row * ep.n_tiles_total + n_blk * 4 + cg] = make_float2(gmax, gsum);
            }
            if (EPI == EPI_SCATTER) {
                sq_acc = warp_sum(sq_acc);
                if (lane == 0 && sq_acc != 0.0f) atomicAdd(ep.tgt, sq_acc);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {   // all epilogue warps (of both CTAs when paired) free the accumulator stage for the leader's MMA thread
                if (CL == 2 && cta_rank == 1) mbar_arrive_remote(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]);
            }
            if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();      // no CTA exits while its peer may still read its smem / signal its barriers
    if (warp == 1) { if (CL == 2) tmem_dealloc_2sm(tmem_base, ACC * BN); else tmem_dealloc(tmem_base, ACC * BN); }
}

// combine the per-(row, n-tile) (max, sumexp) partials: lse = M + log(sum_i s_i * exp(m_i - M)); nll = lse - tgt
__global__ void __launch_bounds__(256) lse_combine_kernel(const float2* __restrict__ part, int n_tiles, const float* __restrict__ tgt,
                                                          int64_t row0, int rows, int N, int T, float* __restrict__ lse_out,
                                                          float* __restrict__ nll_out, int* __restrict__ sched) {
    // (also resets the strip scheduler of the background softmax-grad pass that follows: no memset node on the side stream)
    if (sched != nullptr && blockIdx.x == 0)
        for (int i = threadIdx.x; i < 1032; i += blockDim.x) sched[i] = 0;
    // one warp per row: the row's partials are contiguous (n_tiles x 8 B), so every load instruction is one coalesced line
    // (one thread per row walked 2.5 KB-strided addresses: 27 us per 13 k-row chunk, 0.37 ms per step)
    const int lr = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (lr >= rows) return;
    const float2* p = part + (int64_t)lr * n_tiles;
    float m = -INFINITY;
    for (int i = lane; i < n_tiles; i += 32) m = fmaxf(m, __ldcg(&p[i]).x);
    m = warp_max(m);
    float s = 0.0f;
    for (int i = lane; i < n_tiles; i += 32) { const float2 v = __ldcg(&p[i]); s += v.y * expf(v.x - m); }
    s = warp_sum(s);
    if (lane == 0) {
        const float lse = m + logf(s);
        const int64_t r = row0 + lr;
        if (lse_out) lse_out[r] = lse;
        const int t = (int)(r / N), n = (int)(r % N);
        if (nll_out) nll_out[(int64_t)n * T + t] = lse - tgt[lr];
    }
}

// In-place pass over a chunk of fp16 logits (training): dlogits = exp(logit - lse) - onehot(y) (unscaled, fp16) and
// db_s += alpha * column sums of dlogits (softmax_b gradient).  Strip version: a CTA of 128 threads owns a strip of ROWS rows x 1024 columns; each
// thread owns 8 fixed columns and walks down the strip with 8 rows of 16-byte loads in flight, so column sums stay
// in 8 registers and the kernel needs <= 80 registers and 256 B of smem: it can share an SM with a GEMM CTA.
template <int ROWS>
__device__ __forceinline__ void softmax_grad_strip_item(__half* __restrict__ logits, int64_t ld, int vp1, const float* __restrict__ lse_all,
                                                        const int32_t* __restrict__ y, int64_t row0, int rows, float alpha,
                                                        float* __restrict__ db, int bx, int by, float* s_lse, int* s_tgt) {
    const int r_begin = by * ROWS;
    const int n_rows = min(ROWS, rows - r_begin);
    if (threadIdx.x < n_rows) {   // the row-wise log-sum-exp was combined by lse_combine_kernel just before
        s_lse[threadIdx.x] = lse_all[row0 + r_begin + threadIdx.x];
        s_tgt[threadIdx.x] = y[row0 + r_begin + threadIdx.x];
    }
    __syncthreads();
    const int v0 = (bx * 128 + threadIdx.x) * 8;
    if (v0 >= ld) return;
    float csum[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) csum[e] = 0.0f;
    __half* base = logits + (int64_t)r_begin * ld + v0;
    const bool interior = v0 + 8 <= vp1;
    constexpr int RB = 8;
    for (int lr0 = 0; lr0 < n_rows; lr0 += RB) {
        uint4 raw[RB];
#pragma unroll
        for (int b = 0; b < RB; ++b)
            if (lr0 + b < n_rows) raw[b] = __ldcg(reinterpret_cast<const uint4*>(base + (int64_t)(lr0 + b) * ld));
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            if (lr0 + b < n_rows) {
                const float nl2 = -s_lse[lr0 + b] * 1.4426950408889634f;      // exp(x - lse) = ex2(x*log2e - lse*log2e): one FFMA + MUFU
                const int tg = s_tgt[lr0 + b] - v0;
                __half2* h2 = reinterpret_cast<__half2*>(&raw[b]);
                if (interior && (unsigned)tg >= 8u) {      // fast path: no column guards, no target in these 8 columns
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 f = __half22float2(h2[q]);
                        const float a = fast_ex2(fmaf(f.x, 1.4426950408889634f, nl2));
                        const float bb = fast_ex2(fmaf(f.y, 1.4426950408889634f, nl2));
                        h2[q] = __floats2half2_rn(a, bb);
                        csum[2 * q] += a;
                        csum[2 * q + 1] += bb;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 f = __half22float2(h2[q]);
                        const int v = v0 + 2 * q;
                        const float a = v < vp1 ? fast_ex2(fmaf(f.x, 1.4426950408889634f, nl2)) - (2 * q == tg ? 1.0f : 0.0f) : 0.0f;
                        const float bb = v + 1 < vp1 ? fast_ex2(fmaf(f.y, 1.4426950408889634f, nl2)) - (2 * q + 1 == tg ? 1.0f : 0.0f) : 0.0f;
                        h2[q] = __floats2half2_rn(a, bb);
                        csum[2 * q] += a;
                        csum[2 * q + 1] += bb;
                    }
                }
                *reinterpret_cast<uint4*>(base + (int64_t)(lr0 + b) * ld) = raw[b];
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (v0 + e < vp1) atomicAdd(db + v0 + e, alpha * csum[e]);
}

template <int ROWS>
__global__ void __launch_bounds__(128) softmax_grad_strip_kernel(__half* __restrict__ logits, int64_t ld, int vp1,
                                                                    const float* __restrict__ lse_all, const int32_t* __restrict__ y,
                                                                    int64_t row0, int rows, float alpha, float* __restrict__ db) {
    __shared__ float s_lse[ROWS];
    __shared__ int s_tgt[ROWS];
    softmax_grad_strip_item<ROWS>(logits, ld, vp1, lse_all, y, row0, rows, alpha, db, blockIdx.x, blockIdx.y, s_lse, s_tgt);
}

// Streaming flavour (default): ONE wave of CTAs (6 per SM x 128 threads); CTA (bx, by) owns 1024 columns x one long row segment and
// walks down it software-pipelined — batch b+1 (4 rows of 16-byte loads per thread, plus the rows' lse / target) is in flight while
// batch b is exponentiated and stored, so every thread always has loads outstanding; no shared memory, no barriers.  All CTAs are
// co-resident and own equal work, so there is no partial last wave (the strip grids lost up to 20 % to it).
template <int RB>
__global__ void __launch_bounds__(128, 6) softmax_grad_stream_kernel(__half* __restrict__ logits, int64_t ld, int vp1,
                                                                     const float* __restrict__ lse_all, const int32_t* __restrict__ y,
                                                                     int64_t row0, int rows, int seg_rows, float alpha,
                                                                     float* __restrict__ db) {
    const int v0 = (blockIdx.x * 128 + threadIdx.x) * 8;
    if (v0 >= ld) return;
    const int r_begin = blockIdx.y * seg_rows;
    const int r_end = min(rows, r_begin + seg_rows);
    if (r_begin >= r_end) return;
    __half* base = logits + (int64_t)r_begin * ld + v0;
    const float* lse_p = lse_all + row0 + r_begin;
    const int32_t* y_p = y + row0 + r_begin;
    const int n = r_end - r_begin;
    const bool interior = v0 + 8 <= vp1;
    float csum[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) csum[e] = 0.0f;
    uint4 buf[2][RB];
    float nl2[2][RB];
    int tgc[2][RB];
    auto load = [&](int slot, int r) {
#pragma unroll
        for (int b = 0; b < RB; ++b)
            if (r + b < n) {
                buf[slot][b] = __ldcg(reinterpret_cast<const uint4*>(base + (int64_t)(r + b) * ld));
                nl2[slot][b] = -__ldg(lse_p + r + b) * 1.4426950408889634f;
                tgc[slot][b] = __ldg(y_p + r + b) - v0;
            }
    };
    auto process = [&](int slot, int r) {
#pragma unroll
        for (int b = 0; b < RB; ++b)
            if (r + b < n) {
                __half2* h2 = reinterpret_cast<__half2*>(&buf[slot][b]);
                const float l2 = nl2[slot][b];
                const int tg = tgc[slot][b];
                if (interior && (unsigned)tg >= 8u) {      // fast path: no column guards, no target in these 8 columns
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 f = __half22float2(h2[q]);
                        const float a = fast_ex2(fmaf(f.x, 1.4426950408889634f, l2));
                        const float bb = fast_ex2(fmaf(f.y, 1.4426950408889634f, l2));
                        h2[q] = __floats2half2_rn(a, bb);
                        csum[2 * q] += a;
                        csum[2 * q + 1] += bb;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 f = __half22float2(h2[q]);
                        const int v = v0 + 2 * q;
                        const float a = v < vp1 ? fast_ex2(fmaf(f.x, 1.4426950408889634f, l2)) - (2 * q == tg ? 1.0f : 0.0f) : 0.0f;
                        const float bb = v + 1 < vp1 ? fast_ex2(fmaf(f.y, 1.4426950408889634f, l2)) - (2 * q + 1 == tg ? 1.0f : 0.0f) : 0.0f;
                        h2[q] = __floats2half2_rn(a, bb);
                        csum[2 * q] += a;
                        csum[2 * q + 1] += bb;
                    }
                }
                *reinterpret_cast<uint4*>(base + (int64_t)(r + b) * ld) = buf[slot][b];
            }
    };
    load(0, 0);
    for (int r = 0; r < n; r += 2 * RB) {
        load(1, r + RB);
        process(0, r);
        load(0, r + 2 * RB);
        process(1, r + RB);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (v0 + e < vp1) atomicAdd(db + v0 + e, alpha * csum[e]);
}

// Persistent flavour (default): 6 CTAs per SM walk the strips with a static stride (column block fastest, so the CTAs running at
// any moment cover the same rows: DRAM-page friendly).  The plain grid's last partial wave of CTAs ran at a fraction of the
// occupancy (measured with 18 432-row chunks: 64-row strips = 3.24 waves 1.94 ms, 32-row strips = 6.49 waves 1.59 ms per step).
template <int ROWS>
__global__ void __launch_bounds__(128, 6) softmax_grad_strip_loop_kernel(__half* __restrict__ logits, int64_t ld, int vp1,
                                                                         const float* __restrict__ lse_all, const int32_t* __restrict__ y,
                                                                         int64_t row0, int rows, float alpha, float* __restrict__ db) {
    __shared__ float s_lse[ROWS];
    __shared__ int s_tgt[ROWS];
    const int n_bx = (int)((ld + 1023) / 1024);
    const int n_items = n_bx * ((rows + ROWS - 1) / ROWS);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        __syncthreads();                                   // previous strip done with s_lse / s_tgt
        softmax_grad_strip_item<ROWS>(logits, ld, vp1, lse_all, y, row0, rows, alpha, db, item % n_bx, item / n_bx, s_lse, s_tgt);
    }
}

// Same pass as a BACKGROUND kernel that shares the SMs with a persistent GEMM (the dH / dWs GEMMs of the previous chunk run
// on the caller's stream meanwhile): the GEMM CTA leaves 10 240 registers and ~1.5 KB of shared memory per SM — room for
// exactly one of these CTAs (128 threads x 80 registers, 256 B).  Strips are handed out by a global counter, and at most
// `max_per_sm` CTAs stay alive per SM (claimed through %smid; the others exit at once), so that however the block scheduler
// spreads the grid, a GEMM CTA arriving later always finds its registers free.
//   sched[0] = next strip, sched[1 + smid] = CTAs that claimed SM smid   (zeroed by the host before the launch)
template <int ROWS>
__global__ void __launch_bounds__(128, 6) softmax_grad_strip_bg_kernel(__half* __restrict__ logits, int64_t ld, int vp1,
                                                                       const float* __restrict__ lse_all, const int32_t* __restrict__ y,
                                                                       int64_t row0, int rows, float alpha, float* __restrict__ db,
                                                                       int* __restrict__ sched, int max_per_sm) {
    __shared__ float s_lse[ROWS];
    __shared__ int s_tgt[ROWS];
    __shared__ int s_item;
    const int n_bx = (int)((ld + 1023) / 1024);
    const int n_items = n_bx * ((rows + ROWS - 1) / ROWS);
    if (threadIdx.x == 0) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_item = atomicAdd(&sched[1 + (smid & 1023)], 1) < max_per_sm ? 0 : -1;
    }
    __syncthreads();
    if (s_item < 0) return;
    for (;;) {
        __syncthreads();                                   // previous strip done with s_lse / s_tgt / s_item
        if (threadIdx.x == 0) s_item = atomicAdd(&sched[0], 1);
        __syncthreads();
        const int item = s_item;
        if (item >= n_items) break;
        softmax_grad_strip_item<ROWS>(logits, ld, vp1, lse_all, y, row0, rows, alpha, db, item % n_bx, item / n_bx, s_lse, s_tgt);
    }
}

}  // namespace tc

// =====================================================================================================
// host side
// =====================================================================================================
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int STRIP_ROWS = 32;    // rows per strip of the background softmax-grad kernel

struct TcContext {
    bool ready = false;
    int num_sms = 148;
    PFN_tensorMapEncodeTiled encode = nullptr;
    // projection scratch (carved from the caller's workspace)
    float2* part = nullptr;
    float* tgt = nullptr;
    int part_tiles = 0;
    int* counters = nullptr;   // [256] group-progress counters of the persistent recurrent kernels
    int* strip_sched = nullptr;
    int strip_per_sm = 1;      // live background softmax-grad CTAs per SM (FSMG_STRIP_PER_SM)
    int strip_mode = 2, strip_param = 2, strip_waves = 6;   // softmax-grad pass variant (FSMG_STRIP=mode,param,waves; see tc_softmax_grad_launch)
    long long* trace = nullptr; // [128] debug timeline (FSMG_TRACE=1)
    int enabled = 1;
    int cluster = 2;           // CTAs per cluster of the GEMM core (2 = cta_group::2 pairs, 1 = single-SM MMAs)
    int cluster_lse = 0;       // override for the logits + log-sum-exp GEMM (0 = same as `cluster`)
    int astat = 0;             // A-stationary schedule for the logits GEMM when K <= 512 (FSMG_ASTAT=1; measured slower: the kernel is
                               // epilogue-bound, not operand-bound: 1.96 vs 1.77 ms per step)
    int wide = 1;              // 256 x 512 pair tiles for plain-store GEMMs with long K loops (FSMG_WIDE=0 disables)
    int lstm_cluster = 1;      // CTAs per cluster of the persistent recurrent kernels (1, 2 or 4: operand multicast)
    int lstm_split = 1;        // forward recurrent kernel: two interleaved half-groups per CTA (FSMG_LSTM_SPLIT=0: one lock-step group)
    int lstm_pair = 0;         // persistent backward kernel as cta_group::2 pairs (each CTA ingests half of the exchanged rows)
    int lstm_bsplit = -1;      // backward recurrent kernel (FSMG_LSTM_BSPLIT): -1 auto, 0 first-generation pair kernel, 1 lstm_bwd_pair2<NSUB=1>, 2 <NSUB=2>
    int lstm_ks = 0;           // backward pair2 kernel: K chunks per ring stage (FSMG_LSTM_KS = 1, 2, 4, 8; 0 = auto)
    int lstm_half_m = 1;       // backward pair2 NSUB = 2: 128-row pair MMAs when a sub-group has <= 64 rows per CTA (FSMG_LSTM_HALF_M=0: 256-row)
    int lstm_nh = 0;           // forward split kernel: halves in use (FSMG_LSTM_NH = 1, 2; 0 = auto: 1 up to 32 rows per group)
    int lstm_pub_cta = 0;      // forward split kernel: one release per (CTA, half) instead of per warp (FSMG_LSTM_PUB_CTA=1)
    int lstm_fks = 4;          // forward split kernel: K chunks per ring stage (FSMG_LSTM_FKS = 1, 2, 4)
    int lstm_rot = 3;          // rotated K-chunk order per loader in the recurrent kernels (bit 0: backward, bit 1: forward)
    int pdl = 0;               // launch with programmatic stream serialization (set by the decode loop around its GEMMs)
    int narrow = 0;            // plan 128-wide N tiles even for wide N (set by the decode loop: FSMG_SAMPLE_BN=128)
    int xf_diag = 0;           // FSMG_XF_DIAG: timing diagnostics of the operand-transform GEMMs (results are wrong when set)
    int streamk = 1;           // stream-K scheduling of atomically-combined GEMMs when plain tiling quantises badly
    int lstm_reserve_sms = 0;  // SMs the persistent recurrent kernels leave free (for the NCCL kernels of an overlapped gradient all-reduce)
};

template <typename B>
static inline void tc_carve(TcContext& c, B& b, int /*Nmax*/, int /*T*/, int V1, int /*Vp*/, int /*H*/, int chunk_rows) {
    c.part_tiles = 4 * cdiv(V1, 128);   // four column groups per N tile
    c.part = b.template take<float2>((int64_t)chunk_rows * c.part_tiles);
    c.tgt = b.template take<float>(chunk_rows);
    c.counters = b.template take<int>(256);
    c.strip_sched = b.template take<int>(1032);   // [0] next strip, [1 + smid] claims (background softmax-grad pass)
    c.trace = b.template take<long long>(128);
}

static inline int tc_init(TcContext& c) {
    if (c.ready) return 0;
    const char* env = getenv("FSMG_TC");
    c.enabled = env ? atoi(env) : 1;
    const char* envc = getenv("FSMG_CLUSTER");
    c.cluster = envc ? (atoi(envc) == 1 ? 1 : 2) : 2;
    const char* envcl = getenv("FSMG_CLUSTER_LSE");
    c.cluster_lse = envcl ? atoi(envcl) : 0;
    const char* enva = getenv("FSMG_ASTAT");
    c.astat = enva ? atoi(enva) : 0;
    const char* envw = getenv("FSMG_WIDE");
    c.wide = envw ? atoi(envw) : 1;
    const char* envxd = getenv("FSMG_XF_DIAG");
    c.xf_diag = envxd ? atoi(envxd) : 0;
    const char* envs = getenv("FSMG_STREAMK");
    c.streamk = envs ? atoi(envs) : 1;
    const char* envp = getenv("FSMG_LSTM_PAIR");
    c.lstm_pair = envp ? atoi(envp) : 1;   // bit 0: backward (measured 2.64 -> 1.99 ms), bit 1: forward (neutral)
    const char* envbs = getenv("FSMG_LSTM_BSPLIT");
    c.lstm_bsplit = envbs ? atoi(envbs) : -1;
    const char* envks = getenv("FSMG_LSTM_KS");
    if (envks && atoi(envks) > 0) c.lstm_ks = atoi(envks);
    const char* envhm = getenv("FSMG_LSTM_HALF_M");
    if (envhm) c.lstm_half_m = atoi(envhm);
    const char* envnh = getenv("FSMG_LSTM_NH");
    if (envnh) c.lstm_nh = atoi(envnh);
    const char* envpc = getenv("FSMG_LSTM_PUB_CTA");
    if (envpc) c.lstm_pub_cta = atoi(envpc);
    const char* envfks = getenv("FSMG_LSTM_FKS");
    if (envfks && atoi(envfks) > 0) c.lstm_fks = atoi(envfks);
    const char* envr = getenv("FSMG_LSTM_ROT");
    if (envr) c.lstm_rot = atoi(envr);
    const char* envst = getenv("FSMG_STRIP");
    if (envst) sscanf(envst, "%d,%d,%d", &c.strip_mode, &c.strip_param, &c.strip_waves);
    const char* envps = getenv("FSMG_STRIP_PER_SM");
    if (envps && atoi(envps) > 0) c.strip_per_sm = atoi(envps);
    const char* envsp = getenv("FSMG_LSTM_SPLIT");
    c.lstm_split = envsp ? atoi(envsp) : 1;
    const char* envl = getenv("FSMG_LSTM_CLUSTER");
    c.lstm_cluster = envl ? atoi(envl) : 1;
    int dev = 0;
    FSMG_CUDA_OK(cudaGetDevice(&dev));
    FSMG_CUDA_OK(cudaDeviceGetAttribute(&c.num_sms, cudaDevAttrMultiProcessorCount, dev));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FSMG_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return set_error(-2, "cuTensorMapEncodeTiled not available from the driver");
    c.encode = reinterpret_cast<PFN_tensorMapEncodeTiled>(fn);
#define FSMG_SET_SMEM(BN, EPI, AM, BMN)                                                                                  \
    FSMG_CUDA_OK(cudaFuncSetAttribute(tc::tc_gemm_kernel<BN, EPI, AM, BMN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      tc::SmemLayout<BN, 1>::TOTAL));                                                    \
    FSMG_CUDA_OK(cudaFuncSetAttribute(tc::tc_gemm_kernel<BN, EPI, AM, BMN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      tc::SmemLayout<BN, 2>::TOTAL))
    FSMG_CUDA_OK(cudaFuncSetAttribute(tc::tc_gemm_kernel<512, tc::EPI_STORE, false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SmemLayout<512, 2>::TOTAL));
    FSMG_CUDA_OK(cudaFuncSetAttribute(tc::tc_gemm_kernel<512, tc::EPI_STORE, true, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SmemLayout<512, 2>::TOTAL));
    FSMG_CUDA_OK(cudaFuncSetAttribute(tc::tc_gemm_kernel<512, tc::EPI_STORE, false, false, 2, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SmemLayout<512, 2>::TOTAL));
    FSMG_CUDA_OK(cudaFuncSetAttribute(tc::tc_gemm_kernel<512, tc::EPI_STORE, true, true, 2, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SmemLayout<512, 2>::TOTAL));
    FSMG_SET_SMEM(256, tc::EPI_STORE, false, false);
    FSMG_SET_SMEM(256, tc::EPI_STORE, true, true);
    FSMG_SET_SMEM(128, tc::EPI_STORE, false, false);
    FSMG_SET_SMEM(128, tc::EPI_STORE, true, true);
    FSMG_SET_SMEM(256, tc::EPI_LSE, false, false);
    FSMG_CUDA_OK(cudaFuncSetAttribute(tc::tc_gemm_kernel<256, tc::EPI_LSE, false, false, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SmemLayout<256, 2, true>::TOTAL));
    FSMG_SET_SMEM(256, tc::EPI_SCATTER, false, false);
    FSMG_SET_SMEM(128, tc::EPI_SCATTER, false, false);
    FSMG_SET_SMEM(128, tc::EPI_LSE, false, false);
#undef FSMG_SET_SMEM
    // the small kernels that run between / beside the persistent GEMMs of the projection ask for the same (maximum) shared-memory
    // carve-out as the GEMMs: an SM whose carve-out had been shrunk for them would have to drain before a GEMM CTA could land on it
    FSMG_CUDA_OK(cudaFuncSetAttribute(tc::lse_combine_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    FSMG_CUDA_OK(cudaFuncSetAttribute(tc::softmax_grad_strip_bg_kernel<STRIP_ROWS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    FSMG_CUDA_OK(cudaFuncSetAttribute(tc::softmax_grad_strip_kernel<STRIP_ROWS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    c.ready = true;
    return 0;
}

// fp16 2-D tensor map with 128-byte swizzle.  inner = contiguous extent (elements), outer = rows, ld in elements.
static inline int make_map_f16(const TcContext& c, CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                               uint32_t box_inner, uint32_t box_outer) {
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = c.encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(-2, "cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu outer=%llu ld=%llu box=%ux%u", (int)r, base,
                         (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
    return 0;
}

static inline bool tc_operands_ok(const void* p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld % 8) == 0; }

static inline bool tc_gemm_supported(const GemmArgs& g, bool a_mn, bool b_mn) {
    if (a_mn != b_mn) return false;  // only NT (both K-major) and TN-of-tokens (both MN-major) occur on the hot path
    if (!tc_operands_ok(g.A, g.lda) || !tc_operands_ok(g.B, g.ldb)) return false;
    if (g.c_half && g.accumulate) return false;
    const char* env = getenv("FSMG_TC");
    if (env && atoi(env) == 0) return false;
    return true;
}

struct TcPlan {
    int bn;
    tc::GemmShape sh;
    int grid;
    int cl;   // cluster size (1 or 2)
};

static inline TcPlan tc_plan(const TcContext& c, int M, int N, int K, bool allow_split, int cl_pref = 0, bool allow_wide = false) {
    TcPlan p;
    p.bn = (N > 128 && !c.narrow) ? 256 : 128;
    tc::GemmShape& sh = p.sh;
    sh.M = M; sh.N = N; sh.K = K;
    sh.n_m = cdiv(M, tc::BM);
    sh.kb_total = cdiv(K, tc::BK);
    p.cl = ((cl_pref ? cl_pref : c.cluster) == 2 && sh.n_m >= 2) ? 2 : 1;     // pairs need two M tiles that share a B tile
    // 256 x 512 pair tiles (A fetched once per 512 columns, single-buffered accumulator): long K loops only, little N padding
    if (allow_wide && c.wide && p.cl == 2 && N >= 384 && sh.kb_total >= 64 && cdiv(N, 512) * 512 - N < 128) p.bn = 512;
    sh.n_n = cdiv(N, p.bn);
    sh.n_mp = cdiv(sh.n_m, p.cl);
    const int slots = c.num_sms / p.cl;                  // concurrently resident clusters
    int tiles = sh.n_mp * sh.n_n;                        // cluster-tiles
    int split = 1;
    if (allow_split && tiles < slots) {
        split = slots / tiles;                           // fill the machine once
        int max_split = sh.kb_total / 8;                 // keep >= 8 k-blocks (512 k) per slice
        if (split > max_split) split = max_split;
        if (split < 1) split = 1;
    }
    sh.kb_per_split = cdiv(sh.kb_total, split);
    sh.n_s = cdiv(sh.kb_total, sh.kb_per_split);
    int total = tiles * sh.n_s;
    p.grid = (total < slots ? total : slots) * p.cl;
    sh.streamk = 0;
    sh.units_per_cluster = 0;
    if (allow_split && c.streamk) {
        // wave quantisation of the plain schedule; below 80 % switch to stream-K (perfectly balanced k-block ranges)
        const int rounds = cdiv(total, slots);
        const double eff = (double)total / ((double)rounds * slots);
        const int64_t units = (int64_t)tiles * sh.kb_total;
        if (eff < 0.80 && units >= 16) {
            int ncl = slots;
            int upc = (int)cdiv(units, ncl);
            if (upc < 8) { upc = 8; ncl = (int)cdiv(units, upc); }
            sh.streamk = 1;
            sh.units_per_cluster = upc;
            sh.n_s = 2;                       // marks "partial sums are combined with atomics" for the caller
            p.grid = (int)cdiv(units, upc) * p.cl;
        }
    }
    return p;
}

template <typename K>
static inline int tc_launch_kernel(K kernel, const TcPlan& p, int smem, const CUtensorMap& ma, const CUtensorMap& mb,
                                   const tc::EpiParams& ep, cudaStream_t s, int pdl = 0) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(p.grid);
    cfg.blockDim = dim3(tc::NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (p.cl > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = p.cl; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    FSMG_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, ma, mb, p.sh, ep));
    return 0;
}

template <int EPI>
static inline int tc_launch(const TcContext& c, const TcPlan& p, const CUtensorMap& ma, const CUtensorMap& mb, bool mn,
                            const tc::EpiParams& ep, cudaStream_t s) {
#define FSMG_GO(BN, MN)                                                                                                   \
    (p.cl == 2 ? tc_launch_kernel(tc::tc_gemm_kernel<BN, EPI, MN, MN, 2>, p, tc::SmemLayout<BN, 2>::TOTAL, ma, mb, ep, s, c.pdl) \
               : tc_launch_kernel(tc::tc_gemm_kernel<BN, EPI, MN, MN, 1>, p, tc::SmemLayout<BN, 1>::TOTAL, ma, mb, ep, s, c.pdl))
    int rc = 0;
    if constexpr (EPI == tc::EPI_LSE || EPI == tc::EPI_SCATTER) {
        rc = (p.bn == 256) ? FSMG_GO(256, false) : FSMG_GO(128, false);
    } else {
        if (p.bn == 512) {   // cta_group::2 only
            rc = !mn ? tc_launch_kernel(tc::tc_gemm_kernel<512, EPI, false, false, 2>, p, tc::SmemLayout<512, 2>::TOTAL, ma, mb, ep, s, c.pdl)
                     : tc_launch_kernel(tc::tc_gemm_kernel<512, EPI, true, true, 2>, p, tc::SmemLayout<512, 2>::TOTAL, ma, mb, ep, s, c.pdl);
        } else if (!mn) rc = (p.bn == 256) ? FSMG_GO(256, false) : FSMG_GO(128, false);
        else rc = (p.bn == 256) ? FSMG_GO(256, true) : FSMG_GO(128, true);
    }
#undef FSMG_GO
    if (rc) return rc;
    FSMG_LAUNCH_OK();
    return 0;
}

static inline int tc_make_maps(const TcContext& c, const GemmArgs& g, bool mn, int bn, int cl, CUtensorMap* ma, CUtensorMap* mb) {
    int rc;
    if (!mn) {
        if ((rc = make_map_f16(c, ma, g.A, (uint64_t)g.K, (uint64_t)g.M, (uint64_t)g.lda, tc::BK, tc::BM))) return rc;
        if ((rc = make_map_f16(c, mb, g.B, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.ldb, tc::BK, (uint32_t)((bn > 256 ? 256 : bn) / cl)))) return rc;   // each CTA of a pair fetches half of each MMA's B
    } else {
        if ((rc = make_map_f16(c, ma, g.A, (uint64_t)g.M, (uint64_t)g.K, (uint64_t)g.lda, 64, tc::BK))) return rc;
        if ((rc = make_map_f16(c, mb, g.B, (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.ldb, 64, tc::BK))) return rc;
    }
    return 0;
}

// fused softmax gradient: what the consumer GEMMs need to rebuild dlogits from the exponentials chunk (tc_gemm_kernel, "XF")
struct XfArgs {
    const float* cmaxT; int64_t ld_cmax; int n_c16;   // per (16-column chunk, chunk row) maximum written by the logits GEMM
    const float* lse; const int32_t* y; int64_t row0; // per token of the step: log-sum-exp, target id; first token of the chunk
    float* db;                                        // softmax_b gradient (dWs GEMM only)
};

// the transform variants exist for the 256 x 512 pair tiles only (long K loops, single-buffered accumulator: idle epilogue warps)
static inline bool tc_xf_supported(const TcContext& c, const GemmArgs& g) {
    if (!c.ready || !c.enabled) return false;
    if (!tc_operands_ok(g.A, g.lda) || !tc_operands_ok(g.B, g.ldb)) return false;
    const bool can_split = !g.c_half && !g.accumulate;
    const TcPlan p = tc_plan(c, g.M, g.N, g.K, can_split, 0, true);
    return p.bn == 512 && p.cl == 2;
}

static inline int tc_gemm(TcContext& c, const GemmArgs& g, bool a_mn, bool b_mn, cudaStream_t s, const XfArgs* xf = nullptr) {
    (void)b_mn;
    if (!c.ready) return set_error(-3, "tcgen05 context not initialised");
    const bool can_split = !g.c_half && !g.accumulate;   // split-K partials are combined with fp32 atomics
    TcPlan p = tc_plan(c, g.M, g.N, g.K, can_split, 0, true);
    tc::EpiParams ep;
    memset(&ep, 0, sizeof ep);
    ep.C = g.C; ep.ldc = g.ldc; ep.bias = g.bias; ep.alpha = g.alpha; ep.c_half = g.c_half; ep.accumulate = g.accumulate;
    ep.atomic = g.atomic;
    ep.vec_ok = ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0 && (g.ldc % 4) == 0) ? 1 : 0;
    if (xf) {
        if (p.bn != 512 || p.cl != 2) return set_error(-1, "operand-transform GEMM needs a 256 x 512 pair-tile plan (M=%d N=%d K=%d)", g.M, g.N, g.K);
        ep.cmaxT = const_cast<float*>(xf->cmaxT); ep.ld_cmax = xf->ld_cmax; ep.n_c16 = xf->n_c16;
        ep.xf_lse = xf->lse; ep.y = xf->y; ep.row0 = xf->row0; ep.xf_db = xf->db; ep.xf_diag = c.xf_diag;
    }
    if (p.sh.n_s > 1 && !g.atomic) {
        // plain store with split-K: zero the destination, then accumulate atomically
        FSMG_CUDA_OK(cudaMemset2DAsync(g.C, (size_t)g.ldc * 4, 0, (size_t)g.N * 4, (size_t)g.M, s));
        ep.atomic = 1;
    }
    CUtensorMap ma, mb;
    int rc = tc_make_maps(c, g, a_mn, p.bn, p.cl, &ma, &mb);
    if (rc) return rc;
    if (xf) {
        rc = !a_mn ? tc_launch_kernel(tc::tc_gemm_kernel<512, tc::EPI_STORE, false, false, 2, false, 1>, p, tc::SmemLayout<512, 2>::TOTAL, ma, mb, ep, s)
                   : tc_launch_kernel(tc::tc_gemm_kernel<512, tc::EPI_STORE, true, true, 2, false, 2>, p, tc::SmemLayout<512, 2>::TOTAL, ma, mb, ep, s);
        if (rc) return rc;
        FSMG_LAUNCH_OK();
        return 0;
    }
    return tc_launch<tc::EPI_STORE>(c, p, ma, mb, a_mn, ep, s);
}

// dX[r,:] = A[r,:] * B^T scattered into table[ids[r], :] (+= alpha * dX[r,:]) with the per-occurrence square norm: the embedding
// gradient of the hot path (IndexedSlices densification, reference lstm_baseline.py:83-85) without materialising dX
static inline int tc_gemm_scatter(TcContext& c, const GemmArgs& g, const int32_t* ids, float* table, int64_t ld_table, float* occ_sq,
                                  cudaStream_t s) {
    if (!c.ready) return set_error(-3, "tcgen05 context not initialised");
    TcPlan p = tc_plan(c, g.M, g.N, g.K, false);
    tc::EpiParams ep;
    memset(&ep, 0, sizeof ep);
    ep.C = table; ep.ldc = ld_table; ep.alpha = g.alpha; ep.y = ids; ep.row0 = 0; ep.tgt = occ_sq;
    ep.vec_ok = (table && (reinterpret_cast<uintptr_t>(table) & 15) == 0 && (ld_table % 4) == 0) ? 1 : 0;
    CUtensorMap ma, mb;
    int rc = tc_make_maps(c, g, false, p.bn, p.cl, &ma, &mb);
    if (rc) return rc;
    return tc_launch<tc::EPI_SCATTER>(c, p, ma, mb, false, ep, s);
}

// ---- projection forward: logits tile -> online LSE partials (+ fp16 logits when training) -------------
static inline bool tc_projection_supported(TcContext& c, int H, int V1) {
    (void)V1;
    return c.ready && c.enabled && (H % 8 == 0 || true);
}

static inline int tc_projection_gemm(TcContext& c, const __half* hc, int64_t ldh, const __half* WsT16, int64_t ldw, const float* sb,
                                     const int32_t* y, int64_t row0, int mc, int H, int V1, __half* logits16, int64_t ld16,
                                     int* n_part_out, cudaStream_t s, float* cmaxT = nullptr, int64_t ld_cmax = 0) {
    GemmArgs g;
    memset(&g, 0, sizeof g);
    g.M = mc; g.N = V1; g.K = H; g.A = hc; g.lda = ldh; g.B = WsT16; g.ldb = ldw;
    if (!tc_operands_ok(g.A, g.lda) || !tc_operands_ok(g.B, g.ldb)) return set_error(-1, "projection operands not TMA-aligned");
    const bool astat = c.astat && H <= tc::ASTAT_KB * tc::BK && V1 > 128 && mc > tc::BM;
    TcPlan p = tc_plan(c, mc, V1, H, false, astat ? 2 : c.cluster_lse);
    tc::EpiParams ep;
    memset(&ep, 0, sizeof ep);
    ep.bias = sb; ep.part = c.part; ep.n_tiles_total = 4 * p.sh.n_n; ep.y = y; ep.row0 = row0; ep.tgt = c.tgt;
    ep.logits16 = logits16; ep.ld16 = ld16;
    ep.exp_store = (logits16 && cmaxT) ? 1 : 0; ep.cmaxT = cmaxT; ep.ld_cmax = ld_cmax; ep.n_c16 = cdiv(V1, 16);
    ep.vec_ok = (logits16 && (reinterpret_cast<uintptr_t>(logits16) & 31) == 0 && (ld16 % 16) == 0) ? 1 : 0;   // 256-bit row-chunk stores
    *n_part_out = 4 * p.sh.n_n;
    CUtensorMap ma, mb;
    int rc = tc_make_maps(c, g, false, p.bn, p.cl, &ma, &mb);
    if (rc) return rc;
    if (astat && p.cl == 2 && p.bn == 256) {
        // A-stationary schedule: items = (M pair-block, 1/S of the N tiles); pick S for the fewest tile-times on the busiest pair
        const int slots = c.num_sms / 2;
        int best_s = 1; double best = 1e30;
        for (int sp = 1; sp <= 16 && sp <= p.sh.n_n; ++sp) {
            const int items = p.sh.n_mp * sp;
            const double cost = (double)cdiv(items, slots) * (cdiv(p.sh.n_n, sp) + 0.25);   // + A refill / pipeline restart per item
            if (cost < best - 1e-9) { best = cost; best_s = sp; }
        }
        p.sh.astat_s = best_s;
        p.sh.streamk = 0; p.sh.n_s = 1;
        const int items = p.sh.n_mp * best_s;
        p.grid = (items < slots ? items : slots) * 2;
        rc = tc_launch_kernel(tc::tc_gemm_kernel<256, tc::EPI_LSE, false, false, 2, true>, p, tc::SmemLayout<256, 2, true>::TOTAL, ma, mb, ep, s);
        if (rc) return rc;
        FSMG_LAUNCH_OK();
        return 0;
    }
    return tc_launch<tc::EPI_LSE>(c, p, ma, mb, false, ep, s);
}

// LSE combine (+ in-place softmax gradient and bias gradient when training) over the chunk
static inline int tc_projection_combine(TcContext& c, int n_part, int64_t row0, int mc, int N, int T, float* lse, float* nll_out,
                                        cudaStream_t s, bool reset_sched = false) {
    tc::lse_combine_kernel<<<cdiv(mc, 8), 256, 0, s>>>(c.part, n_part, c.tgt, row0, mc, N, T, lse, nll_out,
                                                       reset_sched ? c.strip_sched : nullptr);
    FSMG_LAUNCH_OK();
    return 0;
}

// mode 0: strip grid (param = rows per strip: 16 / 32 / 64 / 96 / 128);  mode 1: persistent strip loop (param = 8 / 16 / 32 rows);
// mode 2: streaming, one co-resident wave of `waves` CTAs per SM (param = rows per batch: 1 / 2 / 4)
static inline int tc_softmax_grad_launch(int num_sms, int mode, int param, int waves, const int32_t* y, int64_t row0, int mc, int V1,
                                         __half* logits16, int64_t ld16, const float* lse, float db_alpha, float* db, cudaStream_t s) {
#define FSMG_STRIP_GO(R)                                                                                                         \
    do {                                                                                                                         \
        dim3 grid(cdiv(ld16, 1024), cdiv(mc, R));                                                                                \
        tc::softmax_grad_strip_kernel<R><<<grid, 128, 0, s>>>(logits16, ld16, V1, lse, y, row0, mc, db_alpha, db);               \
    } while (0)
    if (waves < 1 || waves > 6) waves = 6;
    if (mode == 2) {
        const int n_bx = cdiv(ld16, 1024);
        int segs = (waves * num_sms) / n_bx;                  // one co-resident wave
        if (segs < 1) segs = 1;
        int seg_rows = cdiv(mc, segs);
        if (seg_rows < 8) seg_rows = 8;
        dim3 grid(n_bx, cdiv(mc, seg_rows));
        if (param == 1) tc::softmax_grad_stream_kernel<1><<<grid, 128, 0, s>>>(logits16, ld16, V1, lse, y, row0, mc, seg_rows, db_alpha, db);
        else if (param == 4) tc::softmax_grad_stream_kernel<4><<<grid, 128, 0, s>>>(logits16, ld16, V1, lse, y, row0, mc, seg_rows, db_alpha, db);
        else tc::softmax_grad_stream_kernel<2><<<grid, 128, 0, s>>>(logits16, ld16, V1, lse, y, row0, mc, seg_rows, db_alpha, db);
    } else if (mode == 1) {
        const int grid = waves * num_sms;
        if (param == 8) tc::softmax_grad_strip_loop_kernel<8><<<grid, 128, 0, s>>>(logits16, ld16, V1, lse, y, row0, mc, db_alpha, db);
        else if (param == 16) tc::softmax_grad_strip_loop_kernel<16><<<grid, 128, 0, s>>>(logits16, ld16, V1, lse, y, row0, mc, db_alpha, db);
        else tc::softmax_grad_strip_loop_kernel<32><<<grid, 128, 0, s>>>(logits16, ld16, V1, lse, y, row0, mc, db_alpha, db);
    } else if (param == 16) FSMG_STRIP_GO(16);
    else if (param == 64) FSMG_STRIP_GO(64);
    else if (param == 96) FSMG_STRIP_GO(96);
    else if (param == 128) FSMG_STRIP_GO(128);
    else FSMG_STRIP_GO(32);
#undef FSMG_STRIP_GO
    FSMG_LAUNCH_OK();
    return 0;
}

static inline int tc_projection_strip(TcContext& c, const int32_t* y, int64_t row0, int mc, int V1, __half* logits16, int64_t ld16,
                                      const float* lse, float db_alpha, float* db, cudaStream_t s) {
    return tc_softmax_grad_launch(c.num_sms, c.strip_mode, c.strip_param, c.strip_waves, y, row0, mc, V1, logits16, ld16, lse, db_alpha, db, s);
}

// background flavour: launched on a side stream while persistent GEMMs own the SMs (see softmax_grad_strip_bg_kernel)
static inline int tc_projection_strip_bg(TcContext& c, const int32_t* y, int64_t row0, int mc, int V1, __half* logits16, int64_t ld16,
                                         const float* lse, float db_alpha, float* db, cudaStream_t s) {
    // (c.strip_sched was zeroed by the chunk's lse_combine_kernel)
    // up to 6 CTAs fit an otherwise empty SM: with 6 x num_sms CTAs every SM sees at least one candidate whichever way the
    // scheduler packs them; the surplus exits on its first instruction (claim refused or no strips left)
    const int items = cdiv(ld16, 1024) * cdiv(mc, STRIP_ROWS);
    int grid = 6 * c.num_sms;
    if (grid > items) grid = items;
    tc::softmax_grad_strip_bg_kernel<STRIP_ROWS><<<grid, 128, 0, s>>>(logits16, ld16, V1, lse, y, row0, mc, db_alpha, db, c.strip_sched,
                                                                      c.strip_per_sm);
    FSMG_LAUNCH_OK();
    return 0;
}

static inline int tc_projection_post(TcContext& c, int n_part, const int32_t* y, int64_t row0, int mc, int N, int T, int V1,
                                     __half* logits16, int64_t ld16, float* lse, float* nll_out, float db_alpha, float* db,
                                     cudaStream_t s) {
    int rc = tc_projection_combine(c, n_part, row0, mc, N, T, lse, nll_out, s);
    if (rc || !logits16) return rc;
    return tc_projection_strip(c, y, row0, mc, V1, logits16, ld16, lse, db_alpha, db, s);
}

}  // namespace fsmg
