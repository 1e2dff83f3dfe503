// tc_lstm.cuh — persistent-RNN kernels (sm_100a): the T-step recurrence of BasicLSTMCell
// (reference src/models/lstm_baseline.py:44-55, [TF-lib] A.2/A.3) and its reverse-time backward, ONE
// launch per direction instead of T launches.
//
// Decomposition.  The N sequences are split into G independent groups; a group is served by
// C = H/U CTAs, CTA j owning hidden units [jU, (j+1)U) — all four gates of those units, so the cell
// update is CTA-local.  Its slice of the recurrent weights (fp16; 128 KB for U=32, H=512) is loaded
// ONCE by TMA and stays resident in shared memory for all T steps.  Per step:
//   producer warp : waits until the group's C CTAs have published the previous step (monotonic counter
//                   in L2, release/acquire), then streams the exchanged operand (h_{t-1} forward,
//                   dgates_{t+1} backward) through a TMA ring (SWIZZLE_128B);
//   MMA warp      : tcgen05.mma (M=128, N=4U forward / U backward, K=16), accumulators in TMEM;
//   epilogue warps: tcgen05.ld -> cell math in fp32 -> state (c forward, dc backward) kept in REGISTERS
//                   for the whole sequence -> writes the step's outputs -> publishes.
// Groups never synchronise with each other: no grid-wide barrier.
#pragma once
#include "tc_gemm.cuh"

namespace fsmg {
namespace tc {

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// 2*sigmoid(2x)-1: absolute error ~2e-7 (ex2.approx + rcp.approx), far below the fp16 rounding of h
__device__ __forceinline__ float tanh_fast(float x) { return __fdividef(2.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Release pattern after a CTA barrier (same as CUTLASS' GenericBarrier::red_release): ONE acq_rel fence at gpu scope —
// it is cumulative, so it also publishes what the other threads of the CTA wrote before the barrier — then a relaxed red.
// (__threadfence() + red.release compiled to MEMBAR.SC.GPU + CCTL.IVALL + a second MEMBAR: ~3 us per step on the trace.)
__device__ __forceinline__ void red_release_add(int* p, int v) {
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic-proxy global writes of other SMs (acquired through the counter) -> async-proxy (TMA) global reads: the global-only form is
// enough here and far cheaper than the all-state-space fence (which cost ~1.5 us per step on the in-kernel trace)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5, %6}], [%2], %3;" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// cta_group::2 flavour: the data lands in THIS CTA's smem, the complete_tx goes to the mbarrier at the same offset in CTA 0 (the
// MMA-issuing leader of the pair) — no software relay between the peer's TMA completion and the leader's MMA thread
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar_local) {
    asm volatile(
        "{\n"
        ".reg .b32 rb;\n"
        "mapa.shared::cluster.u32 rb, %2, 0;\n"
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [rb];\n"
        "}\n" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar_local)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns, registers -> TMEM
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
// NW (4, 8 or 16) consecutive columns from the head of a 16-word buffer
template <int NW>
__device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t (&w)[16]) {
    if constexpr (NW == 16) {
        tmem_st16(taddr, w);
    } else if constexpr (NW == 8) {
        uint32_t w8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) w8[e] = w[e];
        tmem_st8(taddr, w8);
    } else {
        static_assert(NW == 4, "4, 8 or 16 columns");
        uint32_t w4[4] = {w[0], w[1], w[2], w[3]};
        tmem_st4(taddr, w4);
    }
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#define FSMG_TR(step, slot)                                                                      \
    do {                                                                                         \
        if (p.trace && blockIdx.x == 0 && (step) >= 8 && (step) < 16) p.trace[((step) - 8) * 16 + (slot)] = clock64(); \
    } while (0)

struct LstmParams {
    // forward
    const __half* pre;     // [T*N, G4p] fp16: x_t * Wx + b (hoisted input contraction; fp16 storage changes the NLL error by < 1e-5)
    const int32_t* pre_ids; // null: row r of `pre` belongs to token r.  Else `pre` is a per-WORD table [V', G4p] (embedding * Wx + b
                            // for every word, one small GEMM per step) and token r reads row pre_ids[r] (time-major input ids)
    // backward
    const float* dh_out;   // [T*N, H] fp32: dL/dh_t from above (projection or upper layer), unscaled
    __half* dgates;        // [T*N, G4p] fp16 out (also the exchanged operand)
    // shared
    __half* gates;         // [T*N, G4p] fp16 post-activation i|j|f|o (stash: written fwd, read bwd)
    float* c;              // [T*N, H]
    __half* hs;            // [T*N, Hp]   (fwd out)
    int* counters;         // [G] monotonic counters (zeroed by the host before the launch)
    int N, T, H, Hp, G4p;
    int ctas_per_group, rows_per_group, box_rows;   // C, M_g, TMA box rows (= M_g rounded to 8)
    long long* trace;          // debug: clock64 timeline of CTA 0 for steps [8, 16) (FSMG_TRACE=1), else null
    int stage_bytes, stages;   // TMA ring geometry (host-computed: stages sized to the rows actually exchanged)
    int row_offset;        // first sequence handled by this launch (batch slicing when N is large)
    int n_rows;            // sequences handled by this launch
    int rotate;            // walk the K chunks in an order rotated per loader (FSMG_LSTM_ROT bit 0: backward, bit 1: forward)
    int ks;                // pair + split backward: 64-column K chunks (TMA boxes) per ring stage = per full/empty barrier round trip
    int box_pitch;         // bytes between the boxes of one stage (box rows x 128 B rounded to the 1024-B swizzle atom)
    int pub_cta;           // forward split kernel: one release per (CTA, half) behind a named barrier instead of one per warp
    int nh;                // forward split kernel: halves in use (2; 1 for groups of a few rows, where a second half only doubles the MMA count)
};

// first element of the hoisted pre-activation row of token (t, row): direct, or through the per-word table
__device__ __forceinline__ const __half* lstm_pre_row(const LstmParams& p, int t, int row) {
    const int64_t r = (int64_t)t * p.N + row;
    return p.pre + (p.pre_ids ? (int64_t)__ldg(p.pre_ids + r) : r) * p.G4p;
}
// table rows are re-read by many tokens (keep them in L2); per-token rows are read exactly once (streaming)
__device__ __forceinline__ uint4 lstm_pre_load(const LstmParams& p, const __half* ptr) {
    return p.pre_ids ? __ldg(reinterpret_cast<const uint4*>(ptr)) : __ldcs(reinterpret_cast<const uint4*>(ptr));
}

__host__ __device__ constexpr int lstm_threads(int mt) { return 64 + 128 * mt; }   // producer warp, MMA warp, 4 epilogue warps per row tile
constexpr int LSTM_MAX_DYN = 227 * 1024;

// smem: [resident weights: KCW chunks x CHUNK_W bytes] [ring: STAGES x (MT*128 rows x 128 B)] [barriers]
__host__ __device__ constexpr int lstm_stage_bytes(int mt) { return mt * 128 * 128; }
__host__ __device__ constexpr int lstm_num_stages(int w_bytes, int mt) {
    int s = (LSTM_MAX_DYN - 1024 - 256 - w_bytes) / lstm_stage_bytes(mt);
    return s > 8 ? 8 : s;
}
__host__ __device__ constexpr int lstm_smem_total(int w_bytes, int mt) {
    return 1024 + 256 + w_bytes + lstm_num_stages(w_bytes, mt) * lstm_stage_bytes(mt);
}
__host__ __device__ constexpr uint32_t tmem_cols_pow2(int n) { return n <= 32 ? 32u : n <= 64 ? 64u : n <= 128 ? 128u : n <= 256 ? 256u : 512u; }

// =====================================================================================================
// forward
// =====================================================================================================
// CLS > 1: the CLS CTAs of a thread-block cluster serve the same group; each fetches every CLS-th K chunk of the
// exchanged operand and TMA-multicasts it to all of them (L2 -> SMEM traffic / CLS).
template <int U, int MT, int CLS, bool PAIR>
__global__ void __launch_bounds__(lstm_threads(PAIR ? 2 : MT), 1)
lstm_fwd_persistent_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_h, LstmParams p) {
    constexpr int NCOL = 4 * U;                 // accumulator columns per row tile (UMMA N)
    constexpr int CHUNK_W = NCOL * 128;         // bytes of the weight slice per 64-wide K chunk
    const int KC = p.H / 64;
    const int W_BYTES = KC * CHUNK_W;
    const int STAGES = p.stages;
    const int STAGE_BYTES = p.stage_bytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
    uint8_t* sW = smem;
    uint8_t* sA = sW + W_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LSTM_MAX_DYN - 1024 - 256);   // fixed slot at the end of the carve-out
    uint64_t* full_bar = bars;                  // [8]
    uint64_t* empty_bar = bars + 8;             // [8]
    uint64_t* w_bar = bars + 16;
    uint64_t* tmem_full = bars + 17;
    uint64_t* pre_ready = bars + 18;            // [2] one per TMEM buffer: "pre-activations of the step that uses this buffer are staged".
                                                // Two barriers (not one) so the epilogue can never run two phases ahead of the MMA
                                                // thread (e.g. while it still waits for the one-time weight load).
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    uint64_t* peer_full = bars + 22;            // [8] PAIR, leader only

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x / p.ctas_per_group, j = blockIdx.x % p.ctas_per_group;
    static_assert(!PAIR || (MT == 1 && CLS == 1), "PAIR mode: one row tile per CTA, the cluster is the pair");
    constexpr int NQ = PAIR ? 2 : MT;           // epilogue warp quartets: row tiles (plain) or unit slices (PAIR)
    const int prank = PAIR ? (int)cluster_ctarank() : 0;
    const int group_row0 = p.row_offset + g * p.rows_per_group;                      // first sequence of the group
    const int group_rows = min(p.rows_per_group, p.row_offset + p.n_rows - group_row0);
    const int row_base = PAIR ? group_row0 + prank * p.box_rows : group_row0;          // first row this CTA streams / owns
    const int rows = PAIR ? max(0, min(p.box_rows, group_rows - prank * p.box_rows)) : group_rows;
    int* counter = p.counters + g;
    const int crank = (CLS > 1) ? (int)cluster_ctarank() : 0;
    constexpr uint16_t CMASK = (uint16_t)((1u << CLS) - 1u);
    constexpr int BUF_COLS = NQ * NCOL;               // one accumulator buffer; two buffers alternate per step
    constexpr uint32_t TMEM_COLS = tmem_cols_pow2(2 * BUF_COLS);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_h);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], CLS); }
        mbar_init(w_bar, 1);
        mbar_init(tmem_full, 1);
        mbar_init(&pre_ready[0], PAIR ? 16 : 4 * MT);   // PAIR: the leader's barrier collects both CTAs' 8 epilogue warps
        mbar_init(&pre_ready[1], PAIR ? 16 : 4 * MT);
        if (PAIR) for (int i = 0; i < STAGES; ++i) mbar_init(&peer_full[i], 1);
        fence_barrier_init();
    }
    if (warp == 1) { if (PAIR) tmem_alloc_2sm(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS); }
    tc_fence_before();
    __syncthreads();
    if (CLS > 1 || PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // resident weights: per K chunk the 4 gate blocks of U rows -> smem rows [i(U) | j(U) | f(U) | o(U)]
            mbar_expect_tx(w_bar, (uint32_t)W_BYTES);
            for (int kc = 0; kc < KC; ++kc)
                for (int q = 0; q < 4; ++q)
                    tma_load_2d(sW + kc * CHUNK_W + q * U * 128, &map_w, kc * 64, q * p.H + j * U, w_bar);
            int stage = 0; uint32_t phase = 0;
            const uint32_t box_bytes = (uint32_t)p.box_rows * 128u;
            for (int t = 1; t < p.T; ++t) {
                const int need = p.ctas_per_group * t;      // all C CTAs of the group have published h_{t-1}
                while (ld_acquire(counter) < need) { }
                FSMG_TR(t, 0);
                fence_proxy_async_all();                    // generic-proxy writes -> async-proxy (TMA) reads
                for (int kc = 0; kc < KC; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (PAIR) {   // the leader's barrier collects both halves: it expects 2 x box bytes, the peer only issues its load
                        if (prank == 0) mbar_expect_tx(&full_bar[stage], 2 * box_bytes);
                        tma_load_3d_2sm(sA + stage * STAGE_BYTES, &map_h, kc * 64, row_base, t - 1, &full_bar[stage]);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_expect_tx(&full_bar[stage], box_bytes);
                    if (CLS == 1) tma_load_3d(sA + stage * STAGE_BYTES, &map_h, kc * 64, row_base, t - 1, &full_bar[stage]);
                    else if (kc % CLS == crank) tma_load_3d_mc(sA + stage * STAGE_BYTES, &map_h, kc * 64, row_base, t - 1, &full_bar[stage], CMASK);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                FSMG_TR(t, 1);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && PAIR && prank == 1) {
            // peer CTA: no MMA issue.  Tell the leader once that this CTA's resident weight slice (half of B) has landed.
            mbar_wait(w_bar, 0);
            mbar_arrive_remote_release(&peer_full[0], 0);
        } else if (lane == 0) {
            constexpr uint32_t idesc = PAIR ? make_idesc_m(256, 2 * NCOL) : make_idesc(NCOL, false, false);
            mbar_wait(w_bar, 0);
            if (PAIR) mbar_wait(&peer_full[0], 0);     // the peer's weight slice is resident too
            tc_fence_after();
            int stage = 0; uint32_t phase = 0;
            uint32_t pr_phase[2] = {0, 0};
            for (int t = 1; t < p.T; ++t) {
                // buffer t&1 already holds pre[t] (x_t*Wx + b), staged by the epilogue warps during step t-1:
                // every MMA accumulates, so the epilogue reads finished gate pre-activations straight from TMEM
                mbar_wait(&pre_ready[t & 1], pr_phase[t & 1]);
                pr_phase[t & 1] ^= 1;
                tc_fence_after();
                const uint32_t d_buf = tmem_base + (uint32_t)((t & 1) * BUF_COLS);
                for (int kc = 0; kc < KC; ++kc) {
                    mbar_wait(&full_bar[stage], phase);
                    if (kc == 0) FSMG_TR(t, 2);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(sA + stage * STAGE_BYTES);
                    const uint32_t sb = smem_u32(sW + kc * CHUNK_W);
                    if (PAIR) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t a_desc = make_smem_desc(sa + k * 32, 16, 1024);
                            const uint64_t b_desc = make_smem_desc(sb + k * 32, 16, 1024);
                            umma_f16_2sm(d_buf, a_desc, b_desc, idesc, 1u);
                        }
                        umma_commit_2sm_mc(&empty_bar[stage], (uint16_t)0x3);
                    } else {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t a_desc = make_smem_desc(sa + mt * 128 * 128 + k * 32, 16, 1024);
                                const uint64_t b_desc = make_smem_desc(sb + k * 32, 16, 1024);
                                umma_f16(d_buf + mt * NCOL, a_desc, b_desc, idesc, 1u);
                            }
                        }
                        if (CLS > 1) umma_commit_mc(&empty_bar[stage], CMASK); else umma_commit(&empty_bar[stage]);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (PAIR) umma_commit_2sm_mc(tmem_full, (uint16_t)0x3); else umma_commit(tmem_full);
                FSMG_TR(t, 3);
            }
        }
    } else {
        // ===== epilogue: thread <-> (row tile mt, TMEM lane); the cell state c stays in registers for all T steps
        const int quad = warp & 3;
        const int q_idx = (warp - 2) >> 2;       // quartet: row tile (plain) or unit slice of the pair (PAIR)
        const int mt = PAIR ? 0 : q_idx;
        const int ju = PAIR ? (j & ~1) + q_idx : j;
        float c_state[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c_state[u] = 0.0f;
        uint32_t tf_phase = 0;
        // stage pre[t] of this thread's rows into TMEM buffer (t & 1): 4 gates x U columns per row tile
        auto stage_pre = [&](int t) {
            {
                const int lrow = mt * 128 + quad * 32 + lane;
                const bool ok = lrow < rows;
                const __half* pre = ok ? lstm_pre_row(p, t, row_base + lrow) + ju * U : p.pre;
                const uint32_t t_row = tmem_base + (uint32_t)((t & 1) * BUF_COLS) + q_idx * NCOL + ((uint32_t)(quad * 32) << 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t rr[U];
#pragma unroll
                    for (int e = 0; e < U; e += 8) {
                        uint4 raw = ok ? lstm_pre_load(p, pre + q * p.H + e) : make_uint4(0, 0, 0, 0);
                        const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            const float2 f = __half22float2(h2[w]);
                            rr[e + 2 * w] = __float_as_uint(f.x);
                            rr[e + 2 * w + 1] = __float_as_uint(f.y);
                        }
                    }
                    if constexpr (U == 32) tmem_st32(t_row + q * U, rr); else tmem_st16(t_row + q * U, rr);
                }
            }
            tmem_st_wait();
        };
        stage_pre(0);
        for (int t = 0; t < p.T; ++t) {
            if (t + 1 < p.T) {
                // while the group exchanges h_t and the tensor core works on step t, stage step t+1's pre-activations
                stage_pre(t + 1);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR && prank == 1) mbar_arrive_remote(&pre_ready[(t + 1) & 1], 0); else mbar_arrive(&pre_ready[(t + 1) & 1]); }
                if (warp == 2 && lane == 0) FSMG_TR(t, 4);
                if (t + 2 < p.T) {   // and pull step t+2's lines towards L2
                    {
                        const int lrow = mt * 128 + quad * 32 + lane;
                        if (lrow < rows) {
                            const __half* nxt = lstm_pre_row(p, t + 2, row_base + lrow) + ju * U;
#pragma unroll
                            for (int q = 0; q < 4; ++q) prefetch_l2(nxt + q * p.H);
                        }
                    }
                }
            }
            if (t > 0) {
                mbar_wait(tmem_full, tf_phase);
                tf_phase ^= 1;
                tc_fence_after();
            }
            if (warp == 2 && lane == 0) FSMG_TR(t, 5);
            const int lrow = mt * 128 + quad * 32 + lane;         // row inside the group
            const bool ok = lrow < rows;
            const int64_t r = (int64_t)t * p.N + row_base + lrow; // time-major token row
            // BPTT stash of this step (gates fp16, c fp32): kept in registers and written AFTER h has been published,
            // so the release fence only has to drain the 64-byte h stores the rest of the group is waiting for
            uint4 g_stash[U / 8][4];
            float4 c_stash[U / 8][2];
            {
                const uint32_t t_row = tmem_base + (uint32_t)((t & 1) * BUF_COLS) + q_idx * NCOL + ((uint32_t)(quad * 32) << 16);
#pragma unroll
                for (int u0 = 0; u0 < U; u0 += 8) {
                    float acc[4][8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t rr[8];
                        tmem_ld8(t_row + q * U + u0, rr);
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[q][e] = __uint_as_float(rr[e]);
                    }
                    tmem_ld_wait();
                    __align__(16) __half hq[4][8];
                    __align__(16) __half hh[8];
                    float cn[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        // 5 ex2 + 3 rcp per unit instead of 5 + 5 (the epilogue is MUFU-bound): sigmoid(a) and tanh(b) share one
                        // reciprocal of (1+e^-a)(1+e^-2b).  Exponents are clamped at 2^57 so the product cannot overflow.
                        const float ea = fast_ex2(fminf(-1.4426950408889634f * acc[0][e], 57.0f));
                        const float eb = fast_ex2(fminf(-2.8853900817779268f * acc[1][e], 57.0f));
                        const float ef = fast_ex2(fminf(-1.4426950408889634f * (acc[2][e] + 1.0f), 57.0f));   // forget_bias = 1 (A.2)
                        const float eo = fast_ex2(fminf(-1.4426950408889634f * acc[3][e], 57.0f));
                        const float r1 = __fdividef(1.0f, (1.0f + ea) * (1.0f + eb));
                        const float i_ = r1 * (1.0f + eb);
                        const float j_ = (1.0f - eb) * (r1 * (1.0f + ea));
                        const float f_ = __fdividef(1.0f, 1.0f + ef);
                        const float cv = c_state[u0 + e] * f_ + i_ * j_;
                        c_state[u0 + e] = cv;
                        cn[e] = cv;
                        const float ec = fast_ex2(fminf(-2.8853900817779268f * cv, 57.0f));
                        const float r2 = __fdividef(1.0f, (1.0f + ec) * (1.0f + eo));
                        const float o_ = r2 * (1.0f + ec);
                        const float tc_ = (1.0f - ec) * (r2 * (1.0f + eo));
                        hq[0][e] = __float2half_rn(i_); hq[1][e] = __float2half_rn(j_);
                        hq[2][e] = __float2half_rn(f_); hq[3][e] = __float2half_rn(o_);
                        hh[e] = __float2half_rn(tc_ * o_);
                    }
                    if (ok) *reinterpret_cast<uint4*>(p.hs + r * p.Hp + ju * U + u0) = *reinterpret_cast<uint4*>(hh);
#pragma unroll
                    for (int q = 0; q < 4; ++q) g_stash[u0 / 8][q] = *reinterpret_cast<uint4*>(hq[q]);
                    c_stash[u0 / 8][0] = make_float4(cn[0], cn[1], cn[2], cn[3]);
                    c_stash[u0 / 8][1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
                }
            }
            // publish: CTA barrier, then ONE gpu-scope release (cumulative over the CTA's h stores)
            tc_fence_before();
            if (warp == 2 && lane == 0) FSMG_TR(t, 6);
            named_bar_sync(1, 128 * NQ);
            if (warp == 2 && lane == 0) { FSMG_TR(t, 7); red_release_add(counter, 1); FSMG_TR(t, 8); }
            if (ok) {
                __half* gout = p.gates + r * p.G4p + ju * U;
                float* cdst = p.c + r * p.H + ju * U;
#pragma unroll
                for (int u0 = 0; u0 < U; u0 += 8) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(gout + q * p.H + u0) = g_stash[u0 / 8][q];
                    *reinterpret_cast<float4*>(cdst + u0) = c_stash[u0 / 8][0];
                    *reinterpret_cast<float4*>(cdst + u0 + 4) = c_stash[u0 / 8][1];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CLS > 1 || PAIR) cluster_sync_all();
    if (warp == 1) { if (PAIR) tmem_dealloc_2sm(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS); }
}

// =====================================================================================================
// forward, split schedule: the group's rows are cut into two independent HALVES (own progress counter, own TMEM double
// buffer, own epilogue warps); the producer and the MMA thread serve them alternately.  While one half waits for its
// exchange (publish -> counter -> TMA round trip through L2, ~4 us) the other half's loads, MMAs and cell epilogue
// run, so the per-step latency chain of one half hides behind the work of the other.  The epilogue additionally splits
// a row's U units over two threads (16 warps: half x unit slice x TMEM quadrant), halving its dependent MUFU chain.
// =====================================================================================================
constexpr int LSTM_SPLIT_THREADS = 64 + 512;

template <int U>
__global__ void __launch_bounds__(LSTM_SPLIT_THREADS, 1)
lstm_fwd_split_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_h, LstmParams p) {
    constexpr int NCOL = 4 * U;                 // accumulator columns of one half's buffer (UMMA N)
    constexpr int CHUNK_W = NCOL * 128;
    constexpr int US = U / 2;                   // units per epilogue thread
    static_assert(4 * NCOL <= 512, "two halves x two buffers must fit the 512 TMEM columns");
    const int KC = p.H / 64;
    const int W_BYTES = KC * CHUNK_W;
    const int STAGES = p.stages;
    const int STAGE_BYTES = p.stage_bytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;
    uint8_t* sA = sW + W_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LSTM_MAX_DYN - 1024 - 256);
    uint64_t* full_bar = bars;                  // [8]
    uint64_t* empty_bar = bars + 8;             // [8]
    uint64_t* w_bar = bars + 16;
    uint64_t* tmem_full = bars + 17;            // [2] per half
    uint64_t* pre_ready = bars + 19;            // [2][2] per half, per TMEM buffer parity
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x / p.ctas_per_group, j = blockIdx.x % p.ctas_per_group;
    const int group_row0 = p.row_offset + g * p.rows_per_group;
    const int group_rows = min(p.rows_per_group, p.row_offset + p.n_rows - group_row0);
    const int hr = p.box_rows;                  // rows per half (multiple of 8, <= 128)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_h);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(w_bar, 1);
        for (int i = 0; i < 2; ++i) mbar_init(&tmem_full[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&pre_ready[i], 8);   // 2 unit slices x 4 quadrant warps
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(w_bar, (uint32_t)W_BYTES);
            for (int kc = 0; kc < KC; ++kc)
                for (int q = 0; q < 4; ++q)
                    tma_load_2d(sW + kc * CHUNK_W + q * U * 128, &map_w, kc * 64, q * p.H + j * U, w_bar);
            int stage = 0; uint32_t phase = 0;
            const uint32_t box_bytes = (uint32_t)hr * 128u;
            for (int t = 1; t < p.T; ++t) {
                const int need = p.ctas_per_group * (p.pub_cta ? 1 : 8) * t;   // publishes of h_{t-1} per half: one per CTA, or one per epilogue warp
                for (int hs_ = 0; hs_ < p.nh; ++hs_) {
                    int* counter = p.counters + 2 * g + hs_;
                    while (ld_acquire(counter) < need) { }
                    if (hs_ == 0) FSMG_TR(t, 0);
                    fence_proxy_async_all();
                    // K chunks in an order rotated by the CTA's slice index: the C CTAs of a group read the same rows, and in
                    // lock step they would all hit the same L2 lines at the same moment.  p.ks chunks (TMA boxes) per ring stage:
                    // one full/empty barrier round trip of the MMA thread per p.ks x 4 MMAs
                    int kcr = (p.rotate ? j : 0) % KC;
                    for (int kb = 0; kb < KC; kb += p.ks) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_expect_tx(&full_bar[stage], p.ks * box_bytes);
                        for (int i = 0; i < p.ks; ++i) {
                            tma_load_3d(sA + stage * STAGE_BYTES + i * p.box_pitch, &map_h, kcr * 64, group_row0 + hs_ * hr, t - 1, &full_bar[stage]);
                            if (++kcr == KC) kcr = 0;
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (hs_ == 0) FSMG_TR(t, 1);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(NCOL, false, false);
            mbar_wait(w_bar, 0);
            tc_fence_after();
            int stage = 0; uint32_t phase = 0;
            uint32_t pr_phase[4] = {0, 0, 0, 0};
            const uint64_t b_desc0 = make_smem_desc(smem_u32(sW), 16, 1024);
            const uint64_t a_step = (uint64_t)(p.box_pitch >> 4);
            for (int t = 1; t < p.T; ++t) {
                for (int hs_ = 0; hs_ < p.nh; ++hs_) {
                    const int b = hs_ * 2 + (t & 1);
                    mbar_wait(&pre_ready[b], pr_phase[b]);     // pre[t] of this half is staged in its buffer: every MMA accumulates
                    pr_phase[b] ^= 1;
                    tc_fence_after();
                    const uint32_t d_buf = tmem_base + (uint32_t)(b * NCOL);
                    int kcr = (p.rotate ? j : 0) % KC;                    // same rotation as the producer
                    for (int kb = 0; kb < KC; kb += p.ks) {
                        mbar_wait(&full_bar[stage], phase);
                        if (kb == 0 && hs_ == 0) FSMG_TR(t, 2);
                        tc_fence_after();
                        // descriptors built once per stage and advanced with 64-bit adds (start-address field = bytes >> 4)
                        uint64_t a_desc = make_smem_desc(smem_u32(sA + stage * STAGE_BYTES), 16, 1024);
                        for (int i = 0; i < p.ks; ++i) {
                            const uint64_t b_desc = b_desc0 + (uint64_t)(kcr * (CHUNK_W >> 4));
                            umma_f16(d_buf, a_desc, b_desc, idesc, 1u);
                            umma_f16(d_buf, a_desc + 2, b_desc + 2, idesc, 1u);
                            umma_f16(d_buf, a_desc + 4, b_desc + 4, idesc, 1u);
                            umma_f16(d_buf, a_desc + 6, b_desc + 6, idesc, 1u);
                            a_desc += a_step;
                            if (++kcr == KC) kcr = 0;
                        }
                        umma_commit(&empty_bar[stage]);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tmem_full[hs_]);
                    if (hs_ == 0) FSMG_TR(t, 3);
                }
            }
        }
    } else {
        // ===== epilogue: 16 warps = half (2) x unit slice (2) x TMEM lane quadrant (4); thread <-> one row, US units
        const int quad = warp & 3;
        const int idx = (warp - 2) >> 2;
        const int hs_ = idx >> 1, us = idx & 1;
        const int t_end = hs_ < p.nh ? p.T : 0;               // warps of an unused half have nothing to do (no barrier waits on them)
        const int row_base = group_row0 + hs_ * hr;
        const int rows = max(0, min(hr, group_rows - hs_ * hr));
        const int lrow = quad * 32 + lane;
        const bool ok = lrow < rows;
        const int ucol = j * U + us * US;                     // first hidden unit of this thread
        int* counter = p.counters + 2 * g + hs_;
        const bool tracer = (hs_ == 0 && us == 0 && quad == 0 && lane == 0);
        float c_state[US];
#pragma unroll
        for (int u = 0; u < US; ++u) c_state[u] = 0.0f;
        uint32_t tf_phase = 0;
        auto stage_pre = [&](int t) {
            const __half* pre = ok ? lstm_pre_row(p, t, row_base + lrow) + ucol : p.pre;
            const uint32_t t_row = tmem_base + (uint32_t)((hs_ * 2 + (t & 1)) * NCOL) + us * US + ((uint32_t)(quad * 32) << 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t rr[US];
#pragma unroll
                for (int e = 0; e < US; e += 8) {
                    uint4 raw = ok ? lstm_pre_load(p, pre + q * p.H + e) : make_uint4(0, 0, 0, 0);
                    const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const float2 f = __half22float2(h2[w]);
                        rr[e + 2 * w] = __float_as_uint(f.x);
                        rr[e + 2 * w + 1] = __float_as_uint(f.y);
                    }
                }
                if constexpr (US == 16) tmem_st16(t_row + q * U, rr); else tmem_st8(t_row + q * U, rr);
            }
            tmem_st_wait();
        };
        if (t_end > 0) stage_pre(0);
        for (int t = 0; t < t_end; ++t) {
            if (t + 1 < p.T) {
                stage_pre(t + 1);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&pre_ready[hs_ * 2 + ((t + 1) & 1)]);
                if (tracer) FSMG_TR(t, 4);
                if (t + 2 < p.T && ok) {
                    const __half* nxt = lstm_pre_row(p, t + 2, row_base + lrow) + ucol;
#pragma unroll
                    for (int q = 0; q < 4; ++q) prefetch_l2(nxt + q * p.H);
                }
            }
            if (t > 0) {
                mbar_wait(&tmem_full[hs_], tf_phase);
                tf_phase ^= 1;
                tc_fence_after();
            }
            if (tracer) FSMG_TR(t, 5);
            const int64_t r = (int64_t)t * p.N + row_base + lrow;
            uint4 g_stash[US / 8][4];
            float4 c_stash[US / 8][2];
            {
                const uint32_t t_row = tmem_base + (uint32_t)((hs_ * 2 + (t & 1)) * NCOL) + us * US + ((uint32_t)(quad * 32) << 16);
#pragma unroll
                for (int u0 = 0; u0 < US; u0 += 8) {
                    float acc[4][8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t rr[8];
                        tmem_ld8(t_row + q * U + u0, rr);
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[q][e] = __uint_as_float(rr[e]);
                    }
                    tmem_ld_wait();
                    __align__(16) __half hq[4][8];
                    __align__(16) __half hh[8];
                    float cn[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        // 5 ex2 + 3 rcp per unit: sigmoid(a) and tanh(b) share one reciprocal of (1+e^-a)(1+e^-2b); exponents clamped at 2^57
                        const float ea = fast_ex2(fminf(-1.4426950408889634f * acc[0][e], 57.0f));
                        const float eb = fast_ex2(fminf(-2.8853900817779268f * acc[1][e], 57.0f));
                        const float ef = fast_ex2(fminf(-1.4426950408889634f * (acc[2][e] + 1.0f), 57.0f));   // forget_bias = 1 (A.2)
                        const float eo = fast_ex2(fminf(-1.4426950408889634f * acc[3][e], 57.0f));
                        const float r1 = __fdividef(1.0f, (1.0f + ea) * (1.0f + eb));
                        const float i_ = r1 * (1.0f + eb);
                        const float j_ = (1.0f - eb) * (r1 * (1.0f + ea));
                        const float f_ = __fdividef(1.0f, 1.0f + ef);
                        const float cv = c_state[u0 + e] * f_ + i_ * j_;
                        c_state[u0 + e] = cv;
                        cn[e] = cv;
                        const float ec = fast_ex2(fminf(-2.8853900817779268f * cv, 57.0f));
                        const float r2 = __fdividef(1.0f, (1.0f + ec) * (1.0f + eo));
                        const float o_ = r2 * (1.0f + ec);
                        const float tc_ = (1.0f - ec) * (r2 * (1.0f + eo));
                        hq[0][e] = __float2half_rn(i_); hq[1][e] = __float2half_rn(j_);
                        hq[2][e] = __float2half_rn(f_); hq[3][e] = __float2half_rn(o_);
                        hh[e] = __float2half_rn(tc_ * o_);
                    }
                    if (ok) *reinterpret_cast<uint4*>(p.hs + r * p.Hp + ucol + u0) = *reinterpret_cast<uint4*>(hh);
#pragma unroll
                    for (int q = 0; q < 4; ++q) g_stash[u0 / 8][q] = *reinterpret_cast<uint4*>(hq[q]);
                    c_stash[u0 / 8][0] = make_float4(cn[0], cn[1], cn[2], cn[3]);
                    c_stash[u0 / 8][1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
                }
            }
            // publish per WARP (the counter counts warps): no CTA-wide barrier, and the gpu-scope release of lane 0 only has to
            // drain this warp's own h stores (made visible to it by the warp barrier)
            tc_fence_before();
            if (tracer) FSMG_TR(t, 6);
            if (p.pub_cta) {     // barrier among the half's 8 warps, then one cumulative gpu-scope release
                named_bar_sync(1 + hs_, 256);
                if (us == 0 && quad == 0 && lane == 0) { if (tracer) FSMG_TR(t, 7); red_release_add(counter, 1); if (tracer) FSMG_TR(t, 8); }
            } else {
                __syncwarp();
                if (lane == 0) { if (tracer) FSMG_TR(t, 7); red_release_add(counter, 1); if (tracer) FSMG_TR(t, 8); }
            }
            if (ok) {
                __half* gout = p.gates + r * p.G4p + ucol;
                float* cdst = p.c + r * p.H + ucol;
#pragma unroll
                for (int u0 = 0; u0 < US; u0 += 8) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(gout + q * p.H + u0) = g_stash[u0 / 8][q];
                    *reinterpret_cast<float4*>(cdst + u0) = c_stash[u0 / 8][0];
                    *reinterpret_cast<float4*>(cdst + u0 + 4) = c_stash[u0 / 8][1];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================
// backward (reverse time).  Resident: rows [jU, (jU+U)) of kernel[in:, :] (each 4H long, K-major);
// exchanged operand: dgates_{t+1} [rows, 4H];  accumulator: dh_rec[rows, U].
//   dh = dh_out[t] + dh_rec ; do = dh*tanh(c) ; dc = dh*o*(1-tanh(c)^2) + dc_next
//   dgi = dc*j*i(1-i) ; dgj = dc*i*(1-j^2) ; dgf = dc*c_prev*f(1-f) ; dgo = do*o(1-o) ; dc_next' = dc*f
// =====================================================================================================
// PAIR: the two CTAs of a cluster form one cta_group::2 MMA of 256 rows: each CTA streams only ITS half of the group's rows of
// the exchanged operand (the measured bottleneck is per-SM operand ingest) and contributes its own resident weight slice as half of
// B; each CTA's epilogue then owns (its rows) x (both CTAs' units), one warp quartet per unit slice.
template <int U, int MT, int CLS, bool PAIR>
__global__ void __launch_bounds__(lstm_threads(PAIR ? 2 : MT), 1)
lstm_bwd_persistent_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_dg, LstmParams p) {
    constexpr int NCOL = U;
    constexpr int CHUNK_W = U * 128;
    const int KC = (4 * p.H) / 64;
    const int W_BYTES = KC * CHUNK_W;
    const int STAGES = p.stages;
    const int STAGE_BYTES = p.stage_bytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
    uint8_t* sW = smem;
    uint8_t* sA = sW + W_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LSTM_MAX_DYN - 1024 - 256);   // fixed slot at the end of the carve-out
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + 8;
    uint64_t* w_bar = bars + 16;
    uint64_t* tmem_full = bars + 17;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
    uint64_t* peer_full = bars + 20;            // [8] PAIR, leader only: the peer CTA's stage holds its half of the rows

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x / p.ctas_per_group, j = blockIdx.x % p.ctas_per_group;
    static_assert(!PAIR || (MT == 1 && CLS == 1), "PAIR mode: one row tile per CTA, the cluster is the pair");
    constexpr int NQ = PAIR ? 2 : MT;           // epilogue warp quartets: row tiles (plain) or unit slices (PAIR)
    const int prank = PAIR ? (int)cluster_ctarank() : 0;
    const int group_row0 = p.row_offset + g * p.rows_per_group;
    const int group_rows = min(p.rows_per_group, p.row_offset + p.n_rows - group_row0);
    const int row_base = PAIR ? group_row0 + prank * p.box_rows : group_row0;          // first row this CTA streams / owns
    const int rows = PAIR ? max(0, min(p.box_rows, group_rows - prank * p.box_rows)) : group_rows;
    int* counter = p.counters + g;
    const int crank = (CLS > 1) ? (int)cluster_ctarank() : 0;
    constexpr uint16_t CMASK = (uint16_t)((1u << CLS) - 1u);
    // K chunks are walked in an order rotated per loader (pair / cluster / CTA): the CTAs of a group read the same rows, and in lock
    // step they would all hit the same L2 lines at the same moment (measured: the load phase ran at half of the L2 throughput cap)
    const int n_loaders = p.ctas_per_group / (PAIR ? 2 : CLS);
    const int rot = p.rotate ? (((PAIR ? (j >> 1) : (j / CLS)) * KC) / (n_loaders > 0 ? n_loaders : 1)) % KC : 0;
    constexpr int STG_COLS = 5 * U;             // per row tile: gates (2U words) | c (U) | c_prev (U) | dh_out (U), staged by the epilogue warps
    constexpr uint32_t TMEM_COLS = tmem_cols_pow2(NQ * NCOL + NQ * STG_COLS);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_dg);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], CLS); }
        mbar_init(w_bar, 1);
        mbar_init(tmem_full, 1);
        if (PAIR) for (int i = 0; i < STAGES; ++i) mbar_init(&peer_full[i], 1);
        fence_barrier_init();
    }
    if (warp == 1) { if (PAIR) tmem_alloc_2sm(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS); }
    tc_fence_before();
    __syncthreads();
    if (CLS > 1 || PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(w_bar, (uint32_t)W_BYTES);
            for (int kc = 0; kc < KC; ++kc) tma_load_2d(sW + kc * CHUNK_W, &map_w, kc * 64, j * U, w_bar);
            int stage = 0; uint32_t phase = 0;
            const uint32_t box_bytes = (uint32_t)p.box_rows * 128u;
            for (int s = 1; s < p.T; ++s) {             // s-th processed step handles t = T-1-s and needs dgates_{t+1}
                const int t = p.T - 1 - s;
                while (ld_acquire(counter) < p.ctas_per_group * s) { }
                FSMG_TR(s, 0);
                fence_proxy_async_all();
                for (int kc = 0; kc < KC; ++kc) {
                    const int kcr = (kc + rot) % KC;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (PAIR) {
                        if (prank == 0) mbar_expect_tx(&full_bar[stage], 2 * box_bytes);
                        tma_load_3d_2sm(sA + stage * STAGE_BYTES, &map_dg, kcr * 64, row_base, t + 1, &full_bar[stage]);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_expect_tx(&full_bar[stage], box_bytes);
                    if (CLS == 1) tma_load_3d(sA + stage * STAGE_BYTES, &map_dg, kcr * 64, row_base, t + 1, &full_bar[stage]);
                    else if (kc % CLS == crank) tma_load_3d_mc(sA + stage * STAGE_BYTES, &map_dg, kcr * 64, row_base, t + 1, &full_bar[stage], CMASK);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                FSMG_TR(s, 1);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && PAIR && prank == 1) {
            mbar_wait(w_bar, 0);
            mbar_arrive_remote_release(&peer_full[0], 0);
        } else if (lane == 0) {
            constexpr uint32_t idesc = PAIR ? make_idesc_m(256, 2 * NCOL) : make_idesc(NCOL, false, false);
            mbar_wait(w_bar, 0);
            if (PAIR) mbar_wait(&peer_full[0], 0);
            tc_fence_after();
            int stage = 0; uint32_t phase = 0;
            for (int s = 1; s < p.T; ++s) {
                for (int kc = 0; kc < KC; ++kc) {
                    mbar_wait(&full_bar[stage], phase);
                    if (kc == 0) FSMG_TR(s, 2);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(sA + stage * STAGE_BYTES);
                    const uint32_t sb = smem_u32(sW + ((kc + rot) % KC) * CHUNK_W);   // same rotation as the producer
                    if (PAIR) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t a_desc = make_smem_desc(sa + k * 32, 16, 1024);
                            const uint64_t b_desc = make_smem_desc(sb + k * 32, 16, 1024);
                            umma_f16_2sm(tmem_base, a_desc, b_desc, idesc, (kc > 0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit_2sm_mc(&empty_bar[stage], (uint16_t)0x3);
                    } else {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t a_desc = make_smem_desc(sa + mt * 128 * 128 + k * 32, 16, 1024);
                                const uint64_t b_desc = make_smem_desc(sb + k * 32, 16, 1024);
                                umma_f16(tmem_base + mt * NCOL, a_desc, b_desc, idesc, (kc > 0 || k > 0) ? 1u : 0u);
                            }
                        }
                        if (CLS > 1) umma_commit_mc(&empty_bar[stage], CMASK); else umma_commit(&empty_bar[stage]);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (PAIR) umma_commit_2sm_mc(tmem_full, (uint16_t)0x3); else umma_commit(tmem_full);
                FSMG_TR(s, 3);
            }
        }
    } else {
        const int quad = warp & 3;
        const int q_idx = (warp - 2) >> 2;                 // quartet: row tile (plain) or unit slice of the pair (PAIR)
        const int mt = PAIR ? 0 : q_idx;
        const int ju = PAIR ? (j & ~1) + q_idx : j;        // unit slice whose gates this quartet differentiates
        float dc_state[U];
#pragma unroll
        for (int u = 0; u < U; ++u) dc_state[u] = 0.0f;
        uint32_t tf_phase = 0;
        const int lrow = mt * 128 + quad * 32 + lane;
        const bool ok = lrow < rows;
        const uint32_t t_acc = tmem_base + q_idx * NCOL + ((uint32_t)(quad * 32) << 16);
        const uint32_t t_stg = tmem_base + NQ * NCOL + q_idx * STG_COLS + ((uint32_t)(quad * 32) << 16);
        for (int s = 0; s < p.T; ++s) {
            const int t = p.T - 1 - s;
            const int64_t r = (int64_t)t * p.N + row_base + lrow;
            // The cell-backward operands of this step (forward stash + upstream gradient) do not depend on the tensor-core
            // result: fetch them NOW, while the dgates exchange and the MMAs of this step are in flight, and park them in
            // spare TMEM columns (TMEM as an explicit spill space: 5U words per row would not fit in registers).
            {
                const __half* gin = p.gates + r * p.G4p + ju * U;
                const float* cc = p.c + r * p.H + ju * U;
                const float* dho = p.dh_out + r * p.H + ju * U;
                if constexpr (U == 32) {
                    uint32_t w[32];
#pragma unroll
                    for (int q = 0; q < 4; q += 2) {        // two gates (2 x 16 words) per 32-column store
#pragma unroll
                        for (int qq = 0; qq < 2; ++qq)
#pragma unroll
                            for (int e = 0; e < 16; e += 4) {
                                const uint4 v = ok ? __ldcs(reinterpret_cast<const uint4*>(gin + (q + qq) * p.H + 2 * e)) : make_uint4(0, 0, 0, 0);
                                w[qq * 16 + e] = v.x; w[qq * 16 + e + 1] = v.y; w[qq * 16 + e + 2] = v.z; w[qq * 16 + e + 3] = v.w;
                            }
                        tmem_st32(t_stg + q * 16, w);
                    }
#pragma unroll
                    for (int which = 0; which < 3; ++which) {
                        const float* src = which == 0 ? cc : which == 1 ? cc - (int64_t)p.N * p.H : dho;
                        const bool have = ok && !(which == 1 && t == 0);
#pragma unroll
                        for (int e = 0; e < 32; e += 4) {
                            const float4 v = have ? __ldcs(reinterpret_cast<const float4*>(src + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            w[e] = __float_as_uint(v.x); w[e + 1] = __float_as_uint(v.y); w[e + 2] = __float_as_uint(v.z); w[e + 3] = __float_as_uint(v.w);
                        }
                        tmem_st32(t_stg + 2 * U + which * U, w);
                    }
                } else {   // U == 16
                    uint32_t w[16];
#pragma unroll
                    for (int q = 0; q < 4; q += 2) {
#pragma unroll
                        for (int qq = 0; qq < 2; ++qq)
#pragma unroll
                            for (int e = 0; e < 8; e += 4) {
                                const uint4 v = ok ? __ldcs(reinterpret_cast<const uint4*>(gin + (q + qq) * p.H + 2 * e)) : make_uint4(0, 0, 0, 0);
                                w[qq * 8 + e] = v.x; w[qq * 8 + e + 1] = v.y; w[qq * 8 + e + 2] = v.z; w[qq * 8 + e + 3] = v.w;
                            }
                        tmem_st16(t_stg + q * 8, w);
                    }
#pragma unroll
                    for (int which = 0; which < 3; ++which) {
                        const float* src = which == 0 ? cc : which == 1 ? cc - (int64_t)p.N * p.H : dho;
                        const bool have = ok && !(which == 1 && t == 0);
#pragma unroll
                        for (int e = 0; e < 16; e += 4) {
                            const float4 v = have ? __ldcs(reinterpret_cast<const float4*>(src + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            w[e] = __float_as_uint(v.x); w[e + 1] = __float_as_uint(v.y); w[e + 2] = __float_as_uint(v.z); w[e + 3] = __float_as_uint(v.w);
                        }
                        tmem_st16(t_stg + 2 * U + which * U, w);
                    }
                }
                tmem_st_wait();
                if (t > 0 && ok) {   // and pull the next processed step's lines towards L2
                    const int64_t rn = r - p.N;
#pragma unroll
                    for (int q = 0; q < 4; ++q) prefetch_l2(p.gates + rn * p.G4p + q * p.H + ju * U);
                    prefetch_l2(p.dh_out + rn * p.H + ju * U);
                    if (t > 1) prefetch_l2(p.c + (rn - p.N) * p.H + ju * U);
                }
            }
            if (s > 0) {
                mbar_wait(tmem_full, tf_phase);
                tf_phase ^= 1;
                tc_fence_after();
            }
            if (warp == 2 && lane == 0) FSMG_TR(s, 5);
            {
                __half* dgo = p.dgates + r * p.G4p + ju * U;
#pragma unroll
                for (int u0 = 0; u0 < U; u0 += 8) {
                    float acc[8];
                    uint32_t gw[4][4], cw[8], pw[8], dw[8];
                    if (s > 0) {
                        uint32_t rr[8];
                        tmem_ld8(t_acc + u0, rr);
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[e] = __uint_as_float(rr[e]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) tmem_ld4(t_stg + q * (U / 2) + u0 / 2, gw[q]);
                    tmem_ld8(t_stg + 2 * U + u0, cw);
                    tmem_ld8(t_stg + 3 * U + u0, pw);
                    tmem_ld8(t_stg + 4 * U + u0, dw);
                    tmem_ld_wait();
                    if (ok) {
                        __align__(16) __half dq[4][8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const __half2 hi2 = *reinterpret_cast<const __half2*>(&gw[0][e >> 1]);
                            const __half2 hj2 = *reinterpret_cast<const __half2*>(&gw[1][e >> 1]);
                            const __half2 hf2 = *reinterpret_cast<const __half2*>(&gw[2][e >> 1]);
                            const __half2 ho2 = *reinterpret_cast<const __half2*>(&gw[3][e >> 1]);
                            const float i_ = (e & 1) ? __high2float(hi2) : __low2float(hi2);
                            const float j_ = (e & 1) ? __high2float(hj2) : __low2float(hj2);
                            const float f_ = (e & 1) ? __high2float(hf2) : __low2float(hf2);
                            const float o_ = (e & 1) ? __high2float(ho2) : __low2float(ho2);
                            const float dhv = __uint_as_float(dw[e]) + acc[e];
                            const float tcv = tanh_fast(__uint_as_float(cw[e]));
                            const float d_o = dhv * tcv;
                            const float dc = dhv * o_ * (1.0f - tcv * tcv) + dc_state[u0 + e];
                            dq[0][e] = __float2half_rn(dc * j_ * i_ * (1.0f - i_));
                            dq[1][e] = __float2half_rn(dc * i_ * (1.0f - j_ * j_));
                            dq[2][e] = __float2half_rn(dc * __uint_as_float(pw[e]) * f_ * (1.0f - f_));
                            dq[3][e] = __float2half_rn(d_o * o_ * (1.0f - o_));
                            dc_state[u0 + e] = dc * f_;
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(dgo + q * p.H + u0) = *reinterpret_cast<uint4*>(dq[q]);
                    }
                }
            }
            tc_fence_before();
            if (warp == 2 && lane == 0) FSMG_TR(s, 6);
            named_bar_sync(1, 128 * NQ);
            if (warp == 2 && lane == 0) { FSMG_TR(s, 7); red_release_add(counter, 1); FSMG_TR(s, 8); }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CLS > 1 || PAIR) cluster_sync_all();
    if (warp == 1) { if (PAIR) tmem_dealloc_2sm(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS); }
}


// =====================================================================================================
// backward, second-generation pair kernel (default): lstm_bwd_pair2_kernel<U, NSUB>.
//
// What bounds the reverse-time recurrence (measured, profiles/r2_optimization_log.md).  The accumulator tile of a pair is only
// 256 rows x 2U columns (all the resident weight slice allows), so one step is 4H/16 = 128 small tcgen05.mma (at configs[1]);
// each costs ~85 cycles however it is issued — fetching the 128-row A tile from shared memory is not amortised over 64 output
// columns — and every full/empty barrier round trip of the ring costs the issuing thread another ~190 cycles.  The first pair
// kernel paid 128 x 90 + 32 x 190 = ~17.6k cycles per step for that, then ran the cell epilogue, the publish and the wait for the
// other 15 CTAs of the group (~13k cycles) strictly afterwards.  This kernel
//   * groups KS 64-column K chunks (TMA boxes) per ring stage: KS x fewer barrier round trips;
//   * builds the UMMA descriptors once per stage and advances them with 64-bit adds;
//   * writes the upstream gradient dh_out[t] INTO the accumulator before the step's MMAs start (every MMA accumulates), so the
//     epilogue reads dh = dh_out + dgates_{t+1} Wh^T straight from TMEM;
//   * NSUB = 1: splits a row's U units over two threads (16 epilogue warps = unit slice of the pair x unit half x TMEM lane
//     quadrant), halving the dependent epilogue chain; all cell operands (gates, c, c_prev) are prefetched into spare TMEM columns;
//   * publishes per warp (the group counter counts warps): no CTA-wide barrier before the release.
//   * NSUB = 2 (FSMG_LSTM_BSPLIT=2): additionally cuts the group's rows into two sub-groups that the producer and the MMA thread
//     serve alternately (one sub-group's epilogue / exchange hides behind the other's MMAs).  It doubles the number of MMAs per
//     step — the bound resource — and measured slower than NSUB = 1 wherever the group has more than a few rows.
// =====================================================================================================
constexpr int LSTM_BSPLIT_THREADS = 64 + 512;
constexpr int LSTM_BSPLIT_MAX_STAGES = 16;
constexpr int LSTM_BSPLIT_BAR_BYTES = 512;

// HALF_M (NSUB = 2 only, <= 64 rows per CTA and sub-group): the pair's MMA is 128 x 2U x 16 instead of 256 x 2U x 16 — each CTA feeds 64
// rows of A instead of 128 (the sub-group only has 40 at configs[1]), which is what a small-N MMA costs.  Measured with
// fsmg_debug_mma_probe: tcgen05.mma.cta_group::2 with M = 128 puts row r of CTA c's 64 rows in TMEM lane r for output columns
// [0, N/2) and in lane 64 + r for columns [N/2, N), both at TMEM columns [0, N/2).  With N = 2U that is: lanes 0-63 = the rows x the
// first CTA's unit slice, lanes 64-127 = the same rows x the second CTA's unit slice — every lane quadrant has rows to work on, the
// accumulator takes U columns instead of 2U, and the columns saved hold c_{t-1} for a 2-way unit split of the epilogue.
template <int U, int NSUB, bool HALF_M = false>
__global__ void __launch_bounds__(LSTM_BSPLIT_THREADS, 1)
lstm_bwd_pair2_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_dg, LstmParams p) {
    constexpr int CHUNK_W = U * 128;
    constexpr int UH = (NSUB == 1 || HALF_M) ? U / 2 : U;   // units per epilogue thread
    constexpr int ACC_COLS = HALF_M ? U : 2 * U;            // one sub-group's accumulator (see the M = 128 layout above)
    constexpr bool STAGE_CPREV = NSUB == 1 || HALF_M;       // c_{t-1} staged in TMEM too (else no columns left: read from L2)
    constexpr int STG_COLS = (STAGE_CPREV ? 4 : 3) * UH;   // per epilogue quartet: gates (2 UH words) | c (UH) [| c_prev (UH)]
    constexpr uint32_t TMEM_COLS = tmem_cols_pow2(NSUB * ACC_COLS + 4 * STG_COLS);
    static_assert(NSUB * ACC_COLS + 4 * STG_COLS <= 512, "TMEM budget");
    static_assert(UH % 8 == 0, "epilogue works on 8-unit chunks");
    constexpr int WARPS_PER_SUB = 16 / NSUB;    // epilogue warps of one CTA that belong to one sub-group
    const int KC = (4 * p.H) / 64;
    const int W_BYTES = KC * CHUNK_W;
    const int STAGES = p.stages;
    const int STAGE_BYTES = p.stage_bytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;
    uint8_t* sA = sW + W_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LSTM_MAX_DYN - 1024 - LSTM_BSPLIT_BAR_BYTES);
    uint64_t* full_bar = bars;                          // [16]
    uint64_t* empty_bar = bars + LSTM_BSPLIT_MAX_STAGES;   // [16]
    uint64_t* w_bar = bars + 2 * LSTM_BSPLIT_MAX_STAGES;
    uint64_t* tmem_full = w_bar + 1;                    // [2] per sub-group
    uint64_t* pre_ready = w_bar + 3;                    // [2] per sub-group (leader only): dh_out of the next step is staged in the accumulator
    uint64_t* peer_w = w_bar + 5;                       // leader only: the peer's weight slice is resident
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x / p.ctas_per_group, j = blockIdx.x % p.ctas_per_group;
    const int prank = (int)cluster_ctarank();
    const int group_row0 = p.row_offset + g * p.rows_per_group;
    const int group_rows = min(p.rows_per_group, p.row_offset + p.n_rows - group_row0);
    const int hr = p.box_rows;                          // rows per (sub-group, CTA): multiple of 8, <= 128
    const int n_loaders = p.ctas_per_group / 2;
    const int rot = p.rotate ? (((j >> 1) * KC) / (n_loaders > 0 ? n_loaders : 1)) % KC : 0;
    const int pub_per_step = p.ctas_per_group;          // one release per (CTA, sub-group) and step

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_dg);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(w_bar, 1);
        mbar_init(&tmem_full[0], 1);
        mbar_init(&tmem_full[1], 1);
        mbar_init(&pre_ready[0], 2 * WARPS_PER_SUB);    // the sub-group's epilogue warps of both CTAs of the pair
        mbar_init(&pre_ready[1], 2 * WARPS_PER_SUB);
        mbar_init(peer_w, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(w_bar, (uint32_t)W_BYTES);
            for (int kc = 0; kc < KC; ++kc) tma_load_2d(sW + kc * CHUNK_W, &map_w, kc * 64, j * U, w_bar);
            int stage = 0; uint32_t phase = 0;
            const uint32_t box_bytes = (uint32_t)hr * 128u;
            for (int s = 1; s < p.T; ++s) {             // s-th processed step handles t = T-1-s and needs dgates_{t+1}
                const int t = p.T - 1 - s;
                for (int sub = 0; sub < NSUB; ++sub) {
                    const int* counter = p.counters + 2 * g + sub;
                    while (ld_acquire(counter) < pub_per_step * s) { }
                    if (sub == 0) FSMG_TR(s, 0);
                    fence_proxy_async_all();
                    const int row0 = group_row0 + sub * 2 * hr + prank * hr;
                    int kcr = rot;
                    for (int kb = 0; kb < KC; kb += p.ks) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (prank == 0) mbar_expect_tx(&full_bar[stage], 2 * p.ks * box_bytes);
                        for (int i = 0; i < p.ks; ++i) {
                            tma_load_3d_2sm(sA + stage * STAGE_BYTES + i * p.box_pitch, &map_dg, kcr * 64, row0, t + 1, &full_bar[stage]);
                            if (++kcr == KC) kcr = 0;
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (sub == 0) FSMG_TR(s, 1);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && prank == 1) {
            mbar_wait(w_bar, 0);
            mbar_arrive_remote_release(peer_w, 0);
        } else if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_m(HALF_M ? 128 : 256, 2 * U);
            mbar_wait(w_bar, 0);
            mbar_wait(peer_w, 0);
            tc_fence_after();
            int stage = 0; uint32_t phase = 0;
            uint32_t pr_phase[2] = {0, 0};
            const uint64_t b_desc0 = make_smem_desc(smem_u32(sW), 16, 1024);
            const uint64_t a_step = (uint64_t)(p.box_pitch >> 4);
            for (int s = 1; s < p.T; ++s) {
                for (int sub = 0; sub < NSUB; ++sub) {
                    mbar_wait(&pre_ready[sub], pr_phase[sub]);      // dh_out[t] sits in the accumulator: every MMA accumulates
                    pr_phase[sub] ^= 1;
                    tc_fence_after();
                    const uint32_t d_buf = tmem_base + (uint32_t)(sub * ACC_COLS);
                    int kcr = rot;
                    for (int kb = 0; kb < KC; kb += p.ks) {
                        mbar_wait(&full_bar[stage], phase);
                        if (kb == 0 && sub == 0) FSMG_TR(s, 2);
                        tc_fence_after();
                        // descriptors: built once per stage, advanced with plain 64-bit adds (start-address field = bytes >> 4)
                        uint64_t a_desc = make_smem_desc(smem_u32(sA + stage * STAGE_BYTES), 16, 1024);
                        for (int i = 0; i < p.ks; ++i) {
                            const uint64_t b_desc = b_desc0 + (uint64_t)(kcr * (CHUNK_W >> 4));
                            umma_f16_2sm(d_buf, a_desc, b_desc, idesc, 1u);
                            umma_f16_2sm(d_buf, a_desc + 2, b_desc + 2, idesc, 1u);
                            umma_f16_2sm(d_buf, a_desc + 4, b_desc + 4, idesc, 1u);
                            umma_f16_2sm(d_buf, a_desc + 6, b_desc + 6, idesc, 1u);
                            a_desc += a_step;
                            if (++kcr == KC) kcr = 0;
                        }
                        umma_commit_2sm_mc(&empty_bar[stage], (uint16_t)0x3);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit_2sm_mc(&tmem_full[sub], (uint16_t)0x3);
                    if (sub == 0) FSMG_TR(s, 3);
                }
            }
        }
    } else {
        // ===== epilogue: 16 warps x 32 lanes; thread <-> one row, UH units.  quartet qd = (warp - 2) / 4, TMEM lane quadrant quad = warp % 4
        //   NSUB = 1:          unit slice of the pair us = qd & 1, unit half qd >> 1,              row = quad * 32 + lane
        //   NSUB = 2:          sub-group qd >> 1, unit slice us = qd & 1 (all U units),            row = quad * 32 + lane
        //   NSUB = 2, HALF_M:  sub-group qd >> 1, unit half qd & 1, unit slice us = quad >> 1,     row = (quad & 1) * 32 + lane
        //   NSUB = 1, HALF_M:  unit half qd & 1 (quartets 2, 3 idle), unit slice us = quad >> 1,   row = (quad & 1) * 32 + lane
        const int quad = warp & 3;
        const int qd = (warp - 2) >> 2;
        const int sub = NSUB == 1 ? 0 : qd >> 1;
        const int us = HALF_M ? quad >> 1 : qd & 1;
        const int uhalf = HALF_M ? qd & 1 : NSUB == 1 ? qd >> 1 : 0;
        const int ucol = ((j & ~1) + us) * U + uhalf * UH;   // first hidden unit of this thread
        const int row_base = group_row0 + sub * 2 * hr + prank * hr;
        const int rows = max(0, min(hr, group_rows - sub * 2 * hr - prank * hr));
        const int lrow0 = HALF_M ? (quad & 1) * 32 : quad * 32;
        const int lrow = lrow0 + lane;
        const bool ok = lrow < rows && !(HALF_M && NSUB == 1 && qd >= 2);
        const bool warp_active = lrow0 < rows && !(HALF_M && NSUB == 1 && qd >= 2);   // warp-uniform: warps without work only keep the barriers in step
        int* counter = p.counters + 2 * g + sub;
        const bool tracer = (qd == 0 && quad == 0 && lane == 0);
        const uint32_t t_acc = tmem_base + (uint32_t)(sub * ACC_COLS + (HALF_M ? 0 : us * U) + uhalf * UH) + ((uint32_t)(quad * 32) << 16);
        const uint32_t t_stg = tmem_base + (uint32_t)(NSUB * ACC_COLS + qd * STG_COLS) + ((uint32_t)(quad * 32) << 16);
        float dc_state[UH];
#pragma unroll
        for (int u = 0; u < UH; ++u) dc_state[u] = 0.0f;
        uint32_t tf_phase = 0;
        // forward stash of step t (gates, c [, c_prev]) -> staging columns, upstream gradient dh_out[t] -> accumulator; 16 words per store
        auto stage_step = [&](int t) {
            const int64_t r = (int64_t)t * p.N + row_base + lrow;
            const __half* gin = p.gates + r * p.G4p + ucol;
            uint32_t w[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {                   // gate q: UH halves = UH/2 words
#pragma unroll
                for (int e0 = 0; e0 < UH / 2; e0 += 16) {
                    constexpr int NW = UH / 2 < 16 ? UH / 2 : 16;
#pragma unroll
                    for (int e = 0; e < NW; e += 4) {
                        const uint4 v = ok ? __ldcs(reinterpret_cast<const uint4*>(gin + q * p.H + 2 * (e0 + e))) : make_uint4(0, 0, 0, 0);
                        w[e] = v.x; w[e + 1] = v.y; w[e + 2] = v.z; w[e + 3] = v.w;
                    }
                    tmem_st_n<NW>(t_stg + q * (UH / 2) + e0, w);
                }
            }
#pragma unroll
            for (int which = 0; which < (STAGE_CPREV ? 3 : 2); ++which) {
                const float* src = which == 0 ? p.c + r * p.H + ucol : which == 1 ? p.dh_out + r * p.H + ucol : p.c + (r - p.N) * p.H + ucol;
                const bool have = ok && !(which == 2 && t == 0);
                const uint32_t dst = which == 0 ? t_stg + 2 * UH : which == 1 ? t_acc : t_stg + 3 * UH;
#pragma unroll
                for (int e0 = 0; e0 < UH; e0 += 16) {
                    constexpr int NW = UH < 16 ? UH : 16;
#pragma unroll
                    for (int e = 0; e < NW; e += 4) {
                        const float4 v = have ? __ldcs(reinterpret_cast<const float4*>(src + e0 + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        w[e] = __float_as_uint(v.x); w[e + 1] = __float_as_uint(v.y); w[e + 2] = __float_as_uint(v.z); w[e + 3] = __float_as_uint(v.w);
                    }
                    tmem_st_n<NW>(dst + e0, w);
                }
            }
            tmem_st_wait();
        };
        if (warp_active) stage_step(p.T - 1);
        for (int s = 0; s < p.T; ++s) {
            const int t = p.T - 1 - s;
            const int64_t r = (int64_t)t * p.N + row_base + lrow;
            const bool have_prev = ok && t > 0;
            const float* cprev = p.c + (r - p.N) * p.H + ucol;         // NSUB = 2 only: c_{t-1} from L2 (prefetched one step ago)
            float4 nx0 = make_float4(0.f, 0.f, 0.f, 0.f), nx1 = nx0;
            if (!STAGE_CPREV && have_prev) { nx0 = __ldg(reinterpret_cast<const float4*>(cprev)); nx1 = __ldg(reinterpret_cast<const float4*>(cprev + 4)); }
            if (s > 0) {
                mbar_wait(&tmem_full[sub], tf_phase);
                tf_phase ^= 1;
                tc_fence_after();
            }
            if (tracer) FSMG_TR(s, 5);
            if (warp_active) {
                __half* dgo = p.dgates + r * p.G4p + ucol;
#pragma unroll
                for (int u0 = 0; u0 < UH; u0 += 8) {
                    uint32_t dh[8], gw[4][4], cw[8], pw[8];
                    if constexpr (!STAGE_CPREV) {
                        pw[0] = __float_as_uint(nx0.x); pw[1] = __float_as_uint(nx0.y); pw[2] = __float_as_uint(nx0.z); pw[3] = __float_as_uint(nx0.w);
                        pw[4] = __float_as_uint(nx1.x); pw[5] = __float_as_uint(nx1.y); pw[6] = __float_as_uint(nx1.z); pw[7] = __float_as_uint(nx1.w);
                        if (u0 + 8 < UH && have_prev) {
                            nx0 = __ldg(reinterpret_cast<const float4*>(cprev + u0 + 8));
                            nx1 = __ldg(reinterpret_cast<const float4*>(cprev + u0 + 12));
                        }
                    }
                    tmem_ld8(t_acc + u0, dh);
#pragma unroll
                    for (int q = 0; q < 4; ++q) tmem_ld4(t_stg + q * (UH / 2) + u0 / 2, gw[q]);
                    tmem_ld8(t_stg + 2 * UH + u0, cw);
                    if constexpr (STAGE_CPREV) tmem_ld8(t_stg + 3 * UH + u0, pw);
                    tmem_ld_wait();
                    if (ok) {
                        __align__(16) __half dq[4][8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const __half2 hi2 = *reinterpret_cast<const __half2*>(&gw[0][e >> 1]);
                            const __half2 hj2 = *reinterpret_cast<const __half2*>(&gw[1][e >> 1]);
                            const __half2 hf2 = *reinterpret_cast<const __half2*>(&gw[2][e >> 1]);
                            const __half2 ho2 = *reinterpret_cast<const __half2*>(&gw[3][e >> 1]);
                            const float i_ = (e & 1) ? __high2float(hi2) : __low2float(hi2);
                            const float j_ = (e & 1) ? __high2float(hj2) : __low2float(hj2);
                            const float f_ = (e & 1) ? __high2float(hf2) : __low2float(hf2);
                            const float o_ = (e & 1) ? __high2float(ho2) : __low2float(ho2);
                            const float dhv = __uint_as_float(dh[e]);              // dh_out[t] + dgates_{t+1} * Wh^T (accumulated in TMEM)
                            const float tcv = tanh_fast(__uint_as_float(cw[e]));
                            const float d_o = dhv * tcv;
                            const float dc = dhv * o_ * (1.0f - tcv * tcv) + dc_state[u0 + e];
                            dq[0][e] = __float2half_rn(dc * j_ * i_ * (1.0f - i_));
                            dq[1][e] = __float2half_rn(dc * i_ * (1.0f - j_ * j_));
                            dq[2][e] = __float2half_rn(dc * __uint_as_float(pw[e]) * f_ * (1.0f - f_));
                            dq[3][e] = __float2half_rn(d_o * o_ * (1.0f - o_));
                            dc_state[u0 + e] = dc * f_;
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(dgo + q * p.H + u0) = *reinterpret_cast<uint4*>(dq[q]);
                    }
                }
            }
            // publish: barrier among the sub-group's epilogue warps, then ONE gpu-scope release per (CTA, sub-group) — it is cumulative over
            // the dgates stores the other warps made before the barrier.  (A release per warp, 16 fences per CTA and step, measured
            // slower: the group's counter completed 7-10k cycles after the first warp's publish instead of 3k.)
            tc_fence_before();
            if (tracer) FSMG_TR(s, 6);
            named_bar_sync(1 + sub, 32 * WARPS_PER_SUB);
            if (qd == (NSUB == 1 ? 0 : 2 * sub) && quad == 0 && lane == 0) { if (tracer) FSMG_TR(s, 7); red_release_add(counter, 1); if (tracer) FSMG_TR(s, 8); }
            if (s + 1 < p.T) {
                // stage the next processed step while the exchange is in flight (and, NSUB = 2, the other sub-group's MMAs run)
                if (warp_active) stage_step(t - 1);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (prank == 1) mbar_arrive_remote(&pre_ready[sub], 0); else mbar_arrive(&pre_ready[sub]); }
                if (t > 1 && ok) {   // and pull the step after that towards L2
                    const int64_t rn = r - 2 * (int64_t)p.N;
#pragma unroll
                    for (int q = 0; q < 4; ++q) prefetch_l2(p.gates + rn * p.G4p + q * p.H + ucol);
                    prefetch_l2(p.dh_out + rn * p.H + ucol);
                    if (t > 2 || !STAGE_CPREV) prefetch_l2(p.c + (rn - (STAGE_CPREV ? p.N : 0)) * p.H + ucol);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
}

}  // namespace tc

// =====================================================================================================
// host side
// =====================================================================================================
// ring geometry: stages hold exactly the exchanged rows (rounded to the 1024-B swizzle atom); the last stage keeps room for
// the full MT*128-row window the UMMA descriptors address
static inline void lstm_ring(int w_bytes, int box_rows, int mt, int* stage_bytes, int* stages) {
    const int avail = tc::LSTM_MAX_DYN - 1024 - 256 - w_bytes;
    int sb = (int)round_up((int64_t)box_rows * 128, 1024);
    int window = mt * 128 * 128;
    int n = (avail - window) / sb + 1;
    if (n > 8) n = 8;
    if (n < 2) { sb = window; n = avail / window; }
    *stage_bytes = sb;
    *stages = n;
}

// pair + split backward: 16-deep ring of the narrow sub-group boxes, barrier block of LSTM_BSPLIT_BAR_BYTES
static inline void lstm_ring_bsplit(int w_bytes, int box_rows, int ks, int tile_rows, int* box_pitch, int* stage_bytes, int* stages) {
    const int avail = tc::LSTM_MAX_DYN - 1024 - tc::LSTM_BSPLIT_BAR_BYTES - w_bytes;
    const int pitch = (int)round_up((int64_t)box_rows * 128, 1024);
    const int window = tile_rows * 128;   // the UMMA descriptors address a full tile (128 rows per CTA, 64 with HALF_M) from every box base
    const int sb = ks * pitch;
    int n = (avail - (window - pitch)) / sb;
    if (n > tc::LSTM_BSPLIT_MAX_STAGES) n = tc::LSTM_BSPLIT_MAX_STAGES;
    *box_pitch = pitch;
    *stage_bytes = sb;
    *stages = n;
}

struct LstmPlan {
    int U, MT, C, G, rows_per_group, box_rows, rows_per_launch, cls;
    bool ok, pair;
};

static inline LstmPlan lstm_plan(const TcContext& c, int N, int H, bool want_pair = false, int cls_req = -1) {
    LstmPlan pl;
    memset(&pl, 0, sizeof pl);
    pl.ok = false;
    if (H % 64 != 0 || H < 64) return pl;
    int U = 0;
    if (4 * 32 * H * 2 <= 128 * 1024 && H % 32 == 0) U = 32;
    else if (4 * 16 * H * 2 <= 128 * 1024 && H % 16 == 0) U = 16;
    if (!U) return pl;
    pl.U = U;
    pl.C = H / U;
    int cls = cls_req >= 0 ? cls_req : c.lstm_cluster;
    if (cls != 1 && cls != 2 && cls != 4) cls = 1;
    while (cls > 1 && (pl.C % cls) != 0) cls >>= 1;
    pl.cls = cls;
    // clusters of 4 cannot use every SM (GPC sizes are not multiples of 4): 132 co-resident CTAs at most
    // lstm_reserve_sms: every CTA of these cooperative kernels must be resident at once; the SMs left free host the (few-CTA) NCCL
    // kernels of a gradient all-reduce that overlaps the backward pass (fsmg_set_stage_events)
    const int usable = c.num_sms - c.lstm_reserve_sms > 0 ? c.num_sms - c.lstm_reserve_sms : c.num_sms;
    const int sms = cls == 4 ? (usable < 132 ? usable / 4 * 4 : 132) : usable;   // (clusters of 8 do not fit a cooperative launch at 227 KB/CTA)
    int gmax = sms / pl.C;
    if (gmax < 1) return pl;
    int mg = cdiv(N, gmax);
    if (mg > 256) mg = 256;                       // larger batches are processed in slices of gmax*256 sequences
    mg = (int)round_up(mg, 8);
    pl.rows_per_group = mg;
    pl.box_rows = mg;
    pl.MT = mg > 128 ? 2 : 1;
    pl.pair = want_pair && (pl.C % 2 == 0) && mg > 8;
    if (pl.pair) {            // cta_group::2: each CTA of a pair streams half of the group's rows
        pl.cls = 1;
        pl.MT = 1;
        pl.box_rows = (int)round_up(cdiv(mg, 2), 8);
    }
    pl.rows_per_launch = gmax * mg;
    pl.G = gmax;
    pl.ok = true;
    return pl;
}

static inline int make_map_f16_3d(const TcContext& c, CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                                  uint64_t ld1, uint64_t ld2, uint32_t b0, uint32_t b1) {
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {ld1 * 2, ld2 * 2};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = c.encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(-2, "cuTensorMapEncodeTiled(3d) failed (%d)", (int)r);
    return 0;
}

template <typename K>
static inline int lstm_launch(K kernel, int grid, int threads, int cls, int smem_bytes, const CUtensorMap& mw, const CUtensorMap& mx, const tc::LstmParams& p,
                              cudaStream_t s) {
    FSMG_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: they wait on each other's counters
    // FSMG_COOP=0 drops the attribute (profiling only: ncu cannot replay a cooperative launch that also carries a cluster
    // dimension — the launch list stopped at the pair-mode backward kernel; the grid never exceeds the SM count, so with the
    // stream otherwise idle every CTA is resident anyway)
    static const int coop = [] { const char* e = getenv("FSMG_COOP"); return e ? atoi(e) : 1; }();
    attr[0].val.cooperative = coop ? 1 : 0;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cls; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cls > 1 ? 2 : 1;
    FSMG_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, mw, mx, p));
    return 0;
}

#define FSMG_LSTM_GO(KERNEL, UU, MM, CC, PP) RC_ = lstm_launch(tc::KERNEL<UU, MM, CC, PP>, G * pl.C, tc::lstm_threads((PP) ? 2 : (MM)), (PP) ? 2 : (CC), smem, mw, mx, p, s)
#define FSMG_LSTM_DISPATCH(KERNEL, RC, ALLOW_PAIR)                                                                \
    do {                                                                                                          \
        int RC_ = 0;                                                                                              \
        if (pl.pair && (ALLOW_PAIR)) {                                                                            \
            if (pl.U == 32) FSMG_LSTM_GO(KERNEL, 32, 1, 1, true); else FSMG_LSTM_GO(KERNEL, 16, 1, 1, true);      \
        } else if (pl.U == 32 && pl.MT == 2) {                                                                    \
            if (pl.cls == 4) FSMG_LSTM_GO(KERNEL, 32, 2, 4, false);                                               \
            else if (pl.cls == 2) FSMG_LSTM_GO(KERNEL, 32, 2, 2, false);                                          \
            else FSMG_LSTM_GO(KERNEL, 32, 2, 1, false);                                                           \
        } else if (pl.U == 32) {                                                                                  \
            if (pl.cls == 4) FSMG_LSTM_GO(KERNEL, 32, 1, 4, false);                                               \
            else if (pl.cls == 2) FSMG_LSTM_GO(KERNEL, 32, 1, 2, false);                                          \
            else FSMG_LSTM_GO(KERNEL, 32, 1, 1, false);                                                           \
        } else if (pl.MT == 2) {                                                                                  \
            if (pl.cls == 4) FSMG_LSTM_GO(KERNEL, 16, 2, 4, false);                                               \
            else if (pl.cls == 2) FSMG_LSTM_GO(KERNEL, 16, 2, 2, false);                                          \
            else FSMG_LSTM_GO(KERNEL, 16, 2, 1, false);                                                           \
        } else {                                                                                                  \
            if (pl.cls == 4) FSMG_LSTM_GO(KERNEL, 16, 1, 4, false);                                               \
            else if (pl.cls == 2) FSMG_LSTM_GO(KERNEL, 16, 1, 2, false);                                          \
            else FSMG_LSTM_GO(KERNEL, 16, 1, 1, false);                                                           \
        }                                                                                                         \
        RC = RC_;                                                                                                 \
    } while (0)

static inline void lstm_trace_dump(TcContext& c, const char* what, cudaStream_t s) {
    cudaStreamSynchronize(s);
    long long host[8 * 16];
    cudaMemcpy(host, c.trace, sizeof host, cudaMemcpyDeviceToHost);
    static const char* names[9] = {"P:counter_ok", "P:tma_issued", "M:first_full", "M:commit", "E:staged", "E:tmem_full", "E:computed", "E:barrier", "E:published"};
    fprintf(stderr, "[fsmg trace] %s (CTA 0, clock64 cycles relative to step's E:tmem_full of previous row)\n", what);
    for (int st = 1; st < 8; ++st) {
        long long base = host[(st - 1) * 16 + 8];   // previous step's publish
        fprintf(stderr, "  step %2d:", st + 8);
        for (int k = 0; k < 9; ++k) fprintf(stderr, " %s=%lld", names[k], host[st * 16 + k] ? host[st * 16 + k] - base : -1);
        fprintf(stderr, "\n");
    }
}

static inline bool tc_recurrent_supported(TcContext& c, int N, int H) {
    if (!c.ready || !c.enabled || !c.counters) return false;
    const char* env = getenv("FSMG_PERSISTENT");
    if (env && atoi(env) == 0) return false;
    return lstm_plan(c, N, H).ok;
}

// pre [T*N,4H] fp32, WhT16 [4H,Hp] fp16 -> gates [T*N,G4p], c [T*N,H], hs [T*N,Hp]
static inline int tc_lstm_forward(TcContext& c, const __half* pre, const int32_t* pre_ids, const __half* WhT16, __half* gates, float* cbuf,
                                  __half* hs, int N, int T, int H, int Hp, int G4p, cudaStream_t s) {
    LstmPlan pl = lstm_plan(c, N, H, (c.lstm_pair & 2) != 0);
    if (!pl.ok) return set_error(-1, "persistent LSTM: unsupported shape N=%d H=%d", N, H);
    CUtensorMap mw, mh;
    int rc = make_map_f16(c, &mw, WhT16, (uint64_t)H, (uint64_t)4 * H, (uint64_t)Hp, 64, (uint32_t)pl.U);
    if (rc) return rc;
    rc = make_map_f16_3d(c, &mh, hs, (uint64_t)H, (uint64_t)N, (uint64_t)T, (uint64_t)Hp, (uint64_t)N * Hp, 64, (uint32_t)pl.box_rows);
    if (rc) return rc;
    // split schedule (default): two independently progressing halves per group hide each other's exchange latency
    const bool split = c.lstm_split && !pl.pair && pl.cls == 1 && pl.rows_per_group >= 16;
    // groups of a few rows (configs[2]: 24) run ONE "half": a second one would only double the number of 128-row MMAs per step, which is
    // what bounds the kernel there (measured: FSMG_LSTM_NH)
    const int nh = c.lstm_nh > 0 ? c.lstm_nh : (pl.rows_per_group > 32 ? 2 : 1);
    const int half_rows = nh == 2 ? (int)round_up(cdiv(pl.rows_per_group, 2), 8) : (int)round_up(pl.rows_per_group, 8);
    CUtensorMap mh_split = mh;
    if (split) {
        rc = make_map_f16_3d(c, &mh_split, hs, (uint64_t)H, (uint64_t)N, (uint64_t)T, (uint64_t)Hp, (uint64_t)N * Hp, 64, (uint32_t)half_rows);
        if (rc) return rc;
    }
    for (int off = 0; off < N; off += pl.rows_per_launch) {
        int n_rows = N - off < pl.rows_per_launch ? N - off : pl.rows_per_launch;
        int G = cdiv(n_rows, pl.rows_per_group);
        FSMG_CUDA_OK(cudaMemsetAsync(c.counters, 0, sizeof(int) * 256, s));
        tc::LstmParams p;
        memset(&p, 0, sizeof p);
        p.pre = pre; p.pre_ids = pre_ids; p.gates = gates; p.c = cbuf; p.hs = hs; p.counters = c.counters; p.rotate = (c.lstm_rot & 2) != 0;
        p.N = N; p.T = T; p.H = H; p.Hp = Hp; p.G4p = G4p;
        p.ctas_per_group = pl.C; p.rows_per_group = pl.rows_per_group; p.box_rows = pl.box_rows; p.row_offset = off; p.n_rows = n_rows;
        const int w_bytes = (H / 64) * 4 * pl.U * 128;
        const int smem = tc::LSTM_MAX_DYN;
        lstm_ring(w_bytes, pl.box_rows, pl.MT, &p.stage_bytes, &p.stages);
        const CUtensorMap& mx = mh;
        const bool trace = getenv("FSMG_TRACE") != nullptr && c.trace != nullptr;
        if (trace) { cudaMemsetAsync(c.trace, 0, 8 * 16 * sizeof(long long), s); p.trace = c.trace; }
        if (split) {
            // two independent halves per group (own counters: 2 per group), one CTA per (group, unit slice)
            p.box_rows = half_rows;
            int ks = c.lstm_fks;
            while (ks > 1 && ((H / 64) % ks) != 0) ks >>= 1;
            for (;; ks >>= 1) {      // ring of (ks boxes)-stages: at least 2 stages, at most 8 (barrier array)
                p.box_pitch = (int)round_up((int64_t)half_rows * 128, 1024);
                p.stage_bytes = ks * p.box_pitch;
                p.stages = (tc::LSTM_MAX_DYN - 1024 - 256 - w_bytes - (128 * 128 - p.box_pitch)) / p.stage_bytes;
                if (p.stages > 8) p.stages = 8;
                if (p.stages >= 2 || ks == 1) break;
            }
            p.ks = ks;
            p.pub_cta = c.lstm_pub_cta;
            p.nh = nh;
            if (pl.U == 32) rc = lstm_launch(tc::lstm_fwd_split_kernel<32>, G * pl.C, tc::LSTM_SPLIT_THREADS, 1, smem, mw, mh_split, p, s);
            else rc = lstm_launch(tc::lstm_fwd_split_kernel<16>, G * pl.C, tc::LSTM_SPLIT_THREADS, 1, smem, mw, mh_split, p, s);
        } else {
            FSMG_LSTM_DISPATCH(lstm_fwd_persistent_kernel, rc, true);
        }
        if (trace && !rc) lstm_trace_dump(c, "lstm_fwd_persistent", s);
        if (rc) return rc;
    }
    return 0;
}

// dh_out [T*N,H] fp32, Wh_rows = kernel[in:, :] fp16 [H, G4p] -> dgates [T*N,G4p]
static inline int tc_lstm_backward(TcContext& c, const float* dh_out, const __half* Wh_rows, const __half* gates, const float* cbuf,
                                   __half* dgates, int N, int T, int H, int G4p, cudaStream_t s) {
    LstmPlan pl = lstm_plan(c, N, H, (c.lstm_pair & 1) != 0);
    if (!pl.ok) return set_error(-1, "persistent LSTM: unsupported shape N=%d H=%d", N, H);
    CUtensorMap mw, md;
    int rc = make_map_f16(c, &mw, Wh_rows, (uint64_t)4 * H, (uint64_t)H, (uint64_t)G4p, 64, (uint32_t)pl.U);
    if (rc) return rc;
    // second-generation pair kernel (default): lstm_bwd_pair2_kernel<U, NSUB>; FSMG_LSTM_BSPLIT = 0 old pair kernel, 1 (default) NSUB = 1,
    // 2 NSUB = 2 (two alternating sub-groups per pair)
    // auto (measured, profiles/r2_optimization_log.md): two alternating sub-groups pay off only when a group holds many rows
    // (configs[1]: 160 rows per group, 2.16 -> 2.05 ms); with few rows (configs[2]: 24) they only double the MMA count (3.5 -> 5.5 ms)
    const int mode = c.lstm_bsplit >= 0 ? c.lstm_bsplit : (pl.rows_per_group > 64 ? 2 : 1);
    const int nsub = mode == 2 ? 2 : 1;
    const int sub_rows = (int)round_up(cdiv(pl.rows_per_group, 2 * nsub), 8);
    int bs_stage = 0, bs_stages = 0, bs_pitch = 0;
    int ks = c.lstm_ks > 0 ? c.lstm_ks : (sub_rows <= 64 ? 8 : 4);
    while (ks > 1 && ((4 * H / 64) % ks) != 0) ks >>= 1;
    const bool half_m = sub_rows <= 64 && c.lstm_half_m;      // 128-row pair MMAs (64 rows of A per CTA)
    const int tile_rows = half_m ? 64 : 128;
    lstm_ring_bsplit((4 * H / 64) * pl.U * 128, sub_rows, ks, tile_rows, &bs_pitch, &bs_stage, &bs_stages);
    while (bs_stages < 2 && ks > 1) { ks >>= 1; lstm_ring_bsplit((4 * H / 64) * pl.U * 128, sub_rows, ks, tile_rows, &bs_pitch, &bs_stage, &bs_stages); }
    const bool bsplit = mode != 0 && pl.pair && pl.rows_per_group >= 16 && bs_stages >= 2 && sub_rows <= 128;
    rc = make_map_f16_3d(c, &md, dgates, (uint64_t)4 * H, (uint64_t)N, (uint64_t)T, (uint64_t)G4p, (uint64_t)N * G4p, 64,
                         (uint32_t)(bsplit ? sub_rows : pl.box_rows));
    if (rc) return rc;
    for (int off = 0; off < N; off += pl.rows_per_launch) {
        int n_rows = N - off < pl.rows_per_launch ? N - off : pl.rows_per_launch;
        int G = cdiv(n_rows, pl.rows_per_group);
        FSMG_CUDA_OK(cudaMemsetAsync(c.counters, 0, sizeof(int) * 256, s));
        tc::LstmParams p;
        memset(&p, 0, sizeof p);
        p.dh_out = dh_out; p.dgates = dgates; p.gates = const_cast<__half*>(gates); p.c = const_cast<float*>(cbuf); p.counters = c.counters; p.rotate = (c.lstm_rot & 1) != 0;
        p.N = N; p.T = T; p.H = H; p.Hp = 0; p.G4p = G4p;
        p.ctas_per_group = pl.C; p.rows_per_group = pl.rows_per_group; p.box_rows = pl.box_rows; p.row_offset = off; p.n_rows = n_rows;
        const int w_bytes = (4 * H / 64) * pl.U * 128;
        const int smem = tc::LSTM_MAX_DYN;
        lstm_ring(w_bytes, pl.box_rows, pl.MT, &p.stage_bytes, &p.stages);
        const CUtensorMap& mx = md;
        const bool trace = getenv("FSMG_TRACE") != nullptr && c.trace != nullptr;
        if (trace) { cudaMemsetAsync(c.trace, 0, 8 * 16 * sizeof(long long), s); p.trace = c.trace; }
        if (bsplit) {
            p.box_rows = sub_rows;
            p.stage_bytes = bs_stage;
            p.stages = bs_stages;
            p.ks = ks;
            p.box_pitch = bs_pitch;
            if (half_m && pl.U == 32 && nsub == 2) rc = lstm_launch(tc::lstm_bwd_pair2_kernel<32, 2, true>, G * pl.C, tc::LSTM_BSPLIT_THREADS, 2, smem, mw, md, p, s);
            else if (half_m && nsub == 2) rc = lstm_launch(tc::lstm_bwd_pair2_kernel<16, 2, true>, G * pl.C, tc::LSTM_BSPLIT_THREADS, 2, smem, mw, md, p, s);
            else if (half_m && pl.U == 32) rc = lstm_launch(tc::lstm_bwd_pair2_kernel<32, 1, true>, G * pl.C, tc::LSTM_BSPLIT_THREADS, 2, smem, mw, md, p, s);
            else if (half_m) rc = lstm_launch(tc::lstm_bwd_pair2_kernel<16, 1, true>, G * pl.C, tc::LSTM_BSPLIT_THREADS, 2, smem, mw, md, p, s);
            else if (pl.U == 32 && nsub == 1) rc = lstm_launch(tc::lstm_bwd_pair2_kernel<32, 1>, G * pl.C, tc::LSTM_BSPLIT_THREADS, 2, smem, mw, md, p, s);
            else if (pl.U == 32) rc = lstm_launch(tc::lstm_bwd_pair2_kernel<32, 2>, G * pl.C, tc::LSTM_BSPLIT_THREADS, 2, smem, mw, md, p, s);
            else if (nsub == 1) rc = lstm_launch(tc::lstm_bwd_pair2_kernel<16, 1>, G * pl.C, tc::LSTM_BSPLIT_THREADS, 2, smem, mw, md, p, s);
            else rc = lstm_launch(tc::lstm_bwd_pair2_kernel<16, 2>, G * pl.C, tc::LSTM_BSPLIT_THREADS, 2, smem, mw, md, p, s);
        } else {
            FSMG_LSTM_DISPATCH(lstm_bwd_persistent_kernel, rc, true);
        }
        if (trace && !rc) lstm_trace_dump(c, bsplit ? (half_m ? (nsub == 2 ? "lstm_bwd_pair2<NSUB=2, M=128>" : "lstm_bwd_pair2<NSUB=1, M=128>") : nsub == 2 ? "lstm_bwd_pair2<NSUB=2>" : "lstm_bwd_pair2<NSUB=1>") : "lstm_bwd_persistent", s);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace fsmg
