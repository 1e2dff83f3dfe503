// engine.cu — libfsmg: handle, workspace layout, orchestration of the episodic-LSTM hot path and
// the C-ABI declared in include/fsmg.h.
//
// Replaces (reference file:line): LSTMBaseline._build_graph + train/eval/sample
// (src/models/lstm_baseline.py:38-156) and the TensorFlow-1.x runtime underneath it.
// Data layout in HBM: token index is TIME-MAJOR, r = t*N + n, so that every
// recurrent step touches one contiguous [N, .] block.
#include <cuda.h>
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "../../include/fsmg.h"
#include "common.cuh"
#include "simt_kernels.cuh"
#include "tc_gemm.cuh"
#include "tc_lstm.cuh"
#include "unigram.cuh"
#include "probe.cuh"

namespace fsmg {

std::string& last_error() {
    static thread_local std::string e;
    return e;
}
int set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

struct LayerBuf {
    int in = 0, inp = 0;          // input width and its padded leading dimension
    int64_t k_off = 0, b_off = 0; // offsets into the flat parameter buffer
    __half* K16 = nullptr;        // [(in+H), G4p]  row-major copy of kernel   (B of dX / dh_rec)
    __half* WxT16 = nullptr;      // [4H, inp]      transpose of kernel[:in]   (B of the input GEMM)
    __half* WhT16 = nullptr;      // [4H, Hp]       transpose of kernel[in:]   (B of the recurrent GEMM)
    __half* gates = nullptr;      // [NT, G4p] post-activation i,j,f,o (stash for BPTT)
    float* c = nullptr;           // [NT, H]
    __half* hs = nullptr;         // [NT, Hp]
};

}  // namespace fsmg

using namespace fsmg;

enum ProfPhase { PH_PREP = 0, PH_INPUT_GEMM, PH_REC_FWD, PH_PROJ_FWD, PH_SOFTMAX_GRAD, PH_DH, PH_DWS, PH_REC_BWD, PH_WGRAD, PH_DX, PH_UPDATE, PH_COUNT };
static const char* kPhaseNames[PH_COUNT] = {"prep_gather", "input_gemm", "recurrent_fwd", "proj_logits_lse", "softmax_grad_bias", "proj_dh",
                                            "proj_dws", "recurrent_bwd", "wgrad_kernel_bias", "dx_scatter", "clip_adam_refresh"};
struct ProfState {
    bool on = false;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    struct Rec { int ph; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    cudaEvent_t get() {
        if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        return pool[used++];
    }
};

struct fsmg_handle {
    fsmg_config cfg;
    ProfState prof;
    std::string scope;
    int V = 0, V1 = 0, E = 0, H = 0, L = 0, T = 0, Nmax = 0;
    int Ep = 0, Hp = 0, G4 = 0, G4p = 0, Vp = 0;
    std::vector<fsmg_param_info> infos;
    int64_t n_params = 0;       // padded flat count
    int64_t emb_off = 0, sw_off = 0, sb_off = 0;
    int64_t dense_begin = 0;    // first element after the embedding segment
    std::vector<LayerBuf> layers;
    // bound buffers
    float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr;
    char* ws = nullptr;
    int64_t ws_bytes = 0, ws_need = 0;
    bool bound = false;
    // workspace carve-up
    int32_t *x_ids = nullptr, *y_ids = nullptr, *tok_stage = nullptr, *tok_graph = nullptr, *samp_ids = nullptr, *samp_out = nullptr, *samp_out2 = nullptr;
    __half *emb16 = nullptr, *Ws16 = nullptr, *WsT16 = nullptr, *xemb = nullptr, *dgates = nullptr, *dlogits = nullptr;
    __half* pre16 = nullptr;   // hoisted x*Wx+b, fp16 [NT, G4p]
    __half* pre_tab = nullptr; // embedding*Wx+b for every word, fp16 [V', G4p]
    int pre_table = 1;         // FSMG_PRE_TABLE=0: always the per-token GEMM
    int32_t *tok_counts = nullptr, *sorted_rows = nullptr, *sorted_tok = nullptr;
    float* seg32 = nullptr;
    __half* seg16 = nullptr;
    int seg_grad = 1;          // FSMG_SEG_GRAD=0: per-token weight-gradient GEMM + scatter epilogue for the layer-0 input side
    float *gbuf = nullptr, *dact[2] = {nullptr, nullptr}, *dh_rec = nullptr, *dc_next = nullptr, *logits32 = nullptr;
    float *lse = nullptr, *nll = nullptr, *scalars = nullptr, *dws_acc = nullptr;
    float *s_x = nullptr, *s_g = nullptr, *s_logits = nullptr;
    std::vector<float*> s_c, s_h;
    // sampler v2 (split-fp16 tensor-core contractions, fp32-grade)
    __half *emb3 = nullptr, *WsT3 = nullptr;
    std::vector<__half*> KxT3, KhT3, h3;
    float *P0 = nullptr, *s_logits2 = nullptr;
    int* s_step = nullptr;
    bool samp_stale = true;
    cudaGraphExec_t samp_graph = nullptr;   // one captured decode step
    cudaGraphExec_t samp_graph_multi = nullptr;   // 16 consecutive decode steps
    int64_t samp_step_launches = 0;
    int samp_graph_n = 0;
    int chunk_rows = 0;
    // projection backward overlap: dH / dWs GEMMs of chunk i run on two auxiliary streams while the logits GEMM of
    // chunk i+1 runs on the caller's stream (double-buffered dlogits); fills the tail waves of the persistent GEMMs
    __half* dlogits_b[2] = {nullptr, nullptr};
    cudaStream_t aux[2] = {nullptr, nullptr};
    cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_dh[2] = {nullptr, nullptr}, ev_dws[2] = {nullptr, nullptr};
    cudaEvent_t ev_sort_fork = nullptr, ev_sort_join = nullptr;   // token sort of the layer-0 segment sums runs beside the forward pass
    bool sort_forked = false;
    int overlap = 1;
    int dws_transposed = 1;  // softmax_w gradient accumulated as [V', H] (FSMG_DWS_T=0: [H, V'])
    int strip_overlap = 0;   // background softmax-gradient pass beside the dH / dWs GEMMs (FSMG_STRIP_OVERLAP=1; measured slower)
    int fused_sg = 0;        // softmax gradient rebuilt inside the dH / dWs GEMMs from stored exponentials (FSMG_FUSED_SG; no HBM pass)
    float* cmaxT = nullptr;  // [ceil(V'/16), chunk_rows] per-(16-column chunk, row) maxima of the logits chunk (fused softmax gradient)
    int samp_max = 0;
    // pinned host staging
    int32_t* h_tok = nullptr;
    float* h_scal = nullptr;
    int64_t h_tok_elems = 0;
    int64_t launches = 0;
    // whole forward+backward pass captured once per (n_seqs, loss scale, nll buffer) over a fixed token buffer and replayed: ~740 kernel
    // nodes become ONE launch, so the host cannot fall behind the device (the e2e arm synchronises on the loss every step)
    struct StepGraph { int32_t n; uint32_t ls_bits; float* nll; cudaGraphExec_t exec; int64_t launches; int seen; int failed; };
    std::vector<StepGraph> step_graphs;
    int use_graph = 1;
    // caller-owned events recorded INSIDE forward_backward (fsmg_set_stage_events): [0] softmax_w / softmax_b gradients final (after the
    // projection backward), [1] embedding gradient final.  A data-parallel caller all-reduces those slices on a side stream while the
    // recurrent backward still runs.
    cudaEvent_t stage_ev[3] = {nullptr, nullptr, nullptr};   // [2]: sum of the per-token NLL final (fsmg_set_loss_event)
    fsmg::TcContext tc;
};

namespace fsmg {

#define LAUNCH_COUNT(h) ((h)->launches++)

// brackets a phase with CUDA events on the launching stream when profiling is enabled
struct ProfScope {
    fsmg_handle* h; cudaStream_t s; int ph; cudaEvent_t a = nullptr;
    ProfScope(fsmg_handle* h_, int ph_, cudaStream_t s_) : h(h_), s(s_), ph(ph_) {
        if (h->prof.on) { a = h->prof.get(); cudaEventRecord(a, s); }
    }
    ~ProfScope() {
        if (a) { cudaEvent_t b = h->prof.get(); cudaEventRecord(b, s); h->prof.recs.push_back({ph, a, b}); }
    }
};

// records a caller-owned stage event on s; inside a stream capture it becomes an external event-record node of the graph
static int record_stage_event(fsmg_handle* h, int which, cudaStream_t s) {
    cudaEvent_t ev = h->stage_ev[which];
    if (!ev) return FSMG_OK;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    FSMG_CUDA_OK(cudaStreamIsCapturing(s, &st));
    FSMG_CUDA_OK(cudaEventRecordWithFlags(ev, s, st == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault));
    return FSMG_OK;
}

// bump allocator used twice: sizing (base == nullptr) and carving
struct Bump {
    char* base;
    int64_t off = 0;
    template <typename T>
    T* take(int64_t count) {
        off = round_up(off, 256);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * (int64_t)sizeof(T);
        return p;
    }
};

static void carve(fsmg_handle* h, char* base) {
    Bump b{base};
    const int64_t NT = (int64_t)h->Nmax * h->T;
    h->scalars = b.take<float>(512);   // [0..7] results, [32] token-range error count (int), [64..64+SQNORM_BLOCKS) norm partials
    h->x_ids = b.take<int32_t>(NT);
    h->y_ids = b.take<int32_t>(NT);
    h->tok_stage = b.take<int32_t>(NT);
    h->tok_graph = b.take<int32_t>(NT);
    h->emb16 = b.take<__half>((int64_t)h->V1 * h->Ep);
    h->Ws16 = b.take<__half>((int64_t)h->H * h->Vp);
    h->WsT16 = b.take<__half>((int64_t)h->V1 * h->Hp);
    h->xemb = b.take<__half>(NT * h->Ep);
    h->pre16 = b.take<__half>(NT * h->G4p);
    h->pre_tab = b.take<__half>((int64_t)h->V1 * h->G4p);   // per-word pre-activation table of layer 0 (see forward_lstm)
    // token-sorted segment sums of dgates (layer-0 input gradients, see backward_lstm)
    h->tok_counts = b.take<int32_t>(3 * ((int64_t)h->V1 + 8));
    h->sorted_rows = b.take<int32_t>(NT);
    h->sorted_tok = b.take<int32_t>(NT);
    h->seg32 = b.take<float>((int64_t)h->V1 * h->G4);
    h->seg16 = b.take<__half>((int64_t)h->V1 * h->G4p);
    h->gbuf = b.take<float>((int64_t)h->Nmax * h->G4);   // per-step route: recurrent contraction of one step
    h->dgates = b.take<__half>(NT * h->G4p);
    int wmax = h->E > h->H ? h->E : h->H;
    h->dact[0] = b.take<float>(NT * wmax);
    h->dact[1] = b.take<float>(NT * wmax);
    h->dh_rec = b.take<float>((int64_t)h->Nmax * h->H);
    h->dc_next = b.take<float>((int64_t)h->Nmax * h->H);
    h->lse = b.take<float>(NT);
    h->nll = b.take<float>(NT);
    for (auto& l : h->layers) {
        l.K16 = b.take<__half>((int64_t)(l.in + h->H) * h->G4p);
        l.WxT16 = b.take<__half>((int64_t)h->G4 * l.inp);
        l.WhT16 = b.take<__half>((int64_t)h->G4 * h->Hp);
        l.gates = b.take<__half>(NT * h->G4p);
        l.c = b.take<float>(NT * h->H);
        l.hs = b.take<__half>(NT * h->Hp);
    }
    // projection chunk: rows sized so the fp16 logits chunk stays L2-resident (<= ~48 MB)
    const char* env_mb = getenv("FSMG_CHUNK_MB");
    const char* env_ov = getenv("FSMG_OVERLAP");
    const char* env_gr = getenv("FSMG_GRAPH");
    const char* env_sg = getenv("FSMG_SEG_GRAD");
    if (env_sg) h->seg_grad = atoi(env_sg);
    const char* env_pt = getenv("FSMG_PRE_TABLE");
    if (env_pt) h->pre_table = atoi(env_pt);
    const char* env_dt = getenv("FSMG_DWS_T");
    if (env_dt) h->dws_transposed = atoi(env_dt);
    const char* env_so = getenv("FSMG_STRIP_OVERLAP");
    { const char* env_fs = getenv("FSMG_FUSED_SG"); h->fused_sg = env_fs ? atoi(env_fs) : 1; }
    h->strip_overlap = env_so ? atoi(env_so) : 0;   // measured: 14.17 ms (1 background CTA/SM: 1.2 TB/s) / 12.65 (6/SM) vs 12.47 serial
    h->use_graph = env_gr ? atoi(env_gr) : 1;
    h->overlap = env_ov ? atoi(env_ov) : 0;   // measured: with 256 MB chunks and stream-K balanced GEMMs, overlapping streams lose (16.97 vs 14.35 ms)
    // Chunk rows: the dH GEMM of a chunk has ONE 512-wide N tile, so its tile count is rows/256 (cta_group::2 pair tiles).  Pick the
    // number of chunks so that a chunk is at most one full wave of pair tiles (74 on 148 SMs) and divide the tokens EVENLY: no ragged
    // last chunk, no stream-K partial sums in dH (measured at cfg 2: 14 chunks of 13 312 rows 12.47 ms -> 10 chunks of 18 432 rows
    // 12.04 ms, dH 1.85 -> 1.40 ms).  FSMG_CHUNK_MB caps the fp16 logits chunk (default 1 GB); FSMG_CHUNK_ROWS overrides.
    int n_sm = 148;
    { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
    const int64_t wave_rows = (int64_t)(n_sm / 2) * 256;
    const int64_t n_chunks = NT > 0 ? cdiv(NT, wave_rows) : 1;
    int64_t rows = round_up(cdiv(NT, n_chunks), 256);
    const int64_t chunk_mb = env_mb ? atoi(env_mb) : 1024;
    const int64_t cap_rows = ((chunk_mb << 20) / ((int64_t)h->Vp * 2)) / 128 * 128;
    if (rows > cap_rows) rows = cap_rows;
    const char* env_cr = getenv("FSMG_CHUNK_ROWS");   // explicit row count (tile-count experiments)
    if (env_cr && atoll(env_cr) > 0) rows = atoll(env_cr);
    if (rows < 128) rows = 128;
    if (rows > NT) rows = round_up(NT, 128);
    h->chunk_rows = (int)rows;
    h->logits32 = b.take<float>(rows * h->Vp);
    h->dlogits = b.take<__half>(rows * h->Vp);
    h->dlogits_b[0] = h->dlogits;
    h->dlogits_b[1] = b.take<__half>(rows * h->Vp);
    h->cmaxT = b.take<float>((int64_t)cdiv(h->V1, 16) * rows);
    // dWs accumulated with 16-byte aligned rows (V' is odd): [H, Vp], or transposed [V', Hq] (default, see projection())
    { int64_t a = (int64_t)h->H * h->Vp, bt = (int64_t)h->V1 * round_up(h->H, 4); h->dws_acc = b.take<float>(a > bt ? a : bt); }
    // sampler (fp32 route)
    h->samp_max = h->Nmax;
    h->samp_ids = b.take<int32_t>(h->samp_max);
    h->samp_out = b.take<int32_t>((int64_t)h->samp_max * 4096);
    h->samp_out2 = b.take<int32_t>((int64_t)h->samp_max * 4096);
    h->s_x = b.take<float>((int64_t)h->samp_max * wmax);
    h->s_g = b.take<float>((int64_t)h->samp_max * h->G4);
    h->s_logits = b.take<float>((int64_t)h->samp_max * h->V1);
    h->s_c.resize(h->L);
    h->s_h.resize(h->L);
    for (int l = 0; l < h->L; ++l) {
        h->s_c[l] = b.take<float>((int64_t)h->samp_max * h->H);
        h->s_h[l] = b.take<float>((int64_t)h->samp_max * h->H);
    }
    h->emb3 = b.take<__half>((int64_t)h->V1 * 3 * h->Ep);
    h->WsT3 = b.take<__half>((int64_t)h->V1 * 3 * h->Hp);
    h->P0 = b.take<float>((int64_t)h->V1 * h->G4);
    h->s_logits2 = b.take<float>((int64_t)h->samp_max * h->Vp);
    h->s_step = b.take<int>(64);
    h->KxT3.resize(h->L); h->KhT3.resize(h->L); h->h3.resize(h->L);
    for (int l = 0; l < h->L; ++l) {
        h->KxT3[l] = b.take<__half>((int64_t)h->G4 * 3 * h->layers[l].inp);
        h->KhT3[l] = b.take<__half>((int64_t)h->G4 * 3 * h->Hp);
        h->h3[l] = b.take<__half>((int64_t)h->samp_max * 3 * h->Hp);
    }
    tc_carve(h->tc, b, h->Nmax, h->T, h->V1, h->Vp, h->H, h->chunk_rows);
    h->ws_need = round_up(b.off, 256);
}

// ---- GEMM dispatch: tcgen05 route unless the debug flag (or an unsupported shape) says SIMT -------
static int gemm_f16(fsmg_handle* h, const GemmArgs& g, bool a_mn, bool b_mn, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0) return FSMG_OK;
    if (g.K <= 0) return FSMG_OK;
    if (!(h->cfg.flags & FSMG_FLAG_SIMT_GEMM) && tc_gemm_supported(g, a_mn, b_mn)) {
        LAUNCH_COUNT(h);
        return tc_gemm(h->tc, g, a_mn, b_mn, s);
    }
    launch_simt_gemm<__half, __half>(g, a_mn, b_mn, s);
    LAUNCH_COUNT(h);
    FSMG_LAUNCH_OK();
    return FSMG_OK;
}

static GemmArgs mk(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                   float alpha = 1.0f, const float* bias = nullptr, int c_half = 0, int accumulate = 0, int atomic = 0) {
    GemmArgs g;
    g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
    g.bias = bias; g.alpha = alpha; g.c_half = c_half; g.accumulate = accumulate; g.atomic = atomic;
    return g;
}

static int refresh_weights(fsmg_handle* h, cudaStream_t s) {
    const int TB = 256;
    h->samp_stale = true;   // the sampler's split-fp16 operand copies are rebuilt lazily
    auto conv = [&](const float* in, int64_t ldi, __half* out, int64_t ldo, int rows, int cols) {
        int64_t total = (int64_t)rows * ldo;
        convert_f16_kernel<<<cdiv(total, TB), TB, 0, s>>>(in, ldi, out, ldo, rows, cols);
        LAUNCH_COUNT(h);
    };
    auto tr = [&](const float* in, int64_t ldi, __half* out, int64_t ldo, int rows, int cols) {
        dim3 grid(cdiv(cols, 32), cdiv(ldo, 32));
        transpose_f16_kernel<<<grid, dim3(32, 8), 0, s>>>(in, ldi, out, ldo, rows, cols);
        LAUNCH_COUNT(h);
    };
    conv(h->params + h->emb_off, h->E, h->emb16, h->Ep, h->V1, h->E);
    for (auto& l : h->layers) {
        const float* K = h->params + l.k_off;
        conv(K, h->G4, l.K16, h->G4p, l.in + h->H, h->G4);
        tr(K, h->G4, l.WxT16, l.inp, l.in, h->G4);                          // [in,4H] -> [4H, inp]
        tr(K + (int64_t)l.in * h->G4, h->G4, l.WhT16, h->Hp, h->H, h->G4);  // [H,4H]  -> [4H, Hp]
    }
    conv(h->params + h->sw_off, h->V1, h->Ws16, h->Vp, h->H, h->V1);
    tr(h->params + h->sw_off, h->V1, h->WsT16, h->Hp, h->H, h->V1);         // [H,V'] -> [V', Hp]
    FSMG_LAUNCH_OK();
    return FSMG_OK;
}

// ---- forward through the LSTM stack (embedding -> layers), stashing what BPTT needs -----------
// layer 0 handled per WORD instead of per token (pre-activation table forward, segment sums backward)?
static bool layer0_word_forward(fsmg_handle* h, int N) {
    return !(h->cfg.flags & (FSMG_FLAG_SIMT_RECURRENT | FSMG_FLAG_SIMT_GEMM)) && tc_recurrent_supported(h->tc, N, h->H) && h->pre_table &&
           (int64_t)h->V1 < (int64_t)N * h->T;
}
static bool layer0_word_backward(fsmg_handle* h, int N) {
    return !(h->cfg.flags & (FSMG_FLAG_SIMT_RECURRENT | FSMG_FLAG_SIMT_GEMM)) && tc_recurrent_supported(h->tc, N, h->H) && h->seg_grad &&
           (int64_t)h->V1 < (int64_t)N * h->T && h->Ep == h->E;
}

// counting sort of the token rows by input word (layer-0 segment sums, see backward_lstm): depends on the tokens only
static int token_sort(fsmg_handle* h, int N, cudaStream_t s) {
    const int TB = 256, V1 = h->V1;
    const int64_t NT = (int64_t)N * h->T;
    int32_t* counts = h->tok_counts;
    int32_t* offsets = counts + (V1 + 8);
    int32_t* cursor = offsets + (V1 + 8);
    FSMG_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)(V1 + 8), s));
    token_hist_kernel<<<cdiv(NT, TB), TB, 0, s>>>(h->x_ids, NT, counts);
    token_scan_kernel<<<1, 1024, 0, s>>>(counts, V1, offsets, cursor);
    token_fill_kernel<<<cdiv(NT, TB), TB, 0, s>>>(h->x_ids, NT, cursor, h->sorted_rows, h->sorted_tok);
    h->launches += 3;
    FSMG_LAUNCH_OK();
    return FSMG_OK;
}

static bool layer0_word_backward(fsmg_handle* h, int N);

static int forward_lstm(fsmg_handle* h, const int32_t* d_tokens, int N, bool train, cudaStream_t s) {
    const int T = h->T, H = h->H, TB = 256;
    const int64_t NT = (int64_t)N * T;
    {
    ProfScope ps(h, PH_PREP, s);
    prep_tokens_kernel<<<cdiv(NT, TB), TB, 0, s>>>(d_tokens, h->x_ids, h->y_ids, N, T, h->V, h->V, reinterpret_cast<int*>(h->scalars + 32),
                                                   train && h->grads ? h->grads + h->n_params + 2 : nullptr);
    LAUNCH_COUNT(h);
        // the gathered embedding rows are only needed by the per-token input GEMM (forward) / weight-gradient GEMM (backward)
        if (!layer0_word_forward(h, N) || (train && !layer0_word_backward(h, N))) {
            int cols8 = h->Ep / 8;
            gather_rows_f16_kernel<<<cdiv(NT * cols8, TB), TB, 0, s>>>(h->emb16, h->Ep, h->x_ids, h->xemb, h->Ep, NT, cols8);
            LAUNCH_COUNT(h);
        }
    }
    FSMG_LAUNCH_OK();
    // the token sort the backward pass needs depends on x_ids only: fork it onto an auxiliary stream here (also inside a graph
    // capture: a cross-stream fork/join) so that its three small kernels run beside the forward pass instead of after the recurrent backward
    h->sort_forked = false;
    if (train && !h->prof.on && h->aux[0] != nullptr && layer0_word_backward(h, N)) {
        FSMG_CUDA_OK(cudaEventRecord(h->ev_sort_fork, s));
        FSMG_CUDA_OK(cudaStreamWaitEvent(h->aux[0], h->ev_sort_fork, 0));
        int rc0 = token_sort(h, N, h->aux[0]);
        if (rc0) return rc0;
        FSMG_CUDA_OK(cudaEventRecord(h->ev_sort_join, h->aux[0]));
        h->sort_forked = true;
    }
    for (int li = 0; li < h->L; ++li) {
        LayerBuf& l = h->layers[li];
        const __half* in = li == 0 ? h->xemb : h->layers[li - 1].hs;
        const float* bias = h->params + l.b_off;
        // hoisted input contraction (K3 in SURVEY §2.1): pre[NT,4H] = in[NT,in] * Wx + b — or, for layer 0 on the persistent route
        // when the batch holds more tokens than the vocabulary has words, the per-WORD table P[V',4H] = embedding * Wx + b (one small
        // GEMM: 10 001 rows instead of 184 320 at cfg 2) that the recurrent kernel reads through the input ids.  Same operands, same
        // K order, same fp16 rounding: bit-identical pre-activations, without the [NT,4H] write + read and 95 % of the GEMM.
        const bool persistent_fwd = !(h->cfg.flags & FSMG_FLAG_SIMT_RECURRENT) && !(h->cfg.flags & FSMG_FLAG_SIMT_GEMM) &&
                                    tc_recurrent_supported(h->tc, N, H);
        const bool word_table = li == 0 && layer0_word_forward(h, N);
        int rc;
        {
            ProfScope ps(h, PH_INPUT_GEMM, s);
            if (word_table)
                rc = gemm_f16(h, mk(h->V1, h->G4, l.in, h->emb16, h->Ep, l.WxT16, l.inp, h->pre_tab, h->G4p, 1.0f, bias, /*c_half=*/1), false, false, s);
            else
                rc = gemm_f16(h, mk((int)NT, h->G4, l.in, in, l.inp, l.WxT16, l.inp, h->pre16, h->G4p, 1.0f, bias, /*c_half=*/1), false, false, s);
        }
        if (rc) return rc;
        ProfScope ps_rec(h, PH_REC_FWD, s);
        if (persistent_fwd) {
            rc = tc_lstm_forward(h->tc, word_table ? h->pre_tab : h->pre16, word_table ? h->x_ids : nullptr, l.WhT16, l.gates, l.c, l.hs, N, T, H,
                                 h->Hp, h->G4p, s);
            LAUNCH_COUNT(h);
            if (rc) return rc;
            continue;
        }
        for (int t = 0; t < T; ++t) {
            if (t > 0) {
                rc = gemm_f16(h, mk(N, h->G4, H, l.hs + (int64_t)(t - 1) * N * h->Hp, h->Hp, l.WhT16, h->Hp, h->gbuf, h->G4), false, false, s);
                if (rc) return rc;
            }
            lstm_pointwise_fwd_kernel<__half><<<cdiv((int64_t)N * H, TB), TB, 0, s>>>(
                t ? h->gbuf : nullptr, h->G4, h->pre16 + (int64_t)t * N * h->G4p, h->G4p, t ? l.c + (int64_t)(t - 1) * N * H : nullptr,
                l.gates + (int64_t)t * N * h->G4p, h->G4p, l.c + (int64_t)t * N * H, l.hs + (int64_t)t * N * h->Hp, h->Hp, N, H);
            LAUNCH_COUNT(h);
        }
        FSMG_LAUNCH_OK();
    }
    return FSMG_OK;
}

// ---- projection + softmax/NLL (+ its backward when train) over L2-sized token chunks ------------
static int projection(fsmg_handle* h, int N, bool train, float loss_scale, float* d_nll_user, cudaStream_t s) {
    const int T = h->T, H = h->H;
    const int64_t NT = (int64_t)N * T;
    const __half* hs = h->layers[h->L - 1].hs;
    const float* sb = h->params + h->sb_off;
    float* g_sw = h->grads ? h->grads + h->sw_off : nullptr;
    float* g_sb = h->grads ? h->grads + h->sb_off : nullptr;
    float* nll_out = d_nll_user ? d_nll_user : h->nll;
    const bool use_tc = !(h->cfg.flags & FSMG_FLAG_SIMT_GEMM) && tc_projection_supported(h->tc, H, h->V1);
    // dWs: computed TRANSPOSED by default, dWs^T[V', H] += dlogits^T * hs.  With H <= 512 that GEMM has a single (256 x 512 pair) N
    // tile, so every dlogits element is fetched exactly once; the [H, V'] orientation read the chunk once per 256 rows of H (ncu: 766 MB
    // of DRAM reads per 369 MB chunk, 69 % DRAM utilisation — the kernel had become HBM-bound).  One fp32 transpose per step undoes it.
    const int Hq = (int)round_up(H, 4);
    const bool dws_t = h->dws_transposed != 0;
    if (train) FSMG_CUDA_OK(cudaMemsetAsync(h->dws_acc, 0, sizeof(float) * (size_t)(dws_t ? (int64_t)h->V1 * Hq : (int64_t)H * h->Vp), s));
    auto dws_gemm = [&](const __half* hc_, const __half* dl_, int mc_, cudaStream_t st) {
        if (dws_t) return gemm_f16(h, mk(h->V1, H, mc_, dl_, h->Vp, hc_, h->Hp, h->dws_acc, Hq, loss_scale, nullptr, 0, 0, 1), true, true, st);
        return gemm_f16(h, mk(H, h->V1, mc_, hc_, h->Hp, dl_, h->Vp, h->dws_acc, h->Vp, loss_scale, nullptr, 0, 0, 1), true, true, st);
    };
    auto dws_finish = [&](cudaStream_t st) {
        if (dws_t) {
            dim3 grid(cdiv(H, 32), cdiv(h->V1, 32));
            transpose_f32_kernel<<<grid, dim3(32, 8), 0, st>>>(h->dws_acc, Hq, g_sw, h->V1, h->V1, H);
        } else {
            int64_t total = (int64_t)H * h->V1;
            unpad_rows_kernel<<<cdiv(total, 256), 256, 0, st>>>(h->dws_acc, h->Vp, g_sw, h->V1, H, h->V1);
        }
        LAUNCH_COUNT(h);
    };
    // overlap is switched off while profiling so that per-phase event brackets do not double count
    const bool overlap = train && use_tc && h->overlap && !h->prof.on && h->aux[0] != nullptr;
    // Optional training schedule (FSMG_STRIP_OVERLAP=1) on the tcgen05 route: the HBM-bound softmax-gradient pass of chunk i runs as a background kernel
    // on a side stream, sharing the SMs with the tensor-bound dH / dWs GEMMs of chunk i-1 (main stream, after the logits GEMM of
    // chunk i); dlogits is double-buffered.  Per-phase profiling uses the serial schedule below (brackets would overlap).
    const int64_t n_chunks = cdiv(NT, (int64_t)h->chunk_rows);
    if (train && use_tc && !overlap && h->strip_overlap && !h->prof.on && h->aux[0] != nullptr && n_chunks >= 2) {
        cudaStream_t sx = h->aux[0];
        // dH is zeroed once for the whole step and every chunk's GEMM adds into it (stream-K partials are combined with REDs
        // anyway): no memset node between the GEMMs while the background pass is resident
        FSMG_CUDA_OK(cudaMemsetAsync(h->dact[0], 0, sizeof(float) * (size_t)NT * H, s));
        for (int64_t i = 0; i <= n_chunks; ++i) {
            int rc;
            if (i < n_chunks) {
                const int64_t r0 = i * h->chunk_rows;
                const int mc = (int)((NT - r0 < h->chunk_rows) ? NT - r0 : h->chunk_rows);
                const int buf = (int)(i & 1);
                int n_part = 0;
                // buffer `buf` was last read by the dH / dWs GEMMs of chunk i-2, already enqueued on s
                rc = tc_projection_gemm(h->tc, hs + r0 * h->Hp, h->Hp, h->WsT16, h->Hp, sb, h->y_ids, r0, mc, H, h->V1, h->dlogits_b[buf],
                                        h->Vp, &n_part, s);
                if (rc) return rc;
                rc = tc_projection_combine(h->tc, n_part, r0, mc, N, T, h->lse, nll_out, s, /*reset_sched=*/true);
                if (rc) return rc;
                FSMG_CUDA_OK(cudaEventRecord(h->ev_ready[buf], s));
                FSMG_CUDA_OK(cudaStreamWaitEvent(sx, h->ev_ready[buf], 0));
                rc = tc_projection_strip_bg(h->tc, h->y_ids, r0, mc, h->V1, h->dlogits_b[buf], h->Vp, h->lse, loss_scale, g_sb, sx);
                if (rc) return rc;
                FSMG_CUDA_OK(cudaEventRecord(h->ev_dh[buf], sx));
                h->launches += 3;
            }
            if (i >= 1) {
                const int64_t j = i - 1, r0 = j * h->chunk_rows;
                const int mc = (int)((NT - r0 < h->chunk_rows) ? NT - r0 : h->chunk_rows);
                const int buf = (int)(j & 1);
                FSMG_CUDA_OK(cudaStreamWaitEvent(s, h->ev_dh[buf], 0));     // dlogits of chunk j are final
                rc = gemm_f16(h, mk(mc, H, h->V1, h->dlogits_b[buf], h->Vp, h->Ws16, h->Vp, h->dact[0] + r0 * H, H, 1.0f, nullptr, 0, 0, 1), false, false, s);
                if (rc) return rc;
                rc = dws_gemm(hs + r0 * h->Hp, h->dlogits_b[buf], mc, s);
                if (rc) return rc;
            }
        }
        dws_finish(s);
        FSMG_LAUNCH_OK();
        return FSMG_OK;
    }
    int64_t chunk_idx = 0;
    for (int64_t r0 = 0; r0 < NT; r0 += h->chunk_rows, ++chunk_idx) {
        int mc = (int)((NT - r0 < h->chunk_rows) ? NT - r0 : h->chunk_rows);
        const __half* hc = hs + r0 * h->Hp;
        const int buf = overlap ? (int)(chunk_idx & 1) : 0;
        h->dlogits = h->dlogits_b[buf];
        cudaStream_t s_dh = overlap ? h->aux[0] : s, s_dws = overlap ? h->aux[1] : s;
        if (overlap && chunk_idx >= 2) {   // the buffer is free once both consumers of chunk i-2 are done
            FSMG_CUDA_OK(cudaStreamWaitEvent(s, h->ev_dh[buf], 0));
            FSMG_CUDA_OK(cudaStreamWaitEvent(s, h->ev_dws[buf], 0));
        }
        int rc;
        // Fused softmax gradient (FSMG_FUSED_SG): the logits GEMM leaves e = exp(logit - chunk max) in the fp16 chunk and the two
        // consumer GEMMs rebuild softmax - onehot on the tiles TMA lands in their shared memory (tc_gemm_kernel "XF"): the in-place
        // HBM pass over the chunk and its 4 B per logit of traffic disappear.  Needs the 256 x 512 pair-tile plan for both GEMMs.
        const GemmArgs g_dh = mk(mc, H, h->V1, h->dlogits, h->Vp, h->Ws16, h->Vp, h->dact[0] + r0 * H, H);
        const GemmArgs g_dws = mk(h->V1, H, mc, h->dlogits, h->Vp, hc, h->Hp, h->dws_acc, Hq, loss_scale, nullptr, 0, 0, 1);
        const bool fused = train && use_tc && h->fused_sg && dws_t && !overlap && tc_xf_supported(h->tc, g_dh) && tc_xf_supported(h->tc, g_dws);
        XfArgs xf;
        xf.cmaxT = h->cmaxT; xf.ld_cmax = h->chunk_rows; xf.n_c16 = cdiv(h->V1, 16); xf.lse = h->lse; xf.y = h->y_ids; xf.row0 = r0; xf.db = nullptr;
        if (use_tc) {
            // fused: logits tile -> online (max,sumexp) partials + target logit; when training also the fp16 chunk for the backward
            // (exponentials + chunk maxima on the fused-softmax-gradient route, logits on the fallback route)
            int n_part = 0;
            {
                ProfScope ps(h, PH_PROJ_FWD, s);
                rc = tc_projection_gemm(h->tc, hc, h->Hp, h->WsT16, h->Hp, sb, h->y_ids, r0, mc, H, h->V1,
                                        train ? h->dlogits : nullptr, h->Vp, &n_part, s, fused ? h->cmaxT : nullptr, h->chunk_rows);
            }
            if (rc) return rc;
            {
                ProfScope ps(h, PH_SOFTMAX_GRAD, s);
                rc = tc_projection_post(h->tc, n_part, h->y_ids, r0, mc, N, T, h->V1, (train && !fused) ? h->dlogits : nullptr, h->Vp,
                                        h->lse, nll_out, loss_scale, g_sb, s);
            }
            h->launches += (train && !fused) ? 3 : 2;
            if (rc) return rc;
        } else {
            ProfScope ps(h, PH_PROJ_FWD, s);
            rc = gemm_f16(h, mk(mc, h->V1, H, hc, h->Hp, h->WsT16, h->Hp, h->logits32, h->Vp, 1.0f, sb), false, false, s);
            if (rc) return rc;
            rowwise_nll_kernel<<<mc, 256, 0, s>>>(h->logits32, h->Vp, h->V1, h->y_ids, r0, N, T, h->lse, nll_out,
                                                  train ? h->dlogits : nullptr, h->Vp);
            LAUNCH_COUNT(h);
            FSMG_LAUNCH_OK();
        }
        if (!train) continue;
        // db_s += loss_scale * colsum(dlogits)   (the tcgen05 route sums it in the dWs GEMM's operand transform, or in its softmax-grad pass)
        if (!use_tc) {
            ProfScope ps(h, PH_SOFTMAX_GRAD, s);
            int rpb = 64;
            dim3 grid(cdiv(h->V1, 128), cdiv(mc, rpb));
            colsum_f16_kernel<<<grid, 128, 0, s>>>(h->dlogits, h->Vp, mc, h->V1, loss_scale, g_sb, rpb);
            LAUNCH_COUNT(h);
        }
        if (overlap) {
            FSMG_CUDA_OK(cudaEventRecord(h->ev_ready[buf], s));
            FSMG_CUDA_OK(cudaStreamWaitEvent(s_dh, h->ev_ready[buf], 0));
            FSMG_CUDA_OK(cudaStreamWaitEvent(s_dws, h->ev_ready[buf], 0));
        }
        // dH[chunk] = dlogits * Ws^T   (unscaled; fp32)
        {
            ProfScope ps(h, PH_DH, s_dh);
            if (fused) { LAUNCH_COUNT(h); rc = tc_gemm(h->tc, g_dh, false, false, s_dh, &xf); }
            else rc = gemm_f16(h, g_dh, false, false, s_dh);
        }
        if (rc) return rc;
        // dWs += loss_scale * hs_chunk^T * dlogits   (contraction over the chunk's tokens)
        {
            ProfScope ps_dws(h, PH_DWS, s_dws);
            // accumulated across chunks with fire-and-forget vector reductions into the L2-resident buffer (no read latency in the epilogue)
            if (fused) { xf.db = g_sb; LAUNCH_COUNT(h); rc = tc_gemm(h->tc, g_dws, true, true, s_dws, &xf); }   // + bias gradient
            else rc = dws_gemm(hc, h->dlogits, mc, s_dws);
        }
        if (rc) return rc;
        if (overlap) {
            FSMG_CUDA_OK(cudaEventRecord(h->ev_dh[buf], s_dh));
            FSMG_CUDA_OK(cudaEventRecord(h->ev_dws[buf], s_dws));
        }
    }
    if (overlap) {   // join: everything after the projection (BPTT) is ordered behind both auxiliary streams
        for (int b2 = 0; b2 < 2 && b2 < chunk_idx; ++b2) {
            FSMG_CUDA_OK(cudaStreamWaitEvent(s, h->ev_dh[b2], 0));
            FSMG_CUDA_OK(cudaStreamWaitEvent(s, h->ev_dws[b2], 0));
        }
    }
    h->dlogits = h->dlogits_b[0];
    if (train) dws_finish(s);
    FSMG_LAUNCH_OK();
    return FSMG_OK;
}

static int backward_lstm(fsmg_handle* h, int N, float loss_scale, cudaStream_t s) {
    const int T = h->T, H = h->H, TB = 256;
    const int64_t NT = (int64_t)N * T;
    int cur = 0;  // dact[cur] holds dL/dh_t of the current layer ([NT,H] fp32)
    for (int li = h->L - 1; li >= 0; --li) {
        LayerBuf& l = h->layers[li];
        const __half* in = li == 0 ? h->xemb : h->layers[li - 1].hs;
        const __half* Wh_rows = l.K16 + (int64_t)l.in * h->G4p;  // kernel[in:, :] as [H, 4H] K-major
        float* dh_all = h->dact[cur];
        int rc;
        bool persistent = !(h->cfg.flags & FSMG_FLAG_SIMT_RECURRENT) && !(h->cfg.flags & FSMG_FLAG_SIMT_GEMM) &&
                          tc_recurrent_supported(h->tc, N, H);
        {
        ProfScope ps_rec(h, PH_REC_BWD, s);
        if (persistent) {
            rc = tc_lstm_backward(h->tc, dh_all, Wh_rows, l.gates, l.c, h->dgates, N, T, H, h->G4p, s);
            LAUNCH_COUNT(h);
            if (rc) return rc;
        } else {
            for (int t = T - 1; t >= 0; --t) {
                if (t < T - 1) {
                    rc = gemm_f16(h, mk(N, H, h->G4, h->dgates + (int64_t)(t + 1) * N * h->G4p, h->G4p, Wh_rows, h->G4p,
                                        h->dh_rec, H), false, false, s);
                    if (rc) return rc;
                }
                lstm_pointwise_bwd_kernel<<<cdiv((int64_t)N * H, TB), TB, 0, s>>>(
                    dh_all + (int64_t)t * N * H, H, t < T - 1 ? h->dh_rec : nullptr, l.gates + (int64_t)t * N * h->G4p,
                    h->G4p, l.c + (int64_t)t * N * H, t ? l.c + (int64_t)(t - 1) * N * H : nullptr, h->dc_next,
                    h->dgates + (int64_t)t * N * h->G4p, h->G4p, N, H, t == T - 1);
                LAUNCH_COUNT(h);
            }
            FSMG_LAUNCH_OK();
        }
        }
        // Layer-0 input side through token-sorted segment sums S[v,:] = sum_{x[r]=v} dgates[r,:] when the batch holds more tokens than
        // the vocabulary has words (persistent tcgen05 route): bias gradient = column sums of S (82 MB instead of 755 MB at cfg 2),
        // dK[:E] = embedding^T * S and dEmbedding = S * K[:E]^T over V' rows instead of N*T tokens; the per-token dX GEMM survives
        // only as the per-occurrence square norm of TF's clip — an epilogue without a single global write.
        const bool seg = li == 0 && persistent && layer0_word_backward(h, N);
        if (seg) {
            ProfScope ps_seg(h, PH_WGRAD, s);
            const int V1 = h->V1;
            if (h->sort_forked) {                 // sorted beside the forward pass: join
                FSMG_CUDA_OK(cudaStreamWaitEvent(s, h->ev_sort_join, 0));
                h->sort_forked = false;
            } else if ((rc = token_sort(h, N, s))) return rc;
            FSMG_CUDA_OK(cudaMemsetAsync(h->seg32, 0, sizeof(float) * (size_t)V1 * h->G4, s));
            constexpr int RPB = 64;
            segsum_rows_kernel<RPB><<<dim3(cdiv(h->G4, 1024), cdiv(NT, RPB)), 128, 0, s>>>(h->dgates, h->G4p, h->G4, h->sorted_tok, h->sorted_rows, NT,
                                                                                          h->seg32, h->G4);
            // fp16 operand copy of S + db = loss_scale * colsum(S)  (= colsum(dgates))
            seg_finish_kernel<64><<<dim3(cdiv(h->G4p, 1024), cdiv(V1, 64)), 128, 0, s>>>(h->seg32, h->G4, V1, h->G4, h->seg16, h->G4p, loss_scale,
                                                                                        h->grads + l.b_off);
            h->launches += 2;
            FSMG_LAUNCH_OK();
            float* gK0 = h->grads + l.k_off;
            // dEmbedding = loss_scale * S * K[:E]^T   (dense [V', E]; rows of words absent from the batch come out zero).  First of the three
            // GEMMs: the embedding gradient is the largest slice of the flat buffer, and a data-parallel caller starts its all-reduce from the
            // stage event while the two kernel-gradient GEMMs and the occurrence-norm pass below still run
            rc = gemm_f16(h, mk(V1, h->E, h->G4, h->seg16, h->G4p, l.K16, h->G4p, h->grads + h->emb_off, h->E, loss_scale), false, false, s);
            if (rc) return rc;
            if ((rc = record_stage_event(h, 1, s))) return rc;
            // dK[:E] = loss_scale * embedding^T * S   (contraction over the V' words)
            rc = gemm_f16(h, mk(l.in, h->G4, V1, h->emb16, h->Ep, h->seg16, h->G4p, gK0, h->G4, loss_scale, nullptr, 0, 0, 1), true, true, s);
            if (rc) return rc;
            // dK[E:] = loss_scale * h_{t-1}^T * dgates_t  (unchanged: contraction over the tokens)
            if (T > 1) {
                rc = gemm_f16(h, mk(H, h->G4, (int)(NT - N), l.hs, h->Hp, h->dgates + (int64_t)N * h->G4p, h->G4p,
                                    gK0 + (int64_t)l.in * h->G4, h->G4, loss_scale, nullptr, 0, 0, 1), true, true, s);
                if (rc) return rc;
            }
        }
        if (seg) {
            // per-occurrence square norm of the IndexedSlices rows: ||loss_scale * dgates[r,:] * K[:E]^T||^2 summed over tokens (A.6)
            ProfScope ps_dx(h, PH_DX, s);
            GemmArgs gn = mk((int)NT, l.in, h->G4, h->dgates, h->G4p, l.K16, h->G4p, nullptr, l.in);
            gn.alpha = loss_scale;
            rc = tc_gemm_scatter(h->tc, gn, h->x_ids, nullptr, h->E, h->grads + h->n_params + 1, s);
            LAUNCH_COUNT(h);
            if (rc) return rc;
            cur ^= 1;
            continue;
        }
        // db = loss_scale * colsum(dgates)
        ProfScope* ps_w = new ProfScope(h, PH_WGRAD, s);
        {
            int rpb = 256;
            dim3 grid(cdiv(h->G4, 128), cdiv(NT, rpb));
            colsum_f16_kernel<<<grid, 128, 0, s>>>(h->dgates, h->G4p, NT, h->G4, loss_scale, h->grads + l.b_off, rpb);
            LAUNCH_COUNT(h);
        }
        // dK[:in]  = loss_scale * in^T * dgates        (contraction over all tokens)
        float* gK = h->grads + l.k_off;
        rc = gemm_f16(h, mk(l.in, h->G4, (int)NT, in, l.inp, h->dgates, h->G4p, gK, h->G4, loss_scale, nullptr, 0, 0, 1), true, true, s);
        if (rc) { delete ps_w; return rc; }
        // dK[in:]  = loss_scale * h_{t-1}^T * dgates_t  (tokens of steps 1..T-1; h_{-1} = 0)
        if (T > 1) {
            rc = gemm_f16(h, mk(H, h->G4, (int)(NT - N), l.hs, h->Hp, h->dgates + (int64_t)N * h->G4p, h->G4p,
                                gK + (int64_t)l.in * h->G4, h->G4, loss_scale, nullptr, 0, 0, 1), true, true, s);
            if (rc) { delete ps_w; return rc; }
        }
        delete ps_w;
        // dInput[NT,in] = dgates * kernel[:in,:]^T
        ProfScope ps_dx(h, PH_DX, s);
        float* dnext = h->dact[cur ^ 1];
        GemmArgs gdx = mk((int)NT, l.in, h->G4, h->dgates, h->G4p, l.K16, h->G4p, dnext, l.in);
        if (li == 0 && !(h->cfg.flags & FSMG_FLAG_SIMT_GEMM) && tc_gemm_supported(gdx, false, false)) {
            // layer 0: the rows of dX are the IndexedSlices values of the embedding gradient — scatter-add them (x loss_scale)
            // into the dense gradient straight from the GEMM epilogue, accumulating the per-occurrence square norm (A.6)
            gdx.alpha = loss_scale;
            rc = tc_gemm_scatter(h->tc, gdx, h->x_ids, h->grads + h->emb_off, h->E, h->grads + h->n_params + 1, s);
            LAUNCH_COUNT(h);
            if (rc) return rc;
        } else {
            rc = gemm_f16(h, gdx, false, false, s);
            if (rc) return rc;
            if (li == 0) {
                scatter_emb_grad_kernel<<<(unsigned)NT, 128, 0, s>>>(dnext, l.in, h->x_ids, NT, h->E, loss_scale,
                                                                     h->grads + h->emb_off, h->grads + h->n_params + 1);
                LAUNCH_COUNT(h);
            }
        }
        if (li == 0 && (rc = record_stage_event(h, 1, s))) return rc;   // embedding gradient final (scatter / epilogue route)
        cur ^= 1;
    }
    FSMG_LAUNCH_OK();
    return FSMG_OK;
}

// ---- greedy sampler v2: every contraction on the tensor cores with split-fp16 operands (fp32-grade logits) ----
static int sampler_prepare(fsmg_handle* h, cudaStream_t s) {
    const int TB = 256, H = h->H;
    auto wsplit = [&](const float* in, int64_t ldi, __half* out, int Kp, int K, int N) {
        dim3 grid(cdiv(N, 32), cdiv(Kp, 32));
        split_weight_t_kernel<<<grid, dim3(32, 8), 0, s>>>(in, ldi, out, Kp, K, N);
        LAUNCH_COUNT(h);
    };
    split_act_kernel<<<cdiv((int64_t)h->V1 * h->Ep, TB), TB, 0, s>>>(h->params + h->emb_off, h->E, h->emb3, h->Ep, h->V1, h->E);
    LAUNCH_COUNT(h);
    for (int l = 0; l < h->L; ++l) {
        LayerBuf& lb = h->layers[l];
        const float* K = h->params + lb.k_off;
        wsplit(K, h->G4, h->KxT3[l], lb.inp, lb.in, h->G4);
        wsplit(K + (int64_t)lb.in * h->G4, h->G4, h->KhT3[l], h->Hp, H, h->G4);
    }
    wsplit(h->params + h->sw_off, h->V1, h->WsT3, h->Hp, H, h->V1);
    FSMG_LAUNCH_OK();
    // P0[v,:] = embedding[v,:] * Wx_0 + b_0 : the input contraction of layer 0 for every possible word, once
    int rc = gemm_f16(h, mk(h->V1, h->G4, 3 * h->Ep, h->emb3, 3 * h->Ep, h->KxT3[0], 3 * h->Ep, h->P0, h->G4, 1.0f / 2048.0f,
                            h->params + h->layers[0].b_off), false, false, s);
    if (rc) return rc;
    h->samp_stale = false;
    return FSMG_OK;
}

// plain launch, optionally as a programmatic dependent of the previous kernel on the stream (simt_kernels.cuh: griddep_wait)
template <typename... KArgs, typename... Args>
static int launch_maybe_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, int pdl, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    FSMG_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
    return FSMG_OK;
}

// one decode step for n songs: recurrent (+ lower-layer) contractions, cell, projection, argmax, step counter.
// FSMG_SAMPLE_PDL=1 (off by default): the four kernels of a step are chained by programmatic dependent launches — each one's
// block scheduling and prologue (for the GEMMs: TMEM allocation, barrier initialisation, cluster sync, descriptor prefetch)
// overlaps the previous kernel's execution; the data dependence is enforced by griddepcontrol.wait inside the kernels.
// Measured on a B200 (256 songs, E=H=1024, V=4708; gpurun_out r3c): token indices unchanged, 42.0 us per step against 38.1 us
// with ordinary graph edges — the early-resident 226 KB GEMM CTAs cost the running cell / argmax kernels more than the hidden
// prologues save — so the plain edges stay.
static int sample_step_split(fsmg_handle* h, int n, cudaStream_t s) {
    const int TB = 256, H = h->H;
    const float a = 1.0f / 2048.0f;
    static const int SAMP_PDL = [] { const char* e = getenv("FSMG_SAMPLE_PDL"); return e ? atoi(e) : 0; }();
    // 128-wide N tiles for the step's 256-row GEMMs (default; FSMG_SAMPLE_BN=256 restores the 256-wide plan): 32 / 37 tiles instead of
    // 16 / 19, so the 74 CTA pairs are filled by a 2-way split of K with 8-10 MB of fp32 partial sums per GEMM instead of a 4-way
    // split / 3.7-way stream-K cut with 17-18 MB — measured 38.2 -> 35.2 us per 256-song decode step (gpurun_out r3f), token indices
    // unchanged (plans: fsmg_debug_plan, pinned in tests/test_abi_and_host.py)
    static const int SAMP_NARROW = [] { const char* e = getenv("FSMG_SAMPLE_BN"); return e && atoi(e) == 256 ? 0 : 1; }();
    struct PdlScope {
        TcContext& c;
        PdlScope(TcContext& c_, int v, int nw) : c(c_) { c.pdl = v; c.narrow = nw; }
        ~PdlScope() { c.pdl = 0; c.narrow = 0; }
    } pdl_scope(h->tc, SAMP_PDL, SAMP_NARROW);
    int rc;
    for (int l = 0; l < h->L; ++l) {
        // accumulating GEMMs into buffers their consumers (cell / argmax kernels) leave zeroed: no memset node per token
        rc = gemm_f16(h, mk(n, h->G4, 3 * h->Hp, h->h3[l], 3 * h->Hp, h->KhT3[l], 3 * h->Hp, h->s_g, h->G4, a, nullptr, 0, 0, 1), false, false, s);
        if (rc) return rc;
        if (l > 0) {
            rc = gemm_f16(h, mk(n, h->G4, 3 * h->Hp, h->h3[l - 1], 3 * h->Hp, h->KxT3[l], 3 * h->Hp, h->s_g, h->G4, a, nullptr, 0, 0, 1), false, false, s);
            if (rc) return rc;
        }
        if ((H & 3) == 0)
            rc = launch_maybe_pdl(sample_cell_kernel, dim3((unsigned)cdiv((int64_t)n * (H / 4), TB)), dim3(TB), s, SAMP_PDL, h->s_g, h->G4,
                                  l == 0 ? h->P0 : nullptr, h->G4, h->samp_ids, l == 0 ? nullptr : h->params + h->layers[l].b_off, h->s_c[l],
                                  h->h3[l], h->Hp, n, H);
        else
            rc = launch_maybe_pdl(sample_cell_scalar_kernel, dim3((unsigned)cdiv((int64_t)n * H, TB)), dim3(TB), s, SAMP_PDL, h->s_g, h->G4,
                                  l == 0 ? h->P0 : nullptr, h->G4, h->samp_ids, l == 0 ? nullptr : h->params + h->layers[l].b_off, h->s_c[l],
                                  h->h3[l], h->Hp, n, H);
        if (rc) return rc;
        LAUNCH_COUNT(h);
    }
    rc = gemm_f16(h, mk(n, h->V1, 3 * h->Hp, h->h3[h->L - 1], 3 * h->Hp, h->WsT3, 3 * h->Hp, h->s_logits2, h->Vp, a,
                        h->params + h->sb_off, 0, 0, 1), false, false, s);
    if (rc) return rc;
    rc = launch_maybe_pdl(argmax_rows_step_kernel, dim3(n), dim3(256), s, SAMP_PDL, h->s_logits2, h->Vp, h->V1, h->samp_ids, h->samp_out, 4096,
                          h->s_step);
    if (rc) return rc;
    h->launches += 1;
    FSMG_LAUNCH_OK();
    return FSMG_OK;
}

static int sample_greedy_split(fsmg_handle* h, int n, int n_tokens, int32_t* d_out, cudaStream_t s) {
    const int TB = 256, H = h->H;
    int rc;
    if (n_tokens > 4096) return set_error(FSMG_ERR_CAPACITY, "n_tokens > 4096");
    if (h->samp_stale && (rc = sampler_prepare(h, s))) return rc;
    fill_i32_kernel<<<cdiv(n, TB), TB, 0, s>>>(h->samp_ids, n, h->V);   // word = start word (lstm_baseline.py:138)
    LAUNCH_COUNT(h);
    FSMG_CUDA_OK(cudaMemsetAsync(h->s_step, 0, 2 * sizeof(int), s));
    FSMG_CUDA_OK(cudaMemsetAsync(h->s_g, 0, sizeof(float) * (size_t)n * h->G4, s));          // accumulation targets of the decode-step GEMMs:
    FSMG_CUDA_OK(cudaMemsetAsync(h->s_logits2, 0, sizeof(float) * (size_t)n * h->Vp, s));     // zeroed once, then kept zero by their consumers
    for (int l = 0; l < h->L; ++l) {   // zero_state (lstm_baseline.py:140)
        FSMG_CUDA_OK(cudaMemsetAsync(h->s_c[l], 0, sizeof(float) * n * H, s));
        FSMG_CUDA_OK(cudaMemsetAsync(h->h3[l], 0, sizeof(__half) * (size_t)n * 3 * h->Hp, s));
    }
    const char* env_g = getenv("FSMG_SAMPLE_GRAPH");
    const bool use_graph = (env_g ? atoi(env_g) != 0 : true) && h->aux[0] != nullptr && n_tokens >= 4;
    if (use_graph) {
        // the decode step is identical for every token (the step index lives on the device): capture it once on an
        // internal stream and replay the graph — one launch per generated token instead of ~8
        // two graphs: SAMP_UNROLL consecutive steps (bulk of the sequence: 1/16th of the graph launches and of the
        // launch-to-launch gaps) and a single step (remainder)
        static const int SAMP_UNROLL = [] { const char* e = getenv("FSMG_SAMPLE_UNROLL"); int v = e ? atoi(e) : 16; return v < 1 ? 1 : v; }();
        if (!h->samp_graph || h->samp_graph_n != n) {
            if (h->samp_graph) { cudaGraphExecDestroy(h->samp_graph); h->samp_graph = nullptr; }
            if (h->samp_graph_multi) { cudaGraphExecDestroy(h->samp_graph_multi); h->samp_graph_multi = nullptr; }
            for (int which = 0; which < 2; ++which) {
                cudaGraph_t graph = nullptr;
                FSMG_CUDA_OK(cudaStreamBeginCapture(h->aux[0], cudaStreamCaptureModeThreadLocal));
                int64_t saved = h->launches;
                rc = 0;
                for (int k = 0; k < (which ? SAMP_UNROLL : 1) && !rc; ++k) rc = sample_step_split(h, n, h->aux[0]);
                if (which == 0) h->samp_step_launches = h->launches - saved;
                h->launches = saved;
                cudaError_t ce = cudaStreamEndCapture(h->aux[0], &graph);
                if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                if (ce != cudaSuccess) return set_error(FSMG_ERR_CUDA, "graph capture of the decode step failed: %s", cudaGetErrorString(ce));
                ce = cudaGraphInstantiate(which ? &h->samp_graph_multi : &h->samp_graph, graph, 0);
                cudaGraphDestroy(graph);
                if (ce != cudaSuccess) return set_error(FSMG_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
            }
            h->samp_graph_n = n;
        }
        int step = 0;
        for (; step + SAMP_UNROLL <= n_tokens; step += SAMP_UNROLL) FSMG_CUDA_OK(cudaGraphLaunch(h->samp_graph_multi, s));
        for (; step < n_tokens; ++step) FSMG_CUDA_OK(cudaGraphLaunch(h->samp_graph, s));
        h->launches += (int64_t)n_tokens * h->samp_step_launches;    // kernels executed (graph nodes), not host launches
    } else {
        for (int step = 0; step < n_tokens; ++step)
            if ((rc = sample_step_split(h, n, s))) return rc;
    }
    FSMG_CUDA_OK(cudaMemcpy2DAsync(d_out, (size_t)n_tokens * 4, h->samp_out, 4096 * 4, (size_t)n_tokens * 4, n, cudaMemcpyDeviceToDevice, s));
    return FSMG_OK;
}

static int check_call(fsmg_handle* h, int n_seqs) {
    if (!h) return set_error(FSMG_ERR_INVALID, "null handle");
    if (!h->bound) return set_error(FSMG_ERR_STATE, "fsmg_bind has not been called");
    if (n_seqs <= 0 || n_seqs > h->Nmax)
        return set_error(FSMG_ERR_CAPACITY, "n_seqs=%d outside [1, max_seqs=%d]", n_seqs, h->Nmax);
    return FSMG_OK;
}

}  // namespace fsmg

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* fsmg_last_error(void) { return fsmg::last_error().c_str(); }
int fsmg_abi_version(void) { return FSMG_ABI_VERSION; }

int fsmg_create(const fsmg_config* cfg, const char* scope_name, fsmg_handle** out) {
    if (!cfg || !out) return set_error(FSMG_ERR_INVALID, "null argument");
    if (cfg->vocab <= 0 || cfg->embed <= 0 || cfg->hidden <= 0 || cfg->layers <= 0 || cfg->max_len <= 0 || cfg->max_seqs <= 0)
        return set_error(FSMG_ERR_INVALID, "non-positive dimension in fsmg_config");
    if (cfg->max_len > 4096) return set_error(FSMG_ERR_INVALID, "max_len > 4096 unsupported");
    // several kernels put the token count / 64 in gridDim.y (limit 65535) and index tokens with 32-bit ints
    if ((int64_t)cfg->max_len * cfg->max_seqs > (int64_t)65535 * 64)
        return set_error(FSMG_ERR_INVALID, "max_seqs * max_len = %lld tokens per call exceeds the supported %lld", (long long)cfg->max_len * cfg->max_seqs,
                         (long long)65535 * 64);
    fsmg_handle* h = new fsmg_handle();
    h->cfg = *cfg;
    h->scope = scope_name && *scope_name ? scope_name : "lstm_baseline";
    h->V = cfg->vocab; h->V1 = cfg->vocab + 1; h->E = cfg->embed; h->H = cfg->hidden; h->L = cfg->layers;
    h->T = cfg->max_len; h->Nmax = cfg->max_seqs;
    h->Ep = (int)round_up(h->E, 8); h->Hp = (int)round_up(h->H, 8); h->G4 = 4 * h->H; h->G4p = (int)round_up(h->G4, 8);
    h->Vp = (int)round_up(h->V1, 16);   // fp16 logits rows start on 32-byte boundaries (256-bit stores in the LSE epilogue)
    // flat parameter layout, TF get_vars() order (reference tf_model.py:99-104)
    int64_t off = 0;
    auto add = [&](const std::string& name, int rows, int cols) {
        fsmg_param_info pi;
        memset(&pi, 0, sizeof pi);
        snprintf(pi.name, sizeof pi.name, "%s", name.c_str());
        pi.offset = off; pi.rows = rows; pi.cols = cols;
        h->infos.push_back(pi);
        int64_t o = off;
        off = round_up(off + (int64_t)rows * cols, 64);
        return o;
    };
    h->emb_off = add(h->scope + "/embedding", h->V1, h->E);
    h->dense_begin = off;
    h->layers.resize(h->L);
    for (int l = 0; l < h->L; ++l) {
        LayerBuf& lb = h->layers[l];
        lb.in = l == 0 ? h->E : h->H;
        lb.inp = (int)round_up(lb.in, 8);
        std::string base = h->scope + "/rnn/multi_rnn_cell/cell_" + std::to_string(l) + "/basic_lstm_cell";
        lb.k_off = add(base + "/kernel", lb.in + h->H, h->G4);
        lb.b_off = add(base + "/bias", h->G4, 1);
    }
    h->sw_off = add(h->scope + "/softmax_w", h->H, h->V1);
    h->sb_off = add(h->scope + "/softmax_b", h->V1, 1);
    h->n_params = off;
    carve(h, nullptr);
    *out = h;
    return FSMG_OK;
}

void fsmg_destroy(fsmg_handle* h) {
    if (!h) return;
    if (h->h_tok) cudaFreeHost(h->h_tok);
    if (h->h_scal) cudaFreeHost(h->h_scal);
    for (auto ev : h->prof.pool) cudaEventDestroy(ev);
    if (h->samp_graph) cudaGraphExecDestroy(h->samp_graph);
    if (h->samp_graph_multi) cudaGraphExecDestroy(h->samp_graph_multi);
    for (auto& e : h->step_graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
    for (int i = 0; i < 2; ++i) {
        if (h->aux[i]) cudaStreamDestroy(h->aux[i]);
        if (h->ev_ready[i]) cudaEventDestroy(h->ev_ready[i]);
        if (h->ev_dh[i]) cudaEventDestroy(h->ev_dh[i]);
        if (h->ev_dws[i]) cudaEventDestroy(h->ev_dws[i]);
    }
    if (h->ev_sort_fork) cudaEventDestroy(h->ev_sort_fork);
    if (h->ev_sort_join) cudaEventDestroy(h->ev_sort_join);
    delete h;
}

int64_t fsmg_param_count(const fsmg_handle* h) { return h ? h->n_params : 0; }
int64_t fsmg_grad_count(const fsmg_handle* h) { return h ? h->n_params + FSMG_GRAD_EXTRA : 0; }
int64_t fsmg_workspace_bytes(const fsmg_handle* h) { return h ? h->ws_need : 0; }
int fsmg_num_params(const fsmg_handle* h) { return h ? (int)h->infos.size() : 0; }
int fsmg_param_info_at(const fsmg_handle* h, int index, fsmg_param_info* out) {
    if (!h || !out || index < 0 || index >= (int)h->infos.size()) return set_error(FSMG_ERR_INVALID, "bad param index");
    *out = h->infos[index];
    return FSMG_OK;
}
int64_t fsmg_last_launch_count(const fsmg_handle* h) { return h ? h->launches : 0; }

int fsmg_set_profile(fsmg_handle* h, int enable) {
    if (!h) return set_error(FSMG_ERR_INVALID, "null handle");
    h->prof.on = enable != 0;
    return FSMG_OK;
}
int fsmg_read_profile(fsmg_handle* h, float* ms_out, int32_t* count_out, void* stream) {
    if (!h || !ms_out) return set_error(FSMG_ERR_INVALID, "null argument");
    FSMG_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    for (int i = 0; i < PH_COUNT; ++i) { ms_out[i] = 0.0f; if (count_out) count_out[i] = 0; }
    for (auto& r : h->prof.recs) {
        float ms = 0.0f;
        FSMG_CUDA_OK(cudaEventElapsedTime(&ms, r.a, r.b));
        ms_out[r.ph] += ms;
        if (count_out) count_out[r.ph]++;
    }
    h->prof.recs.clear();
    h->prof.used = 0;
    return FSMG_OK;
}
const char* fsmg_profile_phase_name(int phase) { return (phase >= 0 && phase < PH_COUNT) ? kPhaseNames[phase] : ""; }

int fsmg_bind(fsmg_handle* h, float* d_params, float* d_grads, float* d_adam_m, float* d_adam_v, void* d_workspace,
              int64_t workspace_bytes) {
    if (!h || !d_params || !d_workspace) return set_error(FSMG_ERR_INVALID, "null argument");
    if (workspace_bytes < h->ws_need)
        return set_error(FSMG_ERR_CAPACITY, "workspace %lld < required %lld bytes", (long long)workspace_bytes, (long long)h->ws_need);
    if ((reinterpret_cast<uintptr_t>(d_workspace) & 255) || (reinterpret_cast<uintptr_t>(d_params) & 15))
        return set_error(FSMG_ERR_INVALID, "workspace must be 256-byte aligned, params 16-byte aligned");
    h->params = d_params; h->grads = d_grads; h->adam_m = d_adam_m; h->adam_v = d_adam_v;
    h->ws = (char*)d_workspace; h->ws_bytes = workspace_bytes;
    carve(h, h->ws);
    int rc = tc_init(h->tc);
    if (rc) return rc;
    if (!h->aux[0]) {
        for (int i = 0; i < 2; ++i) {
            FSMG_CUDA_OK(cudaStreamCreateWithFlags(&h->aux[i], cudaStreamNonBlocking));
            FSMG_CUDA_OK(cudaEventCreateWithFlags(&h->ev_ready[i], cudaEventDisableTiming));
            FSMG_CUDA_OK(cudaEventCreateWithFlags(&h->ev_dh[i], cudaEventDisableTiming));
            FSMG_CUDA_OK(cudaEventCreateWithFlags(&h->ev_dws[i], cudaEventDisableTiming));
        }
        FSMG_CUDA_OK(cudaEventCreateWithFlags(&h->ev_sort_fork, cudaEventDisableTiming));
        FSMG_CUDA_OK(cudaEventCreateWithFlags(&h->ev_sort_join, cudaEventDisableTiming));
    }
    for (auto& e : h->step_graphs) if (e.exec) cudaGraphExecDestroy(e.exec);   // graphs hold the old buffer addresses
    h->step_graphs.clear();
    if (h->samp_graph) { cudaGraphExecDestroy(h->samp_graph); h->samp_graph = nullptr; }
    if (h->samp_graph_multi) { cudaGraphExecDestroy(h->samp_graph_multi); h->samp_graph_multi = nullptr; }
    h->samp_stale = true;
    h->bound = true;
    FSMG_CUDA_OK(cudaMemset(h->scalars, 0, 512 * sizeof(float)));
    return FSMG_OK;
}

int fsmg_refresh_weights(fsmg_handle* h, void* stream) {
    if (!h || !h->bound) return set_error(FSMG_ERR_STATE, "not bound");
    return refresh_weights(h, (cudaStream_t)stream);
}

int fsmg_forward_nll(fsmg_handle* h, const int32_t* d_tokens, int32_t n_seqs, float* d_nll, float* d_sum_nll, void* stream) {
    int rc = check_call(h, n_seqs);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    h->launches = 0;
    rc = forward_lstm(h, d_tokens, n_seqs, false, s);
    if (rc) return rc;
    rc = projection(h, n_seqs, false, 0.0f, d_nll, s);
    if (rc) return rc;
    if (d_sum_nll) {
        FSMG_CUDA_OK(cudaMemsetAsync(d_sum_nll, 0, sizeof(float), s));
        sum_f32_kernel<<<148, 256, 0, s>>>(d_nll ? d_nll : h->nll, (int64_t)n_seqs * h->T, d_sum_nll);
        LAUNCH_COUNT(h);
        FSMG_LAUNCH_OK();
    }
    return FSMG_OK;
}

static int enqueue_forward_backward(fsmg_handle* h, const int32_t* d_tokens, int32_t n_seqs, float loss_scale, float* d_nll, cudaStream_t s) {
    h->launches = 0;
    FSMG_CUDA_OK(cudaMemsetAsync(h->grads, 0, sizeof(float) * (h->n_params + FSMG_GRAD_EXTRA), s));
    int rc = forward_lstm(h, d_tokens, n_seqs, true, s);
    if (rc) return rc;
    rc = projection(h, n_seqs, true, loss_scale, d_nll, s);
    if (rc) return rc;
    if ((rc = record_stage_event(h, 0, s))) return rc;   // softmax_w / softmax_b gradients are final
    sum_f32_kernel<<<148, 256, 0, s>>>(d_nll ? d_nll : h->nll, (int64_t)n_seqs * h->T, h->grads + h->n_params);
    LAUNCH_COUNT(h);
    if ((rc = record_stage_event(h, 2, s))) return rc;   // the step's loss (sum of NLL, token count) can be read back from here on
    rc = backward_lstm(h, n_seqs, loss_scale, s);
    if (rc) return rc;
    FSMG_LAUNCH_OK();
    return FSMG_OK;
}

int fsmg_forward_backward(fsmg_handle* h, const int32_t* d_tokens, int32_t n_seqs, float loss_scale, float* d_nll, void* stream) {
    int rc = check_call(h, n_seqs);
    if (rc) return rc;
    if (!h->grads) return set_error(FSMG_ERR_STATE, "no gradient buffer bound");
    cudaStream_t s = (cudaStream_t)stream;
    const bool graphable = h->use_graph && !h->prof.on && !h->overlap && h->aux[1] != nullptr && getenv("FSMG_TRACE") == nullptr;
    if (!graphable) return enqueue_forward_backward(h, d_tokens, n_seqs, loss_scale, d_nll, s);
    uint32_t ls_bits;
    memcpy(&ls_bits, &loss_scale, 4);
    fsmg_handle::StepGraph* g = nullptr;
    for (auto& e : h->step_graphs)
        if (e.n == n_seqs && e.ls_bits == ls_bits && e.nll == d_nll) { g = &e; break; }
    if (!g) {
        if (h->step_graphs.size() >= 8) {   // callers that rotate through many batch sizes: drop the oldest graph
            if (h->step_graphs.front().exec) cudaGraphExecDestroy(h->step_graphs.front().exec);
            h->step_graphs.erase(h->step_graphs.begin());
        }
        h->step_graphs.push_back({n_seqs, ls_bits, d_nll, nullptr, 0, 0, 0});
        g = &h->step_graphs.back();
    }
    // first sight of a configuration: plain launches (module loading, function attributes); second: capture; then replay
    if (g->failed || g->seen++ == 0) return enqueue_forward_backward(h, d_tokens, n_seqs, loss_scale, d_nll, s);
    // the graph reads its tokens from a fixed internal buffer, so any caller buffer replays the same graph
    FSMG_CUDA_OK(cudaMemcpyAsync(h->tok_graph, d_tokens, sizeof(int32_t) * (size_t)n_seqs * h->T, cudaMemcpyDeviceToDevice, s));
    d_tokens = h->tok_graph;
    if (!g->exec) {
        cudaStream_t cs = h->aux[1];
        cudaGraph_t graph = nullptr;
        FSMG_CUDA_OK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        rc = enqueue_forward_backward(h, d_tokens, n_seqs, loss_scale, d_nll, cs);
        cudaError_t ce = cudaStreamEndCapture(cs, &graph);
        if (!rc && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&g->exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (rc || ce != cudaSuccess || !g->exec) {   // not capturable here: keep the plain path for this configuration
            cudaGetLastError();
            g->exec = nullptr;
            g->failed = 1;
            if (getenv("FSMG_GRAPH_DEBUG")) fprintf(stderr, "[fsmg] step graph capture failed (rc=%d, %s)\n", rc, cudaGetErrorString(ce));
            return enqueue_forward_backward(h, d_tokens, n_seqs, loss_scale, d_nll, s);
        }
        g->launches = h->launches;
    }
    FSMG_CUDA_OK(cudaGraphLaunch(g->exec, s));
    h->launches = g->launches;
    return FSMG_OK;
}

int fsmg_set_stage_events(fsmg_handle* h, void* ev_softmax_grads, void* ev_embedding_grads, int32_t reserve_sms) {
    if (!h) return set_error(FSMG_ERR_INVALID, "null handle");
    if (reserve_sms < 0 || reserve_sms > 64) return set_error(FSMG_ERR_INVALID, "reserve_sms outside [0, 64]");
    h->stage_ev[0] = (cudaEvent_t)ev_softmax_grads;
    h->stage_ev[1] = (cudaEvent_t)ev_embedding_grads;
    h->tc.lstm_reserve_sms = reserve_sms;
    for (auto& e : h->step_graphs) if (e.exec) cudaGraphExecDestroy(e.exec);   // captured graphs hold the old events / grid sizes
    h->step_graphs.clear();
    return FSMG_OK;
}

int fsmg_set_loss_event(fsmg_handle* h, void* ev_loss_ready) {
    if (!h) return set_error(FSMG_ERR_INVALID, "null handle");
    h->stage_ev[2] = (cudaEvent_t)ev_loss_ready;
    for (auto& e : h->step_graphs) if (e.exec) cudaGraphExecDestroy(e.exec);   // captured graphs hold the old event
    h->step_graphs.clear();
    return FSMG_OK;
}

int fsmg_param_range(const fsmg_handle* h, int32_t which, int64_t* begin, int64_t* end) {
    if (!h || !begin || !end) return set_error(FSMG_ERR_INVALID, "null argument");
    switch (which) {
        case 0: *begin = h->emb_off; *end = h->dense_begin; break;          // embedding
        case 1: *begin = h->dense_begin; *end = h->sw_off; break;           // LSTM kernels and biases
        case 2: *begin = h->sw_off; *end = h->n_params; break;              // softmax_w, softmax_b
        case 3: *begin = h->n_params; *end = h->n_params + FSMG_GRAD_EXTRA; break;   // piggy-backed scalars (gradient buffer only)
        default: return set_error(FSMG_ERR_INVALID, "range index outside [0, 3]");
    }
    return FSMG_OK;
}

int fsmg_apply_update(fsmg_handle* h, int64_t step, float* d_out_norm, void* stream) {
    if (!h || !h->bound) return set_error(FSMG_ERR_STATE, "not bound");
    if (!h->grads || !h->adam_m || !h->adam_v) return set_error(FSMG_ERR_STATE, "optimizer buffers not bound");
    cudaStream_t s = (cudaStream_t)stream;
    // exponential_decay on the pre-increment step, fp32 like TF (A.8); Adam bias correction (A.7)
    float lr_k = h->cfg.lr * powf(0.5f, (float)step / (float)h->cfg.n_decay);
    double t = (double)step + 1.0;
    double alpha = (double)lr_k * sqrt(1.0 - pow((double)h->cfg.beta2, t)) / (1.0 - pow((double)h->cfg.beta1, t));
    float* dense_sq = h->scalars + 64;   // SQNORM_BLOCKS per-block partials, combined in a fixed order by clip_adam_kernel
    ProfScope ps(h, PH_UPDATE, s);
    sqnorm_f32_kernel<<<SQNORM_BLOCKS, 256, 0, s>>>(h->grads, h->dense_begin, h->n_params, dense_sq);
    clip_adam_kernel<<<592, 256, 0, s>>>(h->params, h->grads, h->adam_m, h->adam_v, h->n_params, dense_sq,
                                         h->grads + h->n_params + 1, h->cfg.max_grad_norm, (float)alpha, h->cfg.beta1,
                                         h->cfg.beta2, h->cfg.eps, d_out_norm);
    h->launches += 2;
    FSMG_LAUNCH_OK();
    return refresh_weights(h, s);
}

int fsmg_sample_greedy(fsmg_handle* h, int32_t n_songs, int32_t n_tokens, int32_t* d_out, void* stream) {
    if (!h || !h->bound) return set_error(FSMG_ERR_STATE, "not bound");
    if (n_songs <= 0 || n_songs > h->samp_max) return set_error(FSMG_ERR_CAPACITY, "n_songs=%d outside [1,%d]", n_songs, h->samp_max);
    if (n_tokens <= 0) return set_error(FSMG_ERR_INVALID, "n_tokens must be positive");
    cudaStream_t s = (cudaStream_t)stream;
    h->launches = 0;
    const int H = h->H, TB = 256, n = n_songs;
    if (!(h->cfg.flags & FSMG_FLAG_SIMT_RECURRENT)) return sample_greedy_split(h, n, n_tokens, d_out, s);
    fill_i32_kernel<<<cdiv(n, TB), TB, 0, s>>>(h->samp_ids, n, h->V);  // word = start word (lstm_baseline.py:138)
    LAUNCH_COUNT(h);
    for (int l = 0; l < h->L; ++l) {  // zero_state (lstm_baseline.py:140)
        FSMG_CUDA_OK(cudaMemsetAsync(h->s_c[l], 0, sizeof(float) * n * H, s));
        FSMG_CUDA_OK(cudaMemsetAsync(h->s_h[l], 0, sizeof(float) * n * H, s));
    }
    for (int step = 0; step < n_tokens; ++step) {
        gather_rows_f32_kernel<<<cdiv((int64_t)n * h->E, TB), TB, 0, s>>>(h->params + h->emb_off, h->E, h->samp_ids, h->s_x, h->E, n, h->E);
        LAUNCH_COUNT(h);
        const float* in = h->s_x;
        int in_w = h->E;
        for (int l = 0; l < h->L; ++l) {
            LayerBuf& lb = h->layers[l];
            const float* K = h->params + lb.k_off;
            // fp32 weights, fp32 FMA: argmax near-ties need fp32-grade logits
            launch_simt_gemm<float, float>(mk(n, h->G4, in_w, in, in_w, K, h->G4, h->s_g, h->G4, 1.0f, h->params + lb.b_off), false, true, s);
            launch_simt_gemm<float, float>(mk(n, h->G4, H, h->s_h[l], H, K + (int64_t)in_w * h->G4, h->G4, h->s_g, h->G4, 1.0f, nullptr, 0, 1), false, true, s);
            lstm_pointwise_fwd_kernel<float><<<cdiv((int64_t)n * H, TB), TB, 0, s>>>(h->s_g, h->G4, nullptr, 0, h->s_c[l], nullptr, 0, h->s_c[l], h->s_h[l], H, n, H);
            h->launches += 3;
            in = h->s_h[l];
            in_w = H;
        }
        launch_simt_gemm<float, float>(mk(n, h->V1, H, in, H, h->params + h->sw_off, h->V1, h->s_logits, h->V1, 1.0f, h->params + h->sb_off), false, true, s);
        argmax_rows_kernel<<<n, 256, 0, s>>>(h->s_logits, h->V1, h->V1, h->samp_ids, d_out, n_tokens, step);
        h->launches += 2;
    }
    FSMG_LAUNCH_OK();
    return FSMG_OK;
}

// ---- host-buffer entry points ---------------------------------------------------------------------
// TensorFlow raises InvalidArgument for an id outside the embedding table (reference lstm_baseline.py:41, tf.nn.embedding_lookup)
static int check_host_tokens(const fsmg_handle* h, const int32_t* tok, int64_t n) {
    if (!tok) return set_error(FSMG_ERR_INVALID, "null token buffer");
    for (int64_t i = 0; i < n; ++i)
        if (tok[i] < 0 || tok[i] > h->V)
            return set_error(FSMG_ERR_INVALID, "token id %d at flat index %lld outside [0, %d] (input_size = %d)", tok[i], (long long)i, h->V, h->V);
    return FSMG_OK;
}

static int ensure_host_staging(fsmg_handle* h, int64_t tok_elems) {
    if (!h->h_scal) FSMG_CUDA_OK(cudaHostAlloc((void**)&h->h_scal, 64 * sizeof(float), cudaHostAllocDefault));
    if (tok_elems > h->h_tok_elems) {
        if (h->h_tok) cudaFreeHost(h->h_tok);
        h->h_tok = nullptr;
        FSMG_CUDA_OK(cudaHostAlloc((void**)&h->h_tok, tok_elems * sizeof(int32_t), cudaHostAllocDefault));
        h->h_tok_elems = tok_elems;
    }
    return FSMG_OK;
}

int fsmg_eval_host(fsmg_handle* h, const int32_t* h_tokens, int32_t n_seqs, float* h_mean_nll, float* h_nll, void* stream) {
    int rc = check_call(h, n_seqs);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    int64_t n = (int64_t)n_seqs * h->T;
    rc = check_host_tokens(h, h_tokens, n);
    if (rc) return rc;
    rc = ensure_host_staging(h, (int64_t)h->Nmax * h->T);
    if (rc) return rc;
    memcpy(h->h_tok, h_tokens, n * sizeof(int32_t));
    FSMG_CUDA_OK(cudaMemcpyAsync(h->tok_stage, h->h_tok, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    rc = fsmg_forward_nll(h, h->tok_stage, n_seqs, nullptr, h->scalars, s);
    if (rc) return rc;
    FSMG_CUDA_OK(cudaMemcpyAsync(h->h_scal, h->scalars, sizeof(float), cudaMemcpyDeviceToHost, s));
    if (h_nll) FSMG_CUDA_OK(cudaMemcpyAsync(h_nll, h->nll, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    FSMG_CUDA_OK(cudaStreamSynchronize(s));
    if (h_mean_nll) *h_mean_nll = (float)((double)h->h_scal[0] / ((double)n + 1e-12));
    return FSMG_OK;
}

int fsmg_train_host(fsmg_handle* h, const int32_t* h_tokens, int32_t n_seqs, int64_t step, float* h_mean_loss, void* stream) {
    int rc = check_call(h, n_seqs);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    int64_t n = (int64_t)n_seqs * h->T;
    rc = check_host_tokens(h, h_tokens, n);
    if (rc) return rc;
    rc = ensure_host_staging(h, (int64_t)h->Nmax * h->T);
    if (rc) return rc;
    memcpy(h->h_tok, h_tokens, n * sizeof(int32_t));
    FSMG_CUDA_OK(cudaMemcpyAsync(h->tok_stage, h->h_tok, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    float loss_scale = (float)(1.0 / ((double)n + 1e-12));
    rc = fsmg_forward_backward(h, h->tok_stage, n_seqs, loss_scale, nullptr, s);
    if (rc) return rc;
    rc = fsmg_apply_update(h, step, nullptr, s);  // keeps accumulating h->launches
    if (rc) return rc;
    FSMG_CUDA_OK(cudaMemcpyAsync(h->h_scal, h->grads + h->n_params, sizeof(float), cudaMemcpyDeviceToHost, s));
    FSMG_CUDA_OK(cudaStreamSynchronize(s));
    if (h_mean_loss) *h_mean_loss = (float)((double)h->h_scal[0] / ((double)n + 1e-12));
    return FSMG_OK;
}

int fsmg_sample_host(fsmg_handle* h, int32_t n_songs, int32_t n_tokens, int32_t* h_out, void* stream) {
    if (!h || !h->bound) return set_error(FSMG_ERR_STATE, "not bound");
    if (n_tokens > 4096) return set_error(FSMG_ERR_CAPACITY, "n_tokens > 4096");
    cudaStream_t s = (cudaStream_t)stream;
    int rc = fsmg_sample_greedy(h, n_songs, n_tokens, h->samp_out2, s);
    if (rc) return rc;
    FSMG_CUDA_OK(cudaMemcpyAsync(h_out, h->samp_out2, (int64_t)n_songs * n_tokens * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    FSMG_CUDA_OK(cudaStreamSynchronize(s));
    return FSMG_OK;
}

// device-pointer callers: how many token ids of the calls so far were outside [0, V] (they were clamped); resets the count
int fsmg_token_range_errors(fsmg_handle* h, int64_t* h_count, void* stream) {
    if (!h || !h->bound || !h_count) return set_error(FSMG_ERR_STATE, "not bound / null argument");
    int v = 0;
    FSMG_CUDA_OK(cudaMemcpyAsync(&v, h->scalars + 32, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    FSMG_CUDA_OK(cudaMemsetAsync(h->scalars + 32, 0, sizeof(int), (cudaStream_t)stream));
    FSMG_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    *h_count = v;
    return FSMG_OK;
}

// the device-side input/target shift on its own (parity hook for base_model.py:63-86): time-major x / y ids of a token batch
int fsmg_debug_prep_tokens(fsmg_handle* h, const int32_t* d_tokens, int32_t n_seqs, int32_t* d_x_out, int32_t* d_y_out, void* stream) {
    int rc = check_call(h, n_seqs);
    if (rc) return rc;
    if (!d_tokens || !d_x_out || !d_y_out) return set_error(FSMG_ERR_INVALID, "null argument");
    const int64_t NT = (int64_t)n_seqs * h->T;
    prep_tokens_kernel<<<cdiv(NT, 256), 256, 0, (cudaStream_t)stream>>>(d_tokens, d_x_out, d_y_out, n_seqs, h->T, h->V, h->V,
                                                                          reinterpret_cast<int*>(h->scalars + 32));
    FSMG_LAUNCH_OK();
    return FSMG_OK;
}

// hardware probe: TMEM placement of D for tcgen05.mma.cta_group::2 with the given M, N (d_out: fp32 [2 CTAs][128 lanes][N])
int fsmg_debug_mma_probe(int32_t m, int32_t n, float* d_out, void* stream) {
    if (!d_out) return set_error(FSMG_ERR_INVALID, "null output");
    return mma_probe_2sm(m, n, d_out, (cudaStream_t)stream);
}

int fsmg_debug_gemm(int32_t m, int32_t n, int32_t k, const void* d_a_f16, const void* d_b_f16, float* d_c,
                    int32_t a_mn_major, int32_t b_mn_major, int32_t use_simt, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int64_t lda = a_mn_major ? m : k, ldb = b_mn_major ? n : k;
    GemmArgs g = mk(m, n, k, d_a_f16, lda, d_b_f16, ldb, d_c, n);
    if (use_simt) {
        launch_simt_gemm<__half, __half>(g, a_mn_major != 0, b_mn_major != 0, s);
        FSMG_LAUNCH_OK();
        return FSMG_OK;
    }
    static fsmg::TcContext ctx;
    int rc = tc_init(ctx);
    if (rc) return rc;
    if (!tc_gemm_supported(g, a_mn_major != 0, b_mn_major != 0)) return set_error(FSMG_ERR_INVALID, "shape unsupported by the tcgen05 GEMM");
    return tc_gemm(ctx, g, a_mn_major != 0, b_mn_major != 0, s);
}

// host-side planning of the tcgen05 GEMM core for a shape, without touching the device (defaults of a 148-SM B200): which tile
// width, cluster size, K split / stream-K schedule and grid a launch would get.  out[0..7] = BN, cluster size, grid (CTAs),
// M pair-tiles, N tiles, K splits, stream-K flag, k-block units per cluster.  narrow = the decode loop's 128-wide-tile request.
// CPU tests pin the plans the design rests on (which shapes take the 256 x 512 pair tiles that carry the operand transform, how
// the decode step's 256-row GEMMs are split) with it.
int fsmg_debug_plan(int32_t m, int32_t n, int32_t k, int32_t allow_split, int32_t narrow, int32_t* out8) {
    if (m <= 0 || n <= 0 || k <= 0 || !out8) return set_error(FSMG_ERR_INVALID, "fsmg_debug_plan: bad argument");
    fsmg::TcContext ctx;              // defaults only: no driver call, no environment
    ctx.narrow = narrow ? 1 : 0;
    const TcPlan p = tc_plan(ctx, m, n, k, allow_split != 0, 0, true);
    out8[0] = p.bn; out8[1] = p.cl; out8[2] = p.grid; out8[3] = p.sh.n_mp; out8[4] = p.sh.n_n; out8[5] = p.sh.n_s;
    out8[6] = p.sh.streamk; out8[7] = p.sh.units_per_cluster;
    return FSMG_OK;
}

// the two projection-backward GEMMs with the softmax gradient rebuilt on their A operand in shared memory (tc_gemm_kernel "XF"),
// on caller-provided buffers: parity test of the transform on its own.  a_mn_major = 0: C[M,N] = dl[M,K] * B[N,K]^T with
// dl[r, v] = E[r, v] * exp(cmaxT[v / 16, r] - lse[r]) - (v == y[r]) (M token rows, K vocabulary); a_mn_major = 1: C[M,N] +=
// alpha * dl^T * B with E stored [K tokens, lda >= M vocabulary], B stored [K tokens, ldb >= N], db[v] += alpha * column sums
// of dl (C and db are accumulated atomically: the caller zeroes them).
int fsmg_debug_gemm_xf(int32_t m, int32_t n, int32_t k, const void* d_e_f16, int64_t lda, const void* d_b_f16, int64_t ldb, float* d_c,
                       int32_t a_mn_major, const float* d_cmaxT, int64_t ld_cmax, const float* d_lse, const int32_t* d_y, float alpha,
                       float* d_db, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (m <= 0 || n <= 0 || k <= 0 || !d_e_f16 || !d_b_f16 || !d_c || !d_cmaxT || !d_lse || !d_y)
        return set_error(FSMG_ERR_INVALID, "fsmg_debug_gemm_xf: bad argument");
    GemmArgs g = a_mn_major ? mk(m, n, k, d_e_f16, lda, d_b_f16, ldb, d_c, n, alpha, nullptr, 0, 0, 1)
                            : mk(m, n, k, d_e_f16, lda, d_b_f16, ldb, d_c, n);
    static fsmg::TcContext ctx;
    int rc = tc_init(ctx);
    if (rc) return rc;
    if (!tc_xf_supported(ctx, g)) return set_error(FSMG_ERR_INVALID, "shape has no 256 x 512 pair-tile plan");
    XfArgs xf;
    xf.cmaxT = d_cmaxT; xf.ld_cmax = ld_cmax; xf.n_c16 = cdiv(a_mn_major ? m : k, 16); xf.lse = d_lse; xf.y = d_y; xf.row0 = 0;
    xf.db = a_mn_major ? d_db : nullptr;
    return tc_gemm(ctx, g, a_mn_major != 0, a_mn_major != 0, s, &xf);
}

// one in-place softmax-gradient pass (dlogits = exp(logit - lse) - onehot(y), db += alpha * column sums) over a caller-provided
// fp16 logits block, in an explicit kernel variant: parity test of the pass on its own + micro-benchmark of the variants
int fsmg_debug_softmax_grad(int32_t rows, int32_t vocab1, int64_t ld, void* d_logits_f16, const float* d_lse, const int32_t* d_y,
                            float alpha, float* d_db, int32_t mode, int32_t param, int32_t waves, void* stream) {
    if (rows <= 0 || vocab1 <= 0 || ld < vocab1 || (ld % 8) != 0 || !d_logits_f16 || !d_lse || !d_y || !d_db)
        return set_error(FSMG_ERR_INVALID, "fsmg_debug_softmax_grad: bad argument");
    int dev = 0, sms = 148;
    FSMG_CUDA_OK(cudaGetDevice(&dev));
    FSMG_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    return tc_softmax_grad_launch(sms, mode, param, waves, d_y, 0, rows, vocab1, (__half*)d_logits_f16, ld, d_lse, alpha, d_db, (cudaStream_t)stream);
}

// ---- device-side episode assembly -------------------------------------------------------------------------
int fsmg_gather_token_rows(const int32_t* d_corpus, int64_t n_corpus_rows, int32_t row_len, const int32_t* d_row_ids, int32_t n_ids,
                           int32_t* d_tokens_out, void* stream) {
    if (!d_corpus || !d_row_ids || !d_tokens_out || n_corpus_rows <= 0 || row_len <= 0 || n_ids <= 0)
        return set_error(FSMG_ERR_INVALID, "fsmg_gather_token_rows: bad argument");
    if ((reinterpret_cast<uintptr_t>(d_corpus) & 15) || (reinterpret_cast<uintptr_t>(d_tokens_out) & 15))
        return set_error(FSMG_ERR_INVALID, "fsmg_gather_token_rows: corpus and output must be 16-byte aligned");
    gather_token_rows_kernel<<<cdiv(n_ids, 8), 256, 0, (cudaStream_t)stream>>>(d_corpus, n_corpus_rows, row_len, d_row_ids, n_ids, d_tokens_out);
    FSMG_LAUNCH_OK();
    return FSMG_OK;
}

// ---- unigram baseline (reference src/models/unigram_model.py) ------------------------------------------------------------
int fsmg_unigram_step(float* d_counts, int32_t vocab, const int32_t* d_tokens, int32_t n_rows, int32_t row_len, int32_t col_begin,
                      int32_t col_end, int32_t update, float* d_scratch4, float* d_mean_nll, void* stream) {
    float* d_scratch2 = d_scratch4;
    if ((reinterpret_cast<uintptr_t>(d_scratch4) & 7) != 0) return set_error(FSMG_ERR_INVALID, "fsmg_unigram_step: scratch must be 8-byte aligned");
    if (!d_counts || !d_tokens || !d_scratch2 || !d_mean_nll) return set_error(FSMG_ERR_INVALID, "fsmg_unigram_step: null argument");
    if (vocab <= 0 || n_rows <= 0 || row_len <= 0 || col_begin < 0 || col_end > row_len || col_begin >= col_end)
        return set_error(FSMG_ERR_INVALID, "fsmg_unigram_step: bad shape (vocab=%d rows=%d row_len=%d cols=[%d,%d))", vocab, n_rows, row_len,
                         col_begin, col_end);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = (int64_t)n_rows * (col_end - col_begin);
    if (n <= (1 << 16)) {
        unigram_fused_kernel<<<1, 1024, 0, s>>>(d_counts, vocab, d_tokens, n_rows, row_len, col_begin, col_end, update, d_mean_nll);
    } else {
        int sms = 148, dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        unigram_sum_kernel<<<1, 1024, 0, s>>>(d_counts, vocab, d_scratch2);
        unigram_nll_kernel<<<sms * 4, 256, 0, s>>>(d_counts, d_tokens, n_rows, row_len, col_begin, col_end, d_scratch2);
        unigram_update_kernel<<<sms * 4, 256, 0, s>>>(update ? d_counts : nullptr, d_tokens, n_rows, row_len, col_begin, col_end, d_scratch2,
                                                      d_mean_nll);
    }
    FSMG_LAUNCH_OK();
    return FSMG_OK;
}

int fsmg_unigram_argmax(const float* d_counts, int32_t vocab, int32_t* d_out, void* stream) {
    if (!d_counts || !d_out || vocab <= 0) return set_error(FSMG_ERR_INVALID, "fsmg_unigram_argmax: bad argument");
    unigram_argmax_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_counts, vocab, d_out);
    FSMG_LAUNCH_OK();
    return FSMG_OK;
}

}  // extern "C"
