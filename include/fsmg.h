/*
 * fsmg.h — C-ABI of libfsmg.so, the B200 (sm_100a) engine for the episodic LSTM baseline of
 * AI-ON/Few-Shot-Music-Generation.
 *
 * The reference has no native boundary: its model/compute boundary is
 *     feed_dict -> tf.Session.run        (reference src/models/lstm_baseline.py:98-105,120-125,144-151)
 * inside models.lstm_baseline.LSTMBaseline.{train,eval,sample}.  Every entry point below
 * replaces one use of that boundary and cites it.  The Python class
 * `models.lstm_baseline.LSTMBaseline` (few-shot-music-generation_b200/src/models/lstm_baseline.py)
 * binds these symbols with ctypes; see INTEGRATION.md for the stub a maintainer adds.
 *
 * Conventions
 *   - plain C types only; every pointer named d_* is a DEVICE pointer owned by the caller
 *     (PyTorch tensors), every pointer named h_* is a HOST pointer owned by the caller;
 *   - the library allocates no device memory: parameters, Adam slots, gradients and the
 *     workspace are bound once with fsmg_bind();
 *   - all device work is enqueued on the cudaStream_t passed as `stream` (a void*; NULL = the
 *     legacy default stream); calls are asynchronous unless stated otherwise;
 *   - return value: 0 = FSMG_OK, negative = error (fsmg_last_error() gives the text); nothing
 *     throws across the ABI;
 *   - one handle per device; a handle is not thread-safe.
 *
 * Token layout: `tokens` is int32 [n_seqs, T] row-major, exactly the arrays
 * EpisodeSampler.get_episode() yields after flatten_first_two_dims
 * (reference src/models/base_model.py:57-60, src/data/episode.py:62-74).  The start-word shift of
 * convert_tokens_to_input_and_target (base_model.py:63-86) happens on the device.
 */
#ifndef FSMG_H_
#define FSMG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSMG_ABI_VERSION 1

enum {
    FSMG_OK = 0,
    FSMG_ERR_INVALID = -1,   /* bad argument / config */
    FSMG_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed */
    FSMG_ERR_STATE = -3,     /* call order (e.g. not bound) */
    FSMG_ERR_CAPACITY = -4   /* n_seqs / workspace larger than what was bound */
};

/* flags for fsmg_config.flags */
enum {
    FSMG_FLAG_SIMT_GEMM = 1,     /* debug: route every contraction through the fp32-accumulate SIMT GEMM
                                    instead of the tcgen05 kernels (same data flow and dtypes) */
    FSMG_FLAG_SIMT_RECURRENT = 2 /* debug: per-step launches instead of the persistent recurrent kernel */
};

/* Hyper-parameters read by the reference model:
 *   lstm_baseline.py:21-29 (input_size, max_len, embedding_size, hidden_size, n_layers, lr,
 *   max_grad_norm), :77-81 (n_decay); Adam defaults of tf.train.AdamOptimizer (:82). */
typedef struct fsmg_config {
    int32_t vocab;         /* config['input_size'] = V; start word = V; V' = V + 1 */
    int32_t embed;         /* embedding_size E */
    int32_t hidden;        /* hidden_size H */
    int32_t layers;        /* n_layers */
    int32_t max_len;       /* max_len T (time steps of the unrolled graph) */
    int32_t max_seqs;      /* capacity: sequences per call (B*(S+Q) * episodes_per_step) */
    int32_t n_decay;       /* exponential_decay steps */
    int32_t flags;         /* FSMG_FLAG_* */
    float lr;              /* base learning rate */
    float max_grad_norm;   /* clip_by_global_norm threshold */
    float beta1, beta2, eps; /* Adam (0.9, 0.999, 1e-8 in the reference) */
    float reserved;
} fsmg_config;

typedef struct fsmg_handle fsmg_handle;

/* One trainable tensor inside the flat fp32 parameter buffer, in the order
 * TFModel.get_vars() returns them (reference src/models/tf_model.py:99-104):
 * embedding, cell_l/kernel, cell_l/bias ..., softmax_w, softmax_b. */
typedef struct fsmg_param_info {
    char name[96];     /* TF variable name, e.g. "lstm_baseline/rnn/multi_rnn_cell/cell_0/basic_lstm_cell/kernel" */
    int64_t offset;    /* element offset into the flat buffer */
    int32_t rows, cols;/* cols = 1 for vectors */
} fsmg_param_info;

const char* fsmg_last_error(void);
int fsmg_abi_version(void);

/* Host-only construction (replaces TFModel.__init__ graph building, tf_model.py:80-97). */
int fsmg_create(const fsmg_config* cfg, const char* scope_name, fsmg_handle** out);
void fsmg_destroy(fsmg_handle* h);

/* Sizes the caller must allocate: flat parameter count (padded), gradient buffer count
 * (= params + FSMG_GRAD_EXTRA scalar slots), workspace bytes. */
#define FSMG_GRAD_EXTRA 8  /* [0]=sum of per-token NLL, [1]=per-occurrence embedding-grad square norm (TF clip quirk),
                              [2]=tokens in this call (summed by the data-parallel all-reduce: the global token count) */
int64_t fsmg_param_count(const fsmg_handle* h);
int64_t fsmg_grad_count(const fsmg_handle* h);
int64_t fsmg_workspace_bytes(const fsmg_handle* h);
int fsmg_num_params(const fsmg_handle* h);
int fsmg_param_info_at(const fsmg_handle* h, int index, fsmg_param_info* out);

/* Bind caller-owned device buffers (all fp32 except the opaque workspace).  Builds TMA
 * descriptors.  d_params/d_adam_m/d_adam_v: fsmg_param_count() floats; d_grads: fsmg_grad_count(). */
int fsmg_bind(fsmg_handle* h, float* d_params, float* d_grads, float* d_adam_m, float* d_adam_v,
              void* d_workspace, int64_t workspace_bytes);

/* Re-derive the fp16 operand copies of the weights from the fp32 master parameters.  Must be
 * called after the caller writes d_params directly (init / checkpoint restore); fsmg_apply_update
 * does it itself. */
int fsmg_refresh_weights(fsmg_handle* h, void* stream);

/* Forward + per-token NLL — LSTMBaseline.eval's sess.run(self._avg_neg_log)
 * (lstm_baseline.py:115-125) and the loss half of train (:104).
 *   d_tokens [n_seqs,T] int32; d_nll [n_seqs,T] fp32 (may be NULL);
 *   d_sum_nll: 1 float, receives sum over tokens (mean = sum / (n_seqs*T + 1e-12), A.5). */
int fsmg_forward_nll(fsmg_handle* h, const int32_t* d_tokens, int32_t n_seqs,
                     float* d_nll, float* d_sum_nll, void* stream);

/* Forward + backward (tf.gradients, lstm_baseline.py:83-84) into d_grads: dense gradients of
 * sum(nll) * loss_scale for every trainable, plus the two scalars of FSMG_GRAD_EXTRA.
 * loss_scale = 1/(global token count + 1e-12): for data-parallel shards the caller passes the
 * GLOBAL count, all-reduces d_grads (sum) and then calls fsmg_apply_update on every rank. */
int fsmg_forward_backward(fsmg_handle* h, const int32_t* d_tokens, int32_t n_seqs,
                          float loss_scale, float* d_nll, void* stream);

/* Overlapping the data-parallel all-reduce with the backward pass.  The caller hands in two of its own cudaEvent_t
 * (void*, may be NULL): fsmg_forward_backward records ev_softmax_grads on `stream` as soon as the softmax_w / softmax_b
 * gradients are final (after the projection backward, before the recurrent backward) and ev_embedding_grads as soon as
 * the embedding gradient is final; inside the library's CUDA graph they are external event-record nodes.  The caller
 * makes a side stream wait on them and all-reduces fsmg_param_range(2) / (0) of d_grads there while the rest of the
 * backward pass runs; ranges (1) and (3) follow after the call.  reserve_sms: SMs the cooperative persistent recurrent
 * kernels leave free so that the collective's few CTAs are never queued behind them (0 = use every SM).
 * fsmg_param_range: [begin, end) element ranges of the flat buffers: 0 = embedding, 1 = LSTM kernels + biases,
 * 2 = softmax_w + softmax_b, 3 = the FSMG_GRAD_EXTRA scalars (gradient buffer only). */
int fsmg_set_stage_events(fsmg_handle* h, void* ev_softmax_grads, void* ev_embedding_grads, int32_t reserve_sms);
/* Same mechanism for the loss: ev_loss_ready (a caller-owned cudaEvent_t, NULL to disable) is recorded inside
 * fsmg_forward_backward as soon as d_grads[param_count + 0] (sum of the per-token NLL) and [+ 2] (token count) are final —
 * after the forward pass and the projection, before the recurrent backward.  A single-GPU caller reads the step's loss back
 * from a side stream behind this event and returns it to the training loop (sess.run's fetched avg_neg_log,
 * lstm_baseline.py:104-105) while the backward pass and the update still run; every later call is ordered behind them on
 * `stream`. */
int fsmg_set_loss_event(fsmg_handle* h, void* ev_loss_ready);
int fsmg_param_range(const fsmg_handle* h, int32_t which, int64_t* begin, int64_t* end);

/* clip_by_global_norm + Adam + exponential_decay + global_step++ (lstm_baseline.py:77-87) on the
 * (already reduced) d_grads; refreshes the fp16 operand copies.  `step` = global_step before
 * this update.  d_out_norm (1 float, may be NULL) receives the global norm. */
int fsmg_apply_update(fsmg_handle* h, int64_t step, float* d_out_norm, void* stream);

/* Greedy autoregressive decode fully on the device — LSTMBaseline.sample's python loop
 * (lstm_baseline.py:135-156) for n_songs independent songs at once: start word V, zero state,
 * argmax with first-index tie-break.  d_out [n_songs, n_tokens] int32. */
int fsmg_sample_greedy(fsmg_handle* h, int32_t n_songs, int32_t n_tokens, int32_t* d_out, void* stream);

/* Host-buffer convenience entry points (what a ctypes/cgo caller with numpy/host arrays binds):
 * pinned staging + H2D of tokens, the step, D2H of the result, stream-synchronised on return. */
int fsmg_eval_host(fsmg_handle* h, const int32_t* h_tokens, int32_t n_seqs, float* h_mean_nll, float* h_nll, void* stream);
int fsmg_train_host(fsmg_handle* h, const int32_t* h_tokens, int32_t n_seqs, int64_t step, float* h_mean_loss, void* stream);
int fsmg_sample_host(fsmg_handle* h, int32_t n_songs, int32_t n_tokens, int32_t* h_out, void* stream);

/* Per-phase device timing for benchmarks: when enabled, every phase of forward_backward /
 * apply_update is bracketed by CUDA events on the caller's stream.  fsmg_read_profile synchronises
 * the stream, returns accumulated milliseconds and bracket counts per phase, and resets. */
#define FSMG_PROF_PHASES 11
int fsmg_set_profile(fsmg_handle* h, int enable);
int fsmg_read_profile(fsmg_handle* h, float* ms_out, int32_t* count_out, void* stream);
const char* fsmg_profile_phase_name(int phase);

/* Token ids outside [0, V] (TensorFlow's embedding_lookup raises InvalidArgument for them, lstm_baseline.py:41): the
 * host entry points reject them with FSMG_ERR_INVALID before anything is copied; the device-pointer entry points clamp
 * them (memory safety) and count them in a device flag.  fsmg_token_range_errors synchronises the stream, returns the
 * count accumulated since the last query and resets it. */
int fsmg_token_range_errors(fsmg_handle* h, int64_t* h_count, void* stream);

/* The device-side input/target shift on its own — convert_tokens_to_input_and_target with start word V
 * (reference src/models/base_model.py:63-86) — for parity tests: d_tokens [n_seqs, T] -> d_x_out, d_y_out, both
 * TIME-MAJOR [T, n_seqs] (element t*n_seqs + n), the layout every kernel of the path reads. */
int fsmg_debug_prep_tokens(fsmg_handle* h, const int32_t* d_tokens, int32_t n_seqs, int32_t* d_x_out, int32_t* d_y_out,
                           void* stream);

/* Introspection for benchmarks/tests: number of kernel launches issued by the last call,
 * and a GEMM self-test entry (C[M,N] = A[M,K] * B[N,K]^T, fp16 in, fp32 out) that drives the
 * tcgen05 core directly. */
int64_t fsmg_last_launch_count(const fsmg_handle* h);
int fsmg_debug_gemm(int32_t m, int32_t n, int32_t k, const void* d_a_f16, const void* d_b_f16,
                    float* d_c, int32_t a_mn_major, int32_t b_mn_major, int32_t use_simt, void* stream);

/* Hardware probe (test infrastructure): where tcgen05.mma.cta_group::2 places the rows of D in each CTA's tensor memory for
 * the given instruction shape.  D[r, n] = 256 r + n is computed by one CTA pair; d_out receives fp32 [2 CTAs][128 lanes][n]
 * (-1 = cell not written by the instruction).  The recurrent kernels' epilogues rely on this layout. */
int fsmg_debug_mma_probe(int32_t m, int32_t n, float* d_out, void* stream);

/* One in-place softmax-gradient pass over an fp16 logits block [rows, ld] (K7/K8 of the hot path, reference
 * lstm_baseline.py:70-75 + tf.gradients through sequence_loss): logits[r, v] <- exp(logits[r, v] - lse[r]) - (v == y[r]),
 * db[v] += alpha * sum_r of that, for v < vocab1 (columns vocab1..ld-1 are zeroed).  mode/param/waves select the kernel
 * variant (0 = strip grid, 1 = persistent strips, 2 = streaming; see csrc/tc_gemm.cuh): test + micro-benchmark hook. */
int fsmg_debug_softmax_grad(int32_t rows, int32_t vocab1, int64_t ld, void* d_logits_f16, const float* d_lse,
                            const int32_t* d_y, float alpha, float* d_db, int32_t mode, int32_t param, int32_t waves,
                            void* stream);

/* Host-side launch planning of the GEMM core for C[M,N] = A[M,K] * B[N,K]^T on a 148-SM B200, no device access (test hook).
 * out8 = { N tile width, CTAs per cluster, grid size, M pair-tiles, N tiles, K splits, stream-K flag, k-block units per cluster }.
 * allow_split: partial sums may be combined atomically (fp32 output, no accumulate); narrow: the decode loop's 128-wide tiles. */
int fsmg_debug_plan(int32_t m, int32_t n, int32_t k, int32_t allow_split, int32_t narrow, int32_t* out8);

/* The projection-backward GEMMs with the softmax gradient rebuilt on their A operand inside the kernel (no HBM pass; reference
 * lstm_baseline.py:70-75 + tf.gradients through sequence_loss and xw_plus_b), on caller-provided device buffers: test hook.
 * E holds exp(logit - max of the logit's 16-column chunk) in fp16, cmaxT[chunk, row] those maxima, lse/y the per-row
 * log-sum-exp / target.  a_mn_major = 0: C[M,N] = dl * B[N,K]^T, dl[r, v] = E[r, v] * exp(cmaxT[v / 16, r] - lse[r]) - (v == y[r])
 * (M rows, K vocabulary).  a_mn_major = 1: C[M,N] += alpha * dl^T * B with E stored [K rows, lda >= M vocabulary], B stored
 * [K rows, ldb >= N]; db[v] += alpha * column sums of dl (caller zeroes C and db). */
int fsmg_debug_gemm_xf(int32_t m, int32_t n, int32_t k, const void* d_e_f16, int64_t lda, const void* d_b_f16, int64_t ldb,
                       float* d_c, int32_t a_mn_major, const float* d_cmaxT, int64_t ld_cmax, const float* d_lse,
                       const int32_t* d_y, float alpha, float* d_db, void* stream);

/* ---- device-side episode assembly (replaces the host loop of EpisodeSampler.get_episode, reference
 * src/data/episode.py:62-74, once the tokenised corpus is resident in HBM) ---------------------------------------------
 * d_corpus is int32 [n_corpus_rows, row_len] (every song of the split, zero-padded to max_len like
 * base_loader.py:52-64); d_row_ids are the n_ids song rows of the step in batch order; d_tokens_out [n_ids, row_len]
 * is then a valid `tokens` argument for fsmg_forward_nll / fsmg_forward_backward.  Only the indices cross PCIe. */
int fsmg_gather_token_rows(const int32_t* d_corpus, int64_t n_corpus_rows, int32_t row_len, const int32_t* d_row_ids,
                           int32_t n_ids, int32_t* d_tokens_out, void* stream);

/* ---- unigram baseline: the reference's second registered model (src/models/unigram_model.py) -------------------------
 * d_counts is the `word_count` variable (:27-30: fp32 [vocab], initialised to alpha = 1 by the caller, not trainable).
 * fsmg_unigram_step replaces sess.run([train_op, avg_neg_log]) of train() (:41-55, update = 1, words = tokens[:, 0:T-1])
 * and sess.run(avg_neg_log) of eval() (:57-67, update = 0, words = tokens[:, 1:T]):
 *     *d_mean_nll = -mean(log(word_count[w] / sum(word_count)))  over tokens[:, col_begin:col_end]   (:35-37)
 *     then, if update: word_count[w] += 1 per occurrence                                              (:31-33)
 * The loss is taken on the counts BEFORE the update (TF1 leaves the order unspecified).  d_tokens is int32
 * [n_rows, row_len]; d_scratch4 is 4 floats (8-byte aligned) of device scratch.  fsmg_unigram_argmax replaces sample()'s
 * np.argmax(prob_all) (:69-78): first maximal index. */
int fsmg_unigram_step(float* d_counts, int32_t vocab, const int32_t* d_tokens, int32_t n_rows, int32_t row_len,
                      int32_t col_begin, int32_t col_end, int32_t update, float* d_scratch4, float* d_mean_nll,
                      void* stream);
int fsmg_unigram_argmax(const float* d_counts, int32_t vocab, int32_t* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FSMG_H_ */
