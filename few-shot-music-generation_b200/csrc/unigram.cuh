// unigram.cuh — the reference's second registered model (src/models/unigram_model.py:26-39) on the device:
//   word_count[v] (fp32, initialised to alpha = 1, never trained by gradients)
//   train : loss = -mean(log(word_count[w] / sum(word_count))) over the fed words, word_count[w] += 1 per occurrence
//   eval  : the same mean NLL, no update
//   sample: argmax(word_count / sum) (first maximal index), `num` times
// The work is a histogram (atomics into an L2-resident 40 KB table) and a gather/log/mean: HBM/latency-bound byte work, no GEMM.
// Which of the two the reference evaluates first inside `sess.run([train_op, avg_neg_log])` is unspecified in TF1 (no control
// dependency between the scatter_add and the gather): this path defines it as LOSS FIRST, on the counts before the update.
#pragma once
#include "common.cuh"

namespace fsmg {

// fp32 sum of the table in a fixed order (thread-strided partials, then a tree): deterministic, |error| < 1e-6 relative at V = 10k
__device__ __forceinline__ float unigram_table_sum(const float* __restrict__ counts, int vocab, float* red) {
    float s = 0.0f;
    for (int v = threadIdx.x; v < vocab; v += blockDim.x) s += counts[v];
    return block_sum(s, red);
}

// One CTA does a whole call when the token block is small (one episode: 45 x 49 words): sum, NLL, update, in that order.
__global__ void __launch_bounds__(1024) unigram_fused_kernel(float* __restrict__ counts, int vocab, const int32_t* __restrict__ tokens,
                                                             int n_rows, int row_len, int col_begin, int col_end, int update,
                                                             float* __restrict__ out_mean_nll) {
    __shared__ float red[32];
    const float total = unigram_table_sum(counts, vocab, red);
    const float log_total = logf(total);
    const int w = col_end - col_begin;
    const int64_t n = (int64_t)n_rows * w;
    float acc = 0.0f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const int r = (int)(i / w), c = (int)(i % w) + col_begin;
        const int tok = tokens[(int64_t)r * row_len + c];
        acc += log_total - logf(counts[tok]);            // -log(count / total)
    }
    acc = block_sum(acc, red);                           // (barriers inside: every read of `counts` above precedes the update)
    if (threadIdx.x == 0) *out_mean_nll = acc / (float)n;
    if (update)
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
            const int r = (int)(i / w), c = (int)(i % w) + col_begin;
            atomicAdd(counts + tokens[(int64_t)r * row_len + c], 1.0f);
        }
}

// Large token blocks: three launches (table sum -> NLL partials -> histogram update), grid sized to the SM count.
__global__ void __launch_bounds__(1024) unigram_sum_kernel(const float* __restrict__ counts, int vocab, float* __restrict__ scratch) {
    __shared__ float red[32];
    const float total = unigram_table_sum(counts, vocab, red);
    if (threadIdx.x == 0) { scratch[0] = total; *reinterpret_cast<double*>(scratch + 2) = 0.0; }   // [2..3]: fp64 sum of the NLL partials
}
__global__ void __launch_bounds__(256) unigram_nll_kernel(const float* __restrict__ counts, const int32_t* __restrict__ tokens, int n_rows,
                                                          int row_len, int col_begin, int col_end, float* __restrict__ scratch) {
    __shared__ float red[32];
    const float log_total = logf(scratch[0]);
    const int w = col_end - col_begin;
    const int64_t n = (int64_t)n_rows * w;
    float acc = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / w), c = (int)(i % w) + col_begin;
        acc += log_total - logf(counts[tokens[(int64_t)r * row_len + c]]);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(reinterpret_cast<double*>(scratch + 2), (double)acc);
}
__global__ void __launch_bounds__(256) unigram_update_kernel(float* __restrict__ counts, const int32_t* __restrict__ tokens, int n_rows,
                                                             int row_len, int col_begin, int col_end, const float* __restrict__ scratch,
                                                             float* __restrict__ out_mean_nll) {
    const int w = col_end - col_begin;
    const int64_t n = (int64_t)n_rows * w;
    if (blockIdx.x == 0 && threadIdx.x == 0) *out_mean_nll = (float)(*reinterpret_cast<const double*>(scratch + 2) / (double)n);
    if (counts == nullptr) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / w), c = (int)(i % w) + col_begin;
        atomicAdd(counts + tokens[(int64_t)r * row_len + c], 1.0f);
    }
}

// first maximal index of the table (np.argmax semantics, reference unigram_model.py:70-72)
__global__ void __launch_bounds__(1024) unigram_argmax_kernel(const float* __restrict__ counts, int vocab, int32_t* __restrict__ out) {
    __shared__ float sv[32];
    __shared__ int si[32];
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int v = threadIdx.x; v < vocab; v += blockDim.x) {
        const float c = counts[v];
        if (c > best || (c == best && v < bi)) { best = c; bi = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sv[w] = best; si[w] = bi; }
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        best = lane < nw ? sv[lane] : -INFINITY;
        bi = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) *out = bi;
    }
}

}  // namespace fsmg
