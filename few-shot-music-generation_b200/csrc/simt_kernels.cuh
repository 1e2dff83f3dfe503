// simt_kernels.cuh — CUDA-core kernels of libfsmg: the element-wise / gather / reduction stages
// of the hot path (used by every route) and a generic fp32-accumulate SIMT GEMM that serves
// (a) as the debug route for every contraction (FSMG_FLAG_SIMT_GEMM) and (b) as the fp32-exact
// contraction of the greedy sampler, where argmax ties need fp32-grade logits.
//
// Reference semantics: src/models/lstm_baseline.py:38-87,135-156 and SURVEY.md Appendix A.
#pragma once
#include "common.cuh"

namespace fsmg {

// ============================================================================================
// Generic SIMT GEMM:  C[m,n] (op)= alpha * sum_k A(m,k) * B(n,k) (+ bias[n])
//   A is K-major (A[m*lda+k]) or MN-major (A[k*lda+m]); same for B (B[n*ldb+k] / B[k*ldb+n]).
//   fp32 accumulation in a fixed k order.
// ============================================================================================
struct GemmArgs {
    int M, N, K;
    const void* A; int64_t lda;
    const void* B; int64_t ldb;
    void* C; int64_t ldc;
    const float* bias;   // per-n, may be null
    float alpha;
    int c_half;          // 1: C is __half, 0: float
    int accumulate;      // 1: C += (float C only, non-atomic read-modify-write)
    int atomic;          // 1: atomicAdd into float C
};

__device__ __forceinline__ float ld_elem(const float* p, int64_t i) { return p[i]; }
__device__ __forceinline__ float ld_elem(const __half* p, int64_t i) { return __half2float(p[i]); }

template <typename TA, typename TB, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(256) simt_gemm_kernel(GemmArgs g) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const TA* A = reinterpret_cast<const TA*>(g.A);
    const TB* B = reinterpret_cast<const TB*>(g.B);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            int i = tid + p * 256;
            int mm, kk;
            if (A_MN) { mm = i & 63; kk = i >> 6; } else { kk = i & 15; mm = i >> 4; }
            int gm = m0 + mm, gk = k0 + kk;
            float v = 0.0f;
            if (gm < g.M && gk < g.K) v = A_MN ? ld_elem(A, (int64_t)gk * g.lda + gm) : ld_elem(A, (int64_t)gm * g.lda + gk);
            As[kk][mm] = v;
            int nn;
            if (B_MN) { nn = i & 63; kk = i >> 6; } else { kk = i & 15; nn = i >> 4; }
            int gn = n0 + nn; gk = k0 + kk;
            v = 0.0f;
            if (gn < g.N && gk < g.K) v = B_MN ? ld_elem(B, (int64_t)gk * g.ldb + gn) : ld_elem(B, (int64_t)gn * g.ldb + gk);
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gm = m0 + ty * 4 + i;
        if (gm >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
            if (gn >= g.N) continue;
            float v = g.alpha * acc[i][j];
            if (g.bias) v += g.bias[gn];
            int64_t idx = (int64_t)gm * g.ldc + gn;
            if (g.c_half) {
                reinterpret_cast<__half*>(g.C)[idx] = __float2half_rn(v);
            } else {
                float* c = reinterpret_cast<float*>(g.C);
                if (g.atomic) atomicAdd(c + idx, v);
                else if (g.accumulate) c[idx] += v;
                else c[idx] = v;
            }
        }
    }
}

template <typename TA, typename TB>
static inline void launch_simt_gemm(const GemmArgs& g, bool a_mn, bool b_mn, cudaStream_t s) {
    dim3 grid(cdiv(g.N, 64), cdiv(g.M, 64));
    if (!a_mn && !b_mn) simt_gemm_kernel<TA, TB, false, false><<<grid, 256, 0, s>>>(g);
    else if (a_mn && b_mn) simt_gemm_kernel<TA, TB, true, true><<<grid, 256, 0, s>>>(g);
    else if (a_mn) simt_gemm_kernel<TA, TB, true, false><<<grid, 256, 0, s>>>(g);
    else simt_gemm_kernel<TA, TB, false, true><<<grid, 256, 0, s>>>(g);
}

// ============================================================================================
// Token preparation: convert_tokens_to_input_and_target (reference base_model.py:63-86) on the
// device, re-ordered time-major: r = t*N + n.   x[r] = t ? tok[n,t-1] : V ;  y[r] = tok[n,t]
// ============================================================================================
__global__ void prep_tokens_kernel(const int32_t* __restrict__ tok, int32_t* __restrict__ x,
                                   int32_t* __restrict__ y, int N, int T, int start_word, int vmax, int* __restrict__ range_errors,
                                   float* __restrict__ token_count = nullptr) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (int64_t)N * T) return;
    if (r == 0 && token_count) *token_count = (float)((int64_t)N * T);   // piggy-backed on the gradient all-reduce (FSMG_GRAD_EXTRA slot 2)
    int t = (int)(r / N), n = (int)(r % N);
    int cur = tok[(int64_t)n * T + t];
    int prev = t ? tok[(int64_t)n * T + t - 1] : start_word;
    // ids outside [0, V] would index out of the tables (TensorFlow raises InvalidArgument): count them in a device flag the
    // host entry points turn into an error, and clamp so that the step itself stays memory-safe
    if ((cur < 0 || cur > vmax) && range_errors) atomicAdd(range_errors, 1);
    cur = min(max(cur, 0), vmax);
    prev = min(max(prev, 0), vmax);
    x[r] = prev;
    y[r] = cur;
}

// embedding_lookup (lstm_baseline.py:41): out[r,:] = table16[ids[r],:]  (fp16 rows, 16-byte vectors)
__global__ void gather_rows_f16_kernel(const __half* __restrict__ table, int64_t ld_table,
                                       const int32_t* __restrict__ ids, __half* __restrict__ out,
                                       int64_t ld_out, int64_t rows, int cols8) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = rows * cols8;
    if (i >= total) return;
    int64_t r = i / cols8;
    int c = (int)(i % cols8);
    const uint4* src = reinterpret_cast<const uint4*>(table + (int64_t)ids[r] * ld_table) + c;
    reinterpret_cast<uint4*>(out + r * ld_out)[c] = __ldg(src);
}

// fp32 gather (sampler): out[r,:] = table[ids[r],:]
__global__ void gather_rows_f32_kernel(const float* __restrict__ table, int64_t ld_table,
                                       const int32_t* __restrict__ ids, float* __restrict__ out,
                                       int64_t ld_out, int rows, int cols) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)rows * cols) return;
    int r = (int)(i / cols), c = (int)(i % cols);
    out[(int64_t)r * ld_out + c] = table[(int64_t)ids[r] * ld_table + c];
}

// ============================================================================================
// BasicLSTMCell element-wise stage ([TF-lib] A.2): gates i,j,f,o ; forget_bias 1
//   c = c_prev*sigmoid(f+1) + sigmoid(i)*tanh(j) ; h = tanh(c)*sigmoid(o)
// G: pre-activations fp32 [N, ldg] (columns [i|j|f|o], each H wide).
// ============================================================================================
template <typename TH>
__global__ void lstm_pointwise_fwd_kernel(const float* __restrict__ G, int64_t ldg,     // recurrent part (or the whole thing), may be null
                                          const __half* __restrict__ pre16, int64_t ldp, // hoisted x*Wx+b in fp16, may be null
                                          const float* __restrict__ c_prev,   // [N,H] or null (zeros)
                                          __half* __restrict__ gates_out, int64_t ldgo,  // may be null
                                          float* __restrict__ c_out,          // [N,H]
                                          TH* __restrict__ h_out, int64_t ldh, int N, int H) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * H) return;
    int n = (int)(idx / H), u = (int)(idx % H);
    float gi = 0.0f, gj = 0.0f, gf = 0.0f, go = 0.0f;
    if (G) {
        const float* g = G + (int64_t)n * ldg;
        gi = g[u]; gj = g[H + u]; gf = g[2 * H + u]; go = g[3 * H + u];
    }
    if (pre16) {
        const __half* pp = pre16 + (int64_t)n * ldp;
        gi += __half2float(pp[u]); gj += __half2float(pp[H + u]); gf += __half2float(pp[2 * H + u]); go += __half2float(pp[3 * H + u]);
    }
    float i_ = sigmoidf_(gi);
    float j_ = tanhf_(gj);
    float f_ = sigmoidf_(gf + 1.0f);
    float o_ = sigmoidf_(go);
    float cp = c_prev ? c_prev[(int64_t)n * H + u] : 0.0f;
    float c = cp * f_ + i_ * j_;
    float h = tanhf_(c) * o_;
    c_out[(int64_t)n * H + u] = c;
    if (sizeof(TH) == 2) reinterpret_cast<__half*>(h_out)[(int64_t)n * ldh + u] = __float2half_rn(h);
    else reinterpret_cast<float*>(h_out)[(int64_t)n * ldh + u] = h;
    if (gates_out) {
        __half* go = gates_out + (int64_t)n * ldgo;
        go[u] = __float2half_rn(i_);
        go[H + u] = __float2half_rn(j_);
        go[2 * H + u] = __float2half_rn(f_);
        go[3 * H + u] = __float2half_rn(o_);
    }
}

// Reverse-time cell backward for one step t (what tf.gradients derives for A.2):
//   dh = dh_out[t] + dh_rec ; do = dh*tanh(c) ; dc = dh*o*(1-tanh(c)^2) + dc_next
//   dgi = dc*j*i(1-i) ; dgj = dc*i*(1-j^2) ; dgf = dc*c_prev*f(1-f) ; dgo = do*o(1-o) ; dc_next' = dc*f
// Activations are "unscaled" (loss = sum nll); the 1/(N*T) factor is applied in the weight-grad epilogues.
__global__ void lstm_pointwise_bwd_kernel(const float* __restrict__ dh_out, int64_t ld_dho,  // [N,H] block t
                                          const float* __restrict__ dh_rec,                  // [N,H] or null
                                          const __half* __restrict__ gates, int64_t ldg,      // block t
                                          const float* __restrict__ c, const float* __restrict__ c_prev,  // block t, t-1 (null -> 0)
                                          float* __restrict__ dc_next,                        // [N,H] in/out
                                          __half* __restrict__ dgates, int64_t lddg, int N, int H, int first) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * H) return;
    int n = (int)(idx / H), u = (int)(idx % H);
    const __half* g = gates + (int64_t)n * ldg;
    float i_ = __half2float(g[u]), j_ = __half2float(g[H + u]);
    float f_ = __half2float(g[2 * H + u]), o_ = __half2float(g[3 * H + u]);
    float dh = dh_out[(int64_t)n * ld_dho + u] + (dh_rec ? dh_rec[(int64_t)n * H + u] : 0.0f);
    float tc = tanhf_(c[(int64_t)n * H + u]);
    float cp = c_prev ? c_prev[(int64_t)n * H + u] : 0.0f;
    float dcn = first ? 0.0f : dc_next[(int64_t)n * H + u];
    float d_o = dh * tc;
    float dc = dh * o_ * (1.0f - tc * tc) + dcn;
    __half* dg = dgates + (int64_t)n * lddg;
    dg[u] = __float2half_rn(dc * j_ * i_ * (1.0f - i_));
    dg[H + u] = __float2half_rn(dc * i_ * (1.0f - j_ * j_));
    dg[2 * H + u] = __float2half_rn(dc * cp * f_ * (1.0f - f_));
    dg[3 * H + u] = __float2half_rn(d_o * o_ * (1.0f - o_));
    dc_next[(int64_t)n * H + u] = dc * f_;
}

// ============================================================================================
// Row-wise softmax / NLL over materialised fp32 logits (SIMT route): one CTA per token row.
//   lse = logsumexp(logits[r,:V']) ; nll = lse - logits[r,y]   ([TF-lib] A.5, natural log)
//   if dlogits: dlogits[r,v] = exp(logit-lse) - [v==y]   (unscaled, fp16)
// nll is written sequence-major: nll_out[n*T + t] with r = t*N + n (row0 = global row of chunk row 0).
// ============================================================================================
__global__ void rowwise_nll_kernel(const float* __restrict__ logits, int64_t ld, int vp1,
                                   const int32_t* __restrict__ y, int64_t row0, int N, int T,
                                   float* __restrict__ lse_out, float* __restrict__ nll_out,
                                   __half* __restrict__ dlogits, int64_t ldd) {
    __shared__ float red[32];
    int64_t lr = blockIdx.x;          // row inside the chunk
    int64_t r = row0 + lr;            // global time-major row
    const float* row = logits + lr * ld;
    float m = -INFINITY;
    for (int v = threadIdx.x; v < vp1; v += blockDim.x) m = fmaxf(m, row[v]);
    m = block_max(m, red);
    float s = 0.0f;
    for (int v = threadIdx.x; v < vp1; v += blockDim.x) s += expf(row[v] - m);
    s = block_sum(s, red);
    float lse = m + logf(s);
    int tgt = y[r];
    if (threadIdx.x == 0) {
        if (lse_out) lse_out[r] = lse;
        int t = (int)(r / N), n = (int)(r % N);
        if (nll_out) nll_out[(int64_t)n * T + t] = lse - row[tgt];
    }
    if (dlogits) {
        __half* d = dlogits + lr * ldd;
        for (int v = threadIdx.x; v < vp1; v += blockDim.x) {
            float p = expf(row[v] - lse) - (v == tgt ? 1.0f : 0.0f);
            d[v] = __float2half_rn(p);
        }
        for (int64_t v = vp1 + threadIdx.x; v < ldd; v += blockDim.x) d[v] = __float2half_rn(0.0f);
    }
}

// Column sums of an fp16 matrix [rows, cols] (ld) scaled by alpha, atomically added to out[cols]:
// bias gradients (db = sum_r dgates[r,:], db_s = sum_r dlogits[r,:]).
__global__ void colsum_f16_kernel(const __half* __restrict__ X, int64_t ld, int64_t rows, int cols,
                                  float alpha, float* __restrict__ out, int rows_per_block) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    int64_t r1 = min(rows, r0 + rows_per_block);
    if (c >= cols) return;
    float s = 0.0f;
    for (int64_t r = r0; r < r1; ++r) s += __half2float(X[r * ld + c]);
    atomicAdd(out + c, alpha * s);
}

__global__ void colsum_f32_kernel(const float* __restrict__ X, int64_t ld, int64_t rows, int cols, float alpha, float* __restrict__ out,
                                  int rows_per_block) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    int64_t r1 = min(rows, r0 + rows_per_block);
    if (c >= cols) return;
    float s = 0.0f;
    for (int64_t r = r0; r < r1; ++r) s += X[r * ld + c];
    atomicAdd(out + c, alpha * s);
}

// Embedding gradient: IndexedSlices(values = dX[r,:], indices = x[r]) densified by scatter-add,
// plus the per-occurrence square norm TF's clip_by_global_norm sees ([TF-lib] A.6).
__global__ void scatter_emb_grad_kernel(const float* __restrict__ dX, int64_t ld, const int32_t* __restrict__ x,
                                        int64_t rows, int E, float alpha, float* __restrict__ gemb,
                                        float* __restrict__ occ_sq) {
    __shared__ float red[32];
    int64_t r = blockIdx.x;
    float sq = 0.0f;
    if (r < rows) {
        const float* src = dX + r * ld;
        float* dst = gemb + (int64_t)x[r] * E;
        for (int e = threadIdx.x; e < E; e += blockDim.x) {
            float v = alpha * src[e];
            atomicAdd(dst + e, v);
            sq += v * v;
        }
    }
    sq = block_sum(sq, red);
    if (threadIdx.x == 0) atomicAdd(occ_sq, sq);
}

// sum of a float array -> atomicAdd into out (grid-stride)
__global__ void sum_f32_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
    __shared__ float red[32];
    float s = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += x[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(out, s);
}

// sum of squares over [begin, end) of a float array, ORDER-DETERMINISTIC: every block writes its partial (fixed grid-stride order,
// fixed shuffle tree) to partials[blockIdx.x]; clip_adam_kernel adds the partials in a fixed order.  Data-parallel replicas hold
// bit-identical all-reduced gradients, so they derive bit-identical clip factors and never drift apart (an atomicAdd combine
// differed by an ulp between ranks).
constexpr int SQNORM_BLOCKS = 296;
__global__ void sqnorm_f32_kernel(const float* __restrict__ x, int64_t begin, int64_t end, float* __restrict__ partials) {
    __shared__ float red[32];
    float s = 0.0f;
    for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (int64_t)gridDim.x * blockDim.x) {
        float v = x[i];
        s = fmaf(v, v, s);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
// fixed-order sum of the SQNORM_BLOCKS partials by one warp, in double (same result in every block, on every rank)
__device__ __forceinline__ float sqnorm_combine(const float* __restrict__ partials) {
    double acc = 0.0;
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < SQNORM_BLOCKS; i += 32) acc += (double)partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return (float)acc;
}

// ============================================================================================
// clip_by_global_norm + TF-Adam ([TF-lib] A.6, A.7) over the flat buffers.
//   scalars[0] = dense square norm (all trainables except the embedding), scalars[1] = occ square norm
//   scale = clip / max(norm, clip) ; m,v update ; theta -= alpha_t * m / (sqrt(v) + eps)
// alpha_t (lr schedule + bias correction, A.7/A.8) is computed on the host in fp32/double.
// ============================================================================================
__global__ void clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, int64_t n, const float* __restrict__ dense_sq,
                                 const float* __restrict__ occ_sq, float clip, float alpha_t, float beta1,
                                 float beta2, float eps, float* __restrict__ norm_out) {
    // dense_sq: SQNORM_BLOCKS per-block partials of sqnorm_f32_kernel, combined here in a fixed order
    __shared__ float dense_total;
    if (threadIdx.x < 32) {
        const float t = sqnorm_combine(dense_sq);
        if (threadIdx.x == 0) dense_total = t;
    }
    __syncthreads();
    float norm = sqrtf(dense_total + *occ_sq);
    float scale = clip / fmaxf(norm, clip);
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
    auto upd = [&](float& pi, float gi_raw, float& mi, float& vi) {
        const float gi = gi_raw * scale;
        mi = beta1 * mi + (1.0f - beta1) * gi;
        vi = beta2 * vi + (1.0f - beta2) * gi * gi;
        pi = pi - alpha_t * mi / (sqrtf(vi) + eps);
    };
    // 28 bytes of HBM traffic per parameter: 16-byte accesses (the flat buffers are 256-B aligned and padded to multiples of 64)
    const int64_t n4 = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15) == 0 ? n / 4 : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
        const float4 g4 = __ldcs(reinterpret_cast<const float4*>(g) + i);
        upd(p4.x, g4.x, m4.x, v4.x); upd(p4.y, g4.y, m4.y, v4.y); upd(p4.z, g4.z, m4.z, v4.z); upd(p4.w, g4.w, m4.w, v4.w);
        reinterpret_cast<float4*>(m)[i] = m4;
        reinterpret_cast<float4*>(v)[i] = v4;
        reinterpret_cast<float4*>(p)[i] = p4;
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float pi = p[i], mi = m[i], vi = v[i];
        upd(pi, g[i], mi, vi);
        m[i] = mi; v[i] = vi; p[i] = pi;
    }
}

// fp32 master -> fp16 operand copies.  out[r*ldo + c] = in[r*ldi + c]  (row-major copy, pad cols zeroed)
__global__ void convert_f16_kernel(const float* __restrict__ in, int64_t ldi, __half* __restrict__ out, int64_t ldo,
                                   int rows, int cols) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)rows * ldo) return;
    int r = (int)(i / ldo), c = (int)(i % ldo);
    out[i] = __float2half_rn(c < cols ? in[(int64_t)r * ldi + c] : 0.0f);
}
// out[c*ldo + r] = in[r*ldi + c]  for r in [0,rows), c in [0,cols)   (32x32 smem transpose)
__global__ void transpose_f16_kernel(const float* __restrict__ in, int64_t ldi, __half* __restrict__ out, int64_t ldo,
                                     int rows, int cols) {
    __shared__ float tile[32][33];
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int r = r0 + j, c = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (r < rows && c < cols) ? in[(int64_t)r * ldi + c] : 0.0f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int c = c0 + j, r = r0 + threadIdx.x;
        if (c < cols && r < ldo) out[(int64_t)c * ldo + r] = __float2half_rn(r < rows ? tile[threadIdx.x][j] : 0.0f);
    }
}

// ============================================================================================
// Greedy decode helpers (lstm_baseline.py:152-153): first-index argmax per row of fp32 logits.
// ============================================================================================
__global__ void argmax_rows_kernel(const float* __restrict__ logits, int64_t ld, int cols,
                                   int32_t* __restrict__ next_ids, int32_t* __restrict__ out, int64_t out_stride,
                                   int64_t out_off) {
    __shared__ float sv[32];
    __shared__ int si[32];
    int r = blockIdx.x;
    const float* row = logits + (int64_t)r * ld;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        float v = row[c];
        if (v > best || (v == best && c < bi)) { best = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sv[w] = best; si[w] = bi; }
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        best = lane < nw ? sv[lane] : -INFINITY;
        bi = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) {
            next_ids[r] = bi;
            out[(int64_t)r * out_stride + out_off] = bi;
        }
    }
}

// dst[r*ldd + c] = src[r*lds + c]  (drop the row padding of an accumulation buffer)
__global__ void unpad_rows_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int rows, int cols) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)rows * cols) return;
    int r = (int)(i / cols), c = (int)(i % cols);
    dst[(int64_t)r * ldd + c] = src[(int64_t)r * lds + c];
}

// Episode assembly on the device: out[i, :] = table[ids[i], :] for int32 token rows — the step's token batch is
// gathered from a corpus that is resident in HBM, so only the song indices cross PCIe (reference data/episode.py:62-74 builds the
// same [B, S+Q, T] arrays on the host, song by song).  One warp per row, 16-byte accesses when the row length allows.
__global__ void gather_token_rows_kernel(const int32_t* __restrict__ table, int64_t n_table_rows, int row_len,
                                         const int32_t* __restrict__ ids, int n_ids, int32_t* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n_ids) return;
    int64_t src = ids[row];
    if (src < 0 || src >= n_table_rows) src = 0;      // ids are validated on the host; never read out of bounds
    const int32_t* in = table + src * row_len;
    int32_t* o = out + (int64_t)row * row_len;
    if ((row_len & 3) == 0) {
        const int4* in4 = reinterpret_cast<const int4*>(in);
        int4* o4 = reinterpret_cast<int4*>(o);
        for (int c = lane; c < (row_len >> 2); c += 32) o4[c] = in4[c];
    } else {
        for (int c = lane; c < row_len; c += 32) o[c] = in[c];
    }
}

// ---- token-sorted segment sums (layer-0 input gradients without per-token GEMM work) --------------------------------------------
// The input of layer 0 is an embedding ROW, so its weight gradient and the dense embedding gradient only need, per distinct word v,
//     S[v, :] = sum over the tokens r with x[r] == v of dgates[r, :]                      ([V', 4H], at most V' non-zero rows)
// then dK[:E] = embedding^T * S and dEmbedding = S * K[:E]^T are GEMMs over V' rows instead of N*T tokens (10 001 vs 184 320 at cfg 2).
// Counting sort of the token rows by word (histogram -> exclusive scan -> fill), then a segmented sum that walks the sorted order.
// (warp-aggregated: lanes holding the same word elect one leader — Zipf-distributed ids put a tenth of all tokens on one counter)
__global__ void token_hist_kernel(const int32_t* __restrict__ x, int64_t n, int32_t* __restrict__ counts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = i < n;
    const unsigned active = __ballot_sync(0xffffffffu, ok);
    if (!ok) return;
    const int v = x[i];
    const unsigned same = __match_any_sync(active, v);
    if ((int)(threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(counts + v, __popc(same));
}
// single block: offsets[v] = sum_{u<v} counts[u]; cursor = copy of offsets (consumed by the fill pass)
__global__ void __launch_bounds__(1024) token_scan_kernel(const int32_t* __restrict__ counts, int vocab, int32_t* __restrict__ offsets,
                                                          int32_t* __restrict__ cursor) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < vocab; base += 1024) {
        const int v = base + threadIdx.x;
        const int c = v < vocab ? counts[v] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_tot[w] = incl;
        __syncthreads();
        if (w == 0) {
            int t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += u; }
            warp_tot[lane] = t;     // inclusive totals of the warps
        }
        __syncthreads();
        const int excl = carry + (w ? warp_tot[w - 1] : 0) + incl - c;
        if (v < vocab) { offsets[v] = excl; cursor[v] = excl; }
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[vocab] = carry;
}
__global__ void token_fill_kernel(const int32_t* __restrict__ x, int64_t n, int32_t* __restrict__ cursor, int32_t* __restrict__ sorted_rows,
                                  int32_t* __restrict__ sorted_tok) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = i < n;
    const unsigned active = __ballot_sync(0xffffffffu, ok);
    if (!ok) return;
    const int v = x[i];
    const int lane = threadIdx.x & 31;
    const unsigned same = __match_any_sync(active, v);
    const int leader = __ffs(same) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(cursor + v, __popc(same));
    base = __shfl_sync(same, base, leader);
    const int pos = base + __popc(same & ((1u << lane) - 1u));
    sorted_rows[pos] = (int32_t)i;
    sorted_tok[pos] = v;
}
// CTA (bx, by): columns [bx*1024, +1024) (8 per thread) of the sorted positions [by*RPB, +RPB); a thread keeps the running sum of the
// current word in registers and flushes it to S with atomics only when the word changes (or at the end of its slice)
template <int RPB>
__global__ void __launch_bounds__(128) segsum_rows_kernel(const __half* __restrict__ X, int64_t ldx, int cols, const int32_t* __restrict__ sorted_tok,
                                                          const int32_t* __restrict__ sorted_rows, int64_t n, float* __restrict__ S, int64_t lds) {
    const int c0 = (blockIdx.x * 128 + threadIdx.x) * 8;
    if (c0 >= cols) return;
    const int64_t p0 = (int64_t)blockIdx.y * RPB;
    const int m = (int)((n - p0) < RPB ? (n - p0) : RPB);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
    int cur = -1;
    auto flush = [&]() {
        if (cur < 0) return;
        float* dst = S + (int64_t)cur * lds + c0;
        if (c0 + 8 <= cols && (lds & 3) == 0) {     // two 16-byte vector reductions instead of eight scalar atomics
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(acc[0]), "f"(acc[1]), "f"(acc[2]), "f"(acc[3]) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(acc[4]), "f"(acc[5]), "f"(acc[6]), "f"(acc[7]) : "memory");
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (c0 + e < cols && acc[e] != 0.0f) atomicAdd(dst + e, acc[e]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
    };
    constexpr int B = 8;
    for (int i0 = 0; i0 < m; i0 += B) {
        uint4 raw[B];
        int tok[B];
#pragma unroll
        for (int b = 0; b < B; ++b)
            if (i0 + b < m) {
                const int row = __ldg(sorted_rows + p0 + i0 + b);
                tok[b] = __ldg(sorted_tok + p0 + i0 + b);
                raw[b] = __ldcs(reinterpret_cast<const uint4*>(X + (int64_t)row * ldx + c0));
            }
#pragma unroll
        for (int b = 0; b < B; ++b)
            if (i0 + b < m) {
                if (tok[b] != cur) { flush(); cur = tok[b]; }
                const __half2* h2 = reinterpret_cast<const __half2*>(&raw[b]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = __half22float2(h2[q]);
                    acc[2 * q] += f.x;
                    acc[2 * q + 1] += f.y;
                }
            }
    }
    flush();
}

// one pass over the fp32 segment sums: fp16 operand copy (pad columns zeroed) + column sums (the layer's bias gradient)
template <int RPB>
__global__ void __launch_bounds__(128) seg_finish_kernel(const float* __restrict__ S, int64_t lds, int rows, int cols, __half* __restrict__ out,
                                                         int64_t ldo, float alpha, float* __restrict__ colsum) {
    const int c0 = (blockIdx.x * 128 + threadIdx.x) * 8;
    if (c0 >= ldo) return;
    const int r0 = blockIdx.y * RPB, r1 = min(rows, r0 + RPB);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
    const bool vec = c0 + 8 <= cols && (lds & 3) == 0;
    constexpr int RB = 4;      // rows in flight per thread
    for (int rb = r0; rb < r1; rb += RB) {
        float v[RB][8];
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const int r = rb + b;
            if (r < r1 && vec) {
                const float4 a = __ldcs(reinterpret_cast<const float4*>(S + (int64_t)r * lds + c0));
                const float4 c = __ldcs(reinterpret_cast<const float4*>(S + (int64_t)r * lds + c0 + 4));
                v[b][0] = a.x; v[b][1] = a.y; v[b][2] = a.z; v[b][3] = a.w; v[b][4] = c.x; v[b][5] = c.y; v[b][6] = c.z; v[b][7] = c.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[b][e] = (r < r1 && c0 + e < cols) ? S[(int64_t)r * lds + c0 + e] : 0.0f;
            }
        }
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const int r = rb + b;
            if (r >= r1) break;
            __align__(16) __half2 h2[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                h2[q] = __floats2half2_rn(v[b][2 * q], v[b][2 * q + 1]);
                acc[2 * q] += v[b][2 * q];
                acc[2 * q + 1] += v[b][2 * q + 1];
            }
            *reinterpret_cast<uint4*>(out + (int64_t)r * ldo + c0) = *reinterpret_cast<const uint4*>(h2);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (c0 + e < cols) atomicAdd(colsum + c0 + e, alpha * acc[e]);
}

// dst[c*ldd + r] = src[r*lds + c] : 32 x 32 tiles through shared memory, both sides coalesced
__global__ void transpose_f32_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int rows, int cols) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(int64_t)r * lds + c] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[(int64_t)c * ldd + r] = tile[threadIdx.x][i];
    }
}

// ============================================================================================
// Split-fp16 operands for fp32-grade tensor-core contractions (greedy sampler):
//   x = hi + 2^-11 * lo ,  hi = fp16(x) ,  lo = fp16((x - hi) * 2^11)      (about 22 mantissa bits)
//   a.b ~= 2^-11 * ( (2^11 a_hi).b_hi + a_hi.b_lo + a_lo.b_hi )            (ONE fp32-accumulating GEMM over 3K)
// Activations are stored [rows, 3*Kp] = [2^11*hi | hi | lo], weights [N, 3*Kp] = [hi | lo | hi]; the GEMM epilogue
// applies alpha = 2^-11.  2^11*hi is exact in fp16 for |x| < 32 (activations live in (-1,1), embeddings are small).
// ============================================================================================
__device__ __forceinline__ void split_hi_lo(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn((x - __half2float(hi)) * 2048.0f);
}
// activations: in [rows, cols] fp32 (ldi) -> out [rows, 3*Kp], pads zero
__global__ void split_act_kernel(const float* __restrict__ in, int64_t ldi, __half* __restrict__ out, int Kp, int rows, int cols) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)rows * Kp) return;
    int r = (int)(i / Kp), c = (int)(i % Kp);
    __half hi = __float2half_rn(0.0f), lo = hi;
    if (c < cols) split_hi_lo(in[(int64_t)r * ldi + c], hi, lo);
    __half* o = out + (int64_t)r * 3 * Kp;
    o[c] = __float2half_rn(__half2float(hi) * 2048.0f);
    o[Kp + c] = hi;
    o[2 * Kp + c] = lo;
}
// weights given as [K rows, N cols] fp32 (TF layout, ldi) -> out [N, 3*Kp] = [hi | lo | hi] of in[k, n]
__global__ void split_weight_t_kernel(const float* __restrict__ in, int64_t ldi, __half* __restrict__ out, int Kp, int K, int N) {
    __shared__ float tile[32][33];
    int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int kk = k0 + j, nn = n0 + threadIdx.x;
        tile[j][threadIdx.x] = (kk < K && nn < N) ? in[(int64_t)kk * ldi + nn] : 0.0f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int nn = n0 + j, kk = k0 + threadIdx.x;
        if (nn < N && kk < Kp) {
            __half hi = __float2half_rn(0.0f), lo = hi;
            if (kk < K) split_hi_lo(tile[threadIdx.x][j], hi, lo);
            __half* o = out + (int64_t)nn * 3 * Kp;
            o[kk] = hi;
            o[Kp + kk] = lo;
            o[2 * Kp + kk] = hi;
        }
    }
}

// Sampler cell step: gates = G[n,:] (+ P[word[n],:] when P != null: the embedding row already multiplied by Wx, bias
// included) -> c (in place, fp32) -> h -> split [hi | lo] for the next contractions.
//
// Programmatic dependent launch (decode loop, FSMG_SAMPLE_PDL): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor still runs; griddep_wait() blocks until the predecessor has completed and its writes
// are visible (a no-op for ordinary launches), griddep_launch() lets the successor's CTAs become resident behind this grid.
// Every kernel of the decode step calls wait first and launch right after, so at most two grids overlap: the running one and the
// next one's prologue (block scheduling, TMEM allocation, barrier initialisation, descriptor prefetch).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// scalar variant (one unit per thread) for hidden sizes that are not a multiple of 4
__global__ void sample_cell_scalar_kernel(float* __restrict__ G, int64_t ldg, const float* __restrict__ P, int64_t ldp,
                                          const int32_t* __restrict__ words, const float* __restrict__ bias,
                                          float* __restrict__ c_state, __half* __restrict__ h_split, int Hp, int n, int H) {
    griddep_wait();
    griddep_launch();
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * H) return;
    int r = (int)(idx / H), u = (int)(idx % H);
    float* g = G + (int64_t)r * ldg;
    float gi = g[u], gj = g[H + u], gf = g[2 * H + u], go = g[3 * H + u];
    g[u] = 0.0f; g[H + u] = 0.0f; g[2 * H + u] = 0.0f; g[3 * H + u] = 0.0f;   // consumed: left zeroed for the next step's accumulating GEMM
    if (P) {
        const float* p = P + (int64_t)words[r] * ldp;
        gi += p[u]; gj += p[H + u]; gf += p[2 * H + u]; go += p[3 * H + u];
    } else if (bias) {
        gi += bias[u]; gj += bias[H + u]; gf += bias[2 * H + u]; go += bias[3 * H + u];
    }
    const float i_ = sigmoidf_(gi), j_ = tanhf_(gj), f_ = sigmoidf_(gf + 1.0f), o_ = sigmoidf_(go);
    const float c = c_state[(int64_t)r * H + u] * f_ + i_ * j_;
    c_state[(int64_t)r * H + u] = c;
    const float h = tanhf_(c) * o_;
    __half hi, lo;
    split_hi_lo(h, hi, lo);
    __half* o = h_split + (int64_t)r * 3 * Hp;
    o[u] = __float2half_rn(__half2float(hi) * 2048.0f);
    o[Hp + u] = hi;
    o[2 * Hp + u] = lo;
}
__global__ void sample_cell_kernel(float* __restrict__ G, int64_t ldg, const float* __restrict__ P, int64_t ldp,
                                   const int32_t* __restrict__ words, const float* __restrict__ bias,
                                   float* __restrict__ c_state, __half* __restrict__ h_split, int Hp, int n, int H) {
    // four consecutive units per thread (16-byte loads of the four gate blocks, the per-word table row and the cell state; 8-byte stores of
    // the three fp16 planes): 4x fewer threads, 4x the bytes in flight per thread.  H % 4 == 0 and 16-byte aligned rows are checked by the caller.
    griddep_wait();
    griddep_launch();
    const int H4 = H >> 2;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * H4) return;
    const int r = (int)(idx / H4), u = (int)(idx % H4) * 4;
    float* g = G + (int64_t)r * ldg + u;
    float4 gi = *reinterpret_cast<const float4*>(g), gj = *reinterpret_cast<const float4*>(g + H);
    float4 gf = *reinterpret_cast<const float4*>(g + 2 * H), go = *reinterpret_cast<const float4*>(g + 3 * H);
    {   // consumed: left zeroed for the next step's accumulating (stream-K, RED) GEMM — no memset node per token
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(g) = z; *reinterpret_cast<float4*>(g + H) = z;
        *reinterpret_cast<float4*>(g + 2 * H) = z; *reinterpret_cast<float4*>(g + 3 * H) = z;
    }
    const float* add = P ? P + (int64_t)words[r] * ldp + u : bias ? bias + u : nullptr;
    float4 c4 = *reinterpret_cast<const float4*>(c_state + (int64_t)r * H + u);
    if (add) {
        const float4 ai = __ldg(reinterpret_cast<const float4*>(add)), aj = __ldg(reinterpret_cast<const float4*>(add + H));
        const float4 af = __ldg(reinterpret_cast<const float4*>(add + 2 * H)), ao = __ldg(reinterpret_cast<const float4*>(add + 3 * H));
        gi.x += ai.x; gi.y += ai.y; gi.z += ai.z; gi.w += ai.w;
        gj.x += aj.x; gj.y += aj.y; gj.z += aj.z; gj.w += aj.w;
        gf.x += af.x; gf.y += af.y; gf.z += af.z; gf.w += af.w;
        go.x += ao.x; go.y += ao.y; go.z += ao.z; go.w += ao.w;
    }
    const float vi[4] = {gi.x, gi.y, gi.z, gi.w}, vj[4] = {gj.x, gj.y, gj.z, gj.w}, vf[4] = {gf.x, gf.y, gf.z, gf.w}, vo[4] = {go.x, go.y, go.z, go.w};
    float vc[4] = {c4.x, c4.y, c4.z, c4.w};
    __align__(8) __half p0[4], p1[4], p2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float i_ = sigmoidf_(vi[e]), j_ = tanhf_(vj[e]), f_ = sigmoidf_(vf[e] + 1.0f), o_ = sigmoidf_(vo[e]);
        vc[e] = vc[e] * f_ + i_ * j_;
        const float h = tanhf_(vc[e]) * o_;
        __half hi, lo;
        split_hi_lo(h, hi, lo);
        p0[e] = __float2half_rn(__half2float(hi) * 2048.0f);
        p1[e] = hi;
        p2[e] = lo;
    }
    *reinterpret_cast<float4*>(c_state + (int64_t)r * H + u) = make_float4(vc[0], vc[1], vc[2], vc[3]);
    __half* o = h_split + (int64_t)r * 3 * Hp + u;
    *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(p0);
    *reinterpret_cast<uint2*>(o + Hp) = *reinterpret_cast<const uint2*>(p1);
    *reinterpret_cast<uint2*>(o + 2 * Hp) = *reinterpret_cast<const uint2*>(p2);
}

// argmax with a device-resident step counter (so a captured graph of one decode step can be replayed); the last block to finish
// advances the counter (no separate bump kernel).  step_counter[0] = current step, step_counter[1] = arrival ticket of this launch.
// (Zeroing the consumed logits / gate rows here and in the cell kernel, instead of the memset in front of each split-K GEMM, was
// measured 3-5x slower per kernel: 12 -> 60 us and 6 -> 17 us under ncu, 49 -> 69 us per token.)
__global__ void __launch_bounds__(256) argmax_rows_step_kernel(float* __restrict__ logits, int64_t ld, int cols, int32_t* __restrict__ next_ids,
                                                               int32_t* __restrict__ out, int64_t out_stride, int* __restrict__ step_counter) {
    __shared__ float sv[32];
    __shared__ int si[32];
    griddep_wait();
    griddep_launch();
    int r = blockIdx.x;
    float* row = logits + (int64_t)r * ld;       // consumed rows are left zeroed for the next step's accumulating GEMM
    float best = -INFINITY;
    int bi = 0x7fffffff;
    auto take = [&](float v, int c) { if (v > best || (v == best && c < bi)) { best = v; bi = c; } };
    // 16-byte loads, VEC_ITERS of them in flight per thread before the first compare (the scalar loop was one dependent L2 round trip per
    // element: 14.8 us for 256 x 4709 logits in the ncu capture of round 2); rows start 16-byte aligned when ld % 4 == 0
    constexpr int VEC_ITERS = 4;
    const int cols4 = ((ld & 3) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0) ? cols / 4 : 0;
    for (int c0 = threadIdx.x; c0 < cols4; c0 += VEC_ITERS * blockDim.x) {
        float4 v[VEC_ITERS];
#pragma unroll
        for (int i = 0; i < VEC_ITERS; ++i) {
            const int c = c0 + i * blockDim.x;
            v[i] = c < cols4 ? __ldcs(reinterpret_cast<const float4*>(row) + c) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
#pragma unroll
        for (int i = 0; i < VEC_ITERS; ++i) {
            const int c = c0 + i * blockDim.x;
            if (c < cols4) reinterpret_cast<float4*>(row)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < VEC_ITERS; ++i) {
            const int c = 4 * (c0 + i * blockDim.x);
            take(v[i].x, c); take(v[i].y, c + 1); take(v[i].z, c + 2); take(v[i].w, c + 3);
        }
    }
    for (int c = cols4 * 4 + threadIdx.x; c < cols; c += blockDim.x) { take(row[c], c); row[c] = 0.0f; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sv[w] = best; si[w] = bi; }
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        best = lane < nw ? sv[lane] : -INFINITY;
        bi = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) {
            const int step = *reinterpret_cast<volatile int*>(step_counter);
            next_ids[r] = bi;
            out[(int64_t)r * out_stride + step] = bi;
            __threadfence();
            if (atomicAdd(step_counter + 1, 1) == (int)gridDim.x - 1) {   // every block has read `step`: advance it
                step_counter[1] = 0;
                step_counter[0] = step + 1;
            }
        }
    }
}

__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace fsmg
