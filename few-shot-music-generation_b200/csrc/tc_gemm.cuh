// tc_gemm.cuh — tcgen05 / TMA / TMEM route of libfsmg (sm_100a).  STUB: filled in next.
#pragma once
#include "common.cuh"
#include "simt_kernels.cuh"

namespace fsmg {
struct Bump;
struct TcContext { bool ready = false; };
template <typename B> static inline void tc_carve(TcContext&, B&, int, int, int, int, int, int) {}
static inline int tc_init(TcContext& c) { c.ready = true; return 0; }
static inline bool tc_gemm_supported(const GemmArgs&, bool, bool) { return false; }
static inline int tc_gemm(TcContext&, const GemmArgs&, bool, bool, cudaStream_t) { return set_error(-1, "tc_gemm stub"); }
static inline bool tc_recurrent_supported(TcContext&, int, int) { return false; }
static inline int tc_lstm_forward(TcContext&, const float*, const __half*, __half*, float*, __half*, int, int, int, int, int, cudaStream_t) { return set_error(-1, "stub"); }
static inline int tc_lstm_backward(TcContext&, const float*, const __half*, const __half*, const float*, __half*, int, int, int, int, cudaStream_t) { return set_error(-1, "stub"); }
static inline bool tc_projection_supported(TcContext&, int, int) { return false; }
static inline int tc_projection_fwd(TcContext&, const __half*, int64_t, const __half*, int64_t, const float*, const int32_t*, int64_t, int, int, int, int, int, __half*, int64_t, float*, float*, cudaStream_t) { return set_error(-1, "stub"); }
static inline bool tc_sampler_supported(TcContext&, int, int) { return false; }
static inline int tc_sample_greedy(TcContext&, int, int, int32_t*, cudaStream_t) { return set_error(-1, "stub"); }
}  // namespace fsmg
