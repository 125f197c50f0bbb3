// common.cuh — context object, error plumbing and launch declarations shared by libnvd's
// translation units.  Nothing here is visible through the C ABI (include/nvd.h).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nvd.h"

namespace nvd {

constexpr int kMaxPeers = 16;                   // ranks of one NVLink domain
constexpr int kXchgValues = 448;                // uint64 values per exchange (two alternating data areas per block)
constexpr int kXchgBytes = 8192;                // [0, 3584) area 0 | [3584, 7168) area 1 | 7168: arrival counter


void set_error(const char* fmt, ...);

#define NVD_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::nvd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,     \
                       __LINE__);                                                            \
      return _e == cudaErrorMemoryAllocation ? NVD_ERR_OOM : NVD_ERR_CUDA;                   \
    }                                                                                        \
  } while (0)

#define NVD_CHECK_LAUNCH() NVD_CUDA(cudaGetLastError())

#define NVD_REQUIRE(cond, ...)                                                               \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      ::nvd::set_error(__VA_ARGS__);                                                         \
      return NVD_ERR_INVALID;                                                                \
    }                                                                                        \
  } while (0)

#define NVD_TRY(expr)                                                                        \
  do {                                                                                       \
    int _s = (expr);                                                                         \
    if (_s != NVD_OK) return _s;                                                             \
  } while (0)

constexpr int kMaxLayers = 8;

// A loaded feed-forward decoder (weights resident in HBM for the context's lifetime).
struct MlpModel {
  bool live = false;
  int n_layers = 0;
  int dims[kMaxLayers + 1] = {0};
  int acts[kMaxLayers] = {0};
  float* W[kMaxLayers] = {nullptr};  // fp32 [in, out] row-major
  float* b[kMaxLayers] = {nullptr};
  // tensor-core image: bf16 weights pre-arranged in the UMMA canonical K-major layout with the
  // bias rows folded in (k3_mlp_tc.cu); nullptr when the shape is not covered.
  void* tc_image = nullptr;
  size_t tc_image_bytes = 0;
  // layer-per-launch tensor-core path for wide layers (k_mlp_wide_tc.cu): one bf16 image per layer
  bool wide = false;
  void* wide_image[kMaxLayers] = {nullptr};
  // K-streamed tensor-core path for tall-and-skinny two-layer nets (k_mlp_kstream.cu): bf16 W1 image + fp32 tail
  void* ks_image = nullptr;
  float* ks_l2 = nullptr;
  int ks_Kpad = 0, ks_N1pad = 0;
  float* tc_bias_out = nullptr;  // reserved
  float tc_out_offset = 0.0f;    // power of two added to the output accumulators (positive-key argmax)
};

struct LstmModel {
  bool live = false;
  int H = 0;
  float* blob = nullptr;  // [enc Wp | enc b | dec Wp | dec b | lin_w | lin_b], gate-interleaved
  size_t blob_floats = 0;
  void* tc_image = nullptr;  // bf16 B operands of both cells + linear layer (tensor-core recurrence), H = 16 / 20
};

}  // namespace nvd

// state of one nvd_train_begin .. nvd_train_end run (k_train.cu): activations of a step, gradients, Adam moments
struct nvd_trainer {
  int model = -1;
  nvd_train_config cfg{};
  int64_t max_batch = 0;
  int64_t step = 0;
  float* act[nvd::kMaxLayers + 1] = {nullptr};   // act[l + 1] = activation of layer l, [max_batch, dims[l + 1]]
  float* delta[2] = {nullptr, nullptr};          // ping-pong d(activation) buffers, widest layer
  float* gW[nvd::kMaxLayers] = {nullptr};
  float* gb[nvd::kMaxLayers] = {nullptr};
  float* mW[nvd::kMaxLayers] = {nullptr};
  float* vW[nvd::kMaxLayers] = {nullptr};
  float* mb[nvd::kMaxLayers] = {nullptr};
  float* vb[nvd::kMaxLayers] = {nullptr};
  float* cost = nullptr;                         // device scalar
  bool live = false;
};

struct nvd_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  size_t total_mem = 0;
  uint64_t seed = 0;
  cudaStream_t stream = nullptr;
  std::vector<nvd::MlpModel> mlps;
  std::vector<nvd::LstmModel> lstms;
  std::vector<nvd_trainer> trainers;
  // device-side abort flag written by kernels whose bounded waits expire
  bool isolated_launches = false;  // set by callers that synchronise after every fused launch (k_mlp_tc.cu: launch_mlp_tc_fused)
  int* abort_flag = nullptr;       // device
  int* abort_flag_host = nullptr;  // pinned mirror
  // scratch owned by the context (grown on demand)
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  void* mc_buf = nullptr;                // staged Monte-Carlo buffers (info | coded | y | pred)
  size_t mc_buf_bytes = 0;
  void* pin_buf = nullptr;               // reserved pinned staging
  uint64_t* counters_scratch = nullptr;  // device uint64[4]
  uint64_t* counters_pinned = nullptr;   // pinned uint64[4]
  // host-buffer fused pipeline: copy-in / copy-out streams and per-slice events (created on first use)
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  std::vector<cudaEvent_t> slice_ev;
  uint32_t* slice_flags = nullptr;       // device: one word per input slice of the streaming host step
  uint32_t* slice_ones = nullptr;        // pinned: the value copied into a flag
  uint32_t* slice_done = nullptr;        // device: finished tiles per output slice
  uint32_t* done_flags = nullptr;        // mapped pinned: slice s of the decisions is complete in device memory
  uint32_t* done_flags_dev = nullptr;    // ... its device-side address
  static constexpr int kMaxSlices = 1024;
  // multi-GPU exchange step: NCCL communicator of this context's rank (comm.cu); null = single GPU
  void* comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  // ... and the same exchange over NVLink peer memory (comm.cu, nvd_comm_p2p_*): every rank owns one exchange block
  // (cudaMalloc, exported with CUDA IPC) and maps the blocks of all the others
  uint64_t* xchg_local = nullptr;                 // this rank's block (kXchgBytes)
  uint64_t* xchg_peer[nvd::kMaxPeers] = {};            // every rank's block as seen from here ([comm_rank] = xchg_local)
  int xchg_world = 0;                             // 0 = peer exchange not set up
  uint64_t xchg_epoch = 0;                        // exchanges queued so far (all ranks call them in the same order)
  // staged-pipeline stage timing
  cudaEvent_t stage_ev[6] = {nullptr};
  float stage_ms[5] = {0, 0, 0, 0, 0};
  bool stage_timing = false;
};

namespace nvd {

int ensure_scratch(nvd_ctx* ctx, size_t bytes);

// comm.cu: in-place sum of n uint64 device values over the context's communicator, on ctx->stream (no-op without one)
int comm_allreduce_u64(nvd_ctx* ctx, uint64_t* values_dev, size_t n);
void comm_release(nvd_ctx* ctx);
void trainers_release(nvd_ctx* ctx);            // k_train.cu

// ---- kernel launchers (each validates nothing; the API layer does) -----------------------
int launch_source(nvd_ctx* ctx, int L, uint64_t frame0, int64_t B, uint8_t* info_packed);
int launch_encode(nvd_ctx* ctx, int code, const uint8_t* info, int L, int64_t B, uint8_t* coded);
// row_stride: elements between rows of y (0 = dense, n); a stride that is a multiple of 8 bf16 lets the K-streamed
// decoder take the rows by TMA
int launch_channel(nvd_ctx* ctx, const uint8_t* coded, int n, int64_t B, float sigma, int snr_idx,
                   uint64_t frame0, int out_dtype, void* y, int64_t row_stride = 0);
int launch_mlp_fp32(nvd_ctx* ctx, const MlpModel& m, const void* y, int y_dtype, int64_t B,
                    float* soft_out, void* hard_out, int hard_bytes);
int launch_count(nvd_ctx* ctx, const void* pred, const void* label, int stride, int L, int64_t B,
                 int flags, uint64_t* counters);
int launch_lstm(nvd_ctx* ctx, const LstmModel& m, const void* y, int y_dtype, const float* dec_in,
                int T_enc, int T_dec, int64_t B, int fast_math, float* logits, uint8_t* hard);

// wide-layer tensor-core path (k_mlp_wide_tc.cu): widths <= 511, output <= 128 classes
bool wide_tc_supported(const MlpModel& m);
int wide_tc_build_image(MlpModel& m, const float* const* W_host, const float* const* b_host);
size_t wide_tc_scratch_bytes(const MlpModel& m, int64_t rows);
int launch_mlp_wide_tc(nvd_ctx* ctx, const MlpModel& m, const void* src, int src_dtype, int64_t src_len, int64_t row_stride,
                       int64_t B, void* scratch, float* soft_out, void* hard_out, int hard_bytes);

// K-streamed tensor-core path (k_mlp_kstream.cu): [K, N1 <= 64, n_out <= 8] nets such as the bitwise 1036-30-2
bool kstream_supported(const MlpModel& m);
int kstream_build_image(MlpModel& m, const float* const* W_host, const float* const* b_host);
// staged: bf16 rows [B, K] at `row_stride` elements -> soft [B, n_out] fp32 / hard class index
int launch_mlp_kstream(nvd_ctx* ctx, const MlpModel& m, const __nv_bfloat16* y, int64_t row_stride, int64_t B, float* soft_out,
                       void* hard_out, int hard_bytes);
// fused Monte-Carlo: frames generated on chip (prefix ++ Philox info bits -> encoder -> channel), decision compared
// with info bit `bit_position`, counters accumulated
int launch_mlp_kstream_fused(nvd_ctx* ctx, const MlpModel& m, int code, int L, uint32_t prefix_bits, int n_prefix, int bit_position,
                             int trunc, float sigma, int snr_idx, uint64_t frame0, uint64_t frames, uint64_t* counters);

// K0 helpers / K5 variants for long frames
int launch_prepend_bits(nvd_ctx* ctx, const uint8_t* info, int L, uint32_t prefix_bits, int n_prefix, int64_t B, uint8_t* out);
int launch_count_bit(nvd_ctx* ctx, const uint8_t* pred, const uint8_t* info, int info_stride, int bit_position, int64_t B,
                     uint64_t* counters);
int launch_count_rows(nvd_ctx* ctx, const uint8_t* pred, int pred_stride, const uint8_t* label, int label_stride, int L, int64_t B,
                      uint64_t* counters);
int launch_teacher_inputs(nvd_ctx* ctx, const uint8_t* info, int info_stride, int L, int T_dec, int64_t B, float* dec_in);

// tensor-core recurrence (k_lstm_tc.cu)
bool lstm_tc_supported(int H);
int lstm_tc_build_image(LstmModel& m, const float* enc_W, const float* enc_b, const float* dec_W, const float* dec_b,
                        const float* lin_w, const float* lin_b);
int launch_lstm_tc(nvd_ctx* ctx, const LstmModel& m, const void* y, int y_dtype, const float* dec_in,
                   int T_enc, int T_dec, int64_t B, float* logits, uint8_t* hard);

// tensor-core path (k3_mlp_tc.cu)
bool tc_shape_supported(const MlpModel& m);
int tc_build_image(nvd_ctx* ctx, MlpModel& m, const float* const* W_host, const float* const* b_host);
// staged: y bf16 [B, 14] -> soft/hard
int launch_mlp_tc_staged(nvd_ctx* ctx, const MlpModel& m, const __nv_bfloat16* y, int64_t B,
                         float* soft_out, void* hard_out, int hard_bytes);
// fused: frames [frame0, frame0+frames) generated on chip, counters accumulated
// info != nullptr: the info words come from that device buffer (one byte per frame) and the decisions go to
// pred (one byte per frame, may be null) instead of being a function of the frame index
// streaming host step: the info words arrive in n_in slices (slice i = frames [in_end[i-1], in_end[i]), valid once
// slice_flags[i] != 0) and the decisions are counted complete in n_out slices (slice_done / done_flags); every
// boundary except the last is a multiple of 128 frames
struct HostStepSlices {
  static constexpr int kMax = 16;
  int n_in = 0, n_out = 0;
  int64_t in_end[kMax] = {0}, out_end[kMax] = {0};
};
int launch_mlp_tc_fused(nvd_ctx* ctx, const MlpModel& m, int code, int L, float sigma, int snr_idx,
                        uint64_t frame0, uint64_t frames, uint64_t* counters, const uint8_t* info = nullptr,
                        uint8_t* pred = nullptr, const uint32_t* slice_flags = nullptr, const HostStepSlices* slices = nullptr,
                        uint32_t* slice_done = nullptr, uint32_t* done_flags = nullptr);

}  // namespace nvd
