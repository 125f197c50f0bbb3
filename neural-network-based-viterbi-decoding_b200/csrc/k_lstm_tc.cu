// K4 (throughput path): the LSTM seq2seq recurrence of build_seq2seq_model
// (/root/reference/src/rnn_viterbi.py:278-305; TF-1.x LSTMCell: z = [x, h] W + b, gates i, j, f, o,
// c' = sigmoid(f + 1) c + sigmoid(i) tanh(j), h' = sigmoid(o) tanh(c'); linear H -> 1 on the decoder
// outputs) as ONE persistent tcgen05 kernel: every time step of 128 frames is a
//     [128 x (1 + H + 1)] x [(1 + H + 1) x 4H]      bf16 x bf16 -> fp32
// contraction on the tensor core, with the recurrent operand never leaving the SM:
//   * weights of both cells (bf16, gate columns interleaved per unit, sigmoid's 1/2 and the forget bias
//     folded in, bias row met by a constant 1.0 in the A operand) stay in shared memory for the whole
//     launch;
//   * the cell state c stays in fp32 registers of the thread that owns the frame (thread r = frame r
//     = TMEM lane r) for all T_enc + T_dec steps; h goes back as the next step's A operand through
//     tcgen05.st into TENSOR memory (TS-mode MMA): no shared-memory round trip, no proxy fence;
//   * four independent 128-frame groups per CTA interleave, one group's MMA latency and hand-off
//     hiding behind the others' gate math (100 MUFU.TANH per frame and step: the bound).
// The fp32 kernel (k_lstm.cu) stays the parity path; this one is selected with fast_math = 2 and is
// held to the bf16 tolerance.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace nvd {

namespace {

constexpr int kGroups = 4;
constexpr int kThreads = 128 * kGroups;
constexpr int KP = 32;                 // padded contraction length: x, h[0..H), 1.0, zeros
constexpr uint32_t kMaxWaitMs = 2000;
constexpr uint32_t COL_D = 0, COL_A = 96;   // tensor-memory columns of a group: gates 0..4H-1, A operand 96..111

struct LstmTcParams {
  const uint8_t* image;   // [enc B operand | dec B operand] bf16, canonical K-major, then lin_w[H], lin_b (fp32)
  int* abort_flag;
  const void* y;
  int y_dtype;
  const float* dec_in;
  int T_enc, T_dec;
  int64_t B;
  float* logits;
  uint8_t* hard;
  int hard_stride;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float load_as_f32(const void* p, int dtype, int64_t i) {
  if (dtype == NVD_F32) return __ldg(static_cast<const float*>(p) + i);
  if (dtype == NVD_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
  return static_cast<float>(__ldg(static_cast<const int32_t*>(p) + i));
}

template <int H>
__global__ void __launch_bounds__(kThreads, 1) lstm_tc_kernel(const LstmTcParams p) {
  constexpr int N = 4 * H;
  constexpr int kCellBytes = N * KP * 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  const int g = warp >> 2, wq = warp & 3;
  const bool issuer_warp = wq == 0;
  const int r = tid & 127;
  constexpr int LIN_OFF = 2 * kCellBytes;
  constexpr int BAR_OFF = LIN_OFF + ((H + 1 + 3) & ~3) * 4;
  const uint32_t bar_mma = sbase + BAR_OFF + 8 * g;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 64);
  const float* s_lin = reinterpret_cast<const float*>(smem + LIN_OFF);

  for (int i = tid; i < (LIN_OFF + (H + 1) * 4) / 4; i += kThreads)
    reinterpret_cast<uint32_t*>(smem)[i] = __ldg(reinterpret_cast<const uint32_t*>(p.image) + i);
  if (tid == 0) {
    for (int i = 0; i < kGroups; ++i) ptx::mbar_init(sbase + BAR_OFF + 8 * i, 1);
    ptx::mbar_fence_init();
  }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xFFFFFFFFu, *tmem_slot, 0);
  const uint32_t tmem_grp = tmem_base + g * 128;
  const uint32_t tmem_d = tmem_grp + (static_cast<uint32_t>(wq * 32) << 16);
  constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, N);

  const int64_t num_tiles = (p.B + 127) >> 7;
  uint32_t mma_phase = 0;
  bool ok = true;
  for (int64_t tile = static_cast<int64_t>(blockIdx.x) * kGroups + g; ok && tile < num_tiles;
       tile += static_cast<int64_t>(gridDim.x) * kGroups) {
    const int64_t f = tile * 128 + r;
    const bool live = f < p.B;
    float c[H], h[H];
#pragma unroll
    for (int u = 0; u < H; ++u) { c[u] = 0.0f; h[u] = 0.0f; }
    float prev = 1.0f;   // EOS (rnn_viterbi.py:93)
    uint32_t bits = 0;
    float x = live ? load_as_f32(p.y, p.y_dtype, f * p.T_enc) : 0.0f;
    const int T = p.T_enc + p.T_dec;
    for (int t = 0; t < T; ++t) {
      const bool dec = t >= p.T_enc;
      const int td = t - p.T_enc;
      // ---- A operand of this step: (x, h0) (h1, h2) ... (h[H-1], 1.0) 0 ... as bf16 pairs, two per column
      {
        uint32_t w[16];
        w[0] = pack_bf16x2(x, h[0]);
#pragma unroll
        for (int i = 1; i < H / 2; ++i) w[i] = pack_bf16x2(h[2 * i - 1], h[2 * i]);
        w[H / 2] = pack_bf16x2(h[H - 1], 1.0f);
#pragma unroll
        for (int i = H / 2 + 1; i < 16; ++i) w[i] = 0u;
        ptx::tmem_st16(tmem_d + COL_A, w);
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::bar_sync(1 + g, 128);
      if (issuer_warp && ptx::elect_one()) {
        ptx::tc_fence_after();
        const uint32_t wbase = sbase + (dec ? kCellBytes : 0);
#pragma unroll
        for (int j = 0; j < KP / 16; ++j)
          ptx::umma_f16_ts(tmem_grp + COL_D, tmem_grp + COL_A + 8 * j,
                           ptx::make_smem_desc(wbase + j * 2 * (N * 16), N * 16, 128), idesc, j > 0);
        ptx::umma_commit(bar_mma);
      }
      // next step's input while the MMA runs
      float x_next = 0.0f;
      if (t + 1 < p.T_enc) x_next = live ? load_as_f32(p.y, p.y_dtype, f * p.T_enc + t + 1) : 0.0f;
      else if (t + 1 < T && p.dec_in != nullptr) x_next = live ? __ldg(p.dec_in + f * p.T_dec + (t + 1 - p.T_enc)) : 0.0f;
      if (!ptx::mbar_wait_bounded(bar_mma, mma_phase, kMaxWaitMs)) { atomicExch(p.abort_flag, 20); ok = false; break; }
      mma_phase ^= 1;
      ptx::tc_fence_after();
      // ---- gates: 16 columns = 4 units x (i/2, j, (f+1)/2, o/2) ---------------------------------
#pragma unroll
      for (int q = 0; q < H / 4; ++q) {
        uint32_t v[16];
        ptx::tmem_ld16(tmem_d + COL_D + 16 * q, v);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int u = 4 * q + i;
          const float si = fmaf(ptx::tanh_approx(__uint_as_float(v[4 * i])), 0.5f, 0.5f);
          const float tj = ptx::tanh_approx(__uint_as_float(v[4 * i + 1]));
          const float sf = fmaf(ptx::tanh_approx(__uint_as_float(v[4 * i + 2])), 0.5f, 0.5f);
          const float so = fmaf(ptx::tanh_approx(__uint_as_float(v[4 * i + 3])), 0.5f, 0.5f);
          c[u] = fmaf(sf, c[u], si * tj);
          h[u] = so * ptx::tanh_approx(c[u]);
        }
      }
      ptx::tc_fence_before();   // gate reads done before the next step's MMA overwrites them (after the barrier)
      if (dec) {
        float lg = s_lin[H];
#pragma unroll
        for (int u = 0; u < H; ++u) lg = fmaf(h[u], s_lin[u], lg);
        const bool one = lg > 0.0f;
        bits |= (one ? 1u : 0u) << (td & 31);
        prev = one ? 1.0f : -1.0f;
        if (p.logits != nullptr && live) p.logits[f * p.T_dec + td] = lg;
        if (p.hard != nullptr && live && ((td & 31) == 31 || td == p.T_dec - 1)) {
          const int byte0 = (td >> 5) << 2;
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (byte0 + b < p.hard_stride) p.hard[f * p.hard_stride + byte0 + b] = static_cast<uint8_t>(bits >> (8 * b));
          bits = 0;
        }
      }
      x = (t + 1 >= p.T_enc && p.dec_in == nullptr) ? prev : x_next;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 512);
}

uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

// B operand [N x KP] of one cell: row n = 4u + gate (i, j, f, o), column k = 0: x, 1..H: h, H+1: bias
void fill_cell(uint8_t* img, int H, const float* W /* [1+H, 4H] */, const float* b /* [4H] */) {
  const int N = 4 * H;
  for (int u = 0; u < H; ++u)
    for (int gt = 0; gt < 4; ++gt) {
      const int n = 4 * u + gt, col = gt * H + u;
      const float scale = (gt == 1) ? 1.0f : 0.5f;   // sigmoid(z) = 0.5 tanh(z / 2) + 0.5
      for (int k = 0; k < KP; ++k) {
        float v = 0.0f;
        if (k <= H) v = scale * W[static_cast<size_t>(k) * 4 * H + col];
        else if (k == H + 1) v = scale * (b[col] + (gt == 2 ? 1.0f : 0.0f));   // forget_bias = 1
        const uint16_t q = f32_to_bf16_rn(v);
        memcpy(img + static_cast<size_t>(k >> 3) * (N * 16) + n * 16 + (k & 7) * 2, &q, 2);
      }
    }
}

template <int H>
int launch_h(nvd_ctx* ctx, const LstmTcParams& p) {
  constexpr int N = 4 * H;
  const int smem = 2 * N * KP * 2 + ((H + 1 + 3) & ~3) * 4 + 128;
  auto kern = lstm_tc_kernel<H>;
  NVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int64_t tiles = (p.B + 127) / 128;
  const int grid = static_cast<int>(std::min<int64_t>(ctx->sm_count, (tiles + kGroups - 1) / kGroups));
  kern<<<grid, kThreads, smem, ctx->stream>>>(p);
  NVD_CHECK_LAUNCH();
  return NVD_OK;
}

}  // namespace

bool lstm_tc_supported(int H) { return H == 16 || H == 20; }

int lstm_tc_build_image(LstmModel& m, const float* enc_W, const float* enc_b, const float* dec_W, const float* dec_b,
                        const float* lin_w, const float* lin_b) {
  const int H = m.H, N = 4 * H;
  const size_t cell = static_cast<size_t>(N) * KP * 2;
  std::vector<uint8_t> img(2 * cell + static_cast<size_t>((H + 1 + 3) & ~3) * 4, 0);
  fill_cell(img.data(), H, enc_W, enc_b);
  fill_cell(img.data() + cell, H, dec_W, dec_b);
  memcpy(img.data() + 2 * cell, lin_w, static_cast<size_t>(H) * 4);
  memcpy(img.data() + 2 * cell + static_cast<size_t>(H) * 4, lin_b, 4);
  NVD_CUDA(cudaMalloc(&m.tc_image, img.size()));
  NVD_CUDA(cudaMemcpy(m.tc_image, img.data(), img.size(), cudaMemcpyHostToDevice));
  return NVD_OK;
}

int launch_lstm_tc(nvd_ctx* ctx, const LstmModel& m, const void* y, int y_dtype, const float* dec_in, int T_enc,
                   int T_dec, int64_t B, float* logits, uint8_t* hard) {
  if (B == 0) return NVD_OK;
  if (!m.tc_image) { set_error("LSTM hidden size %d has no tensor-core image (16 or 20)", m.H); return NVD_ERR_UNSUPPORTED; }
  LstmTcParams p{};
  p.image = static_cast<const uint8_t*>(m.tc_image);
  p.abort_flag = ctx->abort_flag;
  p.y = y;
  p.y_dtype = y_dtype;
  p.dec_in = dec_in;
  p.T_enc = T_enc;
  p.T_dec = T_dec;
  p.B = B;
  p.logits = logits;
  p.hard = hard;
  p.hard_stride = (T_dec + 7) / 8;
  return m.H == 16 ? launch_h<16>(ctx, p) : launch_h<20>(ctx, p);
}

}  // namespace nvd
