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

#ifndef NVD_LSTM_GROUPS
#define NVD_LSTM_GROUPS 5
#endif
constexpr int kGroups = NVD_LSTM_GROUPS;
constexpr int kGroupCols = kGroups == 4 ? 128 : 96;       // tensor-memory columns per group
constexpr int kThreads = 128 * kGroups;
constexpr int KP = 32;                 // padded contraction length: x, h[0..H), 1.0, zeros
constexpr uint32_t kMaxWaitMs = 2000;
constexpr int kWin = 32;                                  // time steps per staged input / logit window
constexpr int kTileBytes = 128 * (kWin + 1) * 4;          // [128 frames][32 steps] fp32, padded rows
constexpr uint32_t COL_D = 0, COL_A = kGroups == 4 ? 96 : 80;   // tensor-memory columns of a group: gates 0..4H-1, A operand 96..111

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

// tanh of two accumulators on the FMA pipe (packed fp32x2), same polynomial as the feed-forward decoder
// (k_mlp_tc.cu, tools/fit_tanh_poly.py): tanh(x) ~ sign(x) u Q(u), u = min(|x| / 3.5, 1), max |error| 1.14e-3
// Of the 8 gate pairs of every 4 units ((i, j) and (f, o) per unit) this many are evaluated by the polynomial,
// (i, j) pairs first; tanh(c') always goes to the MUFU.  Measured on B200 (tools/lstm_bench.py, H = 20, T = 200 + 200):
// 0: 89 Mframes/s, 4: 112 Mframes/s, see DESIGN.md for the rest of the sweep.
#ifndef NVD_LSTM_POLY_PAIRS
#define NVD_LSTM_POLY_PAIRS 4
#endif
__device__ __forceinline__ uint64_t dup2(float c) {
  const uint32_t u = __float_as_uint(c);
  return (static_cast<uint64_t>(u) << 32) | u;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ void tanh_pair_poly(uint32_t xa, uint32_t xb, float& ta, float& tb) {
  float ua, ub;
  asm("{\n\t.reg .f32 t;\n\tabs.f32 t, %1;\n\tmul.rn.sat.f32 %0, t, 0f3E924925;\n\t}" : "=f"(ua) : "f"(__uint_as_float(xa)));
  asm("{\n\t.reg .f32 t;\n\tabs.f32 t, %1;\n\tmul.rn.sat.f32 %0, t, 0f3E924925;\n\t}" : "=f"(ub) : "f"(__uint_as_float(xb)));
  const uint64_t u = (static_cast<uint64_t>(__float_as_uint(ub)) << 32) | __float_as_uint(ua);
  uint64_t q = fma2(u, dup2(8.383429305e+00f), dup2(-2.840123474e+01f));
  q = fma2(q, u, dup2(3.558778774e+01f));
  q = fma2(q, u, dup2(-1.784809705e+01f));
  q = fma2(q, u, dup2(-2.634002747e-01f));
  q = fma2(q, u, dup2(3.540826384e+00f));
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(q), "l"(u));
  uint32_t pa = static_cast<uint32_t>(r), pb = static_cast<uint32_t>(r >> 32);
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(pa) : "r"(xa), "r"(pa));
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(pb) : "r"(xb), "r"(pb));
  ta = __uint_as_float(pa);
  tb = __uint_as_float(pb);
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
  constexpr int TILE_OFF = LIN_OFF + ((H + 1 + 3) & ~3) * 4;          // per group: input window + logit window
  constexpr int BAR_OFF = TILE_OFF + kGroups * 2 * kTileBytes;
  float (*xs)[kWin + 1] = reinterpret_cast<float (*)[kWin + 1]>(smem + TILE_OFF + g * 2 * kTileBytes);
  float (*ls)[kWin + 1] = reinterpret_cast<float (*)[kWin + 1]>(smem + TILE_OFF + g * 2 * kTileBytes + kTileBytes);
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
  const uint32_t tmem_grp = tmem_base + g * kGroupCols;
  const uint32_t tmem_d = tmem_grp + (static_cast<uint32_t>(wq * 32) << 16);
  constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, N);

  const int64_t num_tiles = (p.B + 127) >> 7;
  uint32_t mma_phase = 0;
  bool ok = true;
  for (int64_t tile = static_cast<int64_t>(blockIdx.x) * kGroups + g; ok && tile < num_tiles;
       tile += static_cast<int64_t>(gridDim.x) * kGroups) {
    const int64_t f = tile * 128 + r;
    const bool live = f < p.B;
    float c[H];
#pragma unroll
    for (int u = 0; u < H; ++u) c[u] = 0.0f;
    uint32_t bits = 0;
    const int T = p.T_enc + p.T_dec;
    // A operand of a step: (x, h0) (h1, h2) ... (h[H-1], 1.0) 0 ... as bf16 pairs, two per 32-bit column
    uint32_t w[H / 2 + 1];
    // inputs travel through a [128 x 32] shared window filled with coalesced row reads (one frame's 32 consecutive
    // time steps per warp request); per-thread strided reads made the run time depend on where the buffer landed
    auto load_window = [&](int t0) {
#pragma unroll 4
      for (int idx = r; idx < 128 * kWin; idx += 128) {
        const int rr = idx >> 5, sidx = idx & 31, tt = t0 + sidx;
        const int64_t ff = tile * 128 + rr;
        float v = 0.0f;
        if (ff < p.B) {
          if (tt < p.T_enc) v = load_as_f32(p.y, p.y_dtype, ff * p.T_enc + tt);
          else if (p.dec_in != nullptr && tt < T) v = __ldg(p.dec_in + ff * p.T_dec + (tt - p.T_enc));
        }
        xs[rr][sidx] = v;
      }
    };
    ptx::bar_sync(1 + g, 128);   // the previous tile's readers are done with the windows
    load_window(0);
    ptx::bar_sync(1 + g, 128);
    w[0] = pack_bf16x2(xs[r][0], 0.0f);
#pragma unroll
    for (int i = 1; i < H / 2; ++i) w[i] = 0u;
    w[H / 2] = pack_bf16x2(0.0f, 1.0f);
    for (int t = 0; t < T; ++t) {
      const bool dec = t >= p.T_enc;
      const int td = t - p.T_enc;
      // ---- hand the A operand over and start the step's two K = 16 MMAs -------------------------
      {
        uint32_t wa[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) wa[i] = (i <= H / 2) ? w[i] : 0u;
        ptx::tmem_st16(tmem_d + COL_A, wa);
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before();   // also orders the previous step's gate reads before this step's MMA
      ptx::bar_sync(1 + g, 128);
      if (issuer_warp && ptx::elect_one()) {
        ptx::tc_fence_after();
        const uint32_t wbase16 = ((sbase & 0x3FFFFu) >> 4) + (dec ? (kCellBytes >> 4) : 0);   // one uniform add per descriptor
#pragma unroll
        for (int j = 0; j < KP / 16; ++j)
          ptx::umma_f16_ts(tmem_grp + COL_D, tmem_grp + COL_A + 8 * j,
                           ptx::make_smem_desc16(wbase16, j * 2 * (N * 16), N * 16, 128), idesc, j > 0);
        ptx::umma_commit(bar_mma);
      }
      // next window of inputs while the MMA runs (every 32 steps; the hand-off barrier above separates it from
      // the last reads of the previous window)
      if (((t + 1) & (kWin - 1)) == 0 && t + 1 < T) { load_window(t + 1); ptx::bar_sync(1 + g, 128); }
      float x_next = xs[r][(t + 1) & (kWin - 1)];
      if (t + 1 == p.T_enc && p.dec_in == nullptr) x_next = 1.0f;   // EOS opens the autoregressive decoder (rnn_viterbi.py:93)
      if (!ptx::mbar_spin_bounded(bar_mma, mma_phase, kMaxWaitMs)) { atomicExch(p.abort_flag, 20); ok = false; break; }
      mma_phase ^= 1;
      ptx::tc_fence_after();
      // ---- gates: 16 columns = 4 units x (i/2, j, (f+1)/2, o/2); rolling double buffer of tcgen05.ld ----
      float lg = s_lin[H], h_carry = 0.0f, h0 = 0.0f;
      uint32_t v0[16], v1[16];
      ptx::tmem_ld16(tmem_d + COL_D, v0);
      ptx::tmem_ld16(tmem_d + COL_D + 16, v1);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int q = 0; q < H / 4; ++q) {
        float hq[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int u = 4 * q + i;
          const uint32_t* v = (q & 1) ? v1 : v0;
          float ti, tj, tf, to;
          if (i < NVD_LSTM_POLY_PAIRS) tanh_pair_poly(v[4 * i], v[4 * i + 1], ti, tj);
          else { ti = ptx::tanh_approx(__uint_as_float(v[4 * i])); tj = ptx::tanh_approx(__uint_as_float(v[4 * i + 1])); }
          if (i + 4 < NVD_LSTM_POLY_PAIRS) tanh_pair_poly(v[4 * i + 2], v[4 * i + 3], tf, to);
          else { tf = ptx::tanh_approx(__uint_as_float(v[4 * i + 2])); to = ptx::tanh_approx(__uint_as_float(v[4 * i + 3])); }
          const float si = fmaf(ti, 0.5f, 0.5f);
          const float sf = fmaf(tf, 0.5f, 0.5f);
          const float so = fmaf(to, 0.5f, 0.5f);
          c[u] = fmaf(sf, c[u], si * tj);
          hq[i] = so * ptx::tanh_approx(c[u]);
          lg = fmaf(hq[i], s_lin[u], lg);   // every step: no branch on the phase inside the unrolled gate code
        }
        if (q >= 1 && q + 1 < H / 4) ptx::tmem_wait_ld();          // chunk q+1 (issued one chunk ago) has landed
        if (q + 2 < H / 4) { if (q & 1) ptx::tmem_ld16(tmem_d + COL_D + 16 * (q + 2), v1); else ptx::tmem_ld16(tmem_d + COL_D + 16 * (q + 2), v0); }
        // pack: words (h[4q-1], h[4q]) (h[4q+1], h[4q+2]); h[4q+3] is carried into the next chunk
        if (q == 0) h0 = hq[0]; else w[2 * q] = pack_bf16x2(h_carry, hq[0]);
        w[2 * q + 1] = pack_bf16x2(hq[1], hq[2]);
        h_carry = hq[3];
      }
      w[H / 2] = pack_bf16x2(h_carry, 1.0f);
      float x = x_next;
      if (dec) {
        const bool one = lg > 0.0f;
        bits |= (one ? 1u : 0u) << (td & 31);
        if (p.dec_in == nullptr) x = one ? 1.0f : -1.0f;   // autoregressive: the decision is the next input
        if (p.logits != nullptr) {
          ls[r][td & (kWin - 1)] = lg;
          if ((td & (kWin - 1)) == kWin - 1 || td == p.T_dec - 1) {   // flush the window with coalesced row writes
            const int td0 = td & ~(kWin - 1);
            ptx::bar_sync(1 + g, 128);
            for (int idx = r; idx < 128 * kWin; idx += 128) {
              const int rr = idx >> 5, sidx = idx & 31;
              const int64_t ff = tile * 128 + rr;
              if (ff < p.B && td0 + sidx <= td) p.logits[ff * p.T_dec + td0 + sidx] = ls[rr][sidx];
            }
          }
        }
        if (p.hard != nullptr && live && ((td & 31) == 31 || td == p.T_dec - 1)) {
          const int byte0 = (td >> 5) << 2;
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (byte0 + b < p.hard_stride) p.hard[f * p.hard_stride + byte0 + b] = static_cast<uint8_t>(bits >> (8 * b));
          bits = 0;
        }
      }
      w[0] = pack_bf16x2(x, h0);
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
  const int smem = 2 * N * KP * 2 + ((H + 1 + 3) & ~3) * 4 + kGroups * 2 * kTileBytes + 128;
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
