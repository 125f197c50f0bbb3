// K4: persistent LSTM seq2seq recurrence (encoder LSTM -> decoder LSTM -> linear), one frame per
// thread, weights resident in shared memory and state (c, h) in registers for all T_enc + T_dec
// time steps of the launch.
//
// Replaces the forward of build_seq2seq_model (/root/reference/src/rnn_viterbi.py:278-305):
// two tf.contrib.rnn.LSTMCell(H) unrolled by tf.nn.dynamic_rnn (time-major [T, B, 1]) and
// tf.contrib.layers.linear(., 1).  TF-1.x LSTMCell: z = [x, h] @ W + b, gates (i, j, f, o),
// c' = sigmoid(f + 1) c + sigmoid(i) tanh(j), h' = sigmoid(o) tanh(c').
//
// Weight image (host-prepared in nvd_lstm_load): per cell, Wp[k][u] = float4 of the four gate
// weights of input k (k = 0 is x, k = 1 + j is h_j) for unit u, so one 128-bit broadcast LDS feeds
// four FFMAs; the forget bias 1.0 is folded into b.  Inputs and logits move through 128 x 32
// shared tiles so global traffic stays coalesced although the recurrence is per-thread.
#include "common.cuh"

namespace nvd {

namespace {

constexpr int kFrames = 128;  // frames (threads) per CTA
constexpr int kTile = 32;     // time steps per staged tile

template <bool FAST>
__device__ __forceinline__ float act_tanh(float x) {
  if constexpr (FAST) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  } else {
    return tanhf(x);
  }
}
template <bool FAST>
__device__ __forceinline__ float act_sigmoid(float x) {
  if constexpr (FAST) return fmaf(act_tanh<true>(0.5f * x), 0.5f, 0.5f);
  else return 1.0f / (1.0f + expf(-x));
}

template <int H, bool FAST>
__device__ __forceinline__ void cell_step(const float4* __restrict__ Wp, const float4* __restrict__ b4,
                                          float x, float (&h)[H], float (&c)[H]) {
  float hn[H];
#pragma unroll
  for (int u = 0; u < H; ++u) {
    float4 a = b4[u];
    {
      const float4 w = Wp[u];
      a.x = fmaf(x, w.x, a.x); a.y = fmaf(x, w.y, a.y); a.z = fmaf(x, w.z, a.z); a.w = fmaf(x, w.w, a.w);
    }
#pragma unroll
    for (int k = 0; k < H; ++k) {
      const float4 w = Wp[(1 + k) * H + u];
      a.x = fmaf(h[k], w.x, a.x); a.y = fmaf(h[k], w.y, a.y); a.z = fmaf(h[k], w.z, a.z); a.w = fmaf(h[k], w.w, a.w);
    }
    const float ig = act_sigmoid<FAST>(a.x), jg = act_tanh<FAST>(a.y);
    const float fg = act_sigmoid<FAST>(a.z), og = act_sigmoid<FAST>(a.w);
    c[u] = fmaf(fg, c[u], ig * jg);
    hn[u] = og * act_tanh<FAST>(c[u]);
  }
#pragma unroll
  for (int u = 0; u < H; ++u) h[u] = hn[u];
}

__device__ __forceinline__ float load_as_f32(const void* p, int dtype, int64_t i) {
  if (dtype == NVD_F32) return static_cast<const float*>(p)[i];
  if (dtype == NVD_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
  return static_cast<float>(static_cast<const int32_t*>(p)[i]);
}

// cooperative [kFrames x kTile] tile load: rows are frames, columns time steps t0..t0+31
__device__ __forceinline__ void load_tile(float (*xs)[kTile + 1], const void* src, int dtype, int64_t f0,
                                          int64_t B, int T, int t0) {
#pragma unroll 1
  for (int idx = threadIdx.x; idx < kFrames * kTile; idx += kFrames) {
    const int r = idx >> 5, s = idx & 31;
    const int64_t f = f0 + r;
    const int t = t0 + s;
    xs[r][s] = (f < B && t < T) ? load_as_f32(src, dtype, f * T + t) : 0.0f;
  }
}

// hide a loop-invariant shared pointer from LICM: otherwise nvcc hoists all (1+H)*H weight loads
// of the recurrence out of the time loop and spills them
template <typename P>
__device__ __forceinline__ P* opaque(P* p) {
  asm volatile("" : "+l"(p));
  return p;
}

template <int H, bool FAST>
__global__ void __launch_bounds__(kFrames) lstm_seq2seq_kernel(const float* __restrict__ blob, const void* __restrict__ y, int y_dtype,
                                                               const float* __restrict__ dec_in, int T_enc, int T_dec,
                                                               int64_t B, float* __restrict__ logits,
                                                               uint8_t* __restrict__ hard, int hard_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kCellF4 = (1 + H) * H + H;  // Wp + b4, in float4
  float4* sW = reinterpret_cast<float4*>(smem_raw);
  float* sLin = reinterpret_cast<float*>(sW + 2 * kCellF4);
  float (*xs)[kTile + 1] = reinterpret_cast<float (*)[kTile + 1]>(sLin + ((H + 1 + 3) & ~3));
  float (*ls)[kTile + 1] = xs + kFrames;

  const int blob_floats = 2 * kCellF4 * 4 + H + 1;
  for (int i = threadIdx.x; i < blob_floats; i += kFrames) reinterpret_cast<float*>(smem_raw)[i] = blob[i];
  const float4* encW = sW;
  const float4* encB = sW + (1 + H) * H;
  const float4* decW = sW + kCellF4;
  const float4* decB = decW + (1 + H) * H;

  const int64_t f0 = static_cast<int64_t>(blockIdx.x) * kFrames;
  const int r = threadIdx.x;
  const int64_t f = f0 + r;
  float h[H], c[H];
#pragma unroll
  for (int u = 0; u < H; ++u) { h[u] = 0.0f; c[u] = 0.0f; }

  // ---- encoder: T_enc steps over the noisy symbols --------------------------------------
  for (int t0 = 0; t0 < T_enc; t0 += kTile) {
    __syncthreads();
    load_tile(xs, y, y_dtype, f0, B, T_enc, t0);
    __syncthreads();
    const int steps = min(kTile, T_enc - t0);
#pragma unroll 1
    for (int s = 0; s < steps; ++s) cell_step<H, FAST>(opaque(encW), opaque(encB), xs[r][s], h, c);
  }
  // ---- decoder: T_dec steps, teacher-forced or autoregressive ----------------------------
  float prev = 1.0f;  // EOS (rnn_viterbi.py:93)
  for (int t0 = 0; t0 < T_dec; t0 += kTile) {
    __syncthreads();
    if (dec_in) load_tile(xs, dec_in, NVD_F32, f0, B, T_dec, t0);
    __syncthreads();
    const int steps = min(kTile, T_dec - t0);
    uint32_t bits = 0;
#pragma unroll 1
    for (int s = 0; s < steps; ++s) {
      const float x = dec_in ? xs[r][s] : prev;
      cell_step<H, FAST>(opaque(decW), opaque(decB), x, h, c);
      float lg = sLin[H];
#pragma unroll
      for (int u = 0; u < H; ++u) lg = fmaf(h[u], sLin[u], lg);
      ls[r][s] = lg;
      const bool one = lg > 0.0f;
      bits |= (one ? 1u : 0u) << s;
      prev = one ? 1.0f : -1.0f;
    }
    if (hard && f < B) {
      for (int b = 0; b < 4; ++b) {
        const int byte = (t0 >> 3) + b;
        if (byte < hard_stride) hard[f * hard_stride + byte] = static_cast<uint8_t>(bits >> (8 * b));
      }
    }
    if (logits) {
      __syncthreads();
      for (int idx = threadIdx.x; idx < kFrames * kTile; idx += kFrames) {
        const int rr = idx >> 5, s = idx & 31;
        const int64_t ff = f0 + rr;
        if (ff < B && t0 + s < T_dec) logits[ff * T_dec + t0 + s] = ls[rr][s];
      }
    }
  }
}

template <int H>
size_t lstm_smem_bytes() {
  const size_t w = static_cast<size_t>(2) * ((1 + H) * H + H) * 16;
  const size_t lin = static_cast<size_t>((H + 1 + 3) & ~3) * 4;
  const size_t tiles = static_cast<size_t>(2) * kFrames * (kTile + 1) * 4;
  return w + lin + tiles;
}

template <int H, bool FAST>
int launch_one(nvd_ctx* ctx, const LstmModel& m, const void* y, int y_dtype, const float* dec_in, int T_enc,
               int T_dec, int64_t B, float* logits, uint8_t* hard) {
  auto kern = lstm_seq2seq_kernel<H, FAST>;
  const size_t smem = lstm_smem_bytes<H>();
  NVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const unsigned grid = static_cast<unsigned>((B + kFrames - 1) / kFrames);
  kern<<<grid, kFrames, smem, ctx->stream>>>(m.blob, y, y_dtype, dec_in, T_enc, T_dec, B, logits, hard,
                                             (T_dec + 7) / 8);
  NVD_CHECK_LAUNCH();
  return NVD_OK;
}

template <int H>
int launch_h(nvd_ctx* ctx, const LstmModel& m, const void* y, int y_dtype, const float* dec_in, int T_enc,
             int T_dec, int64_t B, int fast, float* logits, uint8_t* hard) {
  return fast ? launch_one<H, true>(ctx, m, y, y_dtype, dec_in, T_enc, T_dec, B, logits, hard)
              : launch_one<H, false>(ctx, m, y, y_dtype, dec_in, T_enc, T_dec, B, logits, hard);
}

}  // namespace

int launch_lstm(nvd_ctx* ctx, const LstmModel& m, const void* y, int y_dtype, const float* dec_in, int T_enc,
                int T_dec, int64_t B, int fast_math, float* logits, uint8_t* hard) {
  if (B == 0) return NVD_OK;
  if (y_dtype != NVD_F32 && y_dtype != NVD_BF16 && y_dtype != NVD_I32_TRUNC) {
    set_error("nvd_lstm_forward: unknown y_dtype %d", y_dtype);
    return NVD_ERR_INVALID;
  }
  switch (m.H) {
    case 16: return launch_h<16>(ctx, m, y, y_dtype, dec_in, T_enc, T_dec, B, fast_math, logits, hard);
    case 20: return launch_h<20>(ctx, m, y, y_dtype, dec_in, T_enc, T_dec, B, fast_math, logits, hard);
    case 32: return launch_h<32>(ctx, m, y, y_dtype, dec_in, T_enc, T_dec, B, fast_math, logits, hard);
  }
  set_error("nvd_lstm_forward: hidden size %d has no compiled kernel (16, 20, 32)", m.H);
  return NVD_ERR_UNSUPPORTED;
}

}  // namespace nvd
