// frame_gen.cuh — device functions shared by the staged kernels (K0/K1/K2) and the fused
// pipeline kernel, so that both produce bit-identical frames:
//   * Philox4x32-10 counter-based generator (checked word-for-word against oracle/philox.py)
//   * Box–Muller on MUFU (lg2 / sqrt / sin / cos approximations), eight normals per Philox call
//   * shift-XOR convolutional encoder on time-packed words
//       encode_133171  /root/reference/src/fnn_viterbi.py:157-217  taps {0,2,3,5,6} / {0,1,2,3,6}
//       encode_57      /root/reference/src/fnn_viterbi.py:122-154  taps {0,2} / {0,1,2}, g=5 first,
//                      plus the non-causal c[1] = b0^b1 of :143
//   * class label <-> time-packed word (generate_class_bits is MSB-first, fnn_viterbi.py:283)
#pragma once

#include <cstdint>

namespace nvd {

constexpr uint32_t kStreamNoise = 1u;
constexpr uint32_t kStreamInfo = 2u;
constexpr uint32_t kStreamNoiseTail = 3u;   // refinement bits of the rare far-tail radii (= kStreamNoise | 2)

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    // round keys as key + r * Weyl constant: with a warp-uniform key (kernel parameter) they are
    // computed once in the uniform datapath instead of 20 vector adds per call
    const uint32_t rk0 = k0 + static_cast<uint32_t>(r) * 0x9E3779B9u;
    const uint32_t rk1 = k1 + static_cast<uint32_t>(r) * 0xBB67AE85u;
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c.x;
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c.z;
    c = make_uint4(static_cast<uint32_t>(p1 >> 32) ^ c.y ^ rk0, static_cast<uint32_t>(p1),
                   static_cast<uint32_t>(p0 >> 32) ^ c.w ^ rk1, static_cast<uint32_t>(p0));
  }
  return c;
}

// The same generator with the 20 round keys (key + r * Weyl constant) precomputed on the host and
// passed as kernel parameters: they arrive as constant-bank operands of the XOR itself, so the
// per-round key additions disappear from the instruction stream altogether.
struct PhiloxKeys {
  uint32_t k[20];   // k[2r] = seed_lo + r * 0x9E3779B9, k[2r+1] = seed_hi + r * 0xBB67AE85
};
__host__ inline PhiloxKeys make_philox_keys(uint64_t seed) {
  PhiloxKeys rk;
  for (uint32_t r = 0; r < 10; ++r) {
    rk.k[2 * r] = static_cast<uint32_t>(seed) + r * 0x9E3779B9u;
    rk.k[2 * r + 1] = static_cast<uint32_t>(seed >> 32) + r * 0xBB67AE85u;
  }
  return rk;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys& rk) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c.x;
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c.z;
    c = make_uint4(static_cast<uint32_t>(p1 >> 32) ^ c.y ^ rk.k[2 * r], static_cast<uint32_t>(p1),
                   static_cast<uint32_t>(p0 >> 32) ^ c.w ^ rk.k[2 * r + 1], static_cast<uint32_t>(p0));
  }
  return c;
}

// counter convention (oracle/philox.py): (frame_lo, frame_hi, block, (snr_idx << 8) | stream)
__device__ __forceinline__ uint4 frame_words(uint64_t seed, uint64_t frame, uint32_t block,
                                             uint32_t stream_word) {
  return philox4x32_10(make_uint4(static_cast<uint32_t>(frame), static_cast<uint32_t>(frame >> 32),
                                  block, stream_word),
                       static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
}

__device__ __forceinline__ uint4 frame_words(const PhiloxKeys& rk, uint64_t frame, uint32_t block, uint32_t stream_word) {
  return philox4x32_10(make_uint4(static_cast<uint32_t>(frame), static_cast<uint32_t>(frame >> 32), block, stream_word), rk);
}

__host__ __device__ __forceinline__ uint32_t noise_stream_word(int snr_idx) {
  return (static_cast<uint32_t>(snr_idx) << 8) | kStreamNoise;
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One Philox call (128 bits) feeds EIGHT symbols 8*block .. 8*block+7 (oracle/philox.py block_normals): four
// Box-Muller pairs, pair p made of word p alone — a 20-bit radius field k = w >> 12 and a 12-bit angle field
// a = w & 0xFFF:
//     u = (k + 0.5) * 2^-20,   theta = a * 2 pi / 4096 - pi,   r = sqrt(-2 ln u),   z = (r cos theta, r sin theta).
// Twenty bits alone would end the radius at 5.4 sigma on a coarse grid, so the sixteen lowest cells (k < 16, i.e.
// w < 2^16, probability 1.5e-5 per pair) are refined by 32 more bits, word p of a SECOND Philox call on the stream
// kStreamNoiseTail made only when a thread meets such a word:  u = (k * 2^32 + t + 0.5) * 2^-52  (radius up to
// 8.6 sigma, relative cell width <= 6 % everywhere).  The fast path costs four MUFU operations and eleven other
// instructions per pair; Philox is the dominant cost of every Monte-Carlo kernel, and the same 128 bits fed 4
// symbols in round 1 and 6 before this form.
// The symbol is produced as  y = fma(sigma r, cos theta, +-1): sigma is folded into the radius (c_scale below), so
// there is no separate multiply of r with the trigonometric factor.
constexpr int kNormalsPerBlock = 8;
constexpr uint32_t kTailWordLimit = 1u << 16;          // w < 2^16  <=>  radius field k < 16

__host__ __device__ __forceinline__ float bm_scale(float sigma) { return -1.3862943611198906f * sigma * sigma; }   // -2 ln 2 sigma^2
__device__ __forceinline__ uint32_t tail_stream_word(uint32_t noise_stream_word) { return noise_stream_word | 2u; }

__device__ __forceinline__ float lg2_approx(float x) {
  // the arguments here are never subnormal (>= 2^-53): the .ftz form is the same function without the four
  // instructions of subnormal scaling that __log2f wraps around MUFU.LG2
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
  return l;
}
__device__ __forceinline__ float bm_lg2_u(uint32_t w) {                   // lg2((k + 0.5) * 2^-20)
  return lg2_approx(fmaf(__uint2float_rn(w >> 12), 9.5367431640625e-7f, 4.76837158203125e-7f));
}
__device__ __forceinline__ float bm_lg2_u_tail(uint32_t w, uint32_t t) {  // lg2((k * 2^32 + t + 0.5) * 2^-52)
  return lg2_approx(fmaf(__uint2float_rn(w >> 12), 9.5367431640625e-7f,
                         fmaf(__uint2float_rn(t), 2.220446049250313e-16f, 1.1102230246251565e-16f)));
}
__device__ __forceinline__ float bm_theta(uint32_t w) {                   // (a / 4096) * 2 pi - pi
  uint32_t fb;   // float in [1, 2): mantissa = a << 11
  asm("lop3.b32 %0, %1, 0x007FF800, 0x3F800000, 0xEA;" : "=r"(fb) : "r"(w << 11));   // (x & mask) | exponent
  return fmaf(__uint_as_float(fb), 6.283185307179586f, -9.42477796076938f);
}
// sigma * r of pair words w[0..NP) of one block, the far-tail refinement included (rare: one more Philox call)
template <int NP>
__device__ __forceinline__ void block_radii(const PhiloxKeys& rk, uint64_t frame, uint32_t block, uint32_t stream_word,
                                            const uint32_t (&w)[NP], float c_scale, float (&r)[NP]) {
  float l[NP];
  uint32_t mn = w[0];
#pragma unroll
  for (int i = 0; i < NP; ++i) { l[i] = bm_lg2_u(w[i]); mn = min(mn, w[i]); }
  if (mn < kTailWordLimit) {
    const uint4 t = frame_words(rk, frame, block, tail_stream_word(stream_word));
    const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < NP; ++i)
      if (w[i] < kTailWordLimit) l[i] = bm_lg2_u_tail(w[i], tw[i]);
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) r[i] = sqrt_approx(c_scale * l[i]);
}

// symbols 8*block .. 8*block + 2*NP - 1 of `frame`:  y[j] = fma(sigma r, cos/sin theta, bit j of `bits` ? +1 : -1)
template <int NP>
__device__ __forceinline__ void block_symbols(const PhiloxKeys& rk, uint64_t frame, uint32_t block, uint32_t stream_word,
                                              float c_scale, uint32_t bits, float* y) {
  const uint4 x = frame_words(rk, frame, block, stream_word);
  const uint32_t xw[4] = {x.x, x.y, x.z, x.w};
  uint32_t w[NP];
  float r[NP], th[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) { w[i] = xw[i]; th[i] = bm_theta(xw[i]); }
  block_radii<NP>(rk, frame, block, stream_word, w, c_scale, r);
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    y[2 * i] = fmaf(r[i], __cosf(th[i]), ((bits >> (2 * i)) & 1u) ? 1.0f : -1.0f);
    y[2 * i + 1] = fmaf(r[i], __sinf(th[i]), ((bits >> (2 * i + 1)) & 1u) ? 1.0f : -1.0f);
  }
}

// The 14 symbols of an L = 7 frame (block 0 and the first three pairs of block 1) in three stages, so that a caller
// with idle time in three places (the fused decoder's epilogue threads, between a hand-off to the tensor core and its
// completion) can run one stage in each; frame14_symbols is their composition, so both forms give the same bits.
struct Noise14Words {
  uint32_t w[7];
};
__device__ __forceinline__ void noise14_words(const PhiloxKeys& rk, uint64_t frame, uint32_t stream_word, Noise14Words& n) {
  const uint4 x0 = frame_words(rk, frame, 0u, stream_word);
  const uint4 x1 = frame_words(rk, frame, 1u, stream_word);
  n.w[0] = x0.x; n.w[1] = x0.y; n.w[2] = x0.z; n.w[3] = x0.w; n.w[4] = x1.x; n.w[5] = x1.y; n.w[6] = x1.z;
}
__device__ __forceinline__ void noise14_polar(const PhiloxKeys& rk, uint64_t frame, uint32_t stream_word, const Noise14Words& n,
                                              float c_scale, float (&r)[7], float (&th)[7]) {
  float l[7];
  uint32_t mn = n.w[0];
#pragma unroll
  for (int i = 0; i < 7; ++i) { l[i] = bm_lg2_u(n.w[i]); th[i] = bm_theta(n.w[i]); mn = min(mn, n.w[i]); }
  if (mn < kTailWordLimit) {
    const uint4 t0 = frame_words(rk, frame, 0u, tail_stream_word(stream_word));
    const uint4 t1 = frame_words(rk, frame, 1u, tail_stream_word(stream_word));
    const uint32_t tw[7] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z};
#pragma unroll
    for (int i = 0; i < 7; ++i)
      if (n.w[i] < kTailWordLimit) l[i] = bm_lg2_u_tail(n.w[i], tw[i]);
  }
#pragma unroll
  for (int i = 0; i < 7; ++i) r[i] = sqrt_approx(c_scale * l[i]);
}
__device__ __forceinline__ void noise14_symbols(const float (&r)[7], const float (&th)[7], uint32_t bits, float y[14]) {
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    y[2 * i] = fmaf(r[i], __cosf(th[i]), ((bits >> (2 * i)) & 1u) ? 1.0f : -1.0f);
    y[2 * i + 1] = fmaf(r[i], __sinf(th[i]), ((bits >> (2 * i + 1)) & 1u) ? 1.0f : -1.0f);
  }
}
__device__ __forceinline__ void frame14_symbols(const PhiloxKeys& rk, uint64_t frame, uint32_t stream_word, float c_scale,
                                                uint32_t bits, float y[14]) {
  Noise14Words n;
  float r[7], th[7];
  noise14_words(rk, frame, stream_word, n);
  noise14_polar(rk, frame, stream_word, n, c_scale, r, th);
  noise14_symbols(r, th, bits, y);
}

// ---- encoder on packed words ---------------------------------------------------------------
// v = (u << 6) | hist6, hist6 bit 5 = info bit at time -1 ... bit 0 = time -6.
// Returns generator outputs for the (up to 32) time steps held in u.
__device__ __forceinline__ void conv_outputs(uint64_t v, int code, uint32_t& oa, uint32_t& ob) {
  if (code == NVD_CODE_133171) {
    oa = static_cast<uint32_t>((v >> 6) ^ (v >> 4) ^ (v >> 3) ^ (v >> 1) ^ v);
    ob = static_cast<uint32_t>((v >> 6) ^ (v >> 5) ^ (v >> 4) ^ (v >> 3) ^ v);
  } else {
    oa = static_cast<uint32_t>((v >> 6) ^ (v >> 4));
    ob = static_cast<uint32_t>((v >> 6) ^ (v >> 5) ^ (v >> 4));
  }
}

// spread the low 16 bits of x to the even bit positions
__device__ __forceinline__ uint32_t spread16(uint32_t x) {
  x &= 0xFFFFu;
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}

// coded bits (interleaved [g_a(i), g_b(i)]) of `nvalid` (<= 16) info bits `u` (already masked
// to its valid bits) with history hist6; frame_start applies the (7,5) reference quirk.  Bits
// at and above 2*nvalid are returned zero.
__device__ __forceinline__ uint32_t encode16(uint32_t u, uint32_t hist6, int code, bool frame_start,
                                             int nvalid) {
  uint32_t oa, ob;
  conv_outputs((static_cast<uint64_t>(u) << 6) | hist6, code, oa, ob);
  if (code == NVD_CODE_57_REF && frame_start) ob ^= (u >> 1) & 1u;
  const uint32_t w = spread16(oa) | (spread16(ob) << 1);
  return nvalid >= 16 ? w : (w & ((1u << (2 * nvalid)) - 1u));
}

// MSB-first class index <-> time-packed word (bit reversal over L bits, L <= 32)
__device__ __forceinline__ uint32_t bitrev_L(uint32_t x, int L) { return __brev(x) >> (32 - L); }

}  // namespace nvd
