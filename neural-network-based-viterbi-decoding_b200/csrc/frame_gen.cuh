// frame_gen.cuh — device functions shared by the staged kernels (K0/K1/K2) and the fused
// pipeline kernel, so that both produce bit-identical frames:
//   * Philox4x32-10 counter-based generator (checked word-for-word against oracle/philox.py)
//   * Box–Muller on MUFU (lg2 / sqrt / sin / cos approximations), six normals per Philox call
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

// One Philox call (128 bits) feeds SIX symbols (oracle/philox.py block_normals): three Box-Muller pairs, pair i
// taking a 27-bit radius uniform from the top of word i and a 15-bit angle made of the low 5 bits of word i and bits
// [10i, 10i+10) of word 3.  Four MUFU operations per pair: lg2, sqrt, cos, sin.
constexpr int kNormalsPerBlock = 6;

// radius / angle fields of pair i (0..2) of a block
__device__ __forceinline__ float bm_log2_u1(uint32_t xi) {          // lg2((r_bits + 0.5) * 2^-27)
  // the argument is >= 2^-28, never subnormal: the .ftz form is the same function here without the four instructions
  // of subnormal scaling that __log2f wraps around MUFU.LG2
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fmaf(__uint2float_rn(xi >> 5), 7.450580596923828e-9f, 3.725290298461914e-9f)));
  return l;
}
template <int I>
__device__ __forceinline__ float bm_theta(uint32_t xi, uint32_t x3) {   // int32(a_bits << 17) * pi * 2^-31
  const uint32_t lo = (I == 0) ? (x3 << 17) : (I == 1) ? (x3 << 7) : (x3 >> 3);
  uint32_t t;
  asm("lop3.b32 %0, %1, 0x07FE0000, %2, 0xEA;" : "=r"(t) : "r"(lo), "r"(xi << 27));   // (lo & mask) | (xi << 27)
  return __int2float_rn(static_cast<int>(t)) * 1.4629180792671596e-9f;                 // pi * 2^-31
}

// the first NP (1..3) pairs of one block -> z[0 .. 2*NP); written stage by stage (all lg2, all sqrt, all sin/cos) so
// that the MUFU operations of the pairs are in flight together instead of queueing one by one
template <int NP>
__device__ __forceinline__ void block_normals(const uint4 x, float* z) {
  const uint32_t w[3] = {x.x, x.y, x.z};
  float l[NP], th[NP], r[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) l[i] = bm_log2_u1(w[i]);
  th[0] = bm_theta<0>(w[0], x.w);
  if constexpr (NP > 1) th[1] = bm_theta<1>(w[1], x.w);
  if constexpr (NP > 2) th[2] = bm_theta<2>(w[2], x.w);
#pragma unroll
  for (int i = 0; i < NP; ++i) r[i] = sqrt_approx(-1.3862943611198906f * l[i]);   // sqrt(-2 ln u1)
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    z[2 * i] = r[i] * __cosf(th[i]);
    z[2 * i + 1] = r[i] * __sinf(th[i]);
  }
}

// 6 standard normals of symbols 6*block .. 6*block+5 of `frame`
__device__ __forceinline__ void noise6(const PhiloxKeys& rk, uint64_t frame, uint32_t block, uint32_t stream_word, float z[6]) {
  block_normals<3>(frame_words(rk, frame, block, stream_word), z);
}

// the 14 normals of an L = 7 frame (blocks 0, 1 and the first pair of block 2): the three Philox chains first, then
// the seven Box-Muller pairs stage by stage.  The three stages are separate functions so that a caller with idle time
// in three places (the fused decoder's epilogue threads, between a hand-off to the tensor core and its completion)
// can run one stage in each; noise14 is their composition, so both forms give the same bits.
struct Noise14Words {
  uint32_t w[7];   // radius words of the seven pairs (low 5 bits: angle)
  uint32_t a[3];   // word 3 of the three blocks (angle fields)
};
__device__ __forceinline__ void noise14_words(const PhiloxKeys& rk, uint64_t frame, uint32_t stream_word, Noise14Words& n) {
  const uint4 x0 = frame_words(rk, frame, 0u, stream_word);
  const uint4 x1 = frame_words(rk, frame, 1u, stream_word);
#ifdef NVD_EXP_TWO_CALLS   // timing experiment only (wrong noise): what a third Philox call costs
  const uint4 x2 = make_uint4(x0.x ^ x1.y, x0.y ^ x1.z, x0.z ^ x1.w, x0.w ^ x1.x);
#else
  const uint4 x2 = frame_words(rk, frame, 2u, stream_word);
#endif
  n.w[0] = x0.x; n.w[1] = x0.y; n.w[2] = x0.z; n.w[3] = x1.x; n.w[4] = x1.y; n.w[5] = x1.z; n.w[6] = x2.x;
  n.a[0] = x0.w; n.a[1] = x1.w; n.a[2] = x2.w;
}
__device__ __forceinline__ void noise14_polar(const Noise14Words& n, float (&r)[7], float (&th)[7]) {
  float l[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) l[i] = bm_log2_u1(n.w[i]);
  th[0] = bm_theta<0>(n.w[0], n.a[0]); th[1] = bm_theta<1>(n.w[1], n.a[0]); th[2] = bm_theta<2>(n.w[2], n.a[0]);
  th[3] = bm_theta<0>(n.w[3], n.a[1]); th[4] = bm_theta<1>(n.w[4], n.a[1]); th[5] = bm_theta<2>(n.w[5], n.a[1]);
  th[6] = bm_theta<0>(n.w[6], n.a[2]);
#pragma unroll
  for (int i = 0; i < 7; ++i) r[i] = sqrt_approx(-1.3862943611198906f * l[i]);
}
__device__ __forceinline__ void noise14_finish(const float (&r)[7], const float (&th)[7], float z[14]) {
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    z[2 * i] = r[i] * __cosf(th[i]);
    z[2 * i + 1] = r[i] * __sinf(th[i]);
  }
}
__device__ __forceinline__ void noise14(const PhiloxKeys& rk, uint64_t frame, uint32_t stream_word, float z[14]) {
  Noise14Words n;
  float r[7], th[7];
  noise14_words(rk, frame, stream_word, n);
  noise14_polar(n, r, th);
  noise14_finish(r, th, z);
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
