// K3 (throughput path) and the fused Monte-Carlo pipeline: the reference FNN decoder
//   out = tanh(tanh(tanh(tanh(y W1 + b1) W2 + b2) W3 + b3) Wout + bout),  argmax
// (/root/reference/src/fnn_viterbi.py:495-508, 606-610; 14 -> 128 -> 64 -> 32 -> 128) as ONE
// persistent tcgen05 kernel per launch.
//
// Mapping to the SM:
//   * 128 frames = one M=128 UMMA tile; layer widths are the UMMA N (128/64/32/128); every layer
//     is a chain of K=16 `tcgen05.mma.kind::f16` (bf16 x bf16 -> fp32) with the accumulator in TMEM.
//   * Weights live in shared memory for the kernel's lifetime, bf16, in the UMMA canonical
//     K-major / no-swizzle layout.  Biases ride in the contraction: layer 1 uses the spare input
//     column 14 (A holds 1.0 there), layers 2-4 add one extra K=16 step against a constant
//     "ones" A-chunk, so the epilogue is tanh + pack only.
//   * Four independent 128-thread groups per CTA, each owning 128 TMEM columns and its own
//     activation buffers, run tile streams that interleave on the SM: while one group waits for
//     its MMA the others keep the MUFU (tanh) and FMA pipes busy.  Thread r of a group owns frame
//     r of the tile == TMEM lane r (tcgen05.ld 32x32b), so argmax needs no shuffles.
//   * MODE_STAGED: y (bf16 [B,14]) arrives from HBM by TMA bulk copies (cp.async.bulk +
//     mbarrier complete_tx), prefetched one tile ahead; decisions / soft outputs go back to HBM.
//     MODE_FUSED: info word, encoder, Philox noise and BPSK are generated in registers
//     (frame_gen.cuh, bit-identical to K0/K1/K2), decisions are compared with the label and only
//     four 64-bit counters leave the SM.
// All mbarrier waits are bounded: a wedged pipeline sets ctx->abort_flag and returns.
#include <algorithm>

#include "common.cuh"
#include "frame_gen.cuh"
#include "ptx.cuh"

namespace nvd {

namespace {

constexpr int kGroups = 4;
constexpr int kThreads = 128 * kGroups;
constexpr int NIN = 14, N1 = 128, N2 = 64, N3 = 32, N4 = 128;
constexpr int kRowBytes = NIN * 2;  // bf16 input row

// bf16 weight image: B operands [N x Kpad] in canonical K-major layout, bias chunk appended
constexpr int W1_OFF = 0;                             // [128 x 16]   (k=14 carries b1)
constexpr int W2_OFF = W1_OFF + N1 * 16 * 2;          // [ 64 x 144]  (k=128 carries b2)
constexpr int W3_OFF = W2_OFF + N2 * (N1 + 16) * 2;   // [ 32 x  80]
constexpr int W4_OFF = W3_OFF + N3 * (N2 + 16) * 2;   // [128 x  48]
constexpr int W_BYTES = W4_OFF + N4 * (N3 + 16) * 2;  // 39936
static_assert(W_BYTES == 39936, "weight image size");

// shared memory carve-up
constexpr int ONES_OFF = W_BYTES;                     // [128 x 16] A chunk, column 0 = 1.0
constexpr int GROUP_OFF = ONES_OFF + 4096;
constexpr int A1_OFF = 0;                             // [128 x 16]
constexpr int A2_OFF = 4096;                          // [128 x 128]; A3 / A4 alias into it
constexpr int A3_OFF = A2_OFF;
constexpr int A4_OFF = A2_OFF + 16384;
constexpr int YBUF_OFF = A2_OFF + 32768;              // 128 x 28 B TMA landing buffer
constexpr int GROUP_BYTES = YBUF_OFF + 128 * kRowBytes;
constexpr int BAR_OFF = GROUP_OFF + kGroups * GROUP_BYTES;
constexpr int SMEM_BYTES = BAR_OFF + 128;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(GROUP_BYTES % 128 == 0 && GROUP_OFF % 128 == 0, "alignment");

constexpr int A_LBO = 128 * 16;  // bytes between the two K chunks of an A operand (128 rows)
constexpr int kSBO = 128;        // bytes between 8-row groups
constexpr uint32_t kMaxTries = 1u << 20;

enum { MODE_STAGED = 0, MODE_FUSED = 1 };

struct TcParams {
  const uint8_t* image;
  int* abort_flag;
  int64_t B;  // frames (rows) in this launch
  // staged
  const __nv_bfloat16* y;
  float* soft;
  void* hard;
  int hard_bytes;
  int last_act;
  int use_tma;
  // fused
  uint64_t seed;
  uint64_t frame0;
  uint64_t* counters;
  float sigma;
  uint32_t stream_word;
  int code;
  int L;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// hidden-layer epilogue: TMEM accumulator (bias already inside) -> tanh -> bf16 -> next A operand
template <int N>
__device__ __forceinline__ void epilogue_hidden(uint32_t tmem_d, uint32_t a_next, int r) {
#pragma unroll
  for (int c = 0; c < N; c += 32) {
    uint32_t v[32];
    ptx::tmem_ld32(tmem_d + c, v);
    ptx::tmem_wait_ld();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        w[e] = pack_bf16x2(ptx::tanh_approx(__uint_as_float(v[8 * q + 2 * e])),
                           ptx::tanh_approx(__uint_as_float(v[8 * q + 2 * e + 1])));
      st_shared_v4(a_next + ((c >> 3) + q) * A_LBO + r * 16, w[0], w[1], w[2], w[3]);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) mlp_tc_kernel(const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  const int tid = threadIdx.x;
  const int g = tid >> 7;
  const int r = tid & 127;
  const uint32_t grp = sbase + GROUP_OFF + g * GROUP_BYTES;
  const uint32_t bar_mma = sbase + BAR_OFF + 8 * g;
  const uint32_t bar_tma = sbase + BAR_OFF + 32 + 8 * g;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 64);

  // ---- one-time set-up: weights -> smem, ones chunk, barriers, TMEM --------------------------
  for (int i = tid; i < W_BYTES / 16; i += kThreads)
    reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(p.image) + i);
  for (int i = tid; i < 4096 / 16; i += kThreads) {
    // chunk 0 (columns 0..7) rows 0..127: column 0 = bf16(1.0) = 0x3F80; chunk 1 all zero
    reinterpret_cast<uint4*>(smem + ONES_OFF)[i] = (i < 128) ? make_uint4(0x00003F80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid == 0) {
    for (int i = 0; i < kGroups; ++i) {
      ptx::mbar_init(sbase + BAR_OFF + 8 * i, 1);
      ptx::mbar_init(sbase + BAR_OFF + 32 + 8 * i, 1);
    }
    ptx::mbar_fence_init();
  }
  if (tid < 32) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_grp = tmem_base + g * 128;                         // MMA destination (lane 0)
  const uint32_t tmem_d = tmem_grp + (static_cast<uint32_t>(r & ~31) << 16);  // this warp's lanes

  constexpr uint32_t idesc128 = ptx::make_idesc_bf16_f32(128, 128);
  constexpr uint32_t idesc64 = ptx::make_idesc_bf16_f32(128, 64);
  constexpr uint32_t idesc32 = ptx::make_idesc_bf16_f32(128, 32);

  const int64_t num_tiles = (p.B + 127) >> 7;
  const int64_t tile_stride = static_cast<int64_t>(gridDim.x) * kGroups;
  int64_t tile = static_cast<int64_t>(blockIdx.x) * kGroups + g;
  uint32_t mma_phase = 0, tma_phase = 0;
  uint32_t bit_err = 0, frame_err = 0;
  const uint32_t lmask = (1u << p.L) - 1u;

  auto tile_uses_tma = [&](int64_t t) { return p.use_tma && (t + 1) * 128 <= p.B; };
  if (MODE == MODE_STAGED && r == 0 && tile < num_tiles && tile_uses_tma(tile)) {
    ptx::mbar_arrive_expect_tx(bar_tma, 128 * kRowBytes);
    ptx::bulk_g2s(grp + YBUF_OFF, reinterpret_cast<const uint8_t*>(p.y) + tile * 128 * kRowBytes, 128 * kRowBytes, bar_tma);
  }

  for (; tile < num_tiles; tile += tile_stride) {
    const int64_t row = tile * 128 + r;
    uint32_t cls = 0;
    // ---- stage 0: this frame's 14 noisy symbols -> A1 (K padded to 16, column 14 = 1.0) -------
    uint32_t w[8];
    if (MODE == MODE_STAGED) {
      if (tile_uses_tma(tile)) {
        if (!ptx::mbar_wait_bounded(bar_tma, tma_phase, kMaxTries)) { atomicExch(p.abort_flag, 1); break; }
        tma_phase ^= 1;
#pragma unroll
        for (int i = 0; i < 7; ++i) w[i] = ld_shared_u32(grp + YBUF_OFF + r * kRowBytes + 4 * i);
      } else {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.y) + row * 7;
#pragma unroll
        for (int i = 0; i < 7; ++i) w[i] = (row < p.B) ? __ldg(src + i) : 0u;
      }
    } else {
      const uint64_t f = p.frame0 + static_cast<uint64_t>(row);
      cls = static_cast<uint32_t>(f) & lmask;
      const uint32_t coded = encode16(bitrev_L(cls, p.L), 0u, p.code, true, p.L);
      float yv[16];
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        float z[4];
        noise4(p.seed, f, blk, p.stream_word, z);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = 4 * blk + k;
          yv[j] = fmaf(p.sigma, z[k], ((coded >> j) & 1u) ? 1.0f : -1.0f);
        }
      }
#pragma unroll
      for (int i = 0; i < 7; ++i) w[i] = pack_bf16x2(yv[2 * i], yv[2 * i + 1]);
    }
    w[7] = 0x00003F80u;  // (1.0, 0.0): bias column 14, zero column 15
    st_shared_v4(grp + A1_OFF + r * 16, w[0], w[1], w[2], w[3]);
    st_shared_v4(grp + A1_OFF + A_LBO + r * 16, w[4], w[5], w[6], w[7]);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    ptx::bar_sync(1 + g, 128);

    // ---- layer 1: [128 x 16] x [16 x 128] --------------------------------------------------
    if (r == 0) {
      if (MODE == MODE_STAGED) {
        const int64_t nt = tile + tile_stride;
        if (nt < num_tiles && tile_uses_tma(nt)) {
          ptx::mbar_arrive_expect_tx(bar_tma, 128 * kRowBytes);
          ptx::bulk_g2s(grp + YBUF_OFF, reinterpret_cast<const uint8_t*>(p.y) + nt * 128 * kRowBytes, 128 * kRowBytes, bar_tma);
        }
      }
      ptx::tc_fence_after();
      ptx::umma_f16_ss(tmem_grp, ptx::make_smem_desc(grp + A1_OFF, A_LBO, kSBO),
                       ptx::make_smem_desc(sbase + W1_OFF, N1 * 16, kSBO), idesc128, 0u);
      ptx::umma_commit(bar_mma);
    }
    if (!ptx::mbar_wait_bounded(bar_mma, mma_phase, kMaxTries)) { atomicExch(p.abort_flag, 2); break; }
    mma_phase ^= 1;
    ptx::tc_fence_after();
    epilogue_hidden<N1>(tmem_d, grp + A2_OFF, r);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    ptx::bar_sync(1 + g, 128);

    // ---- layer 2: [128 x 128(+16)] x [.. x 64] ---------------------------------------------
    if (r == 0) {
      ptx::tc_fence_after();
#pragma unroll
      for (int j = 0; j < N1 / 16; ++j)
        ptx::umma_f16_ss(tmem_grp, ptx::make_smem_desc(grp + A2_OFF + j * 2 * A_LBO, A_LBO, kSBO),
                         ptx::make_smem_desc(sbase + W2_OFF + j * 2 * (N2 * 16), N2 * 16, kSBO), idesc64, j > 0);
      ptx::umma_f16_ss(tmem_grp, ptx::make_smem_desc(sbase + ONES_OFF, A_LBO, kSBO),
                       ptx::make_smem_desc(sbase + W2_OFF + (N1 / 16) * 2 * (N2 * 16), N2 * 16, kSBO), idesc64, 1u);
      ptx::umma_commit(bar_mma);
    }
    if (!ptx::mbar_wait_bounded(bar_mma, mma_phase, kMaxTries)) { atomicExch(p.abort_flag, 3); break; }
    mma_phase ^= 1;
    ptx::tc_fence_after();
    epilogue_hidden<N2>(tmem_d, grp + A3_OFF, r);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    ptx::bar_sync(1 + g, 128);

    // ---- layer 3: [128 x 64(+16)] x [.. x 32] ----------------------------------------------
    if (r == 0) {
      ptx::tc_fence_after();
#pragma unroll
      for (int j = 0; j < N2 / 16; ++j)
        ptx::umma_f16_ss(tmem_grp, ptx::make_smem_desc(grp + A3_OFF + j * 2 * A_LBO, A_LBO, kSBO),
                         ptx::make_smem_desc(sbase + W3_OFF + j * 2 * (N3 * 16), N3 * 16, kSBO), idesc32, j > 0);
      ptx::umma_f16_ss(tmem_grp, ptx::make_smem_desc(sbase + ONES_OFF, A_LBO, kSBO),
                       ptx::make_smem_desc(sbase + W3_OFF + (N2 / 16) * 2 * (N3 * 16), N3 * 16, kSBO), idesc32, 1u);
      ptx::umma_commit(bar_mma);
    }
    if (!ptx::mbar_wait_bounded(bar_mma, mma_phase, kMaxTries)) { atomicExch(p.abort_flag, 4); break; }
    mma_phase ^= 1;
    ptx::tc_fence_after();
    epilogue_hidden<N3>(tmem_d, grp + A4_OFF, r);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    ptx::bar_sync(1 + g, 128);

    // ---- layer 4: [128 x 32(+16)] x [.. x 128] ---------------------------------------------
    if (r == 0) {
      ptx::tc_fence_after();
#pragma unroll
      for (int j = 0; j < N3 / 16; ++j)
        ptx::umma_f16_ss(tmem_grp, ptx::make_smem_desc(grp + A4_OFF + j * 2 * A_LBO, A_LBO, kSBO),
                         ptx::make_smem_desc(sbase + W4_OFF + j * 2 * (N4 * 16), N4 * 16, kSBO), idesc128, j > 0);
      ptx::umma_f16_ss(tmem_grp, ptx::make_smem_desc(sbase + ONES_OFF, A_LBO, kSBO),
                       ptx::make_smem_desc(sbase + W4_OFF + (N3 / 16) * 2 * (N4 * 16), N4 * 16, kSBO), idesc128, 1u);
      ptx::umma_commit(bar_mma);
    }
    if (!ptx::mbar_wait_bounded(bar_mma, mma_phase, kMaxTries)) { atomicExch(p.abort_flag, 5); break; }
    mma_phase ^= 1;
    ptx::tc_fence_after();

    // ---- output epilogue: argmax (first index on ties); tanh is monotone so it is elided for
    //      the decision and applied only when soft outputs are requested -----------------------
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int c = 0; c < N4; c += 32) {
      uint32_t v[32];
      ptx::tmem_ld32(tmem_d + c, v);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = __uint_as_float(v[j]);
        if (x > best) { best = x; bi = c + j; }
      }
      if (MODE == MODE_STAGED && p.soft != nullptr && row < p.B) {
        float4* dst = reinterpret_cast<float4*>(p.soft + row * N4 + c);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o;
          o.x = __uint_as_float(v[j]); o.y = __uint_as_float(v[j + 1]);
          o.z = __uint_as_float(v[j + 2]); o.w = __uint_as_float(v[j + 3]);
          if (p.last_act == NVD_ACT_TANH) { o.x = tanhf(o.x); o.y = tanhf(o.y); o.z = tanhf(o.z); o.w = tanhf(o.w); }
          dst[j >> 2] = o;
        }
      }
    }
    ptx::tc_fence_before();  // TMEM reads done before the next tile's MMA (issued after a bar_sync)
    if (MODE == MODE_STAGED) {
      if (p.hard != nullptr && row < p.B) {
        if (p.hard_bytes == 1) static_cast<uint8_t*>(p.hard)[row] = static_cast<uint8_t>(bi);
        else if (p.hard_bytes == 2) static_cast<uint16_t*>(p.hard)[row] = static_cast<uint16_t>(bi);
        else static_cast<int32_t*>(p.hard)[row] = bi;
      }
    } else if (row < p.B) {
      const uint32_t e = (static_cast<uint32_t>(bi) ^ cls) & lmask;
      bit_err += __popc(e);
      frame_err += e ? 1u : 0u;
    }
  }

  if (MODE == MODE_FUSED) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bit_err += __shfl_xor_sync(0xFFFFFFFFu, bit_err, o);
      frame_err += __shfl_xor_sync(0xFFFFFFFFu, frame_err, o);
    }
    if ((tid & 31) == 0) {
      if (bit_err) atomicAdd(reinterpret_cast<unsigned long long*>(p.counters), static_cast<unsigned long long>(bit_err));
      if (frame_err) atomicAdd(reinterpret_cast<unsigned long long*>(p.counters + 1), static_cast<unsigned long long>(frame_err));
    }
    if (blockIdx.x == 0 && tid == 0) {
      atomicAdd(reinterpret_cast<unsigned long long*>(p.counters + 2), static_cast<unsigned long long>(p.B));
      atomicAdd(reinterpret_cast<unsigned long long*>(p.counters + 3), static_cast<unsigned long long>(p.B) * p.L);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem_base, 512);
}

// ---- diagnostics: one MMA on caller-supplied images -------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(int N, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo,
                                                            uint32_t b_sbo, const uint4* __restrict__ a_img,
                                                            const uint4* __restrict__ b_img, float* __restrict__ d,
                                                            int* abort_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  const int tid = threadIdx.x;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 4096 + 8192 + 8);
  const uint32_t bar = sbase + 4096 + 8192;
  for (int i = tid; i < 256; i += 128) reinterpret_cast<uint4*>(smem)[i] = a_img[i];
  for (int i = tid; i < N * 2; i += 128) reinterpret_cast<uint4*>(smem + 4096)[i] = b_img[i];
  if (tid == 0) { ptx::mbar_init(bar, 1); ptx::mbar_fence_init(); }
  if (tid < 32) { ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 128); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    ptx::umma_f16_ss(tmem_base, ptx::make_smem_desc(sbase, a_lbo, a_sbo), ptx::make_smem_desc(sbase + 4096, b_lbo, b_sbo),
                     ptx::make_idesc_bf16_f32(128, N), 0u);
    ptx::umma_commit(bar);
  }
  if (ptx::mbar_wait_bounded(bar, 0, kMaxTries)) {
    ptx::tc_fence_after();
    for (int c = 0; c < N; c += 32) {
      uint32_t v[32];
      ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(tid & ~31) << 16) + c, v);
      ptx::tmem_wait_ld();
      for (int j = 0; j < 32; ++j) d[tid * N + c + j] = __uint_as_float(v[j]);
    }
  } else {
    atomicExch(abort_flag, 9);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem_base, 128);
}

uint16_t f32_to_bf16_host(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

// element (n, k) of a [N x Kpad] K-major B operand in the canonical no-swizzle layout
inline size_t b_off(int N, int n, int k) { return static_cast<size_t>(k >> 3) * (N * 16) + n * 16 + (k & 7) * 2; }

void fill_layer(uint8_t* img, int off, int K, int N, const float* W, const float* b, bool bias_in_pad_col) {
  // W is [K, N] row-major (x @ W); the B operand is W^T.  Bias: at k = K (first column of the pad /
  // of the extra chunk).
  for (int n = 0; n < N; ++n) {
    for (int k = 0; k < K; ++k) {
      const uint16_t v = f32_to_bf16_host(W[static_cast<size_t>(k) * N + n]);
      memcpy(img + off + b_off(N, n, k), &v, 2);
    }
    const uint16_t bv = f32_to_bf16_host(b[n]);
    memcpy(img + off + b_off(N, n, K), &bv, 2);
  }
  (void)bias_in_pad_col;
}

template <int MODE>
int launch_tc(nvd_ctx* ctx, const TcParams& p) {
  auto kern = mlp_tc_kernel<MODE>;
  NVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const int64_t tiles = (p.B + 127) / 128;
  const int grid = static_cast<int>(std::min<int64_t>(ctx->sm_count, (tiles + kGroups - 1) / kGroups));
  kern<<<grid, kThreads, SMEM_BYTES, ctx->stream>>>(p);
  NVD_CHECK_LAUNCH();
  return NVD_OK;
}

}  // namespace

bool tc_shape_supported(const MlpModel& m) {
  if (m.n_layers != 4) return false;
  const int want[5] = {NIN, N1, N2, N3, N4};
  for (int i = 0; i < 5; ++i)
    if (m.dims[i] != want[i]) return false;
  for (int l = 0; l < 3; ++l)
    if (m.acts[l] != NVD_ACT_TANH) return false;
  return m.acts[3] == NVD_ACT_TANH || m.acts[3] == NVD_ACT_IDENTITY;
}

int tc_build_image(nvd_ctx* ctx, MlpModel& m, const float* const* W_host, const float* const* b_host) {
  std::vector<uint8_t> img(W_BYTES, 0);
  fill_layer(img.data(), W1_OFF, NIN, N1, W_host[0], b_host[0], true);
  fill_layer(img.data(), W2_OFF, N1, N2, W_host[1], b_host[1], false);
  fill_layer(img.data(), W3_OFF, N2, N3, W_host[2], b_host[2], false);
  fill_layer(img.data(), W4_OFF, N3, N4, W_host[3], b_host[3], false);
  NVD_CUDA(cudaMalloc(&m.tc_image, W_BYTES));
  NVD_CUDA(cudaMemcpy(m.tc_image, img.data(), W_BYTES, cudaMemcpyHostToDevice));
  m.tc_image_bytes = W_BYTES;
  (void)ctx;
  return NVD_OK;
}

int launch_mlp_tc_staged(nvd_ctx* ctx, const MlpModel& m, const __nv_bfloat16* y, int64_t B, float* soft_out,
                         void* hard_out, int hard_bytes) {
  if (B == 0) return NVD_OK;
  TcParams p{};
  p.image = static_cast<const uint8_t*>(m.tc_image);
  p.abort_flag = ctx->abort_flag;
  p.B = B;
  p.y = y;
  p.soft = soft_out;
  p.hard = hard_out;
  p.hard_bytes = hard_bytes;
  p.last_act = m.acts[3];
  p.use_tma = (reinterpret_cast<uintptr_t>(y) % 16 == 0) ? 1 : 0;
  p.L = 7;
  if (reinterpret_cast<uintptr_t>(y) % 4 != 0) { set_error("y must be 4-byte aligned"); return NVD_ERR_INVALID; }
  if (soft_out && reinterpret_cast<uintptr_t>(soft_out) % 16 != 0) { set_error("soft_out must be 16-byte aligned"); return NVD_ERR_INVALID; }
  return launch_tc<MODE_STAGED>(ctx, p);
}

int launch_mlp_tc_fused(nvd_ctx* ctx, const MlpModel& m, int code, int L, float sigma, int snr_idx, uint64_t frame0,
                        uint64_t frames, uint64_t* counters) {
  if (frames == 0) return NVD_OK;
  if (2 * L != NIN) { set_error("fused tcgen05 pipeline is compiled for L = %d (got %d)", NIN / 2, L); return NVD_ERR_UNSUPPORTED; }
  TcParams p{};
  p.image = static_cast<const uint8_t*>(m.tc_image);
  p.abort_flag = ctx->abort_flag;
  p.B = static_cast<int64_t>(frames);
  p.seed = ctx->seed;
  p.frame0 = frame0;
  p.counters = counters;
  p.sigma = sigma;
  p.stream_word = noise_stream_word(snr_idx);
  p.code = code;
  p.L = L;
  p.last_act = m.acts[3];
  return launch_tc<MODE_FUSED>(ctx, p);
}

}  // namespace nvd

extern "C" int nvd_debug_umma_probe(nvd_ctx* ctx, int N, int a_lbo_bytes, int a_sbo_bytes, int b_lbo_bytes,
                                    int b_sbo_bytes, const void* a_image_host, const void* b_image_host, float* d_host) {
  using namespace nvd;
  NVD_REQUIRE(ctx && a_image_host && b_image_host && d_host, "null argument");
  NVD_REQUIRE(N == 32 || N == 64 || N == 128 || N == 256, "N must be 32, 64, 128 or 256");
  NVD_CUDA(cudaSetDevice(ctx->device));
  void *da = nullptr, *db = nullptr, *dd = nullptr;
  NVD_CUDA(cudaMalloc(&da, 4096));
  NVD_CUDA(cudaMalloc(&db, static_cast<size_t>(N) * 32));
  NVD_CUDA(cudaMalloc(&dd, static_cast<size_t>(128) * N * 4));
  NVD_CUDA(cudaMemcpyAsync(da, a_image_host, 4096, cudaMemcpyHostToDevice, ctx->stream));
  NVD_CUDA(cudaMemcpyAsync(db, b_image_host, static_cast<size_t>(N) * 32, cudaMemcpyHostToDevice, ctx->stream));
  NVD_CUDA(cudaMemsetAsync(dd, 0, static_cast<size_t>(128) * N * 4, ctx->stream));
  const int smem = 4096 + 8192 + 64;
  umma_probe_kernel<<<1, 128, smem, ctx->stream>>>(N, a_lbo_bytes, a_sbo_bytes, b_lbo_bytes, b_sbo_bytes,
                                                    static_cast<const uint4*>(da), static_cast<const uint4*>(db),
                                                    static_cast<float*>(dd), ctx->abort_flag);
  NVD_CHECK_LAUNCH();
  NVD_CUDA(cudaMemcpyAsync(d_host, dd, static_cast<size_t>(128) * N * 4, cudaMemcpyDeviceToHost, ctx->stream));
  const int s = nvd_sync(ctx);
  cudaFree(da); cudaFree(db); cudaFree(dd);
  return s;
}
