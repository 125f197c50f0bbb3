// K3-kstream: tall-and-skinny decoders — a long input row against a narrow first layer — on tcgen05, and the fused
// Monte-Carlo pipeline of the per-bit predictor built on it.
//
// The reference net is fnn_viterbi_bitwise's build_network (/root/reference/src/fnn_viterbi_bitwise.py:326-399) as it is
// actually wired: out = relu(relu(x W1 + b1) Wout + bout), x = the 2(L+6) = 1036 noisy symbols of one frame, W1
// [1036, 30], Wout [30, 2]; decision = argmax (:504) = the predicted value of info bit `bit_position` (:456-457).
// 31 080 MAC per frame over 2 072 B of bf16 input: 30 FLOP/B, HBM-bound when the symbols come from memory and
// RNG-bound when they are generated on chip — never tensor-bound, so the design spends nothing on
// the MMA side and everything on feeding it:
//   * 128 frames = one M=128 UMMA tile; the first layer is a K loop of `tcgen05.mma.kind::f16` (bf16 x bf16 -> fp32,
//     N = 32 or 64) whose B operand, the whole W1 image (68 KB), stays in shared memory for the kernel's lifetime;
//   * the A operand is STREAMED through a ring of shared-memory stages of 48 symbols (3 K steps) in the UMMA
//     canonical K-major layout, SS mode, so nothing passes through registers on its way to the tensor core:
//       MODE_STAGED  rows in HBM (bf16, any 4-byte-aligned stride): every thread issues `cp.async` copies that
//                    scatter 8/16-byte pieces straight into the canonical layout, three stages ahead of the MMAs;
//       MODE_FUSED   thread r generates frame r: Philox info bits (+ the six start-state bits), the convolutional
//                    encoder 24 time steps at a time, 8 Philox noise calls per stage, BPSK (+ the reference's int32
//                    truncation, :309-317, on request), bf16 pack, 16-byte stores into the same layout.  Same device
//                    functions as K0/K1/K2, so fused and staged frames are bit-identical;
//   * four independent 128-thread groups per CTA (one persistent CTA per SM) keep four tile streams in flight;
//   * epilogue in registers: tcgen05.ld of the 32/64 accumulators, + b1, activation, the tiny second layer
//     (<= 64 x 8) as fp32 FMAs against broadcast shared-memory weights, activation, argmax (first index on ties);
//     fused mode compares with the label bit and only four counters leave the SM.
// All mbarrier waits are bounded (abort flag), as in k_mlp_tc.cu.
#include <cuda.h>      // CUtensorMap and its enums only: the encoder is fetched with cudaGetDriverEntryPoint

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "frame_gen.cuh"
#include "ptx.cuh"

namespace nvd {

namespace {

constexpr int kGroups = 4;
constexpr int kThreads = 128 * kGroups;
constexpr int kBatchSym = 48;                       // symbols per stage: 3 K steps = 8 Philox noise calls = 24 time steps
constexpr int kChunksPerStage = kBatchSym / 8;      // 16-byte K chunks per row and stage
constexpr int kChunkBytes = 128 * 16;               // one K chunk of all 128 rows
constexpr int kStageBytes = kChunksPerStage * kChunkBytes;   // 12 KB
constexpr int kStagesStaged = 3, kStagesFused = 2;
constexpr int kMaxOut = 8;
constexpr uint32_t kMaxWaitMs = 2000;
constexpr int kSBO = 128;

enum { KS_STAGED = 0, KS_FUSED = 1 };

struct KsParams {
  const uint8_t* image;     // bf16 W1 image: [Kpad / 8][N1pad][8] (UMMA canonical K-major, no swizzle)
  const float* l2;          // fp32: b1[N1pad] | W2[N1pad][kMaxOut] | b2[kMaxOut]
  int* abort_flag;
  int64_t B;
  int K, Kpad, N1pad, n_out, act1, act2;
  // staged
  const __nv_bfloat16* y;
  int64_t row_stride;       // elements between rows
  int cp_bytes;             // 4, 8 or 16: cp.async piece size allowed by base / stride alignment
  float* soft;
  void* hard;
  int hard_bytes;
  // fused
  PhiloxKeys rk;
  uint64_t frame0;
  uint64_t* counters;
  float sigma;
  uint32_t stream_word;
  int code, L;              // info bits per frame
  uint32_t prefix_bits;     // time-packed start-state bits prepended to the info bits (extract_state_bits, :114-138)
  int n_prefix;
  int bit_position;
  int trunc;                // int32-truncation quirk of the reference's data sets
};

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// asynchronous global -> shared copy of BYTES (4, 8, 16) with zero fill beyond src_bytes
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(uint32_t dst, const void* src, int src_bytes) {
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(dst), "l"(src), "n"(BYTES), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float apply_act(float v, int act) {
  return act == NVD_ACT_TANH ? tanhf(v) : act == NVD_ACT_RELU ? fmaxf(v, 0.0f) : v;
}

// ---- long-frame generator (fused mode): one thread = one frame, 24 time steps = 48 symbols per call -----------------
struct FrameGen {
  uint64_t frame;          // global frame index
  uint64_t fifo;           // pending bits of (prefix ++ info), time order from bit 0
  int fifo_bits;
  uint32_t next_word;      // next 32-bit word of the Philox info stream
  uint4 blk;               // current Philox info block
  uint32_t prev24;         // the previous call's 24 bits (encoder history)
  uint32_t label;          // info bit `bit_position`
};

__device__ __forceinline__ void gen_init(FrameGen& g, const KsParams& p, uint64_t frame) {
  g.frame = frame;
  g.fifo = p.prefix_bits;
  g.fifo_bits = p.n_prefix;
  g.next_word = 0;
  g.blk = make_uint4(0, 0, 0, 0);
  g.prev24 = 0;
  g.label = 0;
}

// 48 symbols of stage `b` of the frame -> 24 packed bf16 pairs
__device__ __forceinline__ void gen_stage(FrameGen& g, const KsParams& p, int b, uint32_t (&w)[24]) {
  // top the FIFO up to >= 24 bits from the Philox info stream (K0: stream kStreamInfo, 128 bits per call)
  while (g.fifo_bits < 24) {
    const uint32_t wi = g.next_word;
    if ((wi & 3u) == 0u) g.blk = frame_words(p.rk, g.frame, wi >> 2, kStreamInfo);
    uint32_t word = (wi & 3u) == 0u ? g.blk.x : (wi & 3u) == 1u ? g.blk.y : (wi & 3u) == 2u ? g.blk.z : g.blk.w;
    const int valid = p.L - 32 * static_cast<int>(wi);
    if (valid <= 0) word = 0u;
    else if (valid < 32) word &= (1u << valid) - 1u;
    if (static_cast<int>(wi) == (p.bit_position >> 5)) g.label = (word >> (p.bit_position & 31)) & 1u;
    g.fifo |= static_cast<uint64_t>(word) << g.fifo_bits;
    g.fifo_bits += 32;
    g.next_word = wi + 1;
  }
  const int total_bits = p.n_prefix + p.L;
  const int remaining = total_bits - 24 * b;                    // time steps left in the frame (may be <= 0)
  uint32_t u = static_cast<uint32_t>(g.fifo) & 0xFFFFFFu;
  if (remaining < 24) u &= remaining > 0 ? ((1u << remaining) - 1u) : 0u;
  g.fifo >>= 24;
  g.fifo_bits -= 24;
  // encoder: 24 time steps at once on the packed word (frame_gen.cuh conv_outputs; K1 does the same per byte)
  uint32_t oa, ob;
  conv_outputs((static_cast<uint64_t>(u) << 6) | (g.prev24 >> 18), p.code, oa, ob);
  if (p.code == NVD_CODE_57_REF && b == 0) ob ^= (u >> 1) & 1u;
  g.prev24 = u;
  const uint32_t c_lo = spread16(oa) | (spread16(ob) << 1);                 // symbols 0..31 of the stage
  const uint32_t c_hi = spread16(oa >> 16) | (spread16(ob >> 16) << 1);     // symbols 32..47
  const int valid_sym = 2 * remaining;                                      // symbols of this stage inside the frame
#pragma unroll
  for (int nb = 0; nb < 6; ++nb) {       // six Philox blocks of eight symbols
    constexpr int kPerBlock = kNormalsPerBlock;
    const int j0 = kPerBlock * nb;                                          // symbol within the stage (compile time)
    const uint32_t cw = j0 < 32 ? c_lo : c_hi;
    float v[8];
    block_symbols<4>(p.rk, g.frame, static_cast<uint32_t>(6 * b + nb), p.stream_word, bm_scale(p.sigma), cw >> (j0 & 31), v);
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      const int j = j0 + k;
      float a = v[k], c = v[k + 1];
      if (p.trunc) { a = truncf(a); c = truncf(c); }
      w[j >> 1] = j < valid_sym ? pack2(a, c) : 0u;
    }
  }
}

// Tile epilogue of one thread = one frame: the N1pad accumulators of its TMEM lane -> + b1, activation, the small
// second layer in fp32 against broadcast shared-memory weights, activation, argmax (first index on ties, tf.argmax).
// Staged outputs are stored here; returns the decision.
__device__ __forceinline__ int tile_epilogue(const KsParams& p, const float* l2s, uint32_t tmem_d, int64_t row, bool store) {
  float out[kMaxOut];
#pragma unroll
  for (int o = 0; o < kMaxOut; ++o) out[o] = l2s[p.N1pad + p.N1pad * kMaxOut + o];
  for (int c = 0; c < p.N1pad; c += 32) {
    uint32_t v[32];
    ptx::tmem_ld32(tmem_d + c, v);
    ptx::tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float h = apply_act(__uint_as_float(v[j]) + l2s[c + j], p.act1);
      const float4 wa = *reinterpret_cast<const float4*>(l2s + p.N1pad + (c + j) * kMaxOut);
      const float4 wb = *reinterpret_cast<const float4*>(l2s + p.N1pad + (c + j) * kMaxOut + 4);
      out[0] = fmaf(h, wa.x, out[0]); out[1] = fmaf(h, wa.y, out[1]); out[2] = fmaf(h, wa.z, out[2]); out[3] = fmaf(h, wa.w, out[3]);
      out[4] = fmaf(h, wb.x, out[4]); out[5] = fmaf(h, wb.y, out[5]); out[6] = fmaf(h, wb.z, out[6]); out[7] = fmaf(h, wb.w, out[7]);
    }
  }
  ptx::tc_fence_before();            // accumulator reads done before the next MMA chain overwrites them
  int best = 0;
  float bv = 0.0f;
#pragma unroll
  for (int o = 0; o < kMaxOut; ++o) {
    if (o < p.n_out) {
      out[o] = apply_act(out[o], p.act2);
      if (o == 0 || out[o] > bv) { bv = out[o]; best = o; }     // strict: first index on ties
    }
  }
  if (store && row < p.B) {
    if (p.soft)
      for (int o = 0; o < p.n_out; ++o) p.soft[row * p.n_out + o] = out[o];
    if (p.hard) {
      if (p.hard_bytes == 1) static_cast<uint8_t*>(p.hard)[row] = static_cast<uint8_t>(best);
      else if (p.hard_bytes == 2) static_cast<uint16_t*>(p.hard)[row] = static_cast<uint16_t>(best);
      else static_cast<int32_t*>(p.hard)[row] = best;
    }
  }
  return best;
}

// G: 128-thread groups per CTA (4 for the staged forms; the fused form, which is bound by its own instruction stream and
// not by shared memory, takes more: NVD_KS_FUSED_GROUPS)
template <int MODE, int CP, int G>
__global__ void __launch_bounds__(128 * G, 1) mlp_kstream_kernel(const KsParams p) {
  constexpr int kGroups = G;
  constexpr int kThreads = 128 * G;
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int S = MODE == KS_FUSED ? kStagesFused : kStagesStaged;
  const uint32_t sbase = ptx::smem_u32(smem);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  const int g = warp >> 2, wq = warp & 3;
  const bool issuer_warp = wq == 0;
  const int r = tid & 127;
  const int w_bytes = (p.Kpad / 8) * p.N1pad * 16;
  const int l2_floats = p.N1pad + p.N1pad * kMaxOut + kMaxOut;
  const int l2_off = w_bytes;
  const int bar_off = (l2_off + l2_floats * 4 + 15) & ~15;
  const int ring_off = (bar_off + 256 + 1023) & ~1023;
  const uint32_t grp_ring = sbase + ring_off + g * S * kStageBytes;
  const uint32_t sbase16 = (sbase & 0x3FFFFu) >> 4, ring16 = (grp_ring & 0x3FFFFu) >> 4;
  const uint32_t bar_free = sbase + bar_off + 32 * g;          // [S] per group: the MMAs reading stage s have completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + bar_off + 192);
  const float* l2s = reinterpret_cast<const float*>(smem + l2_off);

  for (int i = tid; i < w_bytes / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(p.image) + i);
  for (int i = tid; i < l2_floats; i += kThreads) reinterpret_cast<float*>(smem + l2_off)[i] = __ldg(p.l2 + i);
  if (tid == 0) {
    for (int i = 0; i < kGroups; ++i)
      for (int s = 0; s < S; ++s) ptx::mbar_init(sbase + bar_off + 32 * i + 8 * s, 1);
    ptx::mbar_fence_init();
  }
  uint32_t tmem_cols = 32;                                                  // power of two >= G * N1pad
  while (tmem_cols < static_cast<uint32_t>(kGroups * p.N1pad)) tmem_cols *= 2;
  if (tid < 32) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xFFFFFFFFu, *tmem_slot, 0);
  const uint32_t tmem_grp = tmem_base + g * p.N1pad;
  const uint32_t tmem_d = tmem_grp + (static_cast<uint32_t>(wq * 32) << 16);
  const uint32_t idesc = ptx::make_idesc_bf16_f32(128, p.N1pad);
  const uint32_t b_lbo = p.N1pad * 16;

  const int nb = p.Kpad / kBatchSym;                              // stages per tile
  const int64_t num_tiles = (p.B + 127) >> 7;
  const int64_t tile_stride = static_cast<int64_t>(gridDim.x) * kGroups;
  const int64_t tile0 = static_cast<int64_t>(blockIdx.x) * kGroups + g;
  const int64_t my_tiles = tile0 < num_tiles ? (num_tiles - tile0 + tile_stride - 1) / tile_stride : 0;
  const int64_t items = my_tiles * nb;                             // (tile, stage) pairs of this group, in order

  uint32_t bit_err = 0, frame_err = 0;
  bool ok = true;

  // ---- staged: asynchronous copies of item `it` (tile, stage) into ring slot `slot`; 128 threads cover 128 rows x 96 B
  auto issue_copies = [&](int64_t tile, int b, int slot) {
    // a row's 96 bytes of the stage are three 32-byte sectors; 32 / CP neighbouring lanes share one sector (whole
    // sectors from L2), and the 128 / (32 / CP) sectors of one instruction belong to different rows, i.e. to
    // different 16-byte slots of one K chunk (no more than a 2-way shared-memory conflict)
    constexpr int kLanesPerSector = 32 / CP;
    constexpr int kIters = 3 * kLanesPerSector;
    const uint32_t dst0 = grp_ring + slot * kStageBytes;
    const int64_t row0 = tile * 128;
    const int k0 = b * kBatchSym;
#pragma unroll
    for (int i = 0; i < kIters; ++i) {
      const int unit = i * (128 / kLanesPerSector) + r / kLanesPerSector;     // sector-major: unit = sector * 128 + row
      const int row = unit & 127, sector = unit >> 7;
      const int byte_in_stage = sector * 32 + (r % kLanesPerSector) * CP;       // 0 .. 95
      const int k = k0 + (byte_in_stage >> 1);                                  // first element of the piece
      int bytes = (p.K - k) * 2;
      bytes = bytes < 0 ? 0 : bytes > CP ? CP : bytes;
      if (row0 + row >= p.B) bytes = 0;
      const __nv_bfloat16* src = p.y + (bytes ? (row0 + row) * p.row_stride + k : 0);
      cp_async_zfill<CP>(dst0 + (byte_in_stage >> 4) * kChunkBytes + row * 16 + (byte_in_stage & 15), src, bytes);
    }
  };

  FrameGen gen;
  int64_t tile = tile0;
  int b = 0;                                                       // stage within the tile
  if (MODE == KS_STAGED) {
    // prefetch S - 1 items
    int64_t pt = tile0;
    int pb = 0;
    for (int i = 0; i < S - 1; ++i) {
      if (i < items) issue_copies(pt, pb, i % S);
      cp_async_commit();
      if (++pb == nb) { pb = 0; pt += tile_stride; }
    }
  }
  int64_t ptile = tile0;                                           // (tile, stage) of item it + S - 1 (staged prefetch)
  int pb2 = 0;
  if (MODE == KS_STAGED) {
    for (int i = 0; i < S - 1; ++i)
      if (++pb2 == nb) { pb2 = 0; ptile += tile_stride; }
  }

  for (int64_t it = 0; it < items && ok; ++it) {
    const int slot = static_cast<int>(it % S);
    const uint32_t use = static_cast<uint32_t>(it / S);            // this is the use-th time the slot is filled
    if (MODE == KS_FUSED) {
      if (b == 0) gen_init(gen, p, p.frame0 + static_cast<uint64_t>(tile * 128 + r));
      uint32_t w[24];
      gen_stage(gen, p, b, w);
      if (use > 0 && !ptx::mbar_wait_bounded(bar_free + 8 * slot, (use - 1) & 1u, kMaxWaitMs)) { atomicExch(p.abort_flag, 21); ok = false; }
      const uint32_t dst = grp_ring + slot * kStageBytes + r * 16;
#pragma unroll
      for (int q = 0; q < kChunksPerStage; ++q) st_shared_v4(dst + q * kChunkBytes, w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    } else {
      cp_async_wait<S - 2>();                                       // this thread's copies of item `it` have landed
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    ptx::bar_sync(1 + g, 128);
    if (issuer_warp && ptx::elect_one()) {
      ptx::tc_fence_after();
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int kstep = 3 * b + j;
        ptx::umma_f16_ss(tmem_grp, ptx::make_smem_desc16(ring16, slot * kStageBytes + j * 2 * kChunkBytes, kChunkBytes, kSBO),
                         ptx::make_smem_desc16(sbase16, kstep * 2 * b_lbo, b_lbo, kSBO), idesc, (b > 0 || j > 0) ? 1u : 0u);
      }
      ptx::umma_commit(bar_free + 8 * slot);
    }
    if (MODE == KS_STAGED) {
      // refill the slot of item it - 1 (its MMAs were issued one iteration ago) with item it + S - 1
      const int64_t nit = it + S - 1;
      if (nit < items) {
        const int nslot = static_cast<int>(nit % S);
        if (it >= 1) {
          const uint32_t puse = static_cast<uint32_t>((it - 1) / S);
          if (!ptx::mbar_wait_bounded(bar_free + 8 * nslot, puse & 1u, kMaxWaitMs)) { atomicExch(p.abort_flag, 22); ok = false; }
        }
        issue_copies(ptile, pb2, nslot);
      }
      cp_async_commit();
      if (++pb2 == nb) { pb2 = 0; ptile += tile_stride; }
    }

    if (b + 1 == nb) {
      // ---- tile epilogue: all MMAs of the tile complete when the last stage's commit arrives ----
      if (!ptx::mbar_wait_bounded(bar_free + 8 * slot, use & 1u, kMaxWaitMs)) { atomicExch(p.abort_flag, 23); ok = false; break; }
      ptx::tc_fence_after();
      const int64_t row = tile * 128 + r;
      const int best = tile_epilogue(p, l2s, tmem_d, row, MODE == KS_STAGED);
      if (MODE == KS_FUSED && row < p.B) {
        const uint32_t e = (static_cast<uint32_t>(best) ^ gen.label) & 1u;
        bit_err += e;
        frame_err += e;
      }
      b = 0;
      tile += tile_stride;
    } else {
      ++b;
    }
  }
  if (MODE == KS_STAGED) cp_async_wait<0>();

  if (MODE == KS_FUSED) {
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
      atomicAdd(reinterpret_cast<unsigned long long*>(p.counters + 3), static_cast<unsigned long long>(p.B));   // one decoded bit per frame
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem_base, tmem_cols);
}

// ---- MODE_STAGED with TMA: rows whose stride is a multiple of 16 bytes --------------------------------------------
// One `cp.async.bulk.tensor.2d` per stage: a [128 rows x 64 symbols] box of the row-major bf16 matrix lands in shared
// memory in the SWIZZLE_128B K-major layout the UMMA descriptor names (16 KB; symbols beyond K and rows beyond B are
// zero-filled by the copy engine), so the stream costs no LSU instructions at all — the cp.async form above is bound by
// LDGSTS wavefronts at 21 % of HBM.  Per group: a driver warp (waits for a stage, issues its four MMAs, commits, re-arms
// the stage it just freed) and four epilogue warps that only wake up once per tile; two accumulator buffers per group
// let the epilogue of tile t overlap the stream of tile t + 1.
constexpr int kTmaSym = 64;
constexpr int kTmaStageBytes = 128 * kTmaSym * 2;   // 16 KB
constexpr int kTmaStages = 2;
constexpr int kTmaThreads = kThreads + 32 * kGroups;

__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t addr) {
  // K-major, SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart (SBO); the K step inside the atom is selected by the
  // start address (+32 B per 16 elements); LBO is not used by swizzled K-major layouts
  const uint32_t lo = ((addr & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);       // SBO, descriptor version 1, layout type 2
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(kTmaThreads, 1) mlp_kstream_tma_kernel(const KsParams p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int S = kTmaStages;
  const uint32_t sbase = ptx::smem_u32(smem);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  const bool is_driver = warp >= kThreads / 32;
  const int g = is_driver ? warp - kThreads / 32 : warp >> 2;
  const int wq = warp & 3, lane = tid & 31;
  const int r = tid & 127;
  const int w_bytes = (p.Kpad / 8) * p.N1pad * 16;
  const int l2_floats = p.N1pad + p.N1pad * kMaxOut + kMaxOut;
  const int l2_off = w_bytes;
  const int bar_off = (l2_off + l2_floats * 4 + 15) & ~15;
  const int ring_off = (bar_off + 512 + 1023) & ~1023;
  const uint32_t grp_ring = sbase + ring_off + g * S * kTmaStageBytes;
  const uint32_t sbase16 = (sbase & 0x3FFFFu) >> 4;
  const uint32_t bar_full = sbase + bar_off + 64 * g;            // [S] stage landed (expect_tx)
  const uint32_t bar_free = bar_full + 8 * S;                    // [S] the stage's MMAs have completed
  const uint32_t bar_done = bar_free + 8 * S;                    // [2] accumulator buffer d holds a finished tile
  const uint32_t bar_dfree = bar_done + 16;                      // [2] ... and has been read out (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + bar_off + 64 * kGroups);
  const float* l2s = reinterpret_cast<const float*>(smem + l2_off);

  for (int i = tid; i < w_bytes / 16; i += kTmaThreads) reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(p.image) + i);
  for (int i = tid; i < l2_floats; i += kTmaThreads) reinterpret_cast<float*>(smem + l2_off)[i] = __ldg(p.l2 + i);
  if (tid == 0) {
    for (int i = 0; i < kGroups; ++i) {
      const uint32_t b0 = sbase + bar_off + 64 * i;
      for (int s2 = 0; s2 < S; ++s2) { ptx::mbar_init(b0 + 8 * s2, 1); ptx::mbar_init(b0 + 8 * S + 8 * s2, 1); }
      for (int d = 0; d < 2; ++d) { ptx::mbar_init(b0 + 16 * S + 8 * d, 1); ptx::mbar_init(b0 + 16 * S + 16 + 8 * d, 4); }
    }
    ptx::mbar_fence_init();
  }
  const uint32_t tmem_cols = static_cast<uint32_t>(2 * kGroups * p.N1pad);     // 256 or 512: two accumulators per group
  if (tid < 32) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xFFFFFFFFu, *tmem_slot, 0);
  const uint32_t tmem_grp = tmem_base + g * 2 * p.N1pad;
  const uint32_t idesc = ptx::make_idesc_bf16_f32(128, p.N1pad);
  const uint32_t b_lbo = p.N1pad * 16;
  const int nb = (p.K + kTmaSym - 1) / kTmaSym;                  // stages per tile
  const int ksteps = p.Kpad / 16;                                 // K steps the weight image holds (the rest of A is zero)
  const int64_t num_tiles = (p.B + 127) >> 7;
  const int64_t tile_stride = static_cast<int64_t>(gridDim.x) * kGroups;
  const int64_t tile0 = static_cast<int64_t>(blockIdx.x) * kGroups + g;
  const int64_t my_tiles = tile0 < num_tiles ? (num_tiles - tile0 + tile_stride - 1) / tile_stride : 0;

  if (is_driver) {
    const int64_t items = my_tiles * nb;
    bool ok = true;
    auto load_item = [&](int64_t it) {            // elected lane: arm the stage's barrier and start its copy
      const int slot = static_cast<int>(it % S);
      const int64_t tl = it / nb;
      const int b = static_cast<int>(it - tl * nb);
      ptx::mbar_arrive_expect_tx(bar_full + 8 * slot, kTmaStageBytes);
      tma_load_2d(grp_ring + slot * kTmaStageBytes, &tmap, b * kTmaSym, static_cast<int>((tile0 + tl * tile_stride) * 128), bar_full + 8 * slot);
    };
    if (ptx::elect_one())
      for (int64_t it = 0; it < S - 1 && it < items; ++it) load_item(it);
    int64_t it = 0;
    for (int64_t tl = 0; tl < my_tiles && ok; ++tl) {
      const int d = static_cast<int>(tl & 1);
      if (tl >= 2 && !ptx::mbar_wait_bounded(bar_dfree + 8 * d, static_cast<uint32_t>((tl >> 1) - 1) & 1u, kMaxWaitMs)) { atomicExch(p.abort_flag, 24); ok = false; break; }
      for (int b = 0; b < nb && ok; ++b, ++it) {
        const int slot = static_cast<int>(it % S);
        if (!ptx::mbar_wait_bounded(bar_full + 8 * slot, static_cast<uint32_t>(it / S) & 1u, kMaxWaitMs)) { atomicExch(p.abort_flag, 25); ok = false; break; }
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int j = 0; j < kTmaSym / 16; ++j) {
            const int kstep = (kTmaSym / 16) * b + j;
            if (kstep < ksteps)
              ptx::umma_f16_ss(tmem_grp + d * p.N1pad, make_smem_desc_sw128(grp_ring + slot * kTmaStageBytes + j * 32),
                               ptx::make_smem_desc16(sbase16, kstep * 2 * b_lbo, b_lbo, kSBO), idesc, (b > 0 || j > 0) ? 1u : 0u);
          }
          ptx::umma_commit(bar_free + 8 * slot);
          if (b == nb - 1) ptx::umma_commit(bar_done + 8 * d);
        }
        // the stage of item it - 1 (its MMAs were issued one iteration ago) takes item it + S - 1
        const int64_t nit = it + S - 1;
        if (nit < items) {
          if (it >= 1) {
            const int pslot = static_cast<int>((it - 1) % S);
            if (!ptx::mbar_wait_bounded(bar_free + 8 * pslot, static_cast<uint32_t>((it - 1) / S) & 1u, kMaxWaitMs)) { atomicExch(p.abort_flag, 26); ok = false; break; }
          }
          if (ptx::elect_one()) load_item(nit);
        }
      }
    }
  } else {
    const uint32_t tmem_d = tmem_grp + (static_cast<uint32_t>(wq * 32) << 16);
    for (int64_t tl = 0; tl < my_tiles; ++tl) {
      const int d = static_cast<int>(tl & 1);
      if (!ptx::mbar_wait_bounded(bar_done + 8 * d, static_cast<uint32_t>(tl >> 1) & 1u, kMaxWaitMs)) { atomicExch(p.abort_flag, 27); break; }
      ptx::tc_fence_after();
      tile_epilogue(p, l2s, tmem_d + d * p.N1pad, (tile0 + tl * tile_stride) * 128 + r, true);
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_dfree + 8 * d);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem_base, tmem_cols);
}

int ks_tma_smem_bytes(int Kpad, int N1pad) {
  const int w_bytes = (Kpad / 8) * N1pad * 16;
  const int l2_floats = N1pad + N1pad * kMaxOut + kMaxOut;
  const int bar_off = (w_bytes + l2_floats * 4 + 15) & ~15;
  const int ring_off = (bar_off + 512 + 1023) & ~1023;
  return ring_off + kGroups * kTmaStages * kTmaStageBytes;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libnvd links only the CUDA runtime)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// returns NVD_ERR_UNSUPPORTED when the rows cannot be described to the copy engine (the caller then uses cp.async)
int launch_ks_tma(nvd_ctx* ctx, const KsParams& p) {
  const EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || getenv("NVD_KSTREAM_NO_TMA")) return NVD_ERR_UNSUPPORTED;
  if (ks_tma_smem_bytes(p.Kpad, p.N1pad) > 227 * 1024) return NVD_ERR_UNSUPPORTED;
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(p.K), static_cast<cuuint64_t>(p.B)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(p.row_stride) * 2};
  const cuuint32_t box[2] = {kTmaSym, 128};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(p.y), gdim, gstride, box, estride,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return NVD_ERR_UNSUPPORTED;
  const int smem = ks_tma_smem_bytes(p.Kpad, p.N1pad);
  NVD_CUDA(cudaFuncSetAttribute(mlp_kstream_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int64_t tiles = (p.B + 127) / 128;
  const int grid = static_cast<int>(std::min<int64_t>(ctx->sm_count, (tiles + kGroups - 1) / kGroups));
  mlp_kstream_tma_kernel<<<grid, kTmaThreads, smem, ctx->stream>>>(p, tmap);
  NVD_CHECK_LAUNCH();
  return NVD_OK;
}

uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

int ks_smem_bytes(int Kpad, int N1pad, int stages, int groups = kGroups) {
  const int w_bytes = (Kpad / 8) * N1pad * 16;
  const int l2_floats = N1pad + N1pad * kMaxOut + kMaxOut;
  const int bar_off = (w_bytes + l2_floats * 4 + 15) & ~15;
  const int ring_off = (bar_off + 256 + 1023) & ~1023;
  return ring_off + groups * stages * kStageBytes;
}

template <int MODE, int CP, int G = kGroups>
int launch_ks(nvd_ctx* ctx, const KsParams& p) {
  auto kern = mlp_kstream_kernel<MODE, CP, G>;
  const int smem = ks_smem_bytes(p.Kpad, p.N1pad, MODE == KS_FUSED ? kStagesFused : kStagesStaged, G);
  if (smem > 227 * 1024 || G * p.N1pad > 512) { set_error("K-streamed kernel: %d groups do not fit (shared memory %d B)", G, smem); return NVD_ERR_UNSUPPORTED; }
  NVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int64_t tiles = (p.B + 127) / 128;
  const int grid = static_cast<int>(std::min<int64_t>(ctx->sm_count, (tiles + G - 1) / G));
  kern<<<grid, 128 * G, smem, ctx->stream>>>(p);
  NVD_CHECK_LAUNCH();
  return NVD_OK;
}

void fill_common(KsParams& p, nvd_ctx* ctx, const MlpModel& m) {
  p.image = static_cast<const uint8_t*>(m.ks_image);
  p.l2 = m.ks_l2;
  p.abort_flag = ctx->abort_flag;
  p.K = m.dims[0];
  p.Kpad = m.ks_Kpad;
  p.N1pad = m.ks_N1pad;
  p.n_out = m.dims[2];
  p.act1 = m.acts[0];
  p.act2 = m.acts[1];
}

}  // namespace

bool kstream_supported(const MlpModel& m) {
  if (m.n_layers != 2 || m.dims[1] > 64 || m.dims[2] > kMaxOut || m.dims[0] < 16) return false;
  if (m.dims[0] & 1) return false;                      // bf16 rows must be 4-byte aligned: even input width
  const int Kpad = (m.dims[0] + kBatchSym - 1) / kBatchSym * kBatchSym;
  const int N1pad = m.dims[1] <= 32 ? 32 : 64;
  return ks_smem_bytes(Kpad, N1pad, kStagesStaged) <= 227 * 1024;
}

int kstream_build_image(MlpModel& m, const float* const* W_host, const float* const* b_host) {
  const int K = m.dims[0], N1 = m.dims[1], n_out = m.dims[2];
  const int Kpad = (K + kBatchSym - 1) / kBatchSym * kBatchSym;
  const int N1pad = N1 <= 32 ? 32 : 64;
  std::vector<uint8_t> img(static_cast<size_t>(Kpad / 8) * N1pad * 16, 0);
  for (int n = 0; n < N1; ++n)
    for (int k = 0; k < K; ++k) {
      const uint16_t v = f32_to_bf16_rn(W_host[0][static_cast<size_t>(k) * N1 + n]);
      memcpy(img.data() + static_cast<size_t>(k >> 3) * (N1pad * 16) + n * 16 + (k & 7) * 2, &v, 2);
    }
  std::vector<float> l2(N1pad + N1pad * kMaxOut + kMaxOut, 0.0f);
  for (int n = 0; n < N1; ++n) l2[n] = b_host[0][n];
  for (int n = 0; n < N1; ++n)
    for (int o = 0; o < n_out; ++o) l2[N1pad + n * kMaxOut + o] = W_host[1][static_cast<size_t>(n) * n_out + o];
  for (int o = 0; o < n_out; ++o) l2[N1pad + N1pad * kMaxOut + o] = b_host[1][o];
  NVD_CUDA(cudaMalloc(&m.ks_image, img.size()));
  NVD_CUDA(cudaMemcpy(m.ks_image, img.data(), img.size(), cudaMemcpyHostToDevice));
  NVD_CUDA(cudaMalloc(&m.ks_l2, l2.size() * sizeof(float)));
  NVD_CUDA(cudaMemcpy(m.ks_l2, l2.data(), l2.size() * sizeof(float), cudaMemcpyHostToDevice));
  m.ks_Kpad = Kpad;
  m.ks_N1pad = N1pad;
  return NVD_OK;
}

int launch_mlp_kstream(nvd_ctx* ctx, const MlpModel& m, const __nv_bfloat16* y, int64_t row_stride, int64_t B, float* soft_out,
                       void* hard_out, int hard_bytes) {
  if (B == 0) return NVD_OK;
  KsParams p{};
  fill_common(p, ctx, m);
  p.B = B;
  p.y = y;
  p.row_stride = row_stride;
  p.soft = soft_out;
  p.hard = hard_out;
  p.hard_bytes = hard_bytes;
  const uintptr_t a = reinterpret_cast<uintptr_t>(y);
  const int64_t sb = row_stride * 2;
  if (a % 4 != 0 || sb % 4 != 0) { set_error("bf16 rows must be 4-byte aligned (base and stride)"); return NVD_ERR_INVALID; }
  // a piece must not straddle the end of the row's valid data unevenly: K * 2 bytes is a multiple of 4 for even K only
  if ((p.K & 1) != 0) { set_error("K-streamed tensor-core path needs an even input width (got %d)", p.K); return NVD_ERR_UNSUPPORTED; }
  if (a % 16 == 0 && sb % 16 == 0) {
    const int st = launch_ks_tma(ctx, p);              // 16-byte aligned rows: the copy engine streams them
    if (st != NVD_ERR_UNSUPPORTED) return st;
    p.cp_bytes = 16;
    return launch_ks<KS_STAGED, 16>(ctx, p);
  }
  if (a % 8 == 0 && sb % 8 == 0) { p.cp_bytes = 8; return launch_ks<KS_STAGED, 8>(ctx, p); }
  p.cp_bytes = 4;
  return launch_ks<KS_STAGED, 4>(ctx, p);
}

int launch_mlp_kstream_fused(nvd_ctx* ctx, const MlpModel& m, int code, int L, uint32_t prefix_bits, int n_prefix, int bit_position,
                             int trunc, float sigma, int snr_idx, uint64_t frame0, uint64_t frames, uint64_t* counters) {
  if (frames == 0) return NVD_OK;
  KsParams p{};
  fill_common(p, ctx, m);
  p.B = static_cast<int64_t>(frames);
  p.rk = make_philox_keys(ctx->seed);
  p.frame0 = frame0;
  p.counters = counters;
  p.sigma = sigma;
  p.stream_word = noise_stream_word(snr_idx);
  p.code = code;
  p.L = L;
  p.prefix_bits = prefix_bits;
  p.n_prefix = n_prefix;
  p.bit_position = bit_position;
  p.trunc = trunc;
  if (2 * (L + n_prefix) != p.K) { set_error("model input width %d != 2 (L + prefix) = %d", p.K, 2 * (L + n_prefix)); return NVD_ERR_INVALID; }
#ifndef NVD_KS_FUSED_GROUPS
#define NVD_KS_FUSED_GROUPS 6      // measured: 853 vs 832 Mframes/s with 4 (72 KB more ring, 78 registers)
#endif
  if (NVD_KS_FUSED_GROUPS != kGroups && ks_smem_bytes(p.Kpad, p.N1pad, kStagesFused, NVD_KS_FUSED_GROUPS) <= 227 * 1024 &&
      NVD_KS_FUSED_GROUPS * p.N1pad <= 512)
    return launch_ks<KS_FUSED, 16, NVD_KS_FUSED_GROUPS>(ctx, p);
  return launch_ks<KS_FUSED, 16>(ctx, p);
}

}  // namespace nvd
