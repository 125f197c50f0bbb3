// K3-wide: feed-forward decoders whose layers are too wide to keep every weight in shared memory —
// the sliding-window decoder 200 -> 500 -> 500 -> 500 -> 500 -> 2 of
// /root/reference/src/fnn_viterbi_seq_predict.py:343-414 (tanh after every affine, :408-412; 851 000 MAC per
// decoded bit) and any other MLP with layer widths <= 511 — on the tensor core, ONE LAYER PER LAUNCH:
//     Y[128 x N] = act( X[128 x K] W[K x N] + b ),  bf16 x bf16 -> fp32 accumulate, bf16 activations between layers.
// Per CTA (persistent, one per SM) and 128-row tile:
//   * the tile's X rows go HBM -> registers -> TENSOR memory (tcgen05.st, two K elements per column): the A operand
//     of every MMA of the tile comes from TMEM (TS mode), so shared memory carries only the weights;
//   * W streams from L2 through a ring of 16 KB shared-memory stages ([<=128 n-rows] x [64 k]) filled by a producer
//     warp with 1-D TMA bulk copies (the host pre-arranges W in the UMMA canonical K-major layout, chunked by 128
//     output columns, so every stage is one contiguous copy) and handed over with full/empty mbarriers;
//   * one issuer thread runs the K loop (4 MMAs per stage) into one of two 128-column accumulators; the four
//     epilogue warps (thread r = row r = TMEM lane r) read the other one, apply the activation and write bf16 rows
//     of the next layer's input, so the epilogue of chunk c overlaps the MMAs of chunk c+1;
//   * biases ride in the contraction: every X row carries 1.0 at column K (the previous layer's epilogue gets it
//     from a weight trick: bias row of that padding column = 20 -> tanh = 1, or 1 for relu / identity).
// The last layer's epilogue takes the argmax (first index on ties) and optional fp32 soft outputs.
// Row gathering (dense rows, or overlapping stride-2 windows of one long stream: get_single_instance,
// fnn_viterbi_seq_predict.py:293-336) is a small pad/convert kernel in front of the first layer.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace nvd {

namespace {

// The issuer thread reads the clock around each of its waits (bit 0: A ready, 1: accumulator free, 2: stage full,
// 3: issue) for the NVD_WIDE_DEBUG breakdown.  The reads stay in UNCONDITIONALLY (15): they pace the issue loop, and
// taking them out measured slower on B200 — 10^6 windows of the 200-500x4-2 decoder: 2.77 ms with all four,
// 2.85 ms with only the stage-full pair (4), 2.95 ms with none (0) or any other single one; __nanosleep after each
// stage instead is too coarse (3.25 ms); polling with test_wait instead of try_wait changes nothing either way.
#ifndef NVD_WIDE_TIMING_ALWAYS
#define NVD_WIDE_TIMING_ALWAYS 15
#endif
constexpr int kThreads = 320;            // warps 0-7: rows / epilogues (two sets of four), warp 8: weight producer, warp 9: MMA issuer
#ifdef NVD_WIDE_DIRECT_ROWS
constexpr int kStages = 13;             // experiment: no row staging buffer, 208 KB of weights in flight instead
#else
constexpr int kStages = 6;              // 96 KB of weights in flight next to the 128 KB row staging buffer
#endif
constexpr int kStageK = 64;              // k extent of one stage
constexpr int kStageBytes = 128 * kStageK * 2;   // 16 KB (a full 128-column chunk)
constexpr int kMaxKp = 512;
constexpr uint32_t COL_A = 256;          // A operand: columns 256 .. 256 + Kp/2; accumulators: 0..127 and 128..255
constexpr uint32_t kMaxWaitMs = 4000;
#ifdef NVD_WIDE_DIRECT_ROWS
constexpr int kAOperandBytes = 0;
#else
constexpr int kAOperandBytes = 128 * kMaxKp * 2;   // staging buffer of one tile's rows (tile-canonical layout), 128 KB
#endif

struct WideParams {
  const __nv_bfloat16* X;   // [B, Kp] rows, 1.0 at column K
  const uint8_t* W;         // image: per 128-column chunk [Kp/8][Nc][8] bf16
  void* Y;                  // hidden: bf16 [B, Np] ; last: unused
  float* soft;              // last layer: fp32 [B, N] or null
  void* hard;               // last layer: class index or null
  int hard_bytes;
  int64_t B;
  int Kp, N, Np;            // padded K, true N, padded N (hidden: next layer's Kp; last: round_up(N, 16))
  int act;
  int last;
  int* abort_flag;
  unsigned long long* debug;   // experiments: issuer wait-time breakdown of CTA 0 (null = off)
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == NVD_ACT_TANH) return ptx::tanh_approx(v);
  if (act == NVD_ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}

// CLUSTER CTAs (1, 2 or 4) share every weight stage: each loads 1/CLUSTER of it and multicasts that part to all, so
// the L2 -> SM weight traffic (the bound of this kernel: the whole layer image per 128-row tile) drops by CLUSTER.
// A stage is reusable once the MMAs of ALL CTAs have read it: every issuer commits to every CTA's empty barrier.
template <int CLUSTER>
__global__ void __launch_bounds__(kThreads, 1) mlp_wide_layer_kernel(const WideParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  constexpr int A_OFF = kStages * kStageBytes;
  constexpr int BAR_OFF = A_OFF + kAOperandBytes;
  const uint32_t bar_full = sbase + BAR_OFF;                 // [kStages]
  const uint32_t bar_empty = sbase + BAR_OFF + 8 * kStages;  // [kStages]
  const uint32_t bar_dfull = sbase + BAR_OFF + 16 * kStages;        // [2]
  const uint32_t bar_dfree = sbase + BAR_OFF + 16 * kStages + 16;   // [2]
  const uint32_t bar_aready = sbase + BAR_OFF + 16 * kStages + 32;
  const uint32_t bar_xfull = sbase + BAR_OFF + 16 * kStages + 40;    // next tile's rows have landed in the staging buffer
  const uint32_t bar_xempty = sbase + BAR_OFF + 16 * kStages + 48;   // ... have been moved to tensor memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 16 * kStages + 56);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { ptx::mbar_init(bar_full + 8 * s, 1); ptx::mbar_init(bar_empty + 8 * s, CLUSTER); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(bar_dfull + 8 * b, 1); ptx::mbar_init(bar_dfree + 8 * b, 4); }
    ptx::mbar_init(bar_aready, 8);
    ptx::mbar_init(bar_xfull, 1);
    ptx::mbar_init(bar_xempty, 8);
    ptx::mbar_fence_init();
  }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) ptx::cluster_sync();     // barriers initialised cluster-wide before any multicast / remote arrival
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xFFFFFFFFu, *tmem_slot, 0);
  const uint32_t crank = CLUSTER > 1 ? ptx::cluster_ctarank() : 0u;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CLUSTER) - 1u);

  // every CTA of a cluster must walk the same number of tiles (they consume one shared weight stream): tiles past
  // the end are processed as all-masked rows
  const int64_t real_tiles = (p.B + 127) >> 7;
  const int64_t num_tiles = CLUSTER > 1 ? (real_tiles + gridDim.x - 1) / gridDim.x * gridDim.x : real_tiles;
  const int n_chunks = (p.Np + 127) >> 7;
  const int k_stages = p.Kp / kStageK;
  auto chunk_cols = [&](int c) { return min(128, p.Np - 128 * c); };

  if (warp == 8) {
    // ---- weight producer: the whole image once per tile, stage by stage ----------------------------------
    if (ptx::elect_one()) {
      uint32_t it = 0, xt = 0;
      const uint32_t x_bytes = 128u * static_cast<uint32_t>(p.Kp) * 2u;
      auto stage_rows = [&](int64_t t) {     // tile-canonical activations: one contiguous block per tile
#ifdef NVD_WIDE_DIRECT_ROWS
        if (t < real_tiles)
          for (uint32_t o = 0; o < x_bytes; o += 16384)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const uint8_t*>(p.X) + static_cast<size_t>(t) * x_bytes + o), "r"(16384u) : "memory");
        return;
#endif
        if (t >= real_tiles) { ptx::mbar_arrive(bar_xfull); return; }     // cluster padding tile: nothing to fetch
        ptx::mbar_arrive_expect_tx(bar_xfull, x_bytes);
        for (uint32_t o = 0; o < x_bytes; o += 32768)
          ptx::bulk_g2s(sbase + A_OFF + o, reinterpret_cast<const uint8_t*>(p.X) + static_cast<size_t>(t) * x_bytes + o, 32768, bar_xfull);
      };
      if (static_cast<int64_t>(blockIdx.x) < num_tiles) stage_rows(blockIdx.x);
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++xt) {
        // the staging buffer is free once the row warps have moved this tile into tensor memory: fetch the next
        // tile's rows now, a whole tile of MMAs ahead of their use
        if (tile + gridDim.x < num_tiles) {
#ifndef NVD_WIDE_DIRECT_ROWS
          if (!ptx::mbar_wait_bounded(bar_xempty, xt & 1, kMaxWaitMs)) { atomicExch(p.abort_flag, 35); goto done; }
#endif
          stage_rows(tile + gridDim.x);
        }
        size_t off = 0;
        for (int c = 0; c < n_chunks; ++c) {
          const uint32_t bytes = static_cast<uint32_t>(chunk_cols(c)) * kStageK * 2;
          for (int ks = 0; ks < k_stages; ++ks, ++it, off += bytes) {
            const int s = it % kStages;
            if (!ptx::mbar_wait_bounded(bar_empty + 8 * s, ((it / kStages) & 1) ^ 1, kMaxWaitMs)) { atomicExch(p.abort_flag, 30); goto done; }
            ptx::mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
            if (CLUSTER == 1) {
              ptx::bulk_g2s(sbase + s * kStageBytes, p.W + off, bytes, bar_full + 8 * s);
            } else {
              const uint32_t part = bytes / CLUSTER;      // Nc * 128 B / CLUSTER: a multiple of 16 B for Nc % 16 == 0
              ptx::bulk_g2s_multicast(sbase + s * kStageBytes + crank * part, p.W + off + crank * part, part, bar_full + 8 * s, kMask);
            }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ---- MMA issuer -------------------------------------------------------------------------------------
    if (ptx::elect_one()) {
      uint32_t it = 0, tile_it = 0, chunk_it = 0;
      const bool dbgv = p.debug != nullptr;
      const bool dbg_a = (NVD_WIDE_TIMING_ALWAYS & 1) || dbgv, dbg_d = (NVD_WIDE_TIMING_ALWAYS & 2) || dbgv, dbg_f = (NVD_WIDE_TIMING_ALWAYS & 4) || dbgv, dbg_i = (NVD_WIDE_TIMING_ALWAYS & 8) || dbgv;
      long long t_a = 0, t_d = 0, t_f = 0, t_i = 0, t0 = clock64(), tt = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
        if (dbg_a) tt = clock64();
        if (!ptx::mbar_wait_bounded(bar_aready, tile_it & 1, kMaxWaitMs)) { atomicExch(p.abort_flag, 31); goto done; }
        if (dbg_a) t_a += clock64() - tt;
        ptx::tc_fence_after();
        for (int c = 0; c < n_chunks; ++c, ++chunk_it) {
          const int buf = chunk_it & 1;
          const int nc = chunk_cols(c);
          if (dbg_d) tt = clock64();
          if (!ptx::mbar_wait_bounded(bar_dfree + 8 * buf, ((chunk_it >> 1) & 1) ^ 1, kMaxWaitMs)) { atomicExch(p.abort_flag, 32); goto done; }
          if (dbg_d) t_d += clock64() - tt;
          ptx::tc_fence_after();
          const uint32_t idesc = ptx::make_idesc_bf16_f32(128, nc);
          for (int ks = 0; ks < k_stages; ++ks, ++it) {
            const int s = it % kStages;
            if (dbg_f) tt = clock64();
            if (!ptx::mbar_wait_bounded(bar_full + 8 * s, (it / kStages) & 1, kMaxWaitMs)) { atomicExch(p.abort_flag, 33); goto done; }
            if (dbg_f) t_f += clock64() - tt;
            if (dbg_i) tt = clock64();
            ptx::tc_fence_after();
#pragma unroll
            for (int j = 0; j < kStageK / 16; ++j)
#ifdef NVD_WIDE_SS
              ptx::umma_f16_ss(tmem_base + 128 * buf, ptx::make_smem_desc(sbase + A_OFF + (ks * (kStageK / 16) + j) * 2 * 2048, 2048, 128),
                               ptx::make_smem_desc(sbase + s * kStageBytes + j * 2 * (nc * 16), nc * 16, 128), idesc,
                               (ks | j) != 0);
#else
              ptx::umma_f16_ts(tmem_base + 128 * buf, tmem_base + COL_A + 8 * (ks * (kStageK / 16) + j),
                               ptx::make_smem_desc16(((sbase & 0x3FFFFu) >> 4) + s * (kStageBytes >> 4), j * 2 * (nc * 16), nc * 16, 128), idesc,
                               (ks | j) != 0);
#endif
            if (CLUSTER == 1) ptx::umma_commit(bar_empty + 8 * s);          // stage free once these MMAs have read it
            else ptx::umma_commit_multicast(bar_empty + 8 * s, kMask);       // ... in every CTA of the cluster
            if (dbg_i) t_i += clock64() - tt;
          }
          ptx::umma_commit(bar_dfull + 8 * buf);           // accumulator complete
        }
      }
      if (p.debug != nullptr && blockIdx.x == 0) {
        p.debug[0] = clock64() - t0; p.debug[1] = t_a; p.debug[2] = t_d; p.debug[3] = t_f; p.debug[4] = t_i; p.debug[5] = tile_it;
      }
    }
  } else {
    // ---- rows: load X into tensor memory, epilogues --------------------------------------------------------
    // two sets of four warps; set s owns accumulator buffer s (so the epilogues of consecutive chunks run side by
    // side: one warp per scheduler cannot hide the MUFU / tcgen05.ld latencies of a 128-column epilogue) and loads
    // every second 256-element batch of the rows
    const int set = warp >> 2;
    const int r = (warp & 3) * 32 + (tid & 31);            // row of the tile = TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t chunk_it = 0;
    // rows of a tile: staging buffer (filled by the producer's bulk copy a tile ahead) -> registers -> tensor memory.
    // Tile-canonical layout: the 16-byte K chunk kc of row r sits at kc * 2048 + r * 16, so a warp reads 512
    // contiguous bytes per request (no bank conflicts) and the SS-mode experiment can use the buffer as it is.
    uint32_t rows_it = 0;
    auto load_rows = [&](int64_t tile_to_load) {
#ifdef NVD_WIDE_DIRECT_ROWS
      {
        const uint4* src = reinterpret_cast<const uint4*>(p.X) + static_cast<size_t>(tile_to_load) * (p.Kp / 8) * 128 + r;
        const bool have = tile_to_load < real_tiles;
        for (int b = set; b < p.Kp / 256; b += 2) {
          uint4 v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = have ? __ldg(src + (32 * b + i) * 128) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint32_t w[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) { w[4 * q] = v[4 * c + q].x; w[4 * q + 1] = v[4 * c + q].y; w[4 * q + 2] = v[4 * c + q].z; w[4 * q + 3] = v[4 * c + q].w; }
            ptx::tmem_st16(tmem_base + lane_base + COL_A + 128 * b + 16 * c, w);
          }
        }
        ptx::tmem_wait_st();
        ptx::tc_fence_before();
        __syncwarp();
        if ((tid & 31) == 0) ptx::mbar_arrive(bar_aready);
        return;
      }
#endif
      if (!ptx::mbar_wait_bounded(bar_xfull, rows_it & 1, kMaxWaitMs)) { atomicExch(p.abort_flag, 36); return; }
      ++rows_it;
#ifndef NVD_WIDE_SS
      for (int b = set; b < p.Kp / 256; b += 2) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t w[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 v = *reinterpret_cast<const uint4*>(smem + A_OFF + (32 * b + 4 * c + q) * 2048 + r * 16);
            w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
          }
          ptx::tmem_st16(tmem_base + lane_base + COL_A + 128 * b + 16 * c, w);
        }
      }
      ptx::tmem_wait_st();
#endif
      ptx::tc_fence_before();
      __syncwarp();
      if ((tid & 31) == 0) {
#ifndef NVD_WIDE_SS
        ptx::mbar_arrive(bar_xempty);     // SS mode: the MMAs read the staging buffer; it is released with the tile
#endif
        ptx::mbar_arrive(bar_aready);
      }
    };
    auto epilogue = [&](int64_t tile, int c, int buf) {
      const int64_t row = tile * 128 + r;
      const int nc = chunk_cols(c);
      const uint32_t d = tmem_base + lane_base + 128 * buf;
      if (!p.last) {
        // next layer's rows in the same tile-canonical layout: 16 B per (row, 8 columns), 512 contiguous bytes per warp
        __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.Y) + (static_cast<size_t>(tile) * (p.Np / 8) + 16 * c) * 1024 + r * 8;
        for (int c0 = 0; c0 < nc; c0 += 32) {
          uint32_t v[32];
          ptx::tmem_ld32(d + c0, v);
          ptx::tmem_wait_ld();
          if (tile < real_tiles) {     // pad rows of the last tile are written too (the buffers hold whole tiles)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (c0 + 8 * q < nc) {
                uint4 o;
                o.x = pack_bf16x2(apply_act(__uint_as_float(v[8 * q]), p.act), apply_act(__uint_as_float(v[8 * q + 1]), p.act));
                o.y = pack_bf16x2(apply_act(__uint_as_float(v[8 * q + 2]), p.act), apply_act(__uint_as_float(v[8 * q + 3]), p.act));
                o.z = pack_bf16x2(apply_act(__uint_as_float(v[8 * q + 4]), p.act), apply_act(__uint_as_float(v[8 * q + 5]), p.act));
                o.w = pack_bf16x2(apply_act(__uint_as_float(v[8 * q + 6]), p.act), apply_act(__uint_as_float(v[8 * q + 7]), p.act));
                *reinterpret_cast<uint4*>(out + ((c0 >> 3) + q) * 1024) = o;
              }
            }
          }
        }
      } else {
        // output layer (N <= 128 classes in one chunk): argmax on the pre-activation (monotone activations), soft outputs
        float best = -3.0e38f;
        int bi = 0;
        for (int c0 = 0; c0 < nc; c0 += 16) {
          uint32_t v[16];
          ptx::tmem_ld16(d + c0, v);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = 128 * c + c0 + j;
            const float z = __uint_as_float(v[j]);
            if (col < p.N) {
              if (z > best) { best = z; bi = col; }
              if (p.soft != nullptr && row < p.B) p.soft[row * p.N + col] = (p.act == NVD_ACT_TANH) ? tanhf(z) : apply_act(z, p.act);
            }
          }
        }
        if (p.hard != nullptr && row < p.B) {
          if (p.act == NVD_ACT_RELU && best <= 0.0f) bi = 0;    // relu ties at (0, ..., 0) -> class 0 (first index)
          if (p.hard_bytes == 1) static_cast<uint8_t*>(p.hard)[row] = static_cast<uint8_t>(bi);
          else if (p.hard_bytes == 2) static_cast<uint16_t*>(p.hard)[row] = static_cast<uint16_t>(bi);
          else static_cast<int32_t*>(p.hard)[row] = bi;
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if ((tid & 31) == 0) ptx::mbar_arrive(bar_dfree + 8 * buf);
    };

    int64_t tile = blockIdx.x;
    if (tile < num_tiles) load_rows(tile);
    for (; tile < num_tiles; tile += gridDim.x) {
      for (int c = 0; c < n_chunks; ++c, ++chunk_it) {
        const int buf = chunk_it & 1;
        const bool mine = buf == set, last_chunk = c == n_chunks - 1;
        if (!mine && !last_chunk) continue;
        if (!ptx::mbar_wait_bounded(bar_dfull + 8 * buf, (chunk_it >> 1) & 1, kMaxWaitMs)) { atomicExch(p.abort_flag, 34); goto done; }
        ptx::tc_fence_after();
        // the last accumulator of the tile is complete => every MMA of the tile has read its A operand: bring the
        // next tile's rows in first, so that its MMAs run under this epilogue
#ifdef NVD_WIDE_SS
        if (c == n_chunks - 1) { __syncwarp(); if ((tid & 31) == 0) ptx::mbar_arrive(bar_xempty); }
#endif
        if (c == n_chunks - 1 && tile + gridDim.x < num_tiles) load_rows(tile + gridDim.x);
        if (mine) epilogue(tile, c, buf);
      }
    }
  }
done:
  ptx::tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) ptx::cluster_sync();     // nobody leaves while a peer may still multicast into its shared memory
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 512);
}

// rows of the first layer: dense ([B, K] with row stride K) or overlapping windows of one stream (row stride =
// window stride, e.g. 2), converted to bf16, 1.0 at column K, zero padding up to Kp
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ src, int64_t src_len, int64_t row_stride, int K, int Kp, int64_t B,
                                   __nv_bfloat16* __restrict__ X) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;   // one thread = 8 output columns
  const int per_row = Kp / 8;
  const int64_t row = i / per_row;
  if (row >= ((B + 127) & ~int64_t(127))) return;      // whole tiles: the pad rows of the last tile are zero-filled
  const int c0 = static_cast<int>(i - row * per_row) * 8;
  __align__(16) __nv_bfloat16 v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = c0 + j;
    float f = 0.0f;
    if (row >= B) f = 0.0f;
    else if (col < K) { const int64_t s = row * row_stride + col; f = s < src_len ? static_cast<float>(src[s]) : 0.0f; }
    else if (col == K) f = 1.0f;
    v[j] = __float2bfloat16_rn(f);
  }
  // tile-canonical: [tile][Kp / 8][128 rows][8]
  *reinterpret_cast<uint4*>(X + ((row >> 7) * (Kp / 8) + (c0 >> 3)) * 1024 + (row & 127) * 8) = *reinterpret_cast<const uint4*>(v);
}

uint16_t bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

bool wide_tc_supported(const MlpModel& m) {
  for (int l = 0; l <= m.n_layers; ++l)
    if (m.dims[l] + 1 > kMaxKp) return false;
  if (m.dims[m.n_layers] > 128) return false;   // output layer: one accumulator chunk
  return true;
}

// image of layer l: Kp = round_up(K + 1, 256) rows (row K = bias), Np columns (hidden: next layer's Kp, column N is
// the next layer's bias column: its bias-row entry makes act(.) = 1; last: round_up(N, 16)); chunked by 128 columns,
// each chunk [Kp/8][Nc][8] bf16 (UMMA canonical K-major, no swizzle)
int wide_tc_build_image(MlpModel& m, const float* const* W_host, const float* const* b_host) {
  for (int l = 0; l < m.n_layers; ++l) {
    const int K = m.dims[l], N = m.dims[l + 1];
    const int Kp = round_up(K + 1, 256);
    const bool last = l == m.n_layers - 1;
    const int Np = last ? round_up(N, 16) : round_up(N + 1, 256);
    std::vector<uint16_t> img(static_cast<size_t>(Kp) * Np, 0);
    size_t chunk_off = 0;
    for (int c = 0; c * 128 < Np; ++c) {
      const int nc = std::min(128, Np - 128 * c);
      for (int nn = 0; nn < nc; ++nn) {
        const int n = 128 * c + nn;
        for (int k = 0; k < Kp; ++k) {
          float v = 0.0f;
          if (n < N) {
            if (k < K) v = W_host[l][static_cast<size_t>(k) * N + n];
            else if (k == K) v = b_host[l][n];
          } else if (!last && n == N && k == K) {
            v = (m.acts[l] == NVD_ACT_TANH) ? 20.0f : 1.0f;     // the next layer's constant-one column
          }
          img[chunk_off + static_cast<size_t>(k >> 3) * (nc * 8) + nn * 8 + (k & 7)] = bf16_rn(v);
        }
      }
      chunk_off += static_cast<size_t>(Kp) * nc;
    }
    NVD_CUDA(cudaMalloc(&m.wide_image[l], img.size() * 2));
    NVD_CUDA(cudaMemcpy(m.wide_image[l], img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  }
  m.wide = true;
  return NVD_OK;
}

size_t wide_tc_scratch_bytes(const MlpModel& m, int64_t rows) {
  int widest = 0;
  for (int l = 0; l < m.n_layers; ++l) widest = std::max(widest, round_up(m.dims[l] + 1, 256));
  return 2 * static_cast<size_t>((rows + 127) & ~int64_t(127)) * widest * 2 + 512;
}

// rows: `B` rows taken from `src` (dtype NVD_F32 / NVD_BF16 / NVD_I32_TRUNC) at element stride `row_stride`
// (dims[0] for dense rows), `src_len` elements long.  scratch: wide_tc_scratch_bytes(m, B) bytes.
int launch_mlp_wide_tc(nvd_ctx* ctx, const MlpModel& m, const void* src, int src_dtype, int64_t src_len, int64_t row_stride,
                       int64_t B, void* scratch, float* soft_out, void* hard_out, int hard_bytes) {
  if (B == 0) return NVD_OK;
  int widest = 0;
  for (int l = 0; l < m.n_layers; ++l) widest = std::max(widest, round_up(m.dims[l] + 1, 256));
  __nv_bfloat16* ping = static_cast<__nv_bfloat16*>(scratch);
  __nv_bfloat16* pong = ping + static_cast<size_t>((B + 127) & ~int64_t(127)) * widest;
  const int K0 = m.dims[0], Kp0 = round_up(K0 + 1, 256);
  {
    const int64_t threads = ((B + 127) & ~int64_t(127)) * (Kp0 / 8);
    const unsigned grid = static_cast<unsigned>((threads + 255) / 256);
    if (src_dtype == NVD_F32) gather_rows_kernel<float><<<grid, 256, 0, ctx->stream>>>(static_cast<const float*>(src), src_len, row_stride, K0, Kp0, B, ping);
    else if (src_dtype == NVD_BF16) gather_rows_kernel<__nv_bfloat16><<<grid, 256, 0, ctx->stream>>>(static_cast<const __nv_bfloat16*>(src), src_len, row_stride, K0, Kp0, B, ping);
    else gather_rows_kernel<int32_t><<<grid, 256, 0, ctx->stream>>>(static_cast<const int32_t*>(src), src_len, row_stride, K0, Kp0, B, ping);
    NVD_CHECK_LAUNCH();
  }
  const int smem = kStages * kStageBytes + kAOperandBytes + 16 * kStages + 128;   // 229 856 B   // stages, full/empty barriers, accumulator / row barriers, TMEM slot
  const int64_t tiles = (B + 127) / 128;
  // clusters of 4 (or 2) CTAs share the weight stream when there are enough tiles to fill them
  int cluster = 1;
#ifndef NVD_WIDE_CLUSTER
#define NVD_WIDE_CLUSTER 2
#endif
  int want = NVD_WIDE_CLUSTER;
  if (const char* e = getenv("NVD_WIDE_CLUSTER")) want = atoi(e);     // experiments
  if (want >= 4 && ctx->sm_count % 4 == 0 && tiles >= ctx->sm_count) cluster = 4;
  else if (want >= 2 && ctx->sm_count % 2 == 0 && tiles >= ctx->sm_count) cluster = 2;
  auto kern = cluster == 4 ? mlp_wide_layer_kernel<4> : cluster == 2 ? mlp_wide_layer_kernel<2> : mlp_wide_layer_kernel<1>;
  NVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int grid = static_cast<int>(std::min<int64_t>(ctx->sm_count, tiles));
  if (cluster > 1) grid = ctx->sm_count;
  for (int l = 0; l < m.n_layers; ++l) {
    WideParams p{};
    const bool last = l == m.n_layers - 1;
    p.X = ping;
    p.W = static_cast<const uint8_t*>(m.wide_image[l]);
    p.Y = pong;
    p.soft = last ? soft_out : nullptr;
    p.hard = last ? hard_out : nullptr;
    p.hard_bytes = hard_bytes;
    p.B = B;
    p.Kp = round_up(m.dims[l] + 1, 256);
    p.N = m.dims[l + 1];
    p.Np = last ? round_up(p.N, 16) : round_up(p.N + 1, 256);
    p.act = m.acts[l];
    p.last = last ? 1 : 0;
    p.abort_flag = ctx->abort_flag;
    static unsigned long long* dbg = nullptr;
    if (getenv("NVD_WIDE_DEBUG")) {
      if (!dbg) cudaMalloc(&dbg, 64);
      p.debug = dbg;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    NVD_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    if (p.debug) {
      unsigned long long h[6];
      cudaMemcpyAsync(h, dbg, 48, cudaMemcpyDeviceToHost, ctx->stream);
      cudaStreamSynchronize(ctx->stream);
      fprintf(stderr, "wide layer %d: rows %lld  issuer clk total %llu  wait a_ready %llu  wait d_free %llu  wait full %llu  issue %llu  tiles %llu\n",
              l, static_cast<long long>(B), h[0], h[1], h[2], h[3], h[4], h[5]);
    }
    std::swap(ping, pong);
  }
  return NVD_OK;
}

}  // namespace nvd
