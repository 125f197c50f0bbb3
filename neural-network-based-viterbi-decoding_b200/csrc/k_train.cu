// k_train.cu — the training step of the feed-forward decoders on the GPU (the step BEFORE the hot
// path).  Replaces what the reference runs inside a TensorFlow-1.x session:
//   forward            multilayer_perceptron / build_network         src/fnn_viterbi.py:495-508, src/fnn_viterbi_bitwise.py:388-397
//   cost               mean softmax cross-entropy on the (activated) outputs + beta * sum_l l2_loss(W_l), l2_loss = sum(w^2)/2
//                      src/fnn_viterbi.py:549-559, src/fnn_viterbi_mixed_snr.py:722-739;
//                      mean sigmoid cross-entropy for the per-bit predictor, src/fnn_viterbi_bitwise.py:439-448
//   optimizer          tf.train.AdamOptimizer(lr).minimize(cost)      src/fnn_viterbi.py:562-564 (beta1 0.9, beta2 0.999, eps 1e-8;
//                      TensorFlow's form: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t), w -= lr_t m / (sqrt(v) + eps))
// Everything is fp32 on the FMA pipe: the nets are tiny (16 480 to 851 000 parameters, batches of 256) and training is
// off the Monte-Carlo path, so the kernels are plain shared-memory tiled GEMMs (three operand orders) and elementwise
// passes; what matters is that the gradients are right (tests/test_gpu_training.py holds them to the float64 oracle).
// The parameters live in the model's device arrays (MlpModel::W / b) and are updated in place; nvd_train_end writes
// them back to the host and rebuilds the model's tensor-core images from the trained values.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace nvd {

namespace {

constexpr int TS = 32;     // tile of the training GEMMs

// C[M, N] = op(A) op(B)  (+ C when ACC), row-major with leading dimensions; TA: A is stored [K, M], TB: B is stored [N, K]
template <bool TA, bool TB, bool ACC>
__global__ void __launch_bounds__(256) gemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ Cm,
                                                   int M, int N, int K, int lda, int ldb, int ldc) {
  __shared__ float As[TS][TS + 1];
  __shared__ float Bs[TS][TS + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8 threads, 4 rows each
  const int row0 = blockIdx.y * TS, col0 = blockIdx.x * TS;
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (int k0 = 0; k0 < K; k0 += TS) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i;
      const int gm = row0 + r, gk = k0 + tx;                   // As[r][tx] = op(A)[row0 + r][k0 + tx]
      As[r][tx] = (gm < M && gk < K) ? (TA ? A[static_cast<size_t>(gk) * lda + gm] : A[static_cast<size_t>(gm) * lda + gk]) : 0.0f;
      const int gkb = k0 + r, gn = col0 + tx;                  // Bs[r][tx] = op(B)[k0 + r][col0 + tx]
      Bs[r][tx] = (gkb < K && gn < N) ? (TB ? Bm[static_cast<size_t>(gn) * ldb + gkb] : Bm[static_cast<size_t>(gkb) * ldb + gn]) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TS; ++kk) {
      const float b = Bs[kk][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(As[ty + 8 * i][kk], b, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = row0 + ty + 8 * i, gn = col0 + tx;
    if (gm < M && gn < N) {
      float* c = Cm + static_cast<size_t>(gm) * ldc + gn;
      *c = ACC ? *c + acc[i] : acc[i];
    }
  }
}

template <bool TA, bool TB, bool ACC>
int gemm(nvd_ctx* ctx, const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc) {
  dim3 grid((N + TS - 1) / TS, (M + TS - 1) / TS);
  gemm_kernel<TA, TB, ACC><<<grid, 256, 0, ctx->stream>>>(A, B, C, M, N, K, lda, ldb, ldc);
  NVD_CHECK_LAUNCH();
  return NVD_OK;
}

__global__ void bias_act_kernel(float* __restrict__ z, const float* __restrict__ b, int64_t n, int N, int act) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float v = z[i] + b[i % N];
  if (act == NVD_ACT_TANH) v = tanhf(v);
  else if (act == NVD_ACT_RELU) v = fmaxf(v, 0.0f);
  z[i] = v;
}

// d(pre-activation) = d(activation) * act'(.) expressed through the stored activation value
__global__ void act_backward_kernel(float* __restrict__ d, const float* __restrict__ a, int64_t n, int act) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float av = a[i];
  if (act == NVD_ACT_TANH) d[i] *= 1.0f - av * av;
  else if (act == NVD_ACT_RELU) d[i] = av > 0.0f ? d[i] : 0.0f;
}

// softmax cross-entropy per row on the network's outputs used as logits: cost += -sum_c y log softmax(out);
// d_out = (softmax(out) - y) / B.  One warp per row.
__global__ void __launch_bounds__(256) softmax_ce_kernel(const float* __restrict__ out, const float* __restrict__ y, float* __restrict__ d_out,
                                                         int64_t B, int N, float inv_B, float* __restrict__ cost) {
  const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* o = out + row * N;
  const float* t = y + row * N;
  float mx = -INFINITY;
  for (int c = lane; c < N; c += 32) mx = fmaxf(mx, o[c]);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, s));
  float se = 0.0f, sy = 0.0f, syo = 0.0f;
  for (int c = lane; c < N; c += 32) { se += expf(o[c] - mx); sy += t[c]; syo += t[c] * (o[c] - mx); }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    se += __shfl_xor_sync(0xFFFFFFFFu, se, s);
    sy += __shfl_xor_sync(0xFFFFFFFFu, sy, s);
    syo += __shfl_xor_sync(0xFFFFFFFFu, syo, s);
  }
  const float lse = logf(se);
  for (int c = lane; c < N; c += 32) d_out[row * N + c] = (expf(o[c] - mx - lse) * sy - t[c]) * inv_B;
  if (lane == 0) atomicAdd(cost, (lse * sy - syo) * inv_B);
}

// sigmoid cross-entropy, mean over every element (tf.reduce_mean over [B, n_out]):
// cost = max(z, 0) - z y + log(1 + exp(-|z|));  d = (sigmoid(z) - y) / (B n_out)
__global__ void __launch_bounds__(256) sigmoid_ce_kernel(const float* __restrict__ out, const float* __restrict__ y, float* __restrict__ d_out,
                                                         int64_t n, float inv_n, float* __restrict__ cost) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  float c = 0.0f;
  if (i < n) {
    const float z = out[i], t = y[i];
    c = (fmaxf(z, 0.0f) - z * t + log1pf(expf(-fabsf(z)))) * inv_n;
    d_out[i] = (1.0f / (1.0f + expf(-z)) - t) * inv_n;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, s);
  if ((threadIdx.x & 31) == 0 && c != 0.0f) atomicAdd(cost, c);
}

__global__ void col_sum_kernel(const float* __restrict__ d, int64_t B, int N, float* __restrict__ db) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float s = 0.0f;
  for (int64_t r = 0; r < B; ++r) s += d[r * N + c];
  db[c] = s;
}

// L2 term of a layer: g += beta * w, cost += beta * sum(w^2) / 2
__global__ void __launch_bounds__(256) l2_kernel(const float* __restrict__ w, float* __restrict__ g, int64_t n, float beta, float* __restrict__ cost) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  float c = 0.0f;
  if (i < n) { g[i] += beta * w[i]; c = 0.5f * beta * w[i] * w[i]; }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, s);
  if ((threadIdx.x & 31) == 0 && c != 0.0f) atomicAdd(cost, c);
}

__global__ void adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                            float lr_t, float b1, float b2, float eps) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.0f - b1) * gi;
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  w[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

inline unsigned blocks(int64_t n) { return static_cast<unsigned>((n + 255) / 256); }

}  // namespace

}  // namespace nvd

using namespace nvd;

static void trainer_free(nvd_trainer& t) {
  for (int l = 0; l <= kMaxLayers; ++l) if (l > 0) { cudaFree(t.act[l]); t.act[l] = nullptr; }
  for (int i = 0; i < 2; ++i) { cudaFree(t.delta[i]); t.delta[i] = nullptr; }
  for (int l = 0; l < kMaxLayers; ++l) {
    cudaFree(t.gW[l]); cudaFree(t.gb[l]); cudaFree(t.mW[l]); cudaFree(t.vW[l]); cudaFree(t.mb[l]); cudaFree(t.vb[l]);
    t.gW[l] = t.gb[l] = t.mW[l] = t.vW[l] = t.mb[l] = t.vb[l] = nullptr;
  }
  cudaFree(t.cost);
  t.cost = nullptr;
  t.live = false;
}

// bf16 images of the tensor-core paths are derived from the fp32 parameters: rebuild them from the trained values
static int rebuild_images(nvd_ctx* ctx, MlpModel& m) {
  if (!m.tc_image && !m.wide && !m.ks_image) return NVD_OK;
  std::vector<std::vector<float>> W(m.n_layers), b(m.n_layers);
  const float* Wp[kMaxLayers];
  const float* bp[kMaxLayers];
  for (int l = 0; l < m.n_layers; ++l) {
    W[l].resize(static_cast<size_t>(m.dims[l]) * m.dims[l + 1]);
    b[l].resize(m.dims[l + 1]);
    NVD_CUDA(cudaMemcpy(W[l].data(), m.W[l], W[l].size() * 4, cudaMemcpyDeviceToHost));
    NVD_CUDA(cudaMemcpy(b[l].data(), m.b[l], b[l].size() * 4, cudaMemcpyDeviceToHost));
    Wp[l] = W[l].data();
    bp[l] = b[l].data();
  }
  cudaFree(m.tc_image); m.tc_image = nullptr;
  cudaFree(m.ks_image); m.ks_image = nullptr;
  cudaFree(m.ks_l2); m.ks_l2 = nullptr;
  for (int l = 0; l < kMaxLayers; ++l) { cudaFree(m.wide_image[l]); m.wide_image[l] = nullptr; }
  m.wide = false;
  if (tc_shape_supported(m)) return tc_build_image(ctx, m, Wp, bp);
  if (kstream_supported(m)) return kstream_build_image(m, Wp, bp);
  if (wide_tc_supported(m)) return wide_tc_build_image(m, Wp, bp);
  return NVD_OK;
}

extern "C" {

int nvd_train_begin(nvd_ctx* ctx, int model, const nvd_train_config* cfg, int64_t max_batch, int* trainer) {
  NVD_REQUIRE(ctx && cfg && trainer, "null argument");
  NVD_CUDA(cudaSetDevice(ctx->device));
  NVD_REQUIRE(cfg->struct_bytes == static_cast<int32_t>(sizeof(nvd_train_config)), "bad nvd_train_config (struct_bytes)");
  NVD_REQUIRE(model >= 0 && model < static_cast<int>(ctx->mlps.size()) && ctx->mlps[model].live, "bad MLP handle %d", model);
  NVD_REQUIRE(cfg->loss == NVD_LOSS_SOFTMAX_CE || cfg->loss == NVD_LOSS_SIGMOID_CE, "unknown loss %d", cfg->loss);
  NVD_REQUIRE(max_batch >= 1 && max_batch <= (int64_t(1) << 20), "max_batch out of range");
  NVD_REQUIRE(cfg->learning_rate > 0.0f && cfg->beta1 >= 0.0f && cfg->beta1 < 1.0f && cfg->beta2 >= 0.0f && cfg->beta2 < 1.0f &&
              cfg->epsilon > 0.0f, "bad optimizer constants");
  const MlpModel& m = ctx->mlps[model];
  nvd_trainer t;
  t.model = model;
  t.cfg = *cfg;
  t.max_batch = max_batch;
  int widest = 0;
  for (int l = 0; l < m.n_layers; ++l) {
    const size_t nw = static_cast<size_t>(m.dims[l]) * m.dims[l + 1], nb = m.dims[l + 1];
    widest = std::max(widest, std::max(m.dims[l], m.dims[l + 1]));
    cudaError_t e = cudaMalloc(&t.act[l + 1], static_cast<size_t>(max_batch) * m.dims[l + 1] * 4);
    if (e == cudaSuccess) e = cudaMalloc(&t.gW[l], nw * 4);
    if (e == cudaSuccess) e = cudaMalloc(&t.gb[l], nb * 4);
    if (e == cudaSuccess) e = cudaMalloc(&t.mW[l], nw * 4);
    if (e == cudaSuccess) e = cudaMalloc(&t.vW[l], nw * 4);
    if (e == cudaSuccess) e = cudaMalloc(&t.mb[l], nb * 4);
    if (e == cudaSuccess) e = cudaMalloc(&t.vb[l], nb * 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(t.mW[l], 0, nw * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(t.vW[l], 0, nw * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(t.mb[l], 0, nb * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(t.vb[l], 0, nb * 4, ctx->stream);
    if (e != cudaSuccess) { trainer_free(t); set_error("trainer allocation failed: %s", cudaGetErrorString(e)); return NVD_ERR_OOM; }
  }
  cudaError_t e = cudaMalloc(&t.delta[0], static_cast<size_t>(max_batch) * widest * 4);
  if (e == cudaSuccess) e = cudaMalloc(&t.delta[1], static_cast<size_t>(max_batch) * widest * 4);
  if (e == cudaSuccess) e = cudaMalloc(&t.cost, 4);
  if (e != cudaSuccess) { trainer_free(t); set_error("trainer allocation failed: %s", cudaGetErrorString(e)); return NVD_ERR_OOM; }
  t.live = true;
  ctx->trainers.push_back(t);
  *trainer = static_cast<int>(ctx->trainers.size()) - 1;
  return NVD_OK;
}

int nvd_train_step(nvd_ctx* ctx, int trainer, const float* x_dev, const float* y_dev, int64_t B, float* cost_host) {
  NVD_REQUIRE(ctx != nullptr, "null context");
  NVD_CUDA(cudaSetDevice(ctx->device));
  NVD_REQUIRE(trainer >= 0 && trainer < static_cast<int>(ctx->trainers.size()) && ctx->trainers[trainer].live, "bad trainer handle %d", trainer);
  nvd_trainer& t = ctx->trainers[trainer];
  NVD_REQUIRE(B >= 1 && B <= t.max_batch && x_dev && y_dev, "bad batch / null buffer");
  MlpModel& m = ctx->mlps[t.model];
  NVD_REQUIRE(m.live, "the trainer's model was freed");
  const int nl = m.n_layers;
  const int Bi = static_cast<int>(B);
  NVD_CUDA(cudaMemsetAsync(t.cost, 0, 4, ctx->stream));
  // ---- forward, every activation kept ----
  const float* a_prev = x_dev;
  for (int l = 0; l < nl; ++l) {
    const int K = m.dims[l], N = m.dims[l + 1];
    NVD_TRY((gemm<false, false, false>(ctx, a_prev, m.W[l], t.act[l + 1], Bi, N, K, K, N, N)));
    bias_act_kernel<<<blocks(B * N), 256, 0, ctx->stream>>>(t.act[l + 1], m.b[l], B * N, N, m.acts[l]);
    NVD_CHECK_LAUNCH();
    a_prev = t.act[l + 1];
  }
  // ---- loss on the activated outputs, d(out activation) ----
  const int n_out = m.dims[nl];
  float* d = t.delta[0];
  if (t.cfg.loss == NVD_LOSS_SOFTMAX_CE)
    softmax_ce_kernel<<<blocks(B * 32), 256, 0, ctx->stream>>>(t.act[nl], y_dev, d, B, n_out, 1.0f / static_cast<float>(B), t.cost);
  else
    sigmoid_ce_kernel<<<blocks(B * n_out), 256, 0, ctx->stream>>>(t.act[nl], y_dev, d, B * n_out, 1.0f / static_cast<float>(B * n_out), t.cost);
  NVD_CHECK_LAUNCH();
  // ---- backward ----
  for (int l = nl - 1; l >= 0; --l) {
    const int K = m.dims[l], N = m.dims[l + 1];
    act_backward_kernel<<<blocks(B * N), 256, 0, ctx->stream>>>(d, t.act[l + 1], B * N, m.acts[l]);
    NVD_CHECK_LAUNCH();
    const float* a_in = l ? t.act[l] : x_dev;
    NVD_TRY((gemm<true, false, false>(ctx, a_in, d, t.gW[l], K, N, Bi, K, N, N)));      // dW = A^T dZ
    col_sum_kernel<<<(N + 127) / 128, 128, 0, ctx->stream>>>(d, B, N, t.gb[l]);
    NVD_CHECK_LAUNCH();
    if (t.cfg.l2_layer_mask & (1u << l)) {
      l2_kernel<<<blocks(static_cast<int64_t>(K) * N), 256, 0, ctx->stream>>>(m.W[l], t.gW[l], static_cast<int64_t>(K) * N, t.cfg.beta, t.cost);
      NVD_CHECK_LAUNCH();
    }
    if (l > 0) {
      float* dn = d == t.delta[0] ? t.delta[1] : t.delta[0];
      NVD_TRY((gemm<false, true, false>(ctx, d, m.W[l], dn, Bi, K, N, N, N, K)));          // dA = dZ W^T
      d = dn;
    }
  }
  // ---- Adam (TensorFlow's form) ----
  ++t.step;
  const double b1t = std::pow(static_cast<double>(t.cfg.beta1), static_cast<double>(t.step));
  const double b2t = std::pow(static_cast<double>(t.cfg.beta2), static_cast<double>(t.step));
  const float lr_t = static_cast<float>(t.cfg.learning_rate * std::sqrt(1.0 - b2t) / (1.0 - b1t));
  for (int l = 0; l < nl; ++l) {
    const int64_t nw = static_cast<int64_t>(m.dims[l]) * m.dims[l + 1], nb = m.dims[l + 1];
    adam_kernel<<<blocks(nw), 256, 0, ctx->stream>>>(m.W[l], t.gW[l], t.mW[l], t.vW[l], nw, lr_t, t.cfg.beta1, t.cfg.beta2, t.cfg.epsilon);
    adam_kernel<<<blocks(nb), 256, 0, ctx->stream>>>(m.b[l], t.gb[l], t.mb[l], t.vb[l], nb, lr_t, t.cfg.beta1, t.cfg.beta2, t.cfg.epsilon);
    NVD_CHECK_LAUNCH();
  }
  if (cost_host) {
    NVD_CUDA(cudaMemcpyAsync(cost_host, t.cost, 4, cudaMemcpyDeviceToHost, ctx->stream));
    NVD_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return NVD_OK;
}

int nvd_train_end(nvd_ctx* ctx, int trainer, float* const* W_host, float* const* b_host) {
  NVD_REQUIRE(ctx != nullptr, "null context");
  NVD_CUDA(cudaSetDevice(ctx->device));
  NVD_REQUIRE(trainer >= 0 && trainer < static_cast<int>(ctx->trainers.size()) && ctx->trainers[trainer].live, "bad trainer handle %d", trainer);
  nvd_trainer& t = ctx->trainers[trainer];
  const MlpModel& m = ctx->mlps[t.model];
  if (W_host && b_host && m.live) {
    for (int l = 0; l < m.n_layers; ++l) {
      NVD_CUDA(cudaMemcpyAsync(W_host[l], m.W[l], static_cast<size_t>(m.dims[l]) * m.dims[l + 1] * 4, cudaMemcpyDeviceToHost, ctx->stream));
      NVD_CUDA(cudaMemcpyAsync(b_host[l], m.b[l], static_cast<size_t>(m.dims[l + 1]) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
  }
  NVD_CUDA(cudaStreamSynchronize(ctx->stream));
  int st = NVD_OK;
  if (ctx->mlps[t.model].live) st = rebuild_images(ctx, ctx->mlps[t.model]);     // the tensor-core images follow the new weights
  trainer_free(t);
  return st;
}

}  // extern "C"

namespace nvd {
void trainers_release(nvd_ctx* ctx) {
  for (auto& t : ctx->trainers)
    if (t.live) trainer_free(t);
  ctx->trainers.clear();
}
}  // namespace nvd
