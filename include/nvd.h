/* nvd.h — C ABI of libnvd.so: the B200-native Monte-Carlo decode path
 *   info bits -> convolutional encoder -> BPSK+AWGN -> neural decoder -> hard decision -> BER/FER
 * of enessinanparildi/Neural-Network-Based-Viterbi-Decoding.
 *
 * The reference is pure Python and has NO FFI / plugin layer: the path sits
 * behind module-level functions that exchange numpy arrays and whose heavy lifting is done by
 * Python loops and a TensorFlow-1.x session.  Each entry point below names the reference
 * function (file:line under /root/reference) whose work it replaces; the Python drop-in
 * modules (the .py files under neural-network-based-viterbi-decoding_b200/src) bind them through ctypes, and
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns NVD_OK (0) or a negative nvd_status; nvd_last_error() returns a
 *     thread-local message; no exception or abort crosses the boundary;
 *   - plain pointers and sizes only; `*_dev` / unsuffixed buffer arguments are DEVICE pointers
 *     (from nvd_alloc or any CUDA allocator of the same device), `*_host` are HOST pointers;
 *   - a context owns one device, one stream, its models and scratch; calls are asynchronous on
 *     that stream unless documented as synchronous; a context is not thread-safe, distinct
 *     contexts are;
 *   - there is NO CPU fallback: without a CUDA device nvd_ctx_create fails.
 *
 * Packed-bit wire format: frames are byte aligned; bit i (little-endian within the byte
 * stream) of a frame is the symbol at time i.  A frame of L info bits occupies ceil(L/8)
 * bytes, its 2L coded bits ceil(2L/8) bytes with coded[2i] = g_a(i), coded[2i+1] = g_b(i).
 * Decoder class indices follow the reference's MSB-first labels (generate_class_bits,
 * src/fnn_viterbi.py:283): class = bit-reverse_L(time-packed info word).
 */
#ifndef NVD_H_
#define NVD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NVD_VERSION 100 /* 0.1.0 */

typedef struct nvd_ctx nvd_ctx;

typedef enum {
  NVD_OK = 0,
  NVD_ERR_INVALID = -1,     /* bad argument */
  NVD_ERR_CUDA = -2,        /* CUDA runtime / launch failure (message has the CUDA error) */
  NVD_ERR_OOM = -3,
  NVD_ERR_UNSUPPORTED = -4, /* shape not covered by a compiled kernel */
  NVD_ERR_DEVICE_TIMEOUT = -5 /* a bounded on-device wait expired (kernel self-aborted) */
} nvd_status;

/* encoder selection — src/fnn_viterbi.py:122 (encode_57), :157 (encode_133171) */
enum { NVD_CODE_57_REF = 0, NVD_CODE_133171 = 1, NVD_CODE_57_CAUSAL = 2 };
/* channel output storage — float32 rows (src/fnn_viterbi.py:373,384), bf16 for the tensor-core
 * path, or the int32 truncation of src/fnn_viterbi_bitwise.py:309-317 / src/rnn_viterbi.py:201-208 */
enum { NVD_F32 = 0, NVD_BF16 = 1, NVD_I32_TRUNC = 2,
       NVD_BF16_TRUNC = 3 /* the truncated integer stored as bf16 (exact): the same quirk on the tensor-core path */ };
/* activations — tanh src/fnn_viterbi.py:496-508, relu src/fnn_viterbi_bitwise.py:388,397 */
enum { NVD_ACT_IDENTITY = 0, NVD_ACT_TANH = 1, NVD_ACT_RELU = 2 };
/* arithmetic of the decoder forward */
enum { NVD_PREC_FP32 = 0, NVD_PREC_BF16_TC = 1 };
/* Monte-Carlo pipeline organisation */
enum { NVD_MC_STAGED = 0, NVD_MC_FUSED = 1 };
/* nvd_count_errors flags */
enum { NVD_COUNT_PRED_IS_MSB_CLASS = 1 };

/* ---- library / context ------------------------------------------------------------------ */
int nvd_version(void);
const char* nvd_last_error(void);
int nvd_device_count(int* count);
/* replaces the implicit `tf.Session()` device/state (src/fnn_viterbi.py:570) */
int nvd_ctx_create(int device, uint64_t seed, nvd_ctx** out);
int nvd_ctx_destroy(nvd_ctx* ctx);
int nvd_set_seed(nvd_ctx* ctx, uint64_t seed);
int nvd_sync(nvd_ctx* ctx);                         /* synchronous; reports device-side aborts */
int nvd_device_info(nvd_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);

/* ---- memory and timing (the caller owns what it allocates) ------------------------------- */
int nvd_alloc(nvd_ctx* ctx, size_t bytes, void** dptr);
int nvd_free(nvd_ctx* ctx, void* dptr);
int nvd_host_alloc(nvd_ctx* ctx, size_t bytes, void** hptr); /* pinned */
int nvd_host_free(nvd_ctx* ctx, void* hptr);
int nvd_memcpy_h2d(nvd_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int nvd_memcpy_d2h(nvd_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int nvd_memset(nvd_ctx* ctx, void* dst_dev, int byte, size_t bytes);
int nvd_event_create(nvd_ctx* ctx, void** event);
int nvd_event_record(nvd_ctx* ctx, void* event);    /* on the context's stream */
int nvd_event_elapsed_ms(nvd_ctx* ctx, void* start, void* stop, float* ms); /* synchronises on stop */
int nvd_event_destroy(nvd_ctx* ctx, void* event);

/* ---- K0 source: class-balanced / random info words ---------------------------------------
 * replaces generate_class_bits rows repeated per class (src/fnn_viterbi.py:380-388) for L <= 16
 * (frame f carries class f mod 2^L) and np.random.randint (src/fnn_viterbi_bitwise.py:239-250)
 * for L > 16 (Philox stream 2).  info_packed: [B, ceil(L/8)] bytes. */
int nvd_source(nvd_ctx* ctx, int L, uint64_t frame0, int64_t B, uint8_t* info_packed);

/* ---- K1 encoder: replaces encode_57 / encode_133171 (src/fnn_viterbi.py:122-217) and, with the
 * 6 state bits prepended by the host, encode_133171_with_state (src/fnn_viterbi_bitwise.py:141-215).
 * info_packed [B, ceil(L/8)] -> coded_packed [B, ceil(2L/8)].  Zero state, unterminated. */
int nvd_encode(nvd_ctx* ctx, int code, const uint8_t* info_packed, int L, int64_t B,
               uint8_t* coded_packed);

/* ---- K2 channel: replaces modulate_awgn (src/fnn_viterbi.py:224-253) + the per-frame Python
 * loops that call it (:380-388).  y[b, j] = (2c-1) + sigma * N(0,1), noise = Philox4x32-10 keyed
 * by (seed; frame0+b, j/8, snr_idx) — eight symbols per Philox call (csrc/frame_gen.cuh).  y is [B, n] dense in out_dtype
 * (NVD_F32/BF16/I32_TRUNC). */
int nvd_channel(nvd_ctx* ctx, const uint8_t* coded_packed, int n, int64_t B, float sigma,
                int snr_idx, uint64_t frame0, int out_dtype, void* y);

/* ---- K3 feed-forward decoder: replaces initialize_weights + multilayer_perceptron + tf.argmax
 * (src/fnn_viterbi.py:413-510,606-610; src/fnn_viterbi_bitwise.py:326-399,504;
 *  src/fnn_viterbi_seq_predict.py:343-414).  W[l] is [dims[l], dims[l+1]] row-major (x@W+b),
 * host pointers; returns a model handle. */
int nvd_mlp_load(nvd_ctx* ctx, int n_layers, const int* dims, const int* acts,
                 const float* const* W_host, const float* const* b_host, int* model);
int nvd_mlp_free(nvd_ctx* ctx, int model);
/* y: [B, dims[0]] in y_dtype.  soft_out: [B, dims[n]] fp32 or NULL.  hard_out: argmax class
 * (first index on ties) as hard_bytes-wide (1, 2 or 4) little-endian integers, or NULL.
 * NVD_PREC_BF16_TC needs a tcgen05-covered shape (see nvd_mlp_tc_supported): the reference FNN 14-128-64-32-128
 * runs on the fused single-kernel path (y must be NVD_BF16); any other model whose layer widths are <= 511 with
 * <= 128 classes runs layer by layer on the wide-layer tensor-core kernel (y in any dtype). */
int nvd_mlp_forward(nvd_ctx* ctx, int model, const void* y, int y_dtype, int64_t B, int precision,
                    float* soft_out, void* hard_out, int hard_bytes);
int nvd_mlp_tc_supported(nvd_ctx* ctx, int model, int* supported);
/* device address of a loaded model's fp32 parameters: layer's weights [dims[l], dims[l+1]] (weight != 0) or biases
 * [dims[l+1]] — what a training step updates in place (nvd_train_step) and a caller may read with nvd_memcpy_d2h */
int nvd_mlp_param(nvd_ctx* ctx, int model, int layer, int weight, void** dptr);
/* Sliding-window decode (src/fnn_viterbi_seq_predict.py:293-336 get_single_instance + the forward of :343-414):
 * window i = stream[i * stride, i * stride + dims[0]) of one long symbol stream in device memory
 * (stream_dtype NVD_F32 / NVD_BF16 / NVD_I32_TRUNC), i < n_windows; soft_out [n_windows, n_out] fp32 or NULL,
 * hard_out class index per window or NULL.  NVD_PREC_BF16_TC needs the wide-layer tensor-core image (all layer
 * widths <= 511, <= 128 classes; nvd_mlp_tc_supported), NVD_PREC_FP32 accepts any model.  Asynchronous. */
int nvd_window_decode(nvd_ctx* ctx, int model, const void* stream, int stream_dtype, int64_t n_symbols,
                      int64_t stride, int64_t n_windows, int precision, float* soft_out, void* hard_out,
                      int hard_bytes);
/* How the tensor-core path evaluates the hidden-layer tanh (so a numerical model can follow it exactly):
 * of every 8 consecutive bf16 pairs (16 columns) of a layer, the first *poly_pairs_of_8 use
 *   tanh(x) ~ sign(x) * u * (c[0] + c[1] u + ... + c[5] u^5),  u = min(|x| * *inv_range, 1)   (fp32, FMA pipe)
 * and the rest use tanh.approx.f32 (MUFU).  Needs no device. */
int nvd_tc_tanh_model(int* poly_pairs_of_8, float* coef6, float* inv_range);
/* host-buffer convenience (synchronous): the `predictions_tensor.eval(feed_dict={x: test_set})`
 * call of src/fnn_viterbi.py:610 — chunked H2D, forward, D2H.  y_host fp32 [B, dims[0]]. */
int nvd_mlp_decode_host(nvd_ctx* ctx, int model, const float* y_host, int64_t B, int precision,
                        float* soft_host, int32_t* hard_host);

/* ---- K4 recurrent decoder: replaces build_seq2seq_model's forward (src/rnn_viterbi.py:278-305).
 * LSTMCell semantics: gates i,j,f,o, forget_bias 1; kernels [1+H, 4H]; linear [H]->1. */
int nvd_lstm_load(nvd_ctx* ctx, int H, const float* enc_W, const float* enc_b, const float* dec_W,
                  const float* dec_b, const float* lin_w, const float* lin_b, int* model);
int nvd_lstm_free(nvd_ctx* ctx, int model);
/* y [B, T_enc] in y_dtype; dec_in [B, T_dec] fp32 (teacher forcing, src/rnn_viterbi.py:362-371)
 * or NULL = autoregressive (previous hard decision fed back, first input EOS=1).
 * logits_out [B, T_dec] fp32 or NULL; hard_packed [B, ceil(T_dec/8)] (bit t = logit_t > 0) or NULL.
 * fast_math: 0 = fp32 with libm-accurate tanh/sigmoid (parity path), 1 = fp32 with MUFU approximations,
 * 2 = bf16 tensor-core recurrence (tcgen05; weights and h in bf16, fp32 accumulation, c in fp32; H = 16 or 20). */
int nvd_lstm_forward(nvd_ctx* ctx, int model, const void* y, int y_dtype, const float* dec_in,
                     int T_enc, int T_dec, int64_t B, int fast_math, float* logits_out,
                     uint8_t* hard_packed);

/* ---- K5 reduction: replaces calculate_error (src/fnn_viterbi.py:619-674) and calculate_ber /
 * calculate_hamming_distance_decimal (src/fnn_viterbi_mixed_snr.py:303-328,634-655).
 * pred, label: [B, stride_bytes] little-endian packed frames, L valid bits each.
 * counters_dev: uint64[4] = {bit_errors, frame_errors, frames, bits}, ACCUMULATED into.
 * flag NVD_COUNT_PRED_IS_MSB_CLASS: pred holds MSB-first class indices (stride <= 4), label
 * time-packed info words. */
int nvd_count_errors(nvd_ctx* ctx, const void* pred, const void* label, int stride_bytes, int L,
                     int64_t B, int flags, uint64_t* counters_dev);

/* ---- whole Monte-Carlo step ---------------------------------------------------------------
 * replaces the test half of train_test_dataset + the test-set forward + calculate_error
 * (src/fnn_viterbi.py:380-388,606-610,619-674) / get_dataset + create_model_graph eval branch
 * (src/fnn_viterbi_mixed_snr.py:452-460,806-845).  Frames [frame0, frame0+frames) at each of the
 * n_snr points; results are independent of how a frame range is split across calls or GPUs. */
typedef struct {
  int32_t struct_bytes;   /* = sizeof(nvd_mc_config) */
  int32_t model;          /* MLP handle */
  int32_t code;           /* NVD_CODE_* */
  int32_t L;              /* info bits per frame (<= 8: class decoders carry one info byte per frame) */
  int32_t precision;      /* NVD_PREC_* */
  int32_t mode;           /* NVD_MC_STAGED | NVD_MC_FUSED (fused needs BF16_TC + covered shape) */
  int32_t y_dtype;        /* staged: storage of the channel output (NVD_F32 / NVD_BF16) */
  int32_t reserved;
  uint64_t frame0;
  uint64_t frames;
} nvd_mc_config;
/* asynchronous; counters_dev: uint64 [n_snr][4], accumulated into. snr_idx[i] keys the noise. */
int nvd_mc_run(nvd_ctx* ctx, const nvd_mc_config* cfg, const float* snr_db_host,
               const int32_t* snr_idx_host, int n_snr, uint64_t* counters_dev);
/* synchronous host-buffer form (the end-to-end call): info words come from the host, hard
 * decisions and counters go back.  info_host [B, ceil(L/8)], pred_host [B] uint8 class indices
 * (may be NULL), counters_host uint64[4] (overwritten).  Pass pinned host buffers (nvd_host_alloc).
 * NVD_MC_STAGED runs the five staged kernels per chunk; NVD_MC_FUSED runs the fused tcgen05 kernel:
 * with pinned buffers ONE launch, the info words streaming in and the decisions streaming out in
 * slices while it runs (1 B per frame each way); otherwise (pageable buffers, NVD_HOST_STEP=sliced,
 * or under a profiler) one launch per slice with the copies of neighbouring slices on two copy
 * streams.  All forms give identical decisions and counters. */
int nvd_mc_run_host(nvd_ctx* ctx, const nvd_mc_config* cfg, float snr_db, int snr_idx,
                    const uint8_t* info_host, uint8_t* pred_host, uint64_t* counters_host);
/* per-stage device milliseconds of the most recent staged nvd_mc_run chunk sequence:
 * ms[5] = {source, encode, channel, decode, count} (sums over chunks), synchronises. */
int nvd_mc_stage_ms(nvd_ctx* ctx, float* ms5);
/* per-stage timing makes staged nvd_mc_run synchronise after every chunk; off by default. */
int nvd_mc_set_stage_timing(nvd_ctx* ctx, int enabled);

/* ---- Monte-Carlo step of the per-bit predictor (BASELINE.json configs[2]) ------------------------------------
 * replaces generate_test_data (src/fnn_viterbi_bitwise.py:290-323: random info bits, encode_133171_with_state from
 * `starting_state`, modulate_awgn into an int32 row), the test forward of build_network (:326-399) and the decision /
 * error count against info bit `bit_position` (:456-457, :504-509).  One decoded bit per frame: counters are
 * {bit errors, frame errors (same number), frames, bits (= frames)}.  The model is a [2(L + n_state_bits), N1 <= 64,
 * n_out <= 8] net (nvd_mlp_tc_supported).  NVD_MC_FUSED (BF16_TC only) generates the frames on chip; NVD_MC_STAGED
 * runs K0 (Philox info bits) -> prefix -> K1 -> K2 -> K3 -> K5 through HBM.  Both give identical counters. */
typedef struct {
  int32_t struct_bytes;   /* = sizeof(nvd_mc_bitwise_config) */
  int32_t model;
  int32_t code;           /* NVD_CODE_* */
  int32_t L;              /* info bits per frame (block_length, 512) */
  int32_t n_state_bits;   /* constraint_length (6): start-state bits prepended to the info bits; 0 = none */
  int32_t state;          /* starting_state, the number extract_state_bits takes (:114-138) */
  int32_t bit_position;   /* the info bit the net predicts */
  int32_t precision;      /* NVD_PREC_FP32 (staged only) | NVD_PREC_BF16_TC */
  int32_t mode;           /* NVD_MC_STAGED | NVD_MC_FUSED */
  int32_t trunc;          /* 1 = the reference data sets' int32 truncation of the channel output (:309-317), 0 = real-valued */
  uint64_t frame0;
  uint64_t frames;
} nvd_mc_bitwise_config;
int nvd_mc_run_bitwise(nvd_ctx* ctx, const nvd_mc_bitwise_config* cfg, const float* snr_db_host, const int32_t* snr_idx_host,
                       int n_snr, uint64_t* counters_dev);

/* ---- Monte-Carlo step of the recurrent decoder (BASELINE.json configs[3]) --------------------------------------
 * replaces generate_datasets' test half (src/rnn_viterbi.py:186-232: random info bits, encode_133171, modulate_awgn
 * into an int32 row), prepare_batch's teacher-forced decoder inputs (:328-377), the forward of build_seq2seq_model
 * (:278-305) and the BER over the first L decoder steps (:471-491), with the decision bit = logit > 0 (B-12).
 * K0 (Philox info bits) -> K1 -> K2 -> K4 (T_enc = T_dec = 2L) -> K5 on packed rows; counters as nvd_mc_run.
 * teacher_forced = 1 feeds [EOS, info, 0...] as the reference does even at test time; 0 feeds the previous decision
 * back.  fast_math as nvd_lstm_forward (2 = tcgen05 recurrence). */
typedef struct {
  int32_t struct_bytes;   /* = sizeof(nvd_mc_lstm_config) */
  int32_t model;          /* LSTM handle */
  int32_t code;
  int32_t L;              /* info bits per frame (100) */
  int32_t fast_math;      /* 0, 1, 2 */
  int32_t teacher_forced;
  int32_t trunc;          /* 1 = int32 truncation of the channel output (:216-222) */
  int32_t reserved;
  uint64_t frame0;
  uint64_t frames;
} nvd_mc_lstm_config;
int nvd_mc_run_lstm(nvd_ctx* ctx, const nvd_mc_lstm_config* cfg, const float* snr_db_host, const int32_t* snr_idx_host,
                    int n_snr, uint64_t* counters_dev);

/* K5 variants the two steps above use, also callable on their own (device buffers, counters accumulated into):
 * decision byte (0/1) per frame against bit `bit_position` of a packed info row (src/fnn_viterbi_bitwise.py:504-509);
 * packed decision rows against packed info rows with different strides, first L bits (src/rnn_viterbi.py:477-491). */
int nvd_count_bit_errors(nvd_ctx* ctx, const uint8_t* pred, const uint8_t* info_packed, int info_stride_bytes, int bit_position,
                         int64_t B, uint64_t* counters_dev);
int nvd_count_row_errors(nvd_ctx* ctx, const uint8_t* pred_packed, int pred_stride_bytes, const uint8_t* info_packed,
                         int info_stride_bytes, int L, int64_t B, uint64_t* counters_dev);
/* the six start-state bits in front of every packed frame (encode_133171_with_state, src/fnn_viterbi_bitwise.py:162-163):
 * out [B, ceil((L + n_state_bits) / 8)] */
int nvd_prepend_state(nvd_ctx* ctx, const uint8_t* info_packed, int L, int state, int n_state_bits, int64_t B, uint8_t* out);

/* ---- training step of the feed-forward decoders (the step before the path) -----------------------
 * replaces the TensorFlow-1.x session loops: forward, cost = mean softmax cross-entropy on the activated outputs (or mean
 * sigmoid cross-entropy, src/fnn_viterbi_bitwise.py:439-448) + beta * sum over the masked layers of sum(w^2)/2
 * (src/fnn_viterbi.py:549-559), backward, tf.train.AdamOptimizer (:562-564; TensorFlow's update
 * lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t), w -= lr_t m / (sqrt(v) + epsilon)).  fp32 throughout.  The model's device
 * parameters are updated in place by every step; nvd_train_end copies them to the host arrays (may be NULL), rebuilds the
 * model's tensor-core images from them and releases the trainer.  x_dev [B, dims[0]] fp32, y_dev [B, n_out] fp32 targets
 * (one-hot rows), B <= max_batch.  cost_host (may be NULL; non-NULL synchronises) receives the step's cost before the update. */
enum { NVD_LOSS_SOFTMAX_CE = 0, NVD_LOSS_SIGMOID_CE = 1 };
typedef struct {
  int32_t struct_bytes;     /* = sizeof(nvd_train_config) */
  int32_t loss;             /* NVD_LOSS_* */
  float learning_rate;
  float beta;               /* L2 weight (beta_regularization) */
  uint32_t l2_layer_mask;   /* bit l: layer l's weights are in the L2 term */
  float beta1, beta2, epsilon;   /* Adam: 0.9, 0.999, 1e-8 are TensorFlow's defaults */
} nvd_train_config;
int nvd_train_begin(nvd_ctx* ctx, int model, const nvd_train_config* cfg, int64_t max_batch, int* trainer);
int nvd_train_step(nvd_ctx* ctx, int trainer, const float* x_dev, const float* y_dev, int64_t B, float* cost_host);
int nvd_train_end(nvd_ctx* ctx, int trainer, float* const* W_host, float* const* b_host);

/* ---- multi-GPU exchange step ------------------------------------------------------------------
 * The reference is single-process; the path shards by global frame range, one context (process or
 * thread) per GPU, and the ONLY exchange is the sum of the error counters that calculate_error / calculate_ber
 * (src/fnn_viterbi.py:619-674, src/fnn_viterbi_mixed_snr.py:634-655) would have computed over the whole test set.
 * It is a sum of uint64 values over NVLink (peer memory, or ncclAllReduce) queued on the context's stream right behind the launches
 * that produce the counters.  NCCL is bound at run time (dlopen of libnccl.so.2; NVD_NCCL_LIB overrides): a host
 * without NCCL can still use every single-GPU entry point.
 *   rank 0:    nvd_comm_unique_id(id)  ->  ship the NVD_COMM_ID_BYTES bytes to the other ranks (any host channel)
 *   all ranks: nvd_comm_init(ctx, rank, world, id)  (collective)  ...  nvd_comm_allreduce_u64  ...  nvd_comm_destroy */
#define NVD_COMM_ID_BYTES 128
int nvd_comm_unique_id(void* id_out);
int nvd_comm_init(nvd_ctx* ctx, int rank, int world, const void* id);
int nvd_comm_destroy(nvd_ctx* ctx);
int nvd_comm_info(nvd_ctx* ctx, int* rank, int* world, int* nccl_version);
/* in-place sum over all ranks of n uint64 values in DEVICE memory (e.g. the [n_snr][4] counters of nvd_mc_run);
 * asynchronous on the context's stream; a no-op on a context without a communicator.  Goes over NVLink peer memory
 * when nvd_comm_p2p_open has succeeded (n <= 448), else through ncclAllReduce. */
int nvd_comm_allreduce_u64(nvd_ctx* ctx, uint64_t* values_dev, size_t n);
/* The same exchange without NCCL, over NVLink / NVSwitch peer memory: every rank owns an 8 KB exchange block; a
 * small kernel adds the rank's values into every rank's block with system-scope atomics and raises every rank's
 * arrival counter, a second one waits for all arrivals (csrc/comm.cu).  ~10 us for 8 ranks.
 *   all ranks: nvd_comm_p2p_handle(ctx, h)  ->  gather the NVD_COMM_IPC_BYTES bytes of every rank, in rank order
 *              (any host channel)  ->  nvd_comm_p2p_open(ctx, rank, world, handles)
 * NVD_ERR_UNSUPPORTED when a peer's block cannot be mapped (GPU not visible to this process, no P2P path): keep
 * using the NCCL transport then.  world = 1 is allowed (the kernels run against the rank's own block). */
#define NVD_COMM_IPC_BYTES 64
int nvd_comm_p2p_handle(nvd_ctx* ctx, void* handle_out);
int nvd_comm_p2p_open(nvd_ctx* ctx, int rank, int world, const void* handles);
int nvd_comm_p2p_close(nvd_ctx* ctx);   /* unmap the peers and free the block (all ranks, after a host barrier) */

/* Deep-BER driver (BASELINE.json configs[4]): per SNR point, chunks of `chunk_frames` GLOBAL frames are added until
 * the point has >= min_frame_errors frame errors (globally, after the all-reduce) or max_frames frames.  With a
 * communicator every rank takes its contiguous share of each chunk (cfg->frame0 + done + [rank, rank+1) * chunk / world)
 * and the counters are all-reduced after every chunk, so all ranks take the same stop decision.  Synchronous.
 * cfg->frames is ignored.  counters_host: uint64 [n_snr][4], overwritten, identical on every rank.
 * replaces the fixed-size test set of train_test_dataset (src/fnn_viterbi.py:380-388) as the stopping rule. */
int nvd_mc_sweep(nvd_ctx* ctx, const nvd_mc_config* cfg, const float* snr_db_host, const int32_t* snr_idx_host, int n_snr,
                 uint64_t min_frame_errors, uint64_t max_frames, uint64_t chunk_frames, uint64_t* counters_host);

/* ---- diagnostics ---------------------------------------------------------------------------
 * One tcgen05.mma (M=128, K=16, bf16 x bf16 -> fp32) on caller-supplied raw shared-memory images
 * and descriptor strides; used by tests/test_gpu_kernels.py::test_umma_layout_probe to pin the UMMA shared-memory
 * layout the tensor-core kernel relies on.  a_image_host: 4096 B, b_image_host: N*32 B,
 * d_host: [128, N] fp32. */
/* phase time stamps (clock64) of CTA 0's groups over their first 64 tiles of a fused run:
 * stamps_host[(tile * groups_per_cta + group) * 16 + phase], and behind them, from index 4096, every group's
 * [start ns, finish ns, tiles, SM id] (global timer), group index = CTA * groups_per_cta + group.  8192 values. */
int nvd_debug_tc_timing(nvd_ctx* ctx, int model, uint64_t frames, uint64_t* stamps_host);
/* same with the A operand in TENSOR MEMORY (written by tcgen05.st, K elements packed two bf16 per
 * column): a_rows_host = [128][8] uint32. */
int nvd_debug_umma_probe_ts(nvd_ctx* ctx, int N, int b_lbo_bytes, int b_sbo_bytes, const void* a_rows_host,
                            const void* b_image_host, float* d_host);
int nvd_debug_umma_probe(nvd_ctx* ctx, int N, int a_lbo_bytes, int a_sbo_bytes, int b_lbo_bytes,
                         int b_sbo_bytes, const void* a_image_host, const void* b_image_host,
                         float* d_host);

#ifdef __cplusplus
}
#endif
#endif /* NVD_H_ */
