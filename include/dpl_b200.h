/*
 * dpl_b200.h — C-ABI of libdpl_b200.so, the sm_100a calibration / rounding-finetune
 * kernels behind the Dipoorlet plugin API.
 *
 * Conventions
 *   - every entry point is extern "C", returns 0 on success or a non-zero status
 *     (a cudaError_t value, or DPL_E_* below); dpl_last_error() gives the text;
 *   - every `d_*` pointer is a DEVICE pointer owned by the caller (in the Python
 *     host: torch tensors' data_ptr()); the library never allocates, frees or
 *     keeps a pointer past the call; work is enqueued on `stream` and the call
 *     returns without synchronising;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - a "blob" is one activation tensor of the calibrated network for the images
 *     of the current batch: n_seg images x seg_len float32, contiguous (NCHW per
 *     image). A "segment" is one image of one blob — the unit the reference takes
 *     its per-image statistics on.
 *
 * Each entry point cites the reference code it replaces (paths relative to the
 * ModelTC/Dipoorlet checkout).
 */
#ifndef DPL_B200_H_
#define DPL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPL_VERSION 100

#define DPL_E_BADARG 10001   /* null pointer, zero size, unsupported bins ... */
#define DPL_E_WORKSPACE 10002 /* scratch buffer too small */
#define DPL_E_UNSUPPORTED 10003

/* Tile sizes (elements) of the launch-wide tilings the planner fills in. */
#define DPL_SEG_TILE 8192u
#define DPL_FLAT_TILE 8192u

/* One blob of the current batch. Filled by the caller except the *_begin fields,
 * which dpl_plan_blobs() computes. The same array is then uploaded to the device
 * and handed to the kernels. All fields are 8 bytes so that the layout is
 * identical from C, ctypes and numpy (dtype '<u8', shape [n_blobs, 8]). */
typedef struct dpl_blob {
  uint64_t ptr;            /* device address of float32[n_seg * seg_len], 4-byte aligned */
  uint64_t n_seg;          /* images in the batch */
  uint64_t seg_len;        /* elements per image */
  uint64_t seg_out_base;   /* index of this blob's first segment in per-segment outputs */
  uint64_t seg_tile_begin; /* first tile of this blob in the segment-respecting tiling */
  uint64_t flat_tile_begin;/* first tile of this blob in the flat (image-agnostic) tiling */
  uint64_t stat_index;     /* row of this blob in per-blob arrays (data_max, counts, ...) */
  uint64_t reserved;
} dpl_blob;

int dpl_version(void);
const char* dpl_last_error(void);

/* Host-side planner: fills seg_out_base / seg_tile_begin / flat_tile_begin of
 * host_blobs[0..n) and returns the totals. Pure host arithmetic, no CUDA call. */
int dpl_plan_blobs(dpl_blob* host_blobs, int n_blobs, uint64_t* n_segments,
                   uint64_t* n_seg_tiles, uint64_t* n_flat_tiles);

/* Scratch bytes dpl_segstats_f32 needs for a plan with n_seg_tiles tiles. */
size_t dpl_segstats_scratch_bytes(uint64_t n_seg_tiles);

/* K1 — per-segment min, max, sum|x| and count(|x|>0) of every blob of the batch in
 * ONE launch (+ a small finalize launch). Replaces the per-image NumPy reductions
 * `ort_outs[i].max()/.min()` of forward_get_minmax (dipoorlet/forward_net.py:220-235)
 * and the first line of the OCTAV loop `abs_x.sum() / abs_x[abs_x > 0].size`
 * (forward_net.py:323-324).
 *   d_min/d_max: float32[n_segments]; d_abssum: float64[n_segments];
 *   d_nnz: uint64[n_segments]   (any of d_abssum/d_nnz may be NULL)
 * Also folds the batch into running per-blob extrema when d_blob_min/d_blob_max
 * (float32[n_stats], indexed by stat_index, caller-initialised to +inf/-inf) are
 * non-NULL — the device-resident equivalent of find_clip_val_minmax
 * (dipoorlet/tensor_cali/basic_algorithm.py:20-21).
 * ctas_per_sm: 0 = the full-bandwidth grid (8 CTAs of 256 threads per SM); 1..8 = a smaller grid
 * for a caller that runs this pass on a side stream underneath the forward's tensor-core kernels. */
int dpl_segstats_f32(const dpl_blob* d_blobs, int n_blobs, uint64_t n_segments,
                     uint64_t n_seg_tiles, float* d_min, float* d_max, double* d_abssum,
                     uint64_t* d_nnz, float* d_blob_min, float* d_blob_max, void* d_scratch,
                     size_t scratch_bytes, int ctas_per_sm, void* stream);

/* data_max[b] = max(blob_max[b], -blob_min[b]) as float32 — forward_net.py:266-267. */
int dpl_absmax_f32(const float* d_blob_min, const float* d_blob_max, float* d_data_max,
                   int n_stats, void* stream);

/* K2 — np.histogram(np.abs(x), bins, (0, data_max)) accumulated over every image of
 * every blob of the batch in ONE launch, bit-exact with NumPy's uniform-bin path
 * (float32 edges, edge-corrected index, right edge inclusive, data_max == 0 widened
 * to (-0.5, 0.5), |x| > data_max dropped). Replaces forward_get_hist's per-image
 * np.histogram + the np.stack(hist).sum(0) of find_clip_val_hist
 * (dipoorlet/forward_net.py:265-280, tensor_cali/basic_algorithm.py:37-38).
 *   d_data_max: float32[n_stats]; d_counts: uint64[n_stats * bins], accumulated into
 *   (caller zeroes before the first batch). variant: 0 = auto, see dpl_stats.cu. */
int dpl_hist_abs_f32(const dpl_blob* d_blobs, int n_blobs, uint64_t n_flat_tiles,
                     const float* d_data_max, int bins, unsigned long long* d_counts,
                     int variant, void* stream);

/* K3 — percentile clip search over the accumulated histograms, one tensor per CTA,
 * sequential float64 accumulation so that the selected bin is the reference's.
 * Replaces find_clip_val_hist's Python loop (tensor_cali/basic_algorithm.py:40-53).
 *   d_clip: float32[n_stats * 2] = {lo, hi}; d_bin: int32[n_stats] (-1 = fallback to
 *   full range, basic_algorithm.py:51-53). */
int dpl_hist_percentile(const unsigned long long* d_counts, int n_stats, int bins,
                        double threshold, const float* d_data_max, const float* d_blob_min,
                        const float* d_blob_max, float* d_clip, int* d_bin, void* stream);

/* K4 — OCTAV fixed point per segment (the 'mse' calibrator): persistent CTAs claim
 * segments longest-first from a device work queue; each segment is read from HBM once,
 * survivors {|x| > s} are compacted into an L2-resident scratch slice.
 * Replaces forward_net_octav's NumPy loop (dipoorlet/forward_net.py:316-330).
 *   k_const = 1 / 4**8 / 3 / unsigned; d_s: float32[n_segments]; d_iters (optional):
 *   int32[n_segments] updates taken; needs d_abssum / d_nnz from dpl_segstats_f32 for
 *   s0 and a scratch of dpl_octav_scratch_bytes(longest seg_len, n_segments) bytes.
 *   A segment with a NaN (or no non-zero element: 0 / 0) yields NaN, as the reference. */
size_t dpl_octav_scratch_bytes(uint64_t max_seg_len, uint64_t n_segments);
int dpl_octav_f32(const dpl_blob* d_blobs, int n_blobs, uint64_t n_segments,
                  uint64_t max_seg_len, const double* d_abssum, const uint64_t* d_nnz,
                  double k_const, int max_iter, float* d_s, int* d_iters, void* d_scratch,
                  size_t scratch_bytes, void* stream);

/* K5 — fake quantisation y = (clamp(rne(x / scale) + zp, qlo, qhi) - zp) * scale.
 * Per-tensor (n_channels = 1) or per-channel along an outer axis: element i belongs
 * to channel (i / inner) % n_channels. Replaces the QuantizeLinear/DequantizeLinear
 * pair emitted by make_quant_dequant (dipoorlet/quantize.py:197-239, executed inside
 * onnxruntime) and quant_acti (weight_transform/ada_quant_layer.py:28-36).
 * drop_prob < 1 keeps x where a counter-based uniform draw >= drop_prob (QDrop). */
int dpl_fakequant_f32(const float* d_x, float* d_y, uint64_t n, const float* d_scale,
                      const int32_t* d_zero_point, int n_channels, uint64_t inner, int qlo,
                      int qhi, float drop_prob, uint64_t seed, void* stream);

/* K7a — per-channel mean of (a - b) over (N, H, W): out[c] (+)= sum / count.
 * Replaces update_conv_node_bias (weight_transform/bias_correction.py:10-13).
 *   a, b: float32[n_img, channels, inner]; d_sum: float64[channels] accumulated. */
int dpl_channel_sumdiff_f32(const float* d_a, const float* d_b, uint64_t n_img,
                            uint64_t channels, uint64_t inner, double* d_sum, void* stream);

/* K7b — per-segment sum(a*b), sum(a*a), sum(b*b) for cos_similarity
 * (dipoorlet/utils.py:273-278). d_out: float64[n_seg * 3]. */
int dpl_cosine3_f32(const float* d_a, const float* d_b, uint64_t n_seg, uint64_t seg_len,
                    double* d_out, void* stream);

/* K6 helpers — AdaRound elementwise pieces around the conv/GEMM re-evaluation
 * (weight_transform/ada_quant_layer.py:39-50,96-110; adaround.py:119-144). */

/* alpha0 = -log((zeta-gamma)/(rest-gamma) - 1), rest = w/s - floor(w/s); also writes
 * floor(w/s). Per-channel scale along axis 0 (inner = elements per out channel). */
int dpl_adaround_init_f32(const float* d_w, const float* d_scale, int n_channels,
                          uint64_t inner, float* d_alpha, float* d_wfloor, void* stream);

/* w_soft = clamp(wfloor + h(alpha), qmin, qmax) * s (soft=1) or
 *          clamp(wfloor + (alpha >= 0), qmin, qmax) * s (soft=0). */
int dpl_adaround_weight_f32(const float* d_wfloor, const float* d_alpha, const float* d_scale,
                            int n_channels, uint64_t inner, float qmin, float qmax, int soft,
                            float* d_wq, void* stream);

/* One fused optimiser step on alpha given dL/dW_soft:
 *   g = dW * s * h'(alpha) * [not clamped] + d/dalpha( reg_alpha * sum(1 - |2h-1|^beta) )
 *   Adam(lr, b1, b2, eps, step) in place on (alpha, m, v).
 * Also accumulates the regulariser value into d_reg (float64[1]) when non-NULL.
 * d_sched (optional, float32[3] = {beta, 1-b1^t, sqrt(1-b2^t)} written by dpl_recon_schedule)
 * overrides beta / step so that a captured CUDA graph can be replayed every iteration. */
int dpl_adaround_step_f32(const float* d_grad_w, const float* d_wfloor, const float* d_scale,
                          int n_channels, uint64_t inner, float qmin, float qmax, float beta,
                          float reg_alpha, float lr, float b1, float b2, float eps, int step,
                          float grad_scale, float* d_alpha, float* d_m, float* d_v,
                          double* d_reg, const float* d_sched, void* stream);

/* K6 step with the gradient all-reduce inside: replaces DistributedDataParallel's NCCL
 * all-reduce of dL/dW (adaround.py:121, brecq.py:163) followed by the step. peer_grads / peer_words are HOST
 * arrays of `world` (<= 8) device pointers, in rank order, into buffers every rank has mapped (CUDA IPC, e.g.
 * torch symmetric memory): peer_grads[r] = rank r's dL/dW slot for this epoch, peer_words[r] = rank r's
 * `world` 32-bit arrival words (zeroed once, before the first call, with a barrier after the zeroing).
 * The kernel announces this rank's arrival at `epoch` (>= 1, +1 per call on every rank, the slot
 * alternating between two buffers), waits for every peer's (bounded: sets *d_error and leaves the state
 * untouched after ~2 s), sums the `world` gradients in rank order (bit-identical on every rank), scales by
 * 1 / world and applies dpl_adaround_step_f32's update. */
int dpl_adaround_step_peer_f32(const void* const* peer_grads, void* const* peer_words, int world, int rank,
                               uint32_t epoch, const float* d_wfloor, const float* d_scale, int n_channels,
                               uint64_t inner, float qmin, float qmax, float beta, float reg_alpha, float lr,
                               float b1, float b2, float eps, int step, float* d_alpha, float* d_m,
                               float* d_v, double* d_reg, const float* d_sched, int* d_error, void* stream);

/* K6 epilogues — layer activation of the reconstruction loop: y = [drop-]fakequant(relu(o))
 * (AdaQLayer.forward tail, weight_transform/ada_quant_layer.py:245-251; quant_acti :28-36),
 * its backward (round() has zero gradient: only non-quantised elements pass), the fused
 * L2 loss + dL/do of the block output (L2_norm :113-114: sum over channels, mean over the
 * rest => inv_count = channels / numel), and the QDrop block input
 * where(u < prob, q_in, fp_in) (brecq.py:169-170). Masks are regenerated from (seed, index). */
int dpl_recon_act_f32(const float* d_o, float* d_y, uint64_t n, int relu, int quant, float scale,
                      float qmin, float qmax, float prob, uint64_t seed,
                      const unsigned long long* d_seed, void* stream);
int dpl_recon_act_bwd_f32(const float* d_o, const float* d_gy, float* d_go, uint64_t n, int relu,
                          int quant, float scale, float qmin, float qmax, float prob,
                          uint64_t seed, const unsigned long long* d_seed, void* stream);
int dpl_recon_loss_f32(const float* d_o, const float* d_tgt, float* d_go, uint64_t n, int relu,
                       int quant, float scale, float qmin, float qmax, float prob, uint64_t seed,
                       float inv_count, double* d_loss, const unsigned long long* d_seed,
                       void* stream);
/* (d_seed, optional: the mask seed is read from device memory instead of `seed`.)
 * dpl_recon_schedule: one-thread kernel computing the scalars of iteration t = *d_iter — the
 * regulariser temperature beta(t) (TempDecay, ada_quant_layer.py:117-130), Adam's bias
 * corrections and one mask seed per layer — then t += 1. With it a whole iteration of the
 * reconstruction loop is a replayable CUDA graph. */
int dpl_recon_schedule(int* d_iter, float* d_sched, unsigned long long* d_seeds, int n_seeds,
                       double t_max, double rel_start, double start_b, double end_b, double b1,
                       double b2, uint64_t seed_base, void* stream);
int dpl_mix_drop_f32(const float* d_a, const float* d_b, float* d_y, uint64_t n, float prob,
                     uint64_t seed, void* stream);

#define DPL_E_TIMEOUT 10004   /* a bounded in-kernel wait expired (reported through d_error_flag) */

/* K6 dense tile — TF32 GEMM on tcgen05 tensor cores (TMA-staged operands, TMEM accumulators):
 *   D[z][m][n] (+)= sum_k A[za][m][k] * B[z][k][n]  (+ bias, relu), fp32 in / out.
 * Replaces the F.conv2d (1x1) / F.linear re-evaluation and their autograd gradients inside
 * AdaQLayer.forward / learning_round_mask (weight_transform/ada_quant_layer.py:224-244,
 * adaround.py:119-135), which torch runs on cuDNN with TF32 allowed.
 *   a_major / b_major: 0 = K-major (element (row, k) at row * ld + k), 1 = MN-major (k * ld + row)
 *   a_batch_stride = 0: A shared by all batch slices (weights)
 *   fold_batch = 1: the batch is folded into K (weight gradient); split_k CTAs along it add
 *                   their partial tiles atomically into D (caller zeroes D when split_k > 1)
 *   bias_mode: 0 none, 1 per row m, 2 per column n
 *   d_error_flag: int32 on the device, set to 1 if a pipeline wait timed out (never hangs)
 * TMA needs 16-byte aligned bases and leading dimensions that are multiples of 4 floats;
 * otherwise DPL_E_UNSUPPORTED is returned and the caller keeps its other path. */
int dpl_gemm_tf32(const float* d_a, int a_major, long long lda, long long a_batch_stride,
                  const float* d_b, int b_major, long long ldb, long long b_batch_stride,
                  float* d_d, long long ldd, long long d_batch_stride, int M, int N, int K,
                  int batch, int fold_batch, int split_k, const float* d_bias, int bias_mode,
                  int relu, int* d_error_flag, void* stream);

/* 3xTF32 variant: fp32-accurate products on the TF32 tensor cores,
 *   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo   (x_hi = top 19 bits, x_lo = x - x_hi),
 * for the calibration forward (activations must stay within fp32 rounding of the reference's
 * fp32 ORT path, dipoorlet/forward_net.py:200-216). The weight operand is split once by
 * dpl_tf32_split_f32 (pass its d_hi as d_a and its d_lo as d_a_lo; dpl_tf32_residual_f32 of the raw
 * weight also works but keeps the truncation bias); the residual of B is formed in shared memory
 * inside the kernel. NCHW 1x1-convolution shapes (a_major 0, b_major 1, shared A, bias per row) run on
 * the persistent kernel with CHUNKED accumulation (csrc/dpl_x3p.cuh: the tensor core accumulates
 * DPL_X3_CHUNK = 2 K blocks in TMEM, eight warps drain every chunk into registers with round-to-nearest
 * adds) - tcgen05's truncating accumulation otherwise shrinks every layer's output by ~1e-6 and the
 * error compounds over a deep network. DPL_X3_CHUNK=0 selects the round-1 kernels.
 * d_d_relu (optional): a second output max(D, 0) with D's layout — the Relu node behind a Conv is
 * a calibration blob of its own, so the epilogue writes both instead of a second pass.
 * d_blob_min / d_blob_max / d_relu_min / d_relu_max (optional, one float each): fused range statistics of D
 * and of the Relu output, as in dpl_clip_f32 (same arguments on dpl_conv_taps_tf32x3). */
int dpl_tf32_residual_f32(const float* d_x, float* d_lo, uint64_t n, void* stream);
/* Unbiased split of a weight for the 3xTF32 kernels: d_hi = RN_tf32(x) (round to nearest instead of the
 * truncation kind::tf32 applies to a raw fp32 pattern), d_lo = RN_tf32(x - d_hi). Pass d_hi as the weight. */
int dpl_tf32_split_f32(const float* d_x, float* d_hi, float* d_lo, uint64_t n, void* stream);
int dpl_gemm_tf32x3(const float* d_a, const float* d_a_lo, int a_major, long long lda,
                    long long a_batch_stride, const float* d_b, int b_major, long long ldb,
                    long long b_batch_stride, float* d_d, long long ldd, long long d_batch_stride,
                    int M, int N, int K, int batch, const float* d_bias, int bias_mode, int relu,
                    float* d_d_relu, float* d_blob_min, float* d_blob_max, float* d_relu_min,
                    float* d_relu_max, int* d_error_flag, void* stream);

/* k x k / strided convolutions of the calibration forward (the Conv nodes that ORT executes in
 * dipoorlet/forward_net.py:200-216), fp32-accurate (3xTF32) on the tcgen05 tensor cores as a
 * shifted-window implicit GEMM over a channel-last staging copy of the input:
 *   dpl_pad_plane_f32     Xp[(plane * n_img + img) * Hp * Wp + hp * Wp + wp][c] =
 *                           X[img][c][stride (hp - origin) + a][stride (wp - origin) + b]   (0 outside),
 *                         plane = a * stride + b < n_planes  (stride 1: one bordered plane; stride 2:
 *                         the four parity planes)
 *   dpl_conv_taps_tf32x3  Y[img][co][ho][wo] = bias[co] + sum_tap sum_ci Wt[tap][co][ci] *
 *                           Xp[q + tap_shift[tap]][ci],  q = img * Hp * Wp + (ho + origin) * Wp + (wo + origin);
 *                         d_w_taps tap-major [n_taps][c_out][c_in], d_w_taps_lo = dpl_tf32_residual_f32
 *                         of it; tap_shift: HOST array of n_taps (<= 9) row offsets; c_in multiple of 4.
 * 3x3 stride 1 pad 1:  Hp = H + 2, Wp = W + 2, origin 1, shift = (kh - 1) Wp + (kw - 1).
 * 3x3 stride 2 pad 1:  Hp = Ho + 1, Wp = Wo + 1, origin 1, tap (kh, kw) -> plane ((kh + 1) & 1, (kw + 1) & 1),
 *                      shift = plane * n_img * Hp * Wp - (kh == 0) * Wp - (kw == 0).
 * 1x1 stride 2:        Hp = Ho, Wp = Wo, origin 0, one plane, one tap, shift 0. */
int dpl_pad_plane_f32(const float* d_x, float* d_xp, int n_img, int channels, int H, int W, int stride,
                      int origin, int Hp, int Wp, int n_planes, void* stream);
int dpl_conv_taps_tf32x3(const float* d_xp, long long total_rows, const float* d_w_taps,
                         const float* d_w_taps_lo, float* d_y, int n_img, int c_in, int c_out, int Ho,
                         int Wo, int Hp, int Wp, int origin, int n_taps, const int* tap_shift,
                         const float* d_bias, int relu, float* d_y_relu, float* d_blob_min, float* d_blob_max,
                         float* d_relu_min, float* d_relu_max, int* d_error_flag, void* stream);

/* Direct fp32 convolution for layers with very few input channels (ResNet's 7x7 / stride 2 stem,
 * MobileNetV2's 3x3 / stride 2 stem on the 3 image channels; the Conv node ORT executes in
 * dipoorlet/forward_net.py:200-216): exact fp32 FMA accumulation like the reference's CPU path,
 * input patch + transposed filter in shared memory, 4 pixels x 16 channels per thread.
 *   Y[img][co][ho][wo] = bias[co] + sum W[co][c][a][b] * X[img][c][ho * stride - pad + a][wo * stride - pad + b]
 * (kh, kw, stride) in {(7,7,2), (5,5,1|2), (3,3,1|2)}, symmetric padding; DPL_E_UNSUPPORTED otherwise or when the
 * patch + filter exceed 100 KB of shared memory. d_y_relu / d_blob_* / d_relu_*: as in dpl_gemm_tf32x3. */
int dpl_conv_direct_f32(const float* d_x, const float* d_w, const float* d_bias, float* d_y, int n_img,
                        int channels, int H, int W, int c_out, int kh, int kw, int stride, int pad, int Ho,
                        int Wo, float* d_y_relu, float* d_blob_min, float* d_blob_max, float* d_relu_min,
                        float* d_relu_max, void* stream);

/* Depthwise convolution (group = channels = c_out, depth multiplier 1; MobileNetV2's 3x3 layers, the
 * same ORT Conv node): one FMA per tap and output, HBM bound, k in {3, 5}, symmetric padding.
 * d_w: [channels][1][k][k]. d_blob_min / d_blob_max: fused range statistics as in dpl_clip_f32. */
int dpl_dwconv2d_f32(const float* d_x, const float* d_w, const float* d_bias, float* d_y, int n_img,
                     int channels, int H, int W, int k, int stride, int pad, int Ho, int Wo,
                     float* d_blob_min, float* d_blob_max, void* stream);

/* im2col staging for a convolution with very few input channels (ResNet's 7x7 / stride 2 stem, 3
 * channels): d_xp[(img * Ho + ho) * Wo + wo][(c * kh + a) * kw + b] = X[img][c][ho * stride - pad + a]
 * [wo * stride - pad + b] (0 outside, 0 for columns >= C kh kw); k_pad a multiple of 4, <= 256. The
 * convolution is then dpl_conv_taps_tf32x3 with ONE tap, c_in = k_pad, Hp = Ho, Wp = Wo, origin 0 over
 * the weight viewed as [c_out][C kh kw] (zero-padded to k_pad columns). */
int dpl_im2col_f32(const float* d_x, float* d_xp, int n_img, int channels, int H, int W, int kh, int kw,
                   int stride, int pad, int Ho, int Wo, int k_pad, void* stream);

/* Pixel-major 1x1 convolution of the calibration forward, 3xTF32, straight from / to NCHW:
 *   Y[img][co][px] = bias[co] + sum_ci W[co][ci] * X[img][ci][px]
 * with the pixels on the TMEM lanes and the activations as the TMEM operand of tcgen05.mma (split
 * into TF32 pattern + residual in registers), weights + host residual by TMA; 128 px x 64 co tiles,
 * two CTAs per SM; by default (DPL_X3_CHUNK > 0, DPL_X3_TS = 1) the persistent kernel with chunked
 * accumulation (128 px x 128 co tiles, csrc/dpl_x3ts.cuh), which also folds the fused range statistics
 * d_blob_* / d_relu_* (see dpl_clip_f32). d_y_relu (optional) also receives max(Y, 0). hw and c_in must be multiples of 4
 * (TMA strides), otherwise DPL_E_UNSUPPORTED. Same reference operator as dpl_gemm_tf32x3
 * (the Conv nodes ORT executes in dipoorlet/forward_net.py:200-216). */
int dpl_conv1x1_px_tf32x3(const float* d_x, const float* d_w, const float* d_w_lo, float* d_y, int n_img,
                          int c_in, int c_out, int hw, const float* d_bias, float* d_y_relu,
                          float* d_blob_min, float* d_blob_max, float* d_relu_min, float* d_relu_max,
                          int* d_error_flag, void* stream);

/* Non-GEMM operators of the calibration forward — the Relu / Clip / Add / MaxPool /
 * GlobalAveragePool nodes onnxruntime executes per image in dipoorlet/forward_net.py:200-216 —
 * each one streaming pass over a whole batch of blobs (HBM bound).
 *   dpl_clip_f32            y = min(max(x, lo), hi), NaN propagates (Relu: lo = 0, hi = +inf)
 *   dpl_add_f32             y = a + b; when d_y_relu is non-NULL also y_relu = max(y, 0): the
 *                           residual Add and the Relu that follows it, both blobs from one read
 *   dpl_maxpool2d_f32       planes = n_img * channels planes of H x W -> Ho x Wo, padding = -inf
 *   dpl_global_avgpool_f32  y[plane] = mean of the plane's hw elements (fp32)
 * Fused range statistics: d_blob_min / d_blob_max (optional, each ONE float32 = this blob's entry of
 * the running per-blob extrema that dpl_segstats_f32 maintains, caller-initialised to +inf / -inf)
 * receive min / max of the values the kernel writes, so a minmax / hist calibration does not read these
 * blobs again for its range pass (find_clip_val_minmax, tensor_cali/basic_algorithm.py:20-21);
 * d_relu_min / d_relu_max: the same for dpl_add_f32's second output. */
int dpl_clip_f32(const float* d_x, float* d_y, uint64_t n, float lo, float hi, float* d_blob_min,
                 float* d_blob_max, void* stream);
int dpl_add_f32(const float* d_a, const float* d_b, float* d_y, float* d_y_relu, uint64_t n,
                float* d_blob_min, float* d_blob_max, float* d_relu_min, float* d_relu_max, void* stream);
int dpl_maxpool2d_f32(const float* d_x, float* d_y, uint64_t planes, int H, int W, int kh, int kw,
                      int sh, int sw, int pad_top, int pad_left, int Ho, int Wo, float* d_blob_min,
                      float* d_blob_max, void* stream);
int dpl_global_avgpool_f32(const float* d_x, float* d_y, uint64_t planes, uint64_t hw, float* d_blob_min,
                           float* d_blob_max, void* stream);

/* ---- K6 dense contraction for k x k / strided / depthwise convolutions (single-pass TF32) ----------------
 * Replaces F.conv2d and the weight / data gradients autograd derives from it inside AdaQLayer.forward /
 * learning_round_mask (weight_transform/ada_quant_layer.py:224-244, adaround.py:119-135,
 * brecq.py:163-186), which torch runs on cuDNN with TF32 allowed. All three contractions work on the
 * channel-last, zero-bordered staging copies written by dpl_pad_plane_f32 (Xp of the layer input, Gp of
 * the output gradient), in which a filter tap is a row shift.
 *
 * dpl_tap_conv_tf32 (forward, and data gradient with the roles of the channels swapped):
 *   Y[img][cn][hq*os + oa][wq*os + ob] = bias[cn] + sum_tap sum_ck Wt[tap_w[tap]][cn][ck] * Xp[q + tap_shift[tap]][ck]
 *   q = img*Hp*Wp + (hq + origin)*Wp + (wq + origin); points whose pixel is outside H x W are dropped.
 *   ck must be a multiple of 4 (else DPL_E_UNSUPPORTED). tap_shift / tap_w: host arrays of n_taps <= 9.
 * dpl_tap_wgrad_tf32: dW[co][ci][tap_col[t]] = sum_q Gp[q][co] * Xp[q + tap_shift[t]][ci]; the whole batch is
 *   the K dimension, split over CTAs that add into the zeroed d_dw [c_out][c_in][t_full].
 * dpl_taps_layout_f32: w [c_out][c_in][T] -> d_wf [T][c_out][c_in] and / or d_wd [T][c_in][c_out].
 * dpl_dwconv2d_wgrad_f32 / _dgrad_f32: depthwise (group = channels) gradients on the FMA pipe, exact fp32,
 *   k in {3, 5}; d_gw [C][k][k] and d_gx [n_img][C][H][W] are fully overwritten. */
int dpl_tap_conv_tf32(const float* d_xp, long long total_rows, const float* d_w_taps, int n_w_taps,
                      float* d_y, int n_img, int ck, int cn, int H, int W, int Hp, int Wp, int origin,
                      int out_stride, int out_a, int out_b, int n_taps, const int* tap_shift,
                      const int* tap_w, const float* d_bias, int* d_error_flag, void* stream);
int dpl_tap_wgrad_tf32(const float* d_gp, long long q_total, const float* d_xp, long long x_rows,
                       float* d_dw, int c_out, int c_in, int t_full, int n_taps, const int* tap_shift,
                       const int* tap_col, int* d_error_flag, void* stream);
int dpl_taps_layout_f32(const float* d_w, float* d_wf, float* d_wd, int c_out, int c_in, int T, void* stream);
int dpl_dwconv2d_wgrad_f32(const float* d_x, const float* d_gy, float* d_gw, int n_img, int C, int H, int W,
                           int k, int stride, int pad, int Ho, int Wo, void* stream);
int dpl_dwconv2d_dgrad_f32(const float* d_gy, const float* d_w, float* d_gx, int n_img, int C, int H, int W,
                           int k, int stride, int pad, int Ho, int Wo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DPL_B200_H_ */
