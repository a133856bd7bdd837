/* libwsunet — C ABI of the B200-native UNet -> Weighted-Stego (WS) hot path.
 *
 * The reference (uibk-uncover/ws-unet) is pure Python and has no FFI of its own: its
 * boundary is three duck-typed Python interfaces. Each entry point below names the reference call it serves;
 * the Python shim in ws_unet_b200/ binds them with ctypes and re-exposes the reference's own signatures.
 *
 * Conventions: every function returns 0 on success or a negative wsu_status; wsu_last_error() returns a
 * thread-local message for the last failure. Pointers named *_dev are device pointers owned by the caller on the
 * handle's device; the library never frees caller memory. `stream` is a cudaStream_t passed as void* (NULL = the
 * legacy default stream). A handle is bound to one device and must be used by one host thread at a time.
 * There is no CPU fallback: on a machine without a CUDA device every compute entry point fails with
 * WSU_ERR_CUDA.
 */
#ifndef WSUNET_H_
#define WSUNET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wsu_context* wsu_handle;

enum wsu_status {
  WSU_OK = 0,
  WSU_ERR_INVALID = -1, /* bad argument / unsupported configuration (Python shim raises ValueError/NotImplementedError) */
  WSU_ERR_CUDA = -2,    /* CUDA runtime/driver failure (RuntimeError) */
  WSU_ERR_STATE = -3    /* weights missing / not committed */
};

enum wsu_dtype { WSU_U8 = 0, WSU_F32 = 1 };

/* linear predictors: src/ws/estimate.py:31-52 and src/filters/evaluate.py:29-50 (NAMED_FILTERS[_2D]) */
enum wsu_predictor { WSU_PRED_KB = 0, WSU_PRED_AVG = 1, WSU_PRED_AVG9 = 2, WSU_PRED_ID = 3 };

/* `weighted` argument of attack(): src/ws/estimate.py:55-110 */
enum wsu_weighting { WSU_UNWEIGHTED = 0, WSU_WEIGHTED = 1, WSU_ANTIWEIGHTED = -1 };

const char* wsu_last_error(void);
int wsu_version(void);

/* ---- predictor module: src/unet/model/__init__.py:8-27 get_model('unet_<nsteps>', in_channels, out_channels) and
 *      src/unet/model/unet.py:54-135 UNet.__init__. out_channels must be 1 (the only value the reference uses). */
int wsu_create(wsu_handle* out, int device, int nsteps, int in_channels, int out_channels);
int wsu_destroy(wsu_handle h);

/* nn.Module.load_state_dict (src/unet/evaluate.py:184-186): one call per state_dict entry, e.g. "e11.weight" with
 * dims (64,in,3,3), "upconv3.weight" with dims (256,128,2,2), "outconv.bias" with dims (1). `data` is HOST fp32 in
 * PyTorch's contiguous layout. wsu_commit_weights() packs (split-bf16, tile-swizzled) and uploads. */
int wsu_load_weights(wsu_handle h, const char* name, const float* data, const int64_t* dims, int ndims);
int wsu_commit_weights(wsu_handle h);

/* options: "micro_batch" (images per pass through the layer chain; 0 = auto);
 *          "fuse_e11" (default 0; 1: the first convolution is computed inside e12's producer warps and its feature map
 *                      never touches HBM - bit-identical but measured slower);
 *          "l2_prefetch" (default 0; 1: 3x3 kernels warm L2 with the boxes of their next work item - measured 1 % slower);
 *          "a_collector" (default 1: the CTA-pair kernel issues hi*hi, hi*lo, lo*hi and keeps A_hi in the tensor core's
 *                         A collector for its second product; 0: every MMA re-reads A from shared memory);
 *          "cta_pair" (default 2: every 3x3 layer runs as CTA pairs, tcgen05 cta_group::2; 1: Cout >= 128 layers only; 0 never);
 *          "precision" (default 0: every layer three-term split-bf16; 1 / 2: the layers whose input lives at UNet level >= 1
 *                       read ONE fp16 activation plane with two / one MMA per MAC; 3: plan 2 plus the full-resolution layers
 *                       on an fp16 plane + an e4m3 correction plane, one kind::f16 + one kind::f8f6f4 MMA per MAC, about 15
 *                       bits per operand for any weights (needs in_channels == 1) - see ws_unet_b200/unet/model.py
 *                       set_precision / calibrate_precision; wsu_get_info "precision" reports what is active);
 *          "upconv_resident" (default 1: transposed convs keep their weights in shared memory; 0: per-phase kernel);
 *          "halo" (default 1: 3x3 layers load one haloed box per channel block; 0: per-tap reload kernel);
 *          "tma_store" (default 1: interior boxes of the 3x3 layers leave through TMA tensor stores, bit-identical, 1-3 % of the chain);
 *          "w_resident" (default 1: under the reduced plans the 3x3 layers with ONE input channel block - e12, d42 under precision 3,
 *                        e21 under precision 2 / 3 - keep all nine taps' weights in shared memory instead of streaming them per item);
 *          "alias_buffers" (default 1: feature maps with disjoint lifetimes share arena bytes; 0 for layer inspection);
 *          "dbg" (default 0, also env WSU_DBG: knock-out switches for TIMING EXPERIMENTS - results are wrong when set; bit 0 no
 *                 pooled output, bit 1 no main-output stores, bit 2 no epilogue work, bit 3 no e4m3 correction MMA, bit 4 staging
 *                 without global stores, bit 5 per-lane stores without staging; profiles/r02_knockout_timings.md);
 *          "profile" (1: record CUDA events around every layer launch of the last micro-batch) */
int wsu_set_option(wsu_handle h, const char* key, int64_t value);

/* UNet.forward (src/unet/model/unet.py:137-189): x_dev (B,in_channels,H,W) in [0,1] as float32, or uint8 pixels
 * (then x/255 is applied as in src/unet/evaluate.py:45) -> y_dev (B,1,H,W) float32 sigmoid output in (0,1).
 * H and W must be divisible by 2^nsteps and >= 2 at the deepest level (same constraint as the reference). */
int wsu_unet_forward(wsu_handle h, const void* x_dev, int x_dtype, float* y_dev, int B, int H, int W, void* stream);

/* Fused UNet predictor -> WS estimator. Semantics:
 *   crop=1, weighted=0, clip=0 : src/unet/evaluate.py:109-139 predict_unet   (beta_hat, l1)
 *   crop=1, weighted in {0,1,-1}, clip=1 : src/ws/estimate.py:55-136 attack with the UNet pixel_estimator
 *   crop=0, weighted=0, clip=1 : src/_defs/losses.py:46-61 WSLoss._error betas_hat (float images allowed)
 *   correct_bias=1 (uint8 images): estimate.py:126-128 - a second pass of the predictor over the LSB-difference image
 *     (x_bar - x) / 255, formed inside the first-layer kernel, whose head accumulates sum w (x - x_bar) x_bias next to the
 *     first pass's sums; beta_hat -= beta_hat * that. No prediction map leaves the GPU's on-chip memory in either pass.
 * img_dev (B,1,H,W) uint8 or float32 in [0,1]; beta_dev (B) float32; l1_dev (B) or NULL; yhat_dev (B,1,H,W) or NULL. */
int wsu_unet_ws_estimate(wsu_handle h, const void* img_dev, int img_dtype, int B, int H, int W, int weighted, int clip,
                         int crop, int correct_bias, float* beta_dev, float* l1_dev, float* yhat_dev, void* stream);

/* Same, but img_host/beta_host/l1_host are HOST buffers: chunks are copied H2D on a side stream overlapping compute,
 * results copied back; returns after everything has completed. This is the end-to-end call bench.py times. */
int wsu_unet_ws_estimate_host(wsu_handle h, const uint8_t* img_host, int B, int H, int W, int weighted, int clip, int crop,
                              int correct_bias, float* beta_host, float* l1_host);

/* ---- linear predictors, no handle needed: src/filters/evaluate.py:136-146 infere_single/get_filter_estimator.
 * xhat_dev (B,H-2,W-2) float32 in pixel units ('valid' 3x3 correlation). */
int wsu_filter_predict(int device, const void* img_dev, int img_dtype, int kind, float* xhat_dev, int B, int H, int W,
                       void* stream);

/* src/ws/estimate.py:55-136 attack with a NAMED_FILTERS pixel_estimator, mean_estimator=AVG; one pass over the image. */
int wsu_filter_ws_estimate(int device, const void* img_dev, int img_dtype, int kind, int weighted, int clip,
                           int correct_bias, float* beta_dev, float* l1_dev, int B, int H, int W, void* stream);
int wsu_filter_ws_estimate_host(int device, const uint8_t* img_host, int kind, int weighted, int clip, int correct_bias,
                                float* beta_host, float* l1_host, int B, int H, int W);

/* WS reduction against caller-supplied predictions in pixel units (any pixel_estimator; also the second pass of
 * correct_bias, src/ws/estimate.py:126-128). xhat_cropped=1: xhat/xbias are (B,H-2,W-2); 0: (B,H,W).
 * xbias_dev may be NULL (no bias correction). */
int wsu_ws_from_prediction(int device, const void* img_dev, int img_dtype, const float* xhat_dev, int xhat_cropped,
                           const float* xbias_dev, int weighted, int clip, int crop, float* beta_dev, float* l1_dev, int B,
                           int H, int W, void* stream);

/* Gradient of beta_hat with respect to the prediction, for the training losses WSLoss / L1WSLoss
 * (src/_defs/losses.py:45-115): beta_hat_b = (1/n) sum_crop (x - x_bar)(x - scale*o) is linear in o, so
 * grad_dev[b,p] = coef_dev[b] * (-scale) * (x_p - x_bar_p) / n inside the crop, 0 outside. coef_dev (B) float32 is
 * dLoss/dbeta_hat_b from the caller's autograd; scale = 255 for outputs in [0,1], 1 for pixel units;
 * grad_dev (B,1,H,W) float32. */
int wsu_ws_grad_prediction(int device, const void* img_dev, int img_dtype, const float* coef_dev, int crop, float scale,
                           float* grad_dev, int B, int H, int W, void* stream);

/* UniformDropout.forward (src/unet/model/unet.py:32-42): out = x * mask + KB(reflect-padded x) * (1 - mask) on the channels
 * whose bit is set in channel_mask, a copy elsewhere. x_dev (B,C,H,W) float32 in [0,1] or uint8 pixels (scaled by 1/255);
 * mask_dev (B,1,H,W) float32 of 0 / 1 (the caller draws it); out_dev (B,C,H,W) float32. */
int wsu_uniform_dropout(int device, const void* x_dev, int x_dtype, const float* mask_dev, float* out_dev, int B, int C,
                        int H, int W, uint32_t channel_mask, void* stream);

/* Matrix form of the linear predictors (src/filters/evaluate.py:53-76 get_filter_residuals over the N x 9 neighbour matrix
 * of src/_defs/filters.py:39-69): resid[i] = mat[i][8] - sum_k coef[k] * mat[i][k], accumulated in float64 left to right.
 * mat_dtype: 0 uint8, 1 float32, 2 float64; coef_dev 8 doubles; resid_dev n_rows doubles. */
int wsu_filter_residual_rows(int device, const void* mat_dev, int mat_dtype, const double* coef_dev, double* resid_dev,
                             int64_t n_rows, void* stream);

/* ---- introspection for the per-layer parity tests: copy a feature map of the LAST micro-batch of the last forward
 * as float32 NCHW (hi+lo recombined). name in {"e11","e12","p1","e21",...,"u3","d31","d32",...}. with_halo=1 returns
 * (B,C,H+2,W+2) including the materialised reflect border. dims_out receives (B,C,H,W) of the returned tensor. */
int wsu_debug_layer(wsu_handle h, const char* name, float* dst_dev, size_t dst_capacity_elems, int with_halo,
                    int64_t* dims_out, void* stream);
/* introspection: "micro_batch" (images per pass of the current plan), "last_images" (images in the last pass that ran),
 * "num_sms", "layers" (kernel launches per pass), "precision" (active plan), "bytes_per_image" (feature-map bytes per image) */
int wsu_get_info(wsu_handle h, const char* key, int64_t* out);
/* per-layer device times (ms) of the last micro-batch when option "profile" is on; returns the layer count (>= 0)
 * or a negative status. wsu_profile_name(i) names entry i ("e11", "e12", ..., "d42"). */
int wsu_profile_read(wsu_handle h, float* ms_out, int cap);
const char* wsu_profile_name(wsu_handle h, int i);
/* number of kernels launched by this library in the calling thread since the last call (bench.py gpu_launches) */
int64_t wsu_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* WSUNET_H_ */
