// Host launchers of the CUDA-core kernels in stencil.cu.
#pragma once
#include <cuda_runtime.h>
#include "wsu_common.cuh"

namespace wsu {

// img_kind: 0 uint8 pixels, 1 float32 in [0,1], 2 uint8 pixels read as the LSB-difference image (x_bar - x) / 255
cudaError_t launch_first_conv(const void* img, int img_kind, int cin, const float* w, const float* bias, Act out,
                              cudaStream_t stream);
int filter_ws_strips(int H);
// register sliding-window fast path (uint8, KB/AVG, no bias term, no x_hat output, W % 4 == 0)
bool filter_ws_fast_ok(const void* img, int img_is_float, int W, int kind, int want_bias, const float* xhat_out);
int filter_ws_fast_records(int H, int W);  // partial records per image written by the fast kernel
cudaError_t launch_filter_ws_fast(const void* img, int B, int H, int W, int kind, int weighted, float* partials,
                                  cudaStream_t stream);
// packed 16-bit-lane variant of the fast path: unweighted, beta_hat only (no l1)
int filter_ws_packed_records(int H, int W);
cudaError_t launch_filter_ws_packed(const void* img, int B, int H, int W, int kind, float* partials, cudaStream_t stream);
// adjoint variant (stencil applied to the parity plane, dp4a against the pixels): unweighted, beta_hat only, W % 16 == 0
bool filter_ws_adjoint_ok(const void* img, int H, int W);
int filter_ws_adjoint_records(int H, int W);
cudaError_t launch_filter_ws_adjoint(const void* img, int B, int H, int W, int kind, float* partials, cudaStream_t stream);
// window (dp4a) variant: weighted / anti-weighted and/or L1-reporting KB/AVG estimator on uint8 images, W % 16 == 0
bool filter_ws_window_ok(const void* img, int H, int W);
int filter_ws_window_records(int H, int W);
cudaError_t launch_filter_ws_window(const void* img, int B, int H, int W, int kind, int weighted, int want_l1, float* partials,
                                    cudaStream_t stream);
cudaError_t launch_filter_ws(const void* img, int img_is_float, int B, int H, int W, int kind, int weighted, int want_bias,
                             float* xhat_out, float* partials, cudaStream_t stream);
cudaError_t launch_ws_from_pred(const void* img, int img_is_float, const float* xhat, int xhat_cropped, const float* xbias,
                                int B, int H, int W, int weighted, int crop, float* partials, int chunks,
                                cudaStream_t stream);
cudaError_t launch_ws_grad_pred(const void* img, int img_is_float, const float* coef, float* grad, int B, int H, int W,
                                int crop, float scale, cudaStream_t stream);
cudaError_t launch_finalize(const float* partials, int records, int B, float npix, int clip, int correct_bias,
                            float* beta_hat, float* l1, cudaStream_t stream);
cudaError_t launch_kb_blend(const void* x, int x_is_float, const float* mask, float* out, int B, int C, int H, int W,
                            unsigned chan_mask, cudaStream_t stream);
// mat_dtype: 0 uint8, 1 float32, 2 float64
cudaError_t launch_residual_matvec(const void* mat, int mat_dtype, const double* coef, double* resid, long long n,
                                   cudaStream_t stream);
cudaError_t launch_pack(const float* src, Act dst, cudaStream_t stream);
cudaError_t launch_unpack(Act src, float* dst, int with_halo, cudaStream_t stream);

}  // namespace wsu
