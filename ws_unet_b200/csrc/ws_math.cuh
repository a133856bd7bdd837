// Per-pixel Weighted-Stego arithmetic shared by the fused UNet head epilogue and the linear-filter
// estimator kernel. Mirrors src/ws/estimate.py:83-121 (attack), src/unet/evaluate.py:128-132 (predict_unet)
// and src/_defs/losses.py:46-61 (WSLoss._error) of the reference:
//   x_bar = x ^ 1 (LSB flip on the integer pixel), residual (x - x_bar)(x - x_hat),
//   weights 1/(5+var) | 5+var | 1 with var = AVG(x^2) - AVG(x)^2 over the 8 neighbours.
#pragma once
#include "wsu_common.cuh"

namespace wsu {

struct WsAcc {
  float wr = 0.f;  // sum w * (x - x_bar) * (x - x_hat)
  float w = 0.f;   // sum w
  float l1 = 0.f;  // sum |x - x_hat|
  float wb = 0.f;  // sum w * (x - x_bar) * x_bias   (bias correction, estimate.py:126-128)
};

// pixel value and its LSB-flipped twin. uint8 input is exact; float input in [0,1] follows
// WSLoss: x*255, round-half-even, int ^ 1 (losses.py:48-50).
__device__ __forceinline__ void ws_load_u8(uint8_t p, float& xv, float& xbar) {
  xv = float(p);
  xbar = float(p ^ 1);
}
__device__ __forceinline__ void ws_load_f32(float p01, float& xv, float& xbar) {
  xv = p01 * 255.f;
  xbar = float(__float2int_rn(xv) ^ 1);
}

// local-variance weight from the 8-neighbour sums S1 = sum x, S2 = sum x^2 (estimate.py:94-103).
__device__ __forceinline__ float ws_weight(int mode, float s1, float s2) {
  if (mode == WS_UNWEIGHTED) return 1.f;
  const float mu = s1 * 0.125f;
  const float mu2 = s2 * 0.125f;
  const float var = __fsub_rn(mu2, __fmul_rn(mu, mu));
  return mode == WS_WEIGHTED ? __fdiv_rn(1.f, 5.f + var) : 5.f + var;
}

__device__ __forceinline__ void ws_accumulate(WsAcc& a, float xv, float xbar, float xhat, float wgt) {
  const float d = xv - xbar;
  const float e = xv - xhat;
  a.wr = fmaf(wgt * d, e, a.wr);
  a.w += wgt;
  a.l1 += fabsf(e);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace wsu
