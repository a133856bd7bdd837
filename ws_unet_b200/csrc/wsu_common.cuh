// Shared device/host definitions for the UNet -> WS hot path.
//
// Activation layout in HBM ("split-bf16 NHWC with reflect halo"):
//   every feature map of the reference's UNet.forward (src/unet/model/unet.py:137-189) is stored as TWO bf16
//   planes hi = bf16(v), lo = bf16(v - hi) (hi + lo carries 16 significand bits of the fp32 value), each laid
//   out [B][H+2][W+2][C] with the 1-pixel reflect border of padding_mode='reflect' (unet.py:73) materialised by
//   the producing kernel. A 3x3 tap is then a plain in-bounds TMA box load and the convolution is
//   hi*Whi + lo*Whi + hi*Wlo on the bf16 tensor cores with fp32 accumulation in TMEM.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace wsu {

// Storage formats of a feature map. ACT_SPLIT: two bf16 planes (hi, lo) = 16 significand bits, read by the three-term
// layers. ACT_F16: ONE fp16 plane (11 significand bits), read by the layers the precision plan runs with one or two MMAs per
// MAC (api.cu: "precision" option) - half the HBM bytes and half the TMA box.
// ACT_F16F8: plane 0 = fp16(v); plane 1 = the two CORRECTION operands of the "fp16 + fp8" scheme as e4m3 bytes, per 16-channel
// group g of a 64-channel block the 32 bytes [ e4m3((v - fp16(v)) * kF8ScaleA2) x 16 | e4m3(v * kF8ScaleA1) x 16 ] - one K = 32
// step of a kind::f8f6f4 MMA against the weight bytes [ e4m3(w * pw) x 16 | e4m3((w - fp16(w)) * sw) x 16 ]:
//   a*w ~= a1*w1 (fp16 MMA) + [ a2*w1 + a1*w2 ] (ONE fp8 MMA, accumulated apart and scaled back in the epilogue).
// The correction terms are 2^-11 of the main term, so the 4 significant bits of e4m3 leave ~15 bits overall: the accuracy
// of the three-term bf16 split at two MMA times instead of three, for any weights. Same bytes per element as ACT_SPLIT.
enum : int { ACT_SPLIT = 0, ACT_F16 = 1, ACT_F16F8 = 2 };
constexpr float kF8ScaleA1 = 8.f;        // activations up to 56 stay inside e4m3's range (448); beyond that they saturate
constexpr float kF8ScaleA2 = 16384.f;    // v - fp16(v) <= 2^-6 for v < 64 -> <= 256

struct Act {
  __nv_bfloat16* base;  // plane 0 (hi, or the fp16 plane); plane 1 (lo) starts at base + plane
  size_t plane;         // elements per plane = B*(H+2)*(W+2)*C
  int B, H, W, C;
  int fmt = ACT_SPLIT;
};

__host__ __device__ inline size_t act_plane_elems(int B, int H, int W, int C) {
  return size_t(B) * size_t(H + 2) * size_t(W + 2) * size_t(C);
}

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// pack two floats as hi-parts / lo-parts bf16x2 words (element 0 in the low half).
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo_elem, float hi_elem) {
  uint32_t r;  // one packed round-to-nearest-even conversion: upper half <- hi_elem, lower half <- lo_elem
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi2, uint32_t& lo2) {
  hi2 = cvt_bf16x2(a, b);
  const float ah = __uint_as_float(hi2 << 16);
  const float bh = __uint_as_float(hi2 & 0xffff0000u);
  lo2 = cvt_bf16x2(a - ah, b - bh);  // a - ah is exact in fp32 (Sterbenz-like: ah is a rounded to 8 bits)
}

// two floats -> one fp16x2 word (element 0 in the low half), round to nearest even
__device__ __forceinline__ uint32_t cvt_f16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}

// the same with ReLU folded into the conversion: max(x, 0) commutes with the (monotone) rounding
__device__ __forceinline__ uint32_t cvt_f16x2_relu(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}

// two floats -> two e4m3 bytes (element 0 in the low byte), round to nearest even, saturating to +-448
__device__ __forceinline__ uint32_t cvt_e4m3x2(float lo_elem, float hi_elem) {
  uint16_t r;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
// Packed fp32 pair arithmetic (Blackwell FFMA2): acc.{lo,hi} = fma.rn(v, w.{lo,hi}, acc.{lo,hi}) - two independent IEEE
// fused multiply-adds (bit-identical to two fmaf calls) in one instruction of the FMA pipe.
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2_bcast(float v, uint64_t w, uint64_t acc) {
  uint64_t vv, d;
  asm("mov.b64 %0, {%1, %1};" : "=l"(vv) : "f"(v));   // ptxas folds this into the scalar-broadcast operand form
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(vv), "l"(w), "l"(acc));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t pack_u32x2(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// fp32 = fp16 * fp16 + fp32 in one instruction (FHFMA): element 0 / 1 of the packed halves h2 times element 0 of k2, plus c
__device__ __forceinline__ float fhfma_lo(uint32_t h2, uint32_t k2, float c) {
  float d;
  asm("{ .reg .b16 hl, hh, kl, kh; mov.b32 {hl, hh}, %1; mov.b32 {kl, kh}, %2; fma.rn.f32.f16 %0, hl, kl, %3; }"
      : "=f"(d) : "r"(h2), "r"(k2), "f"(c));
  return d;
}
__device__ __forceinline__ float fhfma_hi(uint32_t h2, uint32_t k2, float c) {
  float d;
  asm("{ .reg .b16 hl, hh, kl, kh; mov.b32 {hl, hh}, %1; mov.b32 {kl, kh}, %2; fma.rn.f32.f16 %0, hh, kl, %3; }"
      : "=f"(d) : "r"(h2), "r"(k2), "f"(c));
  return d;
}

// 32 channels of one pixel (two 16-channel groups) in the ACT_F16F8 layout: h = 16 fp16x2 words (plane 0), l = the 64 bytes
// [a2s 0..15 | a1q 0..15 | a2s 16..31 | a1q 16..31] of plane 1. Packed f32x2 arithmetic: (v - h) 2^14 = fma(h, -2^14, v 2^14)
// exactly (v - h is representable and the scales are powers of two), so the bytes equal those of the scalar formulation.
__device__ __forceinline__ void pack_f16f8_32(const float (&f)[32], uint32_t (&h)[16], uint32_t (&l)[16]) {
  static_assert(kF8ScaleA2 == 16384.f, "kNegA2h below is -2^14 as an fp16 bit pattern");
  const uint64_t kA2 = pack_f32x2(kF8ScaleA2, kF8ScaleA2);
  const uint64_t kA1 = pack_f32x2(kF8ScaleA1, kF8ScaleA1);
  const uint32_t kNegA2h = 0xF400F400u;   // (-16384, -16384) as packed halves
#pragma unroll
  for (int g = 0; g < 2; ++g) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {     // word i of a 16-byte run holds channels 16g + 4i .. 16g + 4i + 3
      const int c = 16 * g + 4 * i;
      uint32_t a2[2], a1[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint64_t v = pack_f32x2(f[c + 2 * j], f[c + 2 * j + 1]);
        const uint32_t hw = cvt_f16x2(f[c + 2 * j], f[c + 2 * j + 1]);
        h[8 * g + 2 * i + j] = hw;
        float e0, e1, s0, s1;
        unpack_f32x2(mul2(v, kA2), e0, e1);
        const float r0 = fhfma_lo(hw, kNegA2h, e0), r1 = fhfma_hi(hw, kNegA2h, e1);   // (v - h) 2^14, the halves read in place
        unpack_f32x2(mul2(v, kA1), s0, s1);
        a2[j] = cvt_e4m3x2(r0, r1);
        a1[j] = cvt_e4m3x2(s0, s1);
      }
      l[8 * g + i] = a2[0] | (a2[1] << 16);
      l[8 * g + 4 + i] = a1[0] | (a1[1] << 16);
    }
  }
}

// Reflect-halo targets of a logical coordinate v in [0,n): storage index v+1, plus the mirrored border
// rows/cols it also owns (index 0 mirrors logical 1, index n+1 mirrors logical n-2).
__device__ __forceinline__ int halo_targets(int v, int n, int (&t)[3]) {
  int k = 0;
  t[k++] = v + 1;
  if (v == 1) t[k++] = 0;
  if (v == n - 2) t[k++] = n + 1;
  return k;
}

// WS estimator modes (src/ws/estimate.py:55-136)
enum : int { WS_UNWEIGHTED = 0, WS_WEIGHTED = 1, WS_ANTIWEIGHTED = -1 };
// linear predictors (src/ws/estimate.py:31-52, src/filters/evaluate.py:29-50)
enum : int { PRED_KB = 0, PRED_AVG = 1, PRED_AVG9 = 2, PRED_ID = 3 };

// number of float slots per partial-sum record: sum(w*r), sum(w), sum|x-xhat|, sum(w*d*xbias)
constexpr int kPartialSlots = 4;

}  // namespace wsu
