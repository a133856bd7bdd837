// Hardware probe (not product code): sustained cycles per K=16 step for the operand pattern of the row kernel
// (conv_row_kernel): cta_group::2, M = 256 over the pair, N = 192 = the three dy-taps of one dx stacked along N,
// A = a 130-pixel row segment (8-row atoms contiguous, SBO 1024) whose descriptor start is shifted by dx rows.
// Per-CTA floors: tensor N/2 = 96 cycles per MMA; shared-memory reads (4 KB A + 96 rows x 32 B of this CTA's half of B)
// at 128 B per cycle. Also prints the alignment of the dynamic shared-memory base (the row kernel's d41 budget
// leaves less than 1 KB of slack).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../ws_unet_b200/csrc/ptx.cuh"
using namespace wsu;

constexpr int kPlane = 130 * 128;        // one plane of a row segment
constexpr int kBox = 2 * kPlane;         // hi + lo
constexpr int kWTile = 96 * 128;         // this CTA's half of one stacked (dx, plane) weight tile
constexpr int kNPat = 6;

template <int PAT>
__device__ __forceinline__ void issue(uint32_t a0, uint32_t w, uint32_t d0) {
  constexpr uint32_t id192 = make_idesc_bf16_m(256, 192);
  for (int rep = 0; rep < 24; ++rep) {
    for (int dx = 0; dx < 3; ++dx) {
      const uint32_t whi = w + uint32_t(dx) * 2 * kWTile, wlo = whi + kWTile;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ah = make_sw128_desc(a0 + dx * 128 + k * 32), al = make_sw128_desc(a0 + kPlane + dx * 128 + k * 32);
        const uint64_t bh = make_sw128_desc(whi + k * 32), bl = make_sw128_desc(wlo + k * 32);
        if (PAT == 0) {          // row kernel: A_hi kept in the collector for its two products
          umma_bf16_2sm_a_fill(d0, ah, bh, id192, 1); umma_bf16_2sm_a_lastuse(d0, ah, bl, id192, 1); umma_bf16_2sm(d0, al, bh, id192, 1);
        } else if (PAT == 1) {   // same without the collector
          umma_bf16_2sm(d0, ah, bh, id192, 1); umma_bf16_2sm(d0, ah, bl, id192, 1); umma_bf16_2sm(d0, al, bh, id192, 1);
        } else if (PAT == 2) {   // single N = 192
          umma_bf16_2sm(d0, ah, bh, id192, 1);
        } else if (PAT == 3) {   // order hi*hi, lo*hi, hi*lo (no reuse possible)
          umma_bf16_2sm(d0, ah, bh, id192, 1); umma_bf16_2sm(d0, al, bh, id192, 1); umma_bf16_2sm(d0, ah, bl, id192, 1);
        } else if (PAT == 4) {   // two accumulator stages alternating per K step (stage switch costs nothing?)
          const uint32_t d = d0 + ((k & 1) ? 192u : 0u);
          umma_bf16_2sm_a_fill(d, ah, bh, id192, 1); umma_bf16_2sm_a_lastuse(d, ah, bl, id192, 1); umma_bf16_2sm(d, al, bh, id192, 1);
        } else {                 // N = 192 split as 128 + 64 (what a wrapped accumulator ring would need)
          umma_bf16_2sm(d0, ah, bh, make_idesc_bf16_m(256, 128), 1); umma_bf16_2sm(d0 + 128, ah, bl, make_idesc_bf16_m(256, 64), 1);
        }
      }
    }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate(long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + 34 * 1024;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 6 * kWTile);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const bool leader = cluster_ctarank() == 0;
  for (int i = threadIdx.x; i < (34 * 1024 + 6 * kWTile) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc_2sm(slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (leader && threadIdx.x == 0) out[kNPat] = (long long)(smem_u32(raw) & 1023u);
  uint32_t phase = 0;
  for (int pat = 0; pat < kNPat; ++pat) {
    long long t0 = 0;
    if (leader && threadIdx.x < 32) {
      if (elect_one()) {
        const uint32_t a0 = smem_u32(sA), w = smem_u32(sB);
        t0 = clock64();
        switch (pat) {
#define CASE(i) case i: issue<i>(a0, w, tmem); break;
          CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5)
#undef CASE
        }
        umma_commit_2sm(bar, 3);
      }
      __syncwarp();
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    if (t0) out[pat] = clock64() - t0;
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
  }
  if (threadIdx.x < 32) tmem_dealloc_2sm(tmem, 512);
}

int main() {
  long long* dout;
  cudaMalloc(&dout, (kNPat + 1) * sizeof(long long));
  cudaMemset(dout, 0, (kNPat + 1) * sizeof(long long));
  const int smem = 34 * 1024 + 6 * kWTile + 1024 + 64;
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int it = 0; it < 2; ++it) rate<<<2, 128, smem>>>(dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[kNPat + 1];
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  struct Info { const char* name; int tensor; int smem_bytes; };
  const Info info[kNPat] = {{"pair 3 x N192, A_hi in collector", 288, 4096 + 3072 + 3072 + 4096 + 3072},
                            {"pair 3 x N192, no collector", 288, 3 * (4096 + 3072)},
                            {"pair N192", 96, 4096 + 3072},
                            {"pair 3 x N192, hi*hi lo*hi hi*lo", 288, 3 * (4096 + 3072)},
                            {"pair 3 x N192 collector, 2 stages", 288, 4096 + 3072 + 3072 + 4096 + 3072},
                            {"pair N128 + N64 (split window)", 96, 4096 + 2048 + 4096 + 1024}};
  printf("dynamic shared memory base & 1023 = %lld\n", h[kNPat]);
  printf("%-40s %10s %10s %10s\n", "pattern (per K=16 step, per CTA)", "cycles", "tensor", "smem-read");
  for (int i = 0; i < kNPat; ++i)
    printf("%-40s %10.1f %10d %10.1f\n", info[i].name, double(h[i]) / (24 * 12), info[i].tensor, info[i].smem_bytes / 128.0);
  return 0;
}
