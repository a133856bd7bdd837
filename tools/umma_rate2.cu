// Hardware probe (not product code): sustained cycles per tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, K = 16,
// bf16) for the operand patterns of conv_halo2_kernel. Cluster of two CTAs, the leader issues; per-CTA floors:
// tensor N/2 cycles per MMA, shared-memory reads (4 KB A + 16 N bytes of this CTA's half of B) / 128 B per cycle.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../ws_unet_b200/csrc/ptx.cuh"
using namespace wsu;

constexpr int kBox = 2 * 180 * 128;
constexpr int kNPat = 8;

template <int PAT>
__device__ __forceinline__ void issue(uint32_t a0, uint32_t a1, uint32_t lo, uint32_t w, uint32_t w2, uint32_t d0, uint32_t d1) {
  for (int rep = 0; rep < 8; ++rep) {
    for (int tap = 0; tap < 9; ++tap) {
      const uint32_t off = uint32_t((tap / 3) * 10 + tap % 3) * 128;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ah0 = make_sw128_desc(a0 + off + k * 32, 1280), al0 = make_sw128_desc(a0 + lo + off + k * 32, 1280);
        const uint64_t ah1 = make_sw128_desc(a1 + off + k * 32, 1280), al1 = make_sw128_desc(a1 + lo + off + k * 32, 1280);
        const uint64_t bw = make_sw128_desc(w + k * 32), bw2 = make_sw128_desc(w2 + k * 32);
        if (PAT == 0) {  // stacked as conv_halo2_kernel<64> issues it
          umma_bf16_2sm(d0, ah0, bw, make_idesc_bf16_m(256, 128), 1); umma_bf16_2sm(d0, al0, bw2, make_idesc_bf16_m(256, 64), 1);
        } else if (PAT == 1) {  // three N=128
          umma_bf16_2sm(d0, ah0, bw, make_idesc_bf16_m(256, 128), 1); umma_bf16_2sm(d0, al0, bw, make_idesc_bf16_m(256, 128), 1);
          umma_bf16_2sm(d0, ah0, bw2, make_idesc_bf16_m(256, 128), 1);
        } else if (PAT == 2) {  // N=64 only
          umma_bf16_2sm(d0, ah0, bw, make_idesc_bf16_m(256, 64), 1);
        } else if (PAT == 3) {  // N=128 only
          umma_bf16_2sm(d0, ah0, bw, make_idesc_bf16_m(256, 128), 1);
        } else if (PAT == 4) {  // N=256 only
          umma_bf16_2sm(d0, ah0, bw, make_idesc_bf16_m(256, 256), 1);
        } else if (PAT == 5) {  // 4-term stacked: both A planes against [Whi; Wlo]
          umma_bf16_2sm(d0, ah0, bw, make_idesc_bf16_m(256, 128), 1); umma_bf16_2sm(d0, al0, bw, make_idesc_bf16_m(256, 128), 1);
        } else if (PAT == 6) {  // stacked, two box slots interleaved
          umma_bf16_2sm(d0, ah0, bw, make_idesc_bf16_m(256, 128), 1); umma_bf16_2sm(d1, ah1, bw, make_idesc_bf16_m(256, 128), 1);
          umma_bf16_2sm(d0, al0, bw2, make_idesc_bf16_m(256, 64), 1); umma_bf16_2sm(d1, al1, bw2, make_idesc_bf16_m(256, 64), 1);
        } else {  // three N=64
          umma_bf16_2sm(d0, ah0, bw, make_idesc_bf16_m(256, 64), 1); umma_bf16_2sm(d0, al0, bw, make_idesc_bf16_m(256, 64), 1);
          umma_bf16_2sm(d0, ah0, bw2, make_idesc_bf16_m(256, 64), 1);
        }
      }
    }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate(long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * kBox;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 2 * 128 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const bool leader = cluster_ctarank() == 0;
  for (int i = threadIdx.x; i < (2 * kBox + 2 * 128 * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc_2sm(slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  uint32_t phase = 0;
  for (int pat = 0; pat < kNPat; ++pat) {
    long long t0 = 0;
    if (leader && threadIdx.x < 32) {
      if (elect_one()) {
        const uint32_t a0 = smem_u32(sA), a1 = a0 + kBox, lo = 180 * 128;
        const uint32_t w = smem_u32(sB), w2 = w + 128 * 128;
        const uint32_t d0 = tmem, d1 = tmem + 256;
        t0 = clock64();
        switch (pat) {
#define CASE(i) case i: issue<i>(a0, a1, lo, w, w2, d0, d1); break;
          CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7)
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
  cudaMalloc(&dout, kNPat * sizeof(long long));
  cudaMemset(dout, 0, kNPat * sizeof(long long));
  const int smem = 2 * kBox + 2 * 128 * 128 + 1024 + 64;
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int it = 0; it < 2; ++it) rate<<<2, 128, smem>>>(dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[kNPat];
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  struct Info { const char* name; int tensor; int smem_bytes; };
  const Info info[kNPat] = {{"pair stacked N128+N64", 96, 8192 + 2048 + 1024}, {"pair 3 x N128", 192, 12288 + 6144},
                            {"pair N64", 32, 4096 + 1024}, {"pair N128", 64, 4096 + 2048}, {"pair N256", 128, 4096 + 4096},
                            {"pair 2 x N128 (4-term stacked)", 128, 8192 + 4096}, {"pair stacked, 2 slots interleaved", 192, 2 * 11264},
                            {"pair 3 x N64", 96, 12288 + 3072}};
  printf("%-36s %10s %10s %10s\n", "pattern (per K=16 step, per CTA)", "cycles", "tensor", "smem-read");
  for (int i = 0; i < kNPat; ++i)
    printf("%-36s %10.1f %10d %10.1f\n", info[i].name, double(h[i]) / (8 * 36), info[i].tensor, info[i].smem_bytes / 128.0);
  return 0;
}
