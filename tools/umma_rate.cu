// Hardware probe (not product code): sustained cycles per tcgen05.mma for the operand patterns of the halo kernels
// (M = 128, K = 16, bf16, operands in shared memory through SW128 descriptors with the haloed 1280-byte atom stride).
// One CTA, one issuing thread; each pattern issues 8 x 36 bodies, commits, waits, and reports cycles per body next to
// the tensor floor (N/2 cycles per MMA) and the shared-memory read floor ((4 KB + 32 N) / 128 B per cycle).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../ws_unet_b200/csrc/ptx.cuh"
using namespace wsu;

#define UMMA_COL(NAME, SUFFIX)                                                                                    \
  __device__ __forceinline__ void NAME(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {         \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                                 \
                 "tcgen05.mma.cta_group::1.kind::f16" SUFFIX " [%0], %1, %2, %3, p;\n\t}" ::"r"(d),                \
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)                                                              \
                 : "memory");                                                                                      \
  }
UMMA_COL(umma_fill, ".collector::a::fill")
UMMA_COL(umma_use, ".collector::a::use")
UMMA_COL(umma_lastuse, ".collector::a::lastuse")
UMMA_COL(umma_discard, ".collector::a::discard")
#define UMMA_WS(NAME, SUFFIX)                                                                                     \
  __device__ __forceinline__ void NAME(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {         \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                                 \
                 "tcgen05.mma.ws.cta_group::1.kind::f16" SUFFIX " [%0], %1, %2, %3, p;\n\t}" ::"r"(d),             \
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)                                                              \
                 : "memory");                                                                                      \
  }
UMMA_WS(ws_b0_fill, ".collector::b0::fill")
UMMA_WS(ws_b0_use, ".collector::b0::use")
UMMA_WS(ws_b0_last, ".collector::b0::lastuse")
UMMA_WS(ws_b1_fill, ".collector::b1::fill")
UMMA_WS(ws_b1_last, ".collector::b1::lastuse")
UMMA_WS(ws_plain, "")

__device__ __forceinline__ void dummy_() {}
constexpr int kBox = 2 * 180 * 128;   // hi + lo planes of one haloed box
constexpr int kNPat = 18;

template <int PAT>
__device__ __forceinline__ void issue(uint32_t a0, uint32_t a1, uint32_t lo, uint32_t w, uint32_t w2, uint32_t d0, uint32_t d1,
                                      uint32_t d2) {
        for (int rep = 0; rep < 8; ++rep) {
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t off = uint32_t((tap / 3) * 10 + tap % 3) * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ah0 = make_sw128_desc(a0 + off + k * 32, 1280), al0 = make_sw128_desc(a0 + lo + off + k * 32, 1280);
              const uint64_t ah1 = make_sw128_desc(a1 + off + k * 32, 1280), al1 = make_sw128_desc(a1 + lo + off + k * 32, 1280);
              const uint64_t bw = make_sw128_desc(w + k * 32), bw2 = make_sw128_desc(w2 + k * 32);
              switch (PAT) {
                case 0:  // stacked, as the Cout=64 kernel issues it
                  umma_bf16(d0, ah0, bw, make_idesc_bf16(128), 1); umma_bf16(d0, al0, bw, make_idesc_bf16(64), 1); break;
                case 1:  // stacked, two boxes interleaved
                  umma_bf16(d0, ah0, bw, make_idesc_bf16(128), 1); umma_bf16(d1, ah1, bw, make_idesc_bf16(128), 1);
                  umma_bf16(d0, al0, bw, make_idesc_bf16(64), 1); umma_bf16(d1, al1, bw, make_idesc_bf16(64), 1); break;
                case 2:  // stacked, lo term into other columns (no accumulator dependency)
                  umma_bf16(d0, ah0, bw, make_idesc_bf16(128), 1); umma_bf16(d2, al0, bw, make_idesc_bf16(64), 1); break;
                case 3:  // three N=64
                  umma_bf16(d0, ah0, bw, make_idesc_bf16(64), 1); umma_bf16(d0, al0, bw, make_idesc_bf16(64), 1);
                  umma_bf16(d0, ah0, bw2, make_idesc_bf16(64), 1); break;
                case 4:  // three N=128, as the Cout>=128 kernel issues it
                  umma_bf16(d0, ah0, bw, make_idesc_bf16(128), 1); umma_bf16(d0, al0, bw, make_idesc_bf16(128), 1);
                  umma_bf16(d0, ah0, bw2, make_idesc_bf16(128), 1); break;
                case 5:  // three N=128, A_hi kept in the collector for its second use
                  umma_fill(d0, ah0, bw, make_idesc_bf16(128), 1); umma_lastuse(d0, ah0, bw2, make_idesc_bf16(128), 1);
                  umma_bf16(d0, al0, bw, make_idesc_bf16(128), 1); break;
                case 6:  // N=128 only
                  umma_bf16(d0, ah0, bw, make_idesc_bf16(128), 1); break;
                case 7:  // N=64 only
                  umma_bf16(d0, ah0, bw, make_idesc_bf16(64), 1); break;
                case 8:  // N=256 only
                  umma_bf16(d0, ah0, bw, make_idesc_bf16(256), 1); break;
                case 9:  // N=64 alternating accumulators
                  umma_bf16(d0, ah0, bw, make_idesc_bf16(64), 1); umma_bf16(d1, al0, bw, make_idesc_bf16(64), 1); break;
                case 10:  // three N=64 with collector reuse of A_hi
                  umma_fill(d0, ah0, bw, make_idesc_bf16(64), 1); umma_lastuse(d0, ah0, bw2, make_idesc_bf16(64), 1);
                  umma_bf16(d0, al0, bw, make_idesc_bf16(64), 1); break;
                case 11:  // N=128 same A four times through the collector (upper bound of the collector effect)
                  umma_fill(d0, ah0, bw, make_idesc_bf16(128), 1); umma_use(d0, ah0, bw2, make_idesc_bf16(128), 1);
                  umma_use(d0, ah0, bw, make_idesc_bf16(128), 1); umma_lastuse(d0, ah0, bw2, make_idesc_bf16(128), 1); break;
                case 12:  // N=128 same A four times without the collector
                  umma_bf16(d0, ah0, bw, make_idesc_bf16(128), 1); umma_bf16(d0, ah0, bw2, make_idesc_bf16(128), 1);
                  umma_bf16(d0, ah0, bw, make_idesc_bf16(128), 1); umma_bf16(d0, ah0, bw2, make_idesc_bf16(128), 1); break;
                case 13:  // N=192: hi*[Whi;Wlo;?] style wide tile
                  umma_bf16(d0, ah0, bw, make_idesc_bf16(192), 1); umma_bf16(d0, al0, bw, make_idesc_bf16(64), 1); break;
              }
            }
          }
        }
}

__global__ void __launch_bounds__(128, 1) rate(long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // two boxes
  uint8_t* sB = smem + 2 * kBox;            // 256 rows x 128 B (N up to 256), second copy after it
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 2 * 256 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (2 * kBox + 2 * 256 * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  uint32_t phase = 0;
  for (int pat = 0; pat < kNPat; ++pat) {
    long long t0 = 0, t1 = 0;
    if (threadIdx.x < 32) {
      if (elect_one()) {
        const uint32_t a0 = smem_u32(sA), a1 = a0 + kBox, lo = 180 * 128;
        const uint32_t w = smem_u32(sB), w2 = w + 256 * 128;
        const uint32_t d0 = tmem, d1 = tmem + 128, d2 = tmem + 256;
        t0 = clock64();
        switch (pat) {
#define CASE(i) case i: issue<i>(a0, a1, lo, w, w2, d0, d1, d2); break;
          CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16) CASE(17)
#undef CASE
        }
        umma_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    if (threadIdx.x < 32 && t0) { t1 = clock64(); out[pat] = t1 - t0; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* dout;
  cudaMalloc(&dout, kNPat * sizeof(long long));
  cudaMemset(dout, 0, kNPat * sizeof(long long));
  const int smem = 2 * kBox + 2 * 256 * 128 + 1024 + 64;
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int it = 0; it < 2; ++it) rate<<<1, 128, smem>>>(dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[kNPat];
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  struct Info { const char* name; int tensor; int smem_bytes; };
  const Info info[kNPat] = {
      {"stacked N128+N64 (same acc)", 96, 14336}, {"stacked, 2 boxes interleaved", 192, 28672},
      {"stacked, lo term other columns", 96, 14336}, {"3 x N64", 96, 18432}, {"3 x N128", 192, 24576},
      {"3 x N128, A_hi via collector", 192, 20480}, {"N128", 64, 8192}, {"N64", 32, 6144}, {"N256", 128, 12288},
      {"N64 alternating accumulators", 64, 12288}, {"3 x N64, A_hi via collector", 96, 14336},
      {"4 x N128 same A, collector", 256, 20480}, {"4 x N128 same A, no collector", 256, 32768},
      {"N192 + N64", 128, 16384}, {".ws stacked 2 boxes, B collectors", 192, 16384 + 6144},
      {".ws 4 x N128, one B", 256, 16384 + 4096}, {".ws stacked, no reuse", 96, 14336}, {".ws 2 x N64, one B", 64, 8192 + 2048}};
  printf("%-34s %10s %10s %10s\n", "pattern (per K=16 step)", "cycles", "tensor", "smem-read");
  for (int i = 0; i < kNPat; ++i)
    printf("%-34s %10.1f %10d %10.1f\n", info[i].name, double(h[i]) / (8 * 36), info[i].tensor, info[i].smem_bytes / 128.0);
  return 0;
}
