// Hardware probe (not product code): does a tcgen05 SWIZZLE_128B K-major A-operand descriptor read correctly when
//   (a) its start address is offset by a number of 128-byte rows that is not a multiple of 8, and
//   (b) the stride between 8-row atoms (SBO) is not a multiple of 1024 bytes (haloed 10-pixel-wide rows)?
// If yes, one haloed activation box in shared memory can feed all 9 taps of a 3x3 convolution.
// Shared memory is filled in the layout TMA produces: 16-byte chunk index XOR (absolute row & 7).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../ws_unet_b200/csrc/ptx.cuh"
using namespace wsu;

constexpr int ROWS = 256;  // rows of 128 B in the A staging area
struct Variant { int start_row; int sbo_bytes; int base_mode; };  // base_mode: 0 -> base_offset 0, 1 -> (addr>>7)&7

__device__ __forceinline__ float aval(int R, int k) { return float(((R * 7 + k * 3) % 17) - 8); }
__device__ __forceinline__ float bval(int n, int k) { return float(((n * 5 + k * 11) % 13) - 6); }

__global__ void __launch_bounds__(128, 1) probe(const Variant* vars, int nvar, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sB = reinterpret_cast<__nv_bfloat16*>(smem + ROWS * 128);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ROWS * 128 + 64 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < ROWS * 64; i += 128) {
    const int R = i / 64, k = i % 64;
    sA[(R * 128 + (((k >> 3) ^ (R & 7)) << 4) + (k & 7) * 2) / 2] = __float2bfloat16(aval(R, k));
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    sB[(n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) / 2] = __float2bfloat16(bval(n, k));
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 64);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  uint32_t phase = 0;
  for (int v = 0; v < nvar; ++v) {
    if (threadIdx.x == 0) {
      const uint32_t a0 = smem_u32(sA) + vars[v].start_row * 128;
      const uint32_t b0 = smem_u32(sB);
      for (int k = 0; k < 4; ++k) {
        uint64_t da = make_sw128_desc(a0 + k * 32);
        da &= ~(uint64_t(0x3FFF) << 32);
        da |= uint64_t(vars[v].sbo_bytes >> 4) << 32;
        if (vars[v].base_mode == 1) da |= uint64_t((a0 >> 7) & 7) << 49;
        umma_bf16(tmem, da, make_sw128_desc(b0 + k * 32), make_idesc_bf16(64), k != 0);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t r[32];
      tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + cc * 32, r);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) out[(size_t(v) * 128 + warp * 32 + lane) * 64 + cc * 32 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (threadIdx.x < 32) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<Variant> vars;
  // dense rows (SBO 1024) with start offsets 0..3, 8, 9 rows; haloed 10-wide rows (SBO 1280) with (dy*10+dx) offsets
  for (int bm = 0; bm < 2; ++bm) {
    for (int s : {0, 1, 2, 3, 8, 9, 16, 17, 18}) vars.push_back({s, 1024, bm});
    for (int dy = 0; dy < 3; ++dy)
      for (int dx = 0; dx < 3; ++dx) vars.push_back({dy * 10 + dx, 1280, bm});
    for (int s : {0, 1, 18, 19, 36, 37, 38}) vars.push_back({s, 18 * 128, bm});  // 16-wide rows + 2 halo: atoms of 8 inside an 18-row pitch? (expected to fail: 2 atoms per row)
  }
  const int nvar = int(vars.size());
  Variant* dv; float* dout;
  cudaMalloc(&dv, nvar * sizeof(Variant));
  cudaMemcpy(dv, vars.data(), nvar * sizeof(Variant), cudaMemcpyHostToDevice);
  cudaMalloc(&dout, size_t(nvar) * 128 * 64 * 4);
  const int smem = ROWS * 128 + 64 * 128 + 1024 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(dv, nvar, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> out(size_t(nvar) * 128 * 64);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  auto av = [](int R, int k) { return float(((R * 7 + k * 3) % 17) - 8); };
  auto bv = [](int n, int k) { return float(((n * 5 + k * 11) % 13) - 6); };
  for (int v = 0; v < nvar; ++v) {
    int bad = 0;
    const int pitch_rows = vars[v].sbo_bytes / 128;  // smem rows between consecutive 8-row atoms
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        const int R = vars[v].start_row + (m / 8) * pitch_rows + (m % 8);
        float ref = 0;
        for (int k = 0; k < 64; ++k) ref += av(R, k) * bv(n, k);
        if (out[(size_t(v) * 128 + m) * 64 + n] != ref) ++bad;
      }
    printf("start_row=%2d sbo=%4d base_mode=%d : %s (%d / 8192 wrong)\n", vars[v].start_row, vars[v].sbo_bytes, vars[v].base_mode,
           bad ? "MISMATCH" : "ok", bad);
  }
  return 0;
}
