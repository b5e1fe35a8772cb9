// Issue-rate microbenchmark of tcgen05.mma.kind::i8 (M = 128, N = 64 / 128 / 256, K = 32) and kind::tf32 (K = 8) from resident shared-memory tiles:
// cycles per MMA at saturation, one CTA per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/i8_rate tools/i8_rate.cu
#include <cstdio>
#include "../approximategps.jl_b200/csrc/i8emu.cuh"
using namespace agp; using namespace agp::i8e; using namespace agp::t5;
__device__ long long g_cyc[8];
template <int N, bool TF32>
__global__ void __launch_bounds__(64, 1) rate_kernel(int iters, int slot) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 2 * 32768, tslot = bar + 16;
  volatile uint32_t* tp = reinterpret_cast<volatile uint32_t*>(raw + (tslot - smem_u32(raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += 64) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x01010101u;
  if (threadIdx.x == 0) { mbar_init_u32(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tslot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tp;
  if (warp == 1) {
    const uint32_t leader = elect_one();
    const uint32_t idesc = TF32 ? ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24)) : ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24));
    const uint64_t da = kdesc(base), db = kdesc(base + 32768);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 16; u++) {
        if (TF32) {
          asm volatile("{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nsetp.ne.b32 q, %5, 0;\n@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + (uint32_t)((u & 1) * N)), "l"(da + (uint64_t)(2 * (u & 3))), "l"(db + (uint64_t)(2 * (u & 3))), "r"(idesc), "r"(1u), "r"(leader) : "memory");
        } else {
          tc_mma_i8(tmem + (uint32_t)((u & 1) * N), da + (uint64_t)(2 * (u & 3)), db + (uint64_t)(2 * (u & 3)), idesc, 1u, leader);
        }
      }
    }
    tc_commit_pred(bar, leader);
    mbar_wait_u32(bar, 0);
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) g_cyc[slot] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory"); }
}
template <int N, bool TF32> void run(int slot, const char* name) {
  auto k = rate_kernel<N, TF32>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  const int iters = 256;
  k<<<148, 64, 70000>>>(iters, slot);
  cudaDeviceSynchronize();
  long long c[8]; cudaMemcpyFromSymbol(c, g_cyc, sizeof c);
  const double per = (double)c[slot] / (iters * 16.0);
  const double macs = 128.0 * N * (TF32 ? 8 : 32);
  printf("{\"mma\": \"%s\", \"N\": %d, \"cycles_per_mma\": %.1f, \"mac_per_clk_per_sm\": %.0f, \"chip_tops_at_1965MHz\": %.0f, \"err\": \"%s\"}\n", name, N, per, macs / per, 2 * macs / per * 148 * 1.965e9 / 1e12, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  run<64, false>(0, "i8"); run<128, false>(1, "i8"); run<256, false>(2, "i8");
  run<64, true>(3, "tf32"); run<128, true>(4, "tf32"); run<256, true>(5, "tf32");
  return 0;
}
