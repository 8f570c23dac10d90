// Microbenchmark: issue rate of tcgen05.mma kind::tf32 (M=128, K=8) on sm_100a as a function of the tile width N, the
// number of TMEM accumulators the MMAs rotate over, and where the A operand lives (shared memory "SS" or TMEM "TS").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe tools/mma_probe.cu ; run: ./mma_probe
// Used to pick the tile shape / issue order of csrc/gemm_tc.cuh (results: profiles/r01b_mma_probe.txt).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) {   // K-major SWIZZLE_128B, SBO 1024
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// MODE 0: SS, MODE 1: TS (A in TMEM columns 256..).  Warp 1 issues in warp-uniform control flow, one elected lane,
// groups of 4 k-steps (32 B apart inside a 128 B swizzled row) per stage like the real kernel; ROT accumulators.
template <int BN, int ROT, int MODE>
__global__ void __launch_bounds__(128, 1) probe(int ngroups, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int STAGE = 16384 + BN * 128;
  for (int i = tid; i < (4 * STAGE) / 4; i += 128) ((float*)smem)[i] = 1.0f + (i & 7) * 0.125f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = __shfl_sync(0xffffffffu, slot, 0);
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (warp == 1) {
    const uint32_t sb = smem_u32(smem);
    const uint64_t a0 = desc_k(sb), b0 = desc_k(sb + 16384);
    long long t0 = clock64();
    for (int g = 0; g < ngroups; ++g) {
      const int stage = g & 3;
      const uint64_t ad = a0 + (uint64_t)((stage * STAGE) >> 4), bd = b0 + (uint64_t)((stage * STAGE) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t d = tb + ((k % ROT) * BN);
          const uint32_t acc = (g > 0 || k >= ROT) ? 1u : 0u;
          if (MODE == 0) mma_ss(d, ad + 2 * k, bd + 2 * k, idesc, acc);
          else mma_ts(d, tb + 256 + k * 8, bd + 2 * k, idesc, acc);
        }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    long long t2 = clock64();
    if (blockIdx.x == 0 && tid == 32) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
  }
}

template <int BN, int ROT, int MODE>
void run(int ngroups, int grid, long long* d) {
  const int smem = 4 * (16384 + BN * 128) + 1024;
  cudaFuncSetAttribute(probe<BN, ROT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) probe<BN, ROT, MODE><<<grid, 128, smem>>>(ngroups, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const double floor_cyc = 128.0 * BN / 256.0;
  printf("N=%3d rot=%d %s grid=%3d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (pipe floor %5.1f) %s\n", BN, ROT, MODE ? "TS" : "SS", grid,
         (double)h[0] / (4 * ngroups), (double)h[1] / (4 * ngroups), floor_cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const int ng = 256;
  for (int grid : {1, 148}) {
    run<64, 1, 0>(ng, grid, d); run<64, 2, 0>(ng, grid, d); run<64, 4, 0>(ng, grid, d);
    run<128, 1, 0>(ng, grid, d); run<128, 2, 0>(ng, grid, d); run<128, 4, 0>(ng, grid, d);
    run<256, 1, 0>(ng, grid, d); run<256, 2, 0>(ng, grid, d);
    run<64, 1, 1>(ng, grid, d); run<64, 4, 1>(ng, grid, d);
    run<128, 1, 1>(ng, grid, d); run<128, 2, 1>(ng, grid, d);
    run<256, 1, 1>(ng, grid, d);
  }
  return 0;
}
