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
  fflush(stdout);
}


// ---- second probe: the TS-mode k-block loop of csrc/gemm_tc.cuh with its side activities switched on one by one
// flags: 1 = per-block mbarrier waits + tcgen05.fence::after, 2 = two tcgen05.commit per block, 4 = 8 converter warps
// free-running (8 LDS.128 + 2 tcgen05.st.x32 + wait::st per block each), 8 = a producer warp streaming 32 KB per block
// from global into shared memory with cp.async.bulk, 16 = converters only do the LDS part, 32 = only the TMEM stores
__device__ __forceinline__ void tmem_st32z(uint32_t taddr, uint32_t v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(v) : "memory");
}
template <int BN>
__global__ void __launch_bounds__(320, 1) probe2(int nblocks, int flags, const float* gsrc, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bdone, bsink, btma;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int STAGE = 16384 + 2 * BN * 128;
  for (int i = tid; i < (4 * STAGE) / 4; i += 320) ((float*)smem)[i] = 1.0f + (i & 7) * 0.125f;
  if (tid == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bdone)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bsink)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&btma)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = __shfl_sync(0xffffffffu, slot, 0);
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (tid == 0) {   // pre-complete `bar` phase 0 so that waits on parity 0 succeed immediately
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();
  if (warp == 1) {
    const uint32_t sb = smem_u32(smem);
    const uint64_t b0 = desc_k(sb + 16384);
    long long t0 = clock64();
    for (int g = 0; g < nblocks; ++g) {
      const int stage = g & 3, t = g & 3;
      if (flags & 1) {
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const uint64_t bh = b0 + (uint64_t)((stage * STAGE) >> 4), bl = bh + (uint64_t)((BN * 128) >> 4);
      const uint32_t ah = tb + 256 + t * 64, al = ah + 32;
      if (elect_one()) {
        const uint32_t am = tb + (g % 3) * BN, alo = tb + 3 * BN, f0 = (g >= 3) ? 1u : 0u, l0 = (g > 0) ? 1u : 0u;
        if ((flags & 192) == 0) {            // order 0: hh x4, then (lo,hi),(hi,lo) interleaved per k
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ts(am, ah + 8 * k, bh + 2 * k, idesc, k > 0 ? 1u : f0);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            mma_ts(alo, al + 8 * k, bh + 2 * k, idesc, k > 0 ? 1u : l0);
            mma_ts(alo, ah + 8 * k, bl + 2 * k, idesc, 1u);
          }
        } else if (flags & 64) {             // order 1: hh x4, (lo,hi) x4, (hi,lo) x4
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ts(am, ah + 8 * k, bh + 2 * k, idesc, k > 0 ? 1u : f0);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ts(alo, al + 8 * k, bh + 2 * k, idesc, k > 0 ? 1u : l0);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ts(alo, ah + 8 * k, bl + 2 * k, idesc, 1u);
        } else {                             // order 2: per k: (hi,hi), (lo,hi), (hi,lo)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            mma_ts(am, ah + 8 * k, bh + 2 * k, idesc, k > 0 ? 1u : f0);
            mma_ts(alo, al + 8 * k, bh + 2 * k, idesc, k > 0 ? 1u : l0);
            mma_ts(alo, ah + 8 * k, bl + 2 * k, idesc, 1u);
          }
        }
        if (flags & 2) {
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bsink)) : "memory");
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bsink)) : "memory");
        }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bdone)) : "memory");
    __syncwarp();
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bdone)) : "memory");
    long long t2 = clock64();
    stop = 1;
    if (blockIdx.x == 0 && tid == 32) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp >= 2 && (flags & (4 | 16 | 32))) {
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t sb = smem_u32(smem);
    uint32_t acc = 0;
    int it = 0;
    for (int guard = 0; guard < 200000 && !__shfl_sync(0xffffffffu, (int)stop, 0); ++guard) {
      const uint32_t st = sb + (it & 3) * STAGE;
      if (flags & (4 | 16)) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(st + row * 128 + ((c ^ (row & 7)) << 4)) : "memory");
          acc += __float_as_uint(v.x) ^ __float_as_uint(v.w);
        }
      }
      if (flags & (4 | 32)) {
        // TMEM columns 256 + 64*(it&3): the same slots the MMAs read (values are irrelevant for timing)
        tmem_st32z(tb + ((uint32_t)(q * 32) << 16) + 256 + (it & 3) * 64, acc);
        tmem_st32z(tb + ((uint32_t)(q * 32) << 16) + 256 + (it & 3) * 64 + 32, acc + 1);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      ++it;
    }
    if (acc == 0x12345678u) out[3] = acc;
  } else if (warp == 0 && (flags & 8)) {
    int it = 0;
    for (int guard = 0; guard < 200000 && !__shfl_sync(0xffffffffu, (int)stop, 0); ++guard) {
      if (elect_one()) {
        const uint32_t dst = smem_u32(smem) + (it & 3) * STAGE;
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(&btma)), "r"(32768u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(gsrc + (size_t)blockIdx.x * 65536 + (it & 7) * 8192), "r"(32768u), "r"(smem_u32(&btma)) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&btma)), "r"((uint32_t)(it & 1)) : "memory");
      }
      __syncwarp();
      ++it;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
  }
}

template <int BN>
void run2(int nblocks, int flags, int grid, const float* gsrc, long long* d) {
  const int smem = 4 * (16384 + 2 * BN * 128) + 1024;
  cudaFuncSetAttribute(probe2<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) probe2<BN><<<grid, 320, smem>>>(nblocks, flags, gsrc, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("probe2 N=%3d flags=%2d grid=%3d: %7.1f cyc per 12-MMA k-block (issue %7.1f; pipe floor %5.1f) %s\n", BN, flags, grid,
         (double)h[1] / nblocks, (double)h[0] / nblocks, 12 * 128.0 * BN / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  fflush(stdout);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  float* gsrc;
  cudaMalloc(&gsrc, (size_t)148 * 65536 * 4);
  cudaMemset(gsrc, 0, (size_t)148 * 65536 * 4);
  const int ng = 256;
  for (int grid : {1, 148}) {
    run<64, 1, 0>(ng, grid, d); run<64, 4, 0>(ng, grid, d);
    run<128, 1, 0>(ng, grid, d); run<256, 1, 0>(ng, grid, d);
    run<64, 1, 1>(ng, grid, d); run<128, 1, 1>(ng, grid, d);
  }
  for (int grid : {1, 148})
    for (int flags : {0, 64, 128, 15, 15 + 64, 15 + 128}) { run2<64>(ng, flags, grid, gsrc, d); }
  for (int flags : {0, 64, 128, 15, 15 + 64, 15 + 128}) run2<128>(ng, flags, 148, gsrc, d);
  return 0;
}
