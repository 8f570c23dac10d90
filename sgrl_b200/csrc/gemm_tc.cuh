// K3 — fp32-accurate tensor-core GEMM for sm_100a: tcgen05.mma (kind::tf32) with the
// accumulator in TMEM, operands streamed by TMA (cp.async.bulk.tensor, SWIZZLE_128B) through
// a 3-4 stage mbarrier ring, error-compensated "3xTF32" split formed in shared memory.
//
//   C[z][m][n] (op)= epi( alpha * sum_k A(m,k) B(n,k) )        (same contract as gemm_simt.cuh)
//
// fp32 parity: every fp32 operand element x is split as x = hi + lo with hi = tf32(x) and
// lo = tf32(x - hi); the tile accumulates hi*hi + lo*hi + hi*lo in the same fp32 TMEM
// accumulator (3 MMAs per k-step), which recovers ~21 mantissa bits (SURVEY.md §7 "fp32
// parity on tensor cores").  The split is elementwise, so it is done in place on the
// TMA-landed (swizzled) tile without knowing the swizzle.
//
// Operand majors: K-major (nn.Linear weights (N,K) and activations (M,K): the forward
// projections) and MN-major (the same buffers read "transposed": dX = dY W and
// dW = dY^T X) are both expressed through the TMA box shape + UMMA descriptor, so no
// transposed copies of weights or activations are ever made.
//
// CTA = 6 warps: warp0 = TMA producer, warp1 = TMEM owner + MMA issuer (one thread),
// warps2-5 = hi/lo converters during the main loop, then the TMEM->register epilogue
// (bias, relu, /F, column scale, relu-mask, residuals, accumulate / split-K atomics).
#pragma once
#include <cuda.h>

#include <unordered_map>

#include "gemm_simt.cuh"

namespace sgrl {

constexpr int TC_BM = 128, TC_BK = 32, TC_THREADS = 192;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;                 // 16 KiB per A tile (hi or lo)

template <int BN> struct TcCfg {
  static constexpr int B_BYTES = BN * TC_BK * 4;
  static constexpr int STAGE_BYTES = 2 * (TC_A_BYTES + B_BYTES);
  static constexpr int STAGES = BN == 128 ? 3 : 4;
  // The tensor core ACCUMULATES WITH TRUNCATION (measured on B200: signed bias -3e-8 per add, i.e.
  // -1.1e-5 relative at K=1024 with one accumulator; profiles/r01_accumulator_probe.txt).  The k-steps are
  // therefore dealt round-robin onto NMAIN independent TMEM accumulators for the hi*hi terms, plus one for
  // the small lo*hi + hi*lo terms (whose truncation is 2^-11 smaller), and summed in fp32 RN in the epilogue.
  static constexpr int NMAIN = BN == 128 ? 3 : 6;
  static constexpr int TMEM_COLS = 512;             // (NMAIN + 1) * BN <= 512
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a pipeline bug traps (error surfaced to the host) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < 20000000u; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// UMMA shared-memory descriptor with the sm_100 version bits (cute/arch/mma_sm100_desc.hpp).
// layout: 2 = SWIZZLE_128B (16 B chunks, 8-row atoms; K-major operands), 1 = SWIZZLE_128B_BASE32B
// (32 B chunks, 4-row atoms) — the only layout the tensor core accepts for MN-major tf32 operands
// (cutlass/gemm/collective/builders/sm100_common.inl); TMA writes it with SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}

// ---------------------------------------------------------------------------- kernel
template <int BN, bool AMN, bool BMN>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                const __grid_constant__ CUtensorMap mapB, GemmP p) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES, B_BYTES = Cfg::B_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto ready_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  const uint32_t acc_bar = bar_base + 8u * (3 * STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, z = blockIdx.z;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int m0 = (blockIdx.x / tiles_n) * TC_BM, n0 = (blockIdx.x % tiles_n) * BN;
  const int nkb = (p.K + TC_BK - 1) / TC_BK;
  const int per = (nkb + p.splitk - 1) / p.splitk;
  const int kb0 = blockIdx.y * per, kb1 = min(nkb, kb0 + per);
  const int nloc = kb1 - kb0;
  const int zA = p.zsA ? z : 0, zB = p.zsB ? z : 0;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(ready_bar(s), 128); mbar_init(empty_bar(s), 1); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nloc > 0) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        for (int i = 0; i < nloc; ++i) {
          const int s = i % STAGES, ph = (i / STAGES) & 1, k0 = (kb0 + i) * TC_BK;
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_arrive_expect_tx(full_bar(s), TC_A_BYTES + B_BYTES);
          const uint32_t a_dst = smem_base + s * STAGE_BYTES, b_dst = a_dst + 2 * TC_A_BYTES;
          if (!AMN) tma_load_3d(a_dst, &mapA, full_bar(s), k0, m0, zA);
          else
            for (int j = 0; j < TC_BM / 32; ++j) tma_load_3d(a_dst + j * 4096, &mapA, full_bar(s), m0 + 32 * j, k0, zA);
          if (!BMN) tma_load_3d(b_dst, &mapB, full_bar(s), k0, n0, zB);
          else
            for (int j = 0; j < BN / 32; ++j) tma_load_3d(b_dst + j * 4096, &mapB, full_bar(s), n0 + 32 * j, k0, zB);
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (one thread) =====================
      if (lane == 0) {
        // instruction descriptor: D=f32, A=B=tf32, majors, N>>3, M>>4 (cute/arch/mma_sm100_desc.hpp InstrDescriptor)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((AMN ? 1u : 0u) << 15) | ((BMN ? 1u : 0u) << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        // K-major tile: 128 B rows (32 k), 8-row swizzle atoms 1024 B apart (SBO); one MMA (K=8) = 32 B along the row.
        // MN-major tile: 32-wide MN chunks as TMA boxes of [32 k-rows x 128 B] 4096 B apart (LBO), 4-row atoms
        // 512 B apart (SBO); one MMA (K=8) = 8 k-rows = 1024 B.
        constexpr uint32_t A_LBO = AMN ? 4096u : 16u, A_SBO = AMN ? 512u : 1024u, A_KSTEP = AMN ? 1024u : 32u, A_LAY = AMN ? 1u : 2u;
        constexpr uint32_t B_LBO = BMN ? 4096u : 16u, B_SBO = BMN ? 512u : 1024u, B_KSTEP = BMN ? 1024u : 32u, B_LAY = BMN ? 1u : 2u;
        for (int i = 0; i < nloc; ++i) {
          const int s = i % STAGES, ph = (i / STAGES) & 1;
          mbar_wait(ready_bar(s), ph);
          tc_fence_after();
          const uint32_t a_hi = smem_base + s * STAGE_BYTES, a_lo = a_hi + TC_A_BYTES;
          const uint32_t b_hi = a_hi + 2 * TC_A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t dah = umma_desc(a_hi + k * A_KSTEP, A_LBO, A_SBO, A_LAY), dal = umma_desc(a_lo + k * A_KSTEP, A_LBO, A_SBO, A_LAY);
            const uint64_t dbh = umma_desc(b_hi + k * B_KSTEP, B_LBO, B_SBO, B_LAY), dbl = umma_desc(b_lo + k * B_KSTEP, B_LBO, B_SBO, B_LAY);
            const int j = i * (TC_BK / 8) + k;                                  // k-step index within this CTA
            const uint32_t acc_lo = tmem_base + Cfg::NMAIN * BN, acc_hi = tmem_base + (j % Cfg::NMAIN) * BN;
            tc_mma_tf32(acc_lo, dal, dbh, idesc, j > 0 ? 1u : 0u);
            tc_mma_tf32(acc_lo, dah, dbl, idesc, 1u);
            tc_mma_tf32(acc_hi, dah, dbh, idesc, j >= Cfg::NMAIN ? 1u : 0u);
          }
          tc_commit(empty_bar(s));          // smem stage reusable once these MMAs retire
        }
        tc_commit(acc_bar);                 // accumulator complete
      }
      __syncwarp();
    } else {
      // ===================== converters: x -> (hi, lo) in shared memory =====================
      const int ct = tid - 64;
      for (int i = 0; i < nloc; ++i) {
        const int s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(full_bar(s), ph);
        uint8_t* st = smem + s * STAGE_BYTES;
#pragma unroll 4
        for (int c = ct; c < (TC_A_BYTES + B_BYTES) / 16; c += 128) {
          const bool isA = c < TC_A_BYTES / 16;
          float4* hi = reinterpret_cast<float4*>(st + (isA ? 0 : 2 * TC_A_BYTES)) + (isA ? c : c - TC_A_BYTES / 16);
          float4* lo = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(hi) + (isA ? TC_A_BYTES : B_BYTES));
          const float4 v = *hi;
          uint4 h, l;
          h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
          l.x = to_tf32(v.x - __uint_as_float(h.x)); l.y = to_tf32(v.y - __uint_as_float(h.y));
          l.z = to_tf32(v.z - __uint_as_float(h.z)); l.w = to_tf32(v.w - __uint_as_float(h.w));
          *reinterpret_cast<uint4*>(hi) = h;
          *reinterpret_cast<uint4*>(lo) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        mbar_arrive(ready_bar(s));
      }
      // ===================== epilogue: TMEM -> registers -> global =====================
      mbar_wait(acc_bar, 0);
      tc_fence_after();
      const int q = warp & 3;                      // TMEM lane quarter this warp may access
      const int m = m0 + q * 32 + lane;
      float* C = p.C + z * p.zsC;
      const float* bias = p.bias ? p.bias + z * p.zsBias : nullptr;
      const float* mask = p.mask ? p.mask + z * p.zsMask : nullptr;
      const float* res1 = p.res1 ? p.res1 + z * p.zsR1 : nullptr;
      const float* res2 = p.res2 ? p.res2 + z * p.zsR2 : nullptr;
      const float rd = (p.rowdiv && m < p.M) ? (p.rowdiv + z * p.zsRow)[m] : 1.f;
      const bool first_split = blockIdx.y == 0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float sum[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[j] = 0.f;
        const int nused = min(Cfg::NMAIN, nloc * (TC_BK / 8));      // accumulators that received at least one k-step
#pragma unroll 1
        for (int a = 0; a <= nused; ++a) {                          // a == nused: the lo accumulator
          uint32_t r[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((a == nused ? Cfg::NMAIN : a) * BN + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(taddr)
              : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(r[j]);
        }
        if (m < p.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + c0 + j;
            if (n < p.N) {
              float v = p.alpha * sum[j];
              if (bias && first_split) v += bias[n];
              if (p.relu) v = fmaxf(v, 0.f);
              if (p.rowdiv) v = v / rd;
              if (n < p.colscale_n) v *= p.colscale;
              if (mask) v = mask[(long long)m * p.ldmask + n] > 0.f ? v : 0.f;
              if (res1) v += res1[(long long)m * p.ldr1 + n];
              if (res2) v += res2[(long long)m * p.ldr2 + n];
              float* c = C + (long long)m * p.ldc + n;
              if (p.splitk > 1) atomicAdd(c, v);
              else if (p.accumulate) *c += v;
              else *c = v;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TmapEncodeFn tmap_encode_fn() {
  static TmapEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TmapEncodeFn>(p);
  }
  return fn;
}

struct TmapKey {
  const void* ptr; long long inner, outer, nz, ld, zs; int box0, box1, swz;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && nz == o.nz && ld == o.ld && zs == o.zs && box0 == o.box0 && box1 == o.box1 &&
           swz == o.swz;
  }
};
struct TmapHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    auto mix = [&](long long v) { h ^= std::hash<long long>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.inner); mix(k.outer); mix(k.nz); mix(k.ld); mix(k.zs); mix(k.box0); mix(k.box1); mix(k.swz);
    return h;
  }
};

// 3-D fp32 tensor map {inner (contiguous), outer (stride ld), z (stride zs)} with a {box0, box1, 1} box, zero OOB fill;
// atom32 = 0: SWIZZLE_128B (K-major operand tiles), 1: SWIZZLE_128B_ATOM_32B (MN-major tf32 operand tiles)
inline int make_tmap(CUtensorMap* out, const float* ptr, long long inner, long long outer, long long nz, long long ld, long long zs,
                     int box0, int box1, int atom32) {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapHash> cache;
  TmapKey key{ptr, inner, outer, nz, ld, zs, box0, box1, atom32};
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return 0; }
  TmapEncodeFn fn = tmap_encode_fn();
  SGRL_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)nz};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(nz > 1 ? zs : ld * outer) * 4};
  cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SGRL_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

inline bool gemm_tc_eligible(const GemmP& p) {
  if (p.M < 1 || p.N < 16 || p.K < 16) return false;
  if ((long long)p.M * p.N * p.K < (1 << 21)) return false;   // tiny: launch-bound either way (per net: Q1 == forward()[0] bit for bit)
  if (!host_vec_ok(p.A, p.lda, p.zsA) || !host_vec_ok(p.B, p.ldb, p.zsB)) return false;   // TMA: 16 B aligned base and strides
  if (p.nb > 1 && ((p.zsA != 0 && p.zsA < 4) || (p.zsB != 0 && p.zsB < 4))) return false;
  return true;
}

template <int BN, bool AMN, bool BMN>
inline int gemm_tc_launch(const GemmP& p, const CUtensorMap& ma, const CUtensorMap& mb, cudaStream_t st) {
  auto kern = gemm_tc_kernel<BN, AMN, BMN>;
  static bool attr_done = false;
  if (!attr_done) {
    SGRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM));
    attr_done = true;
  }
  dim3 grid(ceil_div(p.M, TC_BM) * ceil_div(p.N, BN), p.splitk, p.nb);
  prof_begin(PC_GEMM_TC, 2.0 * p.M * p.N * (double)p.K * p.nb, st);
  kern<<<grid, TC_THREADS, TcCfg<BN>::SMEM, st>>>(ma, mb, p);
  prof_end(st);
  SGRL_LAUNCH_OK();
  return 0;
}

inline int gemm_tc(const GemmP& p_in, cudaStream_t st) {
  GemmP p = p_in;
  if (p.M <= 0 || p.N <= 0 || p.nb <= 0) return 0;
  SGRL_CHECK(gemm_tc_eligible(p), "gemm_tc: operands not TMA-compatible");
  SGRL_CHECK(p.splitk == 1 || (!p.relu && !p.rowdiv && !p.mask && !p.res1 && !p.res2 && p.colscale_n == 0), "gemm_tc: split-K only with a linear epilogue");
  // tile width: 128 unless that leaves most SMs idle
  const long long ctas128 = (long long)ceil_div(p.M, TC_BM) * ceil_div(p.N, 128) * p.nb * p.splitk;
  const int BN = (p.N > 64 && ctas128 >= 100) ? 128 : 64;
  const int nkb = ceil_div(p.K, TC_BK);
  if (p.splitk > nkb) p.splitk = nkb;
  const long long nzA = p.zsA ? p.nb : 1, nzB = p.zsB ? p.nb : 1;
  CUtensorMap ma, mb;
  // K-major operand: inner = K (contiguous), outer = rows; MN-major: inner = rows (contiguous), outer = K
  if (!p.transA) SGRL_TRY(make_tmap(&ma, p.A, p.K, p.M, nzA, p.lda, p.zsA, 32, TC_BM, 0));
  else SGRL_TRY(make_tmap(&ma, p.A, p.M, p.K, nzA, p.lda, p.zsA, 32, 32, 1));
  if (!p.transB) SGRL_TRY(make_tmap(&mb, p.B, p.K, p.N, nzB, p.ldb, p.zsB, 32, BN, 0));
  else SGRL_TRY(make_tmap(&mb, p.B, p.N, p.K, nzB, p.ldb, p.zsB, 32, 32, 1));
#define SGRL_TC_CASE(bn, amn, bmn) return gemm_tc_launch<bn, amn, bmn>(p, ma, mb, st)
  if (BN == 128) {
    if (!p.transA && !p.transB) SGRL_TC_CASE(128, false, false);
    if (!p.transA && p.transB) SGRL_TC_CASE(128, false, true);
    if (p.transA && p.transB) SGRL_TC_CASE(128, true, true);
    SGRL_TC_CASE(128, true, false);
  } else {
    if (!p.transA && !p.transB) SGRL_TC_CASE(64, false, false);
    if (!p.transA && p.transB) SGRL_TC_CASE(64, false, true);
    if (p.transA && p.transB) SGRL_TC_CASE(64, true, true);
    SGRL_TC_CASE(64, true, false);
  }
#undef SGRL_TC_CASE
}

}  // namespace sgrl
