// K3 — fp32-accurate tensor-core GEMM for sm_100a: tcgen05.mma (kind::tf32) with the accumulators in TMEM, operands
// streamed by TMA (cp.async.bulk.tensor) through an mbarrier ring, error-compensated "3xTF32" split.
//
//   C[z][m][n] (op)= epi( alpha * sum_k A(m,k) B(n,k) )        (same contract as gemm_simt.cuh)
//
// fp32 parity: every fp32 operand element x is split as x = hi + lo (hi = the tf32 the tensor core reads out of the raw
// word, lo = tf32(x - hi)); a tile accumulates hi*hi + lo*hi + hi*lo (3 MMAs per k-step, ~21 mantissa bits), the hi*hi
// terms round-robin over up to three accumulators because the tensor core accumulates with truncation (TcCfg).
//
// Operand majors: K-major (nn.Linear weights (N,K) and activations (M,K): the forward projections) and MN-major (the
// same buffers read "transposed": dX = dY W and dW = dY^T X) are both expressed through the TMA box shape + UMMA
// descriptor, so no transposed copies of weights or activations are ever made.
//
// CTA = 10 warps: warp0 = TMA producer, warp1 = TMEM owner + MMA issuer (one elected lane), warps 2-9 = two converter
// groups that move each landed A k-block into TMEM as [hi | lo] (the A operand is fed from tensor memory, ".ts" form)
// and then run the fused epilogue (TMEM -> registers -> shared transpose -> coalesced global: bias, relu, /F, column
// scale, relu-mask, residuals, accumulate / split-K atomics).  Weights arrive pre-split (hi / lo arenas).
// Variants: SM2 = two CTAs per SM for short-K tiles of inference passes; cluster split-K (p.csk) = a (1, 2..3, 1) cluster
// shares one tile's k-loop and reduces through distributed shared memory.  DESIGN.md §4 has the measurements behind each.
#pragma once
#include <cuda.h>

#include <cstdlib>

#include <unordered_map>

#include "gemm_simt.cuh"

namespace sgrl {

constexpr int TC_BM = 128, TC_BK = 32;
constexpr int TC_CONV_WARPS = 8;                              // converter warps (also the epilogue warps)
constexpr int TC_CONV_THREADS = 32 * TC_CONV_WARPS;
constexpr int TC_THREADS = 64 + TC_CONV_THREADS;              // warp0 TMA, warp1 MMA, warps 2.. converters/epilogue
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;                 // 16 KiB per A tile (hi or lo)

// SM2 = the two-CTAs-per-SM variant for short-K tiles (<= 12 k-blocks) of launches with several tiles per SM: half the
// shared memory (2 ring stages) and half the tensor memory (ONE accumulator for all three product terms + 2 A slots =
// 256 columns), so that one CTA's prologue / pipeline fill / epilogue (TMEM read-back and the output burst) overlaps
// the other CTA's MMAs on the same SM.  Measured motivation (tools/gemm_trace.py, 147456x768x256): 15.1 K cycles per
// 128x128 tile of which 6.1 K are MMA time; 2.4 K startup and 4.8 K epilogue run with the tensor pipe idle.
template <int BN, bool SM2 = false> struct TcCfg {
  static constexpr int B_BYTES = BN * TC_BK * 4;                    // one B tile (hi or lo)
  static constexpr int STAGE_BYTES = TC_A_BYTES + 2 * B_BYTES;      // [A raw | B raw/hi | B lo]
  static constexpr int STAGES = SM2 ? 2 : (BN == 128 ? 4 : 6);      // 192 KiB of operand ring (96 / 64 KiB for SM2)
  // The tensor core ACCUMULATES WITH TRUNCATION (measured on B200: signed bias -3e-8 per add, i.e.
  // -1.1e-5 relative at K=1024 with one accumulator; profiles/r01_accumulator_probe.txt).  The k-blocks are
  // therefore dealt round-robin onto NMAIN independent TMEM accumulators for the hi*hi terms, plus one for
  // the small lo*hi + hi*lo terms (whose truncation is 2^-11 smaller), and summed in fp32 RN in the epilogue.
  static constexpr int NMAIN = SM2 ? 1 : (BN == 128 ? 2 : 3);
  static constexpr int ACC_COLS = SM2 ? BN : (NMAIN + 1) * BN;      // 384 | 256 (SM2: the cross terms share the one accumulator:
                                                                    // <= 144 truncating adds, ~4e-6 relative, measured in tests/test_gemm_gpu.py)
  // The A operand is fed from TENSOR MEMORY (tcgen05.mma "ts" form): with both operands in shared memory the
  // 128 B/cycle shared-memory path bounds the mainloop (MMA operand reads + TMA writes + converter traffic =
  // 136 KB per 12-MMA k-block at BN=64: 1100 cycles measured against a 384-cycle tensor-pipe floor;
  // profiles/r01b_mma_probe.txt).  TA_STAGES slots of 64 columns hold [hi 32 | lo 32] of one k-block.
  static constexpr int TA_STAGES = SM2 ? 2 : (512 - ACC_COLS) / 64; // 2 | 4
  static constexpr int TMEM_COLS = SM2 ? 256 : 512;
  static constexpr int UNROLL = SM2 ? 2 : (BN == 128 ? 4 : 12);     // lcm(STAGES, TA_STAGES, NMAIN)
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
// descriptors passed as {lo word, hi word}; scale_d (accumulate) as a runtime flag
__device__ __forceinline__ void tc_mma_tf32_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
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

// explicit shared-space vector access by 32-bit shared address: the tile base comes out of integer alignment
// arithmetic, so the compiler cannot prove the address space and would emit slow generic LD/ST (measured:
// ~850 cycles per dependent generic load in the epilogue, 7 us per 128x64 tile)
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---------------------------------------------------------------------------- kernel
// lo part of the error-compensated split: the tensor core reads an fp32 word as tf32 by IGNORING its 13 low mantissa
// bits (measured, profiles/r01b_mma_probe.txt and the parity tests), so the raw word already is hi = trunc(x); the
// subtraction x - trunc(x) is exact and the result is rounded to tf32 (round half away, 2 integer ops).
__device__ __forceinline__ uint32_t lo_of_trunc(float x) {
  const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  return (__float_as_uint(x - hi) + 0x1000u) & 0xFFFFE000u;
}
// lo parts of NCH 16-byte chunks of a TMA-landed (swizzled) B tile, written LO_OFF bytes further: elementwise, so the
// swizzled placement is preserved without knowing the swizzle.  NT threads cooperate.
template <int NCH, int LO_OFF, int NT>
__device__ __forceinline__ void split_tile_lo(uint32_t tile, int ct) {
  static_assert(NCH % NT == 0, "chunks per converter thread");
  constexpr int PER = NCH / NT;
  float4 v[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) v[i] = lds128(tile + (uint32_t)(ct + i * NT) * 16u);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    uint4 l;
    l.x = lo_of_trunc(v[i].x); l.y = lo_of_trunc(v[i].y); l.z = lo_of_trunc(v[i].z); l.w = lo_of_trunc(v[i].w);
    sts128u(tile + LO_OFF + (uint32_t)(ct + i * NT) * 16u, l);
  }
}

__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}

#define TC_R32(r) "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), \
  "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),          \
  "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
// registers -> TMEM: lane l of the warp writes r[0..31] to TMEM lane (32*(warp%4) + l), columns [col, col+32)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), TC_R32(r)
      : "memory");
}
// A operand in TMEM (".ts" form): [d] (+)= [a_tmem] x b_desc
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
}

constexpr int TC_EPI_LD = 36;     // floats per staged accumulator row (32 + 4: 16 B aligned, conflict-free float4 access)

enum { EM_STORE = 0, EM_ACCUM = 1, EM_ATOMIC = 2 };
struct EpiArgs {
  float* C; const float* bias; const float* mask; const float* res1; const float* res2;   // lane-offset row pointers (chunk 0)
  float rd[8];        // rowdiv of this lane's 8 rows (1 when absent)
  int mrow, rows, vec;
};

// One 32-column chunk of one warp's 32-row slab: staged fp32 accumulators (shared) -> fused epilogue -> global, each
// warp store covering four 128-byte row segments.  Specialised at compile time on the epilogue class so that the
// common cases cost a few instructions per float4 (the first version tested every option per element and spent
// 7 us per 128x64 tile in dependent integer/branch latency).
template <bool ROWDIV, bool MASK, int NRES, int MODE>
__device__ __forceinline__ void tc_epi_chunk(const GemmP& p, const EpiArgs& ea, uint32_t stg, int rsub, int cq, int n, int c0) {
  if (n >= p.N) return;
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool full4 = ea.vec && (n + 3 < p.N);
  if (MODE == EM_STORE && ea.bias) {
    if (full4) b4 = ldg4(ea.bias + n);
    else {
      b4.x = __ldg(ea.bias + n);
      if (n + 1 < p.N) b4.y = __ldg(ea.bias + n + 1);
      if (n + 2 < p.N) b4.z = __ldg(ea.bias + n + 2);
      if (n + 3 < p.N) b4.w = __ldg(ea.bias + n + 3);
    }
  }
  const float cs = (ROWDIV && n < p.colscale_n) ? p.colscale : 1.f;
  const float alpha = p.alpha;
  const bool relu = MODE == EM_STORE && p.relu;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + rsub;
    if (r >= ea.rows) break;
    float4 v = lds128(stg + (r * TC_EPI_LD + cq * 4) * 4);
    v.x = fmaf(alpha, v.x, b4.x); v.y = fmaf(alpha, v.y, b4.y); v.z = fmaf(alpha, v.z, b4.z); v.w = fmaf(alpha, v.w, b4.w);
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (ROWDIV) {
      const float rd = ea.rd[it];
      v.x = v.x / rd * cs; v.y = v.y / rd * cs; v.z = v.z / rd * cs; v.w = v.w / rd * cs;
    }
    const long long roff = (long long)(4 * it) * p.ldc + c0;
    float* c = ea.C + roff;
    if (full4) {
      if (MASK) {
        const float4 t = *reinterpret_cast<const float4*>(ea.mask + (long long)(4 * it) * p.ldmask + c0);
        v.x = t.x > 0.f ? v.x : 0.f; v.y = t.y > 0.f ? v.y : 0.f; v.z = t.z > 0.f ? v.z : 0.f; v.w = t.w > 0.f ? v.w : 0.f;
      }
      if (NRES >= 1) {
        const float4 t = *reinterpret_cast<const float4*>(ea.res1 + (long long)(4 * it) * p.ldr1 + c0);
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        if (ea.res2) {
          const float4 u = *reinterpret_cast<const float4*>(ea.res2 + (long long)(4 * it) * p.ldr2 + c0);
          v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
      }
      if (MODE == EM_ATOMIC) {
        atomicAdd(c, v.x); atomicAdd(c + 1, v.y); atomicAdd(c + 2, v.z); atomicAdd(c + 3, v.w);
      } else {
        if (MODE == EM_ACCUM) { const float4 t = *reinterpret_cast<const float4*>(c); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
        *reinterpret_cast<float4*>(c) = v;
      }
    } else {
      float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (n + e >= p.N) break;
        if (MASK) o[e] = ea.mask[(long long)(4 * it) * p.ldmask + c0 + e] > 0.f ? o[e] : 0.f;
        if (NRES >= 1) {
          o[e] += ea.res1[(long long)(4 * it) * p.ldr1 + c0 + e];
          if (ea.res2) o[e] += ea.res2[(long long)(4 * it) * p.ldr2 + c0 + e];
        }
        if (MODE == EM_ATOMIC) atomicAdd(c + e, o[e]);
        else if (MODE == EM_ACCUM) c[e] += o[e];
        else c[e] = o[e];
      }
    }
  }
}

// any other combination of epilogue options (same op order as gemm_simt_kernel), element by element
__device__ __noinline__ void tc_epi_generic(const GemmP& p, const EpiArgs& ea, uint32_t stg, int rsub, int cq, int n, int c0) {
  if (n >= p.N) return;
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + rsub;
    if (r >= ea.rows) break;
    const float4 a4 = lds128(stg + (r * TC_EPI_LD + cq * 4) * 4);
    const float av[4] = {a4.x, a4.y, a4.z, a4.w};
    for (int e = 0; e < 4 && n + e < p.N; ++e) {
      float v = p.alpha * av[e];
      if (ea.bias) v += __ldg(ea.bias + n + e);
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.rowdiv) v = v / ea.rd[it];
      if (n + e < p.colscale_n) v *= p.colscale;
      if (ea.mask) v = ea.mask[(long long)(4 * it) * p.ldmask + c0 + e] > 0.f ? v : 0.f;
      if (ea.res1) v += ea.res1[(long long)(4 * it) * p.ldr1 + c0 + e];
      if (ea.res2) v += ea.res2[(long long)(4 * it) * p.ldr2 + c0 + e];
      float* c = ea.C + (long long)(4 * it) * p.ldc + c0 + e;
      if (p.splitk > 1 && !p.csk) atomicAdd(c, v);
      else if (p.accumulate) *c += v;
      else *c = v;
    }
  }
}

// BPRE: the B operand arrives already split (raw/hi + lo tensor maps over the weight arenas); otherwise the converter
// warps also write B's lo part next to the landed tile.
template <int BN, bool AMN, bool BMN, bool BPRE, bool SM2 = false>
__global__ void __launch_bounds__(TC_THREADS, SM2 ? 2 : 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                          const __grid_constant__ CUtensorMap mapB,
                                                                          const __grid_constant__ CUtensorMap mapBlo, GemmP p) {
  using Cfg = TcCfg<BN, SM2>;
  constexpr int STAGES = Cfg::STAGES, B_BYTES = Cfg::B_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES, TA = Cfg::TA_STAGES, NMAIN = Cfg::NMAIN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * TA + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                       // TMA landed stage s (A raw + B)
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };           // MMAs reading stage s retired
  auto ta_ready = [&](int t) { return bar_base + 8u * (2 * STAGES + t); };        // A hi|lo of a k-block stored in TMEM slot t
  auto ta_empty = [&](int t) { return bar_base + 8u * (2 * STAGES + TA + t); };   // MMAs reading TMEM slot t retired
  const uint32_t acc_bar = bar_base + 8u * (2 * STAGES + 2 * TA);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, z = blockIdx.z;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int m0 = (blockIdx.x / tiles_n) * TC_BM, n0 = (blockIdx.x % tiles_n) * BN;
  const int nkb = (p.K + TC_BK - 1) / TC_BK;
  const int per = (nkb + p.splitk - 1) / p.splitk;
  const int kb0 = blockIdx.y * per, kb1 = min(nkb, kb0 + per);
  const int nloc = kb1 - kb0;
  const int zA = p.zsA ? z : 0, zB = p.zsB ? z : 0;
  // accumulators in rotation: at most ~12 k-blocks (48 truncating hi*hi accumulations) per accumulator — the level the
  // K = 1024 projections run at with all NMAIN — so short-K tiles use (and later read back from TMEM) fewer of them
  const int nrot = min(Cfg::NMAIN, (nloc + 11) / 12);
  long long* dbg = (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && z == 0) ? p.dbg : nullptr;
#define TC_STAMP(slot) do { if (dbg) dbg[slot] = clock64(); } while (0)
  if (tid == 0) TC_STAMP(0);
  SGRL_PDL_TRIGGER();      // the next kernel of the stream may begin its own prologue on idle SMs

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int t = 0; t < TA; ++t) { mbar_init(ta_ready(t), TC_CONV_WARPS / 2); mbar_init(ta_empty(t), 1); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    if (BPRE) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBlo) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t ta_base = tmem_base + Cfg::ACC_COLS;          // TA slots of 64 columns: [raw/hi 32 | lo 32]
  SGRL_PDL_WAIT();         // everything above overlapped the previous kernel's tail; its outputs are visible from here on
  if (tid == 0) TC_STAMP(1);

  if (nloc > 0) {
    if (warp == 0) {
      // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
      int s = 0; uint32_t ph = 0;
      for (int i = 0; i < nloc; ++i) {
        const int k0 = (kb0 + i) * TC_BK;
        mbar_wait(empty_bar(s), ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(s), TC_A_BYTES + (BPRE ? 2 : 1) * B_BYTES);
          const uint32_t a_dst = smem_base + s * STAGE_BYTES, b_dst = a_dst + TC_A_BYTES;
          if (!AMN) tma_load_3d(a_dst, &mapA, full_bar(s), k0, m0, zA);          // [128 rows][32 k], SWIZZLE_128B
          else tma_load_3d(a_dst, &mapA, full_bar(s), m0, k0, zA);               // [32 k][128 rows], unswizzled
          if (!BMN) {
            tma_load_3d(b_dst, &mapB, full_bar(s), k0, n0, zB);
            if (BPRE) tma_load_3d(b_dst + B_BYTES, &mapBlo, full_bar(s), k0, n0, zB);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) tma_load_3d(b_dst + j * 4096, &mapB, full_bar(s), n0 + 32 * j, k0, zB);
            if (BPRE) {
#pragma unroll
              for (int j = 0; j < BN / 32; ++j) tma_load_3d(b_dst + B_BYTES + j * 4096, &mapBlo, full_bar(s), n0 + 32 * j, k0, zB);
            }
          }
          if (i < 12) TC_STAMP(8 + i);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      // The whole warp walks the pipeline in warp-uniform control flow (loop state and descriptors stay in uniform
      // registers); one elected lane issues the MMAs and commits of a k-block.
      // instruction descriptor: D=f32, A=B=tf32, A K-major (TMEM), B major, N>>3, M>>4 (cute/arch/mma_sm100_desc.hpp)
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((BMN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(TC_BM >> 4) << 24);
      // B K-major tile: 128 B rows (32 k), 8-row swizzle atoms 1024 B apart (SBO); one MMA (K=8) = 32 B along the row.
      // B MN-major tile: 32-wide MN chunks as TMA boxes of [32 k-rows x 128 B] 4096 B apart (LBO), 4-row atoms
      // 512 B apart (SBO); one MMA (K=8) = 8 k-rows = 1024 B.
      constexpr uint32_t B_LBO = BMN ? 4096u : 16u, B_SBO = BMN ? 512u : 1024u, B_KSTEP = BMN ? 1024u : 32u, B_LAY = BMN ? 1u : 2u;
      // 64-bit descriptor = {lo word: addr>>4 | (LBO>>4)<<16, hi word: SBO>>4 | version 1<<14 | layout<<29}
      constexpr uint32_t B_HIW = (B_SBO >> 4) | (1u << 14) | (B_LAY << 29), B_LOW = (B_LBO >> 4) << 16;
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t acc_lo = SM2 ? tb : tb + NMAIN * BN, tab = tb + Cfg::ACC_COLS;
      const uint32_t bdesc0 = (((smem_base + TC_A_BYTES) >> 4) & 0x3FFFu) | B_LOW;
      // Measured (tools/mma_probe.cu, profiles/r01b_mma_probe.txt): the tensor pipe drains during ANY gap in the issue
      // stream (queue of ~1-2 MMAs), so every instruction between two MMAs is exposed.  The loop is therefore unrolled
      // over UNR = lcm(STAGES, TA_STAGES, NMAIN) k-blocks (ring slots, TMEM slots, accumulators and phase parities fold
      // to constants) and the barrier waits of block i+1 sit between the 8th and 9th MMA of block i.
      constexpr int UNR = Cfg::UNROLL;
      static_assert(UNR % STAGES == 0 && UNR % TA == 0 && UNR % NMAIN == 0, "unroll must cover whole ring turns");
      mbar_wait(full_bar(0), 0);
      mbar_wait(ta_ready(0), 0);
      tc_fence_after();
      for (int it = 0, i0 = 0; i0 < nloc; ++it, i0 += UNR) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int i = i0 + u;
          if (i < nloc) {                                                  // warp-uniform
            const int s = u % STAGES, t = u % TA;
            const uint32_t b_hi = bdesc0 + (uint32_t)(s * (STAGE_BYTES >> 4)), b_lo = b_hi + (B_BYTES >> 4);
            const uint32_t a_hi = tab + t * 64, a_lo = a_hi + 32;
            const int ai = nrot == NMAIN ? (u % NMAIN) : (nrot == 1 ? 0 : (u & 1));   // UNR is even: (i0 + u) & 1 == u & 1
            const uint32_t acc_hi = tb + ai * BN;                          // main accumulators rotate per k-block (see TcCfg)
            const uint32_t first_hi = (it > 0 || u >= nrot) ? 1u : 0u, first_lo = (SM2 || it > 0 || u > 0) ? 1u : 0u;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < TC_BK / 8; ++k)
                tc_mma_tf32_ts(acc_hi, a_hi + 8 * k, b_hi + k * (B_KSTEP >> 4), B_HIW, idesc, k > 0 ? 1u : first_hi);
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                tc_mma_tf32_ts(acc_lo, a_lo + 8 * k, b_hi + k * (B_KSTEP >> 4), B_HIW, idesc, k > 0 ? 1u : first_lo);
                tc_mma_tf32_ts(acc_lo, a_hi + 8 * k, b_lo + k * (B_KSTEP >> 4), B_HIW, idesc, 1u);
              }
            }
            __syncwarp();
            if (i + 1 < nloc) {                                            // operands of the next block (hidden behind queued MMAs)
              const int un = (u + 1) % UNR, itn = it + (u + 1 == UNR ? 1 : 0);
              mbar_wait(full_bar(un % STAGES), (uint32_t)((UNR / STAGES) * itn + un / STAGES) & 1u);
              mbar_wait(ta_ready(un % TA), (uint32_t)((UNR / TA) * itn + un / TA) & 1u);
              tc_fence_after();
            }
            if (elect_one()) {
#pragma unroll
              for (int k = 2; k < TC_BK / 8; ++k) {
                tc_mma_tf32_ts(acc_lo, a_lo + 8 * k, b_hi + k * (B_KSTEP >> 4), B_HIW, idesc, 1u);
                tc_mma_tf32_ts(acc_lo, a_hi + 8 * k, b_lo + k * (B_KSTEP >> 4), B_HIW, idesc, 1u);
              }
              tc_commit(empty_bar(s));                     // smem stage reusable once these MMAs retire
              tc_commit(ta_empty(t));                      // so is the TMEM A slot
              if (i == nloc - 1) tc_commit(acc_bar);       // accumulators complete
            }
            __syncwarp();
          }
        }
      }
    } else {
      // ===================== converters: smem A tile -> (hi, lo) rows in TMEM; B lo in smem when not pre-split =====================
      // two groups of 4 warps take alternate k-blocks; thread = one tile row = one TMEM lane
      const int g = (warp - 2) >> 2, q = warp & 3, row = q * 32 + lane, cgt = ((warp - 2) & 3) * 32 + lane;
      const uint32_t trow = ta_base + ((uint32_t)(q * 32) << 16);
      static_assert(STAGES % 2 == 0 && TA % 2 == 0, "two converter groups take alternate ring slots");
      int s = g, t = g; uint32_t ph = 0, pht = 0;
      // bias gradient of a weight-gradient GEMM (A = dY^T, row m = output feature): every A element passes through this
      // thread's registers anyway, so the row sums cost 32 adds per k-block instead of a column-sum kernel per tensor
      const bool do_rs = p.rowsum != nullptr && n0 == 0;
      float rs = 0.f;
      for (int i = g; i < nloc; i += 2) {
        mbar_wait(full_bar(s), ph);
        if (tid == 64 && i < 12) TC_STAMP(20 + i);
        const uint32_t st = smem_base + s * STAGE_BYTES;
        uint32_t raw[32], lo[32];
        if (!AMN) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = lds128(st + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4));
            raw[4 * c] = __float_as_uint(v.x); raw[4 * c + 1] = __float_as_uint(v.y);
            raw[4 * c + 2] = __float_as_uint(v.z); raw[4 * c + 3] = __float_as_uint(v.w);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) raw[k] = __float_as_uint(lds32(st + (uint32_t)k * 512u + (uint32_t)row * 4u));
        }
        if (!BPRE) split_tile_lo<B_BYTES / 16, B_BYTES, 128>(st + TC_A_BYTES, cgt);
        if (do_rs) {
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; ++j) s4[j & 3] += __uint_as_float(raw[j]);
          rs += (s4[0] + s4[1]) + (s4[2] + s4[3]);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) lo[j] = lo_of_trunc(__uint_as_float(raw[j]));
        mbar_wait(ta_empty(t), pht ^ 1u);
        tc_fence_after();
        tmem_st32(trow + (uint32_t)(t * 64), raw);
        tmem_st32(trow + (uint32_t)(t * 64 + 32), lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (!BPRE) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ta_ready(t));
        if (tid == 64 && i < 12) TC_STAMP(32 + i);
        s += 2; if (s >= STAGES) { s -= STAGES; ph ^= 1u; }
        t += 2; if (t >= TA) { t -= TA; pht ^= 1u; }
      }
      if (do_rs && m0 + row < p.M) atomicAdd(p.rowsum + z * p.zsRowsum + m0 + row, p.alpha * rs);
      mbar_wait(acc_bar, 0);                       // every MMA of this CTA has retired: accumulators complete, ring stages idle
      tc_fence_after();
      if (tid == 64) TC_STAMP(2);
    }
  }
  // ===================== cluster split-K (p.csk): the CTAs of a (1, splitk, 1) cluster each accumulated a K range of the
  // SAME tile; ranks > 0 hand their partial sums to rank 0 through distributed shared memory and rank 0 runs the
  // (possibly non-linear) epilogue.  For projections with few tiles and a long K (K = 544..1024 at 2 304 tokens: 36 CTAs
  // walking 17-32 k-blocks) this spreads the serial k-loop over 2-3 SMs without the pre-zeroed C / atomics that plain
  // split-K needs.  Slot layout in rank 0's idle ring: [32-column chunk][float4 k][128 rows] (conflict-free both ways).
  constexpr uint32_t CSK_SLOT0 = 65536u, CSK_SLOT = 128u * BN * 4u;
  static_assert(SM2 || CSK_SLOT0 + 2 * CSK_SLOT <= (uint32_t)(STAGES * STAGE_BYTES), "two partial-sum slots must fit into the idle ring");
  const bool csk = !SM2 && p.csk != 0;
  uint32_t crank = 0;
  if (csk) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  // sum of this CTA's accumulators for one 32-column chunk (thread = tile row = TMEM lane)
  const int q_e = warp & 3, g_e = (warp - 2) >> 2;
  const int nused_e = min(nrot, nloc);         // main accumulators that received at least one k-block
  const uint32_t arow_e = tmem_base + ((uint32_t)(q_e * 32) << 16);
  auto load_sum = [&](int c0, float (&sum)[32]) {
    uint32_t r0[32], r1[32];
    if (SM2) {
      tmem_ld32(arow_e + (uint32_t)c0, r0);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) sum[j] = __uint_as_float(r0[j]);
    } else {
      // two TMEM loads in flight per wait; summed in fp32 RN
      tmem_ld32(arow_e + (uint32_t)(NMAIN * BN + c0), r0);
      tmem_ld32(arow_e + (uint32_t)c0, r1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) sum[j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
      if (nused_e > 1) {
        tmem_ld32(arow_e + (uint32_t)(BN + c0), r0);
        if (NMAIN > 2 && nused_e > 2) tmem_ld32(arow_e + (uint32_t)(2 * BN + c0), r1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(r0[j]);
        if (NMAIN > 2 && nused_e > 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(r1[j]);
        }
      }
    }
  };
  if (csk) {
    // B1: every CTA of the cluster is past its main loop (rank 0's ring stages may now be overwritten)
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (crank != 0 && warp >= 2 && nloc > 0) {
      const int row = q_e * 32 + lane;
#pragma unroll 1
      for (int c0 = g_e * (BN / 2); c0 < (g_e + 1) * (BN / 2); c0 += 32) {
        float sum[32];
        load_sum(c0, sum);
        const uint32_t local = smem_base + CSK_SLOT0 + (crank - 1) * CSK_SLOT + (uint32_t)(((c0 >> 5) * 8) * 128 + row) * 16u;
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0u));
#pragma unroll
        for (int k = 0; k < 8; ++k)
          asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote + (uint32_t)k * 2048u), "f"(sum[4 * k]),
                       "f"(sum[4 * k + 1]), "f"(sum[4 * k + 2]), "f"(sum[4 * k + 3]) : "memory");
      }
    }
    // B2: the partial sums have landed in rank 0's shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (nloc > 0 && warp >= 2 && (!csk || crank == 0)) {
    {
      // ===================== epilogue: TMEM -> registers -> smem transpose -> coalesced global =====================
      // 8 warps: warp w owns TMEM lane quarter w & 3 (hardware rule) and column half (w - 2) >> 2 of the tile
      const int g = g_e, q = q_e;
      const int hf = g;                            // column half
      const uint32_t stg = smem_base + (warp - 2) * (32 * TC_EPI_LD * 4);   // pipeline stages are idle now (shared address)
      const int cq = lane & 7, rsub = lane >> 3;
      EpiArgs ea;
      ea.mrow = m0 + q * 32 + rsub;                                // first row this lane stores (then +4 per iteration)
      ea.rows = min(32, p.M - (m0 + q * 32));                      // valid rows of this warp's 32-row slab
      ea.C = p.C + z * p.zsC + (long long)ea.mrow * p.ldc + n0 + cq * 4;
      ea.bias = (p.bias && blockIdx.y == 0) ? p.bias + z * p.zsBias : nullptr;
      ea.mask = p.mask ? p.mask + z * p.zsMask + (long long)ea.mrow * p.ldmask + n0 + cq * 4 : nullptr;
      ea.res1 = p.res1 ? p.res1 + z * p.zsR1 + (long long)ea.mrow * p.ldr1 + n0 + cq * 4 : nullptr;
      ea.res2 = p.res2 ? p.res2 + z * p.zsR2 + (long long)ea.mrow * p.ldr2 + n0 + cq * 4 : nullptr;
      ea.vec = p.vecE != 0;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int m = ea.mrow + 4 * it;
        ea.rd[it] = (p.rowdiv && m < p.M) ? __ldg(p.rowdiv + z * p.zsRow + m) : 1.f;
      }
      // which specialised epilogue (all warp-uniform): 0 plain, 1 /F (+column scale), 2 relu-mask, 3 residuals,
      // 4 accumulate, 5 split-K atomics, 6 anything else
      int ekind;
      {
        const bool lin = !p.rowdiv && !p.mask && !p.res1 && !p.res2 && p.colscale_n == 0;
        if (p.splitk > 1 && !csk) ekind = 5;
        else if (p.accumulate) ekind = (lin && !p.relu && !p.bias) ? 4 : 6;
        else if (lin) ekind = 0;
        else if (p.rowdiv && !p.mask && !p.res1 && !p.res2) ekind = 1;
        else if (p.mask && !p.rowdiv && !p.res1 && !p.res2 && p.colscale_n == 0 && !p.relu && !p.bias) ekind = 2;
        else if (p.res1 && !p.rowdiv && !p.mask && p.colscale_n == 0 && !p.relu && !p.bias) ekind = 3;
        else ekind = 6;
      }
#pragma unroll 1
      for (int c0 = hf * (BN / 2); c0 < (hf + 1) * (BN / 2); c0 += 32) {
        if (n0 + c0 >= p.N) break;
        float sum[32];
        load_sum(c0, sum);
        if (csk) {
          const uint32_t slot = smem_base + CSK_SLOT0 + (uint32_t)(((c0 >> 5) * 8) * 128 + q * 32 + lane) * 16u;
          for (uint32_t r = 0; r + 1 < (uint32_t)p.splitk; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 v = lds128(slot + r * CSK_SLOT + (uint32_t)k * 2048u);
              sum[4 * k] += v.x; sum[4 * k + 1] += v.y; sum[4 * k + 2] += v.z; sum[4 * k + 3] += v.w;
            }
          }
        }
        if (tid == 64 && c0 == 0) TC_STAMP(5);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          sts128(stg + (lane * TC_EPI_LD + 4 * k) * 4, make_float4(sum[4 * k], sum[4 * k + 1], sum[4 * k + 2], sum[4 * k + 3]));
        __syncwarp();
        if (tid == 64 && c0 == 0) TC_STAMP(6);
        const int n = n0 + c0 + cq * 4;
        switch (ekind) {
          case 0: tc_epi_chunk<false, false, 0, EM_STORE>(p, ea, stg, rsub, cq, n, c0); break;
          case 1: tc_epi_chunk<true, false, 0, EM_STORE>(p, ea, stg, rsub, cq, n, c0); break;
          case 2: tc_epi_chunk<false, true, 0, EM_STORE>(p, ea, stg, rsub, cq, n, c0); break;
          case 3: tc_epi_chunk<false, false, 2, EM_STORE>(p, ea, stg, rsub, cq, n, c0); break;
          case 4: tc_epi_chunk<false, false, 0, EM_ACCUM>(p, ea, stg, rsub, cq, n, c0); break;
          case 5: tc_epi_chunk<false, false, 0, EM_ATOMIC>(p, ea, stg, rsub, cq, n, c0); break;
          default: tc_epi_generic(p, ea, stg, rsub, cq, n, c0); break;
        }
        __syncwarp();
        if (tid == 64 && c0 == 0) TC_STAMP(7);
      }
    }
  }
  if (tid == 64) TC_STAMP(3);
  tc_fence_before();
  __syncthreads();
  if (tid == 0) TC_STAMP(4);
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
// atom32 = 0: SWIZZLE_128B (K-major operand tiles), 1: SWIZZLE_128B_ATOM_32B (MN-major tf32 B tiles read by the tensor
// core), 2: no swizzle (MN-major A tiles, read only by the converter warps)
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
                  atom32 == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE : atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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

template <int BN, bool AMN, bool BMN, bool BPRE, bool SM2 = false>
inline int gemm_tc_launch(const GemmP& p, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mbl, cudaStream_t st) {
  auto kern = gemm_tc_kernel<BN, AMN, BMN, BPRE, SM2>;
  static bool attr_done = false;
  if (!attr_done) {
    SGRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN, SM2>::SMEM));
    attr_done = true;
  }
  dim3 grid(ceil_div(p.M, TC_BM) * ceil_div(p.N, BN), p.splitk, p.nb);
  prof_begin(PC_GEMM_TC, 2.0 * p.M * p.N * (double)p.K * p.nb, st);
  if (p.csk) {      // thread-block cluster along the split-K dimension (+ programmatic dependent launch as everywhere)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TcCfg<BN, SM2>::SMEM; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = (unsigned)p.splitk; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (pdl_enabled() && pdl_stream_ok(st)) ? 2 : 1;
    cudaLaunchKernelEx(&cfg, kern, ma, mb, mbl, p);
  } else
  launch_k(kern, grid, TC_THREADS, TcCfg<BN, SM2>::SMEM, st, ma, mb, mbl, p);
  prof_end(st);
  SGRL_LAUNCH_OK();
  return 0;
}

template <int BN, bool BPRE>
inline int gemm_tc_dispatch(const GemmP& p, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mbl, cudaStream_t st) {
  if (!p.transA && !p.transB) return gemm_tc_launch<BN, false, false, BPRE>(p, ma, mb, mbl, st);
  if (!p.transA && p.transB) return gemm_tc_launch<BN, false, true, BPRE>(p, ma, mb, mbl, st);
  if (p.transA && p.transB) return gemm_tc_launch<BN, true, true, BPRE>(p, ma, mb, mbl, st);
  return gemm_tc_launch<BN, true, false, BPRE>(p, ma, mb, mbl, st);
}

extern long long* g_gemm_trace;   // sgrl_gemm_trace(): device buffer of 64 int64 for TC_STAMP, or nullptr

inline int gemm_tc(const GemmP& p_in, cudaStream_t st) {
  GemmP p = p_in;
  p.dbg = g_gemm_trace;
  static const int sched_env = getenv("SGRL_TC_SCHED") ? atoi(getenv("SGRL_TC_SCHED")) : 0;
  p.sched = sched_env;
  if (p.M <= 0 || p.N <= 0 || p.nb <= 0) return 0;
  SGRL_CHECK(gemm_tc_eligible(p), "gemm_tc: operands not TMA-compatible");
  SGRL_CHECK(p.splitk == 1 || (!p.relu && !p.rowdiv && !p.mask && !p.res1 && !p.res2 && p.colscale_n == 0), "gemm_tc: split-K only with a linear epilogue");
  SGRL_CHECK((p.Blo == nullptr) == (p.Bhi == nullptr), "gemm_tc: Bhi and Blo go together");
  SGRL_CHECK(p.Blo == nullptr || (host_vec_ok(p.Blo, p.ldb, p.zsB) && host_vec_ok(p.Bhi, p.ldb, p.zsB)), "gemm_tc: pre-split B parts not TMA-compatible");
  const float* Bsrc = p.Blo ? p.Bhi : p.B;
  // Tile width and split-K from a small cost model in SM cycles: ceil(CTAs / wave) x (fixed + k-blocks x block time).
  // Block times are the measured mainloop periods (tools/gemm_trace.py), the fixed part covers prologue, pipeline
  // fill and epilogue.  The step runs several kernels concurrently (three forward chains, dW GEMMs on side streams),
  // so what matters is SM-time (CTAs x duration) more than the duration of one kernel alone: the "wave" is therefore
  // a third of the machine, not 148 (swept on B200, tools/sweep_tiles.sh: 5.36 ms/update at 32, 5.49 at 48, 6.23 at 148; profiles/r01n_tile_sweep.txt).  A caller passes splitk > 1 to say "the epilogue is linear and C is pre-zeroed/accumulated":
  // only then may K be split (atomics); the factor itself is chosen here.
  const int nkb = ceil_div(p.K, TC_BK);
  int BN = 64, sk = 1;
  {
    struct Tune { double blk64, blk128, fix64, fix128, split_fix; int wave, wave_split; };
    static const Tune tn = [] {
      auto env = [](const char* k, double d) { const char* e = getenv(k); return e ? atof(e) : d; };
      Tune t;
      t.blk64 = env("SGRL_TC_BLK64", 860.0); t.blk128 = env("SGRL_TC_BLK128", 960.0);
      t.fix64 = env("SGRL_TC_FIX64", 3000.0); t.fix128 = env("SGRL_TC_FIX128", 4000.0);
      t.split_fix = env("SGRL_TC_SPLITFIX", 1500.0); t.wave = (int)env("SGRL_TC_WAVE", 32.0); t.wave_split = (int)env("SGRL_TC_WAVE_SPLIT", (double)t.wave);
      return t;
    }();
    const bool may_split = p.splitk > 1;
    double best = 1e30;
    for (int bn = 64; bn <= 128; bn += 64) {
      if (bn == 128 && p.N <= 64) break;
      const long long base = (long long)ceil_div(p.M, TC_BM) * ceil_div(p.N, bn) * p.nb;
      const double blk = bn == 128 ? tn.blk128 : tn.blk64, fixed = bn == 128 ? tn.fix128 : tn.fix64;
      const int sk_max = may_split ? (nkb / 4 < 1 ? 1 : (nkb / 4 > 64 ? 64 : nkb / 4)) : 1;
      for (int k = 1; k <= sk_max; ++k) {
        const long long ctas = base * k;
        const int wave = may_split ? tn.wave_split : tn.wave;      // weight-gradient GEMMs (side lanes) may be given a different share
        const double waves = (double)((ctas + wave - 1) / wave);
        const double cost = waves * (fixed + (k > 1 ? tn.split_fix : 0.0) + ceil_div(nkb, k) * blk);
        if (cost < best * 0.98) { best = cost; BN = bn; sk = k; }
      }
    }
  }
  p.splitk = sk;
  p.csk = 0;
  {
    // cluster split-K for launches that may NOT use atomics (non-linear epilogue or C not pre-zeroed), have few tiles and a
    // long K: 2-3 CTAs of a cluster share one tile's k-loop (SGRL_TC_CSK=0 disables; _MAXT: most tiles, _MINKB: fewest k-blocks).
    // Measured (tools/gemm_bench.py, tools/sweep_csk.sh; profiles/r01o_*): alone, dgrad 2304x256x1024 drops 22.2 -> 16.7 us and
    // 2304x256x768 18.2 -> 15.0 us, but the two cluster barriers + DSMEM hand-over cost ~4.5 us, so K = 544 gains nothing
    // (14.5 -> 13.7 us) and with every eligible launch clustered the whole update got SLOWER (5.45 vs 5.37 ms: clusters of
    // full-SM CTAs are harder to place next to the other streams' kernels).  Hence the conservative defaults: K >= 768, <= 40 tiles.
    static const int csk_on = getenv("SGRL_TC_CSK") ? atoi(getenv("SGRL_TC_CSK")) : 1;
    static const int csk_maxt = getenv("SGRL_TC_CSK_MAXT") ? atoi(getenv("SGRL_TC_CSK_MAXT")) : 40;
    static const int csk_minkb = getenv("SGRL_TC_CSK_MINKB") ? atoi(getenv("SGRL_TC_CSK_MINKB")) : 24;
    const long long tiles = (long long)ceil_div(p.M, TC_BM) * ceil_div(p.N, BN) * p.nb;
    if (csk_on && p_in.splitk <= 1 && sk == 1 && nkb >= csk_minkb && tiles <= csk_maxt) {
      p.splitk = (tiles * 3 <= 2 * csk_maxt && nkb >= 3 * csk_minkb / 2) ? 3 : 2;
      p.csk = 1;
    }
  }
  p.vecE = host_vec_ok(p.C, p.ldc, p.zsC) && (!p.mask || host_vec_ok(p.mask, p.ldmask, p.zsMask)) &&
           (!p.res1 || host_vec_ok(p.res1, p.ldr1, p.zsR1)) && (!p.res2 || host_vec_ok(p.res2, p.ldr2, p.zsR2));
  const long long nzA = p.zsA ? p.nb : 1, nzB = p.zsB ? p.nb : 1;
  CUtensorMap ma, mb, mbl;
  // K-major operand: inner = K (contiguous), outer = rows; MN-major: inner = rows (contiguous), outer = K
  if (!p.transA) SGRL_TRY(make_tmap(&ma, p.A, p.K, p.M, nzA, p.lda, p.zsA, 32, TC_BM, 0));
  else SGRL_TRY(make_tmap(&ma, p.A, p.M, p.K, nzA, p.lda, p.zsA, TC_BM, 32, 2));
  if (!p.transB) SGRL_TRY(make_tmap(&mb, Bsrc, p.K, p.N, nzB, p.ldb, p.zsB, 32, BN, 0));
  else SGRL_TRY(make_tmap(&mb, Bsrc, p.N, p.K, nzB, p.ldb, p.zsB, 32, 32, 1));
  mbl = mb;
  if (p.Blo) {
    if (!p.transB) SGRL_TRY(make_tmap(&mbl, p.Blo, p.K, p.N, nzB, p.ldb, p.zsB, 32, BN, 0));
    else SGRL_TRY(make_tmap(&mbl, p.Blo, p.N, p.K, nzB, p.ldb, p.zsB, 32, 32, 1));
  }
  // two-CTAs-per-SM variant: short-K tiles, pre-split weights, K-major A, and enough tiles that SMs hold several of them.
  // Its single accumulator takes 3x the truncating adds (measured: gradients of a B=100 critic step drift to 6e-4 on a few
  // tensors when it is forced everywhere), so only passes that keep nothing for a backward (rollout / target networks:
  // p.sm2_ok) may use it.  SGRL_TC_SM2: 0 never, 1 auto, 2 wherever the shape allows (tests); SGRL_TC_SM2_MIN: tile threshold
  {
    static const int sm2_mode = getenv("SGRL_TC_SM2") ? atoi(getenv("SGRL_TC_SM2")) : 1;
    static const int sm2_min = getenv("SGRL_TC_SM2_MIN") ? atoi(getenv("SGRL_TC_SM2_MIN")) : 2 * NUM_SMS;
    static const int sm2_maxkb = getenv("SGRL_TC_SM2_MAXKB") ? atoi(getenv("SGRL_TC_SM2_MAXKB")) : 12;     // k-blocks on the one accumulator
    const long long ctas = (long long)ceil_div(p.M, TC_BM) * ceil_div(p.N, BN) * p.nb * sk;
    if (!p.csk && sm2_mode && p.Blo && !p.transA && ceil_div(nkb, sk) <= sm2_maxkb && (sm2_mode == 2 || (p.sm2_ok && ctas >= sm2_min))) {
      if (BN == 128) return p.transB ? gemm_tc_launch<128, false, true, true, true>(p, ma, mb, mbl, st) : gemm_tc_launch<128, false, false, true, true>(p, ma, mb, mbl, st);
      return p.transB ? gemm_tc_launch<64, false, true, true, true>(p, ma, mb, mbl, st) : gemm_tc_launch<64, false, false, true, true>(p, ma, mb, mbl, st);
    }
  }
  if (BN == 128) return p.Blo ? gemm_tc_dispatch<128, true>(p, ma, mb, mbl, st) : gemm_tc_dispatch<128, false>(p, ma, mb, mbl, st);
  return p.Blo ? gemm_tc_dispatch<64, true>(p, ma, mb, mbl, st) : gemm_tc_dispatch<64, false>(p, ma, mb, mbl, st);
}

}  // namespace sgrl
