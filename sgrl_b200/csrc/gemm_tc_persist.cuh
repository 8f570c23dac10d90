// K3-P — persistent variant of the tcgen05 projection kernel for inference passes with many tiles per SM (rollout:
// 1 152 .. 3 456 row tiles per launch).  Included by gemm_tc.cuh (shares its PTX wrappers, TcGroup and descriptors).
//
// Why: measured on a steady-state CTA of a 147 456 x 768 x 256 launch (tools/gemm_trace.py, profiles/r04*): a 128 x 128 x 256
// tile holds the tensor pipe for 6.1 K cycles (8 k-blocks x 12 MMAs x 64 cycles) but costs 11.6 K cycles of SM time even with
// two CTAs per SM, because each CTA still pays its own pipeline fill (first operands land 2.5-4 K cycles after entry), a
// 2-stage ring that exposes the TMA latency every other k-block, and an epilogue with the tensor pipe idle.  Here ONE CTA per
// SM walks tiles  t = blockIdx.x, blockIdx.x + gridDim.x, ...  with
//   * a 4-stage operand ring and 4 tensor-memory A slots that run on ACROSS tile boundaries (the producer and the converter
//     warps are already working on tile t+1 while tile t's last MMAs execute),
//   * TWO accumulator buffers in tensor memory (2 x 128 columns): dedicated epilogue warps drain buffer b (tcgen05.ld ->
//     bias / relu / F-division in registers -> swizzled shared box -> cp.async.bulk.tensor store) while the MMA warp fills
//     buffer b^1,
// so that in steady state the SM's tensor pipe only waits for operands.  Like the two-CTAs-per-SM variant it keeps all three
// product terms of the 3xTF32 split in ONE accumulator (<= 12 k-blocks, <= 144 truncating adds: measured 2e-6 relative), so
// only passes that keep nothing for a backward use it (GemmP::sm2_ok), and only the store-class epilogues (GemmP::tma_c).
//
// CTA = 14 warps: warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-9 two converter groups (alternate k-blocks),
// warps 10-13 epilogue (warp w owns TMEM lane quarter w & 3).
#pragma once

namespace sgrl {

constexpr int TCP_BN = 128;
constexpr int TCP_STAGES = 4, TCP_TA = 4;
constexpr int TCP_B_BYTES = TCP_BN * TC_BK * 4;                    // 16 KiB (hi or lo)
constexpr int TCP_STAGE_BYTES = TC_A_BYTES + 2 * TCP_B_BYTES;      // 48 KiB: [A raw | B hi | B lo]
constexpr int TCP_EPI_WARPS = 4;
constexpr int TCP_THREADS = 64 + TC_CONV_THREADS + 32 * TCP_EPI_WARPS;   // 448
constexpr int TCP_RING_BYTES = TCP_STAGES * TCP_STAGE_BYTES;       // 192 KiB
constexpr int TCP_BOX_BYTES = 32 * 32 * 4;                         // one warp's 32 x 32 fp32 store box
constexpr int TCP_SMEM = TCP_RING_BYTES + TCP_EPI_WARPS * 2 * TCP_BOX_BYTES + 256 /*barriers*/ + 1024 /*align slack*/;
constexpr int TCP_MAX_KB = 12;                                     // k-blocks per tile on the one accumulator

struct TcpTile { int gi, z, m0, n0, nkb; };
// virtual tile id -> problem of the group, instance, tile origin (n fastest: the CTAs working side by side share A rows in L2)
__device__ __forceinline__ TcpTile tcp_decode(const TcGroup& grp, int vt) {
  TcpTile t;
  t.gi = 0;
#pragma unroll
  for (int k = 1; k < TC_MAXG; ++k) if (k < grp.n && vt >= grp.pr[k].cta_begin) t.gi = k;
  const GemmP& p = grp.pr[t.gi].p;
  const int local = vt - grp.pr[t.gi].cta_begin;
  const int tiles_n = (p.N + TCP_BN - 1) / TCP_BN, tiles_m = (p.M + TC_BM - 1) / TC_BM;
  const int per_z = tiles_m * tiles_n;
  t.z = local / per_z;
  const int rem = local - t.z * per_z;
  t.m0 = (rem / tiles_n) * TC_BM;
  t.n0 = (rem % tiles_n) * TCP_BN;
  t.nkb = (p.K + TC_BK - 1) / TC_BK;
  return t;
}

#if SGRL_TC_PART == 4      // kernel bodies only in the translation unit that launches them (gemm_tc.cuh: parts)
__global__ void __launch_bounds__(TCP_THREADS, 1) gemm_tc_persist_kernel(const __grid_constant__ TcGroup grp, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t epi_base = smem_base + TCP_RING_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TCP_RING_BYTES + TCP_EPI_WARPS * 2 * TCP_BOX_BYTES);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TCP_STAGES + s); };
  auto ta_ready = [&](int t) { return bar_base + 8u * (2 * TCP_STAGES + t); };
  auto ta_empty = [&](int t) { return bar_base + 8u * (2 * TCP_STAGES + TCP_TA + t); };
  auto acc_full = [&](int b) { return bar_base + 8u * (2 * TCP_STAGES + 2 * TCP_TA + b); };
  auto acc_empty = [&](int b) { return bar_base + 8u * (2 * TCP_STAGES + 2 * TCP_TA + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TCP_STAGES + 2 * TCP_TA + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  SGRL_PDL_TRIGGER();
  if (tid == 0) {
    for (int s = 0; s < TCP_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int t = 0; t < TCP_TA; ++t) { mbar_init(ta_ready(t), TC_CONV_WARPS / 2); mbar_init(ta_empty(t), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), TCP_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int g = 0; g < grp.n; ++g) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&grp.pr[g].mapA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&grp.pr[g].mapB) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&grp.pr[g].mapBlo) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&grp.pr[g].mapC) : "memory");
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;                       // columns [0, 256): two accumulators; [256, 512): 4 A slots [hi 32 | lo 32]
  const uint32_t ta_base = tmem_base + 2 * TCP_BN;
  SGRL_PDL_WAIT();

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t kbg = 0;
    for (int vt = blockIdx.x; vt < total_tiles; vt += gridDim.x) {
      const TcpTile t = tcp_decode(grp, vt);
      const TcProblem& pr = grp.pr[t.gi];
      const int zA = pr.p.zsA ? t.z : 0, zB = pr.p.zsB ? t.z : 0;
      const bool bf = pr.p.prec == 1;          // BF16-input mode: the weights' lo tile is never read
      for (int i = 0; i < t.nkb; ++i, ++kbg) {
        const int s = kbg % TCP_STAGES;
        mbar_wait(empty_bar(s), ((kbg / TCP_STAGES) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(s), bf ? TCP_STAGE_BYTES - TCP_B_BYTES : TCP_STAGE_BYTES);
          const uint32_t a_dst = smem_base + s * TCP_STAGE_BYTES, b_dst = a_dst + TC_A_BYTES;
          tma_load_3d(a_dst, &pr.mapA, full_bar(s), i * TC_BK, t.m0, zA);
          tma_load_3d(b_dst, &pr.mapB, full_bar(s), i * TC_BK, t.n0, zB);
          if (!bf) tma_load_3d(b_dst + TCP_B_BYTES, &pr.mapBlo, full_bar(s), i * TC_BK, t.n0, zB);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    // instruction descriptor: D=f32, A=B=tf32, K-major; the N field is set per tile (a 32-column problem issues N = 32 MMAs:
    // 16 instead of 64 cycles each — the Z = [X P^T | gd] projections are then bound by their 64 KB of A per tile, not by MMAs on zero columns)
    constexpr uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BM >> 4) << 24);
    constexpr uint32_t B_HIW = (1024u >> 4) | (1u << 14) | (2u << 29), B_LOW = (16u >> 4) << 16;     // K-major, SWIZZLE_128B
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t tab = tb + 2 * TCP_BN;
    const uint32_t bdesc0 = (((smem_base + TC_A_BYTES) >> 4) & 0x3FFFu) | B_LOW;
    uint32_t kbg = 0, tcount = 0;
    bool ready = false;                       // barriers of the block about to be issued were already seen complete
    for (int vt = blockIdx.x; vt < total_tiles; vt += gridDim.x, ++tcount) {
      const TcpTile t = tcp_decode(grp, vt);
      const uint32_t b = tcount & 1u;
      mbar_wait(acc_empty(b), ((tcount >> 1) & 1u) ^ 1u);      // the epilogue has drained this buffer (two tiles ago)
      tc_fence_after();
      const uint32_t acc = tb + b * TCP_BN;
      int n_eff = grp.pr[t.gi].p.N - t.n0;                       // columns of this tile, rounded up to the MMA's N granularity of 16
      n_eff = n_eff >= TCP_BN ? TCP_BN : ((n_eff + 15) & ~15);
      const uint32_t idesc = idesc0 | ((uint32_t)(n_eff >> 3) << 17);
      const bool bf = grp.pr[t.gi].p.prec == 1;     // BF16-input mode (warp-uniform): operands rounded to bf16 by the converters, hi*hi only
#pragma unroll 1
      for (int i = 0; i < t.nkb; ++i, ++kbg) {
        const int s = kbg % TCP_STAGES, ts = kbg % TCP_TA;
        if (!ready) {
          mbar_wait(full_bar(s), (kbg / TCP_STAGES) & 1u);
          mbar_wait(ta_ready(ts), (kbg / TCP_TA) & 1u);
          tc_fence_after();
        }
        const uint32_t b_hi = bdesc0 + (uint32_t)(s * (TCP_STAGE_BYTES >> 4)), b_lo = b_hi + (TCP_B_BYTES >> 4);
        const uint32_t a_hi = tab + ts * 64, a_lo = a_hi + 32;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_tf32_ts(acc, a_hi + 8 * k, b_hi + 2 * k, B_HIW, idesc, (i > 0 || k > 0) ? 1u : 0u);
          if (!bf) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              tc_mma_tf32_ts(acc, a_lo + 8 * k, b_hi + 2 * k, B_HIW, idesc, 1u);
              tc_mma_tf32_ts(acc, a_hi + 8 * k, b_lo + 2 * k, B_HIW, idesc, 1u);
            }
          }
        }
        __syncwarp();
        // the next block's barriers are probed behind the 8 queued MMAs; never blocked on here (see gemm_tc_kernel)
        ready = false;
        if (i + 1 < t.nkb) {
          const uint32_t kn = kbg + 1;
          ready = mbar_test_all(full_bar(kn % TCP_STAGES), (kn / TCP_STAGES) & 1u) && mbar_test_all(ta_ready(kn % TCP_TA), (kn / TCP_TA) & 1u);
          if (ready) tc_fence_after();
        }
        if (elect_one()) {
          if (!bf) {
#pragma unroll
            for (int k = 2; k < 4; ++k) {
              tc_mma_tf32_ts(acc, a_lo + 8 * k, b_hi + 2 * k, B_HIW, idesc, 1u);
              tc_mma_tf32_ts(acc, a_hi + 8 * k, b_lo + 2 * k, B_HIW, idesc, 1u);
            }
          }
          tc_commit(empty_bar(s));
          tc_commit(ta_empty(ts));
          if (i == t.nkb - 1) tc_commit(acc_full(b));
        }
        __syncwarp();
      }
    }
  } else if (warp < 2 + TC_CONV_WARPS) {
    // ===================== converters: landed A k-block -> [hi | lo] rows of a tensor-memory slot =====================
    const int g = (warp - 2) >> 2, q = warp & 3, row = q * 32 + lane, cgt = ((warp - 2) & 3) * 32 + lane;
    const uint32_t trow = ta_base + ((uint32_t)(q * 32) << 16);
    uint32_t kbg = 0;
    for (int vt = blockIdx.x; vt < total_tiles; vt += gridDim.x) {
      const TcpTile t = tcp_decode(grp, vt);
      const bool bf = grp.pr[t.gi].p.prec == 1;
      for (int i = 0; i < t.nkb; ++i, ++kbg) {
        if ((int)(kbg & 1u) != g) continue;
        const int s = kbg % TCP_STAGES, ts = kbg % TCP_TA;
        mbar_wait(full_bar(s), (kbg / TCP_STAGES) & 1u);
        const uint32_t st = smem_base + s * TCP_STAGE_BYTES;
        uint32_t raw[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = lds128(st + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4));
          raw[4 * c] = __float_as_uint(v.x); raw[4 * c + 1] = __float_as_uint(v.y);
          raw[4 * c + 2] = __float_as_uint(v.z); raw[4 * c + 3] = __float_as_uint(v.w);
        }
        if (bf) {            // BF16-input mode: the weights' hi tile rounded in place by this group's 128 threads, the A rows in registers
          round_tile_bf16<TCP_B_BYTES / 16, 128>(st + TC_A_BYTES, cgt);
#pragma unroll
          for (int j = 0; j < 32; ++j) raw[j] = bf16_rn_bits(raw[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) lo[j] = lo_of_trunc(__uint_as_float(raw[j]));
        }
        mbar_wait(ta_empty(ts), ((kbg / TCP_TA) & 1u) ^ 1u);
        tc_fence_after();
        tmem_st32(trow + (uint32_t)(ts * 64), raw);
        if (!bf) tmem_st32(trow + (uint32_t)(ts * 64 + 32), lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (bf) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the tensor core
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ta_ready(ts));
      }
    }
  } else {
    // ===================== epilogue warps: accumulator buffer -> registers -> swizzled box -> TMA tensor store =====================
    const int ew = warp - (2 + TC_CONV_WARPS), q = warp & 3;
    const uint32_t arow = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t rsw = (uint32_t)(lane & 7);
    uint32_t tcount = 0, nbox = 0;
    for (int vt = blockIdx.x; vt < total_tiles; vt += gridDim.x, ++tcount) {
      const TcpTile t = tcp_decode(grp, vt);
      const TcProblem& pr = grp.pr[t.gi];
      const GemmP& p = pr.p;
      const uint32_t b = tcount & 1u;
      const int m_w = t.m0 + q * 32;
      const bool rows_ok = m_w < p.M;                            // warp-uniform
      const float* bias = p.bias ? p.bias + t.z * p.zsBias : nullptr;
      float rinv = 1.f;
      if (p.rowdiv && m_w + lane < p.M) rinv = 1.f / __ldg(p.rowdiv + t.z * p.zsRow + m_w + lane);
      if (bias && lane < TCP_BN / 32 && t.n0 + lane * 32 < p.N) asm volatile("prefetch.global.L1 [%0];" ::"l"(bias + t.n0 + lane * 32));
      const float alpha = p.alpha, csv = p.colscale;
      const bool relu = p.relu != 0, rdiv = p.rowdiv != nullptr;
      mbar_wait(acc_full(b), (tcount >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < TCP_BN / 32; ++c) {
        const int nc = t.n0 + 32 * c;
        uint32_t r[32];
        tmem_ld32(arow + b * TCP_BN + (uint32_t)(32 * c), r);
        float bj[32];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int n = nc + 4 * k;
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias && n + 3 < p.N) bv = ldg4(bias + n);
          bj[4 * k] = bv.x; bj[4 * k + 1] = bv.y; bj[4 * k + 2] = bv.z; bj[4 * k + 3] = bv.w;
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c == TCP_BN / 32 - 1) {            // every column of the buffer is in registers: the MMA warp may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(b));
        }
        if (!rows_ok || nc >= p.N) continue;   // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = fmaf(alpha, __uint_as_float(r[j]), bj[j]);
          v[j] = (relu && x < 0.f) ? 0.f : x;
        }
        if (p.gdcols && nc == 0) {             // Z = [X P^T | gd]: columns 30, 31 of row m = 3 t + r are gd[t][r][0..1]  (N == 32)
          const int m = m_w + lane;
          if (m < p.M) {
            const float2 gv = __ldg(reinterpret_cast<const float2*>(p.gdcols + t.z * p.zsGd + (long long)(m / 3) * 6 + (m % 3) * 2));
            v[NPJ] = gv.x; v[NPJ + 1] = gv.y;
          }
        }
        if (rdiv) {
          const int ncs = p.colscale_n - nc;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= rinv * (j < ncs ? csv : 1.f);
        }
        const uint32_t box = epi_base + (uint32_t)(ew * 2 + (nbox & 1u)) * TCP_BOX_BYTES;
        ++nbox;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // the store that last used this box has read it
        __syncwarp();
        const uint32_t rowa = box + (uint32_t)lane * 128u;
#pragma unroll
        for (int k = 0; k < 8; ++k) sts128(rowa + (((uint32_t)k ^ rsw) << 4), make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) { tma_store_3d(&pr.mapC, box, nc, m_w, t.z); tma_store_commit(); }
      }
    }
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


#endif
// ======================================================================================================================
// K3-PG — persistent GRAM projection for inference passes:  C = epi( tri(Z^T Z) W'^T )  with the A operand generated by the
// converter warps (GemmP::gramZ, see gemm_tc_kernel<.., GRAM>) and ALL N <= 256 output columns in one tile, so that each
// token's packed Gram row (17 k-blocks of 32) is generated ONCE instead of once per 128-column tile — generation, not the
// tensor pipe, bounds the 128-wide GRAM tile (measured, tools/gemm_trace.py: ~1 000 cycles per k-block against 768 of MMAs).
// One CTA per SM walks row tiles; B (folded weights, pre-split) streams through a 2-stage ring of [hi | lo] 256 x 32 tiles;
// tensor memory = one 256-column accumulator + 4 A slots; the converter warps already generate the next tile's first k-blocks
// while the epilogue warps drain the accumulator.  Single accumulator for all three product terms: inference passes only.
// ======================================================================================================================
constexpr int TCG_BN = 256, TCG_STAGES = 2, TCG_TA = 4;
constexpr int TCG_B_BYTES = TCG_BN * TC_BK * 4;                    // 32 KiB (hi or lo)
constexpr int TCG_STAGE_BYTES = 2 * TCG_B_BYTES;                   // 64 KiB
constexpr int TCG_RING_BYTES = TCG_STAGES * TCG_STAGE_BYTES;       // 128 KiB
constexpr int TCG_Z_BYTES = TC_BM * GRAM_ZLD * 4;                  // 50 KiB: the tile's Z rows, padded stride
constexpr int TCG_SMEM = TCG_RING_BYTES + TCG_Z_BYTES + 1024 /*partial ||G||^2*/ + TCP_EPI_WARPS * 2 * TCP_BOX_BYTES + 256 + 1024;
constexpr int TCG_NKB = GP_K / TC_BK;                              // 17

#if SGRL_TC_PART == 4
__global__ void __launch_bounds__(TCP_THREADS, 1) gemm_tc_gram_persist_kernel(const __grid_constant__ TcGroup grp, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t gram_z = smem_base + TCG_RING_BYTES;
  float* gram_ss = reinterpret_cast<float*>(smem + TCG_RING_BYTES + TCG_Z_BYTES);               // [2 groups][128 rows]
  const uint32_t epi_base = smem_base + TCG_RING_BYTES + TCG_Z_BYTES + 1024;                    // 1024-aligned: 128 K + 50 K + 1 K
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TCG_RING_BYTES + TCG_Z_BYTES + 1024 + TCP_EPI_WARPS * 2 * TCP_BOX_BYTES);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TCG_STAGES + s); };
  auto ta_ready = [&](int t) { return bar_base + 8u * (2 * TCG_STAGES + t); };
  auto ta_empty = [&](int t) { return bar_base + 8u * (2 * TCG_STAGES + TCG_TA + t); };
  const uint32_t acc_full = bar_base + 8u * (2 * TCG_STAGES + 2 * TCG_TA), acc_empty = acc_full + 8u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TCG_STAGES + 2 * TCG_TA + 2);
  static_assert((TCG_RING_BYTES + TCG_Z_BYTES + 1024) % 1024 == 0, "store boxes need 1024-byte alignment (SWIZZLE_128B)");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TcProblem& pr = grp.pr[0];
  const GemmP& p = pr.p;
  const int tiles_m = (p.M + TC_BM - 1) / TC_BM;                 // tile id = z * tiles_m + row tile
  SGRL_PDL_TRIGGER();
  if (tid == 0) {
    for (int s = 0; s < TCG_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int t = 0; t < TCG_TA; ++t) { mbar_init(ta_ready(t), TC_CONV_WARPS / 2); mbar_init(ta_empty(t), 1); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, TCP_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&pr.mapB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&pr.mapBlo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&pr.mapC) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;                         // columns [0, 256): accumulator; [256, 512): 4 A slots [hi 32 | lo 32]
  const uint32_t ta_base = tmem_base + TCG_BN;
  SGRL_PDL_WAIT();

  if (warp == 0) {
    // ===================== TMA producer: the folded weights' k-blocks (the same 17 for every tile; L2-resident) =====================
    uint32_t kbg = 0;
    for (int vt = blockIdx.x; vt < total_tiles; vt += gridDim.x) {
      const int zB = p.zsB ? vt / tiles_m : 0;
      for (int i = 0; i < TCG_NKB; ++i, ++kbg) {
        const int s = kbg % TCG_STAGES;
        mbar_wait(empty_bar(s), ((kbg / TCG_STAGES) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(s), p.prec == 1 ? TCG_B_BYTES : TCG_STAGE_BYTES);        // BF16-input mode: no lo tile
          const uint32_t b_dst = smem_base + s * TCG_STAGE_BYTES;
          tma_load_3d(b_dst, &pr.mapB, full_bar(s), i * TC_BK, 0, zB);
          if (p.prec != 1) tma_load_3d(b_dst + TCG_B_BYTES, &pr.mapBlo, full_bar(s), i * TC_BK, 0, zB);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int n_eff = (p.N + 15) & ~15;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_eff >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    constexpr uint32_t B_HIW = (1024u >> 4) | (1u << 14) | (2u << 29), B_LOW = (16u >> 4) << 16;     // K-major, SWIZZLE_128B
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t tab = tb + TCG_BN;
    const uint32_t bdesc0 = ((smem_base >> 4) & 0x3FFFu) | B_LOW;
    uint32_t kbg = 0, tcount = 0;
    bool ready = false;
    for (int vt = blockIdx.x; vt < total_tiles; vt += gridDim.x, ++tcount) {
      mbar_wait(acc_empty, (tcount & 1u) ^ 1u);                // the epilogue has drained the previous tile
      tc_fence_after();
#pragma unroll 1
      for (int i = 0; i < TCG_NKB; ++i, ++kbg) {
        const int s = kbg % TCG_STAGES, ts = kbg % TCG_TA;
        if (!ready) {
          mbar_wait(full_bar(s), (kbg / TCG_STAGES) & 1u);
          mbar_wait(ta_ready(ts), (kbg / TCG_TA) & 1u);
          tc_fence_after();
        }
        const uint32_t b_hi = bdesc0 + (uint32_t)(s * (TCG_STAGE_BYTES >> 4)), b_lo = b_hi + (TCG_B_BYTES >> 4);
        const uint32_t a_hi = tab + ts * 64, a_lo = a_hi + 32;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_tf32_ts(tb, a_hi + 8 * k, b_hi + 2 * k, B_HIW, idesc, (i > 0 || k > 0) ? 1u : 0u);
          if (p.prec != 1) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              tc_mma_tf32_ts(tb, a_lo + 8 * k, b_hi + 2 * k, B_HIW, idesc, 1u);
              tc_mma_tf32_ts(tb, a_hi + 8 * k, b_lo + 2 * k, B_HIW, idesc, 1u);
            }
          }
        }
        __syncwarp();
        ready = false;
        if (i + 1 < TCG_NKB) {
          const uint32_t kn = kbg + 1;
          ready = mbar_test_all(full_bar(kn % TCG_STAGES), (kn / TCG_STAGES) & 1u) && mbar_test_all(ta_ready(kn % TCG_TA), (kn / TCG_TA) & 1u);
          if (ready) tc_fence_after();
        }
        if (elect_one()) {
          if (p.prec != 1) {
#pragma unroll
            for (int k = 2; k < 4; ++k) {
              tc_mma_tf32_ts(tb, a_lo + 8 * k, b_hi + 2 * k, B_HIW, idesc, 1u);
              tc_mma_tf32_ts(tb, a_hi + 8 * k, b_lo + 2 * k, B_HIW, idesc, 1u);
            }
          }
          tc_commit(empty_bar(s));
          tc_commit(ta_empty(ts));
          if (i == TCG_NKB - 1) tc_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else if (warp < 2 + TC_CONV_WARPS) {
    // ===================== generators: Z tile -> packed Gram rows -> [hi | lo] in a tensor-memory slot =====================
    const int g = (warp - 2) >> 2, q = warp & 3, row = q * 32 + lane;
    const int cth = (warp - 2) * 32 + lane;                        // 0..255
    const uint32_t trow = ta_base + ((uint32_t)(q * 32) << 16);
    const uint32_t zrow = gram_z + (uint32_t)(row * GRAM_ZLD) * 4u;
    uint32_t kbg = 0;
    for (int vt = blockIdx.x; vt < total_tiles; vt += gridDim.x) {
      const int z = vt / tiles_m, m0 = (vt - z * tiles_m) * TC_BM;
      // both groups are done reading the previous tile's Z rows (and group 0 its partial sums) before they are overwritten
      asm volatile("bar.sync 1, %0;" ::"n"(TC_CONV_THREADS) : "memory");
      const float4* zsrc = reinterpret_cast<const float4*>(p.gramZ + z * p.zsGramZ + (long long)m0 * 96);
#pragma unroll
      for (int c = 0; c < 12; ++c) {
        const int f = cth + c * TC_CONV_THREADS, r = f / 24, c4 = f - r * 24;
        const float4 v = (m0 + r < p.M) ? __ldg(zsrc + f) : make_float4(0.f, 0.f, 0.f, 0.f);
        sts128(gram_z + (uint32_t)(r * GRAM_ZLD + c4 * 4) * 4u, v);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(TC_CONV_THREADS) : "memory");
      float ss = 0.f;
      for (int i = 0; i < TCG_NKB; ++i, ++kbg) {
        if ((int)(kbg & 1u) != g) continue;
        const int ts = kbg % TCG_TA;
        uint32_t raw[32];
        gram_kblock(i, zrow, raw, ss);                     // generated before the slot wait: overlaps the MMAs still reading it
        const bool bf = p.prec == 1;
        if (bf) {            // BF16-input mode: this group's 128 threads round the landed folded-weight tile in place, and the generated rows
          const int s = kbg % TCG_STAGES;
          mbar_wait(full_bar(s), (kbg / TCG_STAGES) & 1u);
          round_tile_bf16<TCG_B_BYTES / 16, 128>(smem_base + s * TCG_STAGE_BYTES, ((warp - 2) & 3) * 32 + lane);
#pragma unroll
          for (int j = 0; j < 32; ++j) raw[j] = bf16_rn_bits(raw[j]);
        }
        mbar_wait(ta_empty(ts), ((kbg / TCG_TA) & 1u) ^ 1u);
        tc_fence_after();
        tmem_st32(trow + (uint32_t)(ts * 64), raw);
        if (!bf) {
#pragma unroll
          for (int j = 0; j < 32; ++j) raw[j] = lo_of_trunc(__uint_as_float(raw[j]));
          tmem_st32(trow + (uint32_t)(ts * 64 + 32), raw);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (bf) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ta_ready(ts));
      }
      gram_ss[g * 128 + row] = ss;
      asm volatile("bar.sync 1, %0;" ::"n"(TC_CONV_THREADS) : "memory");          // both groups' partial ||G||_F^2 are in gram_ss
      if (g == 0 && p.gramF && m0 + row < p.M) p.gramF[z * p.zsGramZ + m0 + row] = sqrtf(gram_ss[row] + gram_ss[128 + row]) + 1.0f;
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - (2 + TC_CONV_WARPS), q = warp & 3;
    const uint32_t arow = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t rsw = (uint32_t)(lane & 7);
    const float alpha = p.alpha;
    const bool relu = p.relu != 0;
    const int nch = (p.N + 31) / 32;
    uint32_t tcount = 0, nbox = 0;
    for (int vt = blockIdx.x; vt < total_tiles; vt += gridDim.x, ++tcount) {
      const int z = vt / tiles_m, m0 = (vt - z * tiles_m) * TC_BM;
      const int m_w = m0 + q * 32;
      const bool rows_ok = m_w < p.M;
      const float* bias = p.bias ? p.bias + z * p.zsBias : nullptr;
      if (bias && lane < nch) asm volatile("prefetch.global.L1 [%0];" ::"l"(bias + lane * 32));
      mbar_wait(acc_full, tcount & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < nch; ++c) {
        const int nc = 32 * c;
        uint32_t r[32];
        tmem_ld32(arow + (uint32_t)nc, r);
        float bj[32];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int n = nc + 4 * k;
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias && n + 3 < p.N) bv = ldg4(bias + n);
          bj[4 * k] = bv.x; bj[4 * k + 1] = bv.y; bj[4 * k + 2] = bv.z; bj[4 * k + 3] = bv.w;
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c == nch - 1) {                    // the whole accumulator is in registers / on its way out: the next tile's MMAs may start
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
        }
        if (!rows_ok) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = fmaf(alpha, __uint_as_float(r[j]), bj[j]);
          v[j] = (relu && x < 0.f) ? 0.f : x;
        }
        const uint32_t box = epi_base + (uint32_t)(ew * 2 + (nbox & 1u)) * TCP_BOX_BYTES;
        ++nbox;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        const uint32_t rowa = box + (uint32_t)lane * 128u;
#pragma unroll
        for (int k = 0; k < 8; ++k) sts128(rowa + (((uint32_t)k ^ rsw) << 4), make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) { tma_store_3d(&pr.mapC, box, nc, m_w, z); tma_store_commit(); }
      }
    }
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

#endif
}  // namespace sgrl
