// K5/K6 — TD3 glue and the fused optimizer pass (src/agent.py:127-187, functional.py:7-10).
#pragma once
#include "common.cuh"

namespace sgrl {

// next_action = clamp(pi_target(s') + clamp(noise, +-c), +-max_action)        agent.py:128-133
__global__ void td3_smooth_action_kernel(const float* __restrict__ a, const float* __restrict__ noise, float* __restrict__ out,
                                         float noise_clip, float max_action, long long n) {
  SGRL_PDL_ENTER();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    const float nz = fminf(fmaxf(noise[i], -noise_clip), noise_clip);
    out[i] = fminf(fmaxf(a[i] + nz, -max_action), max_action);
  }
}

// Philox4x32-10 (Salmon et al., SC'11; the counter-based generator behind torch.cuda's and cuRAND's default streams),
// checked against the Random123 known-answer vectors in tests/test_td3_glue.py through the numpy restatement.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}
// Same smoothing with the noise drawn in the kernel: eps ~ N(0, policy_noise^2) (agent.py:128 torch.randn_like * policy_noise),
// four normals per Philox call (counter = {element / 4, 0, draw, 0}, key = seed; Box-Muller).  `draw` is a device counter the
// caller bumps once per update (sgrl_bump_step), so a replayed CUDA graph draws fresh noise every replay.
__global__ void td3_smooth_action_rng_kernel(const float* __restrict__ a, float* __restrict__ out, float* __restrict__ noise_out,
                                             float policy_noise, float noise_clip, float max_action, long long n,
                                             unsigned long long seed, const int* __restrict__ draw) {
  SGRL_PDL_ENTER();
  const unsigned d = (unsigned)*draw;
  const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
  const long long n4 = (n + 3) >> 2;
  for (long long q = blockIdx.x * 256LL + threadIdx.x; q < n4; q += gridDim.x * 256LL) {
    const uint4 r = philox4x32_10(make_uint4((unsigned)q, (unsigned)(q >> 32), d, 0u), key);
    const float u0 = ((float)r.x + 0.5f) * 2.3283064365386963e-10f, u1 = ((float)r.y + 0.5f) * 2.3283064365386963e-10f;
    const float u2 = ((float)r.z + 0.5f) * 2.3283064365386963e-10f, u3 = ((float)r.w + 0.5f) * 2.3283064365386963e-10f;
    const float r0 = sqrtf(-2.f * logf(fminf(u0, 0.99999994f))), r1 = sqrtf(-2.f * logf(fminf(u2, 0.99999994f)));
    float s0, c0, s1, c1;
    sincospif(2.f * u1, &s0, &c0);
    sincospif(2.f * u3, &s1, &c1);
    const float z[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = q * 4 + k;
      if (i >= n) break;
      const float e = z[k] * policy_noise;
      if (noise_out) noise_out[i] = e;
      out[i] = fminf(fmaxf(a[i] + fminf(fmaxf(e, -noise_clip), noise_clip), -max_action), max_action);
    }
  }
}

// target[t] = r[g(t)] * reward_scale + (1 - done[g(t)]) * discount * min(q1t[t], q2t[t])   agent.py:136-139
// dq1 = 2 (q1 - target)/T, dq2 likewise; loss += sum((q1-y)^2 + (q2-y)^2)/T               agent.py:146-148
__global__ void __launch_bounds__(256) td3_critic_loss_kernel(
    const float* __restrict__ q1, const float* __restrict__ q2, const float* __restrict__ tq1, const float* __restrict__ tq2,
    const float* __restrict__ reward, const float* __restrict__ done, const int* __restrict__ tok_graph,
    const float* __restrict__ tok_w,
    float* __restrict__ target, float* __restrict__ dq1, float* __restrict__ dq2, float* __restrict__ loss,
    float discount, float reward_scale, int T, double* __restrict__ rstats, int G) {
  SGRL_PDL_ENTER();
  __shared__ float red[8];
  if (rstats && blockIdx.x == 0 && threadIdx.x < 32) {
    // sum and sum of squares of the scaled rewards (agent.py:158-161 logs their mean / variance): fp64, one warp
    double s1 = 0.0, s2 = 0.0;
    for (int g = threadIdx.x; g < G; g += 32) { const double x = (double)(reward[g] * reward_scale); s1 += x; s2 += x * x; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if (threadIdx.x == 0) { rstats[0] = s1; rstats[1] = s2; }
  }
  float acc = 0.f;
  const float invT = 1.f / (float)T;
  for (int t = blockIdx.x * 256 + threadIdx.x; t < T; t += gridDim.x * 256) {
    const int g = tok_graph[t];
    const float y = reward[g] * reward_scale + (1.f - done[g]) * discount * fminf(tq1[t], tq2[t]);
    target[t] = y;
    const float e1 = q1[t] - y, e2 = q2[t] - y;
    // tok_w (packed mixed-morphology batches): weight of token t in the loss = 1 / (#morphologies * tokens of its morphology),
    // i.e. the mean over morphologies of the reference's per-morphology mse; single morphology: 1/T
    const float w = tok_w ? tok_w[t] : invT;
    dq1[t] = 2.f * e1 * w; dq2[t] = 2.f * e2 * w;
    acc += (e1 * e1 + e2 * e2) * (tok_w ? w * (float)T : 1.f);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0) atomicAdd(loss, s * invT);
  }
}

// actor loss = -mean(Q1):  dq = -1/T, loss += -sum(q)/T                                    agent.py:167
__global__ void __launch_bounds__(256) td3_actor_loss_kernel(const float* __restrict__ q1, const float* __restrict__ tok_w, float* __restrict__ dq,
                                                             float* __restrict__ loss, int T) {
  SGRL_PDL_ENTER();
  __shared__ float red[8];
  float acc = 0.f;
  const float invT = 1.f / (float)T;
  for (int t = blockIdx.x * 256 + threadIdx.x; t < T; t += gridDim.x * 256) {
    const float w = tok_w ? tok_w[t] : invT;
    dq[t] = -w; acc += q1[t] * (tok_w ? w * (float)T : 1.f);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0) atomicAdd(loss, -s * invT);
  }
}

// sum of squares of a flat gradient buffer                                                  clip_grad_norm_, agent.py:152-155
// Deterministic (no atomics): every block leaves its partial sum in a scratch slot, a second one-block launch adds the
// partials in index order.  The norm feeds the clip coefficient of every data-parallel replica: with an atomic reduction the
// ranks disagreed in its last bit and the replicas drifted apart after ~40 steps (tools/dp_debug.py, bench.py --check-replicas).
constexpr int SUMSQ_MAX_BLOCKS = 2048;
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
  SGRL_PDL_ENTER();
  __shared__ float red[8];
  float acc = 0.f;
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) {
    const float4 v = ldg4(g + i * 4);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[n4 * 4 + threadIdx.x]; acc += v * v; }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
  }
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(const float* __restrict__ partial, int nb, float* __restrict__ out) {
  SGRL_PDL_ENTER();
  __shared__ double red[8];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) acc += (double)partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    *out += (float)s;
  }
}

// lr / betas arrive as doubles, like the Python floats torch.optim.Adam computes its scalars from: 1 - beta2 is 0.001
// rounded once to fp32 (torch), not 1.f - 0.999f = 0.00099998713 (1.3e-5 off: visible in exp_avg_sq, tests/test_k6_gpu.py)
struct AdamCfg { double lr, beta1d, beta2d; float beta1, beta2, omb1, omb2, eps, max_norm; float grad_scale; };

// One pass over the live arena: clip coefficient from the global norm, Adam moment update,
// parameter step.  `step` lives on the device so the pass can be replayed from a CUDA graph.
//   coef = min(1, max_norm / (||g|| + 1e-6))     (torch.nn.utils.clip_grad_norm_)
//   m = m + (1-b1)(g - m);  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// grad_scale pre-multiplies g (1/world_size after a sum all-reduce).
// hi/lo (nullable): the refreshed tf32 split of the parameters for the tcgen05 projections (csrc/gemm_tc.cuh).
__device__ __forceinline__ void split_store4(float4 pv, float* hi, float* lo, long long i4) {
  float4 h, l;
  const float* x = &pv.x; float* hh = &h.x; float* ll = &l.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned hb = (__float_as_uint(x[k]) + 0x1000u) & 0xFFFFE000u;
    hh[k] = __uint_as_float(hb);
    ll[k] = __uint_as_float((__float_as_uint(x[k] - hh[k]) + 0x1000u) & 0xFFFFE000u);
  }
  stg4(hi + i4 * 4, h); stg4(lo + i4 * 4, l);
}

__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  SGRL_PDL_ENTER();
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) split_store4(ldg4(w + i * 4), hi, lo, i);
}

__global__ void __launch_bounds__(256) adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, const float* __restrict__ sumsq,
                                                        const int* __restrict__ step, AdamCfg c, float* __restrict__ hi, float* __restrict__ lo) {
  SGRL_PDL_ENTER();
  __shared__ float sh[3];
  if (threadIdx.x == 0) {
    const float nrm = sqrtf(*sumsq) * c.grad_scale;
    float coef = 1.f;
    if (c.max_norm > 0.f) { coef = c.max_norm / (nrm + 1e-6f); coef = coef > 1.f ? 1.f : coef; }
    const double t = (double)(*step);
    sh[0] = coef * c.grad_scale;
    sh[1] = (float)(c.lr / (1.0 - pow(c.beta1d, t)));
    sh[2] = (float)sqrt(1.0 - pow(c.beta2d, t));
  }
  __syncthreads();
  const float gs = sh[0], step_size = sh[1], bc2s = sh[2];
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) {
    float4 pv = *reinterpret_cast<float4*>(p + i * 4);
    const float4 gv = ldg4(g + i * 4);
    float4 mv = *reinterpret_cast<float4*>(m + i * 4), vv = *reinterpret_cast<float4*>(v + i * 4);
    float* pp = &pv.x; const float* gg = &gv.x; float* mm = &mv.x; float* vq = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gg[k] * gs;
      mm[k] = mm[k] + (gk - mm[k]) * c.omb1;
      vq[k] = vq[k] * c.beta2 + c.omb2 * gk * gk;
      pp[k] -= step_size * (mm[k] / (sqrtf(vq[k]) / bc2s + c.eps));
    }
    stg4(p + i * 4, pv); stg4(m + i * 4, mv); stg4(v + i * 4, vv);
    if (hi) split_store4(pv, hi, lo, i);
  }
}

__global__ void bump_step_kernel(int* step) {
  SGRL_PDL_ENTER(); *step += 1; }

// theta_t <- tau*theta + (1-tau)*theta_t                                                    functional.py:7-10
// Rounded exactly like the reference's three fp32 tensor ops (mul, mul, add — no FMA contraction):
// the per-step change tau*(theta-theta_t) ~ 5e-7 is a few ulps of O(1) weights, so op order is visible.
// hi/lo (nullable) cover the first n_split floats (the live prefix of the arena).
__global__ void __launch_bounds__(256) polyak_kernel(float* __restrict__ tgt, const float* __restrict__ src, long long n,
                                                     float tau, float one_minus_tau, float* __restrict__ hi, float* __restrict__ lo, long long n_split) {
  SGRL_PDL_ENTER();
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) {
    float4 t = *reinterpret_cast<float4*>(tgt + i * 4);
    const float4 s = ldg4(src + i * 4);
    t.x = __fadd_rn(__fmul_rn(tau, s.x), __fmul_rn(one_minus_tau, t.x));
    t.y = __fadd_rn(__fmul_rn(tau, s.y), __fmul_rn(one_minus_tau, t.y));
    t.z = __fadd_rn(__fmul_rn(tau, s.z), __fmul_rn(one_minus_tau, t.z));
    t.w = __fadd_rn(__fmul_rn(tau, s.w), __fmul_rn(one_minus_tau, t.w));
    stg4(tgt + i * 4, t);
    if (hi && i * 4 < n_split) split_store4(t, hi, lo, i);
  }
}

inline int grid_for_flat(long long n) {
  long long b = (n / 4 + 255) / 256;
  if (b > 8LL * NUM_SMS) b = 8LL * NUM_SMS;
  return b < 1 ? 1 : (int)b;
}

}  // namespace sgrl
