// K7 — device-resident replay storage (src/common/buffer.py:35-126).
//
// The reference keeps five numpy arrays on the host and, per TD3 update, fancy-indexes each
// of them and pushes five tensors through pageable H2D copies (buffer.py:103-120).  Here
// one transition is ONE packed row in HBM,
//     row = [ obs (od) | action (ad) | next_obs (od) | reward | done ],   rw = 2*od + ad + 2
// so sampling is a single gather: one warp per sampled row streams the row's 4*rw
// contiguous bytes and scatters them into the five batch tensors (which may be the static
// input buffers of Agent.update's plan).  HBM-bound: 8*rw bytes per sampled row.
#pragma once
#include "common.cuh"

namespace sgrl {

struct ReplayOut {
  float* obs; float* act; float* nobs; float* rew; float* done;
};

// idx (B) int64 row numbers.  Rows are 4-byte aligned only (rw is odd for most
// morphologies), so the copy is scalar but fully coalesced: lane k reads float k, k+32, ...
__global__ void __launch_bounds__(256) replay_gather_kernel(const float* __restrict__ rows, long long rw, const long long* __restrict__ idx,
                                                            int B, int od, int ad, long long cap, ReplayOut o) {
  SGRL_PDL_ENTER();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < B; b += gridDim.x * wpb) {
    long long r = idx[b];
    if (r < 0) r += cap;                       // numpy negative indexing
    if (r < 0 || r >= cap) r = 0;              // host validated the indices; never read out of bounds
    const float* src = rows + r * rw;
    float* d0 = o.obs + (long long)b * od;
    float* d1 = o.act + (long long)b * ad;
    float* d2 = o.nobs + (long long)b * od;
    const int e1 = od, e2 = od + ad, e3 = 2 * od + ad;
    for (int k = lane; k < e3 + 2; k += 32) {
      const float v = __ldg(src + k);
      if (k < e1) d0[k] = v;
      else if (k < e2) d1[k - e1] = v;
      else if (k < e3) d2[k - e2] = v;
      else if (k == e3) o.rew[b] = v;
      else o.done[b] = v;
    }
  }
}

// rows[dst[i]] = staged[i] for i < n: flush of device-side transitions (vectorised envs that
// already live on the GPU); host-side add_transition goes through plain contiguous copies.
__global__ void __launch_bounds__(256) replay_scatter_kernel(float* __restrict__ rows, long long rw, const long long* __restrict__ dst,
                                                             const float* __restrict__ staged, int n, long long cap) {
  SGRL_PDL_ENTER();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const long long r = dst[i];
    if (r < 0 || r >= cap) continue;
    float* d = rows + r * rw;
    const float* s = staged + (long long)i * rw;
    for (int k = lane; k < rw; k += 32) d[k] = __ldg(s + k);
  }
}

inline int replay_gather(const float* rows, long long rw, const long long* idx, int B, int od, int ad, long long cap, const ReplayOut& o,
                         cudaStream_t st) {
  if (B <= 0) return 0;
  int gx = ceil_div(B, 8); if (gx > 8 * NUM_SMS) gx = 8 * NUM_SMS;
  prof_begin(PC_OTHER, 8.0 * rw * B, st);
  launch_k(replay_gather_kernel, gx, 256, 0, st, rows, rw, idx, B, od, ad, cap, o);
  prof_end(st);
  SGRL_LAUNCH_OK();
  return 0;
}

inline int replay_scatter(float* rows, long long rw, const long long* dst, const float* staged, int n, long long cap, cudaStream_t st) {
  if (n <= 0) return 0;
  int gx = ceil_div(n, 8); if (gx > 8 * NUM_SMS) gx = 8 * NUM_SMS;
  launch_k(replay_scatter_kernel, gx, 256, 0, st, rows, rw, dst, staged, n, cap);
  SGRL_LAUNCH_OK();
  return 0;
}

}  // namespace sgrl
