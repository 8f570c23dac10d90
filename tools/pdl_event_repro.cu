// Stand-alone probe of the ordering hazard that csrc/net.cuh works around with fence_kernel:
//   stream A:  K0 (programmatic launch) -> K1 (programmatic launch) -> cudaEventRecord(ev)
//   stream B:  cudaStreamWaitEvent(ev)  -> reader (plain launch) checks what K0 / K1 wrote
// Every kernel triggers launch_dependents at entry and executes griddepcontrol.wait before its first dependent access,
// like the library's kernels.  Variants: event directly after K1 / an empty non-programmatic kernel before the event /
// K1 without griddepcontrol.wait (a kernel that has nothing to wait for, e.g. a zero-fill).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o pdl_event_repro tools/pdl_event_repro.cu ; run: ./pdl_event_repro [iters]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void producer(int* buf, int n, int val, int spin, int do_wait) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (do_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) buf[i] = val;
}
__global__ void fence_kernel() {}
__global__ void reader(const int* b0, const int* b1, int n, int val, int* bad) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (b0[i] != val) atomicAdd(bad, 1);
    if (b1[i] != val) atomicAdd(bad + 1, 1);
  }
}

static void launch_pdl(cudaStream_t st, int* buf, int n, int val, int spin, int do_wait, int grid) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, producer, buf, n, val, spin, do_wait));
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000, n = 1 << 16;
  int *b0, *b1, *bad;
  CK(cudaMalloc(&b0, n * 4)); CK(cudaMalloc(&b1, n * 4)); CK(cudaMalloc(&bad, 8));
  cudaStream_t A, B;
  CK(cudaStreamCreateWithFlags(&A, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&B, cudaStreamNonBlocking));
  cudaEvent_t ev, back;
  CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&back, cudaEventDisableTiming));
  const char* names[4] = {"event right after K1 (K1 waits)", "fence kernel before the event (K1 waits)",
                          "event right after K1 (K1 has no griddepcontrol.wait)", "fence kernel before the event (K1 has no wait)"};
  for (int variant = 0; variant < 4; ++variant) {
    const int fence = variant & 1, k1_wait = variant < 2;
    int fails0 = 0, fails1 = 0;
    for (int it = 1; it <= iters; ++it) {
      CK(cudaMemsetAsync(bad, 0, 8, A));
      launch_pdl(A, b0, n, it, 40000 + (it % 7) * 6000, 1, 36);       // K0: ~20-40 us, 36 CTAs like an update-size GEMM
      launch_pdl(A, b1, n, it, 2000, k1_wait, 36);                    // K1: short
      if (fence) fence_kernel<<<1, 1, 0, A>>>();
      CK(cudaEventRecord(ev, A));
      CK(cudaStreamWaitEvent(B, ev, 0));
      reader<<<64, 256, 0, B>>>(b0, b1, n, it, bad);
      CK(cudaEventRecord(back, B));
      CK(cudaStreamWaitEvent(A, back, 0));                            // the next iteration's writers wait for the reader
      int h[2];
      CK(cudaMemcpyAsync(h, bad, 8, cudaMemcpyDeviceToHost, B));
      CK(cudaStreamSynchronize(B));
      fails0 += h[0] != 0; fails1 += h[1] != 0;
    }
    CK(cudaDeviceSynchronize());
    printf("%-58s iterations %d: reader saw stale K0 output %d times, stale K1 output %d times\n", names[variant], iters, fails0, fails1);
  }
  return 0;
}
