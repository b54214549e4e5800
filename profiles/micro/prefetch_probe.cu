// Does prefetch.global.L1 bring a line into the SM's L1 on sm_100a?  One warp, one CTA: for each mode,
// touch a fresh 256-byte row (8 bytes per lane) in some way, spin ~4000 cycles, then time a dependent load.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o prefetch_probe prefetch_probe.cu && ./prefetch_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }
__global__ void probe(const double* __restrict__ buf, long long* out, int trials, size_t stride) {
  __shared__ double dump[32];
  const int lane = threadIdx.x;
  for (int mode = 0; mode < 6; ++mode) {
    long long tot = 0;
    for (int t = 0; t < trials; ++t) {
      const double* p = buf + ((size_t)(mode * trials + t)) * stride + lane;
      if (mode == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      if (mode == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
      if (mode == 3) {
        unsigned s = (unsigned)__cvta_generic_to_shared(&dump[lane]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(p) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      if (mode == 4) { double v; asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p)); if (v == 123.456) out[63] = 1; }
      if (mode == 5) { double v; asm volatile("ld.global.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p)); if (v == 123.456) out[63] = 1; }
      long long t0 = clk();
      while (clk() - t0 < 4000) {}
      if (mode == 3) asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      long long a = clk();
      double v;
      asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
      long long sink = (long long)(v * 0.0);    // dependency on the loaded value
      long long bq = clk() + sink;
      tot += bq - a;
    }
    if (lane == 0) out[mode] = tot / trials;
  }
}
int main() {
  const size_t stride = 1 << 16;   // doubles between rows: 512 KB apart
  const int trials = 64;
  double* buf; long long* out;
  cudaMalloc(&buf, 6 * trials * stride * sizeof(double));
  cudaMemset(buf, 0, 6 * trials * stride * sizeof(double));
  cudaMallocManaged(&out, 64 * sizeof(long long));
  // flush the L2 with another buffer
  char* junk; cudaMalloc(&junk, 512u << 20); cudaMemset(junk, 1, 512u << 20);
  probe<<<1, 32>>>(buf, out, trials, stride);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  const char* names[6] = {"cold (nothing before)", "prefetch.global.L2", "prefetch.global.L1", "cp.async.ca to a dump row in smem",
                          "ld.global before (true L1 hit)", "ld.global.L1::evict_last before"};
  for (int m = 0; m < 6; ++m) printf("%-40s %lld cycles\n", names[m], out[m]);
  return 0;
}
