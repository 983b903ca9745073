// FP64 pipe microbenchmark for the roofline denominators MEASURED_PEAKS.json does not hold:
// DADD / DMUL / DFMA issue rate of the whole chip (all SMs, enough warps to saturate), the dependent-issue
// latency of one warp, and the throughput of one warp as a function of its ILP.  Prints one JSON line.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP>
__global__ void k_tput(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = a + threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (OP == 0) x[i] = __dadd_rn(x[i], b);
      else if (OP == 1) x[i] = __dmul_rn(x[i], b);
      else x[i] = __fma_rn(x[i], b, a);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  if (s == 123.456) out[0] = s;
}

template <int OP, int ILP>
static double run(int grid, int block, int iters, double* d_out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_tput<OP, ILP><<<grid, block>>>(d_out, iters, 1.0, 1.0000001);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0);
    k_tput<OP, ILP><<<grid, block>>>(d_out, iters, 1.0, 1.0000001);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return (double)grid * block * iters * ILP / (best * 1e-3);   // lane-ops per second
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double* d; cudaMalloc(&d, 64);
  const int it = 1 << 14;
  double add = run<0, 8>(sms * 8, 256, it, d), mul = run<1, 8>(sms * 8, 256, it, d), fma = run<2, 8>(sms * 8, 256, it, d);
  // one warp per SM sub-partition (4 warps per SM), varying ILP: cycles per dependent op = clock / (ops/s per lane)
  double w1 = run<0, 1>(sms, 128, it, d), w2 = run<0, 2>(sms, 128, it, d), w3 = run<0, 3>(sms, 128, it, d), w4 = run<0, 4>(sms, 128, it, d),
         w8 = run<0, 8>(sms, 128, it, d);
  double one = run<0, 1>(1, 32, it, d);
  double clk = clk_khz * 1e3;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f, \"dadd_tops\": %.3f, \"dmul_tops\": %.3f, \"dfma_tinst\": %.3f, "
         "\"dadd_per_clk_per_sm\": %.2f, \"dependent_dadd_latency_cycles\": %.2f, "
         "\"one_warp_per_smsp_frac_of_peak\": {\"ilp1\": %.3f, \"ilp2\": %.3f, \"ilp3\": %.3f, \"ilp4\": %.3f, \"ilp8\": %.3f}}\n",
         p.name, sms, clk / 1e6, add / 1e12, mul / 1e12, fma / 1e12, add / clk / sms, clk / (one / 32.0),
         w1 / add, w2 / add, w3 / add, w4 / add, w8 / add);
  return 0;
}
