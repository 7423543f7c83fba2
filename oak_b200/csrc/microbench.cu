// microbench.cu — measured fp64 pipe peaks of the device (roofline denominators for the fp64-bound
// kernels; MEASURED_PEAKS.json only carries HBM and bf16 tensor figures).
//   mode 0: DFMA   — 16 independent register FMA chains per thread
//   mode 1: DMMA   — mma.sync.aligned.m8n8k4.row.col.f64 (the only fp64 tensor-core shape family on
//                    sm_100a; tcgen05 has no f64 kind), 8 independent accumulators per warp
#include "common.cuh"

namespace {

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_dfma(double *out, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = a + i + threadIdx.x * 1e-3;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = fma(x[i], b, a);
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dmma(double *out, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = a + i; c[i][1] = b - i; }
  const double fa = a + threadIdx.x * 1e-3, fb = b;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(fa), "d"(fb));
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

}  // namespace

int oak_fp64_peak(int mode, double *tflops) {
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double *out = nullptr;
  CUDA_TRY(cudaMalloc(&out, 8));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  const int blocks = sms * 8, threads = 256;
  double best = 0.;
  for (int rep = 0; rep < 5; rep++) {
    CUDA_TRY(cudaEventRecord(e0));
    if (mode == 0) k_dfma<<<blocks, threads>>>(out, 1.0, 0.999999);
    else k_dmma<<<blocks, threads>>>(out, 1.0, 0.999999);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    double flops;
    if (mode == 0) flops = 2.0 * 16 * ITERS * (double)blocks * threads;
    else flops = 2.0 * 8 * 8 * 4 * 8 * ITERS * (double)blocks * (threads / 32);
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return 0;
}
