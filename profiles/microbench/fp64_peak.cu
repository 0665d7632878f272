// FP64 throughput microbenchmark for B200 (sm_100a): plain DFMA vs the DMMA
// (mma.sync f64) shapes.  MEASURED_PEAKS.json has no FP64 figure, and the
// assembly kernels' roofline is the FP64 pipe, so this provides the denominator.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double s)
{
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  double a = s, b = 1.0 - s;
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
  double r = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) r += acc[i];
  if (r == 12345.678) out[0] = r;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma884(double *out, int iters, double s)
{
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
  double a = s, b = 1.0 - s;
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int i = 0; i < NACC; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  double r = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) r += c[i][0] + c[i][1];
  if (r == 12345.678) out[0] = r;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma1684(double *out, int iters, double s)
{
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  double a0 = s, a1 = s * 0.5, b = 1.0 - s;
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int i = 0; i < NACC; ++i)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b));
    }
  double r = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) r += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (r == 12345.678) out[0] = r;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma1688(double *out, int iters, double s)
{
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  double a0 = s, a1 = s * 0.5, a2 = s * 0.25, a3 = s * 0.125, b0 = 1.0 - s, b1 = 0.5 - s;
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int i = 0; i < NACC; ++i)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
  double r = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) r += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (r == 12345.678) out[0] = r;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma16816(double *out, int iters, double s)
{
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = s / (i + 1);
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = 1.0 - s / (i + 1);
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int i = 0; i < NACC; ++i)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                       "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
  double r = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) r += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (r == 12345.678) out[0] = r;
}

template <class F>
static float time_it(F f)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep)
    {
      cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
  return best;
}

int main()
{
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device: %s, %d SMs, cc %d.%d\n", p.name, p.multiProcessorCount, p.major, p.minor);
  double *out; CK(cudaMalloc(&out, 8));
  const int iters = 20000, blocks = p.multiProcessorCount * 4, threads = 256;
  const double nthr = (double)blocks * threads, nwarp = nthr / 32;
#define RUN(name, kern, nacc, flop_per_call) { \
    float ms = time_it([&] { kern<nacc><<<blocks, threads>>>(out, iters, 0.5); }); \
    CK(cudaGetLastError()); \
    printf("{\"kernel\": \"%s\", \"nacc\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", name, nacc, ms, (flop_per_call) * (double)iters * nacc / (ms * 1e-3) / 1e12); }
  RUN("dfma", k_dfma, 8, 2.0 * nthr);
  RUN("dfma", k_dfma, 16, 2.0 * nthr);
  RUN("dmma.m8n8k4", k_dmma884, 4, 2.0 * 8 * 8 * 4 * nwarp);
  RUN("dmma.m8n8k4", k_dmma884, 8, 2.0 * 8 * 8 * 4 * nwarp);
  RUN("dmma.m16n8k4", k_dmma1684, 4, 2.0 * 16 * 8 * 4 * nwarp);
  RUN("dmma.m16n8k4", k_dmma1684, 8, 2.0 * 16 * 8 * 4 * nwarp);
  RUN("dmma.m16n8k8", k_dmma1688, 4, 2.0 * 16 * 8 * 8 * nwarp);
  RUN("dmma.m16n8k8", k_dmma1688, 8, 2.0 * 16 * 8 * 8 * nwarp);
  RUN("dmma.m16n8k16", k_dmma16816, 4, 2.0 * 16 * 8 * 16 * nwarp);
  RUN("dmma.m16n8k16", k_dmma16816, 8, 2.0 * 16 * 8 * 16 * nwarp);
  return 0;
}
