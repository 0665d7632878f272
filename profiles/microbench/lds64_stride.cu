// Shared-memory wavefronts of a warp-wide LDS.64 whose lanes are `stride` doubles apart (+ a common offset):
// cycles per instruction of a stream of independent loads, 4 warps of one CTA on one SM (so the LSU pipe, not the
// latency, is the limit).  Conflict-free 64-bit loads cost 2 wavefronts.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(double *out, long long *cyc, const int stride, const int offset, const int iters)
{
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x)
    sm[i] = i;
  __syncthreads();
  const int     lane = threadIdx.x % 32;
  const double *p    = sm + offset + lane * stride;
  double        a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i)
    {
      const volatile double *q = p + (i & 7);
      a0 += q[0], a1 += q[8], a2 += q[16], a3 += q[24], a4 += q[32], a5 += q[40], a6 += q[48], a7 += q[56];
    }
  const long long t1 = clock64();
  out[threadIdx.x]   = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0)
    *cyc = t1 - t0;
}

int main()
{
  double    *out;
  long long *cyc, h;
  cudaMalloc(&out, 1024 * sizeof(double));
  cudaMalloc(&cyc, sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const int iters = 4000;
  for (int stride : {1, 2, 3, 5, 9, 17, 25, 27, 29, 31, 33, 26, 28, 30})
    for (int offset : {0, 1})
      {
        k<<<1, 128, 8192 * 8 + 1024>>>(out, cyc, stride, offset, iters);
        k<<<1, 128, 8192 * 8 + 1024>>>(out, cyc, stride, offset, iters);
        cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("stride %2d offset %d : %6.2f cycles per LDS.64 warp instruction (4 warps)\n", stride, offset, (double)h / (iters * 8.0 * 4.0));
      }
  return 0;
}
