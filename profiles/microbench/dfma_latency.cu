// DFMA dependent-issue latency and single-warp / multi-warp throughput on one SM (B200, sm_100a):
//   chains = independent accumulators per thread (ILP), warps = resident warps on the SM (TLP).
// Prints cycles per DFMA warp-instruction per SM sub-partition... see the table it writes.
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k(double *out, long long *cyc, const double a, const double b, const int iters)
{
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c)
    x[c] = threadIdx.x * 1e-3 + c;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i)
    {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < CH; ++c)
          x[c] = fma(x[c], a, b);
    }
  const long long t1 = clock64();
  double          s  = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c)
    s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0)
    cyc[blockIdx.x] = t1 - t0;
}

template <int CH>
void run(const int warps)
{
  double    *out;
  long long *cyc, h;
  cudaMalloc(&out, 2048 * sizeof(double));
  cudaMalloc(&cyc, sizeof(long long));
  const int iters = 2000;
  k<CH><<<1, warps * 32>>>(out, cyc, 0.999999, 1e-9, iters);
  k<CH><<<1, warps * 32>>>(out, cyc, 0.999999, 1e-9, iters);
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  const double per_chain_step = (double)h / (iters * 8.0);           // cycles per round of CH independent DFMAs
  const double sm_dfma_per_clk = warps * 32.0 * CH / per_chain_step; // DFMA lanes per clock on this SM
  printf("chains %2d warps %2d : %7.2f cycles per dependent step, %6.1f DFMA/clk/SM\n", CH, warps, per_chain_step, sm_dfma_per_clk);
  cudaFree(out);
  cudaFree(cyc);
}

int main()
{
  for (int w : {1, 4, 8, 12, 16, 32})
    {
      run<1>(w);
      run<2>(w);
      run<3>(w);
      run<4>(w);
      run<8>(w);
    }
  return 0;
}
