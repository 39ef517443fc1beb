// FP64 issue microbenchmark (development aid): can integer / select / shared-memory
// instructions issue in the shadow of DFMAs on sm_100a, and what is the dependent
// DFMA latency?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_issue tools/ubench_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NI, int NS, int CHAINS>
__global__ void __launch_bounds__(1024) mix (double *out, int *iout, double a, double b, int iters, int k)
{
  __shared__ double sm[1024];
  double x[8]; int y[8];
  for (int q = 0; q < 8; q++){ x[q] = threadIdx.x*1e-9 + q; y[q] = threadIdx.x + q; }
  sm[threadIdx.x] = x[0];
  double acc = 0.0;
  long long t0 = clock64 ();
  for (int i = 0; i < iters; i++){
#pragma unroll
    for (int r = 0; r < 8; r++){
#pragma unroll
      for (int q = 0; q < NF; q++) x[q % CHAINS] = fma (x[q % CHAINS], a, b);
#pragma unroll
      for (int q = 0; q < NI; q++) y[q % 8] = (y[q % 8] ^ k) + y[(q + 1) % 8];
#pragma unroll
      for (int q = 0; q < NS; q++) acc += sm[(threadIdx.x + q*32 + r) & 1023];
    }
  }
  long long t1 = clock64 ();
  double s = acc; int z = 0;
  for (int q = 0; q < 8; q++){ s += x[q]; z += y[q]; }
  out[blockIdx.x*blockDim.x + threadIdx.x] = s;
  iout[blockIdx.x*blockDim.x + threadIdx.x] = z + (int)(t1 - t0);
  if (threadIdx.x == 0 && blockIdx.x == 0) ((long long *)out)[0] = t1 - t0;
}

template <int NF, int NI, int NS, int CHAINS>
static void run (const char *name, int warps_per_smsp)
{
  const int tpb = warps_per_smsp*4*32, nb = 148, iters = 2000;
  double *out; int *iout;
  cudaMalloc (&out, sizeof (double)*nb*tpb + 64); cudaMalloc (&iout, sizeof (int)*nb*tpb);
  mix<NF, NI, NS, CHAINS><<<nb, tpb>>>(out, iout, 0.999999, 1e-7, 10, 3);
  mix<NF, NI, NS, CHAINS><<<nb, tpb>>>(out, iout, 0.999999, 1e-7, iters, 3);
  cudaDeviceSynchronize ();
  long long cyc; cudaMemcpy (&cyc, out, 8, cudaMemcpyDeviceToHost);
  double per_iter = (double)cyc/(iters*8.0);
  printf ("%-28s warps/SMSP %d : %.2f cycles per group (%d DFMA + %d INT(2 ops) + %d LDS) -> %.2f cyc/DFMA/warp-slot\n",
          name, warps_per_smsp, per_iter, NF, NI, NS, per_iter/warps_per_smsp/(NF ? NF : 1));
  cudaFree (out); cudaFree (iout);
}

int main ()
{
  for (int w = 1; w <= 4; w++){
    run<8, 0, 0, 8>("dfma x8 (8 chains)", w);
    run<8, 0, 0, 1>("dfma x8 (1 chain)", w);
    run<8, 0, 0, 2>("dfma x8 (2 chains)", w);
    run<8, 4, 0, 8>("dfma x8 + 4 int pairs", w);
    run<8, 8, 0, 8>("dfma x8 + 8 int pairs", w);
    run<8, 0, 4, 8>("dfma x8 + 4 lds+dadd", w);
    run<0, 8, 0, 8>("8 int pairs", w);
    run<8, 8, 0, 2>("dfma x8(2ch) + 8 int pairs", w);
  }
  return 0;
}
