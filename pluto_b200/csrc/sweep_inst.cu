// sweep_inst.cu -- one translation unit per (arithmetic namespace, solver):
//   nvcc -DPG_NS=pg_exact -DPG_SOLVER=0 -fmad=false ...
// so the heavy FP64 kernels compile in parallel.
#include "sweep_kernels.cuh"

namespace PG_NS {
#if PG_SOLVER == 0
int launch_sweep_hlld (int dir, int recon, const SweepArgs &a, cudaStream_t s, bool bf)
{ return launch_sweep_t<SOLVER_HLLD>(dir, recon, a, s, bf); }
#ifdef PG_FAST                 // the fused x1+x2 sweep exists with FAST arithmetic only
int launch_sweep_xy_hlld (int recon, const SweepArgs &a, cudaStream_t s, bool bf)
{ return launch_sweep_xy_t<SOLVER_HLLD>(recon, a, s, bf); }
#endif
#elif PG_SOLVER == 1
int launch_sweep_hll (int dir, int recon, const SweepArgs &a, cudaStream_t s, bool bf)
{ return launch_sweep_t<SOLVER_HLL>(dir, recon, a, s, bf); }
#ifdef PG_FAST                 // the fused x1+x2 sweep exists with FAST arithmetic only
int launch_sweep_xy_hll (int recon, const SweepArgs &a, cudaStream_t s, bool bf)
{ return launch_sweep_xy_t<SOLVER_HLL>(recon, a, s, bf); }
#endif
#elif PG_SOLVER == 3
int launch_sweep_hllc (int dir, int recon, const SweepArgs &a, cudaStream_t s, bool bf)
{ return launch_sweep_t<SOLVER_HLLC>(dir, recon, a, s, bf); }
#ifdef PG_FAST
int launch_sweep_xy_hllc (int recon, const SweepArgs &a, cudaStream_t s, bool bf)
{ return launch_sweep_xy_t<SOLVER_HLLC>(recon, a, s, bf); }
#endif
#elif PG_SOLVER == 4
int launch_sweep_tvdlf (int dir, int recon, const SweepArgs &a, cudaStream_t s, bool bf)
{ return launch_sweep_t<SOLVER_TVDLF>(dir, recon, a, s, bf); }
#ifdef PG_FAST
int launch_sweep_xy_tvdlf (int recon, const SweepArgs &a, cudaStream_t s, bool bf)
{ return launch_sweep_xy_t<SOLVER_TVDLF>(recon, a, s, bf); }
#endif
#else
#ifdef PG_FAST
// self-test of the branch-free IEEE division / square root of this unit (mhd_device.cuh) against div.rn.f64 / sqrt.rn.f64:
// n samples per thread, mantissas from a counter-based generator mixed with adversarial patterns (all ones, one bit, 1 + ulp)
__global__ void arith_selftest_kernel (unsigned long long seed, int n, unsigned long long *bad)
{
  unsigned long long s = seed + 0x9E3779B97F4A7C15ull*((unsigned long long)blockIdx.x*blockDim.x + threadIdx.x + 1);
  auto next = [&] (){ s += 0x9E3779B97F4A7C15ull; unsigned long long z = s; z = (z ^ (z >> 30))*0xBF58476D1CE4E5B9ull;
                      z = (z ^ (z >> 27))*0x94D049BB133111EBull; return z ^ (z >> 31); };
  auto mk = [&] (unsigned long long r, int q){
    unsigned long long m = r & 0xFFFFFFFFFFFFFull;
    if      ((q & 7) == 1) m = 0xFFFFFFFFFFFFFull ^ (r & 7);          // mantissa (almost) all ones
    else if ((q & 7) == 2) m = 1ull << (r % 52);                     // a single bit
    else if ((q & 7) == 3) m = r & 0xFFull;                          // just above a power of two
    else if ((q & 7) == 4) m &= ~0x3FFFFFFull;                       // short mantissas: exact and near-tie quotients
    const unsigned long long e = 1023ull - 40 + (r >> 52) % 81;      // 2^-40 .. 2^40
    return __longlong_as_double ((long long)((e << 52) | m));
  };
  unsigned long long nd = 0, ns = 0, nr = 0;
  for (int q = 0; q < n; q++){
    const double a = mk (next (), q >> 3), b = mk (next (), q);
    if (__double_as_longlong (pg_div (a, b)) != __double_as_longlong (__ddiv_rn (a, b))) nd++;
    if (__double_as_longlong (pg_rcp (b)) != __double_as_longlong (__ddiv_rn (1.0, b))) nr++;
    if (__double_as_longlong (pg_sqrt (b)) != __double_as_longlong (__dsqrt_rn (b))) ns++;
  }
  if (nd) atomicAdd (bad, nd);
  if (nr) atomicAdd (bad + 1, nr);
  if (ns) atomicAdd (bad + 2, ns);
}
int launch_arith_selftest (unsigned long long seed, int nblocks, int n, unsigned long long *bad, cudaStream_t s)
{
  arith_selftest_kernel<<<nblocks, 256, 0, s>>>(seed, n, bad);
  return pg_launch_status ();
}
#endif
int launch_sweep_roe (int dir, int recon, const SweepArgs &a, cudaStream_t s, bool bf)
{ return launch_sweep_t<SOLVER_ROE>(dir, recon, a, s, bf); }
#ifdef PG_FAST                 // the fused x1+x2 sweep exists with FAST arithmetic only
int launch_sweep_xy_roe (int recon, const SweepArgs &a, cudaStream_t s, bool bf)
{ return launch_sweep_xy_t<SOLVER_ROE>(recon, a, s, bf); }
#endif
#endif
}
