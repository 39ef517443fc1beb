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
#else
int launch_sweep_roe (int dir, int recon, const SweepArgs &a, cudaStream_t s, bool bf)
{ return launch_sweep_t<SOLVER_ROE>(dir, recon, a, s, bf); }
#ifdef PG_FAST                 // the fused x1+x2 sweep exists with FAST arithmetic only
int launch_sweep_xy_roe (int recon, const SweepArgs &a, cudaStream_t s, bool bf)
{ return launch_sweep_xy_t<SOLVER_ROE>(recon, a, s, bf); }
#endif
#endif
}
