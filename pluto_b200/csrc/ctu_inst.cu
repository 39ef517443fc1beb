// ctu_inst.cu -- one translation unit per (arithmetic namespace, solver) of the
// corner-transport-upwind sweeps:  nvcc -DPG_NS=pg_exact -DPG_SOLVER=0 -fmad=false ...
#include "ctu_kernels.cuh"

namespace PG_NS {
#if PG_SOLVER == 0
int launch_ctu_sweep_hlld (int dir, int phase, const CtuArgs &a, cudaStream_t s)
{ return launch_ctu_sweep_t<SOLVER_HLLD>(dir, phase, a, s); }
int launch_ctu_half (const CtuArgs &a, cudaStream_t s) { return launch_ctu_half_t (a, s); }
#elif PG_SOLVER == 1
int launch_ctu_sweep_hll (int dir, int phase, const CtuArgs &a, cudaStream_t s)
{ return launch_ctu_sweep_t<SOLVER_HLL>(dir, phase, a, s); }
#elif PG_SOLVER == 3
int launch_ctu_sweep_hllc (int dir, int phase, const CtuArgs &a, cudaStream_t s)
{ return launch_ctu_sweep_t<SOLVER_HLLC>(dir, phase, a, s); }
#elif PG_SOLVER == 4
int launch_ctu_sweep_tvdlf (int dir, int phase, const CtuArgs &a, cudaStream_t s)
{ return launch_ctu_sweep_t<SOLVER_TVDLF>(dir, phase, a, s); }
#else
int launch_ctu_sweep_roe (int dir, int phase, const CtuArgs &a, cudaStream_t s)
{ return launch_ctu_sweep_t<SOLVER_ROE>(dir, phase, a, s); }
#endif
}
