// tools/quick_inst.cu -- the two kernels of the bench workload alone (fused x1+x2 sweep and x3 marching sweep, HLLD + PLM, 3-D):
// compiles in seconds for register / spill / SASS checks of a variant
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xptxas -v -fmad=true -DPG_NS=pg_fast -DPG_FAST=1 \
//        -DPG_SOLVER=0 -Ipluto_b200/csrc -c tools/quick_inst.cu -o build/quick/quick.o
#include "sweep_kernels.cuh"
namespace PG_NS {
template __global__ void sweep_xy_kernel<RECON_PLM, SOLVER_HLLD, 3, false, false, false, false, false>(const __grid_constant__ SweepArgs);
template __global__ void sweep_march_kernel<2, RECON_PLM, SOLVER_HLLD, 3, false, false, false, false>(const __grid_constant__ SweepArgs);
}
namespace PG_NS {
template __global__ void sweep_march_kernel<2, RECON_PLM, SOLVER_HLLD, 3, false, false, false, false, true>(const __grid_constant__ SweepArgs);
}
