// cuda_runtime.h (EMULATION SHIM, TEST INFRASTRUCTURE ONLY)
//
// Lets g++ compile the product's CUDA sources (pluto_b200/csrc/*.cu, *.cuh) unchanged
// into a host library in which every kernel launch is executed by a lock-step
// interpreter of the CUDA execution model: one fibre per thread of a block, warp
// collectives (__shfl_*_sync, __ballot_sync, __syncwarp) and __syncthreads as
// rendezvous points, blocks one after the other.  It exists so that the kernels'
// index logic and arithmetic can be checked against the oracle on the CPU-only
// development box, bit for bit, before GPU time is spent.  It is NOT a CPU path of
// the product: nothing under pluto_b200/ refers to it, it is built only by
// tests/emu/build_emu.py into tests/emu/_build/, and only tests load it.
#pragma once
#define PG_EMU 1
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3 (unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x (x_), y (y_), z (z_) {}
};
extern uint3 threadIdx, blockIdx;
extern dim3  blockDim, gridDim;

// ---- runtime API subset used by pluto_gpu.cu -------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
typedef void *cudaGraph_t;
typedef void *cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaStreamCaptureModeThreadLocal = 1, cudaHostRegisterDefault = 0 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { int type; };
struct cudaDeviceProp { int multiProcessorCount; };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };

inline const char *cudaGetErrorString (cudaError_t e) { return e == cudaSuccess ? "no error" : "emulation: unsupported call"; }
inline cudaError_t cudaGetLastError () { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount (int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice (int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice (int) { return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc (T **p, size_t n) { *p = (T *)calloc (n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorEmu; }
template <class T> inline cudaError_t cudaMallocHost (T **p, size_t n) { *p = (T *)calloc (n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorEmu; }
inline cudaError_t cudaFree (void *p) { free (p); return cudaSuccess; }
inline cudaError_t cudaFreeHost (void *p) { free (p); return cudaSuccess; }
inline cudaError_t cudaMemset (void *p, int v, size_t n) { memset (p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync (void *p, int v, size_t n, cudaStream_t) { memset (p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy (void *d, const void *s, size_t n, int) { memmove (d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync (void *d, const void *s, size_t n, int, cudaStream_t) { memmove (d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags (cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize (cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamDestroy (cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate (cudaEvent_t *e) { *e = (void *)1; return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventCreateWithFlags (cudaEvent_t *e, unsigned) { *e = (void *)1; return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent (cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventDestroy (cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord (cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize (cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime (float *ms, cudaEvent_t, cudaEvent_t) { *ms = 1.f; return cudaSuccess; }
// launches run at once, so a "captured" step has already happened: graphs are refused and the
// host code is run with PLUTO_GPU_NO_GRAPH=1
inline cudaError_t cudaStreamBeginCapture (cudaStream_t, int) { return cudaErrorEmu; }
inline cudaError_t cudaStreamEndCapture (cudaStream_t, cudaGraph_t *g) { *g = nullptr; return cudaErrorEmu; }
inline cudaError_t cudaGraphInstantiate (cudaGraphExec_t *, cudaGraph_t, int) { return cudaErrorEmu; }
inline cudaError_t cudaGraphDestroy (cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy (cudaGraphExec_t) { return cudaSuccess; }
inline cudaError_t cudaGraphLaunch (cudaGraphExec_t, cudaStream_t) { return cudaErrorEmu; }
inline cudaError_t cudaPointerGetAttributes (cudaPointerAttributes *a, const void *) { a->type = cudaMemoryTypeHost; return cudaSuccess; }
inline cudaError_t cudaHostRegister (void *, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister (void *) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties (cudaDeviceProp *p, int) { p->multiProcessorCount = 2; return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute (F, int, int) { return cudaSuccess; }
template <class F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor (int *n, F, int, size_t) { *n = 3; return cudaSuccess; }

// ---- the interpreter ---------------------------------------------------------------------
namespace pg_emu {
  void launch_impl (dim3 grid, dim3 block, size_t smem, const std::function<void ()> &body);
  template <class F> inline void launch (F body, dim3 grid, dim3 block, size_t smem = 0, cudaStream_t = nullptr)
  { launch_impl (grid, block, smem, std::function<void ()>(body)); }
  void *dyn_smem ();
  uint64_t exchange (uint64_t mine, int src_lane);       // warp rendezvous: publish, wait for the warp, read src
  unsigned ballot (int pred);
  void barrier_block ();
  int lane_id ();
  // cp.async model: copies are queued per thread in commit groups and land at the wait that covers them
  void async_copy8 (double *dst, const double *src);
  void async_commit ();
  void async_wait (int keep_groups);
}

// ---- device intrinsics -------------------------------------------------------------------
template <class T> inline T __ldg (const T *p) { return *p; }
template <class T> inline T pg_emu_shfl (T v, int src)
{
  static_assert (sizeof (T) <= 8, "shuffle of at most 8 bytes");
  uint64_t bits = 0; memcpy (&bits, &v, sizeof (T));
  bits = pg_emu::exchange (bits, src);
  T r; memcpy (&r, &bits, sizeof (T));
  return r;
}
template <class T> inline T __shfl_sync (unsigned, T v, int src) { return pg_emu_shfl (v, src & 31); }
template <class T> inline T __shfl_xor_sync (unsigned, T v, int m) { return pg_emu_shfl (v, pg_emu::lane_id () ^ m); }
template <class T> inline T __shfl_up_sync (unsigned, T v, int d)
{ const int l = pg_emu::lane_id (); return pg_emu_shfl (v, l - d >= 0 ? l - d : l); }
template <class T> inline T __shfl_down_sync (unsigned, T v, int d)
{ const int l = pg_emu::lane_id (); return pg_emu_shfl (v, l + d <= 31 ? l + d : l); }
inline unsigned __ballot_sync (unsigned, int pred) { return pg_emu::ballot (pred); }
inline void __syncwarp (unsigned = 0xffffffffu) { pg_emu::exchange (0, pg_emu::lane_id ()); }
inline void __syncthreads () { pg_emu::barrier_block (); }
inline int __popc (unsigned x) { return __builtin_popcount (x); }
inline unsigned long long atomicAdd (unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline unsigned long long atomicMax (unsigned long long *p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
inline long long __double_as_longlong (double x) { long long r; memcpy (&r, &x, 8); return r; }
inline double __longlong_as_double (long long x) { double r; memcpy (&r, &x, 8); return r; }
inline double __ddiv_rn (double a, double b) { return a/b; }
inline double __dsqrt_rn (double a) { return sqrt (a); }
inline void __threadfence_system () {}
inline unsigned __float_as_uint (float f) { unsigned u; memcpy (&u, &f, 4); return u; }
inline unsigned __byte_perm (unsigned x, unsigned y, unsigned s)
{
  const unsigned long long v = ((unsigned long long)y << 32) | x;
  unsigned r = 0;
  for (int q = 0; q < 4; q++) r |= (unsigned)((v >> (8*((s >> (4*q)) & 7))) & 0xff) << (8*q);
  return r;
}
inline long long clock64 () { return 0; }
inline double __dmul_rn (double a, double b) { return a*b; }
inline double __dadd_rn (double a, double b) { return a + b; }
inline float __fdividef (float a, float b) { return a/b; }
inline size_t __cvta_generic_to_shared (const void *p) { return (size_t)p; }
