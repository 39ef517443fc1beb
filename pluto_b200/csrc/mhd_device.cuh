// mhd_device.cuh -- per-interface / per-zone device arithmetic of the ideal-MHD
// Godunov step: limiters, PLM / PPM interface states, prim<->cons mappers,
// flux, wave speeds, HLLD / HLL / Roe.  Everything is FP64 and lives in
// registers: the `double q[8]` arrays are indexed by compile-time constants
// only (DIR and NC are template parameters, loops are fully unrolled).
//
// The operation order follows the reference so that the EXACT build
// (-fmad=false) reproduces its IEEE results bit for bit:
//   limiters        Src/States/plm_coeffs.h:72-123, Src/macros.h:140-151
//   PLM             Src/States/plm_states.c:134-275 (uniform Cartesian path)
//   PPM             Src/States/ppm_states.c:146-207, ppm_coeffs.c:500-547
//   mappers         Src/MHD/mappers.c:25-86, 88-254
//   flux            Src/MHD/fluxes.c:159-215
//   fast speed      Src/MHD/eigenv.c:35-104
//   Davis speeds    Src/MHD/hll_speed.c:76-107
//   HLLD            Src/MHD/hlld.c:98-427
//   HLL             Src/MHD/hll.c:101-135
//   Roe             Src/MHD/roe.c:98-700
#pragma once
#include <math.h>

// PG_FAST selects the FAST kernel STRUCTURE (fused x1+x2 sweep, stage-2 U rebuilt from V);
// PG_FAST_MATH the FAST ARITHMETIC of this file (MUFU-seeded division / square root, FMA forms,
// the flux-form HLLD).  The Roe translation units (PG_SOLVER == 2) keep the reference's IEEE
// arithmetic: Roe's eigenvector normalisation switches on comparisons that round-off decides
// where the transverse field vanishes exactly (roe.c:336-364, see riemann_roe), so the states
// and averages that feed them must be the reference's own, bit for bit.
#if defined(PG_FAST) && !(defined(PG_SOLVER) && PG_SOLVER == 2)
#define PG_FAST_MATH 1
#endif

namespace PG_NS {

enum { RHO = 0, VX1 = 1, VX2 = 2, VX3 = 3, BX1 = 4, BX2 = 5, BX3 = 6, PRS = 7, NV = 8 };
enum { MX1 = VX1, MX2 = VX2, MX3 = VX3, ENG = PRS };
enum { RECON_PLM = 0, RECON_PPM = 1, RECON_PLMW = 2 };      // PLMW: linear with the grid-dependent weights of UNIFORM_CARTESIAN_GRID NO
enum { SOLVER_HLLD = 0, SOLVER_HLL = 1, SOLVER_ROE = 2, SOLVER_HLLC = 3, SOLVER_TVDLF = 4 };

struct Phys {                          // same layout as PhysPar (kernels_common.cuh)
  double gamma, gmm1, small_dn, small_pr;
  double igmm1;                       // 1/(gamma - 1), FAST arithmetic only
};

// direction bookkeeping (reference Src/set_indexes.c:49-123)
template <int DIR> struct Dirs;
template <> struct Dirs<0> { enum { vn = VX1, vt = VX2, vb = VX3, bn = BX1, bt = BX2, bb = BX3 }; };
template <> struct Dirs<1> { enum { vn = VX2, vt = VX1, vb = VX3, bn = BX2, bt = BX1, bb = BX3 }; };
template <> struct Dirs<2> { enum { vn = VX3, vt = VX1, vb = VX2, bn = BX3, bt = BX1, bb = BX2 }; };

// a variable slot is carried when NC == 3 or it is not a third component
template <int NC> __device__ __forceinline__ constexpr bool live (int nv)
{ return NC == 3 || (nv != VX3 && nv != BX3); }

#define PG_UNROLL _Pragma("unroll")
#define PG_FOR_NV(nv) PG_UNROLL for (int nv = 0; nv < NV; nv++) if (live<NC>(nv))
// the same without slot SKIP: the reconstructed cell-centred field component NORMAL to a sweep is never used (the
// interface states take the staggered field, plm_states.c:271-275), so the marching sweeps leave it out (SKIP = -1: none)
#define PG_FOR_NV_SKIP(nv, SKIP) PG_UNROLL for (int nv = 0; nv < NV; nv++) if (live<NC>(nv) && nv != (SKIP))

__device__ __forceinline__ double maxv (double a, double b) { return a >= b ? a : b; }
__device__ __forceinline__ double minv (double a, double b) { return a <= b ? a : b; }
__device__ __forceinline__ double abs_min (double a, double b) { return fabs(a) < fabs(b) ? a : b; }
__device__ __forceinline__ double minmod (double a, double b)
{ return a*b > 0.0 ? (fabs(a) < fabs(b) ? a : b) : 0.0; }

// ---------------------------------------------------------------------------
//  division and square root.  EXACT: IEEE (div.rn.f64 / sqrt.rn.f64), the
//  reference's results bit for bit.  FAST: branch-free MUFU seed
//  (rcp/rsqrt.approx.ftz.f64, >= 20 good bits) + one cubically convergent step in
//  FMA form, accurate to 1-2 ulp, no slow-path branch -> the scheduler can interleave
//  independent chains.  Arguments are physical magnitudes far from the
//  subnormal / overflow range.
//    fq_*  the FAST family itself (exists in every PG_FAST unit),
//    pg_*  what the kernels call: fq_* with PG_FAST_MATH, IEEE otherwise,
//    pq_*  fq_* in every PG_FAST unit, IEEE otherwise: the part of the Roe solver
//          downstream of its eigenvector switches.
// ---------------------------------------------------------------------------
#if defined(PG_FAST) && defined(PG_HOST_EMU)
// host build of the FAST algebra (tools/host_hlld_check.cpp, tests/emu): the quotient is formed
// as a*(1/b) and the root as x*(1/sqrt(x)) -- two roundings, the 1-2 ulp of the device iterations
__device__ __forceinline__ double fq_rcp (double b) { return 1.0/b; }
__device__ __forceinline__ double fq_rsqrt (double x) { return 1.0/sqrt (x); }
__device__ __forceinline__ float pg_sqrtf (float x) { return sqrtf (x); }
#elif defined(PG_FAST)
__device__ __forceinline__ float pg_sqrtf (float x)          // no slow-path call
{
  float y;
  asm ("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ double fq_rcp (double b)
{
  // MUFU seed (>= 20 good bits), then ONE cubically convergent step:
  // r (1 + e + e^2), e = 1 - b r  ->  relative error e^3 <= 2^-60
  double r;
  asm ("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  const double e = fma (-b, r, 1.0);
  return fma (r, fma (e, e, e), r);
}
// 1/sqrt(x): MUFU seed, then one cubically convergent step
// y (1 + e/2 + 3 e^2/8), e = 1 - x y^2  ->  relative error ~ (5/16) e^3 <= 2^-60
__device__ __forceinline__ double fq_rsqrt (double x)
{
  double y;
  asm ("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma (-x*y, y, 1.0);
  return fma (y*e, fma (0.375, e, 0.5), y);
}
#endif
#ifdef PG_FAST
__device__ __forceinline__ double fq_div (double a, double b) { return a*fq_rcp (b); }
__device__ __forceinline__ double fq_sqrt (double x)
{
  const double s = x*fq_rsqrt (x);
  return x > 0.0 ? s : 0.0;
}
__device__ __forceinline__ double pq_rcp (double b) { return fq_rcp (b); }
__device__ __forceinline__ double pq_div (double a, double b) { return fq_div (a, b); }
__device__ __forceinline__ double pq_sqrt (double x) { return fq_sqrt (x); }
#else
__device__ __forceinline__ double pq_rcp (double b) { return 1.0/b; }
__device__ __forceinline__ double pq_div (double a, double b) { return a/b; }
__device__ __forceinline__ double pq_sqrt (double x) { return sqrt (x); }
#endif
#ifdef PG_FAST_MATH
__device__ __forceinline__ double pg_rcp (double b) { return fq_rcp (b); }
__device__ __forceinline__ double pg_div (double a, double b) { return fq_div (a, b); }
__device__ __forceinline__ double pg_rsqrt (double x) { return fq_rsqrt (x); }
__device__ __forceinline__ double pg_sqrt (double x) { return fq_sqrt (x); }
__device__ __forceinline__ double pg_sqrt_pos (double x) { return x*fq_rsqrt (x); }    // x > 0 guaranteed
// sqrt(x) and 1/sqrt(x) together
__device__ __forceinline__ void pg_sqrt_rsqrt (double x, double &s, double &rs)
{
  rs = fq_rsqrt (x);
  s = x*rs;
}
#elif defined(PG_FAST) && !defined(PG_HOST_EMU)
// The Roe units of the FAST library: IEEE results (correctly rounded quotient / root, what div.rn.f64 / sqrt.rn.f64 give),
// but branch-free -- no exponent-range test, no slow-path call, so the surrounding code keeps its registers and schedule.
// Reciprocal: MUFU seed -> one cubic step (2^-60) -> one Markstein step y + y (1 - b y): correctly rounded except when b's
// mantissa is all ones (1/b then lies just above a rounding tie that the iteration resolves downwards): that one pattern is
// patched by an integer increment.  Quotient: q = a y, then twice  r = a - b q (exact, FMA), q <- RN(q + r y): the first pass
// makes q faithful, the second correctly rounded (Markstein 1990; Muller et al., Handbook of Floating-Point Arithmetic, 4.7).
// Root: s = x y, r = x - s^2 exact, s' = RN(s + r y/2).  Arguments are physical magnitudes far from the subnormal / overflow
// range.  pluto_gpu_selftest_arith compares all three with div.rn / sqrt.rn on the device over random and adversarial
// mantissas (tests/test_gpu_parity.py).
__device__ __forceinline__ double pg_rcp (double b)
{
  double y;
  asm ("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
  double e = fma (-b, y, 1.0);
  y = fma (y, fma (e, e, e), y);
  e = fma (-b, y, 1.0);
  y = fma (y, e, y);
  const long long ones = 0x000FFFFFFFFFFFFFLL;
  if ((__double_as_longlong (b) & ones) == ones) y = __longlong_as_double (__double_as_longlong (y) + 1);
  return y;
}
__device__ __forceinline__ double pg_div (double a, double b)
{
  const double y = pg_rcp (b);
  double q = a*y;
  double r = fma (-b, q, a);
  q = fma (r, y, q);
  r = fma (-b, q, a);
  return fma (r, y, q);
}
__device__ __forceinline__ double pg_sqrt (double x)
{
  double y;
  asm ("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma (-x*y, y, 1.0);
  y = fma (y*e, fma (0.375, e, 0.5), y);
  double s = x*y;
  const double r = fma (-s, s, x);
  s = fma (r, 0.5*y, s);
  return x > 0.0 ? s : (x < 0.0 ? __longlong_as_double (0x7ff8000000000000LL) : x);     // sqrt(+-0) = +-0, sqrt(< 0) = NaN
}
#else
__device__ __forceinline__ double pg_rcp (double b) { return 1.0/b; }
__device__ __forceinline__ double pg_div (double a, double b) { return a/b; }
__device__ __forceinline__ double pg_sqrt (double x) { return sqrt (x); }
#endif

// ---------------------------------------------------------------------------
//  limited slope, LIMITER DEFAULT: MC on density, minmod on pressure, van Leer
//  on velocity and field (plm_states.c:192-227)
// ---------------------------------------------------------------------------
template <int NV_ID> __device__ __forceinline__ double plm_slope (double dvp, double dvm)
{
  if (NV_ID == RHO){
    if (dvp*dvm > 0.0){
      double qc = 0.5*(dvm + dvp), scrh = 2.0*abs_min(dvp, dvm);
      return abs_min(qc, scrh);
    }
    return 0.0;
  }else if (NV_ID == PRS){
    return dvp*dvm > 0.0 ? abs_min(dvp, dvm) : 0.0;
  }else{
    return dvp*dvm > 0.0 ? pg_div (2.0*dvp*dvm, dvp + dvm) : 0.0;
  }
}

// one limiter for every variable (LIMITER != DEFAULT, plm_states.c:234-236; macros of
// plm_coeffs.h:72-123 on a uniform Cartesian grid), in the reference's operation order.
// lim: 1 flat, 2 minmod, 3 van Albada, 4 OSPRE, 5 UMIST, 6 van Leer, 7 MC
__device__ __forceinline__ double single_limiter (int lim, double dvp, double dvm)
{
  if (lim == 1 || !(dvp*dvm > 0.0)) return 0.0;
  if (lim == 2) return abs_min (dvp, dvm);
  if (lim == 3){
    const double dpp = dvp*dvp, dmm = dvm*dvm;
    return pg_div (dvp*(dmm + 1.e-18) + dvm*(dpp + 1.e-18), dpp + dmm + 1.e-18);
  }
  if (lim == 4) return pg_div (1.5*dvp*dvm*(dvm + dvp), dvp*dvp + dvm*dvm + dvp*dvm);
  if (lim == 5){
    const double ddp = 0.25*(dvp + 3.0*dvm), ddm = 0.25*(dvm + 3.0*dvp);
    double d2 = 2.0*abs_min (dvp, dvm);
    d2 = abs_min (d2, ddp);
    return abs_min (d2, ddm);
  }
  if (lim == 6) return pg_div (2.0*dvp*dvm, dvp + dvm);
  const double qc = 0.5*(dvm + dvp), scrh = 2.0*abs_min (dvp, dvm);
  return abs_min (qc, scrh);
}
template <int NC, int SKIP = -1>
__device__ __forceinline__ void plm_zone_single (int lim, const double *v, const double *dvm, const double *dvp,
                                                 double *vp, double *vm)
{
  PG_FOR_NV_SKIP(nv, SKIP){
    const double dvl = single_limiter (lim, dvp[nv], dvm[nv]);
    vp[nv] = v[nv] + dvl*0.5;
    vm[nv] = v[nv] - dvl*0.5;
  }
}

// vp = v + dvl/2, vm = v - dvl/2 for one zone from its two one-sided differences
#ifdef PG_FAST_MATH
// FAST: the HALF slope h = dvl/2 is formed directly (the factors 2 and 1/2 of the
// van Leer and MC limiters cancel against it; scaling by powers of two is exact,
// so h is the reference's dvl*0.5 up to the rounding of the division)
template <int NV_ID> __device__ __forceinline__ void plm_half (double v, double dvp, double dvm, double &vp, double &vm)
{
  const double p = dvp*dvm, s = dvp + dvm;
  if (NV_ID == RHO){          // MC: abs_min(0.5 (dvm+dvp), 2 abs_min(dvp, dvm))
    const double q = 0.25*s, m = abs_min (dvp, dvm);
    double h = abs_min (q, m);
    h = p > 0.0 ? h : 0.0;
    vp = v + h; vm = v - h;
  }else if (NV_ID == PRS){    // minmod
    double h = 0.5*abs_min (dvp, dvm);
    h = p > 0.0 ? h : 0.0;
    vp = v + h; vm = v - h;
  }else{                      // van Leer: 2 dvp dvm/(dvp + dvm)
    double r = pg_rcp (s);
    r = p > 0.0 ? r : 0.0;
    vp = fma (p, r, v); vm = fma (-p, r, v);
  }
}
template <int NC, int SKIP = -1>
__device__ __forceinline__ void plm_zone_default (const double *v, const double *dvm, const double *dvp,
                                                  double *vp, double *vm)
{
  plm_half<RHO>(v[RHO], dvp[RHO], dvm[RHO], vp[RHO], vm[RHO]);
  plm_half<VX1>(v[VX1], dvp[VX1], dvm[VX1], vp[VX1], vm[VX1]);
  plm_half<VX1>(v[VX2], dvp[VX2], dvm[VX2], vp[VX2], vm[VX2]);
  if (NC == 3) plm_half<VX1>(v[VX3], dvp[VX3], dvm[VX3], vp[VX3], vm[VX3]);
  if (SKIP != BX1) plm_half<VX1>(v[BX1], dvp[BX1], dvm[BX1], vp[BX1], vm[BX1]);
  if (SKIP != BX2) plm_half<VX1>(v[BX2], dvp[BX2], dvm[BX2], vp[BX2], vm[BX2]);
  if (NC == 3 && SKIP != BX3) plm_half<VX1>(v[BX3], dvp[BX3], dvm[BX3], vp[BX3], vm[BX3]);
  plm_half<PRS>(v[PRS], dvp[PRS], dvm[PRS], vp[PRS], vm[PRS]);
}
#else
template <int NC, int SKIP = -1>
__device__ __forceinline__ void plm_zone_default (const double *v, const double *dvm, const double *dvp,
                                                  double *vp, double *vm)
{
  double dvl[NV];
  dvl[RHO] = plm_slope<RHO>(dvp[RHO], dvm[RHO]);
  dvl[VX1] = plm_slope<VX1>(dvp[VX1], dvm[VX1]);
  dvl[VX2] = plm_slope<VX1>(dvp[VX2], dvm[VX2]);
  if (NC == 3) dvl[VX3] = plm_slope<VX1>(dvp[VX3], dvm[VX3]);
  if (SKIP != BX1) dvl[BX1] = plm_slope<VX1>(dvp[BX1], dvm[BX1]);
  if (SKIP != BX2) dvl[BX2] = plm_slope<VX1>(dvp[BX2], dvm[BX2]);
  if (NC == 3 && SKIP != BX3) dvl[BX3] = plm_slope<VX1>(dvp[BX3], dvm[BX3]);
  dvl[PRS] = plm_slope<PRS>(dvp[PRS], dvm[PRS]);
  PG_FOR_NV_SKIP(nv, SKIP){
    vp[nv] = v[nv] + dvl[nv]*0.5;
    vm[nv] = v[nv] - dvl[nv]*0.5;
  }
}
#endif

// ---------------------------------------------------------------------------
//  CHAR_LIMITING YES, 2 components (DIMENSIONS = COMPONENTS = 2): slopes limited on the characteristic variables
//  (plm_states.c:448-706 on a uniform Cartesian grid: cp = cm = 2, dp = dm = 1/2, cpk = cmk = kstp; PrimEigenvectors
//  eigenv.c:190-470 for the ideal EOS with CT -- six waves, no div.B jump; PrimToChar eigenv.c:1310-1375).
//  The reference fills 6 x 6 matrices; only their non-zero entries are formed here, and the sums run in its order
//  (waves: fast-, fast+, entropy, div.B, slow-, slow+).  The row of the normal field is never needed: the interface
//  states take the staggered component.  (3 components are not offered: the reference's never-cleared eigenvector
//  scratch makes its own 3-D result depend on the sweep order)
//  alpha_s (alpha_f) is the square root of a difference that is pure round-off where the transverse field vanishes
//  exactly -- the same sensitivity as Roe's switches -- so the chain up to the alphas keeps IEEE operations in every
//  build (x_* below: no contraction, div.rn / sqrt.rn).
// ---------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double x_mul (double a, double b) { return __dmul_rn (a, b); }
__device__ __forceinline__ double x_add (double a, double b) { return __dadd_rn (a, b); }
__device__ __forceinline__ double x_div (double a, double b) { return __ddiv_rn (a, b); }
__device__ __forceinline__ double x_sqrt (double a) { return __dsqrt_rn (a); }
#else      // host builds (kernel interpreter): keep the compiler from contracting a product with a following sum
__device__ __forceinline__ double x_mul (double a, double b) { double t = a*b; asm volatile ("" : "+x"(t)); return t; }
__device__ __forceinline__ double x_add (double a, double b) { double t = a + b; asm volatile ("" : "+x"(t)); return t; }
__device__ __forceinline__ double x_div (double a, double b) { double t = a/b; asm volatile ("" : "+x"(t)); return t; }
__device__ __forceinline__ double x_sqrt (double a) { return sqrt (a); }
#endif

__device__ __forceinline__ double gm_limiter (double dwp, double dwm, double ck)
{
  if (dwp*dwm > 0.0){
    const double qc = 0.5*(dwm + dwp), scrh = abs_min (dwp*ck, dwm*ck);
    return abs_min (qc, scrh);
  }
  return 0.0;
}

template <int DIR>
__device__ __forceinline__ void plm_zone_char2 (const Phys &ph, int lim, const double *v, const double *dvm, const double *dvp,
                                                double *vp, double *vm)
{
  typedef Dirs<DIR> D;
  const int VXn = D::vn, VXt = D::vt, BXn = D::bn, BXt = D::bt;
  // ---- PrimEigenvectors up to alpha_f, alpha_s: IEEE operations in the reference's order ----
  const double a2   = x_div (x_mul (ph.gamma, v[PRS]), v[RHO]);           // SoundSpeed2
  const double tau  = x_div (1.0, v[RHO]);
  const double sqrt_rho = x_sqrt (v[RHO]);
  const double bn2  = x_mul (v[BXn], v[BXn]);
  const double bt2  = x_add (0.0, x_mul (v[BXt], v[BXt]));
  const double b2   = x_add (bn2, bt2);
  const double ca2  = x_mul (bn2, tau);
  const double A2   = x_mul (b2, tau);
  const double At2  = x_mul (bt2, tau);
  const double d1   = x_add (a2, -A2);
  const double disc = x_sqrt (x_add (x_mul (d1, d1), x_mul (x_mul (4.0, a2), At2)));
  const double cf2  = x_mul (0.5, x_add (x_add (a2, A2), disc));
  const double cs2  = x_div (x_mul (a2, ca2), cf2);
  const double cf = x_sqrt (cf2), cs = x_sqrt (cs2), a = x_sqrt (a2);
  double alpha_f, alpha_s;
  if (cf == cs){
    alpha_f = 1.0; alpha_s = 0.0;
  }else{
    const double id = x_div (1.0, disc);
    alpha_f = x_mul (x_add (a2, -cs2), id);
    alpha_s = x_mul (x_add (cf2, -a2), id);
    alpha_f = maxv (0.0, alpha_f);
    alpha_s = maxv (0.0, alpha_s);
    alpha_f = x_sqrt (alpha_f);
    alpha_s = x_sqrt (alpha_s);
  }
  double beta_y;
  if (x_sqrt (bt2) > 1.e-9) beta_y = (v[BXt] >= 0.0 ? 1.0 : -1.0);
  else                      beta_y = 1.0;
  const double S = (v[BXn] >= 0.0 ? 1.0 : -1.0);

  // ---- non-zero entries of the right and left eigenvectors ----
  const double h2 = pg_div (0.5, a2), h3 = h2*tau;                          // scrh2, scrh3
  // fast waves
  const double f0 = alpha_s*cs*S, f1 = alpha_s*sqrt_rho*a;
  const double Rf_rho = v[RHO]*alpha_f, Rf_vn = -cf*alpha_f, Rf_vt = f0*beta_y, Rf_bt = f1*beta_y, Rf_p = alpha_f*a2*v[RHO];
  const double Lf_vn = Rf_vn*h2, Lf_vt = Rf_vt*h2, Lf_bt = Rf_bt*h3, Lf_p = alpha_f*h3;
  // entropy wave
  const double Le_p = -pg_div (1.0, a2);
  // slow waves
  const double s0 = alpha_f*cf*S, s1 = alpha_f*sqrt_rho*a;
  const double Rs_rho = v[RHO]*alpha_s, Rs_vn = -cs*alpha_s, Rs_vt = -s0*beta_y, Rs_bt = -s1*beta_y, Rs_p = alpha_s*a2*v[RHO];
  const double Ls_vn = Rs_vn*h2, Ls_vt = Rs_vt*h2, Ls_bt = Rs_bt*h3, Ls_p = alpha_s*h3;

  // ---- PrimToChar of the backward and forward differences, limiter per wave ----
  double wl[6];                                      // limited characteristic slopes (index 3, div.B: zero)
  {
    double wm_[6], wp_[6];
    PG_UNROLL for (int s = 0; s < 2; s++){
      const double *d = (s == 0 ? dvm : dvp);
      double *w = (s == 0 ? wm_ : wp_);
      double wv = Lf_vn*d[VXn] + Lf_vt*d[VXt];
      double wB = Lf_p*d[PRS] + Lf_bt*d[BXt];
      w[0] = wv + wB; w[1] = -wv + wB;
      w[2] = 1.0*d[RHO] + Le_p*d[PRS];
      w[3] = 0.0;
      wv = Ls_vn*d[VXn] + Ls_vt*d[VXt];
      wB = Ls_p*d[PRS] + Ls_bt*d[BXt];
      w[4] = wv + wB; w[5] = -wv + wB;
    }
    PG_UNROLL for (int k = 0; k < 6; k++){
      if (k == 3){ wl[k] = 0.0; continue; }
      if (lim == 0) wl[k] = gm_limiter (wp_[k], wm_[k], k == 2 ? 2.0 : 1.0);        // kstp: 1 for the fast and slow families
      else          wl[k] = single_limiter (lim, wp_[k], wm_[k]);
    }
  }
  // ---- back to primitive slopes (sum over the waves in the reference's order), monotone in the primitive variables too ----
  double dc[NV];
  dc[RHO] = 0.0 + wl[0]*Rf_rho; dc[RHO] += wl[1]*Rf_rho; dc[RHO] += wl[2]*1.0; dc[RHO] += wl[4]*Rs_rho; dc[RHO] += wl[5]*Rs_rho;
  dc[VXn] = 0.0 + wl[0]*Rf_vn;  dc[VXn] += wl[1]*(-Rf_vn);                      dc[VXn] += wl[4]*Rs_vn;  dc[VXn] += wl[5]*(-Rs_vn);
  dc[VXt] = 0.0 + wl[0]*Rf_vt;  dc[VXt] += wl[1]*(-Rf_vt);                      dc[VXt] += wl[4]*Rs_vt;  dc[VXt] += wl[5]*(-Rs_vt);
  dc[BXt] = 0.0 + wl[0]*Rf_bt;  dc[BXt] += wl[1]*Rf_bt;                         dc[BXt] += wl[4]*Rs_bt;  dc[BXt] += wl[5]*Rs_bt;
  dc[PRS] = 0.0 + wl[0]*Rf_p;   dc[PRS] += wl[1]*Rf_p;                          dc[PRS] += wl[4]*Rs_p;   dc[PRS] += wl[5]*Rs_p;
  vp[BXn] = vm[BXn] = v[BXn];                       // replaced by the staggered component (plm_states.c:663-667)
  PG_UNROLL for (int nv = 0; nv < NV; nv++){
    if (nv == VX3 || nv == BX3 || nv == BXn) continue;
    double dvl = 0.0;
    if (dvp[nv]*dvm[nv] > 0.0){
      const double d2v = abs_min (2.0*dvp[nv], 2.0*dvm[nv]);
      dvl = minmod (d2v, dc[nv]);
    }
    vp[nv] = v[nv] + dvl*0.5;
    vm[nv] = v[nv] - dvl*0.5;
  }
}

// UNIFORM_CARTESIAN_GRID NO (plm_coeffs.h:23-29): the weights of PLM_CoefficientsGet (plm_coeffs.c:30-104) for this zone and
// direction -- dvp = dv[i] wp, dvm = dv[i-1] wm (plm_states.c:157-164), the limiters "on irregular grids" with cp, cm
// (plm_coeffs.h:130-152: OSPRE, van Leer, MC; flat, minmod, van Albada, UMIST are the same on every grid),
// vp = v + dv_lim dp, vm = v - dv_lim dm (:240-241), in the reference's operation order.
__device__ __forceinline__ double general_limiter (int lim, double dvp, double dvm, double cp, double cm)
{
  if (!(dvp*dvm > 0.0)) return 0.0;
  if (lim == 4){
    const double den = 2.0*dvp*dvp + 2.0*dvm*dvm + (cp + cm - 2.0)*dvp*dvm;
    return pg_div (dvp*dvm*((1.0 + cp)*dvm + (1.0 + cm)*dvp), den);
  }
  if (lim == 6) return pg_div (dvp*dvm*(cp*dvm + cm*dvp), dvp*dvp + dvm*dvm + (cp + cm - 2.0)*dvp*dvm);
  if (lim == 7){
    const double qc = 0.5*(dvm + dvp), scrh = abs_min (dvp*cp, dvm*cm);
    return abs_min (qc, scrh);
  }
  return single_limiter (lim, dvp, dvm);
}
template <int NC, int SKIP = -1>
__device__ __forceinline__ void plm_zone_w (int lim, const double *const *pc, int n, const double *v, const double *dvm,
                                            const double *dvp, double *vp, double *vm)
{
  const double cp = __ldg (pc[0] + n), cm = __ldg (pc[1] + n), wp = __ldg (pc[2] + n), wm = __ldg (pc[3] + n);
  const double dp = __ldg (pc[4] + n), dm = __ldg (pc[5] + n);
  PG_FOR_NV_SKIP(nv, SKIP){
    // LIMITER DEFAULT: MC on the density, minmod on the pressure, van Leer on velocity and field (plm_states.c:192-227)
    const int l = (lim != 0 ? lim : nv == RHO ? 7 : nv == PRS ? 2 : 6);
    const double dvl = general_limiter (l, dvp[nv]*wp, dvm[nv]*wm, cp, cm);
    vp[nv] = v[nv] + dvl*dp;
    vm[nv] = v[nv] - dvl*dm;
  }
}

// ---------------------------------------------------------------------------
//  TIME_STEPPING CHARACTERISTIC_TRACING, 2 components: the predictor of the corner-transport-upwind step by characteristic
//  tracing (States/char_tracing.c:278-560: LINEAR reconstruction, CARTESIAN, CHAR_LIMITING NO, CHTR_REF_STATE 3, no source
//  terms).  In: the limited interface states vp, vm of the zone (normal field = the staggered one of its two faces), the zone
//  value vc, dt/dx.  The eigenvectors are those of plm_zone_char2 (PrimEigenvectors, eigenv.c:190-470: six waves, no div.B
//  jump, lambda[KDIVB] = 0), the chain up to the alphas in IEEE operations for the same reason; only non-zero entries of
//  the right eigenvectors are summed (the reference's scratch holds zeros elsewhere, and a stale row of the normal field
//  that the staggered component overwrites).  Waves in the reference's order: fast-, fast+, entropy, div.B, slow-, slow+.
// ---------------------------------------------------------------------------
template <int DIR>
__device__ __forceinline__ void char_tracing2 (const Phys &ph, const double *v, double dtdx, double *vp, double *vm)
{
  typedef Dirs<DIR> D;
  const int VXn = D::vn, VXt = D::vt, BXn = D::bn, BXt = D::bt;
  const double a2   = x_div (x_mul (ph.gamma, v[PRS]), v[RHO]);           // SoundSpeed2
  const double u    = v[VXn];
  const double tau  = x_div (1.0, v[RHO]);
  const double sqrt_rho = x_sqrt (v[RHO]);
  const double bn2  = x_mul (v[BXn], v[BXn]);
  const double bt2  = x_add (0.0, x_mul (v[BXt], v[BXt]));
  const double b2   = x_add (bn2, bt2);
  const double ca2  = x_mul (bn2, tau);
  const double A2   = x_mul (b2, tau);
  const double At2  = x_mul (bt2, tau);
  const double d1   = x_add (a2, -A2);
  const double disc = x_sqrt (x_add (x_mul (d1, d1), x_mul (x_mul (4.0, a2), At2)));
  const double cf2  = x_mul (0.5, x_add (x_add (a2, A2), disc));
  const double cs2  = x_div (x_mul (a2, ca2), cf2);
  const double cf = x_sqrt (cf2), cs = x_sqrt (cs2), a = x_sqrt (a2);
  double alpha_f, alpha_s;
  if (cf == cs){
    alpha_f = 1.0; alpha_s = 0.0;
  }else{
    const double id = x_div (1.0, disc);
    alpha_f = x_mul (x_add (a2, -cs2), id);
    alpha_s = x_mul (x_add (cf2, -a2), id);
    alpha_f = maxv (0.0, alpha_f);
    alpha_s = maxv (0.0, alpha_s);
    alpha_f = x_sqrt (alpha_f);
    alpha_s = x_sqrt (alpha_s);
  }
  double beta_y;
  if (x_sqrt (bt2) > 1.e-9) beta_y = (v[BXt] >= 0.0 ? 1.0 : -1.0);
  else                      beta_y = 1.0;
  const double S = (v[BXn] >= 0.0 ? 1.0 : -1.0);
  const double h2 = pg_div (0.5, a2), h3 = h2*tau;
  const double f0 = alpha_s*cs*S, f1 = alpha_s*sqrt_rho*a;
  const double Rf_rho = v[RHO]*alpha_f, Rf_vn = -cf*alpha_f, Rf_vt = f0*beta_y, Rf_bt = f1*beta_y, Rf_p = alpha_f*a2*v[RHO];
  const double Lf_vn = Rf_vn*h2, Lf_vt = Rf_vt*h2, Lf_bt = Rf_bt*h3, Lf_p = alpha_f*h3;
  const double Le_p = -pg_div (1.0, a2);
  const double s0 = alpha_f*cf*S, s1 = alpha_f*sqrt_rho*a;
  const double Rs_rho = v[RHO]*alpha_s, Rs_vn = -cs*alpha_s, Rs_vt = -s0*beta_y, Rs_bt = -s1*beta_y, Rs_p = alpha_s*a2*v[RHO];
  const double Ls_vn = Rs_vn*h2, Ls_vt = Rs_vt*h2, Ls_bt = Rs_bt*h3, Ls_p = alpha_s*h3;

  // characteristic Courant numbers (char_tracing.c:361-362)
  double nu[6];
  nu[0] = dtdx*(u - cf); nu[1] = dtdx*(u + cf); nu[2] = dtdx*u; nu[3] = dtdx*0.0; nu[4] = dtdx*(u - cs); nu[5] = dtdx*(u + cs);
  const double nu_max = maxv (nu[1], 0.0), nu_min = minv (nu[0], 0.0);
  // dv = vp - vm projected on the left eigenvectors (:388-389, PrimToChar eigenv.c:1310-1375)
  const double d_rho = vp[RHO] - vm[RHO], d_vn = vp[VXn] - vm[VXn], d_vt = vp[VXt] - vm[VXt], d_bt = vp[BXt] - vm[BXt], d_p = vp[PRS] - vm[PRS];
  double dw[6];
  {
    double wv = Lf_vn*d_vn + Lf_vt*d_vt;
    double wB = Lf_p*d_p + Lf_bt*d_bt;
    dw[0] = wv + wB; dw[1] = -wv + wB;
    dw[2] = 1.0*d_rho + Le_p*d_p;
    dw[3] = 0.0;
    wv = Ls_vn*d_vn + Ls_vt*d_vt;
    wB = Ls_p*d_p + Ls_bt*d_bt;
    dw[4] = wv + wB; dw[5] = -wv + wB;
  }
  // reference state of the fastest waves (:409-417)
  double qp[5], qm[5];                  // rho, vn, vt, bt, p
  const double dq[5] = {d_rho, d_vn, d_vt, d_bt, d_p};
  const double qc[5] = {v[RHO], v[VXn], v[VXt], v[BXt], v[PRS]};
  PG_UNROLL for (int q = 0; q < 5; q++){
    qp[q] = qc[q] + 0.5*dq[q]*(1.0 - nu_max);
    qm[q] = qc[q] - 0.5*dq[q]*(1.0 + nu_min);
  }
  // right eigenvectors, rows rho, vn, vt, bt, p; columns = waves (zero entries left out)
  const double R[5][6] = {{Rf_rho,  Rf_rho, 1.0, 0.0, Rs_rho,  Rs_rho},
                          {Rf_vn,  -Rf_vn,  0.0, 0.0, Rs_vn,  -Rs_vn},
                          {Rf_vt,  -Rf_vt,  0.0, 0.0, Rs_vt,  -Rs_vt},
                          {Rf_bt,   Rf_bt,  0.0, 0.0, Rs_bt,   Rs_bt},
                          {Rf_p,    Rf_p,   0.0, 0.0, Rs_p,    Rs_p}};
  PG_UNROLL for (int k = 0; k < 6; k++){            // :441-471
    if (k == 3) continue;                           // dw = 0: nothing is added
    if (nu[k] >= 0.0){
      const double w = dw[k]*(0.5*(nu_max - nu[k]));
      PG_UNROLL for (int q = 0; q < 5; q++) if (!(k == 2 && q > 0)) qp[q] += w*R[q][k];
    }else{
      const double w = dw[k]*(0.5*(nu_min - nu[k]));
      PG_UNROLL for (int q = 0; q < 5; q++) if (!(k == 2 && q > 0)) qm[q] += w*R[q][k];
    }
  }
  vp[RHO] = qp[0]; vp[VXn] = qp[1]; vp[VXt] = qp[2]; vp[BXt] = qp[3]; vp[PRS] = qp[4];
  vm[RHO] = qm[0]; vm[VXn] = qm[1]; vm[VXt] = qm[2]; vm[BXt] = qm[3]; vm[PRS] = qm[4];
}

// LIMITER DEFAULT (the fast path) or one limiter for all variables (uniform branch)
template <int NC, int SKIP = -1>
__device__ __forceinline__ void plm_zone (int lim, const double *v, const double *dvm, const double *dvp,
                                          double *vp, double *vm)
{
  if (lim == 0) plm_zone_default<NC, SKIP>(v, dvm, dvp, vp, vm);
  else          plm_zone_single<NC, SKIP>(lim, v, dvm, dvp, vp, vm);
}

// PPM 4th-order interface value at i+1/2, bounded (ppm_states.c:146-157):
// W = v0 + MINMOD(P - v0, v1 - v0), P = -1/12 vm1 + 7/12 v0 + 7/12 v1 - 1/12 v2
// qc: non-uniform grid -- the weights wp[n][-1 .. 2] of zone n (ppm_states.c:147, found by PPM_FindWeights, ppm_coeffs.c:300-480),
// four consecutive doubles per zone; NULL: the analytic ones of a uniform direction (ppm_coeffs.c:500-505)
template <int NC, int SKIP = -1>
__device__ __forceinline__ void ppm_interface (const double *vm1, const double *v0,
                                               const double *v1, const double *v2, double *W,
                                               const double *qc = nullptr, int n = 0)
{
  double wm1 = -1.0/12.0, w0 = 7.0/12.0, w1 = 7.0/12.0, w2 = -1.0/12.0;
  if (qc){ qc += 4*n; wm1 = __ldg (qc); w0 = __ldg (qc + 1); w1 = __ldg (qc + 2); w2 = __ldg (qc + 3); }
  PG_FOR_NV_SKIP(nv, SKIP){
    double p = wm1*vm1[nv] + w0*v0[nv] + w1*v1[nv] + w2*v2[nv];
    double dv = v1[nv] - v0[nv];
    double dvp = p - v0[nv];
    W[nv] = v0[nv] + minmod(dvp, dv);
  }
}

// parabolic limiter on one zone (ppm_states.c:185-207, cm = cp = 2 on a
// uniform Cartesian grid: (hm+1)/(hp-1) with hp = hm = 3)
template <int NC, int SKIP = -1>
__device__ __forceinline__ void ppm_zone (const double *v, const double *Wm, const double *Wp,
                                          double *vp, double *vm)
{
  const double hp = 3.0, hm = 3.0;
  const double cm = (hm + 1.0)/(hp - 1.0), cp = (hp + 1.0)/(hm - 1.0);
  PG_FOR_NV_SKIP(nv, SKIP){
    double dvp = Wp[nv] - v[nv];
    double dvm = Wm[nv] - v[nv];
    if (dvp*dvm >= 0.0) dvp = dvm = 0.0;
    else{
      if      (fabs(dvp) >= cm*fabs(dvm)) dvp = -cm*dvm;
      else if (fabs(dvm) >= cp*fabs(dvp)) dvm = -cp*dvp;
    }
    vp[nv] = v[nv] + dvp;
    vm[nv] = v[nv] + dvm;
  }
}

// SHOCK_FLATTENING MULTID with PARABOLIC reconstruction: a zone flagged FLAG_MINMOD falls back to the minmod-limited linear
// states with the weights of PLM_CoefficientsGet (ppm_states.c:167-181) -- wp, wm, dp, dm of the zone; on a uniform grid they are 1
// and 1/2 only up to the round-off of the grid coordinates, so the host hands the reference's arrays over (pluto_gpu_set_plm_coeffs)
template <int NC, int SKIP = -1>
__device__ __forceinline__ void ppm_flat_zone (const double *const *pc, int n, const double *vl, const double *v, const double *vr,
                                               double *vp, double *vm)
{
  const double wp = __ldg (pc[2] + n), wm = __ldg (pc[3] + n), dp = __ldg (pc[4] + n), dm = __ldg (pc[5] + n);
  PG_FOR_NV_SKIP(nv, SKIP){
    const double dvp = (vr[nv] - v[nv])*wp;
    const double dvm = (v[nv] - vl[nv])*wm;
    const double dv  = minmod (dvp, dvm);
    vp[nv] = v[nv] + dv*dp;
    vm[nv] = v[nv] - dv*dm;
  }
}

// ---------------------------------------------------------------------------
//  mappers
// ---------------------------------------------------------------------------
template <int NC>
__device__ __forceinline__ void prim_to_cons (const Phys &ph, const double *v, double *u)
{
  double kinb2;
  u[RHO] = v[RHO];
  u[MX1] = v[RHO]*v[VX1];
  u[MX2] = v[RHO]*v[VX2];
  if (NC == 3) u[MX3] = v[RHO]*v[VX3];
  u[BX1] = v[BX1]; u[BX2] = v[BX2];
  if (NC == 3) u[BX3] = v[BX3];
  if (NC == 3){
    kinb2 = v[VX1]*v[VX1] + v[VX2]*v[VX2] + v[VX3]*v[VX3];
    kinb2 = v[RHO]*kinb2 + v[BX1]*v[BX1] + v[BX2]*v[BX2] + v[BX3]*v[BX3];
  }else{
    kinb2 = v[VX1]*v[VX1] + v[VX2]*v[VX2];
    kinb2 = v[RHO]*kinb2 + v[BX1]*v[BX1] + v[BX2]*v[BX2];
  }
  kinb2 *= 0.5;
#ifdef PG_FAST_MATH
  u[ENG] = kinb2 + v[PRS]*ph.igmm1;
#else
  u[ENG] = kinb2 + pg_div (v[PRS], ph.gmm1);
#endif
}

// returns 1 when a floor was applied; u is repaired in place as the reference does
template <int NC>
__device__ __forceinline__ int cons_to_prim (const Phys &ph, double *u, double *v)
{
  double m2, b2, tau, kinb2;
  int fail = 0;
  if (NC == 3){
    m2 = u[MX1]*u[MX1] + u[MX2]*u[MX2] + u[MX3]*u[MX3];
    b2 = u[BX1]*u[BX1] + u[BX2]*u[BX2] + u[BX3]*u[BX3];
  }else{
    m2 = u[MX1]*u[MX1] + u[MX2]*u[MX2];
    b2 = u[BX1]*u[BX1] + u[BX2]*u[BX2];
  }
  if (u[RHO] < 0.0){ u[RHO] = ph.small_dn; fail = 1; }
  v[RHO] = u[RHO];
  tau = pg_rcp (u[RHO]);
  v[VX1] = u[MX1]*tau; v[VX2] = u[MX2]*tau;
  if (NC == 3) v[VX3] = u[MX3]*tau;
  v[BX1] = u[BX1]; v[BX2] = u[BX2];
  if (NC == 3) v[BX3] = u[BX3];
  kinb2 = 0.5*(m2*tau + b2);
  if (u[ENG] < 0.0){ u[ENG] = pg_div (ph.small_pr, ph.gmm1) + kinb2; fail = 1; }
  v[PRS] = ph.gmm1*(u[ENG] - kinb2);
  if (v[PRS] < 0.0){
    v[PRS] = ph.small_pr;
    u[ENG] = pg_div (v[PRS], ph.gmm1) + kinb2;
    fail = 1;
  }
  return fail;
}

// ---------------------------------------------------------------------------
//  flux and speeds
// ---------------------------------------------------------------------------
template <int DIR, int NC>
__device__ __forceinline__ void mhd_flux (const double *v, const double *u, double *fx, double &prs)
{
  typedef Dirs<DIR> D;
  double Bmag2, ptot, vB;
  if (NC == 3){
    Bmag2 = v[BX1]*v[BX1] + v[BX2]*v[BX2] + v[BX3]*v[BX3];
    vB    = v[VX1]*v[BX1] + v[VX2]*v[BX2] + v[VX3]*v[BX3];
  }else{
    Bmag2 = v[BX1]*v[BX1] + v[BX2]*v[BX2];
    vB    = v[VX1]*v[BX1] + v[VX2]*v[BX2];
  }
  ptot = v[PRS] + 0.5*Bmag2;
  fx[RHO] = u[D::vn];
  fx[MX1] = v[D::vn]*u[MX1] - v[D::bn]*v[BX1];
  fx[MX2] = v[D::vn]*u[MX2] - v[D::bn]*v[BX2];
  if (NC == 3) fx[MX3] = v[D::vn]*u[MX3] - v[D::bn]*v[BX3];
  fx[D::bn] = 0.0;
  fx[D::bt] = v[D::vn]*v[D::bt] - v[D::bn]*v[D::vt];
  if (NC == 3) fx[D::bb] = v[D::vn]*v[D::bb] - v[D::bn]*v[D::vb];
  fx[ENG] = (u[ENG] + ptot)*v[D::vn] - v[D::bn]*vB;
  prs = ptot;
}

template <int DIR, int NC>
__device__ __forceinline__ void max_signal_speed (const Phys &ph, const double *v,
                                                  double &cmin, double &cmax)
{
  typedef Dirs<DIR> D;
  double gpr, b1, b2, b3, Btmag2, Bmag2, cf;
  gpr = ph.gamma*v[PRS];
  b1 = v[D::bn]; b2 = v[D::bt];
  if (NC == 3){ b3 = v[D::bb]; Btmag2 = b2*b2 + b3*b3; }
  else        { Btmag2 = b2*b2; }
  Bmag2 = b1*b1 + Btmag2;
  cf = gpr - Bmag2;
  cf = gpr + Bmag2 + pg_sqrt (cf*cf + 4.0*gpr*Btmag2);
  cf = pg_sqrt (pg_div (0.5*cf, v[RHO]));
  cmin = v[D::vn] - cf;
  cmax = v[D::vn] + cf;
}

template <int DIR, int NC>
__device__ __forceinline__ void hll_speed (const Phys &ph, const double *vL, const double *vR,
                                           double a2L, double a2R, double &SL, double &SR,
                                           double &mach)
{
  typedef Dirs<DIR> D;
  double slmin, slmax, srmin, srmax, scrh;
  max_signal_speed<DIR, NC>(ph, vL, slmin, slmax);
  max_signal_speed<DIR, NC>(ph, vR, srmin, srmax);
  SL = minv(slmin, srmin);
  SR = maxv(slmax, srmax);
  scrh  = fabs(vL[D::vn]) + fabs(vR[D::vn]);
  scrh = pg_div (scrh, pg_sqrt (a2L) + pg_sqrt (a2R));
  mach = scrh;
}

// UCT_HLL: the Riemann fan speeds of the face, as CT_StoreUpwindEMF keeps them
// (ct_emf.c:144-147): max(0, -SL), max(0, SR).  Stored as soon as they are known so that
// they do not stay in registers through the solve; null pointers = not wanted.
__device__ __forceinline__ void store_fan_speeds (double *pSL, double *pSR, double SL, double SR)
{
  if (pSL){ *pSL = maxv (0.0, -SL); *pSR = maxv (0.0, SR); }
}

// ---------------------------------------------------------------------------
//  Riemann solvers.  in: vL, vR (interface states), uL, uR; out: flux[NV]
//  (slot bn is not meaningful), press, cmax, mach (candidate for g_maxMach).
// ---------------------------------------------------------------------------
// STRICT: hll.c's own tests SL > 0 / SR < 0 (:106,113); !STRICT: the HLL flux that HLLD
// substitutes in flagged zones, after ITS tests SL >= 0 / SR <= 0 (hlld.c:131-160)
template <int DIR, int NC, bool STRICT = true>
__device__ __forceinline__ void riemann_hll (const Phys &ph, const double *vL, const double *vR,
                                             const double *uL, const double *uR,
                                             double *flux, double &press, double &cmax, double &mach,
                                             double *pSL = nullptr, double *pSR = nullptr)
{
  double fL[NV], fR[NV], pL, pR, a2L, a2R, SL, SR, scrh;
  a2L = pg_div (ph.gamma*vL[PRS], vL[RHO]);
  a2R = pg_div (ph.gamma*vR[PRS], vR[RHO]);
  mhd_flux<DIR, NC>(vL, uL, fL, pL);
  mhd_flux<DIR, NC>(vR, uR, fR, pR);
  hll_speed<DIR, NC>(ph, vL, vR, a2L, a2R, SL, SR, mach);
  store_fan_speeds (pSL, pSR, SL, SR);
  scrh = maxv(fabs(SL), fabs(SR));
  cmax = scrh;
  if (STRICT ? SL > 0.0 : SL >= 0.0){
    PG_FOR_NV(nv) flux[nv] = fL[nv];
    press = pL;
  }else if (STRICT ? SR < 0.0 : SR <= 0.0){
    PG_FOR_NV(nv) flux[nv] = fR[nv];
    press = pR;
  }else{
    scrh = pg_rcp (SR - SL);
    PG_FOR_NV(nv){
      flux[nv]  = SL*SR*(uR[nv] - uL[nv]) + SR*fL[nv] - SL*fR[nv];
      flux[nv] *= scrh;
    }
    press = (SR*pL - SL*pR)*scrh;
  }
}

// Lax-Friedrichs (Rusanov) flux, tvdlf.c:51-135: one speed from the arithmetic mean of the two states; the fan speeds that
// UCT_HLL keeps are -cRL and +cRL (:123-124), the Mach number is that of the mean state (:104-105).
template <int DIR, int NC>
__device__ __forceinline__ void riemann_tvdlf (const Phys &ph, const double *vL, const double *vR,
                                               const double *uL, const double *uR,
                                               double *flux, double &press, double &cmax, double &mach,
                                               double *pSL = nullptr, double *pSR = nullptr)
{
  typedef Dirs<DIR> D;
  double fL[NV], fR[NV], vRL[NV], pL, pR, cminRL, cmaxRL;
  mhd_flux<DIR, NC>(vL, uL, fL, pL);
  mhd_flux<DIR, NC>(vR, uR, fR, pR);
  PG_FOR_NV(nv) vRL[nv] = 0.5*(vL[nv] + vR[nv]);
  mach = pg_div (fabs(vRL[D::vn]), pg_sqrt (pg_div (ph.gamma*vRL[PRS], vRL[RHO])));
  max_signal_speed<DIR, NC>(ph, vRL, cminRL, cmaxRL);
  const double cRL = maxv(fabs(cminRL), fabs(cmaxRL));
  store_fan_speeds (pSL, pSR, -cRL, cRL);
  cmax = cRL;
  PG_FOR_NV(nv) flux[nv] = 0.5*(fL[nv] + fR[nv] - cRL*(uR[nv] - uL[nv]));
  press = 0.5*(pL + pR);
}

// HLLC (Li 2005), hllc.c:42-236: the HLL average supplies the field and the contact speed of the two star states.  FLAGGED: the
// interface touches a zone that SHOCK_FLATTENING MULTID marked FLAG_HLL -> the plain HLL flux inside the fan (:140-150).
template <int DIR, int NC, bool FLAGGED = false>
__device__ __forceinline__ void riemann_hllc (const Phys &ph, const double *vL, const double *vR,
                                              const double *uL, const double *uR,
                                              double *flux, double &press, double &cmax, double &mach,
                                              double *pSL = nullptr, double *pSR = nullptr)
{
  typedef Dirs<DIR> D;
  constexpr int mxn = D::vn, mxt = D::vt, mxb = D::vb;            // momentum slots = velocity slots
  double fL[NV], fR[NV], pL, pR, a2L, a2R, SL, SR, scrh;
  a2L = pg_div (ph.gamma*vL[PRS], vL[RHO]);
  a2R = pg_div (ph.gamma*vR[PRS], vR[RHO]);
  mhd_flux<DIR, NC>(vL, uL, fL, pL);
  mhd_flux<DIR, NC>(vR, uR, fR, pR);
  hll_speed<DIR, NC>(ph, vL, vR, a2L, a2R, SL, SR, mach);
  store_fan_speeds (pSL, pSR, SL, SR);
  cmax = maxv(fabs(SL), fabs(SR));
  if (SL >= 0.0){
    PG_FOR_NV(nv) flux[nv] = fL[nv];
    press = pL;
    return;
  }
  if (SR <= 0.0){
    PG_FOR_NV(nv) flux[nv] = fR[nv];
    press = pR;
    return;
  }
  scrh = pg_rcp (SR - SL);
  if (FLAGGED){
    PG_FOR_NV(nv){
      flux[nv]  = SL*SR*(uR[nv] - uL[nv]) + SR*fL[nv] - SL*fR[nv];
      flux[nv] *= scrh;
    }
    press = (SR*pL - SL*pR)*scrh;
    return;
  }
  double Uhll[NV];
  PG_FOR_NV(nv){
    Uhll[nv]  = SR*uR[nv] - SL*uL[nv] + fL[nv] - fR[nv];
    Uhll[nv] *= scrh;
  }
  Uhll[mxn] += (pL - pR)*scrh;
  double Fn  = SL*SR*(uR[mxn] - uL[mxn]) + SR*fL[mxn] - SL*fR[mxn];      // Fhll[MXn], Fhll[RHO]: the only two in use
  Fn *= scrh;
  Fn += (SR*pL - SL*pR)*scrh;
  double Fr  = SL*SR*(uR[RHO] - uL[RHO]) + SR*fL[RHO] - SL*fR[RHO];
  Fr *= scrh;

  double pl, pr, vBl, vBr, vBs;
  if (NC == 3){
    pl  = vL[BX1]*vL[BX1] + vL[BX2]*vL[BX2] + vL[BX3]*vL[BX3];
    pr  = vR[BX1]*vR[BX1] + vR[BX2]*vR[BX2] + vR[BX3]*vR[BX3];
    vBl = vL[VX1]*vL[BX1] + vL[VX2]*vL[BX2] + vL[VX3]*vL[BX3];
    vBr = vR[VX1]*vR[BX1] + vR[VX2]*vR[BX2] + vR[VX3]*vR[BX3];
    vBs = Uhll[BX1]*Uhll[MX1] + Uhll[BX2]*Uhll[MX2] + Uhll[BX3]*Uhll[MX3];
  }else{
    pl  = vL[BX1]*vL[BX1] + vL[BX2]*vL[BX2];
    pr  = vR[BX1]*vR[BX1] + vR[BX2]*vR[BX2];
    vBl = vL[VX1]*vL[BX1] + vL[VX2]*vL[BX2];
    vBr = vR[VX1]*vR[BX1] + vR[VX2]*vR[BX2];
    vBs = Uhll[BX1]*Uhll[MX1] + Uhll[BX2]*Uhll[MX2];
  }
  pl = vL[PRS] + 0.5*pl;
  pr = vR[PRS] + 0.5*pr;
  const double vxl = vL[D::vn], vxr = vR[D::vn];
  const double Bxs = Uhll[D::bn], Bys = Uhll[D::bt];
  const double vxs = pg_div (Uhll[mxn], Uhll[RHO]);
  const double ps  = Fn + Bxs*Bxs - Fr*vxs;
  vBs = pg_div (vBs, Uhll[RHO]);

  // the state on the side of the contact the interface lies on (the reference builds both, :185-209); selects value by
  // value, so that the state arrays stay in registers
  const bool left = (vxs >= 0.0);
  #define PG_SIDE(aL, aR) (left ? (aL) : (aR))
  const double S = PG_SIDE (SL, SR), vx = PG_SIDE (vxl, vxr), pt = PG_SIDE (pl, pr), vB = PG_SIDE (vBl, vBr);
  const double bn = PG_SIDE (vL[D::bn], vR[D::bn]);
#ifdef PG_FAST_MATH
  const double idn = pg_rcp (S - vxs);
  #define PG_HLLC_DIV(x) ((x)*idn)
#else
  const double den = S - vxs;
  #define PG_HLLC_DIV(x) pg_div ((x), den)
#endif
  double us[NV];
  us[RHO] = PG_HLLC_DIV (PG_SIDE (uL[RHO], uR[RHO])*(S - vx));
  us[ENG] = PG_HLLC_DIV (PG_SIDE (uL[ENG], uR[ENG])*(S - vx) + ps*vxs - pt*vx - Bxs*vBs + bn*vB);
  us[mxn] = us[RHO]*vxs;
  us[mxt] = PG_HLLC_DIV (PG_SIDE (uL[mxt], uR[mxt])*(S - vx) - (Bxs*Bys - bn*PG_SIDE (vL[D::bt], vR[D::bt])));
  if (NC == 3){
    const double Bzs = Uhll[D::bb];
    us[mxb] = PG_HLLC_DIV (PG_SIDE (uL[mxb], uR[mxb])*(S - vx) - (Bxs*Bzs - bn*PG_SIDE (vL[D::bb], vR[D::bb])));
    us[D::bb] = Bzs;
  }
  #undef PG_HLLC_DIV
  us[D::bn] = Bxs;
  us[D::bt] = Bys;
  PG_FOR_NV(nv) flux[nv] = PG_SIDE (fL[nv], fR[nv]) + S*(us[nv] - PG_SIDE (uL[nv], uR[nv]));
  press = PG_SIDE (pL, pR);
  #undef PG_SIDE
}

#ifdef PG_FAST_MATH
// ---------------------------------------------------------------------------
//  FAST HLLD.  Same five-wave solver, same wave-speed estimates, same region
//  switches as hlld.c:98-427, evaluated in the form that needs the fewest FP64
//  instructions and the fewest live registers:
//   * every intermediate state of HLLD satisfies the jump conditions exactly, so
//     the flux in a region is the ideal-MHD flux function applied to THAT state
//     with its normal velocity and total pressure (Miyoshi & Kusano 2005,
//     eqs 38-48, 59-63):  F = F(rho, un, vt, Bt, E; pT).  No F_L/F_R, no
//     conservative input vectors, no differences U* - U are formed;
//   * transverse star velocities use the same denominator as the star fields
//     (M&K eq 44/46), which removes two divisions;
//   * the normal field is the staggered one on both sides (plm_states.c:271-275),
//     hence Bx = vL[BXn] = vR[BXn];
//   * energies are evaluated on the side the flux is taken from only.
//  The HLLC fallback of hlld.c:245-258 is not a consistent state: interfaces that
//  need it (rare: Alfven speed within 1e-4 of the fast speed) take the cold path
//  riemann_hlld_fallback, the previous formulation.
//  Returns press = total pressure of the selected state and flux[MXn] without it,
//  the split the callers expect (rhs.c:193-201 adds the two differences).
// ---------------------------------------------------------------------------
template <int DIR, int NC>
__device__ __noinline__ void riemann_hlld_fallback (const Phys &ph, const double *vL, const double *vR,
                                                    const double *uL, const double *uR,
                                                    double *flux, double &press, double &cmax, double &mach);

template <int DIR, int NC, bool LEFT>
__device__ __forceinline__ void hlld_side_flux (const Phys &ph, const double *v, double b2, double pt,
                                                double S, double S1, double SM, double pts, double Bx, double Bx2,
                                                double rs, double sq_s, double iSM,
                                                double vs, double ws, double Bts, double Bbs,
                                                double vss, double wss, double Btss, double Bbss, double vBss,
                                                double *flux, double &press)
{
  typedef Dirs<DIR> D;
  const int VXn = D::vn, VXt = D::vt, VXb = D::vb, BXt = D::bt, BXb = D::bb;
  double R, U, Vt, Vb = 0.0, BT, BB = 0.0, EE, PT, VB;
  double v2 = v[VXn]*v[VXn] + v[VXt]*v[VXt];
  if (NC == 3) v2 = fma (v[VXb], v[VXb], v2);
  const double E = fma (v[PRS], ph.igmm1, 0.5*fma (v[RHO], v2, b2));
  double vB = fma (v[VXt], v[BXt], v[VXn]*Bx);
  if (NC == 3) vB = fma (v[VXb], v[BXb], vB);
  if (LEFT ? S >= 0.0 : S <= 0.0){             // supersonic: the physical flux of this side
    R = v[RHO]; U = v[VXn]; Vt = v[VXt]; BT = v[BXt]; EE = E; PT = pt; VB = vB;
    if (NC == 3){ Vb = v[VXb]; BB = v[BXb]; }
  }else{
    double vBs = fma (vs, Bts, SM*Bx);
    if (NC == 3) vBs = fma (ws, Bbs, vBs);
    const double du = S - v[VXn];
    const double Es = (fma (Bx, vB - vBs, fma (pts, SM, fma (du, E, -pt*v[VXn]))))*iSM;
    // U** energy (hlld.c:396-411): E*L - sqrt(rho*L) (v*.B* - v**.B**) sBx, E*R + ...
    const double Ess = LEFT ? fma (-sq_s, vBs - vBss, Es) : fma (sq_s, vBs - vBss, Es);
    const bool star = LEFT ? S1 >= 0.0 : S1 <= 0.0;
    R = rs; U = SM; PT = pts;
    Vt = star ? vs : vss;  BT = star ? Bts : Btss;
    EE = star ? Es : Ess;  VB = star ? vBs : vBss;
    if (NC == 3){ Vb = star ? ws : wss; BB = star ? Bbs : Bbss; }
  }
  const double Fr = R*U;
  flux[RHO]   = Fr;
  flux[D::vn] = fma (Fr, U, -Bx2);
  flux[D::vt] = fma (Fr, Vt, -Bx*BT);
  flux[D::bn] = 0.0;
  flux[D::bt] = fma (BT, U, -Bx*Vt);
  if (NC == 3){
    flux[D::vb] = fma (Fr, Vb, -Bx*BB);
    flux[D::bb] = fma (BB, U, -Bx*Vb);
  }
  flux[ENG]   = fma (EE + PT, U, -Bx*VB);
  press = PT;
}

template <int DIR, int NC>
__device__ __forceinline__ void riemann_hlld (const Phys &ph, const double *vL, const double *vR,
                                              double *flux, double &press, double &cmax, double &mach,
                                              double *pSL = nullptr, double *pSR = nullptr)
{
  typedef Dirs<DIR> D;
  const int VXn = D::vn, VXt = D::vt, VXb = D::vb, BXn = D::bn, BXt = D::bt, BXb = D::bb;
  const double Bx = vL[BXn], Bx2 = Bx*Bx;
  double SL, SR, b2L, b2R;
  {
    // fast magnetosonic speeds (eigenv.c:87-101) and Davis estimate (hll_speed.c:76-107)
    const double irL = pg_rcp (vL[RHO]), irR = pg_rcp (vR[RHO]);
    const double gpL = ph.gamma*vL[PRS], gpR = ph.gamma*vR[PRS];
    double bt2L = vL[BXt]*vL[BXt], bt2R = vR[BXt]*vR[BXt];
    if (NC == 3){ bt2L = fma (vL[BXb], vL[BXb], bt2L); bt2R = fma (vR[BXb], vR[BXb], bt2R); }
    b2L = Bx2 + bt2L; b2R = Bx2 + bt2R;
    double dL = gpL - b2L, dR = gpR - b2R;
    dL = gpL + b2L + pg_sqrt (fma (dL, dL, 4.0*gpL*bt2L));
    dR = gpR + b2R + pg_sqrt (fma (dR, dR, 4.0*gpR*bt2R));
    const double cfL = pg_sqrt_pos (0.5*dL*irL), cfR = pg_sqrt_pos (0.5*dR*irR);
    SL = minv (vL[VXn] - cfL, vR[VXn] - cfR);
    SR = maxv (vL[VXn] + cfL, vR[VXn] + cfR);
    // g_maxMach is a diagnostic the reference prints with 6 digits: single precision
    const float aL = pg_sqrtf ((float)(gpL*irL)), aR = pg_sqrtf ((float)(gpR*irR));
    mach = (double)__fdividef ((float)(fabs (vL[VXn]) + fabs (vR[VXn])), aL + aR);
  }
  store_fan_speeds (pSL, pSR, SL, SR);
  cmax = maxv (fabs (SL), fabs (SR));
  const double ptL = fma (0.5, b2L, vL[PRS]), ptR = fma (0.5, b2R, vR[PRS]);

  const double duL = SL - vL[VXn], duR = SR - vR[VXn];
  const double rduL = vL[RHO]*duL, rduR = vR[RHO]*duR;            // rho (S - vn)
  const double idn = pg_rcp (rduR - rduL);
  const double SM  = (fma (rduR, vR[VXn], -rduL*vL[VXn]) - ptR + ptL)*idn;
  const double pts = (fma (rduR, ptL, -rduL*ptR) + rduL*rduR*(vR[VXn] - vL[VXn]))*idn;

  const double dSL = SL - SM, dSR = SR - SM;
  const double iSLM = pg_rcp (dSL), iSRM = pg_rcp (dSR);
  const double rsL = rduL*iSLM, rsR = rduR*iSRM;                   // star densities
  double sqrL, sqrR, isqL, isqR;
  pg_sqrt_rsqrt (rsL, sqrL, isqL);
  pg_sqrt_rsqrt (rsR, sqrR, isqR);
  const double aBx = fabs (Bx);
  const double S1L = fma (-aBx, isqL, SM), S1R = fma (aBx, isqR, SM);

  // hlld.c:238-243
  if ( (S1L - SL) < -1.e-4*dSL || (S1R - SR) > -1.e-4*dSR ){
    // cold path: private copies, so that the caller's arrays never have their
    // address taken (they stay in registers on the hot path)
    double a[NV], b[NV], ua[NV], ub[NV], fl[NV], pr, cm, ma;
    PG_UNROLL for (int nv = 0; nv < NV; nv++){ a[nv] = 0.0; b[nv] = 0.0; fl[nv] = 0.0; }
    PG_FOR_NV(nv){ a[nv] = vL[nv]; b[nv] = vR[nv]; }
    prim_to_cons<NC>(ph, a, ua);
    prim_to_cons<NC>(ph, b, ub);
    riemann_hlld_fallback<DIR, NC>(ph, a, b, ua, ub, fl, pr, cm, ma);
    PG_FOR_NV(nv) flux[nv] = fl[nv];
    press = pr;
    return;
  }

  // star states (M&K eqs 44-47 with the common denominator rho (S-u)(S-SM) - Bx^2)
  const double idL = pg_rcp (fma (rduL, dSL, -Bx2)), idR = pg_rcp (fma (rduR, dSR, -Bx2));
  const double qL = fma (rduL, duL, -Bx2)*idL, qR = fma (rduR, duR, -Bx2)*idR;
  const double tL = (SM - vL[VXn])*idL*Bx, tR = (SM - vR[VXn])*idR*Bx;
  const double BtsL = vL[BXt]*qL, BtsR = vR[BXt]*qR;
  const double vsL = fma (-tL, vL[BXt], vL[VXt]), vsR = fma (-tR, vR[BXt], vR[VXt]);
  double BbsL = 0.0, BbsR = 0.0, wsL = 0.0, wsR = 0.0;
  if (NC == 3){
    BbsL = vL[BXb]*qL; BbsR = vR[BXb]*qR;
    wsL = fma (-tL, vL[BXb], vL[VXb]); wsR = fma (-tR, vR[BXb], vR[VXb]);
  }

  // double-star state (hlld.c:349-394); sBx = sign(Bx) folded into the weights
  const double isum = pg_rcp (sqrL + sqrR);
  const double sgL = (Bx > 0.0 ? sqrL : -sqrL);                    // sBx sqrt(rho*L)
  const double sI = (Bx > 0.0 ? isum : -isum);                      // sBx/(sqrt(rho*L) + sqrt(rho*R))
  const double vss  = fma (BtsR - BtsL, sI, fma (sqrL, vsL, sqrR*vsR)*isum);
  const double Btss = fma (sgL*sqrR, (vsR - vsL)*isum, fma (sqrL, BtsR, sqrR*BtsL)*isum);
  double wss = 0.0, Bbss = 0.0;
  if (NC == 3){
    wss  = fma (BbsR - BbsL, sI, fma (sqrL, wsL, sqrR*wsR)*isum);
    Bbss = fma (sgL*sqrR, (wsR - wsL)*isum, fma (sqrL, BbsR, sqrR*BbsL)*isum);
  }
  double vBss = fma (vss, Btss, SM*Bx);
  if (NC == 3) vBss = fma (wss, Bbss, vBss);

  if (SM >= 0.0)
    hlld_side_flux<DIR, NC, true> (ph, vL, b2L, ptL, SL, S1L, SM, pts, Bx, Bx2, rsL, sgL, iSLM,
                                   vsL, wsL, BtsL, BbsL, vss, wss, Btss, Bbss, vBss, flux, press);
  else
    hlld_side_flux<DIR, NC, false>(ph, vR, b2R, ptR, SR, S1R, SM, pts, Bx, Bx2, rsR, (Bx > 0.0 ? sqrR : -sqrR), iSRM,
                                   vsR, wsR, BtsR, BbsR, vss, wss, Btss, Bbss, vBss, flux, press);
}

template <int DIR, int NC>
__device__ __noinline__ void riemann_hlld_fallback (const Phys &ph, const double *vL, const double *vR,
                                              const double *uL, const double *uR,
                                              double *flux, double &press, double &cmax, double &mach)
{
  typedef Dirs<DIR> D;
  const int VXn = D::vn, VXt = D::vt, VXb = D::vb, BXn = D::bn, BXt = D::bt, BXb = D::bb;
  const int MXn = VXn, MXt = VXt, MXb = VXb;
  double fL[NV], fR[NV], ptL, ptR, SL, SR;

  mhd_flux<DIR, NC>(vL, uL, fL, ptL);
  mhd_flux<DIR, NC>(vR, uR, fR, ptR);
  {
    // fast magnetosonic speeds (eigenv.c:87-101) and Davis estimate (hll_speed.c:76-107)
    const double irL = pg_rcp (vL[RHO]), irR = pg_rcp (vR[RHO]);
    const double gpL = ph.gamma*vL[PRS], gpR = ph.gamma*vR[PRS];
    double bt2L = vL[BXt]*vL[BXt], bt2R = vR[BXt]*vR[BXt];
    if (NC == 3){ bt2L = fma (vL[BXb], vL[BXb], bt2L); bt2R = fma (vR[BXb], vR[BXb], bt2R); }
    const double b2L = fma (vL[BXn], vL[BXn], bt2L), b2R = fma (vR[BXn], vR[BXn], bt2R);
    double dL = gpL - b2L, dR = gpR - b2R;
    dL = gpL + b2L + pg_sqrt (fma (dL, dL, 4.0*gpL*bt2L));
    dR = gpR + b2R + pg_sqrt (fma (dR, dR, 4.0*gpR*bt2R));
    const double cfL = pg_sqrt_pos (0.5*dL*irL), cfR = pg_sqrt_pos (0.5*dR*irR);
    SL = minv (vL[VXn] - cfL, vR[VXn] - cfR);
    SR = maxv (vL[VXn] + cfL, vR[VXn] + cfR);
    const float aL = sqrtf ((float)(gpL*irL)), aR = sqrtf ((float)(gpR*irR));
    mach = (double)__fdividef ((float)(fabs (vL[VXn]) + fabs (vR[VXn])), aL + aR);
  }
  cmax = maxv (fabs (SL), fabs (SR));

  if (SL >= 0.0){
    PG_FOR_NV(nv) flux[nv] = fL[nv];
    press = ptL;
    return;
  }else if (SR <= 0.0){
    PG_FOR_NV(nv) flux[nv] = fR[nv];
    press = ptR;
    return;
  }

  const double iSRL = pg_rcp (SR - SL);
  const double Bx = (SR*vR[BXn] - SL*vL[BXn])*iSRL;
  const double sBx = (Bx > 0.0 ? 1.0 : -1.0);
  const double duL = SL - vL[VXn], duR = SR - vR[VXn];
  const double rduL = uL[RHO]*duL, rduR = uR[RHO]*duR;          // rho (S - vn)
  const double idn = pg_rcp (rduR - rduL);
  const double SM  = (duR*uR[MXn] - duL*uL[MXn] - ptR + ptL)*idn;
  const double pts = (rduR*ptL - rduL*ptR + vL[RHO]*vR[RHO]*duR*duL*(vR[VXn] - vL[VXn]))*idn;

  const double iSLM = pg_rcp (SL - SM), iSRM = pg_rcp (SR - SM);
  const double rsL = rduL*iSLM, rsR = rduR*iSRM;                 // star densities
  double sqrL, sqrR, isqL, isqR;
  pg_sqrt_rsqrt (rsL, sqrL, isqL);
  pg_sqrt_rsqrt (rsR, sqrR, isqR);
  double S1L = SM - fabs (Bx)*isqL;
  double S1R = SM + fabs (Bx)*isqR;

  bool revert_to_hllc = false;
  if ( (S1L - SL) <  1.e-4*(SM - SL) ) revert_to_hllc = true;
  if ( (S1R - SR) > -1.e-4*(SR - SM) ) revert_to_hllc = true;

  double BtsL, BtsR, BbsL = 0.0, BbsR = 0.0;                     // star transverse fields
  const double Bx2 = Bx*Bx;
  if (revert_to_hllc){
    BtsL = BtsR = (SR*uR[BXt] - SL*uL[BXt] + fL[BXt] - fR[BXt])*iSRL;
    if (NC == 3) BbsL = BbsR = (SR*uR[BXb] - SL*uL[BXb] + fL[BXb] - fR[BXb])*iSRL;
    S1L = S1R = SM;
  }else{
    const double qL = (rduL*duL - Bx2)*pg_rcp (rduL*(SL - SM) - Bx2);
    const double qR = (rduR*duR - Bx2)*pg_rcp (rduR*(SR - SM) - Bx2);
    BtsL = uL[BXt]*qL; BtsR = uR[BXt]*qR;
    if (NC == 3){ BbsL = uL[BXb]*qL; BbsR = uR[BXb]*qR; }
  }
  const double kL = Bx*pg_rcp (rduL), kR = Bx*pg_rcp (rduR);
  const double vsL = vL[VXt] - kL*(BtsL - uL[BXt]);
  const double vsR = vR[VXt] - kR*(BtsR - uR[BXt]);
  double wsL = 0.0, wsR = 0.0;
  if (NC == 3){
    wsL = vL[VXb] - kL*(BbsL - uL[BXb]);
    wsR = vR[VXb] - kR*(BbsR - uR[BXb]);
  }
  double vBL, vBsL, vBR, vBsR;                                   // v.B and its star value
  if (NC == 3){
    vBL  = vL[VXn]*Bx + vL[VXt]*uL[BXt] + vL[VXb]*uL[BXb];
    vBsL = SM*Bx + vsL*BtsL + wsL*BbsL;
    vBR  = vR[VXn]*Bx + vR[VXt]*uR[BXt] + vR[VXb]*uR[BXb];
    vBsR = SM*Bx + vsR*BtsR + wsR*BbsR;
  }else{
    vBL  = vL[VXn]*Bx + vL[VXt]*uL[BXt];
    vBsL = SM*Bx + vsL*BtsL;
    vBR  = vR[VXn]*Bx + vR[VXt]*uR[BXt];
    vBsR = SM*Bx + vsR*BtsR;
  }
  const double EsL = (duL*uL[ENG] - ptL*vL[VXn] + pts*SM + Bx*(vBL - vBsL))*iSLM;
  const double EsR = (duR*uR[ENG] - ptR*vR[VXn] + pts*SM + Bx*(vBR - vBsR))*iSRM;

  if (S1L >= 0.0){
    flux[RHO] = fL[RHO] + SL*(rsL - uL[RHO]);
    flux[MXn] = fL[MXn] + SL*(rsL*SM - uL[MXn]);
    flux[MXt] = fL[MXt] + SL*(rsL*vsL - uL[MXt]);
    if (NC == 3) flux[MXb] = fL[MXb] + SL*(rsL*wsL - uL[MXb]);
    flux[BXn] = 0.0;
    flux[BXt] = fL[BXt] + SL*(BtsL - uL[BXt]);
    if (NC == 3) flux[BXb] = fL[BXb] + SL*(BbsL - uL[BXb]);
    flux[ENG] = fL[ENG] + SL*(EsL - uL[ENG]);
    press = ptL;
  }else if (S1R <= 0.0){
    flux[RHO] = fR[RHO] + SR*(rsR - uR[RHO]);
    flux[MXn] = fR[MXn] + SR*(rsR*SM - uR[MXn]);
    flux[MXt] = fR[MXt] + SR*(rsR*vsR - uR[MXt]);
    if (NC == 3) flux[MXb] = fR[MXb] + SR*(rsR*wsR - uR[MXb]);
    flux[BXn] = 0.0;
    flux[BXt] = fR[BXt] + SR*(BtsR - uR[BXt]);
    if (NC == 3) flux[BXb] = fR[BXb] + SR*(BbsR - uR[BXb]);
    flux[ENG] = fR[ENG] + SR*(EsR - uR[ENG]);
    press = ptR;
  }else{
    const double isum = pg_rcp (sqrL + sqrR);
    const double vss = (sqrL*vsL + sqrR*vsR + (BtsR - BtsL)*sBx)*isum;
    const double Btss = (sqrL*BtsR + sqrR*BtsL + sqrL*sqrR*(vsR - vsL)*sBx)*isum;
    double wss = 0.0, Bbss = 0.0;
    if (NC == 3){
      wss  = (sqrL*wsL + sqrR*wsR + (BbsR - BbsL)*sBx)*isum;
      Bbss = (sqrL*BbsR + sqrR*BbsL + sqrL*sqrR*(wsR - wsL)*sBx)*isum;
    }
    double vBss;
    if (NC == 3) vBss = SM*Bx + vss*Btss + wss*Bbss; else vBss = SM*Bx + vss*Btss;
    if (SM >= 0.0){
      const double Ess = EsL - sqrL*(vBsL - vBss)*sBx;
      flux[RHO] = fL[RHO] + SL*(rsL - uL[RHO]);
      flux[MXn] = fL[MXn] + SL*(rsL*SM - uL[MXn]);
      flux[MXt] = fL[MXt] + S1L*(rsL*vss - rsL*vsL) + SL*(rsL*vsL - uL[MXt]);
      if (NC == 3) flux[MXb] = fL[MXb] + S1L*(rsL*wss - rsL*wsL) + SL*(rsL*wsL - uL[MXb]);
      flux[BXn] = 0.0;
      flux[BXt] = fL[BXt] + S1L*(Btss - BtsL) + SL*(BtsL - uL[BXt]);
      if (NC == 3) flux[BXb] = fL[BXb] + S1L*(Bbss - BbsL) + SL*(BbsL - uL[BXb]);
      flux[ENG] = fL[ENG] + S1L*(Ess - EsL) + SL*(EsL - uL[ENG]);
      press = ptL;
    }else{
      const double Ess = EsR + sqrR*(vBsR - vBss)*sBx;
      flux[RHO] = fR[RHO] + SR*(rsR - uR[RHO]);
      flux[MXn] = fR[MXn] + SR*(rsR*SM - uR[MXn]);
      flux[MXt] = fR[MXt] + S1R*(rsR*vss - rsR*vsR) + SR*(rsR*vsR - uR[MXt]);
      if (NC == 3) flux[MXb] = fR[MXb] + S1R*(rsR*wss - rsR*wsR) + SR*(rsR*wsR - uR[MXb]);
      flux[BXn] = 0.0;
      flux[BXt] = fR[BXt] + S1R*(Btss - BtsR) + SR*(BtsR - uR[BXt]);
      if (NC == 3) flux[BXb] = fR[BXb] + S1R*(Bbss - BbsR) + SR*(BbsR - uR[BXb]);
      flux[ENG] = fR[ENG] + S1R*(Ess - EsR) + SR*(EsR - uR[ENG]);
      press = ptR;
    }
  }
}
#else
template <int DIR, int NC>
__device__ __forceinline__ void riemann_hlld (const Phys &ph, const double *vL, const double *vR,
                                              const double *uL, const double *uR,
                                              double *flux, double &press, double &cmax, double &mach,
                                              double *pSL = nullptr, double *pSR = nullptr)
{
  typedef Dirs<DIR> D;
  const int VXn = D::vn, VXt = D::vt, VXb = D::vb, BXn = D::bn, BXt = D::bt, BXb = D::bb;
  const int MXn = VXn, MXt = VXt, MXb = VXb;
  double fL[NV], fR[NV], ptL, ptR, a2L, a2R, SL, SR, scrh;
  double usL[NV], usR[NV];
  double vsL, wsL = 0.0, scrhL, S1L, sqrL, duL;
  double vsR, wsR = 0.0, scrhR, S1R, sqrR, duR;
  double Bx, Bx1, SM, sBx, pts;

  a2L = pg_div (ph.gamma*vL[PRS], vL[RHO]);
  a2R = pg_div (ph.gamma*vR[PRS], vR[RHO]);
  mhd_flux<DIR, NC>(vL, uL, fL, ptL);
  mhd_flux<DIR, NC>(vR, uR, fR, ptR);
  hll_speed<DIR, NC>(ph, vL, vR, a2L, a2R, SL, SR, mach);
  store_fan_speeds (pSL, pSR, SL, SR);

  scrh = maxv(fabs(SL), fabs(SR));
  cmax = scrh;

  if (SL >= 0.0){
    PG_FOR_NV(nv) flux[nv] = fL[nv];
    press = ptL;
    return;
  }else if (SR <= 0.0){
    PG_FOR_NV(nv) flux[nv] = fR[nv];
    press = ptR;
    return;
  }

  scrh = pg_rcp (SR - SL);
  Bx1  = Bx = (SR*vR[BXn] - SL*vL[BXn])*scrh;
  sBx  = (Bx > 0.0 ? 1.0 : -1.0);

  duL = SL - vL[VXn];
  duR = SR - vR[VXn];

  scrh = pg_rcp (duR*uR[RHO] - duL*uL[RHO]);
  SM   = (duR*uR[MXn] - duL*uL[MXn] - ptR + ptL)*scrh;

  pts  = duR*uR[RHO]*ptL - duL*uL[RHO]*ptR +
         vL[RHO]*vR[RHO]*duR*duL*(vR[VXn] - vL[VXn]);
  pts *= scrh;

  usL[RHO] = pg_div (uL[RHO]*duL, SL - SM);
  usR[RHO] = pg_div (uR[RHO]*duR, SR - SM);

  sqrL = pg_sqrt (usL[RHO]);
  sqrR = pg_sqrt (usR[RHO]);

  S1L = SM - pg_div (fabs(Bx), sqrL);
  S1R = SM + pg_div (fabs(Bx), sqrR);

  bool revert_to_hllc = false;
  if ( (S1L - SL) <  1.e-4*(SM - SL) ) revert_to_hllc = true;
  if ( (S1R - SR) > -1.e-4*(SR - SM) ) revert_to_hllc = true;

  if (revert_to_hllc){
    scrh = pg_rcp (SR - SL);
    double hn = (SR*uR[BXn] - SL*uL[BXn] + fL[BXn] - fR[BXn])*scrh;
    double ht = (SR*uR[BXt] - SL*uL[BXt] + fL[BXt] - fR[BXt])*scrh;
    usL[BXn] = usR[BXn] = hn;
    usL[BXt] = usR[BXt] = ht;
    if (NC == 3){
      double hb = (SR*uR[BXb] - SL*uL[BXb] + fL[BXb] - fR[BXb])*scrh;
      usL[BXb] = usR[BXb] = hb;
    }
    S1L = S1R = SM;
  }else{
    scrhL = pg_div (uL[RHO]*duL*duL - Bx*Bx, uL[RHO]*duL*(SL - SM) - Bx*Bx);
    scrhR = pg_div (uR[RHO]*duR*duR - Bx*Bx, uR[RHO]*duR*(SR - SM) - Bx*Bx);
    usL[BXn] = Bx1;
    usL[BXt] = uL[BXt]*scrhL;
    usR[BXn] = Bx1;
    usR[BXt] = uR[BXt]*scrhR;
    if (NC == 3){
      usL[BXb] = uL[BXb]*scrhL;
      usR[BXb] = uR[BXb]*scrhR;
    }
  }

  scrhL = pg_div (Bx, uL[RHO]*duL);
  scrhR = pg_div (Bx, uR[RHO]*duR);

  vsL = vL[VXt] - scrhL*(usL[BXt] - uL[BXt]);
  vsR = vR[VXt] - scrhR*(usR[BXt] - uR[BXt]);
  if (NC == 3){
    wsL = vL[VXb] - scrhL*(usL[BXb] - uL[BXb]);
    wsR = vR[VXb] - scrhR*(usR[BXb] - uR[BXb]);
  }

  usL[MXn] = usL[RHO]*SM;
  usR[MXn] = usR[RHO]*SM;
  usL[MXt] = usL[RHO]*vsL;
  usR[MXt] = usR[RHO]*vsR;
  if (NC == 3){
    usL[MXb] = usL[RHO]*wsL;
    usR[MXb] = usR[RHO]*wsR;
  }

  if (NC == 3){
    scrhL  = vL[VXn]*Bx1 + vL[VXt]*uL[BXt] + vL[VXb]*uL[BXb];
    scrhL -=      SM*Bx1 +    vsL*usL[BXt] +    wsL*usL[BXb];
  }else{
    scrhL  = vL[VXn]*Bx1 + vL[VXt]*uL[BXt];
    scrhL -=      SM*Bx1 +    vsL*usL[BXt];
  }
  usL[ENG]  = duL*uL[ENG] - ptL*vL[VXn] + pts*SM + Bx*scrhL;
  usL[ENG] = pg_div (usL[ENG], SL - SM);

  if (NC == 3){
    scrhR  = vR[VXn]*Bx1 + vR[VXt]*uR[BXt] + vR[VXb]*uR[BXb];
    scrhR -=      SM*Bx1 +    vsR*usR[BXt] +    wsR*usR[BXb];
  }else{
    scrhR  = vR[VXn]*Bx1 + vR[VXt]*uR[BXt];
    scrhR -=      SM*Bx1 +    vsR*usR[BXt];
  }
  usR[ENG]  = duR*uR[ENG] - ptR*vR[VXn] + pts*SM + Bx*scrhR;
  usR[ENG] = pg_div (usR[ENG], SR - SM);

  if (S1L >= 0.0){
    PG_FOR_NV(nv) flux[nv] = fL[nv] + SL*(usL[nv] - uL[nv]);
    press = ptL;
  }else if (S1R <= 0.0){
    PG_FOR_NV(nv) flux[nv] = fR[nv] + SR*(usR[nv] - uR[nv]);
    press = ptR;
  }else{
    double ussl[NV], ussr[NV], vss, wss = 0.0;
    ussl[RHO] = usL[RHO];
    ussr[RHO] = usR[RHO];

    vss  = sqrL*vsL + sqrR*vsR + (usR[BXt] - usL[BXt])*sBx;
    vss = pg_div (vss, sqrL + sqrR);
    if (NC == 3){
      wss  = sqrL*wsL + sqrR*wsR + (usR[BXb] - usL[BXb])*sBx;
      wss = pg_div (wss, sqrL + sqrR);
    }

    ussl[MXn] = ussl[RHO]*SM;
    ussr[MXn] = ussr[RHO]*SM;
    ussl[MXt] = ussl[RHO]*vss;
    ussr[MXt] = ussr[RHO]*vss;
    if (NC == 3){
      ussl[MXb] = ussl[RHO]*wss;
      ussr[MXb] = ussr[RHO]*wss;
    }

    ussl[BXn] = ussr[BXn] = Bx1;
    ussl[BXt]  = sqrL*usR[BXt] + sqrR*usL[BXt] + sqrL*sqrR*(vsR - vsL)*sBx;
    ussl[BXt] = pg_div (ussl[BXt], sqrL + sqrR);
    ussr[BXt]  = ussl[BXt];
    if (NC == 3){
      ussl[BXb]  = sqrL*usR[BXb] + sqrR*usL[BXb] + sqrL*sqrR*(wsR - wsL)*sBx;
      ussl[BXb] = pg_div (ussl[BXb], sqrL + sqrR);
      ussr[BXb]  = ussl[BXb];
    }

    if (NC == 3){
      scrhL  = SM*Bx1 + vsL*usL [BXt] + wsL*usL [BXb];
      scrhL -= SM*Bx1 + vss*ussl[BXt] + wss*ussl[BXb];
      scrhR  = SM*Bx1 + vsR*usR [BXt] + wsR*usR [BXb];
      scrhR -= SM*Bx1 + vss*ussr[BXt] + wss*ussr[BXb];
    }else{
      scrhL  = SM*Bx1 + vsL*usL [BXt];
      scrhL -= SM*Bx1 + vss*ussl[BXt];
      scrhR  = SM*Bx1 + vsR*usR [BXt];
      scrhR -= SM*Bx1 + vss*ussr[BXt];
    }

    ussl[ENG] = usL[ENG] - sqrL*scrhL*sBx;
    ussr[ENG] = usR[ENG] + sqrR*scrhR*sBx;

    if (SM >= 0.0){
      PG_FOR_NV(nv) flux[nv] = fL[nv] + S1L*(ussl[nv] - usL[nv]) + SL*(usL[nv] - uL[nv]);
      press = ptL;
    }else{
      PG_FOR_NV(nv) flux[nv] = fR[nv] + S1R*(ussr[nv] - usR[nv]) + SR*(usR[nv] - uR[nv]);
      press = ptR;
    }
  }
}

#endif   // PG_FAST_MATH

// Roe: returns false when a2 < 0 (the reference aborts, roe.c:300-306).
// The eigenvector normalisation switches on `cf == cs`, `a <= cs`, `cf <= a` (roe.c:336-364).  Where the transverse
// field vanishes exactly (a rotor or blast at t = 0) cf2 equals a2 or b2 up to the last rounding, round-off decides the
// branch and the branches differ by sqrt(ulp) ~ 1e-8 in alpha_s / alpha_f: everything UPSTREAM of the switches (the
// interface states, the Roe averages, a2, cf2, cs2 and their roots) is therefore evaluated in the reference's IEEE
// arithmetic and operation order in every build (pg_* are IEEE in the Roe units, -fmad=false); what follows the
// switches is smooth in its inputs and uses the FAST division / square root / FMA (pq_*) in the FAST build.
template <int DIR, int NC>
__device__ __forceinline__ bool riemann_roe (const Phys &ph, const double *vL, const double *vR,
                                             const double *uL, const double *uR,
                                             double *flux, double &press, double &cmax, double &mach,
                                             double *pSL = nullptr, double *pSR = nullptr)
{
  typedef Dirs<DIR> D;
  const int VXn = D::vn, VXt = D::vt, VXb = D::vb, BXn = D::bn, BXt = D::bt, BXb = D::bb;
  const int MXn = VXn, MXt = VXt, MXb = VXb;
  enum { KFASTM, KFASTP, KENTRP, KDIVB, KSLOWM, KSLOWP, KALFVM, KALFVP, NW };
  const double sqrt_1_2 = 0.70710678118654752440;
  double fL[NV], fR[NV], pL, pR;
  double rho, u, v, w = 0.0, vel2, bx, by, bz = 0.0;
  double a2, a, ca2, cf2, cs2, cs, ca, cf, b2;
  double alpha_f, alpha_s, beta_y, beta_z = 0.0, beta_v, scrh, sBx;
  double dV[NV], dU[NV];
  double Rc[NV][NW], eta[NW], lambda[NW], alambda[NW];
  double sqrt_rho, delta = 1.e-6;
  double g1 = ph.gmm1, sl, sr, H, Hgas, HL, HR, Bx, By, Bz = 0.0, X;
  double vdm, BdB, beta_dv, beta_dB, bt2, Btmag, sqr_rho_L, sqr_rho_R;

  mhd_flux<DIR, NC>(vL, uL, fL, pL);
  mhd_flux<DIR, NC>(vR, uR, fR, pR);

  PG_UNROLL for (int nv = 0; nv < NV; nv++){
    PG_UNROLL for (int k = 0; k < NW; k++) Rc[nv][k] = 0.0;
  }
  PG_UNROLL for (int k = 0; k < NW; k++) eta[k] = lambda[k] = 0.0;

  PG_UNROLL for (int nv = 0; nv < NV; nv++){ dV[nv] = 0.0; dU[nv] = 0.0; }
  PG_FOR_NV(nv){
    dV[nv] = vR[nv] - vL[nv];
    dU[nv] = uR[nv] - uL[nv];
  }

  sqr_rho_L = pg_sqrt (vL[RHO]);
  sqr_rho_R = pg_sqrt (vR[RHO]);
  sl = pg_div (sqr_rho_L, sqr_rho_L + sqr_rho_R);
  sr = pg_div (sqr_rho_R, sqr_rho_L + sqr_rho_R);
  rho = sr*vL[RHO] + sl*vR[RHO];
  sqrt_rho = pg_sqrt (rho);

  u = sl*vL[VXn] + sr*vR[VXn];
  v = sl*vL[VXt] + sr*vR[VXt];
  if (NC == 3) w = sl*vL[VXb] + sr*vR[VXb];
  Bx = sr*vL[BXn] + sl*vR[BXn];
  By = sr*vL[BXt] + sl*vR[BXt];
  if (NC == 3) Bz = sr*vL[BXb] + sl*vR[BXb];

  sBx = (Bx >= 0.0 ? 1.0 : -1.0);
  bx = pg_div (Bx, sqrt_rho);
  by = pg_div (By, sqrt_rho);
  if (NC == 3) bz = pg_div (Bz, sqrt_rho);

  if (NC == 3) bt2 = 0.0 + by*by + bz*bz; else bt2 = 0.0 + by*by;
  b2    = bx*bx + bt2;
  Btmag = pg_sqrt (bt2*rho);

  if (NC == 3) X = dV[BXn]*dV[BXn] + dV[BXt]*dV[BXt] + dV[BXb]*dV[BXb];
  else         X = dV[BXn]*dV[BXn] + dV[BXt]*dV[BXt];
  X = pg_div (X, (sqr_rho_L + sqr_rho_R)*(sqr_rho_L + sqr_rho_R)*2.0);

  if (NC == 3){
    vdm = u*dU[MXn] + v*dU[MXt] + w*dU[MXb];
    BdB = Bx*dU[BXn] + By*dU[BXt] + Bz*dU[BXb];
    vel2 = u*u + v*v + w*w;
  }else{
    vdm = u*dU[MXn] + v*dU[MXt];
    BdB = Bx*dU[BXn] + By*dU[BXt];
    vel2 = u*u + v*v;
  }
  dV[PRS] = g1*((0.5*vel2 - X)*dV[RHO] - vdm + dU[ENG] - BdB);

  HL   = pg_div (uL[ENG] + pL, vL[RHO]);
  HR   = pg_div (uR[ENG] + pR, vR[RHO]);
  H    = sl*HL + sr*HR;
  Hgas = H - b2;

  a2 = (2.0 - ph.gamma)*X + g1*(Hgas - 0.5*vel2);
  bool ok = !(a2 < 0.0);

  scrh = a2 - b2;
  ca2  = bx*bx;
  scrh = scrh*scrh + 4.0*bt2*a2;
  scrh = pg_sqrt (scrh);

  cf2 = 0.5*(a2 + b2 + scrh);
  cs2 = pg_div (a2*ca2, cf2);

  cf = pg_sqrt (cf2);
  cs = pg_sqrt (cs2);
  ca = pg_sqrt (ca2);
  a  = pg_sqrt (a2);

  if (cf == cs){
    alpha_f = 1.0; alpha_s = 0.0;
  }else if (a <= cs){
    alpha_f = 0.0; alpha_s = 1.0;
  }else if (cf <= a){
    alpha_f = 1.0; alpha_s = 0.0;
  }else{
    scrh    = pq_rcp (cf2 - cs2);
    alpha_f = (a2  - cs2)*scrh;
    alpha_s = (cf2 -  a2)*scrh;
    alpha_f = maxv(0.0, alpha_f);
    alpha_s = maxv(0.0, alpha_s);
    alpha_f = pq_sqrt (alpha_f);
    alpha_s = pq_sqrt (alpha_s);
  }

  if (Btmag > 1.e-9){
    if (NC == 3){ beta_y = pq_div (By, Btmag); beta_z = pq_div (Bz, Btmag); }
    else          beta_y = (By >= 0.0 ? 1.0 : -1.0);
  }else{
    if (NC == 3) beta_z = beta_y = sqrt_1_2;
    else         beta_y = 1.0;
  }

  int k;
  // ---- fast wave u - cf ----
  k = KFASTM;
  lambda[k] = u - cf;
  scrh = alpha_s*cs*sBx;
  if (NC == 3){
    beta_dv = 0.0 + beta_y*dV[VXt] + beta_z*dV[VXb];
    beta_dB = 0.0 + beta_y*dV[BXt] + beta_z*dV[BXb];
    beta_v  = 0.0 + beta_y*v       + beta_z*w;
  }else{
    beta_dv = 0.0 + beta_y*dV[VXt];
    beta_dB = 0.0 + beta_y*dV[BXt];
    beta_v  = 0.0 + beta_y*v;
  }
  Rc[RHO][k] = alpha_f;
  Rc[MXn][k] = alpha_f*lambda[k];
  Rc[MXt][k] = alpha_f*v + scrh*beta_y;
  if (NC == 3) Rc[MXb][k] = alpha_f*w + scrh*beta_z;
  Rc[BXt][k] = pq_div (alpha_s*a*beta_y, sqrt_rho);
  if (NC == 3) Rc[BXb][k] = pq_div (alpha_s*a*beta_z, sqrt_rho);
  Rc[ENG][k] =   alpha_f*(Hgas - u*cf) + scrh*beta_v
               + pq_div (alpha_s*a*Btmag, sqrt_rho);
  eta[k] =   alpha_f*(X*dV[RHO] + dV[PRS]) + rho*scrh*beta_dv
           - rho*alpha_f*cf*dV[VXn]        + sqrt_rho*alpha_s*a*beta_dB;
  eta[k] *= pq_div (0.5, a2);

  // ---- fast wave u + cf ----
  k = KFASTP;
  lambda[k] = u + cf;
  Rc[RHO][k] = alpha_f;
  Rc[MXn][k] = alpha_f*lambda[k];
  Rc[MXt][k] = alpha_f*v - scrh*beta_y;
  if (NC == 3) Rc[MXb][k] = alpha_f*w - scrh*beta_z;
  Rc[BXt][k] = Rc[BXt][KFASTM];
  if (NC == 3) Rc[BXb][k] = Rc[BXb][KFASTM];
  Rc[ENG][k] =   alpha_f*(Hgas + u*cf) - scrh*beta_v
               + pq_div (alpha_s*a*Btmag, sqrt_rho);
  eta[k] =   alpha_f*(X*dV[RHO] + dV[PRS]) - rho*scrh*beta_dv
           + rho*alpha_f*cf*dV[VXn]        + sqrt_rho*alpha_s*a*beta_dB;
  eta[k] *= pq_div (0.5, a2);

  // ---- entropy wave ----
  k = KENTRP;
  lambda[k] = u;
  Rc[RHO][k] = 1.0;
  Rc[MXn][k] = u;
  Rc[MXt][k] = v;
  if (NC == 3) Rc[MXb][k] = w;
  Rc[ENG][k] = 0.5*vel2 + pq_div (ph.gamma - 2.0, g1)*X;
  eta[k] = pq_div ((a2 - X)*dV[RHO] - dV[PRS], a2);

  // ---- div.B wave: no jump with CT ----
  k = KDIVB;
  lambda[k] = u;
  Rc[BXn][k] = eta[k] = 0.0;

  // ---- slow wave u - cs ----
  scrh = alpha_f*cf*sBx;
  k = KSLOWM;
  lambda[k] = u - cs;
  Rc[RHO][k] = alpha_s;
  Rc[MXn][k] = alpha_s*lambda[k];
  Rc[MXt][k] = alpha_s*v - scrh*beta_y;
  if (NC == 3) Rc[MXb][k] = alpha_s*w - scrh*beta_z;
  Rc[BXt][k] = pq_div (- alpha_f*a*beta_y, sqrt_rho);
  if (NC == 3) Rc[BXb][k] = pq_div (- alpha_f*a*beta_z, sqrt_rho);
  Rc[ENG][k] =   alpha_s*(Hgas - u*cs) - scrh*beta_v
               - pq_div (alpha_f*a*Btmag, sqrt_rho);
  eta[k] =   alpha_s*(X*dV[RHO] + dV[PRS]) - rho*scrh*beta_dv
           - rho*alpha_s*cs*dV[VXn]        - sqrt_rho*alpha_f*a*beta_dB;
  eta[k] *= pq_div (0.5, a2);

  // ---- slow wave u + cs ----
  k = KSLOWP;
  lambda[k] = u + cs;
  Rc[RHO][k] = alpha_s;
  Rc[MXn][k] = alpha_s*lambda[k];
  Rc[MXt][k] = alpha_s*v + scrh*beta_y;
  if (NC == 3) Rc[MXb][k] = alpha_s*w + scrh*beta_z;
  Rc[BXt][k] = Rc[BXt][KSLOWM];
  if (NC == 3) Rc[BXb][k] = Rc[BXb][KSLOWM];
  Rc[ENG][k] =   alpha_s*(Hgas + u*cs) + scrh*beta_v
               - pq_div (alpha_f*a*Btmag, sqrt_rho);
  eta[k] =   alpha_s*(X*dV[RHO] + dV[PRS]) + rho*scrh*beta_dv
           + rho*alpha_s*cs*dV[VXn]        - sqrt_rho*alpha_f*a*beta_dB;
  eta[k] *= pq_div (0.5, a2);

  if (NC == 3){
    // ---- Alfven wave u - ca ----
    k = KALFVM;
    lambda[k] = u - ca;
    Rc[MXt][k] = - rho*beta_z;
    Rc[MXb][k] = + rho*beta_y;
    Rc[BXt][k] = - sBx*sqrt_rho*beta_z;
    Rc[BXb][k] =   sBx*sqrt_rho*beta_y;
    Rc[ENG][k] = - rho*(v*beta_z - w*beta_y);
    eta[k] = + beta_y*dV[VXb]               - beta_z*dV[VXt]
             + pq_div (sBx, sqrt_rho)*(beta_y*dV[BXb] - beta_z*dV[BXt]);
    eta[k] *= 0.5;

    // ---- Alfven wave u + ca ----
    k = KALFVP;
    lambda[k] = u + ca;
    Rc[MXt][k] = - Rc[MXt][KALFVM];
    Rc[MXb][k] = - Rc[MXb][KALFVM];
    Rc[BXt][k] =   Rc[BXt][KALFVM];
    Rc[BXb][k] =   Rc[BXb][KALFVM];
    Rc[ENG][k] = - Rc[ENG][KALFVM];
    eta[k] = - beta_y*dV[VXb]               + beta_z*dV[VXt]
             + pq_div (sBx, sqrt_rho)*(beta_y*dV[BXb] - beta_z*dV[BXt]);
    eta[k] *= 0.5;
  }

  cmax = fabs(u) + cf;
  mach = fabs(pq_div (u, a));
  store_fan_speeds (pSL, pSR, lambda[KFASTM], lambda[KFASTP]);      // roe.c:681-682
  const int nw = (NC == 3 ? 8 : 6);
  PG_UNROLL for (int kk = 0; kk < NW; kk++) alambda[kk] = fabs(lambda[kk]);

  // entropy fix (roe.c:623-640)
  if (alambda[KFASTM] < 0.5*delta) alambda[KFASTM] = pq_div (lambda[KFASTM]*lambda[KFASTM], delta) + 0.25*delta;
  if (alambda[KFASTP] < 0.5*delta) alambda[KFASTP] = pq_div (lambda[KFASTP]*lambda[KFASTP], delta) + 0.25*delta;
  if (alambda[KSLOWM] < 0.5*delta) alambda[KSLOWM] = pq_div (lambda[KSLOWM]*lambda[KSLOWM], delta) + 0.25*delta;
  if (alambda[KSLOWP] < 0.5*delta) alambda[KSLOWP] = pq_div (lambda[KSLOWP]*lambda[KSLOWP], delta) + 0.25*delta;

#ifdef PG_FAST
  PG_UNROLL for (int kk = 0; kk < NW; kk++) alambda[kk] *= eta[kk];
  PG_FOR_NV(nv){
    scrh = 0.0;
    PG_UNROLL for (int kk = 0; kk < NW; kk++) if (kk < nw) scrh = fma (alambda[kk], Rc[nv][kk], scrh);
    flux[nv] = 0.5*(fL[nv] + fR[nv] - scrh);
  }
#else
  PG_FOR_NV(nv){
    scrh = 0.0;
    PG_UNROLL for (int kk = 0; kk < NW; kk++) if (kk < nw) scrh += alambda[kk]*eta[kk]*Rc[nv][kk];
    flux[nv] = 0.5*(fL[nv] + fR[nv] - scrh);
  }
#endif
  press = 0.5*(pL + pR);
  return ok;
}

// SHOCK_FLATTENING MULTID: the flux of an interface with a shocked zone on either side
// (hlld.c:149-160, roe.c:165-189; hll.c is unchanged)
template <int SOLVER, int DIR, int NC>
__device__ __forceinline__ void riemann_flagged (const Phys &ph, const double *vL, const double *vR,
                                                 const double *uL, const double *uR,
                                                 double *flux, double &press, double &cmax, double &mach,
                                                 double *pSL, double *pSR)
{
  if      (SOLVER == SOLVER_HLLD)  riemann_hll<DIR, NC, false>(ph, vL, vR, uL, uR, flux, press, cmax, mach, pSL, pSR);
  else if (SOLVER == SOLVER_HLLC)  riemann_hllc<DIR, NC, true>(ph, vL, vR, uL, uR, flux, press, cmax, mach, pSL, pSR);    // hllc.c:140-150
  else if (SOLVER == SOLVER_TVDLF) riemann_tvdlf<DIR, NC>(ph, vL, vR, uL, uR, flux, press, cmax, mach, pSL, pSR);         // tvdlf.c has no flagged branch
  else                             riemann_hll<DIR, NC, true> (ph, vL, vR, uL, uR, flux, press, cmax, mach, pSL, pSR);
}

// solver dispatch on a compile-time constant
template <int SOLVER, int DIR, int NC>
__device__ __forceinline__ bool riemann (const Phys &ph, const double *vL, const double *vR,
                                         const double *uL, const double *uR,
                                         double *flux, double &press, double &cmax, double &mach,
                                         double *pSL = nullptr, double *pSR = nullptr)
{
#ifdef PG_FAST_MATH
  if (SOLVER == SOLVER_HLLD){ riemann_hlld<DIR, NC>(ph, vL, vR, flux, press, cmax, mach, pSL, pSR); return true; }
#else
  if (SOLVER == SOLVER_HLLD){ riemann_hlld<DIR, NC>(ph, vL, vR, uL, uR, flux, press, cmax, mach, pSL, pSR); return true; }
#endif
  else if (SOLVER == SOLVER_HLL){ riemann_hll<DIR, NC>(ph, vL, vR, uL, uR, flux, press, cmax, mach, pSL, pSR); return true; }
  else if (SOLVER == SOLVER_HLLC){ riemann_hllc<DIR, NC>(ph, vL, vR, uL, uR, flux, press, cmax, mach, pSL, pSR); return true; }
  else if (SOLVER == SOLVER_TVDLF){ riemann_tvdlf<DIR, NC>(ph, vL, vR, uL, uR, flux, press, cmax, mach, pSL, pSR); return true; }
  else return riemann_roe<DIR, NC>(ph, vL, vR, uL, uR, flux, press, cmax, mach, pSL, pSR);
}

} // namespace PG_NS
