// Host build of the device arithmetic (mhd_device.cuh) to check the FAST HLLD
// formulation against the reference-order EXACT formulation on random interface
// states.  Development aid; not part of the product.  Compiled twice:
//   g++ -O2 -std=c++17 -DPASS_EXACT -c tools/host_hlld_check.cpp -o /tmp/hc_exact.o
//   g++ -O2 -std=c++17 -DPASS_FAST  -c tools/host_hlld_check.cpp -o /tmp/hc_fast.o
//   g++ -o /tmp/hlld_check /tmp/hc_exact.o /tmp/hc_fast.o && /tmp/hlld_check
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __host__
static inline float __fdividef (float a, float b) { return a/b; }

#ifdef PASS_EXACT
#define PG_NS pg_exact
#include "../pluto_b200/csrc/mhd_device.cuh"
// dir-major dispatch
template <int DIR, int NC> static void one (double gamma, const double *vL, const double *vR, double *F, double *press, double *cmax, double *mach)
{
  pg_exact::Phys ph; ph.gamma = gamma; ph.gmm1 = gamma - 1.0; ph.small_dn = 1e-12; ph.small_pr = 1e-12; ph.igmm1 = 1.0/ph.gmm1;
  double uL[8], uR[8];
  pg_exact::prim_to_cons<NC>(ph, vL, uL);
  pg_exact::prim_to_cons<NC>(ph, vR, uR);
  pg_exact::riemann_hlld<DIR, NC>(ph, vL, vR, uL, uR, F, *press, *cmax, *mach);
}
void hlld_exact (int dir, int nc, double gamma, const double *vL, const double *vR, double *F, double *press, double *cmax, double *mach)
{
  if (nc == 3){ if (dir == 0) one<0,3>(gamma,vL,vR,F,press,cmax,mach); else if (dir == 1) one<1,3>(gamma,vL,vR,F,press,cmax,mach); else one<2,3>(gamma,vL,vR,F,press,cmax,mach); }
  else        { if (dir == 0) one<0,2>(gamma,vL,vR,F,press,cmax,mach); else one<1,2>(gamma,vL,vR,F,press,cmax,mach); }
}
#else
#define PG_NS pg_fast
#define PG_FAST 1
#define PG_HOST_EMU 1
#include "../pluto_b200/csrc/mhd_device.cuh"
void hlld_exact (int dir, int nc, double gamma, const double *vL, const double *vR, double *F, double *press, double *cmax, double *mach);
static long g_fallback = 0;
template <int DIR, int NC> static void one (double gamma, const double *vL, const double *vR, double *F, double *press, double *cmax, double *mach)
{
  pg_fast::Phys ph; ph.gamma = gamma; ph.gmm1 = gamma - 1.0; ph.small_dn = 1e-12; ph.small_pr = 1e-12; ph.igmm1 = 1.0/ph.gmm1;
  pg_fast::riemann_hlld<DIR, NC>(ph, vL, vR, F, *press, *cmax, *mach);
}
static void hlld_fast (int dir, int nc, double gamma, const double *vL, const double *vR, double *F, double *press, double *cmax, double *mach)
{
  if (nc == 3){ if (dir == 0) one<0,3>(gamma,vL,vR,F,press,cmax,mach); else if (dir == 1) one<1,3>(gamma,vL,vR,F,press,cmax,mach); else one<2,3>(gamma,vL,vR,F,press,cmax,mach); }
  else        { if (dir == 0) one<0,2>(gamma,vL,vR,F,press,cmax,mach); else one<1,2>(gamma,vL,vR,F,press,cmax,mach); }
}

static double urand () { return rand ()/(double)RAND_MAX; }

static double run (int dir, int nc, int ncase, double vscale, double bscale, double pscale, double jump)
{
  const double gamma = 5.0/3.0;
  double worst = 0.0;
  const int bn = 4 + dir, vn = 1 + dir;
  for (int c = 0; c < ncase; c++){
    double vL[8], vR[8];
    for (int nv = 0; nv < 8; nv++){
      double base = (nv == 0 ? 0.5 + urand () : nv == 7 ? pscale*(0.1 + urand ()) : (nv <= 3 ? vscale : bscale)*(2*urand () - 1));
      vL[nv] = base;
      vR[nv] = (nv == 0 || nv == 7) ? base*(1.0 + jump*(2*urand () - 1)*0.9) : base + jump*(nv <= 3 ? vscale : bscale)*(2*urand () - 1);
    }
    if (nc == 2){ vL[3] = vR[3] = vL[6] = vR[6] = 0.0; }
    vR[bn] = vL[bn];
    double Fe[8] = {0}, Ff[8] = {0}, pe_, pf_, ce, cf, me, mf;
    hlld_exact (dir, nc, gamma, vL, vR, Fe, &pe_, &ce, &me);
    hlld_fast (dir, nc, gamma, vL, vR, Ff, &pf_, &cf, &mf);
    Fe[vn] += pe_; Ff[vn] += pf_;              // only the sum enters the update
    double c2 = (gamma*vL[7] + vL[4]*vL[4] + vL[5]*vL[5] + vL[6]*vL[6])/vL[0] + vL[1]*vL[1] + vL[2]*vL[2] + vL[3]*vL[3];
    double c1 = std::sqrt (c2);
    // natural scales: mass rho c, momentum rho c^2, energy rho c^3, induction c B ~ sqrt(rho) c^2
    double sc[8] = {vL[0]*c1, vL[0]*c2, vL[0]*c2, vL[0]*c2, std::sqrt (vL[0])*c2, std::sqrt (vL[0])*c2, std::sqrt (vL[0])*c2, vL[0]*c2*c1};
    double err = 0.0;
    for (int nv = 0; nv < 8; nv++){ if (nv == bn) continue; if (nc == 2 && (nv == 3 || nv == 6)) continue; err = std::max (err, std::fabs (Fe[nv] - Ff[nv])/sc[nv]); }
    err = std::max (err, std::fabs (ce - cf)/ce);
    err = std::max (err, std::fabs (me - mf)/(me + 1e-30)*1e-8);       // Mach is single precision in FAST
    if (!(err <= worst)) worst = err;
  }
  return worst;
}

int main ()
{
  srand (12345);
  const double vs[] = {0.0, 0.3, 3.0, 30.0}, bs[] = {0.0, 1e-3, 1.0, 10.0}, js[] = {1e-8, 1e-2, 0.5, 1.0};
  double all = 0.0;
  for (double v : vs) for (double b : bs) for (double j : js){
    double w = 0.0;
    w = std::max (w, run (0, 3, 20000, v, b, 1.0, j));
    w = std::max (w, run (1, 3, 5000, v, b, 1.0, j));
    w = std::max (w, run (2, 3, 5000, v, b, 1.0, j));
    w = std::max (w, run (0, 2, 5000, v, b, 1.0, j));
    w = std::max (w, run (1, 2, 5000, v, b, 1.0, j));
    printf ("v %.1e B %.1e jump %.1e : worst scaled error %.3e\n", v, b, j, w);
    if (!(w <= all)) all = w;
  }
  printf ("overall worst %.3e\n", all);
  return all < 1e-11 ? 0 : 1;
}
#endif
