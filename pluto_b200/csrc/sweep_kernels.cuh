// sweep_kernels.cuh -- the fused directional sweep: reconstruct -> Riemann ->
// CT face-EMF store -> flux difference -> conservative update -> CFL sum,
// one kernel per direction, no intermediate state/flux arrays in HBM.
//
// Replaces the per-pencil loop of UpdateStage (reference
// Src/Time_Stepping/update_stage.c:134-242): gather (:158-164), States
// (:193), Riemann (:194), CT_StoreUpwindEMF (Src/MHD/CT/ct_emf.c:104-190),
// RightHandSide (Src/MHD/rhs.c:193-201), U += rhs (:214-216) and the C_dt
// accumulation (:229-235) plus the g_maxMach reduction (hll_speed.c:105).
//
// Thread mapping.  x1 is the fastest index in HBM, so in every kernel the
// lanes of a warp run along x1:
//   * sweep_x   : a lane owns one zone and its right face; neighbours'
//                 interface states and fluxes travel by warp shuffles.  A
//                 warp covers 32 consecutive entries of the flattened
//                 (row, i) sequence and produces 32-HL-1 updated zones.
//   * sweep_march (x2, x3): a thread owns one pencil (fixed transverse
//                 position) and MARCHES along the sweep direction keeping
//                 the stencil window, the left interface state and the
//                 previous flux in registers: every zone is reconstructed
//                 once, every face solved once, loads are 256-byte rows.
// Sweep box (update_stage.c:144-148): zones DOM along the sweep, DOM+-1 in
// the transverse directions (the extra pencils only feed the face EMFs).
#pragma once
#include "kernels_common.cuh"
#include "mhd_device.cuh"

#ifndef PG_MINB_X
#define PG_MINB_X 2
#endif
#ifndef PG_MINB_MARCH
#define PG_MINB_MARCH 2
#endif

namespace PG_NS {

__device__ __forceinline__ void atomic_max_pos (unsigned long long *slot, double x)
{
  // non-negative doubles order like their bit patterns
  if (x > 0.0) atomicMax (slot, (unsigned long long)__double_as_longlong(x));
}

__device__ __forceinline__ double warp_max (double x)
{
  PG_UNROLL for (int o = 16; o > 0; o >>= 1){
    double y = __shfl_xor_sync (0xffffffffu, x, o);
    x = y > x ? y : x;
  }
  return x;
}

template <int NC>
__device__ __forceinline__ void load_zone (const SweepArgs &a, long long id, double *v)
{
  PG_FOR_NV(nv) v[nv] = __ldg (a.V[nv] + id);
}

#ifndef PG_PREFETCH
#define PG_PREFETCH 1
#endif
__device__ __forceinline__ void prefetch_l1 (const void *p)
{
#if PG_PREFETCH == 1
  asm volatile ("prefetch.global.L1 [%0];" :: "l"(p));
#elif PG_PREFETCH == 2
  asm volatile ("prefetch.global.L2 [%0];" :: "l"(p));
#endif
}

// face EMFs from the induction flux (ct_emf.c:132-134,155-156,175-176) and
// the sign of the mass flux with the UCT_CONTACT dead band (:137-141)
template <int DIR, int NC>
__device__ __forceinline__ void store_face_emf (const SweepArgs &a, long long id, const double *F)
{
  const double eps = 1.e-6;
  signed char s = 0;
  if      (F[RHO] >  eps) s = 1;
  else if (F[RHO] < -eps) s = -1;
  if (DIR == 0){            // e1 = ezi, e2 = eyi
    a.e1[id] = -F[BX2];
    if (NC == 3) a.e2[id] = F[BX3];
  }else if (DIR == 1){      // e1 = ezj, e2 = exj
    a.e1[id] = F[BX1];
    if (NC == 3) a.e2[id] = -F[BX3];
  }else{                    // e1 = eyk, e2 = exk
    a.e1[id] = -F[BX1];
    a.e2[id] =  F[BX2];
  }
  a.sv[id] = s;
}

// ---------------------------------------------------------------------------
//  x1 sweep
// ---------------------------------------------------------------------------
template <int RECON, int SOLVER, int NC>
__global__ void __launch_bounds__(128, PG_MINB_X)
sweep_x_kernel (const __grid_constant__ SweepArgs a)
{
  constexpr int DIR = 0;
  typedef Dirs<DIR> D;
  constexpr int HL = (RECON == RECON_PPM ? 2 : 1);
  constexpr int STRIDE = 32 - HL - 1;
  const Geom &g = a.g;
  Phys ph; ph.gamma = a.ph.gamma; ph.gmm1 = a.ph.gmm1; ph.small_dn = a.ph.small_dn; ph.small_pr = a.ph.small_pr;

  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5;
  const int L = g.n[0] + HL + 1;                 // zones IBEG-HL .. IEND+1 of one row
  const int nrj = g.n[1] + 2;
  const int nrk = (NC == 3 ? g.n[2] + 2 : 1);
  const long long total = (long long)L*nrj*nrk;

  long long e = warp*STRIDE + lane;
  const bool in_range = e < total;
  if (!in_range) e = total - 1;
  const long long row = e/L;
  const int ii = (int)(e - row*L);
  const int i  = g.beg[0] - HL + ii;
  const int jr = (int)(row % nrj), kr = (int)(row / nrj);
  const int j  = g.beg[1] - 1 + jr;
  const int k  = (NC == 3 ? g.beg[2] - 1 + kr : 0);
  const long long id = gidx (g, k, j, i);

  double v[NV], vp[NV], vm[NV];
  load_zone<NC>(a, id, v);
  bool zone_ok = in_range;
  if (RECON == RECON_PLM){
    double dvm[NV], dvp[NV];
    PG_FOR_NV(nv){
      double vl = __ldg (a.V[nv] + id - 1), vr = __ldg (a.V[nv] + id + 1);
      dvm[nv] = v[nv] - vl;
      dvp[nv] = vr - v[nv];
    }
    plm_zone<NC>(v, dvm, dvp, vp, vm);
  }else{
    double vl[NV], vr[NV], vrr[NV], W[NV], Wm[NV];
    PG_FOR_NV(nv){
      vl[nv]  = __ldg (a.V[nv] + id - 1);
      vr[nv]  = __ldg (a.V[nv] + id + 1);
      vrr[nv] = __ldg (a.V[nv] + id + 2);
    }
    ppm_interface<NC>(vl, v, vr, vrr, W);
    PG_FOR_NV(nv) Wm[nv] = __shfl_up_sync (0xffffffffu, W[nv], 1);
    ppm_zone<NC>(v, Wm, W, vp, vm);
    zone_ok = zone_ok && lane >= 1 && ii >= 1;
  }

  // right interface state of face i+1/2 = minus state of zone i+1
  double vR[NV];
  PG_FOR_NV(nv) vR[nv] = __shfl_down_sync (0xffffffffu, vm[nv], 1);
  const bool face_ok = zone_ok && lane <= 30 && lane >= HL - 1 && ii >= HL - 1 && ii <= L - 2;

  const double bn = __ldg (a.Bn + id);            // plm_states.c:271-275
  vp[D::bn] = bn; vR[D::bn] = bn;

  double uL[NV], uR[NV], F[NV], press, cmax, mach;
  prim_to_cons<NC>(ph, vp, uL);
  prim_to_cons<NC>(ph, vR, uR);
  bool ok = riemann<SOLVER, DIR, NC>(ph, vp, vR, uL, uR, F, press, cmax, mach);

  if (face_ok && (lane >= HL || warp == 0)) store_face_emf<DIR, NC>(a, id, F);
  double my_mach = face_ok ? mach : 0.0;
  if (SOLVER == SOLVER_ROE && face_ok && !ok) atomicAdd (a.red + RED_ROEFAIL, 1ull);

  // left-face flux from the lane below
  double Fm[NV], pm, cm;
  Fm[RHO] = __shfl_up_sync (0xffffffffu, F[RHO], 1);
  Fm[MX1] = __shfl_up_sync (0xffffffffu, F[MX1], 1);
  Fm[MX2] = __shfl_up_sync (0xffffffffu, F[MX2], 1);
  if (NC == 3) Fm[MX3] = __shfl_up_sync (0xffffffffu, F[MX3], 1);
  Fm[ENG] = __shfl_up_sync (0xffffffffu, F[ENG], 1);
  pm = __shfl_up_sync (0xffffffffu, press, 1);
  cm = __shfl_up_sync (0xffffffffu, cmax, 1);

  bool upd = face_ok && lane >= HL && ii >= HL;
  upd = upd && j >= g.beg[1] && j <= g.end[1];
  if (NC == 3) upd = upd && k >= g.beg[2] && k <= g.end[2];

  double cd = 0.0;
  if (upd){
    double u0[NV];
    if (a.u_from_v) prim_to_cons<NC>(ph, v, u0);
    else{
      u0[RHO] = a.U[RHO][id]; u0[MX1] = a.U[MX1][id]; u0[MX2] = a.U[MX2][id];
      if (NC == 3) u0[MX3] = a.U[MX3][id];
      u0[ENG] = a.U[ENG][id];
    }
    const double dtdx = a.dtdx;
    double r;
    r = -dtdx*(F[RHO] - Fm[RHO]);                                 a.U[RHO][id] = u0[RHO] + r;
    r = -dtdx*(F[MX1] - Fm[MX1]); r -= dtdx*(press - pm);          a.U[MX1][id] = u0[MX1] + r;
    r = -dtdx*(F[MX2] - Fm[MX2]);                                 a.U[MX2][id] = u0[MX2] + r;
    if (NC == 3){ r = -dtdx*(F[MX3] - Fm[MX3]);                   a.U[MX3][id] = u0[MX3] + r; }
    r = -dtdx*(F[ENG] - Fm[ENG]);                                 a.U[ENG][id] = u0[ENG] + r;
    if (a.stage1){
      cd = 0.5*(cm + cmax)*a.inv_dl;
      if (!a.last_dir) a.cdt[id] = cd;
    }
  }
  if (a.stage1 && a.last_dir){
    cd = warp_max (cd);
    if (lane == 0) atomic_max_pos (a.red + RED_CDT, cd);
  }
  my_mach = warp_max (my_mach);
  if (lane == 0) atomic_max_pos (a.red + RED_MACH, my_mach);
}

// ---------------------------------------------------------------------------
//  x2 / x3 sweeps: marching pencils
// ---------------------------------------------------------------------------
template <int DIR, int RECON, int SOLVER, int NC>
__global__ void __launch_bounds__(128, PG_MINB_MARCH)
sweep_march_kernel (const __grid_constant__ SweepArgs a)
{
  typedef Dirs<DIR> D;
  const Geom &g = a.g;
  Phys ph; ph.gamma = a.ph.gamma; ph.gmm1 = a.ph.gmm1; ph.small_dn = a.ph.small_dn; ph.small_pr = a.ph.small_pr;
  const int lane = threadIdx.x & 31;

  // transverse enumeration: x1 (fastest) and the other transverse direction
  constexpr int TD = (DIR == 1 ? 2 : 1);               // second transverse dimension
  const int np1 = g.n[0] + 2;
  const int np2 = (NC == 3 ? g.n[TD] + 2 : 1);
  const long long npen = (long long)np1*np2;
  long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  const bool in_range = t < npen*a.nchunk;
  if (!in_range) t = npen*a.nchunk - 1;
  const int chunk = (int)(t/npen);
  const long long p = t - (long long)chunk*npen;
  const int t2 = (int)(p/np1), t1 = (int)(p - (long long)t2*np1);
  const int i  = g.beg[0] - 1 + t1;
  const int o2 = (NC == 3 ? g.beg[TD] - 1 + t2 : 0);
  const int nbeg = g.beg[DIR];
  const int c0 = nbeg + chunk*a.chunk_len;
  int c1 = c0 + a.chunk_len - 1; if (c1 > g.end[DIR]) c1 = g.end[DIR];

  bool upd = in_range && i >= g.beg[0] && i <= g.end[0];
  if (NC == 3) upd = upd && o2 >= g.beg[TD] && o2 <= g.end[TD];

  const long long sD = (DIR == 1 ? g.S1 : g.S12);
  // id of zone c0-1 along the pencil
  long long id = (DIR == 1 ? gidx (g, o2, c0 - 1, i) : gidx (g, c0 - 1, o2, i));

  double vb_[NV], vc_[NV], vpL[NV];     // zone f, zone f+1, plus state of zone f
  double vd_[NV];                        // PPM: zone f+2
  double Wf[NV];                         // PPM: interface value at f+1/2
  if (RECON == RECON_PLM){
    double va_[NV], dvm[NV], dvp[NV], vm_unused[NV];
    load_zone<NC>(a, id - sD, va_);
    load_zone<NC>(a, id, vb_);
    load_zone<NC>(a, id + sD, vc_);
    PG_FOR_NV(nv){ dvm[nv] = vb_[nv] - va_[nv]; dvp[nv] = vc_[nv] - vb_[nv]; }
    plm_zone<NC>(vb_, dvm, dvp, vpL, vm_unused);
  }else{
    double vz_[NV], va_[NV], Wm[NV], vm_unused[NV];
    load_zone<NC>(a, id - 2*sD, vz_);
    load_zone<NC>(a, id - sD, va_);
    load_zone<NC>(a, id, vb_);
    load_zone<NC>(a, id + sD, vc_);
    load_zone<NC>(a, id + 2*sD, vd_);
    ppm_interface<NC>(vz_, va_, vb_, vc_, Wm);      // W[c0-2]
    ppm_interface<NC>(va_, vb_, vc_, vd_, Wf);      // W[c0-1]
    ppm_zone<NC>(vb_, Wm, Wf, vpL, vm_unused);
  }

  double Fp[NV], pp = 0.0, cp = 0.0;              // flux through the previous face
  PG_UNROLL for (int nv = 0; nv < NV; nv++) Fp[nv] = 0.0;
  double my_mach = 0.0, my_cdt = 0.0;

  for (int f = c0 - 1; f <= c1; f++, id += sD){
    // id = zone f; interface f+1/2 lies between zone f (vb_) and zone f+1 (vc_)
    double vR[NV], vpn[NV];
    if (f < c1){          // pull the next iteration's rows towards the SM while this face is solved
      const long long ahead = id + (RECON == RECON_PLM ? 3 : 4)*sD;
      PG_FOR_NV(nv) prefetch_l1 (a.V[nv] + ahead);
      prefetch_l1 (a.Bn + id + sD);
      if (upd){
        prefetch_l1 (a.U[RHO] + id + sD); prefetch_l1 (a.U[MX1] + id + sD); prefetch_l1 (a.U[MX2] + id + sD);
        if (NC == 3) prefetch_l1 (a.U[MX3] + id + sD);
        prefetch_l1 (a.U[ENG] + id + sD);
        if (a.stage1) prefetch_l1 (a.cdt + id + sD);
      }
    }
    if (RECON == RECON_PLM){
      double vnx[NV], dvm[NV], dvp[NV];
      load_zone<NC>(a, id + 2*sD, vnx);
      PG_FOR_NV(nv){ dvm[nv] = vc_[nv] - vb_[nv]; dvp[nv] = vnx[nv] - vc_[nv]; }
      plm_zone<NC>(vc_, dvm, dvp, vpn, vR);
      PG_FOR_NV(nv){ vb_[nv] = vc_[nv]; vc_[nv] = vnx[nv]; }
    }else{
      double vnx[NV], Wn[NV];
      load_zone<NC>(a, id + 3*sD, vnx);
      ppm_interface<NC>(vb_, vc_, vd_, vnx, Wn);    // W[f+1]
      ppm_zone<NC>(vc_, Wf, Wn, vpn, vR);
      PG_FOR_NV(nv){ vb_[nv] = vc_[nv]; vc_[nv] = vd_[nv]; vd_[nv] = vnx[nv]; Wf[nv] = Wn[nv]; }
    }
    const double bn = __ldg (a.Bn + id);
    double vL[NV];
    PG_FOR_NV(nv) vL[nv] = vpL[nv];
    vL[D::bn] = bn; vR[D::bn] = bn;

    double uL[NV], uR[NV], F[NV], press, cmax, mach;
    prim_to_cons<NC>(ph, vL, uL);
    prim_to_cons<NC>(ph, vR, uR);
    bool ok = riemann<SOLVER, DIR, NC>(ph, vL, vR, uL, uR, F, press, cmax, mach);
    if (in_range){
      my_mach = mach > my_mach ? mach : my_mach;
      if (SOLVER == SOLVER_ROE && !ok) atomicAdd (a.red + RED_ROEFAIL, 1ull);
      if (f >= c0 || chunk == 0) store_face_emf<DIR, NC>(a, id, F);
    }
    if (upd && f >= c0){
      const double dtdx = a.dtdx;
      double r;
      r = -dtdx*(F[RHO] - Fp[RHO]);                               a.U[RHO][id] += r;
      r = -dtdx*(F[MX1] - Fp[MX1]); if (D::vn == MX1) r -= dtdx*(press - pp);   a.U[MX1][id] += r;
      r = -dtdx*(F[MX2] - Fp[MX2]); if (D::vn == MX2) r -= dtdx*(press - pp);   a.U[MX2][id] += r;
      if (NC == 3){
        r = -dtdx*(F[MX3] - Fp[MX3]); if (D::vn == MX3) r -= dtdx*(press - pp); a.U[MX3][id] += r;
      }
      r = -dtdx*(F[ENG] - Fp[ENG]);                               a.U[ENG][id] += r;
      if (a.stage1){
        double cd = a.cdt[id] + 0.5*(cp + cmax)*a.inv_dl;
        if (a.last_dir) my_cdt = cd > my_cdt ? cd : my_cdt;
        else            a.cdt[id] = cd;
      }
    }
    PG_FOR_NV(nv){ Fp[nv] = F[nv]; vpL[nv] = vpn[nv]; }
    pp = press; cp = cmax;
  }

  my_mach = warp_max (my_mach);
  if (lane == 0) atomic_max_pos (a.red + RED_MACH, my_mach);
  if (a.stage1 && a.last_dir){
    my_cdt = warp_max (my_cdt);
    if (lane == 0) atomic_max_pos (a.red + RED_CDT, my_cdt);
  }
}

// ---------------------------------------------------------------------------
//  launcher for one solver (one translation unit per solver and arithmetic)
// ---------------------------------------------------------------------------
template <int SOLVER>
static int launch_sweep_t (int dir, int recon, const SweepArgs &a, cudaStream_t s)
{
  const Geom &g = a.g;
  const int nc = g.dims;
  const int TPB = 128;
  if (dir == 0){
    const int HL = (recon == RECON_PPM ? 2 : 1);
    const int stride = 32 - HL - 1;
    const long long L = g.n[0] + HL + 1;
    const long long total = L*(g.n[1] + 2)*(nc == 3 ? g.n[2] + 2 : 1);
    const long long nwarp = (total - HL + stride - 1)/stride;
    const unsigned nb = (unsigned)((nwarp*32 + TPB - 1)/TPB);
#define PG_LX(R, C) sweep_x_kernel<R, SOLVER, C><<<nb, TPB, 0, s>>>(a)
    if      (recon == RECON_PLM && nc == 3) PG_LX(RECON_PLM, 3);
    else if (recon == RECON_PLM && nc == 2) PG_LX(RECON_PLM, 2);
    else if (recon == RECON_PPM && nc == 3) PG_LX(RECON_PPM, 3);
    else                                    PG_LX(RECON_PPM, 2);
#undef PG_LX
  }else{
    const int td = (dir == 1 ? 2 : 1);
    const long long npen = (long long)(g.n[0] + 2)*(nc == 3 ? g.n[td] + 2 : 1);
    const long long nthr = npen*a.nchunk;
    const unsigned nb = (unsigned)((nthr + TPB - 1)/TPB);
#define PG_LM(DD, R, C) sweep_march_kernel<DD, R, SOLVER, C><<<nb, TPB, 0, s>>>(a)
    if (dir == 1){
      if      (recon == RECON_PLM && nc == 3) PG_LM(1, RECON_PLM, 3);
      else if (recon == RECON_PLM && nc == 2) PG_LM(1, RECON_PLM, 2);
      else if (recon == RECON_PPM && nc == 3) PG_LM(1, RECON_PPM, 3);
      else                                    PG_LM(1, RECON_PPM, 2);
    }else{
      if (recon == RECON_PLM) PG_LM(2, RECON_PLM, 3);
      else                    PG_LM(2, RECON_PPM, 3);
    }
#undef PG_LM
  }
  return cudaGetLastError () == cudaSuccess ? 1 : -1;
}

} // namespace PG_NS
