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
#define PG_MINB_X 4
#endif
#ifndef PG_MINB_MARCH
#define PG_MINB_MARCH 3
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

// 8-byte asynchronous global -> shared copy (LDGSTS): no destination register,
// so a thread can pull the rows of its NEXT face while the current one is solved
__device__ __forceinline__ void cp_async8 (double *smem_dst, const double *gsrc)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared (smem_dst);
  asm volatile ("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit () { asm volatile ("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all () { asm volatile ("cp.async.wait_group 0;" ::: "memory"); }

#ifndef PG_PREFETCH
#define PG_PREFETCH 0
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
#ifndef PG_XROWS
#define PG_XROWS 16          // rows a warp walks through (software-pipelined)
#endif

template <int RECON, int SOLVER, int NC>
__global__ void __launch_bounds__(128, PG_MINB_X)
sweep_x_kernel (const __grid_constant__ SweepArgs a)
{
  // A warp owns one 32-entry SEGMENT of the x1 rows (30/29 updated zones + halo
  // lanes) and walks through PG_XROWS consecutive rows.  While it solves the
  // faces of row r, the 9 row segments (8 primitives + face field, 36 entries
  // with the stencil halo) of row r+1 stream into shared memory with cp.async:
  // no register is tied up by loads in flight and neighbours are read from
  // shared memory instead of three global loads per variable.
  constexpr int DIR = 0;
  typedef Dirs<DIR> D;
  constexpr int HL = (RECON == RECON_PPM ? 2 : 1);
  constexpr int STRIDE = 32 - HL - 1;
  constexpr int W = 36;                              // staged entries per array and row
  const Geom &g = a.g;
  Phys ph; ph.gamma = a.ph.gamma; ph.gmm1 = a.ph.gmm1; ph.small_dn = a.ph.small_dn; ph.small_pr = a.ph.small_pr; ph.igmm1 = a.ph.igmm1;

  extern __shared__ double rowbuf_[];
  const int lane = threadIdx.x & 31;
  double *wb = rowbuf_ + (threadIdx.x >> 5)*(2*9*W);
  const long long gw = ((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5;
  const int L = g.n[0] + HL + 1;                     // zones IBEG-HL .. IEND+1 of one row
  const int nseg = (g.n[0] + STRIDE - 1)/STRIDE;
  const int nrj = g.n[1] + 2;
  const int nrows = nrj*(NC == 3 ? g.n[2] + 2 : 1);
  const int seg = (int)(gw % nseg);
  const int r_beg = (int)(gw / nseg)*PG_XROWS;
  if (r_beg >= nrows) return;
  const int r_end = (r_beg + PG_XROWS < nrows ? r_beg + PG_XROWS : nrows);

  const int ii = seg*STRIDE + lane;                  // entry within the row
  const int i  = g.beg[0] - HL + ii;
  bool zone_ok = ii < L;
  if (RECON == RECON_PPM) zone_ok = zone_ok && lane >= 1 && ii >= 1;
  const bool face_ok = zone_ok && lane <= 30 && lane >= HL - 1 && ii >= HL - 1 && ii <= L - 2;
  const bool emf_ok = face_ok && (lane >= HL || seg == 0);
  const bool upd_i = face_ok && lane >= HL && ii >= HL;

  auto row_id = [&] (int r, int &j, int &k) -> long long {
    const int jr = r % nrj, kr = r / nrj;
    j = g.beg[1] - 1 + jr;
    k = (NC == 3 ? g.beg[2] - 1 + kr : 0);
    return gidx (g, k, j, i);
  };
  auto issue = [&] (int r, int buf){
    int j, k;
    const long long id = row_id (r, j, k);
    double *dst = wb + buf*(9*W);
    PG_FOR_NV(nv) cp_async8 (dst + nv*W + lane + 1, a.V[nv] + id);
    cp_async8 (dst + 8*W + lane + 1, a.Bn + id);
    if (lane < 4){                                   // stencil halo: entries -1, 32, 33, 34
      const int off = (lane == 0 ? -1 : 31 + lane);
      PG_FOR_NV(nv) cp_async8 (dst + nv*W + off + 1, a.V[nv] + (id - lane) + off);
    }
    cp_async_commit ();
  };

  double my_mach = 0.0, my_cdt = 0.0;
  issue (r_beg, 0);
  for (int r = r_beg; r < r_end; r++){
    const int buf = (r - r_beg) & 1;
    cp_async_wait_all ();
    __syncwarp ();                                   // the other lanes' copies are visible
    if (r + 1 < r_end) issue (r + 1, buf ^ 1);
    int j, k;
    const long long id = row_id (r, j, k);
    const double *src = wb + buf*(9*W);

    double v[NV], vp[NV], vm[NV];
    PG_FOR_NV(nv) v[nv] = src[nv*W + lane + 1];
    if (RECON == RECON_PLM){
      double dvm[NV], dvp[NV];
      PG_FOR_NV(nv){
        const double vl = src[nv*W + lane], vr = src[nv*W + lane + 2];
        dvm[nv] = v[nv] - vl;
        dvp[nv] = vr - v[nv];
      }
      plm_zone<NC>(v, dvm, dvp, vp, vm);
    }else{
      double vl[NV], vr[NV], vrr[NV], Wi[NV], Wm[NV];
      PG_FOR_NV(nv){
        vl[nv]  = src[nv*W + lane];
        vr[nv]  = src[nv*W + lane + 2];
        vrr[nv] = src[nv*W + lane + 3];
      }
      ppm_interface<NC>(vl, v, vr, vrr, Wi);
      PG_FOR_NV(nv) Wm[nv] = __shfl_up_sync (0xffffffffu, Wi[nv], 1);
      ppm_zone<NC>(v, Wm, Wi, vp, vm);
    }

    // right interface state of face i+1/2 = minus state of zone i+1
    double vR[NV];
    PG_FOR_NV(nv) vR[nv] = __shfl_down_sync (0xffffffffu, vm[nv], 1);
    const double bn = src[8*W + lane + 1];           // plm_states.c:271-275
    vp[D::bn] = bn; vR[D::bn] = bn;

    double uL[NV], uR[NV], F[NV], press, cmax, mach;
    prim_to_cons<NC>(ph, vp, uL);
    prim_to_cons<NC>(ph, vR, uR);
    bool ok = riemann<SOLVER, DIR, NC>(ph, vp, vR, uL, uR, F, press, cmax, mach);

    if (emf_ok) store_face_emf<DIR, NC>(a, id, F);
    if (face_ok) my_mach = mach > my_mach ? mach : my_mach;
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

    bool upd = upd_i && j >= g.beg[1] && j <= g.end[1];
    if (NC == 3) upd = upd && k >= g.beg[2] && k <= g.end[2];
    if (upd){
      double u0[NV];
      if (a.u_from_v) prim_to_cons<NC>(ph, v, u0);
      else{
        u0[RHO] = a.U[RHO][id]; u0[MX1] = a.U[MX1][id]; u0[MX2] = a.U[MX2][id];
        if (NC == 3) u0[MX3] = a.U[MX3][id];
        u0[ENG] = a.U[ENG][id];
      }
      const double dtdx = __ldg (a.dtp + DIR);
      double rr;
      rr = -dtdx*(F[RHO] - Fm[RHO]);                                a.U[RHO][id] = u0[RHO] + rr;
      rr = -dtdx*(F[MX1] - Fm[MX1]); rr -= dtdx*(press - pm);       a.U[MX1][id] = u0[MX1] + rr;
      rr = -dtdx*(F[MX2] - Fm[MX2]);                                a.U[MX2][id] = u0[MX2] + rr;
      if (NC == 3){ rr = -dtdx*(F[MX3] - Fm[MX3]);                  a.U[MX3][id] = u0[MX3] + rr; }
      rr = -dtdx*(F[ENG] - Fm[ENG]);                                a.U[ENG][id] = u0[ENG] + rr;
      if (a.stage1){
        const double cd = 0.5*(cm + cmax)*a.inv_dl;
        if (a.last_dir) my_cdt = cd > my_cdt ? cd : my_cdt;
        else            a.cdt[id] = cd;
      }
    }
  }
  if (a.stage1 && a.last_dir){
    my_cdt = warp_max (my_cdt);
    if (lane == 0) atomic_max_pos (a.red + RED_CDT, my_cdt);
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
  Phys ph; ph.gamma = a.ph.gamma; ph.gmm1 = a.ph.gmm1; ph.small_dn = a.ph.small_dn; ph.small_pr = a.ph.small_pr; ph.igmm1 = a.ph.igmm1;
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

  // State carried from face to face lives in SHARED memory, one column per
  // thread (slot*blockDim.x + threadIdx.x: conflict-free), so that it does not
  // occupy registers during the Riemann solve (the register allocator would
  // otherwise spill it to local memory and thrash L1):
  //   slots 0-7 zone f, 8-15 zone f+1, 16-23 plus state of zone f,
  //   24-28 previous flux (rho, m1, m2, m3, E), 29 previous total pressure,
  //   30 previous cmax, PPM only: 31-38 zone f+2, 39-46 interface value at f+1/2.
  // Landing slots of the asynchronous copies (cp.async, issued one face ahead):
  //   NX 8 slots: the zone entering the stencil, BN: the face field,
  //   UA 2 x 6: conservative variables + C_dt of the zone to update (double
  //   buffered: consumed at the END of an iteration, refilled at its top).
  extern __shared__ double carry_[];
  double *cs = carry_ + threadIdx.x;
  constexpr int CS = 128;                   // = blockDim.x (fixed by the launcher): immediate smem offsets
  constexpr int S_NX = (RECON == RECON_PLM ? 31 : 47), S_BN = S_NX + 8, S_UA = S_BN + 1;
  constexpr int LA = (RECON == RECON_PLM ? 2 : 3);      // look-ahead of the stencil
#define C_VB(nv) cs[(0 + (nv))*CS]
#define C_VC(nv) cs[(8 + (nv))*CS]
#define C_VP(nv) cs[(16 + (nv))*CS]
#define C_FP(q)  cs[(24 + (q))*CS]
#define C_VD(nv) cs[(31 + (nv))*CS]
#define C_WF(nv) cs[(39 + (nv))*CS]
#define C_NX(nv) cs[(S_NX + (nv))*CS]
#define C_BN     cs[S_BN*CS]
#define C_UA(b, q) cs[(S_UA + 6*(b) + (q))*CS]
  {
    double vb_[NV], vc_[NV], vpL[NV];
    if (RECON == RECON_PLM){
      double va_[NV], dvm[NV], dvp[NV], vm_unused[NV];
      load_zone<NC>(a, id - sD, va_);
      load_zone<NC>(a, id, vb_);
      load_zone<NC>(a, id + sD, vc_);
      PG_FOR_NV(nv){ dvm[nv] = vb_[nv] - va_[nv]; dvp[nv] = vc_[nv] - vb_[nv]; }
      plm_zone<NC>(vb_, dvm, dvp, vpL, vm_unused);
    }else{
      double vz_[NV], va_[NV], vd_[NV], Wm[NV], Wf[NV], vm_unused[NV];
      load_zone<NC>(a, id - 2*sD, vz_);
      load_zone<NC>(a, id - sD, va_);
      load_zone<NC>(a, id, vb_);
      load_zone<NC>(a, id + sD, vc_);
      load_zone<NC>(a, id + 2*sD, vd_);
      ppm_interface<NC>(vz_, va_, vb_, vc_, Wm);      // W[c0-2]
      ppm_interface<NC>(va_, vb_, vc_, vd_, Wf);      // W[c0-1]
      ppm_zone<NC>(vb_, Wm, Wf, vpL, vm_unused);
      PG_FOR_NV(nv){ C_VD(nv) = vd_[nv]; C_WF(nv) = Wf[nv]; }
    }
    PG_FOR_NV(nv){ C_VB(nv) = vb_[nv]; C_VC(nv) = vc_[nv]; C_VP(nv) = vpL[nv]; }
    PG_UNROLL for (int q = 0; q < 7; q++) C_FP(q) = 0.0;
  }
  // first face: the entering zone and the face field
  PG_FOR_NV(nv) cp_async8 (&C_NX(nv), a.V[nv] + id + LA*sD);
  cp_async8 (&C_BN, a.Bn + id);
  cp_async_commit ();
  double my_mach = 0.0, my_cdt = 0.0;

  for (int f = c0 - 1; f <= c1; f++, id += sD){
    // id = zone f; interface f+1/2 lies between zone f and zone f+1
    double vL[NV], vR[NV];
    const int cur = (f - c0) & 1;            // UA buffer holding zone f (filled one iteration ago)
    double vnx[NV];
    cp_async_wait_all ();
    PG_FOR_NV(nv) vnx[nv] = C_NX(nv);
    const double bn = C_BN;
    if (f < c1){          // start pulling everything the NEXT face needs
      PG_FOR_NV(nv) cp_async8 (&C_NX(nv), a.V[nv] + id + (LA + 1)*sD);
      cp_async8 (&C_BN, a.Bn + id + sD);
      if (upd){
        cp_async8 (&C_UA(cur ^ 1, 0), a.U[RHO] + id + sD); cp_async8 (&C_UA(cur ^ 1, 1), a.U[MX1] + id + sD);
        cp_async8 (&C_UA(cur ^ 1, 2), a.U[MX2] + id + sD);
        if (NC == 3) cp_async8 (&C_UA(cur ^ 1, 3), a.U[MX3] + id + sD);
        cp_async8 (&C_UA(cur ^ 1, 4), a.U[ENG] + id + sD);
        if (a.stage1) cp_async8 (&C_UA(cur ^ 1, 5), a.cdt + id + sD);
      }
      cp_async_commit ();
    }
    {
      double vb_[NV], vc_[NV], vpn[NV];
      PG_FOR_NV(nv){ vb_[nv] = C_VB(nv); vc_[nv] = C_VC(nv); }
      if (RECON == RECON_PLM){
        double dvm[NV], dvp[NV];
        PG_FOR_NV(nv){ dvm[nv] = vc_[nv] - vb_[nv]; dvp[nv] = vnx[nv] - vc_[nv]; }
        plm_zone<NC>(vc_, dvm, dvp, vpn, vR);
        PG_FOR_NV(nv){ C_VB(nv) = vc_[nv]; C_VC(nv) = vnx[nv]; }
      }else{
        double vd_[NV], Wf[NV], Wn[NV];
        PG_FOR_NV(nv){ vd_[nv] = C_VD(nv); Wf[nv] = C_WF(nv); }
        ppm_interface<NC>(vb_, vc_, vd_, vnx, Wn);    // W[f+1]
        ppm_zone<NC>(vc_, Wf, Wn, vpn, vR);
        PG_FOR_NV(nv){ C_VB(nv) = vc_[nv]; C_VC(nv) = vd_[nv]; C_VD(nv) = vnx[nv]; C_WF(nv) = Wn[nv]; }
      }
      PG_FOR_NV(nv){ vL[nv] = C_VP(nv); C_VP(nv) = vpn[nv]; }
    }
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
    const double pp = C_FP(5), cp = C_FP(6);
    if (upd && f >= c0){
      const double dtdx = __ldg (a.dtp + DIR);
      double r;
      r = -dtdx*(F[RHO] - C_FP(0));                               a.U[RHO][id] = C_UA(cur, 0) + r;
      r = -dtdx*(F[MX1] - C_FP(1)); if (D::vn == MX1) r -= dtdx*(press - pp);   a.U[MX1][id] = C_UA(cur, 1) + r;
      r = -dtdx*(F[MX2] - C_FP(2)); if (D::vn == MX2) r -= dtdx*(press - pp);   a.U[MX2][id] = C_UA(cur, 2) + r;
      if (NC == 3){
        r = -dtdx*(F[MX3] - C_FP(3)); if (D::vn == MX3) r -= dtdx*(press - pp); a.U[MX3][id] = C_UA(cur, 3) + r;
      }
      r = -dtdx*(F[ENG] - C_FP(4));                               a.U[ENG][id] = C_UA(cur, 4) + r;
      if (a.stage1){
        double cd = C_UA(cur, 5) + 0.5*(cp + cmax)*a.inv_dl;
        if (a.last_dir) my_cdt = cd > my_cdt ? cd : my_cdt;
        else            a.cdt[id] = cd;
      }
    }
    C_FP(0) = F[RHO]; C_FP(1) = F[MX1]; C_FP(2) = F[MX2];
    if (NC == 3) C_FP(3) = F[MX3];
    C_FP(4) = F[ENG]; C_FP(5) = press; C_FP(6) = cmax;
  }
#undef C_VB
#undef C_VC
#undef C_VP
#undef C_FP
#undef C_VD
#undef C_WF
#undef C_NX
#undef C_BN
#undef C_UA

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
    const long long nseg = (g.n[0] + stride - 1)/stride;
    const long long nrows = (long long)(g.n[1] + 2)*(nc == 3 ? g.n[2] + 2 : 1);
    const long long nwarp = nseg*((nrows + PG_XROWS - 1)/PG_XROWS);
    const unsigned nb = (unsigned)((nwarp*32 + TPB - 1)/TPB);
    const size_t xsmem = (size_t)(TPB/32)*2*9*36*sizeof (double);
#define PG_LX(R, C) sweep_x_kernel<R, SOLVER, C><<<nb, TPB, xsmem, s>>>(a)
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
    const size_t smem = (size_t)((recon == RECON_PPM ? 47 : 31) + 8 + 1 + 12)*TPB*sizeof (double);
#define PG_LM(DD, R, C) do { auto kfn = sweep_march_kernel<DD, R, SOLVER, C>;                       \
      static bool attr_set = false;                                                                   \
      if (!attr_set){ cudaFuncSetAttribute (kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 72*TPB*8); attr_set = true; } \
      kfn<<<nb, TPB, smem, s>>>(a); } while (0)
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
