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
#include <stdlib.h>
#include "kernels_common.cuh"
#include "mhd_device.cuh"

#ifndef PG_MINB_X
#define PG_MINB_X 3
#endif
#ifndef PG_MINB_MARCH
#define PG_MINB_MARCH 3
#endif

namespace PG_NS {

// resident blocks per SM of a marching sweep: three, except PARABOLIC + roe (50 - 76 B spilled at 168 registers: x3 sweep of
// ot 256^3 3.1 - 3.4 -> 2.56 ms per step with two; LINEAR + roe loses 7 % with two; profiles/r2ag_roe_march_ab.txt)
__host__ __device__ constexpr int march_min_blocks (int recon, int solver)
{ return (recon == 1 /* RECON_PPM */ && solver == 2 /* SOLVER_ROE */) ? 2 : PG_MINB_MARCH; }

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

// Element indices inside the sweeps are 32-bit (pluto_gpu_create refuses blocks with
// 2^31 or more padded zones): one IMAD.WIDE per address instead of 64-bit adds.
__device__ __forceinline__ int gidx32 (const Geom &g, int k, int j, int i)
{
  return (k + g.off[2])*(int)g.S12 + (j + g.off[1])*(int)g.S1 + (i + g.off[0]);
}

template <int NC>
__device__ __forceinline__ void load_zone (const SweepArgs &a, int id, double *v)
{
  PG_FOR_NV(nv) v[nv] = __ldg (a.V[nv] + id);
}

// 8-byte asynchronous global -> shared copy (LDGSTS): no destination register,
// so a thread can pull the rows of its NEXT face while the current one is solved
#ifdef PG_EMU
// host interpreter of the kernels (tests/emu, test infrastructure): the copies of a thread land
// when it waits for them
__device__ __forceinline__ void cp_async8 (double *smem_dst, const double *gsrc) { pg_emu::async_copy8 (smem_dst, gsrc); }
__device__ __forceinline__ void cp_async8_ordered (double *smem_dst, const double *gsrc) { pg_emu::async_copy8 (smem_dst, gsrc); }
__device__ __forceinline__ void cp_async_commit () { pg_emu::async_commit (); }
template <int N> __device__ __forceinline__ void cp_async_wait () { pg_emu::async_wait (N); }
__device__ __forceinline__ void cp_async_wait_all () { pg_emu::async_wait (0); }
#else
__device__ __forceinline__ void cp_async8 (double *smem_dst, const double *gsrc)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared (smem_dst);
  asm volatile ("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(gsrc));
}
// same, as a compiler barrier: smem reads issued before it stay before it (the copy
// overwrites a slot that has just been read)
__device__ __forceinline__ void cp_async8_ordered (double *smem_dst, const double *gsrc)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared (smem_dst);
  asm volatile ("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit () { asm volatile ("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait ()     // all but the N most recent groups
{ asm volatile ("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all () { asm volatile ("cp.async.wait_group 0;" ::: "memory"); }
#endif

// ---------------------------------------------------------------------------
//  TMA staging (fused x1+x2 sweep, SweepArgs.tma): ONE lane of a warp issues one bulk asynchronous copy
//  (cp.async.bulk.shared.global, the TMA engine's 1-D form) per variable and ring row -- 38 consecutive entries of the row,
//  starting at an even index so that source and destination are 16-byte aligned -- and the arrival is counted in bytes on the
//  warp's own mbarrier; the other lanes only wait on its phase.  Replaces 8 (+8 for the four halo lanes) 8-byte cp.async per
//  lane and row together with their address arithmetic.  (A 4-D tensor map over the 8 primitive arrays, one
//  cp.async.bulk.tensor per row, was tried first: cuTensorMapEncodeTiled accepted it, the B200 raised "illegal instruction" on
//  the first UTMALDG with the descriptor in the kernel parameters and in global memory alike -- profiles/r2_tma_ab.txt.)
// ---------------------------------------------------------------------------
#ifdef PG_EMU
__device__ __forceinline__ void mbar_init (unsigned long long *bar) { *bar = 0; }
__device__ __forceinline__ void bulk_copy (double *dst, const double *src, unsigned bytes, unsigned long long *)
{ for (unsigned q = 0; q < bytes/8; q++) dst[q] = src[q]; }             // lands at once
__device__ __forceinline__ void mbar_expect (unsigned long long *, unsigned) {}
__device__ __forceinline__ void mbar_wait (unsigned long long *, unsigned) {}
__device__ __forceinline__ void fence_proxy_async () {}
#else
__device__ __forceinline__ void mbar_init (unsigned long long *bar)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared (bar);
  asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(sa) : "memory");
  asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect (unsigned long long *bar, unsigned bytes)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared (bar);
  asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy (double *dst, const double *src, unsigned bytes, unsigned long long *bar)
{
  const unsigned sd = (unsigned)__cvta_generic_to_shared (dst), sb = (unsigned)__cvta_generic_to_shared (bar);
  asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                :: "r"(sd), "l"(src), "r"(bytes), "r"(sb) : "memory");
}
__device__ __forceinline__ void mbar_wait (unsigned long long *bar, unsigned phase)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared (bar);
  asm volatile ("{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\n D_%=:\n}"
                :: "r"(sa), "r"(phase) : "memory");
}
// generic-proxy reads of a ring row come before the async-proxy write that refills it
__device__ __forceinline__ void fence_proxy_async () { asm volatile ("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

#ifndef PG_MARCH_L2PF
#define PG_MARCH_L2PF 0
#endif
__device__ __forceinline__ void prefetch_l2 (const void *p)
{
#ifndef PG_EMU
  asm volatile ("prefetch.global.L2 [%0];" :: "l"(p));
#endif
}
#ifndef PG_PREFETCH
#define PG_PREFETCH 0
#endif
__device__ __forceinline__ void prefetch_l1 (const void *p)
{
#if defined(PG_EMU)
  (void)p;
#elif PG_PREFETCH == 1
  asm volatile ("prefetch.global.L1 [%0];" :: "l"(p));
#elif PG_PREFETCH == 2
  asm volatile ("prefetch.global.L2 [%0];" :: "l"(p));
#endif
}

// face EMFs from the induction flux (ct_emf.c:132-134,155-156,175-176) and
// the sign of the mass flux with the UCT_CONTACT dead band (:137-141)
template <int DIR, int NC>
__device__ __forceinline__ void store_face_emf_p (double *e1, double *e2, signed char *sv, int id, const double *F)
{
  const double eps = 1.e-6;
  signed char s = 0;
  if      (F[RHO] >  eps) s = 1;
  else if (F[RHO] < -eps) s = -1;
  if (DIR == 0){            // e1 = ezi, e2 = eyi
    e1[id] = -F[BX2];
    if (NC == 3) e2[id] = F[BX3];
  }else if (DIR == 1){      // e1 = ezj, e2 = exj
    e1[id] = F[BX1];
    if (NC == 3) e2[id] = -F[BX3];
  }else{                    // e1 = eyk, e2 = exk
    e1[id] = -F[BX1];
    e2[id] =  F[BX2];
  }
  sv[id] = s;
}
// UCT_HLL: limited velocity slopes of a reconstructed zone (ct_stag_slopes.c:16-44)
template <int NC>
__device__ __forceinline__ void store_vel_slopes (double *const *dv, int id, const double *vp, const double *vm)
{
  dv[0][id] = vp[VX1] - vm[VX1];
  dv[1][id] = vp[VX2] - vm[VX2];
  if (NC == 3) dv[2][id] = vp[VX3] - vm[VX3];
}

// SHOCK_FLATTENING MULTID: minmod for every variable in a zone flagged FLAG_MINMOD
// (plm_states.c:174-180), the HLL flux at an interface next to a zone flagged FLAG_HLL
// CL: CHAR_LIMITING YES (2 components), the slopes of the sweep direction DIR limited on the characteristic variables
// W: grid-dependent weights (RECON_PLMW); pc, n: the weight arrays of the direction and the zone's index along it
template <int NC, bool FLAT, int SKIP = -1, bool CL = false, int DIR = 0, bool W = false>
__device__ __forceinline__ void plm_zone_f (const SweepArgs &a, unsigned fl, const double *v, const double *dvm,
                                            const double *dvp, double *vp, double *vm, const double *const *pc = nullptr, int n = 0)
{
  if (W){ plm_zone_w<NC, SKIP>(a.limiter, pc, n, v, dvm, dvp, vp, vm); return; }
  if (CL && NC == 2) plm_zone_char2<DIR>(*reinterpret_cast<const Phys *>(&a.ph), a.limiter, v, dvm, dvp, vp, vm);
  else if (FLAT && (fl & 1u)) plm_zone_single<NC, SKIP>(2, v, dvm, dvp, vp, vm);
  else                        plm_zone<NC, SKIP>(a.limiter, v, dvm, dvp, vp, vm);
}
template <int SOLVER, int DIR, int NC, bool FLAT>
__device__ __forceinline__ bool riemann_f (const Phys &ph, unsigned fl2, const double *vL, const double *vR,
                                           const double *uL, const double *uR, double *F, double &press,
                                           double &cmax, double &mach, double *pSL, double *pSR)
{
  if (FLAT && (fl2 & 4u) && SOLVER != SOLVER_HLL){
    riemann_flagged<SOLVER, DIR, NC>(ph, vL, vR, uL, uR, F, press, cmax, mach, pSL, pSR);
    return true;
  }
  return riemann<SOLVER, DIR, NC>(ph, vL, vR, uL, uR, F, press, cmax, mach, pSL, pSR);
}

template <int DIR, int NC>
__device__ __forceinline__ void store_face_emf (const SweepArgs &a, int id, const double *F)
{ store_face_emf_p<DIR, NC>(a.e1, a.e2, a.sv, id, F); }

// ---------------------------------------------------------------------------
//  x1 sweep
// ---------------------------------------------------------------------------
#ifndef PG_XROWS
#define PG_XROWS 16          // rows a warp walks through (software-pipelined)
#endif

template <int RECON, int SOLVER, int NC, bool HLL, bool FLAT, bool BF = false, bool CL = false>   // HLL: CT_EMF_AVERAGE == UCT_HLL (fan speeds +
                                                  // velocity slopes); FLAT: SHOCK_FLATTENING MULTID (zone flags); BF: body force;
                                                  // CL: CHAR_LIMITING YES (2 components)
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
  const Phys &ph = *reinterpret_cast<const Phys *>(&a.ph);     // stays in the kernel-parameter constant bank

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

  // rows are enumerated (k, j) with j fastest; (jr, kr) of the row being SOLVED and of
  // the row being FETCHED advance incrementally (no division in the loop)
  const int S1i = (int)g.S1, S12i = (int)g.S12;
  int jr = r_beg % nrj, kr = r_beg / nrj;
  int id = gidx32 (g, (NC == 3 ? g.beg[2] - 1 + kr : 0), g.beg[1] - 1 + jr, i);
  int jr_n = jr, kr_n = kr, id_n = id;               // next row to fetch
  auto advance = [&] (int &jq, int &kq, int &idq){
    if (++jq == nrj){ jq = 0; kq++; idq += S12i - (nrj - 1)*S1i; }
    else idq += S1i;
  };
  auto issue = [&] (int buf){
    double *dst = wb + buf*(9*W);
    PG_FOR_NV(nv) cp_async8 (dst + nv*W + lane + 1, a.V[nv] + id_n);
    cp_async8 (dst + 8*W + lane + 1, a.Bn + id_n);
    if (lane < 4){                                   // stencil halo: entries -1, 32, 33, 34
      const int off = (lane == 0 ? -1 : 31 + lane);
      PG_FOR_NV(nv) cp_async8 (dst + nv*W + off + 1, a.V[nv] + (id_n - lane) + off);
    }
    cp_async_commit ();
    advance (jr_n, kr_n, id_n);
  };

  double my_mach = 0.0, my_cdt = 0.0;
  issue (0);
  for (int r = r_beg; r < r_end; r++, advance (jr, kr, id)){
    const int buf = (r - r_beg) & 1;
    cp_async_wait_all ();
    __syncwarp ();                                   // the other lanes' copies are visible
    if (r + 1 < r_end) issue (buf ^ 1);
    const double *src = wb + buf*(9*W);

    double v[NV], vp[NV], vm[NV];
    PG_FOR_NV(nv) v[nv] = src[nv*W + lane + 1];
    unsigned fl = 0;
    if (FLAT) fl = a.flag[id];
    if (RECON != RECON_PPM){
      double dvm[NV], dvp[NV];
      PG_FOR_NV(nv){
        const double vl = src[nv*W + lane], vr = src[nv*W + lane + 2];
        dvm[nv] = v[nv] - vl;
        dvp[nv] = vr - v[nv];
      }
      plm_zone_f<NC, FLAT, -1, CL, 0, RECON == RECON_PLMW>(a, fl, v, dvm, dvp, vp, vm, a.pc, i);
    }else{
      double vl[NV], vr[NV], vrr[NV], Wi[NV], Wm[NV];
      PG_FOR_NV(nv){
        vl[nv]  = src[nv*W + lane];
        vr[nv]  = src[nv*W + lane + 2];
        vrr[nv] = src[nv*W + lane + 3];
      }
      ppm_interface<NC>(vl, v, vr, vrr, Wi, a.qc, i);
      PG_FOR_NV(nv) Wm[nv] = __shfl_up_sync (0xffffffffu, Wi[nv], 1);
      ppm_zone<NC>(v, Wm, Wi, vp, vm);
      if (FLAT && (fl & 1u)) ppm_flat_zone<NC>(a.pc, i, vl, v, vr, vp, vm);         // FLAG_MINMOD, ppm_states.c:167-181
    }

    // right interface state of face i+1/2 = minus state of zone i+1
    double vR[NV];
    PG_FOR_NV(nv) vR[nv] = __shfl_down_sync (0xffffffffu, vm[nv], 1);
    const double bn = src[8*W + lane + 1];           // plm_states.c:271-275
    vp[D::bn] = bn; vR[D::bn] = bn;

    double uL[NV], uR[NV], F[NV], press, cmax, mach;
    prim_to_cons<NC>(ph, vp, uL);
    prim_to_cons<NC>(ph, vR, uR);
    constexpr bool hll = HLL;
    if (hll && zone_ok && i >= g.beg[0] - 1) store_vel_slopes<NC>(a.dvel, id, vp, vm);
    const unsigned fl2 = FLAT ? (fl | __shfl_down_sync (0xffffffffu, fl, 1)) : 0u;
    bool ok = riemann_f<SOLVER, DIR, NC, FLAT>(ph, fl2, vp, vR, uL, uR, F, press, cmax, mach,
                                               hll && emf_ok ? a.e1 + id : nullptr, hll && emf_ok ? a.e2 + id : nullptr);

    if (emf_ok && !hll) store_face_emf<DIR, NC>(a, id, F);
    if (emf_ok && a.fbn) a.fbn[id] = F[D::bn];
    double pfx = 0.0;                                  // potential of the face i+1/2: gravitational energy flux (rhs.c:388-392)
    if (BF && a.phif && zone_ok){ pfx = __ldg (a.phif + id); F[ENG] += F[RHO]*pfx; }
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

    // interior rows: jr, kr in [1, n] (row 0 and row n+1 are the transverse extension)
    bool upd = upd_i && jr >= 1 && jr <= g.n[1];
    if (NC == 3) upd = upd && kr >= 1 && kr <= g.n[2];
    if (upd){
      double u0[NV];
      if (a.u_from_v) prim_to_cons<NC>(ph, v, u0);
      else{
        u0[RHO] = a.U[RHO][id]; u0[MX1] = a.U[MX1][id]; u0[MX2] = a.U[MX2][id];
        if (NC == 3) u0[MX3] = a.U[MX3][id];
        u0[ENG] = a.U[ENG][id];
      }
      const double dtdx = __ldg (a.dtx + i*a.gs);      // dt/dx[i] (rhs.c:195); a uniform grid has one value (gs = 0)
      double rr;
      rr = -dtdx*(F[RHO] - Fm[RHO]);                                a.U[RHO][id] = u0[RHO] + rr;
      const double r_rho = rr;
      rr = -dtdx*(F[MX1] - Fm[MX1]); rr -= dtdx*(press - pm);
      const double gx = BF ? (a.gf ? __ldg (a.gf + id) : a.grav[DIR]) : 0.0;
      if (BF && a.bfv) rr += __ldg (a.dtp + 3)*v[RHO]*gx;
      if (BF && a.phif) rr -= dtdx*v[RHO]*(pfx - __ldg (a.phif + id - 1));
      a.U[MX1][id] = u0[MX1] + rr;
      rr = -dtdx*(F[MX2] - Fm[MX2]);                                a.U[MX2][id] = u0[MX2] + rr;
      if (NC == 3){ rr = -dtdx*(F[MX3] - Fm[MX3]);                  a.U[MX3][id] = u0[MX3] + rr; }
      rr = -dtdx*(F[ENG] - Fm[ENG]);
      if (BF && a.bfv) rr += __ldg (a.dtp + 3)*0.5*(F[RHO] + Fm[RHO])*gx;
      if (BF && a.phic) rr -= __ldg (a.phic + id)*r_rho;
      a.U[ENG][id] = u0[ENG] + rr;
      if (a.stage1){
        const double cd = 0.5*(cm + cmax)*(a.gs ? __ldg (a.idl + i) : a.inv_dl);
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
#ifndef PG_MARCH_PF
#define PG_MARCH_PF 1
#endif
// Faces of look-ahead of the asynchronous copies, and shared-memory slots (doubles per
// thread).  Measured on B200 (256^3, HLLD+PLM): PF = 2 is 40 % SLOWER than PF = 1 --
// cp.async.ca stages its lines in L1, and three blocks x 67 slots leave only 28 KB of
// L1 next to 206 KB of shared memory (the same cliff appears with four blocks of 52
// slots).  The second face of look-ahead is therefore an L2 prefetch (PG_MARCH_L2PF),
// which needs no landing space.
__host__ __device__ constexpr int march_prefetch (int recon) { return recon == RECON_PPM ? 1 : PG_MARCH_PF; }
// cl: CHAR_LIMITING (the ring then keeps all 8 primitives, otherwise the 7 without the cell-centred normal field);
// r3: the flux difference goes to its own arrays (SweepArgs.R3) -- no U is staged, only C_dt
__host__ __device__ constexpr int march_ring_vars (bool cl) { return cl ? 8 : 7; }
__host__ __device__ constexpr int march_slots (int recon, bool cl = false, bool r3 = false)
{
  return march_ring_vars (cl)*((recon == RECON_PPM ? 3 : 2) + march_prefetch (recon)) + march_ring_vars (cl) + 7 + (recon == RECON_PPM ? 8 : 0)
         + march_prefetch (recon) + (r3 ? 1 : 6)*(march_prefetch (recon) + 1);
}
// R3 (the bench configuration's x3 sweep): 38 slots = 38 KB per block and 128 registers -> FOUR blocks per SM (16 warps, 152 KB of
// shared memory, below the L1 cliff) instead of three: the sweeps are bound by FP64 issue latency, a fourth warp per scheduler
// is what hides it
#ifndef PG_MINB_MARCH_R3
#define PG_MINB_MARCH_R3 4
#endif
template <int DIR, int RECON, int SOLVER, int NC, bool HLL, bool FLAT, bool BF = false, bool CL = false, bool R3 = false>
__global__ void __launch_bounds__(128, R3 ? PG_MINB_MARCH_R3 : march_min_blocks (RECON, SOLVER))
sweep_march_kernel (const __grid_constant__ SweepArgs a)
{
  typedef Dirs<DIR> D;
  const Geom &g = a.g;
  const Phys &ph = *reinterpret_cast<const Phys *>(&a.ph);     // stays in the kernel-parameter constant bank
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

  const int sD = (DIR == 1 ? (int)g.S1 : (int)g.S12);
  // id of zone c0-1 along the pencil
  int id = (DIR == 1 ? gidx32 (g, o2, c0 - 1, i) : gidx32 (g, c0 - 1, o2, i));

  // State carried from face to face lives in SHARED memory, one column per
  // thread (slot*blockDim.x + threadIdx.x: conflict-free), so that it does not
  // occupy registers during the Riemann solve:
  //   ring of zones (8 slots each): the stencil zones f .. f+LA plus the PF-1 zones
  //     already in flight for the next faces.  cp.async lands every zone ONCE,
  //     directly in its ring slot, PF faces ahead of its first use (PF = 2 covers the
  //     HBM latency with one face of work per warp still to spare); the slot of zone
  //     f is refilled as soon as zone f has been read (no copies between slots: the
  //     slot pointers rotate instead);
  //   VP 8 slots: plus state of zone f;  FP 7: previous flux (rho, m1, m2, m3, E),
  //     total pressure and cmax;  PPM only, WF 8: interface value at f+1/2.
  // Other landing slots: BN x PF the face field, UA (PF+1) x 6 conservative variables
  //   + C_dt of the zones to update (consumed at the END of an iteration).
  extern __shared__ double carry_[];
  double *cs = carry_ + threadIdx.x;
  constexpr int CS = 128;                   // = blockDim.x (fixed by the launcher): immediate smem offsets
  constexpr bool PPM = (RECON == RECON_PPM);
  constexpr int SK = CL ? -1 : (int)D::bn;  // the cell-centred normal field is neither staged nor reconstructed (PG_FOR_NV_SKIP),
                                            // except with CHAR_LIMITING, whose eigenvectors are built from it
  constexpr int LA = (PPM ? 3 : 2);         // look-ahead of the stencil
  constexpr int PF = march_prefetch (RECON);
  constexpr int NZ = LA + PF;               // ring slots
  constexpr int NR = march_ring_vars (CL);  // variables per ring zone: slot ZS(nv) of variable nv (the skipped one has none)
  constexpr int NUA = (R3 ? 1 : 6);         // staged per zone to update: 5 U + C_dt, or C_dt alone
  constexpr int S_VP = NR*NZ, S_FP = S_VP + NR, S_WF = S_FP + 7, S_BN = S_WF + (PPM ? 8 : 0), S_UA = S_BN + PF;
  static_assert (S_UA + NUA*(PF + 1) == march_slots (RECON, CL, R3), "shared-memory layout");
#define ZS(nv)   ((SK >= 0 && (nv) > SK) ? (nv) - 1 : (nv))
#define C_VP(nv) cs[(S_VP + ZS(nv))*CS]
#define C_FP(q)  cs[(S_FP + (q))*CS]
#define C_WF(nv) cs[(S_WF + (nv))*CS]
  double *z[NZ], *bnp[PF], *uap[PF + 1];
  PG_UNROLL for (int q = 0; q < NZ; q++) z[q] = cs + NR*q*CS;           // z[q]: zone f+q
  PG_UNROLL for (int q = 0; q < PF; q++) bnp[q] = cs + (S_BN + q)*CS;   // bnp[q]: field of face f+q+1/2
  PG_UNROLL for (int q = 0; q <= PF; q++) uap[q] = cs + (S_UA + NUA*q)*CS; // uap[q]: zone f+q (uap[PF]: free)
  auto fetch_ua = [&] (double *dst, int idz){
    if (R3){ if (a.stage1) cp_async8 (dst, a.cdt + idz); return; }
    cp_async8 (dst, a.U[RHO] + idz); cp_async8 (dst + CS, a.U[MX1] + idz);
    cp_async8 (dst + 2*CS, a.U[MX2] + idz);
    if (NC == 3) cp_async8 (dst + 3*CS, a.U[MX3] + idz);
    cp_async8 (dst + 4*CS, a.U[ENG] + idz);
    if (a.stage1) cp_async8 (dst + 5*CS, a.cdt + idz);
  };
  {
    // group 0: what face c0-1/2 needs (zones c0-1 .. c0-1+LA, its field); groups
    // 1 .. PF-1: the additional zone, field and U of the following faces
    PG_UNROLL for (int q = 0; q <= LA; q++) PG_FOR_NV_SKIP(nv, SK) cp_async8 (z[q] + ZS(nv)*CS, a.V[nv] + id + q*sD);
    cp_async8 (bnp[0], a.Bn + id);
    cp_async_commit ();
    PG_UNROLL for (int q = 1; q < PF; q++){
      PG_FOR_NV_SKIP(nv, SK) cp_async8 (z[LA + q] + ZS(nv)*CS, a.V[nv] + id + (LA + q)*sD);
      cp_async8 (bnp[q], a.Bn + id + q*sD);
      if (upd) fetch_ua (uap[q], id + q*sD);
      cp_async_commit ();
    }
    double vb_[NV], vc_[NV], vpL[NV];
    if (!PPM){
      double va_[NV], dvm[NV], dvp[NV], vm_unused[NV];
      load_zone<NC>(a, id - sD, va_);
      cp_async_wait<PF - 1> ();
      PG_FOR_NV_SKIP(nv, SK){ vb_[nv] = z[0][ZS(nv)*CS]; vc_[nv] = z[1][ZS(nv)*CS]; }
      PG_FOR_NV_SKIP(nv, SK){ dvm[nv] = vb_[nv] - va_[nv]; dvp[nv] = vc_[nv] - vb_[nv]; }
      plm_zone_f<NC, FLAT, SK, CL, DIR, RECON == RECON_PLMW>(a, FLAT ? a.flag[id] : 0u, vb_, dvm, dvp, vpL, vm_unused, a.pc, c0 - 1);
      if (HLL && chunk == 0 && in_range) store_vel_slopes<NC>(a.dvel, id, vpL, vm_unused);
    }else{
      double vz_[NV], va_[NV], vd_[NV], Wm[NV], Wf[NV], vm_unused[NV];
      load_zone<NC>(a, id - 2*sD, vz_);
      load_zone<NC>(a, id - sD, va_);
      cp_async_wait<PF - 1> ();
      PG_FOR_NV_SKIP(nv, SK){ vb_[nv] = z[0][ZS(nv)*CS]; vc_[nv] = z[1][ZS(nv)*CS]; vd_[nv] = z[2][ZS(nv)*CS]; }
      ppm_interface<NC, SK>(vz_, va_, vb_, vc_, Wm, a.qc, c0 - 2);      // W[c0-2]
      ppm_interface<NC, SK>(va_, vb_, vc_, vd_, Wf, a.qc, c0 - 1);      // W[c0-1]
      ppm_zone<NC, SK>(vb_, Wm, Wf, vpL, vm_unused);
      if (FLAT && (a.flag[id] & 1u)) ppm_flat_zone<NC, SK>(a.pc, c0 - 1, va_, vb_, vc_, vpL, vm_unused);
      if (HLL && chunk == 0 && in_range) store_vel_slopes<NC>(a.dvel, id, vpL, vm_unused);
      PG_FOR_NV_SKIP(nv, SK) C_WF(nv) = Wf[nv];
    }
    PG_FOR_NV_SKIP(nv, SK) C_VP(nv) = vpL[nv];
    PG_UNROLL for (int q = 0; q < 7; q++) C_FP(q) = 0.0;
  }
  double my_mach = 0.0, my_cdt = 0.0;
  constexpr bool hll = HLL;
  unsigned flb = 0;                          // flag of zone f
  if (FLAT) flb = a.flag[id];

  for (int f = c0 - 1; f <= c1; f++, id += sD){
    // id = zone f; interface f+1/2 lies between zone f and zone f+1
    double vL[NV], vR[NV];
    unsigned flc = 0;                        // flag of zone f+1
    if (FLAT) flc = a.flag[id + sD];
    cp_async_wait<PF - 1> ();
    const double bn = bnp[0][0];
    const double *ua = uap[0];               // U and C_dt of zone f (landed at least one face ago)
    double rho_f = 0.0;                      // density of zone f (body force)
    {
      double vb_[NV], vc_[NV], vd_[NV], vnx[NV], vpn[NV];
      PG_FOR_NV_SKIP(nv, SK){ vb_[nv] = z[0][ZS(nv)*CS]; vc_[nv] = z[1][ZS(nv)*CS]; vnx[nv] = z[LA][ZS(nv)*CS]; }
      if (BF) rho_f = vb_[RHO];
      if (PPM) PG_FOR_NV_SKIP(nv, SK) vd_[nv] = z[2][ZS(nv)*CS];
      // start pulling what face f+PF+1/2 needs; its new zone replaces zone f, its field
      // the one just read.  Always commit (possibly empty) so that the group count holds.
#if PG_MARCH_L2PF
      if (f + PF + 1 <= c1){                 // one more face ahead: HBM -> L2 only
        PG_FOR_NV(nv) prefetch_l2 (a.V[nv] + id + (LA + PF + 1)*sD);
        prefetch_l2 (a.Bn + id + (PF + 1)*sD);
        if (upd){
          prefetch_l2 (a.U[RHO] + id + (PF + 1)*sD); prefetch_l2 (a.U[MX1] + id + (PF + 1)*sD);
          prefetch_l2 (a.U[MX2] + id + (PF + 1)*sD);
          if (NC == 3) prefetch_l2 (a.U[MX3] + id + (PF + 1)*sD);
          prefetch_l2 (a.U[ENG] + id + (PF + 1)*sD);
          if (a.stage1) prefetch_l2 (a.cdt + id + (PF + 1)*sD);
        }
      }
#endif
      if (f + PF <= c1){
        cp_async8_ordered (z[0], a.V[0] + id + (LA + PF)*sD);
        PG_UNROLL for (int nv = 1; nv < NV; nv++) if (live<NC>(nv) && nv != SK) cp_async8 (z[0] + ZS(nv)*CS, a.V[nv] + id + (LA + PF)*sD);
        cp_async8 (bnp[0], a.Bn + id + PF*sD);
        if (upd) fetch_ua (uap[PF], id + PF*sD);
      }
      cp_async_commit ();
      if (!PPM){
        double dvm[NV], dvp[NV];
        PG_FOR_NV_SKIP(nv, SK){ dvm[nv] = vc_[nv] - vb_[nv]; dvp[nv] = vnx[nv] - vc_[nv]; }
        plm_zone_f<NC, FLAT, SK, CL, DIR, RECON == RECON_PLMW>(a, flc, vc_, dvm, dvp, vpn, vR, a.pc, f + 1);
      }else{
        double Wf[NV], Wn[NV];
        PG_FOR_NV_SKIP(nv, SK) Wf[nv] = C_WF(nv);
        ppm_interface<NC, SK>(vb_, vc_, vd_, vnx, Wn, a.qc, f + 1);    // W[f+1]
        ppm_zone<NC, SK>(vc_, Wf, Wn, vpn, vR);
        if (FLAT && (flc & 1u)) ppm_flat_zone<NC, SK>(a.pc, f + 1, vb_, vc_, vd_, vpn, vR);
        PG_FOR_NV_SKIP(nv, SK) C_WF(nv) = Wn[nv];
      }
      if (hll && in_range) store_vel_slopes<NC>(a.dvel, id + sD, vpn, vR);      // zone f+1
      PG_FOR_NV_SKIP(nv, SK){ vL[nv] = C_VP(nv); C_VP(nv) = vpn[nv]; }
    }
    vL[D::bn] = bn; vR[D::bn] = bn;
    {                        // rotate: zone f+1 becomes zone f, the refilled slots go to the back
      double *t0 = z[0];
      PG_UNROLL for (int q = 0; q + 1 < NZ; q++) z[q] = z[q + 1];
      z[NZ - 1] = t0;
      t0 = bnp[0];
      PG_UNROLL for (int q = 0; q + 1 < PF; q++) bnp[q] = bnp[q + 1];
      bnp[PF - 1] = t0;
      t0 = uap[0];
      PG_UNROLL for (int q = 0; q < PF; q++) uap[q] = uap[q + 1];
      uap[PF] = t0;
    }

    double uL[NV], uR[NV], F[NV], press, cmax, mach;
    prim_to_cons<NC>(ph, vL, uL);
    prim_to_cons<NC>(ph, vR, uR);
    const bool sf = hll && in_range && (f >= c0 || chunk == 0);
    bool ok = riemann_f<SOLVER, DIR, NC, FLAT>(ph, flb | flc, vL, vR, uL, uR, F, press, cmax, mach,
                                               sf ? a.e1 + id : nullptr, sf ? a.e2 + id : nullptr);
    flb = flc;
    double pfd = 0.0;                                  // potential of the face f+1/2
    if (BF && a.phif){ pfd = __ldg (a.phif + id); F[ENG] += F[RHO]*pfd; }
    if (in_range){
      my_mach = mach > my_mach ? mach : my_mach;
      if (SOLVER == SOLVER_ROE && !ok) atomicAdd (a.red + RED_ROEFAIL, 1ull);
      if (!hll && (f >= c0 || chunk == 0)) store_face_emf<DIR, NC>(a, id, F);
      if (a.fbn && (f >= c0 || chunk == 0)) a.fbn[id] = F[D::bn];
    }
    const double pp = C_FP(5), cp = C_FP(6);
    if (upd && f >= c0){
      const double dtdx = __ldg (a.dtx + f*a.gs);      // dt/dx[f] of the zone being updated
      double r;
      // R3: the flux difference alone is stored; the stage completion forms U + r, the same sum
      double *const *Uo = R3 ? a.R3 : a.U;
#define UPD_(q, r) (R3 ? (r) : ua[(q)*CS] + (r))
      r = -dtdx*(F[RHO] - C_FP(0));                               Uo[RHO][id] = UPD_(0, r);
      const double r_rho = r;
      const double dphi = (BF && a.phif) ? pfd - __ldg (a.phif + id - sD) : 0.0;
      const double dtg = BF ? __ldg (a.dtp + 3) : 0.0;
      const double gd = BF ? (a.gf ? __ldg (a.gf + id) : a.grav[DIR]) : 0.0;
      r = -dtdx*(F[MX1] - C_FP(1)); if (D::vn == MX1) r -= dtdx*(press - pp);   Uo[MX1][id] = UPD_(1, r);
      r = -dtdx*(F[MX2] - C_FP(2));
      if (D::vn == MX2){ r -= dtdx*(press - pp); if (BF && a.bfv) r += dtg*rho_f*gd; if (BF && a.phif) r -= dtdx*rho_f*dphi; }
      Uo[MX2][id] = UPD_(2, r);
      if (NC == 3){
        r = -dtdx*(F[MX3] - C_FP(3));
        if (D::vn == MX3){ r -= dtdx*(press - pp); if (BF && a.bfv) r += dtg*rho_f*gd; if (BF && a.phif) r -= dtdx*rho_f*dphi; }
        Uo[MX3][id] = UPD_(3, r);
      }
      r = -dtdx*(F[ENG] - C_FP(4));
      if (BF && a.bfv) r += dtg*0.5*(F[RHO] + C_FP(0))*gd;
      if (BF && a.phic) r -= __ldg (a.phic + id)*r_rho;
      Uo[ENG][id] = UPD_(4, r);
      if (a.stage1){
        double cd = ua[(NUA - 1)*CS] + 0.5*(cp + cmax)*(a.gs ? __ldg (a.idl + f) : a.inv_dl);
        if (a.last_dir) my_cdt = cd > my_cdt ? cd : my_cdt;
        else            a.cdt[id] = cd;
      }
    }
    C_FP(0) = F[RHO]; C_FP(1) = F[MX1]; C_FP(2) = F[MX2];
    if (NC == 3) C_FP(3) = F[MX3];
    C_FP(4) = F[ENG]; C_FP(5) = press; C_FP(6) = cmax;
  }
#undef C_VP
#undef C_FP
#undef C_WF
#undef ZS
#undef UPD_

  my_mach = warp_max (my_mach);
  if (lane == 0) atomic_max_pos (a.red + RED_MACH, my_mach);
  if (a.stage1 && a.last_dir){
    my_cdt = warp_max (my_cdt);
    if (lane == 0) atomic_max_pos (a.red + RED_CDT, my_cdt);
  }
}

// ---------------------------------------------------------------------------
//  fused x1 + x2 sweep (FAST arithmetic): one pass over the primitives does both
//  directions.  A warp owns a 32-entry segment of the x1 rows of one x3 plane and
//  MARCHES along x2: the x2 faces are solved as in sweep_march_kernel (carried state
//  in shared memory), and the x1 faces of the row the march is standing on are solved
//  as in sweep_x_kernel, the neighbours in x1 being the adjacent columns of the SAME
//  shared-memory ring of zone rows (plus four halo columns per warp).  The conservative
//  update of a zone is written once, U = (PrimToCons(V) + rhs_x1) + rhs_x2 (the
//  reference's order, update_stage.c:214-216): no U is read, the primitives and both
//  face fields are read once, the four face EMFs of the two directions are written.
//  HBM bytes per zone: read 8 V + Bx1s + Bx2s, write 5 U + 4 EMF (+ C_dt, 2 sign bytes)
//  = 154 B (+8) instead of 129 + 169 for the two separate sweeps.
// ---------------------------------------------------------------------------
#ifndef PG_MINB_XY
#define PG_MINB_XY 3
#endif
// Resident blocks per SM the fused sweep is compiled for.  LINEAR: three (168 registers, no spills; two blocks with 255 registers
// lose 3 - 9 %).  PARABOLIC: two -- at 168 registers its variants spill 100 - 190 B per thread, with up to 255 they do not and the
// sweep runs 11 - 17 % faster (profiles/r2ad_launch_bounds_ab.txt: rotor 4096^2 PARABOLIC + roe 5.26 -> 4.74 ms per step, + hlld
// 4.85 -> 4.04, blast 256^3 PARABOLIC + hlld 4.20 -> 3.73)
// Roe in 3-D likewise (160 - 190 B spilled at 168 registers): LINEAR + roe 256^3 7.49 -> 6.20 ms per step (profiles/r2af_roe3d_ab.txt);
// in 2-D the LINEAR Roe sweep fits and keeps three blocks (two: - 9 %).
__host__ __device__ constexpr int xy_min_blocks (int recon, int solver, int nc)
{ return (recon == RECON_PPM || (solver == SOLVER_ROE && nc == 3)) ? 2 : PG_MINB_XY; }
__host__ __device__ constexpr int xy_ring_cols () { return 4*38; }             // 4 warps x (32 + 4 halo columns + 2: the bulk copies
                                                                               // of the TMA variant start at an even entry)
__host__ __device__ constexpr int xy_thread_slots (int recon) { return 8 + 7 + (recon == RECON_PPM ? 8 : 0) + 2; }
// ring rows: the stencil rows f .. f+LA, plus (PLM) one free row for the copy in flight so
// that nothing has to be read ahead of its use; with PPM that row would push three blocks
// past the 164 KB carve-out (L1 cliff, see march_prefetch), so row f is read early instead
#ifndef PG_XY_ROWS
#define PG_XY_ROWS 4
#endif
__host__ __device__ constexpr int xy_ring_rows (int recon) { return recon == RECON_PPM ? 4 : PG_XY_ROWS; }
__host__ __device__ constexpr size_t xy_smem_bytes (int recon)
{
  return (size_t)(8*xy_ring_rows (recon)*xy_ring_cols () + xy_thread_slots (recon)*128 + 4)*sizeof (double);   // + 4 mbarriers (TMA)
}

template <int RECON, int SOLVER, int NC, bool HLL, bool FLAT, bool BF = false, bool TMA = false, bool CL = false>
__global__ void __launch_bounds__(128, xy_min_blocks (RECON, SOLVER, NC))
sweep_xy_kernel (const __grid_constant__ SweepArgs a)
{
  typedef Dirs<0> DX;
  typedef Dirs<1> DY;
  constexpr bool PPM = (RECON == RECON_PPM);
  constexpr int HL = (PPM ? 2 : 1);
  constexpr int STRIDE = 32 - HL - 1;
  constexpr int LA = (PPM ? 3 : 2);                  // look-ahead of the x2 stencil
  constexpr int NZ = xy_ring_rows (RECON);           // ring rows: f .. f+LA (+ a free one)
  constexpr int ZF = (NZ > LA + 1 ? NZ - 1 : 0);     // ring row the copy in flight lands in
  constexpr int CW = xy_ring_cols ();                // ring columns per block
  constexpr int VS = 38;                             // ring layout [row][warp][variable][38 entries] (36 used by cp.async)
  constexpr int CS = 128;
  const Geom &g = a.g;
  const Phys &ph = *reinterpret_cast<const Phys *>(&a.ph);     // stays in the kernel-parameter constant bank

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = g.n[0] + HL + 1;                     // zones IBEG-HL .. IEND+1 of one row
  const int nseg = (g.n[0] + STRIDE - 1)/STRIDE;
  const int np2 = (NC == 3 ? g.n[2] + 2 : 1);
  const int gw = (int)(((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5);
  const int seg = gw % nseg;
  const int plane = (gw / nseg) % np2;
  const int chunk = gw / (nseg*np2);
  if (chunk >= a.nchunk) return;                     // whole warps leave together

  const int ii = seg*STRIDE + lane;                  // entry within the row
  const int i  = g.beg[0] - HL + ii;
  const int k  = (NC == 3 ? g.beg[2] - 1 + plane : 0);
  // x1 roles of the lanes (as in sweep_x_kernel)
  bool zone_ok = ii < L;
  if (PPM) zone_ok = zone_ok && lane >= 1 && ii >= 1;
  const bool xface_ok = zone_ok && lane <= 30 && lane >= HL - 1 && ii >= HL - 1 && ii <= L - 2;
  const bool xemf_ok = xface_ok && (lane >= HL || seg == 0);
  const bool upd_i = xface_ok && lane >= HL && ii >= HL;
  // x2 roles: columns IBEG-1 .. IEND+1; overlapping lanes of neighbouring segments are
  // computed twice and stored once
  const bool col_ok = ii < L && i >= g.beg[0] - 1;
  const bool yemf_ok = col_ok && ((lane >= HL && lane <= 30) || (seg == 0 && lane < HL) || (lane == 31 && seg == nseg - 1));
  bool upd = upd_i;
  if (NC == 3) upd = upd && k >= g.beg[2] && k <= g.end[2];

  const int c0 = g.beg[1] + chunk*a.chunk_len;
  int c1 = c0 + a.chunk_len - 1; if (c1 > g.end[1]) c1 = g.end[1];
  const bool last_chunk = (chunk == a.nchunk - 1);
  const int f_end = c1 + (last_chunk ? 1 : 0);       // the last chunk also solves the x1 faces of row JEND+1
  const int sD = (int)g.S1;
  int id = gidx32 (g, k, c0 - 1, i);                 // zone (k, c0-1, i)

  extern __shared__ __align__(128) double carry_[];
  // TMA: the 38-entry window of a row starts at the even entry at or below the warp's entry -1
  const int tpar = TMA ? ((id - lane - 1) & 1) : 0;
  double *ring = carry_ + warp*(8*VS) + 1 + lane + tpar;   // own column; entries -1, 32, 33, 34 of the warp around it
  double *cs = carry_ + 8*NZ*CW + threadIdx.x;
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(carry_ + 8*NZ*CW + xy_thread_slots (RECON)*CS) + warp;
  unsigned tphase = 0;                               // TMA: parity of the warp's mbarrier
  if (TMA){
    if (lane == 0) mbar_init (mbar);
    __syncwarp ();
  }
  constexpr int S_VP = 0, S_FP = 8, S_WF = 15, S_BY = S_WF + (PPM ? 8 : 0), S_BX = S_BY + 1;
  static_assert (S_BX + 1 == xy_thread_slots (RECON), "shared-memory layout");
#define C_VP(nv) cs[(S_VP + (nv))*CS]
#define C_FP(q)  cs[(S_FP + (q))*CS]
#define C_WF(nv) cs[(S_WF + (nv))*CS]
#define C_BY     cs[S_BY*CS]
#define C_BX     cs[S_BX*CS]
  double *z[NZ];
  PG_UNROLL for (int q = 0; q < NZ; q++) z[q] = ring + 8*q*CW;           // z[q]: row f+q
  // TMA: `nrows` consecutive rows starting with the row of zone idr -> ring rows dst, dst + 8*CW, ...; one elected lane,
  // bytes counted on mbar
  constexpr int NLIVE = (NC == 3 ? 8 : 6);
  auto fetch_rows_tma = [&] (double *dst, int idr, int nrows){
    if (lane == 0){
      fence_proxy_async ();
      mbar_expect (mbar, (unsigned)(nrows*NLIVE*VS*sizeof (double)));
      for (int q = 0; q < nrows; q++)
        PG_FOR_NV(nv) bulk_copy (dst - 1 - tpar + q*(8*CW) + nv*VS, a.V[nv] + (idr - 1 - tpar) + q*sD, (unsigned)(VS*sizeof (double)), mbar);
    }
  };
  auto fetch_row = [&] (double *dst, int idr, bool ordered){
    if (ordered) cp_async8_ordered (dst, a.V[0] + idr); else cp_async8 (dst, a.V[0] + idr);
    PG_UNROLL for (int nv = 1; nv < NV; nv++) if (live<NC>(nv)) cp_async8 (dst + nv*VS, a.V[nv] + idr);
    if (lane < 4){                                   // stencil halo: entries -1, 32, 33, 34
      const int off = (lane == 0 ? -1 : 31 + lane);
      PG_FOR_NV(nv) cp_async8 (dst + nv*VS - lane + off, a.V[nv] + (idr - lane) + off);
    }
  };
  {
    constexpr int SKP = CL ? -1 : (int)DY::bn;
    if (TMA) fetch_rows_tma (z[0], id, LA + 1);
    else PG_UNROLL for (int q = 0; q <= LA; q++) fetch_row (z[q], id + q*sD, false);
    cp_async8 (&C_BY, a.Bn2 + id);
    cp_async8 (&C_BX, a.Bn + id);
    cp_async_commit ();
    double vb_[NV], vc_[NV], vpL[NV];
    if (!PPM){
      double va_[NV], dvm[NV], dvp[NV], vm_unused[NV];
      load_zone<NC>(a, id - sD, va_);
      cp_async_wait_all ();
      if (TMA){ mbar_wait (mbar, tphase); tphase ^= 1u; }
      PG_FOR_NV_SKIP(nv, SKP){ vb_[nv] = z[0][nv*VS]; vc_[nv] = z[1][nv*VS]; }
      PG_FOR_NV_SKIP(nv, SKP){ dvm[nv] = vb_[nv] - va_[nv]; dvp[nv] = vc_[nv] - vb_[nv]; }
      plm_zone_f<NC, FLAT, SKP, CL, 1, RECON == RECON_PLMW>(a, FLAT ? a.flag[id] : 0u, vb_, dvm, dvp, vpL, vm_unused, a.pc2, c0 - 1);
      if (HLL && chunk == 0 && col_ok) store_vel_slopes<NC>(a.dvel2, id, vpL, vm_unused);
    }else{
      double vz_[NV], va_[NV], vd_[NV], Wm[NV], Wf[NV], vm_unused[NV];
      load_zone<NC>(a, id - 2*sD, vz_);
      load_zone<NC>(a, id - sD, va_);
      cp_async_wait_all ();
      if (TMA){ mbar_wait (mbar, tphase); tphase ^= 1u; }
      PG_FOR_NV_SKIP(nv, SKP){ vb_[nv] = z[0][nv*VS]; vc_[nv] = z[1][nv*VS]; vd_[nv] = z[2][nv*VS]; }
      ppm_interface<NC, SKP>(vz_, va_, vb_, vc_, Wm, a.qc2, c0 - 2);
      ppm_interface<NC, SKP>(va_, vb_, vc_, vd_, Wf, a.qc2, c0 - 1);
      ppm_zone<NC, SKP>(vb_, Wm, Wf, vpL, vm_unused);
      if (FLAT && (a.flag[id] & 1u)) ppm_flat_zone<NC, SKP>(a.pc2, c0 - 1, va_, vb_, vc_, vpL, vm_unused);
      if (HLL && chunk == 0 && col_ok) store_vel_slopes<NC>(a.dvel2, id, vpL, vm_unused);
      PG_FOR_NV_SKIP(nv, SKP) C_WF(nv) = Wf[nv];
    }
    PG_FOR_NV_SKIP(nv, SKP) C_VP(nv) = vpL[nv];
    PG_UNROLL for (int q = 0; q < 7; q++) C_FP(q) = 0.0;
  }
  double my_mach = 0.0, my_cdt = 0.0;
  constexpr bool hll = HLL;
  bool tpending = false;                             // TMA: a ring row is in flight
  for (int f = c0 - 1; f <= f_end; f++, id += sD){
    // id = zone (k, f, i).  x2 interface f+1/2 lies between rows f and f+1; the x1 faces
    // solved here are those of row f.
    const bool do_y = f <= c1;
    const bool do_x = f >= c0 || chunk == 0;
    cp_async_wait_all ();
    if (TMA && tpending){ mbar_wait (mbar, tphase); tphase ^= 1u; }
    __syncwarp ();                                   // the other lanes' copies are visible
    const double bny = C_BY, bnx = C_BX;
    unsigned flz = 0, fln = 0;               // flags of zone (f, i) and of zone (f+1, i)
    if (FLAT){ flz = a.flag[id]; fln = a.flag[id + sD]; }
    double v[NV], rx[NV], cdx = 0.0;
    double xvl[NV], xvr[NV], xvrr[NV];
    PG_UNROLL for (int nv = 0; nv < NV; nv++) rx[nv] = 0.0;
    auto read_row_f = [&] (){                        // zone (f, i) and its x1 neighbours
      PG_FOR_NV(nv) v[nv] = z[0][nv*VS];
      if (do_x){
        PG_FOR_NV(nv){ xvl[nv] = z[0][nv*VS - 1]; xvr[nv] = z[0][nv*VS + 1]; }
        if (PPM) PG_FOR_NV(nv) xvrr[nv] = z[0][nv*VS + 2];
      }
    };
    if (a.Ec[2] && col_ok){                          // cell-centred EMFs of zone (f, i) for ct_emf_kernel (ct_emf.c:348-388),
      const double *zr = z[0];                       // straight from the ring while few registers are live
      const double ux = zr[VX1*VS], uy = zr[VX2*VS], bx = zr[BX1*VS], by = zr[BX2*VS];
      a.Ec[2][id] = uy*bx - ux*by;
      if (NC == 3){
        const double uz = zr[VX3*VS], bz = zr[BX3*VS];
        a.Ec[0][id] = uz*by - uy*bz;
        a.Ec[1][id] = ux*bz - uz*bx;
      }
    }
    // the copy of row f+LA+1 lands in the free ring row (PLM), or in the row of f itself
    // (PPM), which must then be read first
    if (ZF == 0){ read_row_f (); __syncwarp (); }     // every lane has read its neighbours' columns of row f before they are refilled
    tpending = (f + 1 <= c1);
    if (f + 1 <= c1){ if (TMA) fetch_rows_tma (z[ZF], id + (LA + 1)*sD, 1); else fetch_row (z[ZF], id + (LA + 1)*sD, true); }
    if (f + 1 <= c1) cp_async8 (&C_BY, a.Bn2 + id + sD);
    if (f + 1 <= f_end) cp_async8_ordered (&C_BX, a.Bn + id + sD);
    cp_async_commit ();
    if (ZF != 0) read_row_f ();
    if (do_x){
      double vp[NV], vm[NV];
      if (!PPM){
        double dvm[NV], dvp[NV];
        PG_FOR_NV(nv){ dvm[nv] = v[nv] - xvl[nv]; dvp[nv] = xvr[nv] - v[nv]; }
        plm_zone_f<NC, FLAT, -1, CL, 0, RECON == RECON_PLMW>(a, flz, v, dvm, dvp, vp, vm, a.pc, i);
      }else{
        double Wi[NV], Wm[NV];
        ppm_interface<NC>(xvl, v, xvr, xvrr, Wi, a.qc, i);
        PG_FOR_NV(nv) Wm[nv] = __shfl_up_sync (0xffffffffu, Wi[nv], 1);
        ppm_zone<NC>(v, Wm, Wi, vp, vm);
        if (FLAT && (flz & 1u)) ppm_flat_zone<NC>(a.pc, i, xvl, v, xvr, vp, vm);
      }
      double vR[NV];
      PG_FOR_NV(nv) vR[nv] = __shfl_down_sync (0xffffffffu, vm[nv], 1);
      vp[DX::bn] = bnx; vR[DX::bn] = bnx;
      double uL[NV], uR[NV], F[NV], press, cmax, mach;
      prim_to_cons<NC>(ph, vp, uL);
      prim_to_cons<NC>(ph, vR, uR);
      if (hll && zone_ok && i >= g.beg[0] - 1) store_vel_slopes<NC>(a.dvel, id, vp, vm);
      const unsigned flx2 = FLAT ? (flz | __shfl_down_sync (0xffffffffu, flz, 1)) : 0u;
      bool ok = riemann_f<SOLVER, 0, NC, FLAT>(ph, flx2, vp, vR, uL, uR, F, press, cmax, mach,
                                               hll && xemf_ok ? a.e1 + id : nullptr, hll && xemf_ok ? a.e2 + id : nullptr);
      if (xemf_ok && !hll) store_face_emf_p<0, NC>(a.e1, a.e2, a.sv, id, F);
      double pfx = 0.0;
      if (BF && a.phif && zone_ok){ pfx = __ldg (a.phif + id); F[ENG] += F[RHO]*pfx; }
      if (xface_ok) my_mach = mach > my_mach ? mach : my_mach;
      if (SOLVER == SOLVER_ROE && xface_ok && !ok) atomicAdd (a.red + RED_ROEFAIL, 1ull);
      double Fm[NV], pm, cm;
      Fm[RHO] = __shfl_up_sync (0xffffffffu, F[RHO], 1);
      Fm[MX1] = __shfl_up_sync (0xffffffffu, F[MX1], 1);
      Fm[MX2] = __shfl_up_sync (0xffffffffu, F[MX2], 1);
      if (NC == 3) Fm[MX3] = __shfl_up_sync (0xffffffffu, F[MX3], 1);
      Fm[ENG] = __shfl_up_sync (0xffffffffu, F[ENG], 1);
      pm = __shfl_up_sync (0xffffffffu, press, 1);
      cm = __shfl_up_sync (0xffffffffu, cmax, 1);
      const double dtdx0 = __ldg (a.dtx + i*a.gs);
      rx[RHO] = -dtdx0*(F[RHO] - Fm[RHO]);
      rx[MX1] = -dtdx0*(F[MX1] - Fm[MX1]); rx[MX1] -= dtdx0*(press - pm);
      rx[MX2] = -dtdx0*(F[MX2] - Fm[MX2]);
      if (NC == 3) rx[MX3] = -dtdx0*(F[MX3] - Fm[MX3]);
      rx[ENG] = -dtdx0*(F[ENG] - Fm[ENG]);
      if (BF && a.bfv){
        const double gx = (a.gf ? __ldg (a.gf + id) : a.grav[0]);
        rx[MX1] += __ldg (a.dtp + 3)*v[RHO]*gx;
        rx[ENG] += __ldg (a.dtp + 3)*0.5*(F[RHO] + Fm[RHO])*gx;
      }
      if (BF && a.phif && upd_i){
        rx[MX1] -= dtdx0*v[RHO]*(pfx - __ldg (a.phif + id - 1));
        rx[ENG] -= __ldg (a.phic + id)*rx[RHO];
      }
      cdx = 0.5*(cm + cmax)*(a.gs ? __ldg (a.idl + i) : a.inv_dl);
    }

    // ---------------- x2 face f+1/2 ----------------
    if (do_y){
      double vL[NV], vR[NV], vpn[NV], vc_[NV], vd_[NV], vnx[NV];
      if (ZF != 0) PG_FOR_NV(nv) v[nv] = z[0][nv*VS];      // row f is still in the ring: not kept in registers
      constexpr int SKY = CL ? -1 : (int)DY::bn;       // the cell-centred BX2 is not reconstructed along x2 (PG_FOR_NV_SKIP)
      PG_FOR_NV_SKIP(nv, SKY){ vc_[nv] = z[1][nv*VS]; vnx[nv] = z[LA][nv*VS]; }
      if (PPM) PG_FOR_NV_SKIP(nv, SKY) vd_[nv] = z[2][nv*VS];
      if (!PPM){
        double dvm[NV], dvp[NV];
        PG_FOR_NV_SKIP(nv, SKY){ dvm[nv] = vc_[nv] - v[nv]; dvp[nv] = vnx[nv] - vc_[nv]; }
        plm_zone_f<NC, FLAT, SKY, CL, 1, RECON == RECON_PLMW>(a, fln, vc_, dvm, dvp, vpn, vR, a.pc2, f + 1);
      }else{
        double Wf[NV], Wn[NV];
        PG_FOR_NV_SKIP(nv, SKY) Wf[nv] = C_WF(nv);
        ppm_interface<NC, SKY>(v, vc_, vd_, vnx, Wn, a.qc2, f + 1);       // W[f+1]
        ppm_zone<NC, SKY>(vc_, Wf, Wn, vpn, vR);
        if (FLAT && (fln & 1u)) ppm_flat_zone<NC, SKY>(a.pc2, f + 1, v, vc_, vd_, vpn, vR);
        PG_FOR_NV_SKIP(nv, SKY) C_WF(nv) = Wn[nv];
      }
      if (hll && col_ok) store_vel_slopes<NC>(a.dvel2, id + sD, vpn, vR);       // row f+1
      PG_FOR_NV_SKIP(nv, SKY){ vL[nv] = C_VP(nv); C_VP(nv) = vpn[nv]; }
      vL[DY::bn] = bny; vR[DY::bn] = bny;
      double uL[NV], uR[NV], F[NV], press, cmax, mach;
      prim_to_cons<NC>(ph, vL, uL);
      prim_to_cons<NC>(ph, vR, uR);
      const bool sfy = hll && yemf_ok && (f >= c0 || chunk == 0);
      bool ok = riemann_f<SOLVER, 1, NC, FLAT>(ph, flz | fln, vL, vR, uL, uR, F, press, cmax, mach,
                                               sfy ? a.e3 + id : nullptr, sfy ? a.e4 + id : nullptr);
      double pfy = 0.0;
      if (BF && a.phif2 && col_ok){ pfy = __ldg (a.phif2 + id); F[ENG] += F[RHO]*pfy; }
      if (col_ok){
        my_mach = mach > my_mach ? mach : my_mach;
        if (SOLVER == SOLVER_ROE && !ok) atomicAdd (a.red + RED_ROEFAIL, 1ull);
      }
      if (!hll && yemf_ok && (f >= c0 || chunk == 0)) store_face_emf_p<1, NC>(a.e3, a.e4, a.sv2, id, F);
      const double pp = C_FP(5), cp = C_FP(6);
      if (upd && f >= c0){
        double u0[NV], r;
        const double dtdx1 = __ldg (a.dtx2 + f*a.gs);
        prim_to_cons<NC>(ph, v, u0);
        r = -dtdx1*(F[RHO] - C_FP(0));                                   a.U[RHO][id] = (u0[RHO] + rx[RHO]) + r;
        const double r_rho = r;
        r = -dtdx1*(F[MX1] - C_FP(1));                                   a.U[MX1][id] = (u0[MX1] + rx[MX1]) + r;
        r = -dtdx1*(F[MX2] - C_FP(2)); r -= dtdx1*(press - pp);
        const double gy = BF ? (a.gf2 ? __ldg (a.gf2 + id) : a.grav[1]) : 0.0;
        if (BF && a.bfv) r += __ldg (a.dtp + 3)*v[RHO]*gy;
        if (BF && a.phif2) r -= dtdx1*v[RHO]*(pfy - __ldg (a.phif2 + id - sD));
        a.U[MX2][id] = (u0[MX2] + rx[MX2]) + r;
        if (NC == 3){ r = -dtdx1*(F[MX3] - C_FP(3));                     a.U[MX3][id] = (u0[MX3] + rx[MX3]) + r; }
        r = -dtdx1*(F[ENG] - C_FP(4));
        if (BF && a.bfv) r += __ldg (a.dtp + 3)*0.5*(F[RHO] + C_FP(0))*gy;
        if (BF && a.phic) r -= __ldg (a.phic + id)*r_rho;
        a.U[ENG][id] = (u0[ENG] + rx[ENG]) + r;
        if (a.stage1){
          const double cd = cdx + 0.5*(cp + cmax)*(a.gs ? __ldg (a.idl2 + f) : a.inv_dl2);
          if (a.last_dir) my_cdt = cd > my_cdt ? cd : my_cdt;
          else            a.cdt[id] = cd;
        }
      }
      C_FP(0) = F[RHO]; C_FP(1) = F[MX1]; C_FP(2) = F[MX2];
      if (NC == 3) C_FP(3) = F[MX3];
      C_FP(4) = F[ENG]; C_FP(5) = press; C_FP(6) = cmax;
    }
    {                        // rotate the ring
      double *t0 = z[0];
      PG_UNROLL for (int q = 0; q + 1 < NZ; q++) z[q] = z[q + 1];
      z[NZ - 1] = t0;
    }
  }
#undef C_VP
#undef C_FP
#undef C_WF
#undef C_BY
#undef C_BX
  my_mach = warp_max (my_mach);
  if (lane == 0) atomic_max_pos (a.red + RED_MACH, my_mach);
  if (a.stage1 && a.last_dir){
    my_cdt = warp_max (my_cdt);
    if (lane == 0) atomic_max_pos (a.red + RED_CDT, my_cdt);
  }
}

// ---------------------------------------------------------------------------
//  Chunks of a marching sweep.  A thread (fused sweep: a warp) marches chunk_len zones along the sweep and pays one
//  extra face for its first left state, so long chunks are cheap -- but the blocks of a launch run in rounds of
//  (SMs x resident blocks per SM), and a last round that is nearly empty costs a whole chunk: at 256^3 the fused
//  sweep with 4 chunks of 64 rows is 2322 blocks = 5.2 rounds of 444, i.e. 6 rounds; 3 chunks of 86 rows are 3.9.
//  plan_chunks picks the chunk count with the smallest  rounds x (chunk_len + extra face + set-up)  among those the
//  host allows (a.chunk_len = the longest chunk wanted, a.nchunk = the fewest chunks wanted).
// ---------------------------------------------------------------------------
static int sm_count ()
{
  static int n = 0;
  if (!n){
    int dev = 0; cudaDeviceProp prop;
    cudaGetDevice (&dev);
    n = (cudaGetDeviceProperties (&prop, dev) == cudaSuccess && prop.multiProcessorCount > 0) ? prop.multiProcessorCount : 148;
  }
  return n;
}
static void plan_chunks (SweepArgs &b, int nzones, long long blocks_per_chunk_x1000, int blocks_per_sm)
{
  const long long slots = (long long)sm_count ()*(blocks_per_sm > 0 ? blocks_per_sm : 1);
  const int max_len = b.chunk_len > 0 ? b.chunk_len : nzones;
  int best_n = 0; double best_cost = 0.0;
  for (int nc = 1; nc <= nzones; nc++){
    const int len = (nzones + nc - 1)/nc;
    if (len > max_len) continue;
    if (len < 4 && nc > 1) break;
    const int nce = (nzones + len - 1)/len;               // chunks actually launched with this length
    const long long blocks = (blocks_per_chunk_x1000*nce + 999)/1000;
    const long long rounds = (blocks + slots - 1)/slots;
    // a partly filled last round still costs most of a chunk (the SMs it uses run fewer blocks each, not faster ones)
    const double cost = (double)rounds*(len + 2.0);
    if (best_n == 0 || cost < best_cost*0.995){ best_n = nce; best_cost = cost; b.chunk_len = len; }
  }
  b.nchunk = (nzones + b.chunk_len - 1)/b.chunk_len;
}

template <int SOLVER>
static int launch_sweep_xy_t (int recon, const SweepArgs &a, cudaStream_t s, bool bf)
{
  const Geom &g = a.g;
  const int nc = g.dims;
  const int TPB = 128;
  const int HL = (recon == RECON_PPM ? 2 : 1);
  const int stride = 32 - HL - 1;
  const long long nseg = (g.n[0] + stride - 1)/stride;
  const long long nwarp1 = nseg*(nc == 3 ? g.n[2] + 2 : 1);             // warps of one chunk
  const size_t smem = xy_smem_bytes (recon);
  SweepArgs b = a;
#define PG_LXY1(R, C, H, F) PG_LXY2(R, C, H, F, false)
#define PG_LXY3(R, C) PG_LXYK((sweep_xy_kernel<R, SOLVER, C, false, false, false, true>))
#define PG_LXY2(R, C, H, F, B) PG_LXYK((sweep_xy_kernel<R, SOLVER, C, H, F, B>))
#define PG_LXYK(KF) do { auto kfn = KF;                          \
      static int bps = 0; static unsigned long long devs = 0;                                         \
      if (pg_attr_needed (devs)){ cudaFuncSetAttribute (kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xy_smem_bytes (RECON_PPM)); \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&bps, kfn, TPB, smem) != cudaSuccess || bps < 1) bps = xy_min_blocks (recon, SOLVER, nc); } \
      if (a.plan) plan_chunks (b, g.n[1], nwarp1*32*1000/TPB, bps);                                   \
      const unsigned nb = (unsigned)((nwarp1*b.nchunk*32 + TPB - 1)/TPB);                             \
      kfn<<<nb, TPB, smem, s>>>(b); } while (0)
#define PG_LXY(R, C) do { constexpr bool P = (R == RECON_PLM); const bool fl = a.flag != nullptr;  \
      if (a.char_lim && P && C == 2){                  /* CHAR_LIMITING (2-D, LINEAR): with the UCT_HLL slopes and / or a body force */ \
        if (bf){ if (a.avg == 3) PG_LXYK((sweep_xy_kernel<RECON_PLM, SOLVER, 2, true, false, true, false, true>));               \
                 else            PG_LXYK((sweep_xy_kernel<RECON_PLM, SOLVER, 2, false, false, true, false, true>)); }            \
        else   { if (a.avg == 3) PG_LXYK((sweep_xy_kernel<RECON_PLM, SOLVER, 2, true, false, false, false, true>));              \
                 else            PG_LXYK((sweep_xy_kernel<RECON_PLM, SOLVER, 2, false, false, false, false, true>)); } } else     \
      if (bf){ if (a.avg == 3){ if (fl) PG_LXY2(R, C, true, true, true); else PG_LXY2(R, C, true, false, true); }                         \
               else           { if (fl) PG_LXY2(R, C, false, true, true); else PG_LXY2(R, C, false, false, true); } }                  \
      else if (a.avg == 3){ if (fl) PG_LXY1(R, C, true, true); else PG_LXY1(R, C, true, false); }                       \
      else if (fl)        PG_LXY1(R, C, false, true);                                                                  \
      else if (a.tma)     PG_LXY3(R, C);                               /* TMA staging of the ring rows */              \
      else                PG_LXY1(R, C, false, false); } while (0)
  if      (recon == RECON_PLMW && nc == 3) PG_LXY1(RECON_PLMW, 3, false, false);      // grid weights: plain options only (checked at create)
  else if (recon == RECON_PLMW && nc == 2) PG_LXY1(RECON_PLMW, 2, false, false);
  else if (recon == RECON_PLM && nc == 3) PG_LXY(RECON_PLM, 3);
  else if (recon == RECON_PLM && nc == 2) PG_LXY(RECON_PLM, 2);
  else if (recon == RECON_PPM && nc == 3) PG_LXY(RECON_PPM, 3);
  else                                    PG_LXY(RECON_PPM, 2);
#undef PG_LXY
#undef PG_LXY1
#undef PG_LXY2
#undef PG_LXY3
#undef PG_LXYK
  return pg_launch_status ();
}

// ---------------------------------------------------------------------------
//  launcher for one solver (one translation unit per solver and arithmetic)
// ---------------------------------------------------------------------------
template <int SOLVER>
static int launch_sweep_t (int dir, int recon, const SweepArgs &a, cudaStream_t s, bool bf)
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
#define PG_LX(R, C) do { constexpr bool P = (R == RECON_PLM); const bool fl = a.flag != nullptr;                  \
      if (a.char_lim && P && C == 2){                                                                                  \
        if (bf){ if (a.avg == 3) sweep_x_kernel<RECON_PLM, SOLVER, 2, true, false, true, true><<<nb, TPB, xsmem, s>>>(a);        \
                 else            sweep_x_kernel<RECON_PLM, SOLVER, 2, false, false, true, true><<<nb, TPB, xsmem, s>>>(a); }     \
        else   { if (a.avg == 3) sweep_x_kernel<RECON_PLM, SOLVER, 2, true, false, false, true><<<nb, TPB, xsmem, s>>>(a);       \
                 else            sweep_x_kernel<RECON_PLM, SOLVER, 2, false, false, false, true><<<nb, TPB, xsmem, s>>>(a); } }   \
      else if (bf){ if (a.avg == 3){ if (fl) sweep_x_kernel<R, SOLVER, C, true, true, true><<<nb, TPB, xsmem, s>>>(a);          \
                                     else    sweep_x_kernel<R, SOLVER, C, true, false, true><<<nb, TPB, xsmem, s>>>(a); }    \
                    else           { if (fl) sweep_x_kernel<R, SOLVER, C, false, true, true><<<nb, TPB, xsmem, s>>>(a);      \
                                     else    sweep_x_kernel<R, SOLVER, C, false, false, true><<<nb, TPB, xsmem, s>>>(a); } } \
      else if (a.avg == 3){ if (fl) sweep_x_kernel<R, SOLVER, C, true, true><<<nb, TPB, xsmem, s>>>(a);                     \
                       else    sweep_x_kernel<R, SOLVER, C, true, false><<<nb, TPB, xsmem, s>>>(a); }               \
      else           { if (fl) sweep_x_kernel<R, SOLVER, C, false, true><<<nb, TPB, xsmem, s>>>(a);                 \
                       else    sweep_x_kernel<R, SOLVER, C, false, false><<<nb, TPB, xsmem, s>>>(a); } } while (0)
    if      (recon == RECON_PLMW && nc == 3) sweep_x_kernel<RECON_PLMW, SOLVER, 3, false, false><<<nb, TPB, xsmem, s>>>(a);
    else if (recon == RECON_PLMW && nc == 2) sweep_x_kernel<RECON_PLMW, SOLVER, 2, false, false><<<nb, TPB, xsmem, s>>>(a);
    else if (recon == RECON_PLM && nc == 3) PG_LX(RECON_PLM, 3);
    else if (recon == RECON_PLM && nc == 2) PG_LX(RECON_PLM, 2);
    else if (recon == RECON_PPM && nc == 3) PG_LX(RECON_PPM, 3);
    else                                    PG_LX(RECON_PPM, 2);
#undef PG_LX
  }else{
    const int td = (dir == 1 ? 2 : 1);
    const long long npen = (long long)(g.n[0] + 2)*(nc == 3 ? g.n[td] + 2 : 1);
    size_t smem = (size_t)march_slots (recon)*TPB*sizeof (double);
    SweepArgs b = a;
#define PG_LM1(DD, R, C, H, F) PG_LM2(DD, R, C, H, F, false)
#define PG_LM2(DD, R, C, H, F, B) PG_LMK((sweep_march_kernel<DD, R, SOLVER, C, H, F, B>))
#define PG_LMK(KF) do { auto kfn = KF;             \
      static int bps = 0; static unsigned long long devs = 0;                                         \
      if (pg_attr_needed (devs)){ cudaFuncSetAttribute (kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 72*TPB*8);       \
        if (getenv ("PLUTO_GPU_CARVEOUT")) cudaFuncSetAttribute (kfn, cudaFuncAttributePreferredSharedMemoryCarveout, atoi (getenv ("PLUTO_GPU_CARVEOUT"))); \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&bps, kfn, TPB, smem) != cudaSuccess || bps < 1) bps = march_min_blocks (recon, SOLVER); } \
      if (a.plan) plan_chunks (b, g.n[dir], npen*1000/TPB, bps);                                      \
      const unsigned nb = (unsigned)((npen*b.nchunk + TPB - 1)/TPB);                                  \
      kfn<<<nb, TPB, smem, s>>>(b); } while (0)
#define PG_LM(DD, R, C) do { constexpr bool P = (R == RECON_PLM); const bool fl = a.flag != nullptr;               \
      if (a.char_lim && P && C == 2 && DD == 1){ smem = (size_t)march_slots (RECON_PLM, true)*TPB*sizeof (double);            \
        if (bf){ if (a.avg == 3) PG_LMK((sweep_march_kernel<1, RECON_PLM, SOLVER, 2, true, false, true, true>));             \
                 else            PG_LMK((sweep_march_kernel<1, RECON_PLM, SOLVER, 2, false, false, true, true>)); }          \
        else   { if (a.avg == 3) PG_LMK((sweep_march_kernel<1, RECON_PLM, SOLVER, 2, true, false, false, true>));            \
                 else            PG_LMK((sweep_march_kernel<1, RECON_PLM, SOLVER, 2, false, false, false, true>)); } }        \
      else if (bf){ if (a.avg == 3){ if (fl) PG_LM2(DD, R, C, true, true, true); else PG_LM2(DD, R, C, true, false, true); }   \
                    else           { if (fl) PG_LM2(DD, R, C, false, true, true); else PG_LM2(DD, R, C, false, false, true); } } \
      else if (a.avg == 3){ if (fl) PG_LM1(DD, R, C, true, true); else PG_LM1(DD, R, C, true, false); }                    \
      else           { if (fl) PG_LM1(DD, R, C, false, true); else PG_LM1(DD, R, C, false, false); } } while (0)
    if (dir == 1){
      if      (recon == RECON_PLMW && nc == 3) PG_LM1(1, RECON_PLMW, 3, false, false);
      else if (recon == RECON_PLMW && nc == 2) PG_LM1(1, RECON_PLMW, 2, false, false);
      else if (recon == RECON_PLM && nc == 3) PG_LM(1, RECON_PLM, 3);
      else if (recon == RECON_PLM && nc == 2) PG_LM(1, RECON_PLM, 2);
      else if (recon == RECON_PPM && nc == 3) PG_LM(1, RECON_PPM, 3);
      else                                    PG_LM(1, RECON_PPM, 2);
    }else{
#ifdef PG_FAST
      if (recon == RECON_PLM && a.R3[RHO] && !bf && a.avg != 3 && !a.flag){     // flux difference kept apart, four blocks per SM
        smem = (size_t)march_slots (RECON_PLM, false, true)*TPB*sizeof (double);
        PG_LMK((sweep_march_kernel<2, RECON_PLM, SOLVER, 3, false, false, false, false, true>));
      }else
#endif
      if      (recon == RECON_PLMW) PG_LM1(2, RECON_PLMW, 3, false, false);
      else if (recon == RECON_PLM)  PG_LM(2, RECON_PLM, 3);
      else                          PG_LM(2, RECON_PPM, 3);
    }
#undef PG_LM
#undef PG_LM1
#undef PG_LM2
#undef PG_LMK
  }
  return pg_launch_status ();
}

} // namespace PG_NS
