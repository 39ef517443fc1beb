// ctu_kernels.cuh -- corner-transport-upwind time stepping (TIME_STEPPING HANCOCK with
// CONSTRAINED_TRANSPORT): the reference's AdvanceStep of Src/Time_Stepping/ctu_step.c:142-727
// with the primitive MUSCL-Hancock predictor (Src/States/hancock.c:33-142,
// Src/MHD/prim_eqn.c:26-89), CheckPrimStates (Src/check_states.c:18-72) and CTU_CT_Source
// (ctu_step.c:731-816).
//
// Two sweeps per direction and step, both fused like the RK sweeps (reconstruct -> Riemann ->
// face-EMF store -> flux difference, nothing but the result in HBM):
//   PHASE 0, predictor  (ctu_step.c:283-420): normal predictor states from V^n, Riemann problem,
//            face EMFs, half-step right-hand side rhs[dir] (dt/2) of zones DOM+-1, on the pencils
//            of DOM+-2.
//   PHASE 1, corrector  (ctu_step.c:517-640): the normal predictor states are RECOMPUTED from V^n
//            (PLM + Hancock + PrimToCons + CT source: ~250 flops per zone and direction) instead of
//            being stored and read back as the reference's Up/Um arrays are (16 doubles written and
//            read per zone and direction = 256 B: the recomputation is ~6 times cheaper on a B200
//            and saves 48 arrays of HBM), corrected with the transverse half-step right-hand sides,
//            given the half-step staggered normal field, mapped to primitives; Riemann problem, face
//            EMFs, U += rhs (dt).
// Thread mapping as in sweep_kernels.cuh: lanes always run along x1.  x1 sweep: a lane owns a zone
// and its right face, the neighbour's minus state and the left face's flux travel by warp shuffles
// (30 updated zones per warp).  x2 / x3 sweeps: a thread marches along the sweep with the stencil
// window, the previous zone's plus state and the previous face's flux in registers.
// In between, ctu_half_kernel forms U^n + sum of the half-step right-hand sides with the half-step
// face-averaged field and maps it to V^{n+1/2} on DOM+-1 (ctu_step.c:427-497).
#pragma once
#include "sweep_kernels.cuh"

#ifndef PG_CTU_MINB_X
#define PG_CTU_MINB_X 3
#endif
#ifndef PG_CTU_MINB_M
#define PG_CTU_MINB_M 2
#endif

namespace PG_NS {

// A dV/dx of the primitive MHD equations along DIR (prim_eqn.c:26-89; the Powell term is absent with CT)
template <int DIR, int NC>
__device__ __forceinline__ void prim_rhs (const Phys &ph, const double *v, const double *dv, double *Adv)
{
  typedef Dirs<DIR> D;
  const double tau = pg_rcp (v[RHO]);
  double scrh;
  PG_FOR_NV(nv) Adv[nv] = 0.0;
  Adv[RHO] = v[D::vn]*dv[RHO] + v[RHO]*dv[D::vn];
  if (NC == 3) scrh = 0.0 + v[D::bt]*dv[D::bt] + v[D::bb]*dv[D::bb];
  else         scrh = 0.0 + v[D::bt]*dv[D::bt];
  Adv[D::vn] = v[D::vn]*dv[D::vn] + tau*(dv[PRS] + scrh);
  Adv[D::vt] = v[D::vn]*dv[D::vt] - tau*v[D::bn]*dv[D::bt];
  if (NC == 3) Adv[D::vb] = v[D::vn]*dv[D::vb] - tau*v[D::bn]*dv[D::bb];
  Adv[D::bn] = 0.0;
  Adv[D::bt] = v[D::bt]*dv[D::vn] - v[D::bn]*dv[D::vt] + v[D::vn]*dv[D::bt];
  if (NC == 3) Adv[D::bb] = v[D::bb]*dv[D::vn] - v[D::bn]*dv[D::vb] + v[D::vn]*dv[D::bb];
  Adv[PRS] = ph.gamma*v[PRS]*dv[D::vn] + v[D::vn]*dv[PRS];
}

// normal predictor of one zone: PLM states (plm_states.c:134-275, the normal component takes the
// staggered field of the zone's two faces), evolved by dt/2 with the primitive equations
// (hancock.c:96-124), first order where density or pressure turned negative (check_states.c:35-66)
// chtr_dtdx > 0: TIME_STEPPING CHARACTERISTIC_TRACING (2 components), the states are traced along the characteristics with
// dt/dx = chtr_dtdx (char_tracing2, mhd_device.cuh) instead of the Hancock half step
template <int DIR, int NC, bool FLAT>
__device__ __forceinline__ void ctu_states (const Phys &ph, int limiter, unsigned fl, const double *vl, const double *v,
                                            const double *vr, double bsm, double bsp, double dt_2, double d_dl,
                                            double src_n, double *vp, double *vm, double chtr_dtdx = 0.0, int char_lim = 0)
{
  typedef Dirs<DIR> D;
  double dvm[NV], dvp[NV], dv[NV], Adv[NV];
  PG_FOR_NV(nv){ dvm[nv] = v[nv] - vl[nv]; dvp[nv] = vr[nv] - v[nv]; }
  if (FLAT && (fl & 1u)) plm_zone_single<NC>(2, v, dvm, dvp, vp, vm);
  else if (NC == 2 && DIR < 2 && char_lim) plm_zone_char2<DIR>(ph, limiter, v, dvm, dvp, vp, vm);     // CHAR_LIMITING YES
  else                   plm_zone<NC>(limiter, v, dvm, dvp, vp, vm);
  vp[D::bn] = bsp; vm[D::bn] = bsm;
  if (NC == 2 && DIR < 2 && chtr_dtdx > 0.0){
    char_tracing2<DIR>(ph, v, chtr_dtdx, vp, vm);
  }else{
  PG_FOR_NV(nv) dv[nv] = vp[nv] - vm[nv];
  prim_rhs<DIR, NC>(ph, v, dv, Adv);
  PG_FOR_NV(nv){                                     // src_n: PrimSource of the normal velocity (body force, prim_eqn.c:289-360)
    const double scrh = dt_2*(d_dl*Adv[nv] - (nv == D::vn ? src_n : 0.0));
    vp[nv] -= scrh;
    vm[nv] -= scrh;
  }
  }
  bool sw = (vp[PRS] < 0.0) || (vm[PRS] < 0.0);
  sw = sw || (vp[RHO] < 0.0) || (vm[RHO] < 0.0);
  if (sw){
    const double bp = vp[D::bn], bm = vm[D::bn];
    PG_FOR_NV(nv) vm[nv] = vp[nv] = v[nv];
    vp[D::bn] = bp; vm[D::bn] = bm;
  }
}

// corrector states of one zone: conservative predictor states + CT source (ctu_step.c:731-816) +
// transverse half-step right-hand sides, half-step staggered normal field, ConsToPrim
// (ctu_step.c:545-603).  Returns the number of ConsToPrim repairs.
template <int DIR, int NC>
__device__ __forceinline__ int ctu_correct (const Phys &ph, const double *vc, const double *dU, double dt2_dx,
                                            double bhm, double bhp, double *vp, double *vm, double *up, double *um)
{
  typedef Dirs<DIR> D;
  prim_to_cons<NC>(ph, vp, up);
  prim_to_cons<NC>(ph, vm, um);
  const double db = dt2_dx*(up[D::bn] - um[D::bn]);
  up[MX1] += vc[BX1]*db; um[MX1] += vc[BX1]*db;
  up[MX2] += vc[BX2]*db; um[MX2] += vc[BX2]*db;
  if (NC == 3){ up[MX3] += vc[BX3]*db; um[MX3] += vc[BX3]*db; }
  up[D::bt] += vc[D::vt]*db; um[D::bt] += vc[D::vt]*db;
  if (NC == 3){ up[D::bb] += vc[D::vb]*db; um[D::bb] += vc[D::vb]*db; }
  double scrh;
  if (NC == 3) scrh = vc[VX1]*vc[BX1] + vc[VX2]*vc[BX2] + vc[VX3]*vc[BX3];
  else         scrh = vc[VX1]*vc[BX1] + vc[VX2]*vc[BX2];
  up[ENG] += scrh*db; um[ENG] += scrh*db;
  PG_FOR_NV(nv){ up[nv] = up[nv] + dU[nv]; um[nv] = um[nv] + dU[nv]; }
  up[D::bn] = bhp; um[D::bn] = bhm;
  int nfl = cons_to_prim<NC>(ph, um, vm);
  nfl += cons_to_prim<NC>(ph, up, vp);
  return nfl;
}

// ---------------------------------------------------------------------------
//  x1 sweep.  A warp owns one 32-entry segment of the rows (30 updated zones) and walks through
//  PG_CTU_XROWS consecutive rows; while it solves row r, what row r+1 needs streams into the warp's
//  shared-memory buffer by cp.async (primitives with a two-entry halo, face fields, and in the
//  corrector the transverse right-hand sides), so the ~30 loads per zone are never waited for.
// ---------------------------------------------------------------------------
#ifndef PG_CTU_XROWS
#define PG_CTU_XROWS 8
#endif
__host__ __device__ constexpr int ctu_x_slot (int phase) { return 9*36 + (phase == 0 ? 0 : 36 + 16*32); }   // doubles per warp and row
template <int PHASE, int SOLVER, int NC, bool FLAT>
__global__ void __launch_bounds__(128, PG_CTU_MINB_X)
ctu_sweep_x_kernel (const __grid_constant__ CtuArgs a)
{
  constexpr int DIR = 0;
  typedef Dirs<DIR> D;
  constexpr int E = (PHASE == 0 ? 2 : 1);            // transverse extension of the pencils, states of zones DOM+-E
  constexpr int STRIDE = 30, W = 36;
  constexpr int NQ = ctu_x_slot (PHASE);
  constexpr int Q_BS = 8*W, Q_BH = 9*W, Q_RA = 10*W, Q_RB = 10*W + 8*32;
  const Geom &g = a.g;
  const Phys &ph = *reinterpret_cast<const Phys *>(&a.ph);
  extern __shared__ double xbuf_[];
  const int lane = threadIdx.x & 31;
  double *wb = xbuf_ + (threadIdx.x >> 5)*(2*NQ);
  const int gw = (int)(((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5);
  const int L = g.n[0] + 2*E;                        // zones with states in one row
  const int nseg = (L - 2 + STRIDE - 1)/STRIDE;
  const int nrj = g.n[1] + 2*E;
  const int nrows = nrj*(NC == 3 ? g.n[2] + 2*E : 1);
  const int seg = gw % nseg;
  const int r_beg = (gw / nseg)*PG_CTU_XROWS;
  if (r_beg >= nrows) return;                        // whole warps leave together
  const int r_end = (r_beg + PG_CTU_XROWS < nrows ? r_beg + PG_CTU_XROWS : nrows);
  const int ii = seg*STRIDE + lane;
  const bool zone_ok = ii < L;
  const bool face_ok = zone_ok && lane <= 30 && ii <= L - 2;
  const bool emf_ok = face_ok && (lane >= 1 || seg == 0);
  const bool rhs_ok = face_ok && lane >= 1;
  // zone L (one past the last zone with a state) is the last one read; entries beyond it are clamped
  const int iic = ii < L ? ii : L;
  const int ih = (lane == 0 ? seg*STRIDE - 1 : (seg*STRIDE + 32 < L ? seg*STRIDE + 32 : L));   // halo entries 0 and 33

  // rows are enumerated (k, j) with j fastest; idr = index of zone IBEG-E of the row
  const int S1i = (int)g.S1, S12i = (int)g.S12;
  int jr = r_beg % nrj, kr = r_beg / nrj;
  int idr = gidx32 (g, (NC == 3 ? g.beg[2] - E + kr : 0), g.beg[1] - E + jr, g.beg[0] - E);
  int jr_n = jr, kr_n = kr, idr_n = idr;
  auto advance = [&] (int &jq, int &kq, int &idq){
    if (++jq == nrj){ jq = 0; kq++; idq += S12i - (nrj - 1)*S1i; }
    else idq += S1i;
  };
  auto issue = [&] (int buf){
    double *dst = wb + buf*NQ;
    const int ido = idr_n + iic;
    PG_FOR_NV(nv) cp_async8 (dst + nv*W + lane + 1, a.V0[nv] + ido);
    cp_async8 (dst + Q_BS + lane + 1, a.Bs0[DIR] + ido);
    if (lane < 2){
      const int e = (lane == 0 ? 0 : 33);
      PG_FOR_NV(nv) cp_async8 (dst + nv*W + e, a.V0[nv] + idr_n + ih);
    }
    if (lane == 0) cp_async8 (dst + Q_BS, a.Bs0[DIR] + idr_n + ih);
    if (PHASE == 1){
      cp_async8 (dst + Q_BH + lane + 1, a.Bsh[DIR] + ido);
      if (lane == 0) cp_async8 (dst + Q_BH, a.Bsh[DIR] + idr_n + ih);
      PG_FOR_NV(nv) cp_async8 (dst + Q_RA + nv*32 + lane, a.rhs[1][nv] + ido);
      if (NC == 3) PG_FOR_NV(nv) cp_async8 (dst + Q_RB + nv*32 + lane, a.rhs[2][nv] + ido);
    }
    cp_async_commit ();
    advance (jr_n, kr_n, idr_n);
  };

  // dt/dx, (dt/2)/dx = half of it (exact) and 1/dx of the lane's zone: per-zone arrays on a non-uniform grid (gs = 1; rhs.c:195,
  // ctu_step.c:310, hancock.c:83), the scalars otherwise
  const int izn = g.beg[0] - E + iic;
  const double dtdx = __ldg (a.dtx + izn*a.gs), dt2_dx = 0.5*dtdx, dt_2 = 0.5*__ldg (a.dtp + 3);
  const double d_dl = a.gs ? __ldg (a.idl + izn) : a.inv_dl;
  double my_mach = 0.0, my_cdt = 0.0;
  int nfl_tot = 0;
  issue (0);
  for (int r = r_beg; r < r_end; r++, advance (jr, kr, idr)){
    const int buf = (r - r_beg) & 1;
    cp_async_wait_all ();
    __syncwarp ();                                   // the other lanes' copies are visible, row r-1 has been read
    if (r + 1 < r_end) issue (buf ^ 1);
    const double *src = wb + buf*NQ;
    const int id = idr + iic;

    double vl[NV], v[NV], vr[NV], vp[NV], vm[NV], up[NV], um[NV];
    PG_FOR_NV(nv){ vl[nv] = src[nv*W + lane]; v[nv] = src[nv*W + lane + 1]; vr[nv] = src[nv*W + lane + 2]; }
    const double bsm = src[Q_BS + lane], bsp = src[Q_BS + lane + 1];
    unsigned fl = 0;
    if (FLAT) fl = a.flag[id];
    const double gz = a.bf ? (a.gf ? __ldg (a.gf + id) : a.grav[DIR]) : 0.0;      // body force at the zone centre
    double src_n = 0.0, dphi = 0.0;                    // PrimSource of the normal velocity (prim_eqn.c:289-307)
    if (a.bf) src_n += gz;
    if (a.phif){ dphi = __ldg (a.phif + id) - __ldg (a.phif + id - 1); src_n -= pg_div (dphi, 1.0*(a.dxz ? __ldg (a.dxz + izn) : g.dx[DIR])); }
    ctu_states<DIR, NC, FLAT>(ph, a.limiter, fl, vl, v, vr, bsm, bsp, dt_2, d_dl, src_n, vp, vm, a.chtr ? dtdx : 0.0, a.char_lim);
    // body force: density of stateC -- the half-step zone average of hancock.c:136-141 in the predictor,
    // V^{n+1/2} (ctu_step.c:566-570) in the corrector
    double rho_c = 0.0;
    if (a.bf || a.phif) rho_c = (PHASE == 0 ? 0.5*(vp[RHO] + vm[RHO]) : __ldg (a.Vh[RHO] + id));
    if (PHASE == 0){
      prim_to_cons<NC>(ph, vp, up);
      prim_to_cons<NC>(ph, vm, um);
    }else{
      double dU[NV];
      PG_FOR_NV(nv){                                 // ctu_step.c:551-562, the reference's order
        if (NC == 3) dU[nv] = 0.0 + src[Q_RA + nv*32 + lane] + src[Q_RB + nv*32 + lane];
        else         dU[nv] = 0.0 + src[Q_RA + nv*32 + lane];
      }
      const int n = ctu_correct<DIR, NC>(ph, v, dU, dt2_dx, src[Q_BH + lane], src[Q_BH + lane + 1], vp, vm, up, um);
      if (zone_ok && (lane >= 1 || seg == 0)) nfl_tot += n;    // zones shared by two segments count once
    }

    double vR[NV], uR[NV];
    PG_FOR_NV(nv){ vR[nv] = __shfl_down_sync (0xffffffffu, vm[nv], 1); uR[nv] = __shfl_down_sync (0xffffffffu, um[nv], 1); }
    const unsigned fl2 = FLAT ? (fl | __shfl_down_sync (0xffffffffu, fl, 1)) : 0u;
    double F[NV], press, cmax, mach;
    const bool ok = riemann_f<SOLVER, DIR, NC, FLAT>(ph, fl2, vp, vR, up, uR, F, press, cmax, mach, nullptr, nullptr);
    if (emf_ok) store_face_emf_p<DIR, NC>(a.e1, a.e2, a.sv, id, F);
    if (PHASE == 1 && emf_ok && a.fbn) a.fbn[id] = F[D::bn];
    if (a.phif && zone_ok) F[ENG] += F[RHO]*__ldg (a.phif + id);      // TotalFlux, rhs.c:388-392
    if (face_ok) my_mach = mach > my_mach ? mach : my_mach;
    if (SOLVER == SOLVER_ROE && face_ok && !ok) atomicAdd (a.red + RED_ROEFAIL, 1ull);

    double Fm[NV];
    PG_FOR_NV(nv) Fm[nv] = __shfl_up_sync (0xffffffffu, F[nv], 1);
    const double pm = __shfl_up_sync (0xffffffffu, press, 1);
    if (rhs_ok){
      const double cd = cmax*d_dl;                             // ctu_step.c:416-419, 634-637: cmax[i] inv_dl[i]
      my_cdt = cd > my_cdt ? cd : my_cdt;
      if (PHASE == 0){
        PG_FOR_NV(nv){
          double rr = -dt2_dx*(F[nv] - Fm[nv]);
          if (nv == D::vn){
            rr -= dt2_dx*(press - pm);
            if (a.bf) rr += dt_2*rho_c*gz;
            if (a.phif) rr -= dt2_dx*rho_c*dphi;
          }
          if (nv == ENG && a.bf) rr += dt_2*0.5*(F[RHO] + Fm[RHO])*gz;
          if (nv == ENG && a.phic) rr -= __ldg (a.phic + id)*(-dt2_dx*(F[RHO] - Fm[RHO]));
          a.rhs[DIR][nv][id] = rr;
        }
      }else{
        bool upd = jr >= 1 && jr <= g.n[1];
        if (NC == 3) upd = upd && kr >= 1 && kr <= g.n[2];
        if (upd){
          double u0[NV], rr;
          prim_to_cons<NC>(ph, v, u0);                           // Uc = PrimToCons (V^n), ctu_step.c:257-262
          rr = -dtdx*(F[RHO] - Fm[RHO]);                              a.U[RHO][id] = u0[RHO] + rr;
          const double dtf = 2.0*dt_2;                                // = dt
          rr = -dtdx*(F[MX1] - Fm[MX1]); rr -= dtdx*(press - pm);
          if (a.bf) rr += dtf*rho_c*gz;
          if (a.phif) rr -= dtdx*rho_c*dphi;
          a.U[MX1][id] = u0[MX1] + rr;
          rr = -dtdx*(F[MX2] - Fm[MX2]);                              a.U[MX2][id] = u0[MX2] + rr;
          if (NC == 3){ rr = -dtdx*(F[MX3] - Fm[MX3]);                a.U[MX3][id] = u0[MX3] + rr; }
          rr = -dtdx*(F[ENG] - Fm[ENG]);
          if (a.bf) rr += dtf*0.5*(F[RHO] + Fm[RHO])*gz;
          if (a.phic) rr -= __ldg (a.phic + id)*(-dtdx*(F[RHO] - Fm[RHO]));
          a.U[ENG][id] = u0[ENG] + rr;
        }
      }
    }
  }
  my_cdt = warp_max (my_cdt);
  my_mach = warp_max (my_mach);
  const unsigned mfl = __ballot_sync (0xffffffffu, nfl_tot > 0);
  int tot_fl = nfl_tot;
  if (mfl) PG_UNROLL for (int o = 16; o > 0; o >>= 1) tot_fl += __shfl_xor_sync (0xffffffffu, tot_fl, o);
  if (lane == 0){
    atomic_max_pos (a.red + RED_CDT, my_cdt);
    atomic_max_pos (a.red + RED_MACH, my_mach);
    if (mfl) atomicAdd (a.red + RED_FLOOR, (unsigned long long)tot_fl);
  }
}

// ---------------------------------------------------------------------------
//  x2 / x3 sweeps: marching pencils.  What the iteration of zone z consumes -- V(z+1), the face
//  field(s) of z, and in the corrector the 2 x 8 transverse right-hand sides of z and U(z-1) --
//  is pulled ONE zone ahead by cp.async into the thread's own shared-memory column (two slots,
//  conflict-free, no destination registers): with ~25 global loads per zone consumed at once the
//  kernel otherwise waits on memory (ncu: long_scoreboard 8.6 stalls per issue at 12 % occupancy).
// ---------------------------------------------------------------------------
__host__ __device__ constexpr int ctu_march_slot (int phase) { return phase == 0 ? 9 : 8 + 2 + 16 + 5; }   // doubles per zone
template <int DIR, int PHASE, int SOLVER, int NC, bool FLAT>
__global__ void __launch_bounds__(128, PG_CTU_MINB_M)
ctu_sweep_march_kernel (const __grid_constant__ CtuArgs a)
{
  typedef Dirs<DIR> D;
  constexpr int E = (PHASE == 0 ? 2 : 1);
  constexpr int TD = (DIR == 1 ? 2 : 1);               // second transverse dimension
  constexpr int CS = 128;                              // = blockDim.x
  constexpr int NQ = ctu_march_slot (PHASE);
  constexpr int Q_BS = 8, Q_BH = 9, Q_RA = 10, Q_RB = 18, Q_U = 26;
  constexpr int DA = 0, DB = (DIR == 1 ? 2 : 1);       // the two transverse directions (2-D: only DA)
  const Geom &g = a.g;
  const Phys &ph = *reinterpret_cast<const Phys *>(&a.ph);
  const int lane = threadIdx.x & 31;
  const int np1 = g.n[0] + 2*E;
  const int np2 = (NC == 3 ? g.n[TD] + 2*E : 1);
  const long long npen = (long long)np1*np2;
  const long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  extern __shared__ double ring_[];
  double my_mach = 0.0, my_cdt = 0.0;
  int nfl = 0;
  if (t < npen*a.nchunk){
    const int chunk = (int)(t/npen);
    const long long p = t - (long long)chunk*npen;
    const int t2 = (int)(p/np1), t1 = (int)(p - (long long)t2*np1);
    const int i  = g.beg[0] - E + t1;
    const int o2 = (NC == 3 ? g.beg[TD] - E + t2 : 0);
    // zones that receive a right-hand side: DOM+-1 along the sweep in the predictor, DOM in the corrector
    const int R0 = g.beg[DIR] - (PHASE == 0 ? 1 : 0), R1 = g.end[DIR] + (PHASE == 0 ? 1 : 0);
    const int c0 = R0 + chunk*a.chunk_len;
    int c1 = c0 + a.chunk_len - 1; if (c1 > R1) c1 = R1;
    bool upd = PHASE == 1 && i >= g.beg[0] && i <= g.end[0];
    if (NC == 3) upd = upd && o2 >= g.beg[TD] && o2 <= g.end[TD];
    const int sD = (DIR == 1 ? (int)g.S1 : (int)g.S12);
    int id = (DIR == 1 ? gidx32 (g, o2, c0 - 1, i) : gidx32 (g, c0 - 1, o2, i));      // zone c0-1

    double *cur = ring_ + threadIdx.x, *nxt = cur + NQ*CS;
    // everything the iteration of zone z (index idz) reads from HBM
    auto fetch = [&] (double *dst, int idz, bool first){
      cp_async8_ordered (dst + Q_BS*CS, a.Bs0[DIR] + idz);          // compiler barrier: the slot has just been read
      PG_FOR_NV(nv) cp_async8 (dst + nv*CS, a.V0[nv] + idz + sD);
      if (PHASE == 1){
        cp_async8 (dst + Q_BH*CS, a.Bsh[DIR] + idz);
        PG_FOR_NV(nv) cp_async8 (dst + (Q_RA + nv)*CS, a.rhs[DA][nv] + idz);
        if (NC == 3) PG_FOR_NV(nv) cp_async8 (dst + (Q_RB + nv)*CS, a.rhs[DB][nv] + idz);
        if (upd && !first){                            // U of zone z-1, updated at the end of the iteration
          cp_async8 (dst + (Q_U + 0)*CS, a.U[RHO] + idz - sD); cp_async8 (dst + (Q_U + 1)*CS, a.U[MX1] + idz - sD);
          cp_async8 (dst + (Q_U + 2)*CS, a.U[MX2] + idz - sD);
          if (NC == 3) cp_async8 (dst + (Q_U + 3)*CS, a.U[MX3] + idz - sD);
          cp_async8 (dst + (Q_U + 4)*CS, a.U[ENG] + idz - sD);
        }
      }
    };
    fetch (cur, id, true);
    cp_async_commit ();

    const double dt_2 = 0.5*__ldg (a.dtp + 3);
    double dtdx = 0.0, dt2_dx = 0.0, d_dlL = 0.0;     // dt/dx, (dt/2)/dx, 1/dx of zone z-1 (the zone that receives the right-hand side)
    double rhoL = 0.0, gL = 0.0, dphiL = 0.0;        // body force: stateC density (predictor), force, potential step of zone z-1
    double vl[NV], v[NV], vr[NV];
    PG_FOR_NV(nv){ v[nv] = __ldg (a.V0[nv] + id - sD); vr[nv] = __ldg (a.V0[nv] + id); }
    double bsp = __ldg (a.Bs0[DIR] + id - sD), bhp = 0.0;
    if (PHASE == 1) bhp = __ldg (a.Bsh[DIR] + id - sD);
    double vpL[NV], upL[NV], Fp[NV], pp = 0.0;
    PG_FOR_NV(nv){ vpL[nv] = 0.0; upL[nv] = 0.0; Fp[nv] = 0.0; }
    unsigned flb = 0;

    for (int z = c0 - 1; z <= c1 + 1; z++, id += sD){
      // id = zone z; its data were requested one iteration ago
      if (z + 1 <= c1 + 1) fetch (nxt, id + sD, false);
      cp_async_commit ();
      cp_async_wait<1> ();
      PG_FOR_NV(nv){ vl[nv] = v[nv]; v[nv] = vr[nv]; vr[nv] = cur[nv*CS]; }
      const double bsm = bsp;
      bsp = cur[Q_BS*CS];
      unsigned flz = 0;
      if (FLAT) flz = a.flag[id];
      double vp[NV], vm[NV], up[NV], um[NV];
      const double gz = a.bf ? (a.gf ? __ldg (a.gf + id) : a.grav[DIR]) : 0.0;    // body force at the centre of zone z
      double src_n = 0.0, dphi = 0.0;                  // PrimSource of the normal velocity (prim_eqn.c:289-307)
      if (a.bf) src_n += gz;
      if (a.phif){ dphi = __ldg (a.phif + id) - __ldg (a.phif + id - sD); src_n -= pg_div (dphi, 1.0*(a.dxz ? __ldg (a.dxz + z) : g.dx[DIR])); }
      const double dtdx_z = __ldg (a.dtx + z*a.gs), d_dl = a.gs ? __ldg (a.idl + z) : a.inv_dl;       // of zone z
      ctu_states<DIR, NC, FLAT>(ph, a.limiter, flz, vl, v, vr, bsm, bsp, dt_2, d_dl, src_n, vp, vm, a.chtr ? dtdx_z : 0.0, a.char_lim);
      const double rho_z = ((a.bf || a.phif) && PHASE == 0 ? 0.5*(vp[RHO] + vm[RHO]) : 0.0);   // hancock.c:136-141
      if (PHASE == 0){
        prim_to_cons<NC>(ph, vp, up);
        prim_to_cons<NC>(ph, vm, um);
      }else{
        double dU[NV];
        PG_FOR_NV(nv){                                   // ctu_step.c:551-562, the reference's order
          const double ra = cur[(Q_RA + nv)*CS];
          if (NC == 3){
            const double rb = cur[(Q_RB + nv)*CS];
            if (DIR == 1) dU[nv] = ra + 0.0 + rb;
            else          dU[nv] = ra + rb;
          }else dU[nv] = ra + 0.0;
        }
        const double bhm = bhp;
        bhp = cur[Q_BH*CS];
        const int n = ctu_correct<DIR, NC>(ph, v, dU, 0.5*dtdx_z, bhm, bhp, vp, vm, up, um);
        if ((z >= c0 || chunk == 0) && (z <= c1 || chunk == a.nchunk - 1)) nfl += n;   // zones shared by two chunks count once
      }
      if (z >= c0){
        // face z-1/2, stored with the index of zone z-1
        const int idf = id - sD;
        double F[NV], press, cmax, mach;
        const bool ok = riemann_f<SOLVER, DIR, NC, FLAT>(ph, flb | flz, vpL, vm, upL, um, F, press, cmax, mach, nullptr, nullptr);
        my_mach = mach > my_mach ? mach : my_mach;
        if (SOLVER == SOLVER_ROE && !ok) atomicAdd (a.red + RED_ROEFAIL, 1ull);
        if (z - 1 >= c0 || chunk == 0) store_face_emf_p<DIR, NC>(a.e1, a.e2, a.sv, idf, F);
        if (a.phif) F[ENG] += F[RHO]*__ldg (a.phif + idf);              // TotalFlux, rhs.c:388-392
        if (PHASE == 1 && a.fbn && (z - 1 >= c0 || chunk == 0)) a.fbn[idf] = F[D::bn];
        if (z - 1 >= c0){
          // zone z-1: faces z-3/2 (Fp) and z-1/2 (F)
          const double cd = cmax*d_dlL;                          // cmax[z-1] inv_dl[z-1]
          my_cdt = cd > my_cdt ? cd : my_cdt;
          if (PHASE == 0){
            PG_FOR_NV(nv){
              double r = -dt2_dx*(F[nv] - Fp[nv]);
              if (nv == D::vn){
                r -= dt2_dx*(press - pp);
                if (a.bf) r += dt_2*rhoL*gL;
                if (a.phif) r -= dt2_dx*rhoL*dphiL;
              }
              if (nv == ENG && a.bf) r += dt_2*0.5*(F[RHO] + Fp[RHO])*gL;
              if (nv == ENG && a.phic) r -= __ldg (a.phic + idf)*(-dt2_dx*(F[RHO] - Fp[RHO]));
              a.rhs[DIR][nv][idf] = r;
            }
          }else if (upd){
            const double *ua = cur + Q_U*CS;
            const double dtf = 2.0*dt_2;                                        // = dt
            const double rho_h = (a.bf || a.phif) ? __ldg (a.Vh[RHO] + idf) : 0.0;   // V^{n+1/2} (ctu_step.c:566-570)
            double r;
            r = -dtdx*(F[RHO] - Fp[RHO]);                                           a.U[RHO][idf] = ua[0] + r;
            r = -dtdx*(F[MX1] - Fp[MX1]); if (D::vn == MX1) r -= dtdx*(press - pp); a.U[MX1][idf] = ua[CS] + r;
            r = -dtdx*(F[MX2] - Fp[MX2]);
            if (D::vn == MX2){ r -= dtdx*(press - pp); if (a.bf) r += dtf*rho_h*gL; if (a.phif) r -= dtdx*rho_h*dphiL; }
            a.U[MX2][idf] = ua[2*CS] + r;
            if (NC == 3){
              r = -dtdx*(F[MX3] - Fp[MX3]);
              if (D::vn == MX3){ r -= dtdx*(press - pp); if (a.bf) r += dtf*rho_h*gL; if (a.phif) r -= dtdx*rho_h*dphiL; }
              a.U[MX3][idf] = ua[3*CS] + r;
            }
            r = -dtdx*(F[ENG] - Fp[ENG]);
            if (a.bf) r += dtf*0.5*(F[RHO] + Fp[RHO])*gL;
            if (a.phic) r -= __ldg (a.phic + idf)*(-dtdx*(F[RHO] - Fp[RHO]));
            a.U[ENG][idf] = ua[4*CS] + r;
          }
        }
        PG_FOR_NV(nv) Fp[nv] = F[nv];
        pp = press;
      }
      PG_FOR_NV(nv){ vpL[nv] = vp[nv]; upL[nv] = up[nv]; }
      rhoL = rho_z; gL = gz; dphiL = dphi;
      dtdx = dtdx_z; dt2_dx = 0.5*dtdx_z; d_dlL = d_dl;
      flb = flz;
      double *tmp = cur; cur = nxt; nxt = tmp;
    }
    cp_async_wait_all ();
  }
  my_cdt = warp_max (my_cdt);
  my_mach = warp_max (my_mach);
  const unsigned mfl = __ballot_sync (0xffffffffu, nfl > 0);
  int tot_fl = nfl;
  if (mfl) PG_UNROLL for (int o = 16; o > 0; o >>= 1) tot_fl += __shfl_xor_sync (0xffffffffu, tot_fl, o);
  if (lane == 0){
    atomic_max_pos (a.red + RED_CDT, my_cdt);
    atomic_max_pos (a.red + RED_MACH, my_mach);
    if (mfl) atomicAdd (a.red + RED_FLOOR, (unsigned long long)tot_fl);
  }
}

// ---------------------------------------------------------------------------
//  V^{n+1/2} on DOM+-1: U^n + sum of the half-step right-hand sides, the cell average of the
//  half-step staggered field, ConsToPrim (ctu_step.c:427-497)
// ---------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(128)
ctu_half_kernel (const __grid_constant__ CtuArgs a)
{
  const Geom &g = a.g;
  const Phys &ph = *reinterpret_cast<const Phys *>(&a.ph);
  const int ni = g.n[0] + 2, nj = g.n[1] + 2, nk = (NC == 3 ? g.n[2] + 2 : 1);
  const unsigned t = blockIdx.x*blockDim.x + threadIdx.x;
  int fl = 0;
  if (t < (unsigned)(ni*nj*nk)){
    const unsigned tq = t/(unsigned)ni;
    const int ti = (int)(t - tq*(unsigned)ni), tj = (int)(tq % (unsigned)nj), tk = (int)(tq/(unsigned)nj);
    const int i = g.beg[0] - 1 + ti, j = g.beg[1] - 1 + tj, k = (NC == 3 ? g.beg[2] - 1 + tk : 0);
    const int id = gidx32 (g, k, j, i);
    double v0[NV], u[NV], v[NV];
    PG_FOR_NV(nv) v0[nv] = __ldg (a.V0[nv] + id);
    prim_to_cons<NC>(ph, v0, u);
    PG_FOR_NV(nv){
      double dU;
      if (NC == 3) dU = a.rhs[0][nv][id] + a.rhs[1][nv][id] + a.rhs[2][nv][id];
      else         dU = a.rhs[0][nv][id] + a.rhs[1][nv][id];
      u[nv] = u[nv] + dU;
    }
    double b2_old;                      // of Uh[B] = U^n[B] + sum of the half-step induction right-hand sides
    if (NC == 3) b2_old = u[BX1]*u[BX1] + u[BX2]*u[BX2] + u[BX3]*u[BX3];
    else         b2_old = u[BX1]*u[BX1] + u[BX2]*u[BX2];
    u[BX1] = 0.5*(a.Bsh[0][id] + a.Bsh[0][id - 1]);
    u[BX2] = 0.5*(a.Bsh[1][id] + a.Bsh[1][id - (int)g.S1]);
    if (NC == 3) u[BX3] = 0.5*(a.Bsh[2][id] + a.Bsh[2][id - (int)g.S12]);
    if (a.en_corr){                     // CT_EN_CORRECTION YES (ct_field_average.c:116-129)
      double b2_new;
      if (NC == 3) b2_new = u[BX1]*u[BX1] + u[BX2]*u[BX2] + u[BX3]*u[BX3];
      else         b2_new = u[BX1]*u[BX1] + u[BX2]*u[BX2];
      u[ENG] += 0.5*(b2_new - b2_old);
    }
    fl = cons_to_prim<NC>(ph, u, v);
    PG_FOR_NV(nv) a.Vh[nv][id] = v[nv];
  }
  const unsigned mfl = __ballot_sync (0xffffffffu, fl);
  if ((threadIdx.x & 31) == 0 && mfl) atomicAdd (a.red + RED_FLOOR, (unsigned long long)__popc (mfl));
}

// ---------------------------------------------------------------------------
//  launchers
// ---------------------------------------------------------------------------
template <int SOLVER>
static int launch_ctu_sweep_t (int dir, int phase, const CtuArgs &a, cudaStream_t s)
{
  const Geom &g = a.g;
  const int nc = g.dims, TPB = 128;
  const int E = (phase == 0 ? 2 : 1);
  const bool fl = a.flag != nullptr;
  if (dir == 0){
    const long long nseg = (g.n[0] + 2*E - 2 + 29)/30;
    const long long nrows = (long long)(g.n[1] + 2*E)*(nc == 3 ? g.n[2] + 2*E : 1);
    const long long nwarp = nseg*((nrows + PG_CTU_XROWS - 1)/PG_CTU_XROWS);
    const unsigned nb = (unsigned)((nwarp*32 + TPB - 1)/TPB);
#define PG_CX1(P, C, F) do { auto kfn = ctu_sweep_x_kernel<P, SOLVER, C, F>;                                  \
      const size_t smem = (size_t)(TPB/32)*2*ctu_x_slot (P)*sizeof (double);                                  \
      static unsigned long long devs = 0;                                                                    \
      if (pg_attr_needed (devs)) cudaFuncSetAttribute (kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      kfn<<<nb, TPB, smem, s>>>(a); } while (0)
#define PG_CX(P, C) do { if (fl) PG_CX1(P, C, true); else PG_CX1(P, C, false); } while (0)
    if (phase == 0){ if (nc == 3) PG_CX(0, 3); else PG_CX(0, 2); }
    else           { if (nc == 3) PG_CX(1, 3); else PG_CX(1, 2); }
#undef PG_CX
#undef PG_CX1
  }else{
    const int td = (dir == 1 ? 2 : 1);
    const long long npen = (long long)(g.n[0] + 2*E)*(nc == 3 ? g.n[td] + 2*E : 1);
    const unsigned nb = (unsigned)((npen*a.nchunk + TPB - 1)/TPB);
#define PG_CM1(DD, P, C, F) do { auto kfn = ctu_sweep_march_kernel<DD, P, SOLVER, C, F>;                    \
      const size_t smem = (size_t)2*ctu_march_slot (P)*TPB*sizeof (double);                                  \
      static unsigned long long devs = 0;                                                                    \
      if (pg_attr_needed (devs)) cudaFuncSetAttribute (kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      kfn<<<nb, TPB, smem, s>>>(a); } while (0)
#define PG_CM(DD, P, C) do { if (fl) PG_CM1(DD, P, C, true); else PG_CM1(DD, P, C, false); } while (0)
    if (dir == 1){
      if (phase == 0){ if (nc == 3) PG_CM(1, 0, 3); else PG_CM(1, 0, 2); }
      else           { if (nc == 3) PG_CM(1, 1, 3); else PG_CM(1, 1, 2); }
    }else{
      if (phase == 0) PG_CM(2, 0, 3); else PG_CM(2, 1, 3);
    }
#undef PG_CM
#undef PG_CM1
  }
  return pg_launch_status ();
}

static int launch_ctu_half_t (const CtuArgs &a, cudaStream_t s)
{
  const Geom &g = a.g;
  const long long n = (long long)(g.n[0] + 2)*(g.n[1] + 2)*(g.dims == 3 ? g.n[2] + 2 : 1);
  const unsigned nb = (unsigned)((n + 127)/128);
  if (g.dims == 3) ctu_half_kernel<3><<<nb, 128, 0, s>>>(a);
  else             ctu_half_kernel<2><<<nb, 128, 0, s>>>(a);
  return pg_launch_status ();
}

} // namespace PG_NS
