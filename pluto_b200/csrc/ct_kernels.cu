// ct_kernels.cu -- constrained transport, stage completion, boundary fills and
// halo pack/unpack.  Compiled once per arithmetic namespace (PG_NS).
//
//   ct_emf_kernel     CT_ComputeCenterEMF + CT_EMF_ArithmeticAverage +
//                     CT_EMF_IntegrateToCorner + x0.25 in GATHER form: one
//                     thread per edge, no scatter-add, no cell-centred EMF
//                     array (reference Src/MHD/CT/ct_emf.c:210-254, 348-388,
//                     ct_emf_average.c:13-49, 52-164).  The additions are
//                     issued in the order the reference's (k,j,i) scatter
//                     loop applies them to each edge, so the result is
//                     bit-identical.
//   ct_update_kernel  CT_Update (ct_update.c:79-218) fused with the RK stage
//                     average of the staggered field (rk_step.c:172-174,
//                     231-233).
//   final_kernel      CT_AverageMagneticField (ct_field_average.c:58-124),
//                     RK stage average of U (rk_step.c:168-171, 226-229),
//                     ConsToPrim (mappers.c:88-254) and the floor/NaN counts.
//   bc_kernel / bc_fill_kernel
//                     Boundary (boundary.c:137-293), FillMagneticField
//                     (ct_fill_mag_field.c:38-179), CT_AverageNormalMagField
//                     (ct_field_average.c:134-272).
#include "kernels_common.cuh"
#include "mhd_device.cuh"

namespace PG_NS {

static inline unsigned nblocks (long long n, int tpb) { return (unsigned)((n + tpb - 1)/tpb); }

// ---------------------------------------------------------------------------
//  edge EMFs
// ---------------------------------------------------------------------------
struct CellE { double e1, e2, e3; };

template <int NC>
__device__ __forceinline__ double cell_e3 (const CtArgs &a, long long id)
{
  return a.V[VX2][id]*a.V[BX1][id] - a.V[VX1][id]*a.V[BX2][id];
}
__device__ __forceinline__ double cell_e1 (const CtArgs &a, long long id)
{
  return a.V[VX3][id]*a.V[BX2][id] - a.V[VX2][id]*a.V[BX3][id];
}
__device__ __forceinline__ double cell_e2 (const CtArgs &a, long long id)
{
  return a.V[VX1][id]*a.V[BX3][id] - a.V[VX3][id]*a.V[BX1][id];
}

// upwind selection of a (face - centre) difference pair: s > 0 -> first,
// s < 0 -> second, s == 0 -> mean (ct_emf_average.c:110-162)
__device__ __forceinline__ double upw (signed char s, double d0, double d1)
{
  if (s == 0) return 0.5*(d0 + d1);
  return s > 0 ? d0 : d1;
}

#ifndef PG_CT_MINB
#define PG_CT_MINB 8
#endif
template <int NC, int AVG, bool EC = false>   // AVG: PLUTO_GPU_EMF_* (compile time: the averages share no code); EC: the cell-centred
                                              // EMFs come from the arrays the fused x1+x2 sweep stored (CtArgs.Ec)
__global__ void __launch_bounds__(128, PG_CT_MINB)
ct_emf_kernel (const __grid_constant__ CtArgs a)
{
  const Geom &g = a.g;
  // edges (k,j,i) with i in [IBEG-1,IEND], j in [JBEG-1,JEND], k in [KBEG-1,KEND]
  const int ext = a.ext;
  const int ni = g.n[0] + 1 + 2*ext, nj = g.n[1] + 1 + 2*ext, nk = (NC == 3 ? g.n[2] + 1 + 2*ext : 1);
  const unsigned t = blockIdx.x*blockDim.x + threadIdx.x;          // 32-bit: < 2^31 zones per block
  if (t >= (unsigned)(ni*nj*nk)) return;
  const unsigned tq = t/(unsigned)ni;
  const int ti = (int)(t - tq*(unsigned)ni), tj = (int)(tq % (unsigned)nj), tk = (int)(tq/(unsigned)nj);
  const int i = g.beg[0] - 1 - ext + ti, j = g.beg[1] - 1 - ext + tj, k = (NC == 3 ? g.beg[2] - 1 - ext + tk : 0);
  const long long id = gidx (g, k, j, i);
  const long long sx = 1, sy = g.S1, sz = g.S12;

  if (AVG == 3){
    // UCT_HLL: CT_GetStagSlopes (ct_stag_slopes.c:52-95) + CT_EMF_HLL_Solver (ct_emf_average.c:180-359,
    // Londrillo & Del Zanna 2004, eq. 56) in the reference's operation order.  The face arrays hold
    // the fan speeds max(0,-SL) [e1 of the direction] and max(0,SR) [e2]; the staggered slopes are
    // evaluated where they are used.
    const long long st[3] = {sx, sy, sz};
    auto mc2 = [] (double dp, double dm){
      if (dp*dm < 0.0) return 0.0;
      const double dc = 0.5*(dp + dm), scrh = 2.0*(fabs (dp) < fabs (dm) ? dp : dm);
      return fabs (dc) < fabs (scrh) ? dc : scrh;
    };
    auto DB = [&] (int c, int d, long long q){ return mc2 (a.Bs_in[c][q + st[d]] - a.Bs_in[c][q], a.Bs_in[c][q] - a.Bs_in[c][q - st[d]]); };
    auto BP = [&] (int c, int d, long long q){ return a.Bs_in[c][q] + 0.5*DB (c, d, q); };
    auto BM = [&] (int c, int d, long long q){ return a.Bs_in[c][q] - 0.5*DB (c, d, q); };
    auto VPP = [&] (int c, int x, int y, long long q){ return a.V[VX1 + c][q] + 0.5*(a.dvel[c][x][q] + a.dvel[c][y][q]); };
    auto VPM = [&] (int c, int x, int y, long long q){ return a.V[VX1 + c][q] + 0.5*(a.dvel[c][x][q] - a.dvel[c][y][q]); };
    auto VMM = [&] (int c, int x, int y, long long q){ return a.V[VX1 + c][q] - 0.5*(a.dvel[c][x][q] + a.dvel[c][y][q]); };
    auto VMP = [&] (int c, int x, int y, long long q){ return a.V[VX1 + c][q] - 0.5*(a.dvel[c][x][q] - a.dvel[c][y][q]); };
    auto hll2d = [] (double a_xp, double a_xm, double a_yp, double a_ym, double eSW, double eSE, double eNE, double eNW,
                     double bS, double bN, double bW, double bE){
      double e = a_xp*a_yp*eSW + a_xm*a_yp*eSE + a_xm*a_ym*eNE + a_xp*a_ym*eNW;
      e = pg_div (e, (a_xp + a_xm)*(a_yp + a_ym));
      e -= pg_div (a_yp*a_ym*(bN - bS), a_yp + a_ym);
      e += pg_div (a_xp*a_xm*(bE - bW), a_xp + a_xm);
      return e;
    };
    const double *SL[3] = {a.ezi, a.ezj, a.eyk}, *SR[3] = {a.eyi, a.exj, a.exk};
    const long long ip = id + sx, jp = id + sy, kp = id + sz;
    {
      const double a_xp = maxv (SR[0][id], SR[0][jp]), a_xm = maxv (SL[0][id], SL[0][jp]);
      const double a_yp = maxv (SR[1][id], SR[1][ip]), a_ym = maxv (SL[1][id], SL[1][ip]);
      const double bS = BP (0, 1, id), bW = BP (1, 0, id), bN = BM (0, 1, jp), bE = BM (1, 0, ip);
      const double eSW = VPP (1, 0, 1, id)*bS - VPP (0, 0, 1, id)*bW;
      const double eSE = VMP (1, 0, 1, ip)*bS - VMP (0, 0, 1, ip)*bE;
      const double eNE = VMM (1, 0, 1, ip + sy)*bN - VMM (0, 0, 1, ip + sy)*bE;
      const double eNW = VPM (1, 0, 1, jp)*bN - VPM (0, 0, 1, jp)*bW;
      a.ez[id] = hll2d (a_xp, a_xm, a_yp, a_ym, eSW, eSE, eNE, eNW, bS, bN, bW, bE);
    }
    if (NC == 3){
      {
        const double a_xp = maxv (SR[1][id], SR[1][kp]), a_xm = maxv (SL[1][id], SL[1][kp]);
        const double a_yp = maxv (SR[2][id], SR[2][jp]), a_ym = maxv (SL[2][id], SL[2][jp]);
        const double bS = BP (1, 2, id), bW = BP (2, 1, id), bN = BM (1, 2, kp), bE = BM (2, 1, jp);
        const double eSW = VPP (2, 1, 2, id)*bS - VPP (1, 1, 2, id)*bW;
        const double eSE = VMP (2, 1, 2, jp)*bS - VMP (1, 1, 2, jp)*bE;
        const double eNE = VMM (2, 1, 2, jp + sz)*bN - VMM (1, 1, 2, jp + sz)*bE;
        const double eNW = VPM (2, 1, 2, kp)*bN - VPM (1, 1, 2, kp)*bW;
        a.ex[id] = hll2d (a_xp, a_xm, a_yp, a_ym, eSW, eSE, eNE, eNW, bS, bN, bW, bE);
      }
      {
        const double a_xp = maxv (SR[2][id], SR[2][ip]), a_xm = maxv (SL[2][id], SL[2][ip]);
        const double a_yp = maxv (SR[0][id], SR[0][kp]), a_ym = maxv (SL[0][id], SL[0][kp]);
        const double bS = BP (2, 0, id), bW = BP (0, 2, id), bN = BM (2, 0, ip), bE = BM (0, 2, kp);
        const double eSW = VPP (0, 2, 0, id)*bS - VPP (2, 2, 0, id)*bW;
        const double eSE = VMP (0, 2, 0, kp)*bS - VMP (2, 2, 0, kp)*bE;
        const double eNE = VMM (0, 2, 0, ip + sz)*bN - VMM (2, 2, 0, ip + sz)*bE;
        const double eNW = VPM (0, 2, 0, ip)*bN - VPM (2, 2, 0, ip)*bW;
        a.ey[id] = hll2d (a_xp, a_xm, a_yp, a_ym, eSW, eSE, eNE, eNW, bS, bN, bW, bE);
      }
    }
    return;
  }
  if (AVG != 0){
    // ARITHMETIC: CT_EMF_ArithmeticAverage (emf, 0.25) (ct_emf.c:241-243).  UCT0: the face EMFs
    // are first replaced by 2 face - mean of the two adjacent cell-centred EMFs (ct_emf.c:261-283);
    // evaluated here per use instead of in place.
    constexpr bool u0 = (AVG == 2);
    auto fz_i = [&] (long long q){ double e = a.ezi[q]; if (u0){ e *= 2.0; e -= 0.5*(cell_e3<NC>(a, q) + cell_e3<NC>(a, q + sx)); } return e; };
    auto fz_j = [&] (long long q){ double e = a.ezj[q]; if (u0){ e *= 2.0; e -= 0.5*(cell_e3<NC>(a, q) + cell_e3<NC>(a, q + sy)); } return e; };
    a.ez[id] = 0.25*(fz_i (id) + fz_i (id + sy) + fz_j (id) + fz_j (id + sx));
    if (NC == 3){
      auto fx_j = [&] (long long q){ double e = a.exj[q]; if (u0){ e *= 2.0; e -= 0.5*(cell_e1 (a, q) + cell_e1 (a, q + sy)); } return e; };
      auto fx_k = [&] (long long q){ double e = a.exk[q]; if (u0){ e *= 2.0; e -= 0.5*(cell_e1 (a, q) + cell_e1 (a, q + sz)); } return e; };
      auto fy_i = [&] (long long q){ double e = a.eyi[q]; if (u0){ e *= 2.0; e -= 0.5*(cell_e2 (a, q) + cell_e2 (a, q + sx)); } return e; };
      auto fy_k = [&] (long long q){ double e = a.eyk[q]; if (u0){ e *= 2.0; e -= 0.5*(cell_e2 (a, q) + cell_e2 (a, q + sz)); } return e; };
      a.ex[id] = 0.25*(fx_k (id) + fx_k (id + sy) + fx_j (id) + fx_j (id + sz));
      a.ey[id] = 0.25*(fy_i (id) + fy_i (id + sz) + fy_k (id) + fy_k (id + sx));
    }
    return;
  }

#define CE1(q) (EC ? a.Ec[0][q] : cell_e1 (a, q))
#define CE2(q) (EC ? a.Ec[1][q] : cell_e2 (a, q))
#define CE3(q) (EC ? a.Ec[2][q] : cell_e3<NC>(a, q))
  {   // ---- ez at (i+1/2, j+1/2) ----
    const double ezi0 = a.ezi[id], ezi1 = a.ezi[id + sy];
    const double ezj0 = a.ezj[id], ezj1 = a.ezj[id + sx];
    const double E00 = CE3(id),      E10 = CE3(id + sx);
    const double E01 = CE3(id + sy), E11 = CE3(id + sx + sy);
    double e = ezi0 + ezi1 + ezj0 + ezj1;
    e += upw (a.svx[id],      ezj0 - E00, ezj1 - E10);          // DEZ_DYP(j, i | i+1)
    e += upw (a.svy[id],      ezi0 - E00, ezi1 - E01);          // DEZ_DXP(j | j+1, i)
    e -= upw (a.svy[id + sx], E10 - ezi0, E11 - ezi1);          // DEZ_DXM(j | j+1, i+1)
    e -= upw (a.svx[id + sy], E01 - ezj0, E11 - ezj1);          // DEZ_DYM(j+1, i | i+1)
    a.ez[id] = e*0.25;
  }
  if (NC == 3){
    {   // ---- ex at (j+1/2, k+1/2) ----
      const double exk0 = a.exk[id], exk1 = a.exk[id + sy];
      const double exj0 = a.exj[id], exj1 = a.exj[id + sz];
      const double E00 = CE1(id),      E10 = CE1(id + sy);
      const double E01 = CE1(id + sz), E11 = CE1(id + sy + sz);
      double e = exk0 + exk1 + exj0 + exj1;
      e += upw (a.svy[id],      exk0 - E00, exk1 - E10);        // DEX_DZP(k, j | j+1)
      e += upw (a.svz[id],      exj0 - E00, exj1 - E01);        // DEX_DYP(k | k+1, j)
      e -= upw (a.svz[id + sy], E10 - exj0, E11 - exj1);        // DEX_DYM(k | k+1, j+1)
      e -= upw (a.svy[id + sz], E01 - exk0, E11 - exk1);        // DEX_DZM(k+1, j | j+1)
      a.ex[id] = e*0.25;
    }
    {   // ---- ey at (i+1/2, k+1/2) ----
      const double eyi0 = a.eyi[id], eyi1 = a.eyi[id + sz];
      const double eyk0 = a.eyk[id], eyk1 = a.eyk[id + sx];
      const double E00 = CE2(id),      E10 = CE2(id + sx);
      const double E01 = CE2(id + sz), E11 = CE2(id + sx + sz);
      double e = eyi0 + eyi1 + eyk0 + eyk1;
      e += upw (a.svx[id],      eyk0 - E00, eyk1 - E10);        // DEY_DZP(k, i | i+1)
      e += upw (a.svz[id],      eyi0 - E00, eyi1 - E01);        // DEY_DXP(k | k+1, i)
      e -= upw (a.svz[id + sx], E10 - eyi0, E11 - eyi1);        // DEY_DXM(k | k+1, i+1)
      e -= upw (a.svx[id + sz], E01 - eyk0, E11 - eyk1);        // DEY_DZM(k+1, i | i+1)
      a.ey[id] = e*0.25;
    }
  }
}
#undef CE1
#undef CE2
#undef CE3

// ---------------------------------------------------------------------------
//  staggered update + RK average
// ---------------------------------------------------------------------------
__device__ __forceinline__ double stage_mix (int combine, double w0, double wc, double b0, double b)
{
  if (combine == 1) return w0*b0 + wc*b;
  if (combine == 2) return (b0 + 2.0*b)/3.0;
  return b;
}

template <int NC>
__global__ void __launch_bounds__(128, 16)
ct_update_kernel (const __grid_constant__ CtArgs a)
{
  const Geom &g = a.g;
  const int ext = a.ext;
  const int ni = g.n[0] + 1 + 2*ext, nj = g.n[1] + 1 + 2*ext, nk = (NC == 3 ? g.n[2] + 1 + 2*ext : 1);
  const unsigned t = blockIdx.x*blockDim.x + threadIdx.x;          // 32-bit: < 2^31 zones per block
  if (t >= (unsigned)(ni*nj*nk)) return;
  const unsigned tq = t/(unsigned)ni;
  const int ti = (int)(t - tq*(unsigned)ni), tj = (int)(tq % (unsigned)nj), tk = (int)(tq/(unsigned)nj);
  const int i = g.beg[0] - 1 - ext + ti, j = g.beg[1] - 1 - ext + tj, k = (NC == 3 ? g.beg[2] - 1 - ext + tk : 0);
  const long long id = gidx (g, k, j, i);
  const long long sy = g.S1, sz = g.S12;
  const bool in_i = i >= g.beg[0] - ext, in_j = j >= g.beg[1] - ext, in_k = (NC == 3 ? k >= g.beg[2] - ext : true);
  // dt/dx2[j], dt/dx3[k], ... of the zone the face belongs to (ct_update.c:91-96, 147-152, 202-204); uniform grid: gs = 0
  const double dtdx0 = a.dts*__ldg (a.dtx[0] + i*a.gs), dtdx1 = a.dts*__ldg (a.dtx[1] + j*a.gs);
  const double dtdx2 = (NC == 3 ? a.dts*__ldg (a.dtx[2] + k*a.gs) : 0.0);

  if (in_j && in_k){        // Bx1 at (i+1/2, j, k), i in [IBEG-1, IEND]
    double rhs;
    if (NC == 3) rhs = 0.0 - dtdx1*(a.ez[id] - a.ez[id - sy]) + dtdx2*(a.ey[id] - a.ey[id - sz]);
    else         rhs = 0.0 - dtdx1*(a.ez[id] - a.ez[id - sy]);
    double b = a.Bs_in[0][id] + rhs;
    if (a.combine) b = stage_mix (a.combine, a.w0, a.wc, a.Bs0[0][id], b);
    a.Bs_out[0][id] = b;
  }
  if (in_i && in_k){        // Bx2 at (i, j+1/2, k)
    double rhs;
    if (NC == 3) rhs = dtdx0*(a.ez[id] - a.ez[id - 1]) - dtdx2*(a.ex[id] - a.ex[id - sz]);
    else         rhs = dtdx0*(a.ez[id] - a.ez[id - 1]);
    double b = a.Bs_in[1][id] + rhs;
    if (a.combine) b = stage_mix (a.combine, a.w0, a.wc, a.Bs0[1][id], b);
    a.Bs_out[1][id] = b;
  }
  if (NC == 3 && in_i && in_j){   // Bx3 at (i, j, k+1/2)
    double rhs = - dtdx0*(a.ey[id] - a.ey[id - 1]) + dtdx1*(a.ex[id] - a.ex[id - sy]);
    double b = a.Bs_in[2][id] + rhs;
    if (a.combine) b = stage_mix (a.combine, a.w0, a.wc, a.Bs0[2][id], b);
    a.Bs_out[2][id] = b;
  }
}

// ---------------------------------------------------------------------------
//  stage completion: face -> centre average, RK average, cons -> prim
// ---------------------------------------------------------------------------
// the new staggered field of ONE face, exactly as ct_update_kernel forms it (ct_update.c:79-218 + rk_step.c:172-174)
template <int NC, int D>
__device__ __forceinline__ double face_update (const FinalArgs &a, long long id, long long sy, long long sz,
                                               double dtdx0, double dtdx1, double dtdx2)
{
  double rhs;
  if (D == 0){
    if (NC == 3) rhs = 0.0 - dtdx1*(a.ez[id] - a.ez[id - sy]) + dtdx2*(a.ey[id] - a.ey[id - sz]);
    else         rhs = 0.0 - dtdx1*(a.ez[id] - a.ez[id - sy]);
  }else if (D == 1){
    if (NC == 3) rhs = dtdx0*(a.ez[id] - a.ez[id - 1]) - dtdx2*(a.ex[id] - a.ex[id - sz]);
    else         rhs = dtdx0*(a.ez[id] - a.ez[id - 1]);
  }else{
    rhs = - dtdx0*(a.ey[id] - a.ey[id - 1]) + dtdx1*(a.ex[id] - a.ex[id - sy]);
  }
  double b = a.Bs_in[D][id] + rhs;
  if (a.combine) b = stage_mix (a.combine, a.w0, a.wc, a.Bs0[D][id], b);
  return b;
}

template <int NC, bool EN, bool FUSE = false, bool R3 = false>   // EN: CT_EN_CORRECTION YES (compile time: the default instantiation
                                                // keeps its 32 registers); FUSE: CT_Update evaluated here (FinalArgs.fuse_ct);
                                                // R3: the x3 sweep kept its flux difference apart (FinalArgs.R3): u = U + R3
__global__ void __launch_bounds__(128)
final_kernel (const __grid_constant__ FinalArgs a)
{
  const Geom &g = a.g;
  Phys ph; ph.gamma = a.ph.gamma; ph.gmm1 = a.ph.gmm1; ph.small_dn = a.ph.small_dn; ph.small_pr = a.ph.small_pr; ph.igmm1 = a.ph.igmm1;
  const int *blo = a.nbox ? a.boxes_lo[blockIdx.y] : a.box_lo, *bn = a.nbox ? a.boxes_n[blockIdx.y] : a.box_n;
  const int ni = bn[0], nj = bn[1], nk = (NC == 3 ? bn[2] : 1);
  const unsigned t = blockIdx.x*blockDim.x + threadIdx.x;
  int fl = 0, bad = 0;
  if (t < (unsigned)(ni*nj*nk)){
    const unsigned tq = t/(unsigned)ni;
    const int ti = (int)(t - tq*(unsigned)ni), tj = (int)(tq % (unsigned)nj), tk = (int)(tq/(unsigned)nj);
    const int i = g.beg[0] + blo[0] + ti, j = g.beg[1] + blo[1] + tj, k = (NC == 3 ? g.beg[2] + blo[2] + tk : 0);
    const long long id = gidx (g, k, j, i);
    double u[NV], v[NV];
    u[RHO] = a.U[RHO][id]; u[MX1] = a.U[MX1][id]; u[MX2] = a.U[MX2][id];
    if (NC == 3) u[MX3] = a.U[MX3][id];
    u[ENG] = a.U[ENG][id];
    if (R3){                 // U = ((U0 + rhs_x1) + rhs_x2) + rhs_x3, the sum of update_stage.c:214-216 completed here
      u[RHO] = u[RHO] + a.R3[RHO][id]; u[MX1] = u[MX1] + a.R3[MX1][id]; u[MX2] = u[MX2] + a.R3[MX2][id];
      if (NC == 3) u[MX3] = u[MX3] + a.R3[MX3][id];
      u[ENG] = u[ENG] + a.R3[ENG][id];
    }
    if (a.combine){
      double v0[NV], u0[NV];
      PG_FOR_NV(nv) v0[nv] = a.V0[nv][id];
      prim_to_cons<NC>(ph, v0, u0);
      if (a.combine == 1){
        u[RHO] = a.w0*u0[RHO] + a.wc*u[RHO];
        u[MX1] = a.w0*u0[MX1] + a.wc*u[MX1];
        u[MX2] = a.w0*u0[MX2] + a.wc*u[MX2];
        if (NC == 3) u[MX3] = a.w0*u0[MX3] + a.wc*u[MX3];
        u[ENG] = a.w0*u0[ENG] + a.wc*u[ENG];
      }else{
        const double one_third = 1.0/3.0;
        u[RHO] = one_third*(u0[RHO] + 2.0*u[RHO]);
        u[MX1] = one_third*(u0[MX1] + 2.0*u[MX1]);
        u[MX2] = one_third*(u0[MX2] + 2.0*u[MX2]);
        if (NC == 3) u[MX3] = one_third*(u0[MX3] + 2.0*u[MX3]);
        u[ENG] = one_third*(u0[ENG] + 2.0*u[ENG]);
      }
    }
    double b2_old = 0.0;
    if (EN){
      // Uc[B] as the reference's sweeps leave it: U = ((U_in + rhs_x1) + rhs_x2) + rhs_x3, rhs = -dt/dx (F_i - F_{i-1})
      // (update_stage.c:214-216, rhs.c:193-201) with the induction fluxes F of the stored face EMFs
      // (ct_emf.c:132-176: x1 faces F[BX2] = -ezi, F[BX3] = eyi; x2: F[BX1] = ezj, F[BX3] = -exj; x3: F[BX1] = -eyk,
      // F[BX2] = exk; the flux of the normal component is zero), then the RK average with U0[B] = V0[B]
      const long long sy = g.S1, sz = g.S12;
      const double dtdx0 = __ldg (a.dtx[0] + i*a.gs), dtdx1 = __ldg (a.dtx[1] + j*a.gs), dtdx2 = (NC == 3 ? __ldg (a.dtx[2] + k*a.gs) : 0.0);
      double b1 = a.Vin[BX1][id], b2 = a.Vin[BX2][id], b3 = (NC == 3 ? a.Vin[BX3][id] : 0.0);
      b1 = b1 + -dtdx0*(a.fbn[0] ? a.fbn[0][id] - a.fbn[0][id - 1] : 0.0 - 0.0);
      b2 = b2 + -dtdx0*((-a.ezi[id]) - (-a.ezi[id - 1]));
      if (NC == 3) b3 = b3 + -dtdx0*(a.eyi[id] - a.eyi[id - 1]);
      b1 = b1 + -dtdx1*(a.ezj[id] - a.ezj[id - sy]);
      b2 = b2 + -dtdx1*(a.fbn[1] ? a.fbn[1][id] - a.fbn[1][id - sy] : 0.0 - 0.0);
      if (NC == 3) b3 = b3 + -dtdx1*((-a.exj[id]) - (-a.exj[id - sy]));
      if (NC == 3){
        b1 = b1 + -dtdx2*((-a.eyk[id]) - (-a.eyk[id - sz]));
        b2 = b2 + -dtdx2*(a.exk[id] - a.exk[id - sz]);
        b3 = b3 + -dtdx2*(a.fbn[2] ? a.fbn[2][id] - a.fbn[2][id - sz] : 0.0 - 0.0);
      }
      if (a.combine == 1){
        b1 = a.w0*a.V0[BX1][id] + a.wc*b1; b2 = a.w0*a.V0[BX2][id] + a.wc*b2;
        if (NC == 3) b3 = a.w0*a.V0[BX3][id] + a.wc*b3;
      }else if (a.combine == 2){
        const double one_third = 1.0/3.0;
        b1 = one_third*(a.V0[BX1][id] + 2.0*b1); b2 = one_third*(a.V0[BX2][id] + 2.0*b2);
        if (NC == 3) b3 = one_third*(a.V0[BX3][id] + 2.0*b3);
      }
      if (NC == 3) b2_old = b1*b1 + b2*b2 + b3*b3;
      else         b2_old = b1*b1 + b2*b2;
    }
    if (FUSE){
      const long long sy = g.S1, sz = g.S12;
      const double dtdx0 = __ldg (a.dtp), dtdx1 = __ldg (a.dtp + 1), dtdx2 = (NC == 3 ? __ldg (a.dtp + 2) : 0.0);
      const double b1p = face_update<NC, 0>(a, id, sy, sz, dtdx0, dtdx1, dtdx2), b1m = face_update<NC, 0>(a, id - 1, sy, sz, dtdx0, dtdx1, dtdx2);
      const double b2p = face_update<NC, 1>(a, id, sy, sz, dtdx0, dtdx1, dtdx2), b2m = face_update<NC, 1>(a, id - sy, sy, sz, dtdx0, dtdx1, dtdx2);
      a.Bs_out[0][id] = b1p; a.Bs_out[1][id] = b2p;
      if (i == g.beg[0]) a.Bs_out[0][id - 1] = b1m;          // the face beg-1 has no zone of its own
      if (j == g.beg[1]) a.Bs_out[1][id - sy] = b2m;
      u[BX1] = 0.5*(b1p + b1m);
      u[BX2] = 0.5*(b2p + b2m);
      if (NC == 3){
        const double b3p = face_update<NC, 2>(a, id, sy, sz, dtdx0, dtdx1, dtdx2), b3m = face_update<NC, 2>(a, id - sz, sy, sz, dtdx0, dtdx1, dtdx2);
        a.Bs_out[2][id] = b3p;
        if (k == g.beg[2]) a.Bs_out[2][id - sz] = b3m;
        u[BX3] = 0.5*(b3p + b3m);
      }
    }else{
      u[BX1] = 0.5*(a.Bs[0][id] + a.Bs[0][id - 1]);
      u[BX2] = 0.5*(a.Bs[1][id] + a.Bs[1][id - g.S1]);
      if (NC == 3) u[BX3] = 0.5*(a.Bs[2][id] + a.Bs[2][id - g.S12]);
    }
    if (EN){
      double b2_new;
      if (NC == 3) b2_new = u[BX1]*u[BX1] + u[BX2]*u[BX2] + u[BX3]*u[BX3];
      else         b2_new = u[BX1]*u[BX1] + u[BX2]*u[BX2];
      u[ENG] += 0.5*(b2_new - b2_old);
    }
    fl = cons_to_prim<NC>(ph, u, v);
    if (a.write_u == 1 || (a.write_u == 2 && fl)){
      // the reference keeps Uc across stages (rk_step.c:149-186 has no PrimToCons3D)
      a.Uw[RHO][id] = u[RHO]; a.Uw[MX1][id] = u[MX1]; a.Uw[MX2][id] = u[MX2];
      if (NC == 3) a.Uw[MX3][id] = u[MX3];
      a.Uw[ENG][id] = u[ENG];
    }
    PG_FOR_NV(nv){
      a.Vout[nv][id] = v[nv];
      if (!(fabs(v[nv]) <= 1.7976931348623157e308)) bad = 1;
    }
  }
  // block-level counts
  unsigned mfl = __ballot_sync (0xffffffffu, fl), mbad = __ballot_sync (0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0){
    if (mfl)  atomicAdd (a.red + RED_FLOOR, (unsigned long long)__popc (mfl));
    if (mbad) atomicAdd (a.red + RED_NAN,   (unsigned long long)__popc (mbad));
  }
}

// ---------------------------------------------------------------------------
//  boundary fills: one launch per dimension.  blockIdx.y < nf: copy job (one field,
//  one side); otherwise the div B = 0 fill of one side.  A fill does not read the
//  ghost values the copy jobs of the same launch write: it takes the tangential
//  components through the same source mapping (bc_source), so the two are
//  independent.
// ---------------------------------------------------------------------------
// source zone along d of ghost index n for a boundary of the given type
__device__ __forceinline__ int bc_source (const Geom &g, int type, int d, int hi_side, int n)
{
  if (type == 0) return n + (hi_side ? -g.n[d] : g.n[d]);            // periodic  (boundary.c:480-518)
  if (type == 1) return hi_side ? g.end[d] : g.beg[d];               // outflow   (boundary.c:439-477)
  return hi_side ? 2*g.end[d] - n + 1 : 2*g.beg[d] - n - 1;          // reflective / eqtsymmetric (boundary.c:521-564)
}

__global__ void __launch_bounds__(128)
bc_kernel (const __grid_constant__ BcArgs a)
{
  const Geom &g = a.g;
  if ((int)blockIdx.y < a.nf){
    const BcField &f = a.f[blockIdx.y];
    // 32-bit index arithmetic (a ghost slab has far fewer than 2^31 zones): the three divisions
    // per thread dominate this kernel otherwise
    const int ni = f.hi[0] - f.lo[0] + 1, nj = f.hi[1] - f.lo[1] + 1, nk = f.hi[2] - f.lo[2] + 1;
    const unsigned t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= (unsigned)(ni*nj*nk)) return;
    const unsigned tj = t/(unsigned)ni;
    const int i = f.lo[0] + (int)(t - tj*(unsigned)ni), j = f.lo[1] + (int)(tj % (unsigned)nj), k = f.lo[2] + (int)(tj/(unsigned)nj);
    int c[3] = {i, j, k};
    const int d = f.side >> 1, hi_side = f.side & 1;
    c[d] = bc_source (g, f.type, d, hi_side, c[d]);
    const double x = f.q[gidx (g, c[2], c[1], c[0])];
    f.q[gidx (g, k, j, i)] = (f.type == 2 || f.type == 4 ? (double)f.sign*x : x);
    return;
  }
  // normal staggered component in the ghost zones from div B = 0, marching
  // outwards (sequential along the normal), then the cell-centred normal
  // component as the face average.  One thread per transverse position.
  const BcFill &fl = a.fill[blockIdx.y - a.nf];
  const int d = fl.side >> 1, hi_side = fl.side & 1;
  const int d1 = (d == 0 ? 1 : 0), d2 = (d == 2 ? 1 : 2);     // transverse dims, d1 faster
  const int n1 = g.T[d1], n2 = g.T[d2];
  long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= (long long)n1*n2) return;
  int c[3], cs[3];
  c[d1] = (int)(t % n1); c[d2] = (int)(t / n1);
  double A[3];
  const long long st[3] = {1, g.S1, g.S12};
  const int nbeg = hi_side ? g.end[d] + 1 : g.beg[d] - 1;
  const int nend = hi_side ? g.T[d] - 1 : 0;
  const int dn = hi_side ? 1 : -1;
  for (int n = nbeg; dn*n <= dn*nend; n += dn){
    c[d] = n;
    const long long id = gidx (g, c[2], c[1], c[0]);
    {   // face areas of zone (c[2], c[1], c[0]) (set_geometry.c:149,180,202): on a non-uniform grid they change along the march
      const double dx1 = a.dxa[0] ? a.dxa[0][c[0]] : g.dx[0], dx2 = a.dxa[1] ? a.dxa[1][c[1]] : g.dx[1];
      const double dx3 = (g.dims == 3 ? (a.dxa[2] ? a.dxa[2][c[2]] : g.dx[2]) : 1.0);
      if (g.dims == 3){ A[0] = 1.0*dx2*dx3; A[1] = dx1*1.0*dx3; A[2] = dx1*dx2*1.0; }
      else            { A[0] = 1.0*dx2;     A[1] = dx1*1.0;     A[2] = 0.0; }
    }
    // tangential components of ghost zone n = those of its source zone (what the copy
    // jobs of this side store there; reflective: tangential fields keep their sign)
    cs[0] = c[0]; cs[1] = c[1]; cs[2] = c[2];
    cs[d] = bc_source (g, fl.type, d, hi_side, n);
    const long long ids = gidx (g, cs[2], cs[1], cs[0]);
    double dB[3] = {0.0, 0.0, 0.0};
    double bp[3] = {0.0, 0.0, 0.0}, bm[3] = {0.0, 0.0, 0.0};
    const double tsign = (fl.type == 4 ? -1.0 : 1.0);        // EQTSYMMETRIC: the tangential components change sign
    for (int q = 0; q < g.dims; q++){
      const long long idq = (q == d ? id : ids);
      bp[q] = a.Bs[q][idq]; bm[q] = a.Bs[q][idq - st[q]];
      if (q != d){ bp[q] *= tsign; bm[q] *= tsign; }
      dB[q] = (A[q]*bp[q] - A[q]*bm[q]);
    }
    // sum of the two transverse flux differences in the reference's order
    // (ct_fill_mag_field.c:84-140): x: dBy + dBz, y: dBx + dBz, z: dBx + dBy
    const int qa = (d == 0 ? 1 : 0), qb = (d == 2 ? 1 : 2);
    // low side: (A*b+ + dB_a + dB_b)/A, left-associated; high side: (A*b- - (dB_a + dB_b))/A
    if (!hi_side) a.Bs[d][id - st[d]] = (A[d]*bp[d] + dB[qa] + dB[qb])/A[d];
    else          a.Bs[d][id]         = (A[d]*bm[d] - (dB[qa] + dB[qb]))/A[d];
  }
  if (fl.Bc){
    const int lo = hi_side ? g.end[d] + 1 : 0, hi = hi_side ? g.T[d] - 1 : g.beg[d] - 1;
    for (int n = lo; n <= hi; n++){
      c[d] = n;
      const long long id = gidx (g, c[2], c[1], c[0]);
      fl.Bc[id] = 0.5*(a.Bs[d][id] + a.Bs[d][id - st[d]]);
    }
  }
}

// ---------------------------------------------------------------------------
//  SHOCK_FLATTENING MULTID: FlagShock (flag_shock.c:79-230, Cartesian, ideal EOS).
//  Pass 1: a zone with div v < 0 and |grad p| > 5 min(p) lies in a shock; pass 2 (gather
//  form of the reference's scatter): the zone takes FLAG_HLL | FLAG_MINMOD, its six
//  neighbours FLAG_MINMOD.  Evaluated once per step on the stage-1 state (rk_step.c:86-88).
// ---------------------------------------------------------------------------
template <int NC, int PASS>
__global__ void __launch_bounds__(256)
flag_shock_kernel (const __grid_constant__ FlagArgs a)
{
  const Geom &g = a.g;
  const int ni = g.T[0], nj = g.T[1], nk = (NC == 3 ? g.T[2] : 1);
  long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= (long long)ni*nj*nk) return;
  const int i = (int)(t % ni), j = (int)((t/ni) % nj), k = (int)(t/((long long)ni*nj));
  const long long id = gidx (g, k, j, i);
  const long long sx = 1, sy = g.S1, sz = g.S12;
  const bool inner = i >= 1 && i < ni - 1 && j >= 1 && j < nj - 1 && (NC == 2 || (k >= 1 && k < nk - 1));
  if (PASS == 1){
    unsigned char sh = 0;
    if (inner){
      const double dvx1 = pg_div (a.vx[0][id + sx] - a.vx[0][id - sx], a.dxa[0] ? __ldg (a.dxa[0] + i) : g.dx[0]);
      const double dvx2 = pg_div (a.vx[1][id + sy] - a.vx[1][id - sy], a.dxa[1] ? __ldg (a.dxa[1] + j) : g.dx[1]);
      double divv = dvx1 + dvx2;
      if (NC == 3) divv = dvx1 + dvx2 + pg_div (a.vx[2][id + sz] - a.vx[2][id - sz], a.dxa[2] ? __ldg (a.dxa[2] + k) : g.dx[2]);
      if (divv < 0.0){
        double pt_min = a.prs[id];
        const double p1 = minv (a.prs[id + sx], a.prs[id - sx]), p2 = minv (a.prs[id + sy], a.prs[id - sy]);
        pt_min = minv (pt_min, p1);
        pt_min = minv (pt_min, p2);
        double gradp = fabs (a.prs[id + sx] - a.prs[id - sx]) + fabs (a.prs[id + sy] - a.prs[id - sy]);
        if (NC == 3){
          pt_min = minv (pt_min, minv (a.prs[id + sz], a.prs[id - sz]));
          gradp = fabs (a.prs[id + sx] - a.prs[id - sx]) + fabs (a.prs[id + sy] - a.prs[id - sy])
                + fabs (a.prs[id + sz] - a.prs[id - sz]);
        }
        if (gradp > 5.0*pt_min) sh = 1;                       // EPS_PSHOCK_FLATTEN
      }
    }
    a.shock[id] = sh;
  }else{
    // neighbours outside the array never lie in a shock (pass 1 wrote 0 on the outermost layer)
    unsigned char f = a.shock[id] ? (4 | 1) : 0;
    if ((i > 0 && a.shock[id - sx]) || (i < ni - 1 && a.shock[id + sx]) ||
        (j > 0 && a.shock[id - sy]) || (j < nj - 1 && a.shock[id + sy])) f |= 1;
    if (NC == 3 && ((k > 0 && a.shock[id - sz]) || (k < nk - 1 && a.shock[id + sz]))) f |= 1;
    a.flag[id] = f;
  }
}

int launch_flag_shock (const FlagArgs &a, cudaStream_t s)
{
  const Geom &g = a.g;
  const long long n = (long long)g.T[0]*g.T[1]*(g.dims == 3 ? g.T[2] : 1);
  if (g.dims == 3){
    flag_shock_kernel<3, 1><<<nblocks (n, 256), 256, 0, s>>>(a);
    flag_shock_kernel<3, 2><<<nblocks (n, 256), 256, 0, s>>>(a);
  }else{
    flag_shock_kernel<2, 1><<<nblocks (n, 256), 256, 0, s>>>(a);
    flag_shock_kernel<2, 2><<<nblocks (n, 256), 256, 0, s>>>(a);
  }
  return pg_launch_status (2);
}

// ---------------------------------------------------------------------------
//  run-time diagnostics: per-block partial sums over a grid-stride loop (fixed order
//  within a block: strided thread sums, then a shared-memory tree), summed on the host
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
analysis_kernel (const __grid_constant__ AnalysisArgs a)
{
  const Geom &g = a.g;
  const bool d3 = (g.dims == 3);
  const long long n = (long long)g.n[0]*g.n[1]*(d3 ? g.n[2] : 1);
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x*blockDim.x){
    const int i = g.beg[0] + (int)(t % g.n[0]), j = g.beg[1] + (int)((t/g.n[0]) % g.n[1]);
    const int k = d3 ? g.beg[2] + (int)(t/((long long)g.n[0]*g.n[1])) : 0;
    const long long id = gidx (g, k, j, i);
    const double rho = a.V[RHO][id], vx = a.V[VX1][id], vy = a.V[VX2][id], vz = d3 ? a.V[VX3][id] : 0.0;
    const double bx = a.V[BX1][id], by = a.V[BX2][id], bz = d3 ? a.V[BX3][id] : 0.0;
    s[0] += rho;
    s[1] += 0.5*rho*(vx*vx + vy*vy + vz*vz);
    s[2] += 0.5*(bx*bx + by*by + bz*bz);
    s[3] += a.V[PRS][id]*a.igmm1;
    s[4] += rho*vx; s[5] += rho*vy; s[6] += rho*vz;
    double div = (a.Bs[0][id] - a.Bs[0][id - 1])/g.dx[0] + (a.Bs[1][id] - a.Bs[1][id - g.S1])/g.dx[1];
    if (d3) div += (a.Bs[2][id] - a.Bs[2][id - g.S12])/g.dx[2];
    s[7] = maxv (s[7], fabs (div));
  }
  __shared__ double sh[8][256];
  for (int q = 0; q < 8; q++) sh[q][threadIdx.x] = s[q];
  __syncthreads ();
  for (int w = 128; w > 0; w >>= 1){
    if ((int)threadIdx.x < w){
      for (int q = 0; q < 7; q++) sh[q][threadIdx.x] += sh[q][threadIdx.x + w];
      sh[7][threadIdx.x] = maxv (sh[7][threadIdx.x], sh[7][threadIdx.x + w]);
    }
    __syncthreads ();
  }
  if (threadIdx.x < 8) a.partial[(size_t)blockIdx.x*8 + threadIdx.x] = sh[threadIdx.x][0];
}

// double -> float of the interior zones (Convert_dbl2flt, Src/bin_io.c:51-83: (float)(V*unit), unit = 1), optionally
// byte-swapped as write_vtk.c stores it
__global__ void __launch_bounds__(256)
cvt_float_kernel (const __grid_constant__ CvtArgs a)
{
  const Geom &g = a.g;
  const long long n1 = g.n[0], n2 = g.n[1], n3 = (g.dims == 3 ? g.n[2] : 1), nz = n1*n2*n3;
  const int nv = blockIdx.y;
  if (a.live[nv] < 0) return;
  for (long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x; t < nz; t += (long long)gridDim.x*blockDim.x){
    const int i = (int)(t % n1), j = (int)((t/n1) % n2), k = (int)(t/(n1*n2));
    const float f = (float)a.V[nv][gidx (g, g.beg[2] + k, g.beg[1] + j, g.beg[0] + i)];
    unsigned u = __float_as_uint (f);
    if (a.swap) u = __byte_perm (u, 0u, 0x0123);
    reinterpret_cast<unsigned *>(a.out)[(long long)a.live[nv]*nz + t] = u;
  }
}
int launch_cvt_float (const CvtArgs &a, cudaStream_t s)
{
  const Geom &g = a.g;
  const long long nz = (long long)g.n[0]*g.n[1]*(g.dims == 3 ? g.n[2] : 1);
  unsigned nb = nblocks (nz, 256); if (nb > 4096) nb = 4096;
  cvt_float_kernel<<<dim3 (nb, 8), 256, 0, s>>>(a);
  return pg_launch_status ();
}

int launch_analysis (const AnalysisArgs &a, int nb, cudaStream_t s)
{
  analysis_kernel<<<nb, 256, 0, s>>>(a);
  return pg_launch_status ();
}

// ---------------------------------------------------------------------------
//  halo pack / unpack (contiguous buffers for the inter-GPU exchange)
// ---------------------------------------------------------------------------
template <bool PACK>
__global__ void __launch_bounds__(256)
halo_kernel (const __grid_constant__ HaloArgs a)
{
  const Geom &g = a.g;
  const int f = blockIdx.y;
  const int ni = a.hi[f][0] - a.lo[f][0] + 1, nj = a.hi[f][1] - a.lo[f][1] + 1, nk = a.hi[f][2] - a.lo[f][2] + 1;
  const long long cnt = (long long)ni*nj*nk;
  for (long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x; t < cnt; t += (long long)gridDim.x*blockDim.x){
    const int i = a.lo[f][0] + (int)(t % ni), j = a.lo[f][1] + (int)((t/ni) % nj), k = a.lo[f][2] + (int)(t/((long long)ni*nj));
    const long long id = gidx (g, k, j, i);
    if (PACK) a.buf[a.offset[f] + t] = a.q[f][id];
    else      a.q[f][id] = a.buf[a.offset[f] + t];
  }
}

// all boxes of all neighbours in ONE launch (blockIdx.y = table entry)
template <bool PACK>
__global__ void __launch_bounds__(256)
halo_table_kernel (const HaloEntry *__restrict__ tab, const Geom g)
{
  const HaloEntry e = tab[blockIdx.y];
  const int ni = e.n[0], nj = e.n[1];
  for (long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x; t < e.count; t += (long long)gridDim.x*blockDim.x){
    const int i = e.lo[0] + (int)(t % ni), j = e.lo[1] + (int)((t/ni) % nj), k = e.lo[2] + (int)(t/((long long)ni*nj));
    const long long id = gidx (g, k, j, i);
    if (PACK) e.buf[t] = e.field[id];
    else      e.field[id] = e.buf[t];
  }
}

// ---------------------------------------------------------------------------
//  launchers
// ---------------------------------------------------------------------------

int launch_ct_emf (const CtArgs &a, cudaStream_t s)
{
  const Geom &g = a.g;
  const long long n = (long long)(g.n[0] + 1 + 2*a.ext)*(g.n[1] + 1 + 2*a.ext)*(g.dims == 3 ? g.n[2] + 1 + 2*a.ext : 1);
#define PG_LE(C) do { switch (a.avg){                                                   \
      case 1:  ct_emf_kernel<C, 1><<<nblocks (n, 128), 128, 0, s>>>(a); break;            \
      case 2:  ct_emf_kernel<C, 2><<<nblocks (n, 128), 128, 0, s>>>(a); break;            \
      case 3:  ct_emf_kernel<C, 3><<<nblocks (n, 128), 128, 0, s>>>(a); break;            \
      default: if (a.Ec[2]) ct_emf_kernel<C, 0, true><<<nblocks (n, 128), 128, 0, s>>>(a);  \
               else         ct_emf_kernel<C, 0><<<nblocks (n, 128), 128, 0, s>>>(a); } } while (0)
  if (g.dims == 3) PG_LE(3); else PG_LE(2);
#undef PG_LE
  return pg_launch_status ();
}

int launch_ct_update (const CtArgs &a, cudaStream_t s)
{
  const Geom &g = a.g;
  const long long n = (long long)(g.n[0] + 1 + 2*a.ext)*(g.n[1] + 1 + 2*a.ext)*(g.dims == 3 ? g.n[2] + 1 + 2*a.ext : 1);
  if (g.dims == 3) ct_update_kernel<3><<<nblocks (n, 128), 128, 0, s>>>(a);
  else             ct_update_kernel<2><<<nblocks (n, 128), 128, 0, s>>>(a);
  return pg_launch_status ();
}

int launch_final (const FinalArgs &a, cudaStream_t s)
{
  const Geom &g = a.g;
  long long n = (long long)a.box_n[0]*a.box_n[1]*(g.dims == 3 ? a.box_n[2] : 1);
  for (int b = 0; b < a.nbox; b++){           // several boxes: the grid covers the largest, blockIdx.y = box
    const long long nb = (long long)a.boxes_n[b][0]*a.boxes_n[b][1]*(g.dims == 3 ? a.boxes_n[b][2] : 1);
    if (b == 0 || nb > n) n = nb;
  }
  if (n <= 0) return 0;
  const dim3 grid (nblocks (n, 128), a.nbox ? a.nbox : 1);
  if (a.fuse_ct && !a.en_corr){
    if (g.dims == 3) final_kernel<3, false, true><<<grid, 128, 0, s>>>(a);
    else             final_kernel<2, false, true><<<grid, 128, 0, s>>>(a);
  }else if (a.en_corr){
    if (g.dims == 3) final_kernel<3, true><<<grid, 128, 0, s>>>(a);
    else             final_kernel<2, true><<<grid, 128, 0, s>>>(a);
  }else if (a.R3[RHO] && g.dims == 3){
    final_kernel<3, false, false, true><<<grid, 128, 0, s>>>(a);
  }else{
    if (g.dims == 3) final_kernel<3, false><<<grid, 128, 0, s>>>(a);
    else             final_kernel<2, false><<<grid, 128, 0, s>>>(a);
  }
  return pg_launch_status ();
}

int launch_bc (const BcArgs &a, cudaStream_t s)
{
  const Geom &g = a.g;
  long long nmax = 0;
  for (int f = 0; f < a.nf; f++){
    long long n = 1;
    for (int d = 0; d < 3; d++) n *= (a.f[f].hi[d] - a.f[f].lo[d] + 1);
    if (n > nmax) nmax = n;
  }
  for (int f = 0; f < a.nfill; f++){
    const int d = a.fill[f].side >> 1;
    const int d1 = (d == 0 ? 1 : 0), d2 = (d == 2 ? 1 : 2);
    const long long n = (long long)g.T[d1]*g.T[d2];
    if (n > nmax) nmax = n;
  }
  if (nmax <= 0 || a.nf + a.nfill <= 0) return 0;
  dim3 grid (nblocks (nmax, 128), a.nf + a.nfill);
  bc_kernel<<<grid, 128, 0, s>>>(a);
  return pg_launch_status ();
}

static int launch_halo (const HaloArgs &a, cudaStream_t s, bool pack)
{
  long long nmax = 0;
  for (int f = 0; f < a.nf; f++){
    long long n = 1;
    for (int d = 0; d < 3; d++) n *= (a.hi[f][d] - a.lo[f][d] + 1);
    if (n > nmax) nmax = n;
  }
  if (nmax <= 0 || a.nf <= 0) return 0;
  unsigned nb = nblocks (nmax, 256); if (nb > 4096) nb = 4096;
  dim3 grid (nb, a.nf);
  if (pack) halo_kernel<true><<<grid, 256, 0, s>>>(a);
  else      halo_kernel<false><<<grid, 256, 0, s>>>(a);
  return pg_launch_status ();
}
int launch_halo_table (const HaloEntry *tab, int n, long long maxcount, const Geom &g, bool pack, cudaStream_t s)
{
  if (n <= 0 || maxcount <= 0) return 0;
  unsigned nb = nblocks (maxcount, 256); if (nb > 256) nb = 256;
  dim3 grid (nb, n);
  if (pack) halo_table_kernel<true><<<grid, 256, 0, s>>>(tab, g);
  else      halo_table_kernel<false><<<grid, 256, 0, s>>>(tab, g);
  return pg_launch_status ();
}
int launch_halo_pack   (const HaloArgs &a, cudaStream_t s) { return launch_halo (a, s, true); }
int launch_halo_unpack (const HaloArgs &a, cudaStream_t s) { return launch_halo (a, s, false); }

} // namespace PG_NS
