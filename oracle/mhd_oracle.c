/*
 * mhd_oracle.c -- CPU restatement of the PLUTO 4.3 unsplit RK + CT ideal-MHD
 * step (see mhd_oracle.h for scope and parity status).  TEST INFRASTRUCTURE.
 *
 * The restatement follows the reference's arithmetic operation by operation
 * (same association order, same comparisons) so that, compiled with plain
 * `gcc -O2` on x86-64 (no FMA contraction, no -ffast-math), it reproduces the
 * reference's double-precision results bit for bit.  All `file:line`
 * citations are relative to /root/reference/Src.
 *
 * Internal storage: every 3-D scalar array is padded by one layer on each
 * side so that indices -1 .. T are addressable:
 *     idx(k,j,i) = ((k+1)*S2 + (j+1))*S1 + (i+1),   S = T + 2.
 * Staggered components use the reference's convention: Vs[d] at index i is
 * the face i+1/2 in direction d (valid from -1).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mhd_oracle.h"

#define NV   ORC_NV
#define RHO  ORC_RHO
#define VX1  ORC_VX1
#define VX2  ORC_VX2
#define VX3  ORC_VX3
#define BX1  ORC_BX1
#define BX2  ORC_BX2
#define BX3  ORC_BX3
#define PRS  ORC_PRS
#define MX1  VX1
#define ENG  PRS

/* reference macros.h:140-151 */
#define MAXV(a,b)     ((a) >= (b) ? (a) : (b))
#define MINV(a,b)     ((a) <= (b) ? (a) : (b))
#define ABS_MIN(a,b)  (fabs(a) < fabs(b) ? (a) : (b))
#define MINMOD(a,b)   ((a)*(b) > 0.0 ? (fabs(a) < fabs(b) ? (a):(b)):0.0)

struct Oracle {
  OracleConfig c;
  int ng, T[3], S1, S2, S3, tot;
  int beg[3], end[3];                  /* IBEG..KEND */
  double *Vc[NV], *Uc[NV], *U0[NV];
  double *Vs[3], *Bs0[3];
  double *dxa[3];                      /* zone widths of a non-uniform grid (oracle_set_grid: grid->dx[d][0..T-1]), else NULL */
  double *plmc[3][6];                  /* UNIFORM_CARTESIAN_GRID NO: cp, cm, wp, wm, dp, dm of every direction (oracle_set_plm_coeffs) */
  double *ppmc[3][4];                  /* PARABOLIC on a non-uniform grid: interface weights wp[i][-1 .. 2] of every direction (oracle_set_ppm_coeffs) */
  double *gf[3];                       /* per-zone body force (oracle_set_body_force), else NULL */
  double *phic, *phif[3];              /* body-force potential at centres and faces (oracle_set_body_potential), else NULL */
  double *ppen;                        /* face potential of the current pencil */
  double *exj, *exk, *eyi, *eyk, *ezi, *ezj;
  double *ex, *ey, *ez, *Ex1, *Ex2, *Ex3;
  signed char *svx, *svy, *svz;
  /* UCT_HLL (ct_emf.c:144-147,165-168,186-189, ct_stag_slopes.c:5-45): face speeds and velocity slopes */
  double *SfL[3], *SfR[3];             /* max(0,-SL), max(0,SR) at the faces of each direction */
  double *dvel[3][3];                  /* dvel[c][d] = d v_c / d x_d (limited slope vp - vm)   */
  double cur_SL, cur_SR;               /* Riemann fan speeds of the interface just solved      */
  double *gpen;                        /* body force of the current pencil, sweep direction (PrimSource) */
  unsigned char *flag;                 /* SHOCK_FLATTENING MULTID: FLAG_MINMOD 1, FLAG_HLL 4 (pluto.h:192-194) */
  unsigned char *pflag;                /* ... of the current pencil                            */
  int use_hll;                         /* the interface being solved takes the HLL flux        */
  /* CTU (ctu_step.c:214-224): L/R conservative states and half-step rhs of every direction, half-step U */
  double *Up[3][NV], *Um[3][NV], *rhs3[3][NV], *Uh[NV];
  double (*up)[NV], (*um)[NV], (*vn)[NV];   /* pencil: conservative L/R states, t^n zone values */
  double *C_dt;
  /* pencil scratch */
  int np;
  double (*v)[NV], (*vp)[NV], (*vm)[NV], (*dv)[NV];
  double (*Rp)[NV][NV];                /* CHAR_LIMITING: right eigenvectors per pencil index, never cleared (as stateC->Rp) */
  double (*flux)[NV], *press, *cmax, *bn, *SLp, *SRp;
  double max_mach, inv_dt_hyp;
  int    stage, floor_events;
  int    emf_ibeg, emf_iend, emf_jbeg, emf_jend, emf_kbeg, emf_kend;
  int    debug_stop_after;   /* tests only: leave oracle_advance after this stage */
};

#define IDX(o,k,j,i) ((((k)+1)*(o)->S2 + ((j)+1))*(o)->S1 + ((i)+1))
/* width of zone n in direction d: grid->dx[d][n] (set_grid.c); a uniform grid has the same double in every zone */
#define DXA(o,d,n)   ((o)->dxa[d] ? (o)->dxa[d][n] : (o)->c.dx[d])

static double *dalloc (int n) { return (double *)calloc((size_t)n, sizeof(double)); }

int oracle_nghost (const Oracle *o) { return o->ng; }

/* ********************************************************************* */
Oracle *oracle_create (const OracleConfig *cfg)
/* ghost count: reference get_nghost.c:32-50 (2 for LINEAR, 3 for
   PARABOLIC with MHD).                                                 */
{
  int d, nv;
  Oracle *o = (Oracle *)calloc(1, sizeof(Oracle));
  o->c  = *cfg;
  o->ng = (cfg->recon == ORC_RECON_PPM ? 3 : 2);
  if (cfg->shock_flattening && o->ng < 3) o->ng = 3;       /* get_nghost.c:67-77 */
  if (cfg->ctu) o->ng++;                                   /* CTU + CT, get_nghost.c:86-90 */
  for (d = 0; d < 3; d++){
    if (d < cfg->dims){
      o->T[d]   = cfg->n[d] + 2*o->ng;
      o->beg[d] = o->ng;
      o->end[d] = o->ng + cfg->n[d] - 1;
    }else{
      o->T[d] = 1; o->beg[d] = o->end[d] = 0;
    }
  }
  o->S1 = o->T[0] + 2; o->S2 = o->T[1] + 2; o->S3 = o->T[2] + 2;
  o->tot = o->S1*o->S2*o->S3;
  for (nv = 0; nv < NV; nv++){
    o->Vc[nv] = dalloc(o->tot); o->Uc[nv] = dalloc(o->tot); o->U0[nv] = dalloc(o->tot);
  }
  for (d = 0; d < 3; d++){ o->Vs[d] = dalloc(o->tot); o->Bs0[d] = dalloc(o->tot); }
  o->exj = dalloc(o->tot); o->exk = dalloc(o->tot); o->eyi = dalloc(o->tot);
  o->eyk = dalloc(o->tot); o->ezi = dalloc(o->tot); o->ezj = dalloc(o->tot);
  o->ex  = dalloc(o->tot); o->ey  = dalloc(o->tot); o->ez  = dalloc(o->tot);
  o->Ex1 = dalloc(o->tot); o->Ex2 = dalloc(o->tot); o->Ex3 = dalloc(o->tot);
  o->svx = (signed char *)calloc((size_t)o->tot, 1);
  o->svy = (signed char *)calloc((size_t)o->tot, 1);
  o->svz = (signed char *)calloc((size_t)o->tot, 1);
  o->C_dt = dalloc(o->tot);
  o->flag = (unsigned char *)calloc((size_t)o->tot, 1);
  if (cfg->emf_average == ORC_EMF_UCT_HLL){
    int c;
    for (d = 0; d < 3; d++){
      o->SfL[d] = dalloc(o->tot); o->SfR[d] = dalloc(o->tot);
      for (c = 0; c < 3; c++) o->dvel[c][d] = dalloc(o->tot);
    }
  }
  o->np = o->T[0];
  if (o->T[1] > o->np) o->np = o->T[1];
  if (o->T[2] > o->np) o->np = o->T[2];
  o->np += 8;
  o->v    = calloc((size_t)o->np, sizeof(*o->v));
  o->vp   = calloc((size_t)o->np, sizeof(*o->vp));
  o->vm   = calloc((size_t)o->np, sizeof(*o->vm));
  o->dv   = calloc((size_t)o->np, sizeof(*o->dv));
  o->Rp   = calloc((size_t)o->np, sizeof(*o->Rp));
  o->flux = calloc((size_t)o->np, sizeof(*o->flux));
  o->press = dalloc(o->np); o->cmax = dalloc(o->np); o->bn = dalloc(o->np);
  o->gpen = dalloc(o->np) + 4;
  o->ppen = dalloc(o->np) + 4;
  o->SLp = dalloc(o->np) + 4; o->SRp = dalloc(o->np) + 4;
  if (cfg->ctu){
    int c;
    for (d = 0; d < 3; d++) for (c = 0; c < NV; c++){
      o->Up[d][c] = dalloc(o->tot); o->Um[d][c] = dalloc(o->tot); o->rhs3[d][c] = dalloc(o->tot);
    }
    for (c = 0; c < NV; c++) o->Uh[c] = dalloc(o->tot);
    o->up = calloc((size_t)o->np, sizeof(*o->up)); o->up += 4;
    o->um = calloc((size_t)o->np, sizeof(*o->um)); o->um += 4;
    o->vn = calloc((size_t)o->np, sizeof(*o->vn)); o->vn += 4;
  }
  o->pflag = (unsigned char *)calloc((size_t)o->np, 1) + 4;
  /* shift pencil arrays so that index -2 is addressable */
  o->v += 4; o->vp += 4; o->vm += 4; o->dv += 4; o->flux += 4; o->Rp += 4;
  o->press += 4; o->cmax += 4; o->bn += 4;
  return o;
}

void oracle_destroy (Oracle *o)
{
  int d, nv;
  if (!o) return;
  for (nv = 0; nv < NV; nv++){ free(o->Vc[nv]); free(o->Uc[nv]); free(o->U0[nv]); }
  for (d = 0; d < 3; d++){ free(o->Vs[d]); free(o->Bs0[d]); }
  free(o->exj); free(o->exk); free(o->eyi); free(o->eyk); free(o->ezi); free(o->ezj);
  free(o->ex); free(o->ey); free(o->ez); free(o->Ex1); free(o->Ex2); free(o->Ex3);
  free(o->svx); free(o->svy); free(o->svz); free(o->C_dt);
  for (d = 0; d < 3; d++){ int q; free(o->dxa[d]); for (q = 0; q < 6; q++) free(o->plmc[d][q]); for (q = 0; q < 4; q++) free(o->ppmc[d][q]); }
  free(o->v - 4); free(o->vp - 4); free(o->vm - 4); free(o->dv - 4); free(o->flux - 4);
  free(o->press - 4); free(o->cmax - 4); free(o->bn - 4);
  free(o);
}

/* ********************************************************************* */
void oracle_set_grid (Oracle *o, const double *dx1, const double *dx2, const double *dx3)
/* Non-uniform Cartesian grid: dx_d[0 .. T_d-1] = the reference's grid->dx[d] (ghost zones included).  With the reference's
   default UNIFORM_CARTESIAN_GRID YES (plm_coeffs.h:23-29: every CARTESIAN build) the reconstruction keeps its uniform
   weights; the zone widths enter the flux difference (rhs.c:195), the inverse time step (update_stage.c:229-235 with
   inv_dl = 1/dx, set_geometry.c), CT_Update (ct_update.c:79-218: dt/dx2[j], dt/dx3[k], ...) and the face areas of
   FillMagneticField (ct_fill_mag_field.c:108-114, set_geometry.c:149,180,202). */
{
  const double *src[3] = {dx1, dx2, dx3};
  int d;
  for (d = 0; d < o->c.dims; d++){
    if (!o->dxa[d]) o->dxa[d] = dalloc (o->T[d]);
    memcpy (o->dxa[d], src[d], sizeof(double)*(size_t)o->T[d]);
  }
}

void oracle_set_ppm_coeffs (Oracle *o, int dir, const double *wm1, const double *w0, const double *w1, const double *w2)
/* PARABOLIC reconstruction on a non-uniform grid: the interface weights wp[i][-1], wp[i][0], wp[i][1], wp[i][2] of every zone as
   PPM_CoefficientsGet returns them for direction dir (ppm_coeffs.c:586-609; found numerically by PPM_FindWeights, :300-480, where the
   direction is not uniform; T entries each, set for 1 <= i <= T-3).  hp = hm = 3 on every Cartesian grid (ppm_coeffs.c:544-547). */
{
  const double *src[4] = {wm1, w0, w1, w2};
  int q;
  for (q = 0; q < 4; q++){
    if (!o->ppmc[dir][q]) o->ppmc[dir][q] = dalloc (o->T[dir]);
    memcpy (o->ppmc[dir][q], src[q], sizeof(double)*(size_t)o->T[dir]);
  }
}

void oracle_set_plm_coeffs (Oracle *o, int dir, const double *cp, const double *cm, const double *wp, const double *wm,
                            const double *dp, const double *dm)
/* UNIFORM_CARTESIAN_GRID NO (plm_coeffs.h:23-29): the grid-dependent weights of the linear reconstruction, as
   PLM_CoefficientsGet returns them for direction dir (plm_coeffs.c:30-104; T entries each, first and last unused):
   dvp = dv[i] wp, dvm = dv[i-1] wm (plm_states.c:157-164), the limiters of plm_coeffs.h:130-152 with cp, cm,
   vp = v + dv_lim dp, vm = v - dv_lim dm (:240-241). */
{
  const double *src[6] = {cp, cm, wp, wm, dp, dm};
  int q;
  for (q = 0; q < 6; q++){
    if (!o->plmc[dir][q]) o->plmc[dir][q] = dalloc (o->T[dir]);
    memcpy (o->plmc[dir][q], src[q], sizeof(double)*(size_t)o->T[dir]);
  }
}

void oracle_set_body_force (Oracle *o, const double *g1, const double *g2, const double *g3)
{
  const double *src[3] = {g1, g2, g3};
  int d, i, j, k;
  for (d = 0; d < o->c.dims; d++){
    if (!o->gf[d]) o->gf[d] = dalloc (o->tot);
    for (k = 0; k < o->T[2]; k++) for (j = 0; j < o->T[1]; j++) for (i = 0; i < o->T[0]; i++)
      o->gf[d][IDX(o,k,j,i)] = src[d][((size_t)k*o->T[1] + j)*o->T[0] + i];
  }
  o->c.body_force = 1;
}

void oracle_set_body_potential (Oracle *o, const double *phic, const double *pf1, const double *pf2, const double *pf3)
{
  const double *src[3] = {pf1, pf2, pf3};
  int d, i, j, k;
  if (!o->phic) o->phic = dalloc (o->tot);
  for (k = 0; k < o->T[2]; k++) for (j = 0; j < o->T[1]; j++) for (i = 0; i < o->T[0]; i++)
    o->phic[IDX(o,k,j,i)] = phic[((size_t)k*o->T[1] + j)*o->T[0] + i];
  for (d = 0; d < o->c.dims; d++){
    int e[3] = {0, 0, 0}, n1, n2;
    e[d] = 1;
    n1 = o->T[0] + e[0]; n2 = o->T[1] + e[1];
    if (!o->phif[d]) o->phif[d] = dalloc (o->tot);
    for (k = -e[2]; k < o->T[2]; k++) for (j = -e[1]; j < o->T[1]; j++) for (i = -e[0]; i < o->T[0]; i++)
      o->phif[d][IDX(o,k,j,i)] = src[d][((size_t)(k + e[2])*n2 + (j + e[1]))*n1 + (i + e[0])];
  }
}

void oracle_set_interior (Oracle *o, const double *vc, const double *bx1s,
                          const double *bx2s, const double *bx3s)
{
  int i, j, k, nv;
  int n1 = o->c.n[0], n2 = o->c.n[1], n3 = (o->c.dims == 3 ? o->c.n[2] : 1);
  int ib = o->beg[0], jb = o->beg[1], kb = o->beg[2];
  for (nv = 0; nv < NV; nv++)
  for (k = 0; k < n3; k++) for (j = 0; j < n2; j++) for (i = 0; i < n1; i++)
    o->Vc[nv][IDX(o,k+kb,j+jb,i+ib)] = vc[((size_t)(nv*n3 + k)*n2 + j)*n1 + i];
  for (k = 0; k < n3; k++) for (j = 0; j < n2; j++) for (i = 0; i <= n1; i++)
    o->Vs[0][IDX(o,k+kb,j+jb,i+ib-1)] = bx1s[((size_t)k*n2 + j)*(n1+1) + i];
  for (k = 0; k < n3; k++) for (j = 0; j <= n2; j++) for (i = 0; i < n1; i++)
    o->Vs[1][IDX(o,k+kb,j+jb-1,i+ib)] = bx2s[((size_t)k*(n2+1) + j)*n1 + i];
  if (o->c.dims == 3)
  for (k = 0; k <= n3; k++) for (j = 0; j < n2; j++) for (i = 0; i < n1; i++)
    o->Vs[2][IDX(o,k+kb-1,j+jb,i+ib)] = bx3s[((size_t)k*n2 + j)*n1 + i];
}

void oracle_get_interior (const Oracle *o, double *vc, double *bx1s,
                          double *bx2s, double *bx3s)
{
  int i, j, k, nv;
  int n1 = o->c.n[0], n2 = o->c.n[1], n3 = (o->c.dims == 3 ? o->c.n[2] : 1);
  int ib = o->beg[0], jb = o->beg[1], kb = o->beg[2];
  for (nv = 0; nv < NV; nv++)
  for (k = 0; k < n3; k++) for (j = 0; j < n2; j++) for (i = 0; i < n1; i++)
    vc[((size_t)(nv*n3 + k)*n2 + j)*n1 + i] = o->Vc[nv][IDX(o,k+kb,j+jb,i+ib)];
  for (k = 0; k < n3; k++) for (j = 0; j < n2; j++) for (i = 0; i <= n1; i++)
    bx1s[((size_t)k*n2 + j)*(n1+1) + i] = o->Vs[0][IDX(o,k+kb,j+jb,i+ib-1)];
  for (k = 0; k < n3; k++) for (j = 0; j <= n2; j++) for (i = 0; i < n1; i++)
    bx2s[((size_t)k*(n2+1) + j)*n1 + i] = o->Vs[1][IDX(o,k+kb,j+jb-1,i+ib)];
  if (o->c.dims == 3 && bx3s)
  for (k = 0; k <= n3; k++) for (j = 0; j < n2; j++) for (i = 0; i < n1; i++)
    bx3s[((size_t)k*n2 + j)*n1 + i] = o->Vs[2][IDX(o,k+kb-1,j+jb,i+ib)];
}

/* =====================================================================
   Boundary conditions  (reference boundary.c:41-315, 439-564;
   MHD/CT/ct_fill_mag_field.c:38-179; MHD/CT/ct_field_average.c:134-272)
   ===================================================================== */

typedef struct { int ib, ie, jb, je, kb, ke; } Box;   /* may run backwards */

#define BOXLOOP(b,k,j,i) \
  for (dk_ = ((k = (b).kb) <= (b).ke ? 1:-1); k != (b).ke + dk_; k += dk_) \
  for (dj_ = ((j = (b).jb) <= (b).je ? 1:-1); j != (b).je + dj_; j += dj_) \
  for (di_ = ((i = (b).ib) <= (b).ie ? 1:-1); i != (b).ie + di_; i += di_)

static void outflow_bound (Oracle *o, double *q, Box b, int side)
/* boundary.c:439-477 */
{
  int i, j, k, di_, dj_, dk_;
  BOXLOOP(b,k,j,i){
    int ks = k, js = j, is = i;
    switch (side){
      case 0: is = o->beg[0]; break;  case 1: is = o->end[0]; break;
      case 2: js = o->beg[1]; break;  case 3: js = o->end[1]; break;
      case 4: ks = o->beg[2]; break;  case 5: ks = o->end[2]; break;
    }
    q[IDX(o,k,j,i)] = q[IDX(o,ks,js,is)];
  }
}

static void periodic_bound (Oracle *o, double *q, Box b, int side)
/* boundary.c:480-518 */
{
  int i, j, k, di_, dj_, dk_;
  int n1 = o->c.n[0], n2 = o->c.n[1], n3 = o->c.n[2];
  BOXLOOP(b,k,j,i){
    int ks = k, js = j, is = i;
    switch (side){
      case 0: is = i + n1; break;  case 1: is = i - n1; break;
      case 2: js = j + n2; break;  case 3: js = j - n2; break;
      case 4: ks = k + n3; break;  case 5: ks = k - n3; break;
    }
    q[IDX(o,k,j,i)] = q[IDX(o,ks,js,is)];
  }
}

static void reflective_bound (Oracle *o, double *q, int s, Box b, int side)
/* boundary.c:521-564 */
{
  int i, j, k, di_, dj_, dk_;
  BOXLOOP(b,k,j,i){
    int ks = k, js = j, is = i;
    switch (side){
      case 0: is = 2*o->beg[0]-i-1; break;  case 1: is = 2*o->end[0]-i+1; break;
      case 2: js = 2*o->beg[1]-j-1; break;  case 3: js = 2*o->end[1]-j+1; break;
      case 4: ks = 2*o->beg[2]-k-1; break;  case 5: ks = 2*o->end[2]-k+1; break;
    }
    q[IDX(o,k,j,i)] = s*q[IDX(o,ks,js,is)];
  }
}

static void fill_magnetic_field (Oracle *o, int side)
/* ct_fill_mag_field.c:38-179.  Cartesian areas (set_geometry.c:149,180,202):
   Ax1 = 1.0*dx2*dx3, Ax2 = dx1*1.0*dx3, Ax3 = dx1*dx2*1.0 (2-D: dx3 dropped). */
{
  int ibeg, iend, jbeg, jend, kbeg, kend, di = 1, dj = 1, dk = 1, i, j, k;
  int dims = o->c.dims;
  double *bx = o->Vs[0], *by = o->Vs[1], *bz = o->Vs[2];
  double Ax, Ay, Az;

  ibeg = 0; iend = o->T[0]-1;
  jbeg = 0; jend = o->T[1]-1;
  if (dims == 3){ kbeg = 0; kend = o->T[2]-1; } else { kbeg = kend = 0; }
  if (side == 0){ ibeg = o->beg[0]-1; iend = 0; di = -1; }
  if (side == 1)  ibeg = o->end[0]+1;
  if (side == 2){ jbeg = o->beg[1]-1; jend = 0; dj = -1; }
  if (side == 3)  jbeg = o->end[1]+1;
  if (side == 4){ kbeg = o->beg[2]-1; kend = 0; dk = -1; }
  if (side == 5)  kbeg = o->end[2]+1;

  for (k = kbeg; dk*k <= dk*kend; k += dk){
  for (j = jbeg; dj*j <= dj*jend; j += dj){
  for (i = ibeg; di*i <= di*iend; i += di){
    double bxp = bx[IDX(o,k,j,i)], bxm = bx[IDX(o,k,j,i-1)];
    double byp = by[IDX(o,k,j,i)], bym = by[IDX(o,k,j-1,i)];
    double bzp = 0.0, bzm = 0.0, dBx, dBy, dBz = 0.0;
    double dx1 = DXA(o,0,i), dx2 = DXA(o,1,j), dx3 = (dims == 3 ? DXA(o,2,k) : 1.0);
    if (dims == 3){ Ax = 1.0*dx2*dx3; Ay = dx1*1.0*dx3; Az = dx1*dx2*1.0; }
    else          { Ax = 1.0*dx2;     Ay = dx1*1.0;     Az = 0.0; }
    if (dims == 3){ bzp = bz[IDX(o,k,j,i)]; bzm = bz[IDX(o,k-1,j,i)]; }
    dBx = (Ax*bxp - Ax*bxm);
    dBy = (Ay*byp - Ay*bym);
    if (dims == 3) dBz = (Az*bzp - Az*bzm);
    if      (side == 0) bx[IDX(o,k,j,i-1)] = (Ax*bxp + dBy + dBz)/Ax;
    else if (side == 1) bx[IDX(o,k,j,i)]   = (Ax*bxm - (dBy + dBz))/Ax;
    else if (side == 2) by[IDX(o,k,j-1,i)] = (Ay*byp + dBx + dBz)/Ay;
    else if (side == 3) by[IDX(o,k,j,i)]   = (Ay*bym - (dBx + dBz))/Ay;
    else if (side == 4) bz[IDX(o,k-1,j,i)] = (Az*bzp + dBx + dBy)/Az;
    else if (side == 5) bz[IDX(o,k,j,i)]   = (Az*bzm - (dBx + dBy))/Az;
  }}}
}

static void average_normal_mag_field (Oracle *o, int side)
/* ct_field_average.c:134-272 (Cartesian) */
{
  int i, j, k, d = side/2;
  int lo[3] = {0,0,0}, hi[3];
  double *B = o->Vc[BX1+d], *b = o->Vs[d];
  hi[0] = o->T[0]-1; hi[1] = o->T[1]-1; hi[2] = o->T[2]-1;
  if (side & 1) lo[d] = o->end[d]+1; else hi[d] = o->beg[d]-1;
  for (k = lo[2]; k <= hi[2]; k++) for (j = lo[1]; j <= hi[1]; j++)
  for (i = lo[0]; i <= hi[0]; i++){
    int im = IDX(o, k - (d==2), j - (d==1), i - (d==0));
    B[IDX(o,k,j,i)] = 0.5*(b[IDX(o,k,j,i)] + b[im]);
  }
}

static void boundary (Oracle *o)
/* boundary.c:41-315, serial (no AL_Exchange), ALL_DIR */
{
  int is, nv, d, dims = o->c.dims;
  for (is = 0; is < 2*dims; is++){
    int type = o->c.bc[is];
    Box cb, fb[3];
    cb.ib = 0; cb.ie = o->T[0]-1; cb.jb = 0; cb.je = o->T[1]-1; cb.kb = 0; cb.ke = o->T[2]-1;
    if      (is == 0){ cb.ib = o->beg[0]-1; cb.ie = 0; }
    else if (is == 1){ cb.ib = o->end[0]+1; cb.ie = o->T[0]-1; }
    else if (is == 2){ cb.jb = o->beg[1]-1; cb.je = 0; }
    else if (is == 3){ cb.jb = o->end[1]+1; cb.je = o->T[1]-1; }
    else if (is == 4){ cb.kb = o->beg[2]-1; cb.ke = 0; }
    else if (is == 5){ cb.kb = o->end[2]+1; cb.ke = o->T[2]-1; }
    /* staggered boxes, boundary.c:157-164 */
    fb[0] = fb[1] = fb[2] = cb;
    if (cb.ib < cb.ie) fb[0].ib = cb.ib-1; else fb[0].ie = cb.ie-1;
    if (cb.jb < cb.je) fb[1].jb = cb.jb-1; else fb[1].je = cb.je-1;
    if (cb.kb < cb.ke) fb[2].kb = cb.kb-1; else fb[2].ke = cb.ke-1;
    /* NB: in 2-D kb == ke == 0 -> "kb < ke" false -> ke-1 = -1: the x3 face
       box is never used in 2-D (D_EXPAND drops it). */

    if (type == ORC_BC_OUTFLOW){
      for (nv = 0; nv < NV; nv++) outflow_bound (o, o->Vc[nv], cb, is);
      for (d = 0; d < dims; d++) if (d != is/2) outflow_bound (o, o->Vs[d], fb[d], is);
      fill_magnetic_field (o, is);
      average_normal_mag_field (o, is);
    }else if (type == ORC_BC_REFLECTIVE || type == ORC_BC_EQTSYMMETRIC){
      /* FlipSign, boundary.c:318-436.  REFLECTIVE: normal v and normal B change sign; EQTSYMMETRIC (:423-427): normal v
         and the two TRANSVERSE field components do */
      const int eqt = (type == ORC_BC_EQTSYMMETRIC);
      for (nv = 0; nv < NV; nv++){
        int s = 1;
        if (nv == VX1 + is/2) s = -1;
        if (nv >= BX1 && nv <= BX1 + 2) s = ((nv == BX1 + is/2) != eqt) ? -1 : 1;
        reflective_bound (o, o->Vc[nv], s, cb, is);
      }
      for (d = 0; d < dims; d++) if (d != is/2) reflective_bound (o, o->Vs[d], eqt ? -1 : 1, fb[d], is);
      fill_magnetic_field (o, is);
    }else if (type == ORC_BC_PERIODIC){
      for (nv = 0; nv < NV; nv++) periodic_bound (o, o->Vc[nv], cb, is);
      for (d = 0; d < dims; d++) periodic_bound (o, o->Vs[d], fb[d], is);
    }
  }
}

/* =====================================================================
   Mappers (reference MHD/mappers.c:25-86, 88-254)
   ===================================================================== */

static void prim_to_cons (const Oracle *o, const double *v, double *u)
{
  double gmm1 = o->c.gamma - 1.0, kinb2;
  u[RHO] = v[RHO];
  u[MX1]   = v[RHO]*v[VX1];
  u[MX1+1] = v[RHO]*v[VX2];
  u[MX1+2] = v[RHO]*v[VX3];
  u[BX1] = v[BX1]; u[BX2] = v[BX2]; u[BX3] = v[BX3];
  kinb2  = v[VX1]*v[VX1] + v[VX2]*v[VX2] + v[VX3]*v[VX3];
  kinb2  = v[RHO]*kinb2 + v[BX1]*v[BX1] + v[BX2]*v[BX2] + v[BX3]*v[BX3];  /* left-assoc: EXPAND has no parens */
  kinb2 *= 0.5;
  u[ENG] = kinb2 + v[PRS]/gmm1;
}

static int cons_to_prim (const Oracle *o, double *u, double *v)
/* returns 1 when a floor was applied (FLAG_CONS2PRIM_FAIL) */
{
  double gmm1 = o->c.gamma - 1.0, m2, b2, tau, kinb2;
  int fail = 0;
  m2 = u[MX1]*u[MX1] + u[MX1+1]*u[MX1+1] + u[MX1+2]*u[MX1+2];
  b2 = u[BX1]*u[BX1] + u[BX2]*u[BX2] + u[BX3]*u[BX3];
  if (u[RHO] < 0.0){ u[RHO] = o->c.small_dn; fail = 1; }
  v[RHO] = u[RHO];
  tau = 1.0/u[RHO];
  v[VX1] = u[MX1]*tau; v[VX2] = u[MX1+1]*tau; v[VX3] = u[MX1+2]*tau;
  v[BX1] = u[BX1]; v[BX2] = u[BX2]; v[BX3] = u[BX3];
  kinb2 = 0.5*(m2*tau + b2);
  if (u[ENG] < 0.0){ u[ENG] = o->c.small_pr/gmm1 + kinb2; fail = 1; }
  v[PRS] = gmm1*(u[ENG] - kinb2);
  if (v[PRS] < 0.0){
    v[PRS] = o->c.small_pr;
    u[ENG] = v[PRS]/gmm1 + kinb2;
    fail = 1;
  }
  return fail;
}

/* =====================================================================
   Reconstruction
   ===================================================================== */

static double single_limiter (int lim, double dvp, double dvm)
/* plm_coeffs.h:72-123, uniform Cartesian grid (cp = cm = 2) */
{
  double dv;
  switch (lim){
    case ORC_LIM_FLAT:   return 0.0;
    case ORC_LIM_MINMOD: return (dvp*dvm > 0.0 ? ABS_MIN(dvp, dvm) : 0.0);
    case ORC_LIM_VANALBADA:
      if (dvp*dvm > 0.0){
        double dpp = dvp*dvp, dmm = dvm*dvm;
        dv = (dvp*(dmm + 1.e-18) + dvm*(dpp + 1.e-18))/(dpp + dmm + 1.e-18);
      }else dv = 0.0;
      return dv;
    case ORC_LIM_OSPRE:
      return (dvp*dvm > 0.0 ? 1.5*dvp*dvm*(dvm + dvp)/(dvp*dvp + dvm*dvm + dvp*dvm) : 0.0);
    case ORC_LIM_UMIST:
      if (dvp*dvm > 0.0){
        double ddp = 0.25*(dvp + 3.0*dvm), ddm = 0.25*(dvm + 3.0*dvp);
        double d2 = 2.0*ABS_MIN(dvp, dvm);
        d2 = ABS_MIN(d2, ddp);
        dv = ABS_MIN(d2, ddm);
      }else dv = 0.0;
      return dv;
    case ORC_LIM_VANLEER:
      return (dvp*dvm > 0.0 ? 2.0*dvp*dvm/(dvp + dvm) : 0.0);
    default:             /* MC */
      if (dvp*dvm > 0.0){
        double qc = 0.5*(dvm + dvp), scrh = 2.0*ABS_MIN(dvp, dvm);
        dv = ABS_MIN(qc, scrh);
      }else dv = 0.0;
      return dv;
  }
}

static double general_limiter (int lim, int nv, double dvp, double dvm, double cp, double cm)
/* plm_coeffs.h:72-152 with UNIFORM_CARTESIAN_GRID NO: the limiters "on irregular or non-Cartesian grids" (OS, VL, MC take the
   weights cp, cm; FL, MM, VA, UM are the same on every grid).  lim == DEFAULT: MC on the density, minmod on the pressure,
   van Leer on velocity and field (plm_states.c:192-227). */
{
  double dv;
  if (lim == ORC_LIM_DEFAULT) lim = (nv == RHO ? ORC_LIM_MC : nv == PRS ? ORC_LIM_MINMOD : ORC_LIM_VANLEER);
  switch (lim){
    case ORC_LIM_OSPRE:
      if (dvp*dvm > 0.0){
        double den = 2.0*dvp*dvp + 2.0*dvm*dvm + (cp + cm - 2.0)*dvp*dvm;
        dv = dvp*dvm*((1.0+cp)*dvm + (1.0+cm)*dvp)/den;
      }else dv = 0.0;
      return dv;
    case ORC_LIM_VANLEER:
      return (dvp*dvm > 0.0 ? dvp*dvm*(cp*dvm + cm*dvp)/(dvp*dvp + dvm*dvm + (cp + cm - 2.0)*dvp*dvm) : 0.0);
    case ORC_LIM_MC:
      if (dvp*dvm > 0.0){
        double qc = 0.5*(dvm + dvp), scrh = ABS_MIN(dvp*cp, dvm*cm);
        dv = ABS_MIN(qc, scrh);
      }else dv = 0.0;
      return dv;
    default:
      return single_limiter (lim, dvp, dvm);
  }
}

static void states_plm (Oracle *o, int beg, int end, int bxn)
/* plm_states.c:80-312, CHAR_LIMITING NO, LIMITER DEFAULT,
   UNIFORM_CARTESIAN_GRID YES (cp=cm=2, wp=wm=1, dp=dm=0.5).
   Limiter macros: plm_coeffs.h:72-123.                                  */
{
  int i, nv;
  double (*v)[NV] = o->v, (*vp)[NV] = o->vp, (*vm)[NV] = o->vm, (*dv)[NV] = o->dv;
  for (i = beg-1; i <= end; i++)
    for (nv = 0; nv < NV; nv++) dv[i][nv] = v[i+1][nv] - v[i][nv];

  if (o->plmc[bxn - BX1][0]){                 /* UNIFORM_CARTESIAN_GRID NO: grid-dependent weights (plm_states.c:122-124, 156-164) */
    double *const *c = o->plmc[bxn - BX1];
    for (i = beg; i <= end; i++){
      const double cp = c[0][i], cm = c[1][i], wp = c[2][i], wm = c[3][i], dp = c[4][i], dm = c[5][i];
      for (nv = 0; nv < NV; nv++){
        const double dvp = dv[i][nv]*wp, dvm = dv[i-1][nv]*wm;
        const double lim = general_limiter (o->c.limiter, nv, dvp, dvm, cp, cm);
        vp[i][nv] = v[i][nv] + lim*dp;
        vm[i][nv] = v[i][nv] - lim*dm;
      }
    }
    for (i = beg-1; i <= end; i++) vp[i][bxn] = vm[i+1][bxn] = o->bn[i];
    return;
  }
  for (i = beg; i <= end; i++){
    double dvl[NV];
    for (nv = 0; nv < NV; nv++){
      double dvp = dv[i][nv], dvm = dv[i-1][nv], lim;
      if (o->pflag[i] & 1){                   /* FLAG_MINMOD, plm_states.c:174-180 */
        lim = (dvp*dvm > 0.0 ? ABS_MIN(dvp, dvm) : 0.0);
      }else if (o->c.limiter != ORC_LIM_DEFAULT){   /* same limiter for all variables, plm_states.c:234-236 */
        lim = single_limiter (o->c.limiter, dvp, dvm);
      }else if (nv == RHO){                   /* MC */
        if (dvp*dvm > 0.0){
          double qc = 0.5*(dvm + dvp), scrh = 2.0*ABS_MIN(dvp, dvm);
          lim = ABS_MIN(qc, scrh);
        }else lim = 0.0;
      }else if (nv == PRS){                   /* minmod */
        lim = (dvp*dvm > 0.0 ? ABS_MIN(dvp, dvm) : 0.0);
      }else{                                  /* van Leer */
        lim = (dvp*dvm > 0.0 ? 2.0*dvp*dvm/(dvp + dvm) : 0.0);
      }
      dvl[nv] = lim;
    }
    for (nv = 0; nv < NV; nv++){
      vp[i][nv] = v[i][nv] + dvl[nv]*0.5;
      vm[i][nv] = v[i][nv] - dvl[nv]*0.5;
    }
  }
  for (i = beg-1; i <= end; i++) vp[i][bxn] = vm[i+1][bxn] = o->bn[i];
}

typedef struct { int vn, vt, vb, bn, bt, bb; } Dirs;

enum { KFASTM = 0, KFASTP, KENTRP, KDIVB, KSLOWM, KSLOWP, KALFVM, KALFVP, NWAVE };    /* MHD/mod_defs.h:54-74 */

static void prim_eigenvectors (const Oracle *o, const double *qv, Dirs q, double (*RR)[NV], double LL[NWAVE][NV], double *lambda)
/* PrimEigenvectors (eigenv.c:190-560) for one zone, ideal EOS, CT (no div.B wave: lambda[KDIVB] = 0, eigenv.c:410-415).
   Only the entries this sweep direction defines are written into RR (= stateC->Rp[i], persistent in the reference);
   LL is cleared by the reference for every zone (eigenv.c:262-266 sets the left eigenvectors it uses). */
{
  const int nc = o->c.dims;
  const double sqrt_1_2 = 0.70710678118654752440;
  int k;
  double a2 = o->c.gamma*qv[PRS]/qv[RHO];            /* SoundSpeed2, eos.c:16-42 */
  double u, tau, sqrt_rho, scrh0, scrh1, scrh2, scrh3, scrh4, b2, ca2, A2, At2, cf2, cs2, cf, cs, ca, a;
  double alpha_f, alpha_s, beta_y, beta_z = 0.0, S;
  memset (LL, 0, sizeof (double)*NWAVE*NV);

  u   = qv[q.vn];
  tau = 1.0/qv[RHO];
  sqrt_rho = sqrt(qv[RHO]);
  scrh2 = qv[q.bn]*qv[q.bn];
  if (nc == 3) scrh3 = 0.0 + qv[q.bt]*qv[q.bt] + qv[q.bb]*qv[q.bb];
  else         scrh3 = 0.0 + qv[q.bt]*qv[q.bt];
  b2  = scrh2 + scrh3;
  ca2 = scrh2*tau;
  A2  = b2*tau;
  At2 = scrh3*tau;
  scrh1 = a2 - A2;
  scrh0 = sqrt(scrh1*scrh1 + 4.0*a2*At2);
  cf2 = 0.5*(a2 + A2 + scrh0);
  cs2 = a2*ca2/cf2;
  cf = sqrt(cf2); cs = sqrt(cs2); ca = sqrt(ca2); a = sqrt(a2);
  if (cf == cs){
    alpha_f = 1.0; alpha_s = 0.0;
  }else{
    scrh0   = 1.0/scrh0;
    alpha_f = (a2 - cs2)*scrh0;
    alpha_s = (cf2 - a2)*scrh0;
    alpha_f = MAXV(0.0, alpha_f);
    alpha_s = MAXV(0.0, alpha_s);
    alpha_f = sqrt(alpha_f);
    alpha_s = sqrt(alpha_s);
  }
  scrh0 = sqrt(scrh3);
  if (scrh0 > 1.e-9){
    if (nc == 3){ beta_y = qv[q.bt]/scrh0; beta_z = qv[q.bb]/scrh0; }
    else          beta_y = (qv[q.bt] >= 0.0 ? 1.0 : -1.0);
  }else{
    if (nc == 3) beta_z = beta_y = sqrt_1_2;
    else         beta_y = 1.0;
  }
  S = (qv[q.bn] >= 0.0 ? 1.0 : -1.0);
  (void)ca;

  /* fast wave u - cf */
  k = KFASTM;
  scrh0 = alpha_s*cs*S;
  scrh1 = alpha_s*sqrt_rho*a;
  scrh2 = 0.5/a2;
  scrh3 = scrh2*tau;
  RR[RHO][k] = qv[RHO]*alpha_f;
  RR[q.vn][k] = -cf*alpha_f;
  RR[q.vt][k] = scrh0*beta_y;
  if (nc == 3) RR[q.vb][k] = scrh0*beta_z;
  RR[q.bt][k] = scrh1*beta_y;
  if (nc == 3) RR[q.bb][k] = scrh1*beta_z;
  scrh4 = alpha_f*a2*qv[RHO];
  RR[PRS][k] = scrh4;
  LL[k][q.vn] = RR[q.vn][k]*scrh2;
  LL[k][q.vt] = RR[q.vt][k]*scrh2;
  if (nc == 3) LL[k][q.vb] = RR[q.vb][k]*scrh2;
  LL[k][q.bt] = RR[q.bt][k]*scrh3;
  if (nc == 3) LL[k][q.bb] = RR[q.bb][k]*scrh3;
  LL[k][PRS] = alpha_f*scrh3;
  /* fast wave u + cf */
  k = KFASTP;
  RR[RHO][k] = RR[RHO][KFASTM];
  RR[q.vn][k] = -RR[q.vn][KFASTM];
  RR[q.vt][k] = -RR[q.vt][KFASTM];
  if (nc == 3) RR[q.vb][k] = -RR[q.vb][KFASTM];
  RR[q.bt][k] = RR[q.bt][KFASTM];
  if (nc == 3) RR[q.bb][k] = RR[q.bb][KFASTM];
  RR[PRS][k] = RR[PRS][KFASTM];
  /* entropy wave */
  k = KENTRP;
  RR[RHO][k] = 1.0;
  LL[k][RHO] = 1.0;
  LL[k][PRS] = -1.0/a2;
  /* slow wave u - cs */
  k = KSLOWM;
  scrh0 = alpha_f*cf*S;
  scrh1 = alpha_f*sqrt_rho*a;
  RR[RHO][k] = qv[RHO]*alpha_s;
  RR[q.vn][k] = -cs*alpha_s;
  RR[q.vt][k] = -scrh0*beta_y;
  if (nc == 3) RR[q.vb][k] = -scrh0*beta_z;
  RR[q.bt][k] = -scrh1*beta_y;
  if (nc == 3) RR[q.bb][k] = -scrh1*beta_z;
  scrh4 = alpha_s*a2*qv[RHO];
  RR[PRS][k] = scrh4;
  LL[k][q.vn] = RR[q.vn][k]*scrh2;
  LL[k][q.vt] = RR[q.vt][k]*scrh2;
  if (nc == 3) LL[k][q.vb] = RR[q.vb][k]*scrh2;
  LL[k][q.bt] = RR[q.bt][k]*scrh3;
  if (nc == 3) LL[k][q.bb] = RR[q.bb][k]*scrh3;
  LL[k][PRS] = alpha_s*scrh3;
  /* slow wave u + cs */
  k = KSLOWP;
  RR[RHO][k] = RR[RHO][KSLOWM];
  RR[q.vn][k] = -RR[q.vn][KSLOWM];
  RR[q.vt][k] = -RR[q.vt][KSLOWM];
  if (nc == 3) RR[q.vb][k] = -RR[q.vb][KSLOWM];
  RR[q.bt][k] = RR[q.bt][KSLOWM];
  if (nc == 3) RR[q.bb][k] = RR[q.bb][KSLOWM];
  RR[PRS][k] = scrh4;
  if (nc == 3){
    /* Alfven waves */
    k = KALFVM;
    scrh2 = beta_y*sqrt_1_2;
    scrh3 = beta_z*sqrt_1_2;
    RR[q.vt][k] = -scrh3;
    RR[q.vb][k] =  scrh2;
    RR[q.bt][k] = -scrh3*sqrt_rho*S;
    RR[q.bb][k] =  scrh2*sqrt_rho*S;
    LL[k][q.vt] = RR[q.vt][k];
    LL[k][q.vb] = RR[q.vb][k];
    LL[k][q.bt] = RR[q.bt][k]*tau;
    LL[k][q.bb] = RR[q.bb][k]*tau;
    k = KALFVP;
    RR[q.vt][k] =   RR[q.vt][KALFVM];
    RR[q.vb][k] =   RR[q.vb][KALFVM];
    RR[q.bt][k] = - RR[q.bt][KALFVM];
    RR[q.bb][k] = - RR[q.bb][KALFVM];
  }

  if (lambda){
    lambda[KFASTM] = u - cf; lambda[KFASTP] = u + cf; lambda[KENTRP] = u; lambda[KDIVB] = 0.0;
    lambda[KSLOWM] = u - cs; lambda[KSLOWP] = u + cs;
    if (nc == 3){ lambda[KALFVM] = u - ca; lambda[KALFVP] = u + ca; }
  }
}

static void states_plm_char (Oracle *o, int beg, int end, Dirs q)
/* plm_states.c:448-706 (CHAR_LIMITING YES, UNIFORM_CARTESIAN_GRID YES: cp = cm = 2, dp = dm = 0.5, cpk = cmk = kstp),
   PrimEigenvectors eigenv.c:190-560 (ideal EOS, CT: no div.B wave), PrimToChar eigenv.c:1310-1400.
   2 components (DIMENSIONS = COMPONENTS = 2): 6 waves, beta_y = sign(Bt); 3 components: 8 waves.
   o->Rp[i] is persistent, like stateC->Rp of the reference: only the entries a sweep direction defines are written. */
{
  const int nc = o->c.dims, nw = (nc == 3 ? 8 : 6);
  const double sqrt_1_2 = 0.70710678118654752440;
  int i, nv, k;
  double (*v)[NV] = o->v, (*vp)[NV] = o->vp, (*vm)[NV] = o->vm, (*dv)[NV] = o->dv;
  double kstp[NWAVE];
  for (i = beg-1; i <= end; i++)
    for (nv = 0; nv < NV; nv++) dv[i][nv] = v[i+1][nv] - v[i][nv];
  for (k = 0; k < NWAVE; k++) kstp[k] = 2.0;
  kstp[KFASTP] = kstp[KFASTM] = 1.0;
  kstp[KSLOWP] = kstp[KSLOWM] = 1.0;

  for (i = beg; i <= end; i++){
    double (*RR)[NV] = o->Rp[i];           /* RR[nv][k] */
    double LL[NWAVE][NV];
    double dvp[NV], dvm[NV], dwp[NWAVE], dwm[NWAVE], dw_lim[NWAVE], dv_lim[NV];
    prim_eigenvectors (o, v[i], q, RR, LL, NULL);

    /* 2a. undivided differences projected on the characteristics (PrimToChar) */
    for (nv = 0; nv < NV; nv++){ dvp[nv] = dv[i][nv]; dvm[nv] = dv[i-1][nv]; }
    {
      const double *pass[2] = {dvm, dvp};
      double *out[2] = {dwm, dwp};
      int s;
      for (s = 0; s < 2; s++){
        const double *d = pass[s]; double *w = out[s], wv, wB; const double *L;
        for (k = 0; k < NWAVE; k++) w[k] = 0.0;
        L = LL[KFASTM];
        if (nc == 3){ wv = L[q.vn]*d[q.vn] + L[q.vt]*d[q.vt] + L[q.vb]*d[q.vb]; wB = L[PRS]*d[PRS] + L[q.bt]*d[q.bt] + L[q.bb]*d[q.bb]; }
        else        { wv = L[q.vn]*d[q.vn] + L[q.vt]*d[q.vt];                   wB = L[PRS]*d[PRS] + L[q.bt]*d[q.bt]; }
        w[KFASTM] =  wv + wB;
        w[KFASTP] = -wv + wB;
        L = LL[KENTRP];
        w[KENTRP] = L[RHO]*d[RHO] + L[PRS]*d[PRS];
        w[KDIVB] = 0.0;
        L = LL[KSLOWM];
        if (nc == 3){ wv = L[q.vn]*d[q.vn] + L[q.vt]*d[q.vt] + L[q.vb]*d[q.vb]; wB = L[PRS]*d[PRS] + L[q.bt]*d[q.bt] + L[q.bb]*d[q.bb]; }
        else        { wv = L[q.vn]*d[q.vn] + L[q.vt]*d[q.vt];                   wB = L[PRS]*d[PRS] + L[q.bt]*d[q.bt]; }
        w[KSLOWM] =  wv + wB;
        w[KSLOWP] = -wv + wB;
        if (nc == 3){
          L = LL[KALFVM];
          wv = L[q.vt]*d[q.vt] + L[q.vb]*d[q.vb];
          wB = L[q.bt]*d[q.bt] + L[q.bb]*d[q.bb];
          w[KALFVM] = wv + wB;
          w[KALFVP] = wv - wB;
        }
      }
    }
    /* 2b. limiter on the characteristic differences */
    for (k = nw; k--;   ){
      if (o->c.limiter == ORC_LIM_DEFAULT){              /* SET_GM_LIMITER (plm_coeffs.h:96-100) with cpk = cmk = kstp[k] */
        if (dwp[k]*dwm[k] > 0.0){
          double qc = 0.5*(dwm[k] + dwp[k]), scrh = ABS_MIN(dwp[k]*kstp[k], dwm[k]*kstp[k]);
          dw_lim[k] = ABS_MIN(qc, scrh);
        }else dw_lim[k] = 0.0;
      }else dw_lim[k] = single_limiter (o->c.limiter, dwp[k], dwm[k]);
    }
    /* 2c. back to primitive slopes, monotonicity in the primitive variables as well */
    for (nv = NV; nv--;   ){
      double dc = 0.0, d2v;
      if (nc == 2 && (nv == VX3 || nv == BX3)){ dv_lim[nv] = 0.0; continue; }
      for (k = 0; k < nw; k++) dc += dw_lim[k]*RR[nv][k];
      if (dvp[nv]*dvm[nv] > 0.0){
        d2v = ABS_MIN(2.0*dvp[nv], 2.0*dvm[nv]);
        dv_lim[nv] = MINMOD(d2v, dc);
      }else dv_lim[nv] = 0.0;
    }
    for (nv = NV; nv--;   ){
      vp[i][nv] = v[i][nv] + dv_lim[nv]*0.5;
      vm[i][nv] = v[i][nv] - dv_lim[nv]*0.5;
    }
  }
  for (i = beg-1; i <= end; i++) vp[i][q.bn] = vm[i+1][q.bn] = o->bn[i];
}

static void states_ppm (Oracle *o, int beg, int end, int bxn);   /* below */

/* =====================================================================
   Physics kernels shared by the Riemann solvers
   ===================================================================== */

static Dirs set_vector_indices (int dir)       /* set_indexes.c:49-123 */
{
  Dirs q;
  if (dir == 0){ q.vn = VX1; q.vt = VX2; q.vb = VX3; q.bn = BX1; q.bt = BX2; q.bb = BX3; }
  else if (dir == 1){ q.vn = VX2; q.vt = VX1; q.vb = VX3; q.bn = BX2; q.bt = BX1; q.bb = BX3; }
  else { q.vn = VX3; q.vt = VX1; q.vb = VX2; q.bn = BX3; q.bt = BX1; q.bb = BX2; }
  return q;
}

static void mhd_flux (const double *v, const double *u, Dirs q, double *fx, double *prs)
/* MHD/fluxes.c:159-215 */
{
  double Bmag2, ptot, vB;
  Bmag2 = v[BX1]*v[BX1] + v[BX2]*v[BX2] + v[BX3]*v[BX3];
  ptot  = v[PRS] + 0.5*Bmag2;
  vB    = v[VX1]*v[BX1] + v[VX2]*v[BX2] + v[VX3]*v[BX3];
  fx[RHO]   = u[q.vn];
  fx[MX1]   = v[q.vn]*u[MX1]   - v[q.bn]*v[BX1];
  fx[MX1+1] = v[q.vn]*u[MX1+1] - v[q.bn]*v[BX2];
  fx[MX1+2] = v[q.vn]*u[MX1+2] - v[q.bn]*v[BX3];
  fx[q.bn] = 0.0;
  fx[q.bt] = v[q.vn]*v[q.bt] - v[q.bn]*v[q.vt];
  fx[q.bb] = v[q.vn]*v[q.bb] - v[q.bn]*v[q.vb];
  fx[ENG]  = (u[ENG] + ptot)*v[q.vn] - v[q.bn]*vB;
  *prs = ptot;
}

static void max_signal_speed (const Oracle *o, const double *v, Dirs q,
                              double *cmin, double *cmax)
/* MHD/eigenv.c:35-104 (EOS IDEAL: gpr = g_gamma*p) */
{
  double gpr, b1, b2, b3, Btmag2, Bmag2, cf;
  gpr = o->c.gamma*v[PRS];
  b1 = v[q.bn]; b2 = v[q.bt]; b3 = v[q.bb];
  Btmag2 = b2*b2 + b3*b3;
  Bmag2  = b1*b1 + Btmag2;
  cf = gpr - Bmag2;
  cf = gpr + Bmag2 + sqrt(cf*cf + 4.0*gpr*Btmag2);
  cf = sqrt(0.5*cf/v[RHO]);
  *cmin = v[q.vn] - cf;
  *cmax = v[q.vn] + cf;
}

static void hll_speed (Oracle *o, const double *vL, const double *vR, Dirs q,
                       double a2L, double a2R, double *SL, double *SR)
/* MHD/hll_speed.c:76-107 (DAVIS_ESTIMATE) */
{
  double slmin, slmax, srmin, srmax, scrh;
  max_signal_speed (o, vL, q, &slmin, &slmax);
  max_signal_speed (o, vR, q, &srmin, &srmax);
  *SL = MINV(slmin, srmin);
  *SR = MAXV(slmax, srmax);
  o->cur_SL = *SL; o->cur_SR = *SR;      /* sweep->SL[i], SR[i] (hll_speed.c:92-93) */
  scrh  = fabs(vL[q.vn]) + fabs(vR[q.vn]);
  scrh /= sqrt(a2L) + sqrt(a2R);
  o->max_mach = MAXV(scrh, o->max_mach);
}

/* =====================================================================
   Riemann solvers: one interface at a time.
   in: vL,vR,uL,uR; out: flux[NV], press, cmax
   ===================================================================== */

static void riemann_hll (Oracle *o, const double *vL, const double *vR,
                         const double *uL, const double *uR, Dirs q,
                         double *flux, double *press, double *cmax)
/* MHD/hll.c:30-135 */
{
  double fL[NV], fR[NV], pL, pR, a2L, a2R, SL, SR, scrh;
  int nv;
  a2L = o->c.gamma*vL[PRS]/vL[RHO];                 /* EOS/Ideal/eos.c:33 */
  a2R = o->c.gamma*vR[PRS]/vR[RHO];
  mhd_flux (vL, uL, q, fL, &pL);
  mhd_flux (vR, uR, q, fR, &pR);
  hll_speed (o, vL, vR, q, a2L, a2R, &SL, &SR);
  scrh = MAXV(fabs(SL), fabs(SR));
  *cmax = scrh;
  if (SL > 0.0){
    for (nv = 0; nv < NV; nv++) flux[nv] = fL[nv];
    *press = pL;
  }else if (SR < 0.0){
    for (nv = 0; nv < NV; nv++) flux[nv] = fR[nv];
    *press = pR;
  }else{
    scrh = 1.0/(SR - SL);
    for (nv = 0; nv < NV; nv++){
      flux[nv]  = SL*SR*(uR[nv] - uL[nv]) + SR*fL[nv] - SL*fR[nv];
      flux[nv] *= scrh;
    }
    *press = (SR*pL - SL*pR)*scrh;
  }
}

static void riemann_tvdlf (Oracle *o, const double *vL, const double *vR,
                           const double *uL, const double *uR, Dirs q,
                           double *flux, double *press, double *cmax)
/* MHD/tvdlf.c:51-135 (Lax-Friedrichs / Rusanov; EOS IDEAL) */
{
  double fL[NV], fR[NV], vRL[NV], pL, pR, cminRL, cmaxRL, cRL, scrh;
  int nv;
  mhd_flux (vL, uL, q, fL, &pL);
  mhd_flux (vR, uR, q, fR, &pR);
  for (nv = 0; nv < NV; nv++) vRL[nv] = 0.5*(vL[nv] + vR[nv]);                       /* :102 */
  scrh = fabs(vRL[q.vn])/sqrt(o->c.gamma*vRL[PRS]/vRL[RHO]);                          /* :104-105 */
  o->max_mach = MAXV(o->max_mach, scrh);
  max_signal_speed (o, vRL, q, &cminRL, &cmaxRL);                                     /* :119 */
  cRL = MAXV(fabs(cminRL), fabs(cmaxRL));
  o->cur_SL = -cRL; o->cur_SR = cRL;                                                  /* :123-124 */
  *cmax = cRL;
  for (nv = 0; nv < NV; nv++) flux[nv] = 0.5*(fL[nv] + fR[nv] - cRL*(uR[nv] - uL[nv]));
  *press = 0.5*(pL + pR);
}

static void riemann_hllc (Oracle *o, const double *vL, const double *vR,
                          const double *uL, const double *uR, Dirs q,
                          double *flux, double *press, double *cmax)
/* MHD/hllc.c:42-236 (EOS IDEAL; Li 2005 / Gurski 2004) */
{
  double fL[NV], fR[NV], Uhll[NV], Fhll[NV], usl[NV], usr[NV];
  double pL, pR, a2L, a2R, SL, SR, scrh;
  double pl, pr, vBl, vBr, vxl, vxr, Bxs, Bys, Bzs, vxs, ps, vBs;
  int nv, mxn = MX1 + (q.vn - VX1), mxt = MX1 + (q.vt - VX1), mxb = MX1 + (q.vb - VX1);
  a2L = o->c.gamma*vL[PRS]/vL[RHO];
  a2R = o->c.gamma*vR[PRS]/vR[RHO];
  mhd_flux (vL, uL, q, fL, &pL);
  mhd_flux (vR, uR, q, fR, &pR);
  hll_speed (o, vL, vR, q, a2L, a2R, &SL, &SR);
  scrh = MAXV(fabs(SL), fabs(SR));
  *cmax = scrh;
  if (SL >= 0.0){                                         /* :106-111 */
    for (nv = 0; nv < NV; nv++) flux[nv] = fL[nv];
    *press = pL;
    return;
  }
  if (SR <= 0.0){                                         /* :113-118 */
    for (nv = 0; nv < NV; nv++) flux[nv] = fR[nv];
    *press = pR;
    return;
  }
  scrh = 1.0/(SR - SL);                                   /* :127-138 */
  for (nv = 0; nv < NV; nv++){
    Uhll[nv]  = SR*uR[nv] - SL*uL[nv] + fL[nv] - fR[nv];
    Uhll[nv] *= scrh;
    Fhll[nv]  = SL*SR*(uR[nv] - uL[nv]) + SR*fL[nv] - SL*fR[nv];
    Fhll[nv] *= scrh;
  }
  Uhll[mxn] += (pL - pR)*scrh;
  Fhll[mxn] += (SR*pL - SL*pR)*scrh;
  if (o->use_hll){                                        /* SHOCK_FLATTENING MULTID, :140-150 */
    for (nv = 0; nv < NV; nv++){
      flux[nv]  = SL*SR*(uR[nv] - uL[nv]) + SR*fL[nv] - SL*fR[nv];
      flux[nv] *= scrh;
    }
    *press = (SR*pL - SL*pR)*scrh;
    return;
  }
  pl = vL[BX1]*vL[BX1] + vL[BX2]*vL[BX2] + vL[BX3]*vL[BX3];       /* :154-164 */
  pr = vR[BX1]*vR[BX1] + vR[BX2]*vR[BX2] + vR[BX3]*vR[BX3];
  pl = vL[PRS] + 0.5*pl;
  pr = vR[PRS] + 0.5*pr;
  vBl = vL[VX1]*vL[BX1] + vL[VX2]*vL[BX2] + vL[VX3]*vL[BX3];
  vBr = vR[VX1]*vR[BX1] + vR[VX2]*vR[BX2] + vR[VX3]*vR[BX3];
  vxl = vL[q.vn];
  vxr = vR[q.vn];
  Bxs = Uhll[q.bn]; Bys = Uhll[q.bt]; Bzs = Uhll[q.bb];           /* :168-170 */
  vxs = Uhll[mxn]/Uhll[RHO];                                       /* :174-175 */
  ps  = Fhll[mxn] + Bxs*Bxs - Fhll[RHO]*vxs;
  vBs = Uhll[BX1]*Uhll[MX1] + Uhll[BX2]*Uhll[MX1+1] + Uhll[BX3]*Uhll[MX1+2];   /* :179-183 */
  vBs /= Uhll[RHO];
  usl[RHO] = uL[RHO]*(SL - vxl)/(SL - vxs);                        /* :185-191 */
  usr[RHO] = uR[RHO]*(SR - vxr)/(SR - vxs);
  usl[ENG] = (uL[ENG]*(SL - vxl) + ps*vxs - pl*vxl - Bxs*vBs + vL[q.bn]*vBl)/(SL - vxs);
  usr[ENG] = (uR[ENG]*(SR - vxr) + ps*vxs - pr*vxr - Bxs*vBs + vR[q.bn]*vBr)/(SR - vxs);
  usl[mxn] = usl[RHO]*vxs;                                         /* :193-205 */
  usr[mxn] = usr[RHO]*vxs;
  usl[mxt] = (uL[mxt]*(SL - vxl) - (Bxs*Bys - vL[q.bn]*vL[q.bt]))/(SL - vxs);
  usr[mxt] = (uR[mxt]*(SR - vxr) - (Bxs*Bys - vR[q.bn]*vR[q.bt]))/(SR - vxs);
  usl[mxb] = (uL[mxb]*(SL - vxl) - (Bxs*Bzs - vL[q.bn]*vL[q.bb]))/(SL - vxs);
  usr[mxb] = (uR[mxb]*(SR - vxr) - (Bxs*Bzs - vR[q.bn]*vR[q.bb]))/(SR - vxs);
  usl[q.bn] = usr[q.bn] = Bxs;                                     /* :207-209 */
  usl[q.bt] = usr[q.bt] = Bys;
  usl[q.bb] = usr[q.bb] = Bzs;
  if (vxs >= 0.0){                                                 /* :215-225 */
    for (nv = 0; nv < NV; nv++) flux[nv] = fL[nv] + SL*(usl[nv] - uL[nv]);
    *press = pL;
  }else{
    for (nv = 0; nv < NV; nv++) flux[nv] = fR[nv] + SR*(usr[nv] - uR[nv]);
    *press = pR;
  }
}

static void riemann_hlld (Oracle *o, const double *vL, const double *vR,
                          const double *uL, const double *uR, Dirs q,
                          double *flux, double *press, double *cmax)
/* MHD/hlld.c:44-452 (EOS IDEAL, no background field) */
{
  double fL[NV], fR[NV], ptL, ptR, a2L, a2R, SL, SR, scrh;
  double usL[NV], usR[NV], ussl[NV], ussr[NV], Uhll[NV];
  double vsL, wsL, scrhL, S1L, sqrL, duL;
  double vsR, wsR, scrhR, S1R, sqrR, duR;
  double Bx, Bx1, SM, sBx, pts, vss, wss;
  int nv, revert_to_hllc;
  int VXn = q.vn, VXt = q.vt, VXb = q.vb, BXn = q.bn, BXt = q.bt, BXb = q.bb;
  int MXn = VXn, MXt = VXt, MXb = VXb;

  a2L = o->c.gamma*vL[PRS]/vL[RHO];
  a2R = o->c.gamma*vR[PRS]/vR[RHO];
  mhd_flux (vL, uL, q, fL, &ptL);
  mhd_flux (vR, uR, q, fR, &ptR);
  hll_speed (o, vL, vR, q, a2L, a2R, &SL, &SR);

  scrh  = MAXV(fabs(SL), fabs(SR));
  *cmax = scrh;

  if (SL >= 0.0){
    for (nv = 0; nv < NV; nv++) flux[nv] = fL[nv];
    *press = ptL;
    return;
  }else if (SR <= 0.0){
    for (nv = 0; nv < NV; nv++) flux[nv] = fR[nv];
    *press = ptR;
    return;
  }

  if (o->use_hll){                       /* hlld.c:149-160 */
    scrh = 1.0/(SR - SL);
    for (nv = 0; nv < NV; nv++){
      flux[nv]  = SR*SL*(uR[nv] - uL[nv]) + SR*fL[nv] - SL*fR[nv];
      flux[nv] *= scrh;
    }
    *press = (SR*ptL - SL*ptR)*scrh;
    return;
  }

  for (nv = 0; nv < NV; nv++) usL[nv] = usR[nv] = 0.0;

  scrh = 1.0/(SR - SL);
  Bx1  = Bx = (SR*vR[BXn] - SL*vL[BXn])*scrh;
  sBx  = (Bx > 0.0 ? 1.0 : -1.0);

  duL = SL - vL[VXn];
  duR = SR - vR[VXn];

  scrh = 1.0/(duR*uR[RHO] - duL*uL[RHO]);
  SM   = (duR*uR[MXn] - duL*uL[MXn] - ptR + ptL)*scrh;

  pts  = duR*uR[RHO]*ptL - duL*uL[RHO]*ptR +
         vL[RHO]*vR[RHO]*duR*duL*(vR[VXn] - vL[VXn]);
  pts *= scrh;

  usL[RHO] = uL[RHO]*duL/(SL - SM);
  usR[RHO] = uR[RHO]*duR/(SR - SM);

  sqrL = sqrt(usL[RHO]);
  sqrR = sqrt(usR[RHO]);

  S1L = SM - fabs(Bx)/sqrL;
  S1R = SM + fabs(Bx)/sqrR;

  revert_to_hllc = 0;
  if ( (S1L - SL) <  1.e-4*(SM - SL) ) revert_to_hllc = 1;
  if ( (S1R - SR) > -1.e-4*(SR - SM) ) revert_to_hllc = 1;

  if (revert_to_hllc){
    scrh = 1.0/(SR - SL);
    for (nv = 0; nv < NV; nv++){
      Uhll[nv]  = SR*uR[nv] - SL*uL[nv] + fL[nv] - fR[nv];
      Uhll[nv] *= scrh;
    }
    usL[BXn] = usR[BXn] = Uhll[BXn];
    usL[BXt] = usR[BXt] = Uhll[BXt];
    usL[BXb] = usR[BXb] = Uhll[BXb];
    S1L = S1R = SM;
  }else{
    scrhL = (uL[RHO]*duL*duL - Bx*Bx)/(uL[RHO]*duL*(SL - SM) - Bx*Bx);
    scrhR = (uR[RHO]*duR*duR - Bx*Bx)/(uR[RHO]*duR*(SR - SM) - Bx*Bx);
    usL[BXn] = Bx1;
    usL[BXt] = uL[BXt]*scrhL;
    usL[BXb] = uL[BXb]*scrhL;
    usR[BXn] = Bx1;
    usR[BXt] = uR[BXt]*scrhR;
    usR[BXb] = uR[BXb]*scrhR;
  }

  scrhL = Bx/(uL[RHO]*duL);
  scrhR = Bx/(uR[RHO]*duR);

  vsL = vL[VXt] - scrhL*(usL[BXt] - uL[BXt]);
  vsR = vR[VXt] - scrhR*(usR[BXt] - uR[BXt]);
  wsL = vL[VXb] - scrhL*(usL[BXb] - uL[BXb]);
  wsR = vR[VXb] - scrhR*(usR[BXb] - uR[BXb]);

  usL[MXn] = usL[RHO]*SM;
  usR[MXn] = usR[RHO]*SM;
  usL[MXt] = usL[RHO]*vsL;
  usR[MXt] = usR[RHO]*vsR;
  usL[MXb] = usL[RHO]*wsL;
  usR[MXb] = usR[RHO]*wsR;

  scrhL  = vL[VXn]*Bx1 + vL[VXt]*uL[BXt] + vL[VXb]*uL[BXb];
  scrhL -=      SM*Bx1 +    vsL*usL[BXt] +    wsL*usL[BXb];
  usL[ENG]  = duL*uL[ENG] - ptL*vL[VXn] + pts*SM + Bx*scrhL;
  usL[ENG] /= SL - SM;

  scrhR  = vR[VXn]*Bx1 + vR[VXt]*uR[BXt] + vR[VXb]*uR[BXb];
  scrhR -=      SM*Bx1 +    vsR*usR[BXt] +    wsR*usR[BXb];
  usR[ENG]  = duR*uR[ENG] - ptR*vR[VXn] + pts*SM + Bx*scrhR;
  usR[ENG] /= SR - SM;

  if (S1L >= 0.0){
    for (nv = 0; nv < NV; nv++) flux[nv] = fL[nv] + SL*(usL[nv] - uL[nv]);
    *press = ptL;
  }else if (S1R <= 0.0){
    for (nv = 0; nv < NV; nv++) flux[nv] = fR[nv] + SR*(usR[nv] - uR[nv]);
    *press = ptR;
  }else{
    ussl[RHO] = usL[RHO];
    ussr[RHO] = usR[RHO];

    vss  = sqrL*vsL + sqrR*vsR + (usR[BXt] - usL[BXt])*sBx;
    vss /= sqrL + sqrR;
    wss  = sqrL*wsL + sqrR*wsR + (usR[BXb] - usL[BXb])*sBx;
    wss /= sqrL + sqrR;

    ussl[MXn] = ussl[RHO]*SM;
    ussr[MXn] = ussr[RHO]*SM;
    ussl[MXt] = ussl[RHO]*vss;
    ussr[MXt] = ussr[RHO]*vss;
    ussl[MXb] = ussl[RHO]*wss;
    ussr[MXb] = ussr[RHO]*wss;

    ussl[BXn] = ussr[BXn] = Bx1;
    ussl[BXt]  = sqrL*usR[BXt] + sqrR*usL[BXt] + sqrL*sqrR*(vsR - vsL)*sBx;
    ussl[BXt] /= sqrL + sqrR;
    ussr[BXt]  = ussl[BXt];
    ussl[BXb]  = sqrL*usR[BXb] + sqrR*usL[BXb] + sqrL*sqrR*(wsR - wsL)*sBx;
    ussl[BXb] /= sqrL + sqrR;
    ussr[BXb]  = ussl[BXb];

    scrhL  = SM*Bx1 + vsL*usL [BXt] + wsL*usL [BXb];
    scrhL -= SM*Bx1 + vss*ussl[BXt] + wss*ussl[BXb];
    scrhR  = SM*Bx1 + vsR*usR [BXt] + wsR*usR [BXb];
    scrhR -= SM*Bx1 + vss*ussr[BXt] + wss*ussr[BXb];

    ussl[ENG] = usL[ENG] - sqrL*scrhL*sBx;
    ussr[ENG] = usR[ENG] + sqrR*scrhR*sBx;

    if (SM >= 0.0){
      for (nv = 0; nv < NV; nv++)
        flux[nv] = fL[nv] + S1L*(ussl[nv] - usL[nv]) + SL*(usL[nv] - uL[nv]);
      *press = ptL;
    }else{
      for (nv = 0; nv < NV; nv++)
        flux[nv] = fR[nv] + S1R*(ussr[nv] - usR[nv]) + SR*(usR[nv] - uR[nv]);
      *press = ptR;
    }
  }
}

static void riemann_roe (Oracle *o, const double *vL, const double *vR,
                         const double *uL, const double *uR, Dirs q,
                         double *flux, double *press, double *cmax);  /* below */

/* =====================================================================
   UpdateStage  (reference Time_Stepping/update_stage.c:37-314)
   ===================================================================== */

#define EPS_UCT_CONTACT 1.e-6      /* MHD/CT/ct_emf.c:102 */

static void ct_compute_emf (Oracle *o);
static void ct_update (Oracle *o, double dt);
static void ct_update_from (Oracle *o, double *const *Bs, double dt);
static void ct_average_magnetic_field (Oracle *o);
static void cons_to_prim_3d (Oracle *o);
static void flag_shock (Oracle *o);

static void update_stage (Oracle *o, double dt)
{
  int dir, dims = o->c.dims, nv;
  int i, j, k;

  if (o->stage == 1) memset (o->C_dt, 0, sizeof(double)*(size_t)o->tot);  /* :86-90 */

  for (dir = 0; dir < dims; dir++){
    Dirs q = set_vector_indices (dir);
    int lo[3], hi[3], t1, t2, tb, te, bb, be, a, b, n;
    int nbeg = o->beg[dir], nend = o->end[dir], ntot = o->T[dir];
    for (a = 0; a < 3; a++){ lo[a] = o->beg[a]; hi[a] = o->end[a]; }
    /* transverse extension by one zone for CT, update_stage.c:144-148 */
    for (a = 0; a < dims; a++) if (a != dir){ lo[a]--; hi[a]++; }
    /* transverse loop order: BOX_TRANSVERSE_LOOP (macros.h:107-112):
       IDIR: t=j,b=k ; JDIR: t=i,b=k ; KDIR: t=i,b=j  (b outer, t inner) */
    if (dir == 0){ t1 = 1; t2 = 2; } else if (dir == 1){ t1 = 0; t2 = 2; } else { t1 = 0; t2 = 1; }
    tb = lo[t1]; te = hi[t1]; bb = lo[t2]; be = hi[t2];

    for (b = bb; b <= be; b++) for (a = tb; a <= te; a++){
      int idx3[3];
      idx3[t1] = a; idx3[t2] = b;
      /* gather pencil, update_stage.c:158-164 */
      for (n = 0; n < ntot; n++){
        int id;
        idx3[dir] = n;
        id = IDX(o, idx3[2], idx3[1], idx3[0]);
        for (nv = 0; nv < NV; nv++) o->v[n][nv] = o->Vc[nv][id];
        o->bn[n] = o->Vs[dir][id];
        o->pflag[n] = o->flag[id];
      }
      /* States (nbeg-1 .. nend+1), :193 */
      if (o->c.recon == ORC_RECON_PLM && o->c.char_limiting) states_plm_char (o, nbeg-1, nend+1, q);
      else if (o->c.recon == ORC_RECON_PLM) states_plm (o, nbeg-1, nend+1, q.bn);
      else                             states_ppm (o, nbeg-1, nend+1, q.bn);

      /* Riemann (nbeg-1 .. nend), :194 ; stateL = vp[n], stateR = vm[n+1] */
      for (n = nbeg-1; n <= nend; n++){
        double uL[NV], uR[NV];
        prim_to_cons (o, o->vp[n],   uL);          /* plm_states.c:310-311 */
        prim_to_cons (o, o->vm[n+1], uR);
        /* SHOCK_FLATTENING MULTID: HLL flux where either zone lies in a shock (hlld.c:149-160, roe.c:165-189) */
        o->use_hll = ((o->pflag[n] & 4) || (o->pflag[n+1] & 4));
        if      (o->c.solver == ORC_SOLVER_HLLD)
          riemann_hlld (o, o->vp[n], o->vm[n+1], uL, uR, q, o->flux[n], &o->press[n], &o->cmax[n]);
        else if (o->c.solver == ORC_SOLVER_HLLC)
          riemann_hllc (o, o->vp[n], o->vm[n+1], uL, uR, q, o->flux[n], &o->press[n], &o->cmax[n]);
        else if (o->c.solver == ORC_SOLVER_TVDLF)
          riemann_tvdlf (o, o->vp[n], o->vm[n+1], uL, uR, q, o->flux[n], &o->press[n], &o->cmax[n]);
        else if (o->c.solver == ORC_SOLVER_HLL || o->use_hll)
          riemann_hll  (o, o->vp[n], o->vm[n+1], uL, uR, q, o->flux[n], &o->press[n], &o->cmax[n]);
        else
          riemann_roe  (o, o->vp[n], o->vm[n+1], uL, uR, q, o->flux[n], &o->press[n], &o->cmax[n]);
        o->SLp[n] = o->cur_SL; o->SRp[n] = o->cur_SR;
      }

      /* CT_StoreUpwindEMF (nbeg-1 .. nend), ct_emf.c:104-190 */
      for (n = nbeg-1; n <= nend; n++){
        int id; signed char s;
        idx3[dir] = n;
        id = IDX(o, idx3[2], idx3[1], idx3[0]);
        if (o->c.emf_average == ORC_EMF_UCT_HLL){
          o->SfL[dir][id] = MAXV(0.0, -o->SLp[n]);
          o->SfR[dir][id] = MAXV(0.0,  o->SRp[n]);
        }
        if      (o->flux[n][RHO] >  EPS_UCT_CONTACT) s = 1;
        else if (o->flux[n][RHO] < -EPS_UCT_CONTACT) s = -1;
        else s = 0;
        if (dir == 0){
          o->ezi[id] = -o->flux[n][BX2];
          if (dims == 3) o->eyi[id] = o->flux[n][BX3];
          o->svx[id] = s;
        }else if (dir == 1){
          o->ezj[id] = o->flux[n][BX1];
          if (dims == 3) o->exj[id] = -o->flux[n][BX3];
          o->svy[id] = s;
        }else{
          o->eyk[id] = -o->flux[n][BX1];
          o->exk[id] =  o->flux[n][BX2];
          o->svz[id] = s;
        }
      }

      if (o->c.emf_average == ORC_EMF_UCT_HLL){
        /* CT_StoreVelSlopes (beg = nbeg-1 .. end+1 = nend+1), ct_stag_slopes.c:5-45 */
        for (n = nbeg-1; n <= nend+1; n++){
          int id, c;
          idx3[dir] = n;
          id = IDX(o, idx3[2], idx3[1], idx3[0]);
          for (c = 0; c < dims; c++) o->dvel[c][dir][id] = o->vp[n][VX1+c] - o->vm[n][VX1+c];
        }
      }

      if (o->phic){                    /* TotalFlux, rhs.c:388-392: gravitational energy flux */
        for (n = nbeg-1; n <= nend; n++){
          idx3[dir] = n;
          o->ppen[n] = o->phif[dir][IDX(o, idx3[2], idx3[1], idx3[0])];
          o->flux[n][ENG] += o->flux[n][RHO]*o->ppen[n];
        }
      }
      /* RightHandSide + update (nbeg .. nend), rhs.c:193-201,
         update_stage.c:214-216, 229-235 */
      for (n = nbeg; n <= nend; n++){
        int id;
        double rhs[NV];
        idx3[dir] = n;
        id = IDX(o, idx3[2], idx3[1], idx3[0]);
        const double dtdx   = dt/DXA(o,dir,n);    /* rhs.c:195  scrh = dt/dx[i] */
        const double inv_dl = 1.0/DXA(o,dir,n);   /* set_geometry.c: inv_dx     */
        for (nv = 0; nv < NV; nv++) rhs[nv] = -dtdx*(o->flux[n][nv] - o->flux[n-1][nv]);
        rhs[q.vn] -= dtdx*(o->press[n] - o->press[n-1]);
        if (o->c.body_force){          /* RightHandSideSource, rhs_source.c:214-217 (x1), :277-280 (x2), :342-345 (x3) */
          const double g = (o->gf[dir] ? o->gf[dir][id] : o->c.grav[dir]);
          rhs[q.vn] += dt*o->v[n][RHO]*g;
          rhs[ENG]  += dt*0.5*(o->flux[n][RHO] + o->flux[n-1][RHO])*g;
        }
        if (o->phic){                  /* rhs_source.c:233-237, 316-320, 358-362 */
          rhs[q.vn] -= dtdx*o->v[n][RHO]*(o->ppen[n] - o->ppen[n-1]);
          rhs[ENG]  -= o->phic[id]*rhs[RHO];
        }
        for (nv = 0; nv < NV; nv++) o->Uc[nv][id] += rhs[nv];
        if (o->stage == 1)
          o->C_dt[id] += 0.5*(o->cmax[n-1] + o->cmax[n])*inv_dl;
      }
    }
  }
  /* emf index ranges as left by the last sweeps, ct_emf.c:130,152,173 */
  o->emf_ibeg = o->beg[0]-1; o->emf_iend = o->end[0];
  o->emf_jbeg = o->beg[1]-1; o->emf_jend = o->end[1];
  if (dims == 3){ o->emf_kbeg = o->beg[2]-1; o->emf_kend = o->end[2]; }
  else          { o->emf_kbeg = o->emf_kend = 0; }

  ct_compute_emf (o);                                   /* :250 */
  ct_update (o, dt);                                    /* :284 */

  if (o->stage == 1){                                   /* :308-312 */
    for (k = o->beg[2]; k <= o->end[2]; k++)
    for (j = o->beg[1]; j <= o->end[1]; j++)
    for (i = o->beg[0]; i <= o->end[0]; i++)
      o->inv_dt_hyp = MAXV(o->inv_dt_hyp, o->C_dt[IDX(o,k,j,i)]);
    o->inv_dt_hyp /= (double)dims;
  }
}

/* =====================================================================
   Constrained transport
   ===================================================================== */

static double mc_lim2 (double dp, double dm)      /* ct_stag_slopes.c:97-108 */
{
  double dc, scrh;
  if (dp*dm < 0.0) return 0.0;
  dc   = 0.5*(dp + dm);
  scrh = 2.0*(fabs(dp) < fabs(dm) ? dp : dm);
  return (fabs(dc) < fabs(scrh) ? dc : scrh);
}

static void ct_emf_hll (Oracle *o)
/* UCT_HLL: CT_GetStagSlopes (ct_stag_slopes.c:52-95) + CT_EMF_HLL_Solver
   (ct_emf_average.c:180-359, Londrillo & Del Zanna 2004 eq. 56).  The staggered slopes
   are evaluated where they are used instead of being stored.                       */
{
  int i, j, k, dims = o->c.dims;
  int ibeg = o->emf_ibeg, iend = o->emf_iend, jbeg = o->emf_jbeg, jend = o->emf_jend;
  int kbeg = o->emf_kbeg, kend = o->emf_kend;
  const int st[3] = {1, o->S1, o->S1*o->S2};
#define I3(k,j,i) IDX(o,k,j,i)
  /* limited slope of staggered component c along direction d at zone id */
#define DB(c, d, id) mc_lim2 (o->Vs[c][(id) + st[d]] - o->Vs[c][id], o->Vs[c][id] - o->Vs[c][(id) - st[d]])
#define BP(c, d, id) (o->Vs[c][id] + 0.5*DB(c, d, id))
#define BM(c, d, id) (o->Vs[c][id] - 0.5*DB(c, d, id))
  /* velocity component c integrated from the centre of zone id to the edge, transverse
     directions (x, y) of the edge: dPP, dPM, dMM, dMP of ct_emf_average.c:196-206 */
#define VPP(c, x, y, id) (o->Vc[VX1+(c)][id] + 0.5*(o->dvel[c][x][id] + o->dvel[c][y][id]))
#define VPM(c, x, y, id) (o->Vc[VX1+(c)][id] + 0.5*(o->dvel[c][x][id] - o->dvel[c][y][id]))
#define VMM(c, x, y, id) (o->Vc[VX1+(c)][id] - 0.5*(o->dvel[c][x][id] + o->dvel[c][y][id]))
#define VMP(c, x, y, id) (o->Vc[VX1+(c)][id] - 0.5*(o->dvel[c][x][id] - o->dvel[c][y][id]))
  for (k = kbeg; k <= kend; k++) for (j = jbeg; j <= jend; j++) for (i = ibeg; i <= iend; i++){
    int id = I3(k,j,i), ip = id + st[0], jp = id + st[1], kp = id + st[2];
    double a_xp, a_xm, a_yp, a_ym, bS, bN, bW, bE, eSW, eSE, eNE, eNW, e;
    /* ---- ez at (i+1/2, j+1/2, k) ---- */
    a_xp = MAXV(o->SfR[0][id], o->SfR[0][jp]); a_xm = MAXV(o->SfL[0][id], o->SfL[0][jp]);
    a_yp = MAXV(o->SfR[1][id], o->SfR[1][ip]); a_ym = MAXV(o->SfL[1][id], o->SfL[1][ip]);
    bS = BP(0, 1, id); bW = BP(1, 0, id);
    bN = BM(0, 1, jp); bE = BM(1, 0, ip);
    eSW = VPP(1,0,1,id)*bS - VPP(0,0,1,id)*bW;
    eSE = VMP(1,0,1,ip)*bS - VMP(0,0,1,ip)*bE;
    eNE = VMM(1,0,1,ip + st[1])*bN - VMM(0,0,1,ip + st[1])*bE;
    eNW = VPM(1,0,1,jp)*bN - VPM(0,0,1,jp)*bW;
    e  = a_xp*a_yp*eSW + a_xm*a_yp*eSE + a_xm*a_ym*eNE + a_xp*a_ym*eNW;
    e /= (a_xp + a_xm)*(a_yp + a_ym);
    e -= a_yp*a_ym*(bN - bS)/(a_yp + a_ym);
    e += a_xp*a_xm*(bE - bW)/(a_xp + a_xm);
    o->ez[id] = e;
    if (dims == 3){
      /* ---- ex at (i, j+1/2, k+1/2) ---- */
      a_xp = MAXV(o->SfR[1][id], o->SfR[1][kp]); a_xm = MAXV(o->SfL[1][id], o->SfL[1][kp]);
      a_yp = MAXV(o->SfR[2][id], o->SfR[2][jp]); a_ym = MAXV(o->SfL[2][id], o->SfL[2][jp]);
      bS = BP(1, 2, id); bW = BP(2, 1, id);
      bN = BM(1, 2, kp); bE = BM(2, 1, jp);
      eSW = VPP(2,1,2,id)*bS - VPP(1,1,2,id)*bW;
      eSE = VMP(2,1,2,jp)*bS - VMP(1,1,2,jp)*bE;
      eNE = VMM(2,1,2,jp + st[2])*bN - VMM(1,1,2,jp + st[2])*bE;
      eNW = VPM(2,1,2,kp)*bN - VPM(1,1,2,kp)*bW;
      e  = a_xp*a_yp*eSW + a_xm*a_yp*eSE + a_xm*a_ym*eNE + a_xp*a_ym*eNW;
      e /= (a_xp + a_xm)*(a_yp + a_ym);
      e -= a_yp*a_ym*(bN - bS)/(a_yp + a_ym);
      e += a_xp*a_xm*(bE - bW)/(a_xp + a_xm);
      o->ex[id] = e;
      /* ---- ey at (i+1/2, j, k+1/2) ---- */
      a_xp = MAXV(o->SfR[2][id], o->SfR[2][ip]); a_xm = MAXV(o->SfL[2][id], o->SfL[2][ip]);
      a_yp = MAXV(o->SfR[0][id], o->SfR[0][kp]); a_ym = MAXV(o->SfL[0][id], o->SfL[0][kp]);
      bS = BP(2, 0, id); bW = BP(0, 2, id);
      bN = BM(2, 0, ip); bE = BM(0, 2, kp);
      eSW = VPP(0,2,0,id)*bS - VPP(2,2,0,id)*bW;
      eSE = VMP(0,2,0,kp)*bS - VMP(2,2,0,kp)*bE;
      eNE = VMM(0,2,0,ip + st[2])*bN - VMM(2,2,0,ip + st[2])*bE;
      eNW = VPM(0,2,0,ip)*bN - VPM(2,2,0,ip)*bW;
      e  = a_xp*a_yp*eSW + a_xm*a_yp*eSE + a_xm*a_ym*eNE + a_xp*a_ym*eNW;
      e /= (a_xp + a_xm)*(a_yp + a_ym);
      e -= a_yp*a_ym*(bN - bS)/(a_yp + a_ym);
      e += a_xp*a_xm*(bE - bW)/(a_xp + a_xm);
      o->ey[id] = e;
    }
  }
#undef I3
#undef DB
#undef BP
#undef BM
#undef VPP
#undef VPM
#undef VMM
#undef VMP
}

static void ct_compute_emf (Oracle *o)
/* MHD/CT/ct_emf.c:210-254 (UCT_CONTACT): CT_ComputeCenterEMF (:348-388),
   CT_EMF_ArithmeticAverage(w = 1) (ct_emf_average.c:13-49),
   CT_EMF_IntegrateToCorner (ct_emf_average.c:52-164, scatter form as in
   the reference, loop order k,j,i), then x0.25.                          */
{
  int i, j, k, dims = o->c.dims;
  int koff = (dims == 3 ? 1 : 0);
  int ibeg = o->emf_ibeg, iend = o->emf_iend, jbeg = o->emf_jbeg, jend = o->emf_jend;
  int kbeg = o->emf_kbeg, kend = o->emf_kend;
  double *exj = o->exj, *exk = o->exk, *eyi = o->eyi, *eyk = o->eyk, *ezi = o->ezi, *ezj = o->ezj;
  double *ex = o->ex, *ey = o->ey, *ez = o->ez, *Ex1 = o->Ex1, *Ex2 = o->Ex2, *Ex3 = o->Ex3;

  for (k = 0; k < o->T[2]; k++) for (j = 0; j < o->T[1]; j++) for (i = 0; i < o->T[0]; i++){
    int id = IDX(o,k,j,i);
    double vx1 = o->Vc[VX1][id], vx2 = o->Vc[VX2][id], vx3 = o->Vc[VX3][id];
    double Bx1 = o->Vc[BX1][id], Bx2 = o->Vc[BX2][id], Bx3 = o->Vc[BX3][id];
    Ex1[id] = (vx3*Bx2 - vx2*Bx3);
    Ex2[id] = (vx1*Bx3 - vx3*Bx1);
    Ex3[id] = (vx2*Bx1 - vx1*Bx2);
  }

#define I3(k,j,i) IDX(o,k,j,i)
  if (o->c.emf_average == ORC_EMF_UCT_HLL){ ct_emf_hll (o); return; }
  if (o->c.emf_average != ORC_EMF_UCT_CONTACT){
    /* ARITHMETIC: CT_EMF_ArithmeticAverage (emf, 0.25) (ct_emf.c:241-243);
       UCT0: face EMFs <- 2 face - mean of the two adjacent cell-centred EMFs, over
       k..kend+KOFFSET etc., then the same average (ct_emf.c:261-283) */
    if (o->c.emf_average == ORC_EMF_UCT0){
      for (k = kbeg; k <= kend + koff; k++) for (j = jbeg; j <= jend + 1; j++) for (i = ibeg; i <= iend + 1; i++){
        if (dims == 3){
          exj[I3(k,j,i)] *= 2.0; exk[I3(k,j,i)] *= 2.0; eyi[I3(k,j,i)] *= 2.0; eyk[I3(k,j,i)] *= 2.0;
          exj[I3(k,j,i)] -= 0.5*(Ex1[I3(k,j,i)] + Ex1[I3(k,j+1,i)]);
          exk[I3(k,j,i)] -= 0.5*(Ex1[I3(k,j,i)] + Ex1[I3(k+1,j,i)]);
          eyi[I3(k,j,i)] -= 0.5*(Ex2[I3(k,j,i)] + Ex2[I3(k,j,i+1)]);
          eyk[I3(k,j,i)] -= 0.5*(Ex2[I3(k,j,i)] + Ex2[I3(k+1,j,i)]);
        }
        ezi[I3(k,j,i)] *= 2.0; ezj[I3(k,j,i)] *= 2.0;
        ezi[I3(k,j,i)] -= 0.5*(Ex3[I3(k,j,i)] + Ex3[I3(k,j,i+1)]);
        ezj[I3(k,j,i)] -= 0.5*(Ex3[I3(k,j,i)] + Ex3[I3(k,j+1,i)]);
      }
    }
    for (k = kbeg; k <= kend; k++) for (j = jbeg; j <= jend; j++) for (i = ibeg; i <= iend; i++){
      if (dims == 3){
        ex[I3(k,j,i)] = 0.25*(  exk[I3(k,j,i)] + exk[I3(k,j+1,i)]
                              + exj[I3(k,j,i)] + exj[I3(k+1,j,i)]);
        ey[I3(k,j,i)] = 0.25*(  eyi[I3(k,j,i)] + eyi[I3(k+1,j,i)]
                              + eyk[I3(k,j,i)] + eyk[I3(k,j,i+1)]);
      }
      ez[I3(k,j,i)] = 0.25*(  ezi[I3(k,j,i)] + ezi[I3(k,j+1,i)]
                            + ezj[I3(k,j,i)] + ezj[I3(k,j,i+1)]);
    }
    return;
  }
  for (k = kbeg; k <= kend; k++) for (j = jbeg; j <= jend; j++) for (i = ibeg; i <= iend; i++){
    if (dims == 3){
      ex[I3(k,j,i)] = 1.0*(  exk[I3(k,j,i)] + exk[I3(k,j+1,i)]
                           + exj[I3(k,j,i)] + exj[I3(k+1,j,i)]);
      ey[I3(k,j,i)] = 1.0*(  eyi[I3(k,j,i)] + eyi[I3(k+1,j,i)]
                           + eyk[I3(k,j,i)] + eyk[I3(k,j,i+1)]);
    }
    ez[I3(k,j,i)] = 1.0*(  ezi[I3(k,j,i)] + ezi[I3(k,j+1,i)]
                         + ezj[I3(k,j,i)] + ezj[I3(k,j,i+1)]);
  }

#define DEX_DYP(k,j,i) (exj[I3(k,j,i)] - Ex1[I3(k,j,i)])
#define DEX_DZP(k,j,i) (exk[I3(k,j,i)] - Ex1[I3(k,j,i)])
#define DEY_DXP(k,j,i) (eyi[I3(k,j,i)] - Ex2[I3(k,j,i)])
#define DEY_DZP(k,j,i) (eyk[I3(k,j,i)] - Ex2[I3(k,j,i)])
#define DEZ_DXP(k,j,i) (ezi[I3(k,j,i)] - Ex3[I3(k,j,i)])
#define DEZ_DYP(k,j,i) (ezj[I3(k,j,i)] - Ex3[I3(k,j,i)])
#define DEX_DYM(k,j,i) (Ex1[I3(k,j,i)] - exj[I3(k,j-1,i)])
#define DEX_DZM(k,j,i) (Ex1[I3(k,j,i)] - exk[I3(k-1,j,i)])
#define DEY_DXM(k,j,i) (Ex2[I3(k,j,i)] - eyi[I3(k,j,i-1)])
#define DEY_DZM(k,j,i) (Ex2[I3(k,j,i)] - eyk[I3(k-1,j,i)])
#define DEZ_DXM(k,j,i) (Ex3[I3(k,j,i)] - ezi[I3(k,j,i-1)])
#define DEZ_DYM(k,j,i) (Ex3[I3(k,j,i)] - ezj[I3(k,j-1,i)])

  /* CT_EMF_IntegrateToCorner returns at once in the CTU predictor (ct_emf_average.c:90-92) */
  for (k = kbeg; k <= (o->c.ctu && o->stage == 1 ? kbeg - 1 : kend + koff); k++)
  for (j = jbeg; j <= jend + 1; j++)
  for (i = ibeg; i <= iend + 1; i++){
    signed char sx = o->svx[I3(k,j,i)], sy = o->svy[I3(k,j,i)], sz = 0;
    int iu, ju, ku = k;
    if (dims == 3) sz = o->svz[I3(k,j,i)];
    iu = sx > 0 ? i : i+1;
    ju = sy > 0 ? j : j+1;
    if (dims == 3) ku = sz > 0 ? k : k+1;

    if (sx == 0){
      ez[I3(k,j,i)]   += 0.5*(DEZ_DYP(k,j,i) + DEZ_DYP(k,j,i+1));
      ez[I3(k,j-1,i)] -= 0.5*(DEZ_DYM(k,j,i) + DEZ_DYM(k,j,i+1));
      if (dims == 3){
        ey[I3(k,j,i)]   += 0.5*(DEY_DZP(k,j,i) + DEY_DZP(k,j,i+1));
        ey[I3(k-1,j,i)] -= 0.5*(DEY_DZM(k,j,i) + DEY_DZM(k,j,i+1));
      }
    }else{
      ez[I3(k,j,i)]   += DEZ_DYP(k,j,iu);
      ez[I3(k,j-1,i)] -= DEZ_DYM(k,j,iu);
      if (dims == 3){
        ey[I3(k,j,i)]   += DEY_DZP(k,j,iu);
        ey[I3(k-1,j,i)] -= DEY_DZM(k,j,iu);
      }
    }

    if (sy == 0){
      ez[I3(k,j,i)]   += 0.5*(DEZ_DXP(k,j,i) + DEZ_DXP(k,j+1,i));
      ez[I3(k,j,i-1)] -= 0.5*(DEZ_DXM(k,j,i) + DEZ_DXM(k,j+1,i));
      if (dims == 3){
        ex[I3(k,j,i)]   += 0.5*(DEX_DZP(k,j,i) + DEX_DZP(k,j+1,i));
        ex[I3(k-1,j,i)] -= 0.5*(DEX_DZM(k,j,i) + DEX_DZM(k,j+1,i));
      }
    }else{
      ez[I3(k,j,i)]   += DEZ_DXP(k,ju,i);
      ez[I3(k,j,i-1)] -= DEZ_DXM(k,ju,i);
      if (dims == 3){
        ex[I3(k,j,i)]   += DEX_DZP(k,ju,i);
        ex[I3(k-1,j,i)] -= DEX_DZM(k,ju,i);
      }
    }

    if (dims == 3){
      if (sz == 0){
        ex[I3(k,j,i)]   += 0.5*(DEX_DYP(k,j,i) + DEX_DYP(k+1,j,i));
        ex[I3(k,j-1,i)] -= 0.5*(DEX_DYM(k,j,i) + DEX_DYM(k+1,j,i));
        ey[I3(k,j,i)]   += 0.5*(DEY_DXP(k,j,i) + DEY_DXP(k+1,j,i));
        ey[I3(k,j,i-1)] -= 0.5*(DEY_DXM(k,j,i) + DEY_DXM(k+1,j,i));
      }else{
        ex[I3(k,j,i)]   += DEX_DYP(ku,j,i);
        ex[I3(k,j-1,i)] -= DEX_DYM(ku,j,i);
        ey[I3(k,j,i)]   += DEY_DXP(ku,j,i);
        ey[I3(k,j,i-1)] -= DEY_DXM(ku,j,i);
      }
    }
  }

  for (k = kbeg; k <= kend; k++) for (j = jbeg; j <= jend; j++) for (i = ibeg; i <= iend; i++){
    if (dims == 3){ ex[I3(k,j,i)] *= 0.25; ey[I3(k,j,i)] *= 0.25; }
    ez[I3(k,j,i)] *= 0.25;
  }
}

static void ct_update (Oracle *o, double dt)
/* MHD/CT/ct_update.c:79-218 (Cartesian), in place on Vs */
{
  int i, j, k, dims = o->c.dims, koff = (dims == 3 ? 1 : 0);
  int ibeg = o->emf_ibeg, iend = o->emf_iend, jbeg = o->emf_jbeg, jend = o->emf_jend;
  int kbeg = o->emf_kbeg, kend = o->emf_kend;
  double *Ex1 = o->ex, *Ex2 = o->ey, *Ex3 = o->ez;
  double rhs;

  for (k = kbeg + koff; k <= kend; k++) for (j = jbeg + 1; j <= jend; j++)
  for (i = ibeg; i <= iend; i++){
    if (dims == 3)
      rhs = 0.0 - dt/DXA(o,1,j)*(Ex3[I3(k,j,i)] - Ex3[I3(k,j-1,i)])
                + dt/DXA(o,2,k)*(Ex2[I3(k,j,i)] - Ex2[I3(k-1,j,i)]);
    else
      rhs = 0.0 - dt/DXA(o,1,j)*(Ex3[I3(k,j,i)] - Ex3[I3(k,j-1,i)]);
    o->Vs[0][I3(k,j,i)] = o->Vs[0][I3(k,j,i)] + rhs;
  }
  for (k = kbeg + koff; k <= kend; k++) for (j = jbeg; j <= jend; j++)
  for (i = ibeg + 1; i <= iend; i++){
    if (dims == 3)
      rhs =   dt/DXA(o,0,i)*(Ex3[I3(k,j,i)] - Ex3[I3(k,j,i-1)])
            - dt/DXA(o,2,k)*(Ex1[I3(k,j,i)] - Ex1[I3(k-1,j,i)]);
    else
      rhs =   dt/DXA(o,0,i)*(Ex3[I3(k,j,i)] - Ex3[I3(k,j,i-1)]);
    o->Vs[1][I3(k,j,i)] = o->Vs[1][I3(k,j,i)] + rhs;
  }
  if (dims == 3)
  for (k = kbeg; k <= kend; k++) for (j = jbeg + 1; j <= jend; j++)
  for (i = ibeg + 1; i <= iend; i++){
    rhs = - dt/DXA(o,0,i)*(Ex2[I3(k,j,i)] - Ex2[I3(k,j,i-1)])
          + dt/DXA(o,1,j)*(Ex1[I3(k,j,i)] - Ex1[I3(k,j-1,i)]);
    o->Vs[2][I3(k,j,i)] = o->Vs[2][I3(k,j,i)] + rhs;
  }
}

static void ct_update_from (Oracle *o, double *const *Bs, double dt)
/* CT_Update (d, Bs, dt): d->Vs = Bs + dt curl E on the faces of the emf ranges (ct_update.c:79-218) */
{
  int i, j, k, dims = o->c.dims, koff = (dims == 3 ? 1 : 0);
  int ibeg = o->emf_ibeg, iend = o->emf_iend, jbeg = o->emf_jbeg, jend = o->emf_jend;
  int kbeg = o->emf_kbeg, kend = o->emf_kend;
  double *Ex1 = o->ex, *Ex2 = o->ey, *Ex3 = o->ez;
  double rhs;
  for (k = kbeg + koff; k <= kend; k++) for (j = jbeg + 1; j <= jend; j++)
  for (i = ibeg; i <= iend; i++){
    if (dims == 3)
      rhs = 0.0 - dt/DXA(o,1,j)*(Ex3[I3(k,j,i)] - Ex3[I3(k,j-1,i)])
                + dt/DXA(o,2,k)*(Ex2[I3(k,j,i)] - Ex2[I3(k-1,j,i)]);
    else
      rhs = 0.0 - dt/DXA(o,1,j)*(Ex3[I3(k,j,i)] - Ex3[I3(k,j-1,i)]);
    o->Vs[0][I3(k,j,i)] = Bs[0][I3(k,j,i)] + rhs;
  }
  for (k = kbeg + koff; k <= kend; k++) for (j = jbeg; j <= jend; j++)
  for (i = ibeg + 1; i <= iend; i++){
    if (dims == 3)
      rhs =   dt/DXA(o,0,i)*(Ex3[I3(k,j,i)] - Ex3[I3(k,j,i-1)])
            - dt/DXA(o,2,k)*(Ex1[I3(k,j,i)] - Ex1[I3(k-1,j,i)]);
    else
      rhs =   dt/DXA(o,0,i)*(Ex3[I3(k,j,i)] - Ex3[I3(k,j,i-1)]);
    o->Vs[1][I3(k,j,i)] = Bs[1][I3(k,j,i)] + rhs;
  }
  if (dims == 3)
  for (k = kbeg; k <= kend; k++) for (j = jbeg + 1; j <= jend; j++)
  for (i = ibeg + 1; i <= iend; i++){
    rhs = - dt/DXA(o,0,i)*(Ex2[I3(k,j,i)] - Ex2[I3(k,j,i-1)])
          + dt/DXA(o,1,j)*(Ex1[I3(k,j,i)] - Ex1[I3(k,j-1,i)]);
    o->Vs[2][I3(k,j,i)] = Bs[2][I3(k,j,i)] + rhs;
  }
}

static void ct_average_magnetic_field (Oracle *o)
/* MHD/CT/ct_field_average.c:58-124 (Cartesian): DOM +/- 1 in every active direction, writes into Uc;
   with CT_EN_CORRECTION YES the energy is redefined with the averaged field (:116-129) */
{
  int i, j, k, dims = o->c.dims, koff = (dims == 3 ? 1 : 0);
  for (k = o->beg[2]-koff; k <= o->end[2]+koff; k++)
  for (j = o->beg[1]-1; j <= o->end[1]+1; j++)
  for (i = o->beg[0]-1; i <= o->end[0]+1; i++){
    double bx_ave, by_ave, bz_ave = 0.0, b2_old = 0.0, b2_new;
    bx_ave = 0.5*(o->Vs[0][I3(k,j,i)] + o->Vs[0][I3(k,j,i-1)]);
    by_ave = 0.5*(o->Vs[1][I3(k,j,i)] + o->Vs[1][I3(k,j-1,i)]);
    if (dims == 3) bz_ave = 0.5*(o->Vs[2][I3(k,j,i)] + o->Vs[2][I3(k-1,j,i)]);
    if (o->c.en_correction){
      if (dims == 3) b2_old = o->Uc[BX1][I3(k,j,i)]*o->Uc[BX1][I3(k,j,i)] + o->Uc[BX2][I3(k,j,i)]*o->Uc[BX2][I3(k,j,i)]
                            + o->Uc[BX3][I3(k,j,i)]*o->Uc[BX3][I3(k,j,i)];
      else           b2_old = o->Uc[BX1][I3(k,j,i)]*o->Uc[BX1][I3(k,j,i)] + o->Uc[BX2][I3(k,j,i)]*o->Uc[BX2][I3(k,j,i)];
    }
    o->Uc[BX1][I3(k,j,i)] = bx_ave;
    o->Uc[BX2][I3(k,j,i)] = by_ave;
    if (dims == 3) o->Uc[BX3][I3(k,j,i)] = bz_ave;
    if (o->c.en_correction){
      if (dims == 3) b2_new = bx_ave*bx_ave + by_ave*by_ave + bz_ave*bz_ave;
      else           b2_new = bx_ave*bx_ave + by_ave*by_ave;
      o->Uc[ENG][I3(k,j,i)] += 0.5*(b2_new - b2_old);
    }
  }
}

/* =====================================================================
   AdvanceStep  (reference Time_Stepping/rk_step.c:27-254)
   ===================================================================== */

static void prim_to_cons_3d (Oracle *o)      /* mappers3D.c:74-119, DOM box */
{
  int i, j, k, nv;
  for (k = o->beg[2]; k <= o->end[2]; k++) for (j = o->beg[1]; j <= o->end[1]; j++)
  for (i = o->beg[0]; i <= o->end[0]; i++){
    double v[NV], u[NV]; int id = I3(k,j,i);
    for (nv = 0; nv < NV; nv++) v[nv] = o->Vc[nv][id];
    prim_to_cons (o, v, u);
    for (nv = 0; nv < NV; nv++) o->Uc[nv][id] = u[nv];
  }
}

static void cons_to_prim_3d (Oracle *o)      /* mappers3D.c:16-72, DOM box */
{
  int i, j, k, nv;
  for (k = o->beg[2]; k <= o->end[2]; k++) for (j = o->beg[1]; j <= o->end[1]; j++)
  for (i = o->beg[0]; i <= o->end[0]; i++){
    double v[NV], u[NV]; int id = I3(k,j,i);
    for (nv = 0; nv < NV; nv++) u[nv] = o->Uc[nv][id];
    o->floor_events += cons_to_prim (o, u, v);
    for (nv = 0; nv < NV; nv++){ o->Uc[nv][id] = u[nv]; o->Vc[nv][id] = v[nv]; }
  }
}

static void flag_shock (Oracle *o)
/* flag_shock.c:79-230, SHOCK_FLATTENING MULTID, Cartesian, ideal EOS; flags zeroed every
   step (main.c:329), evaluated once per step after the first Boundary call (rk_step.c:86-88) */
{
  int i, j, k, dims = o->c.dims;
  int koff = (dims == 3 ? 1 : 0);
  const double *vx1 = o->Vc[VX1], *vx2 = o->Vc[VX2], *vx3 = o->Vc[VX3], *pt = o->Vc[PRS];
  memset (o->flag, 0, (size_t)o->tot);
  for (k = koff; k < o->T[2] - koff; k++)
  for (j = 1; j < o->T[1] - 1; j++)
  for (i = 1; i < o->T[0] - 1; i++){
    double dvx1, dvx2, dvx3 = 0.0, divv, gradp, pt_min, pt_min1, pt_min2, pt_min3, dpx1, dpx2, dpx3;
    dvx1 = (vx1[I3(k,j,i+1)] - vx1[I3(k,j,i-1)])/DXA(o,0,i);       /* dx1[i], dx2[j], dx3[k]: flag_shock.c:143-145 */
    dvx2 = (vx2[I3(k,j+1,i)] - vx2[I3(k,j-1,i)])/DXA(o,1,j);
    if (dims == 3){
      dvx3 = (vx3[I3(k+1,j,i)] - vx3[I3(k-1,j,i)])/DXA(o,2,k);
      divv = dvx1 + dvx2 + dvx3;
    }else divv = dvx1 + dvx2;
    if (divv < 0.0){
      pt_min  = pt[I3(k,j,i)];
      pt_min1 = MINV(pt[I3(k,j,i+1)], pt[I3(k,j,i-1)]);
      pt_min2 = MINV(pt[I3(k,j+1,i)], pt[I3(k,j-1,i)]);
      pt_min  = MINV(pt_min, pt_min1);
      pt_min  = MINV(pt_min, pt_min2);
      dpx1 = fabs(pt[I3(k,j,i+1)] - pt[I3(k,j,i-1)]);
      dpx2 = fabs(pt[I3(k,j+1,i)] - pt[I3(k,j-1,i)]);
      if (dims == 3){
        pt_min3 = MINV(pt[I3(k+1,j,i)], pt[I3(k-1,j,i)]);
        pt_min  = MINV(pt_min, pt_min3);
        dpx3 = fabs(pt[I3(k+1,j,i)] - pt[I3(k-1,j,i)]);
        gradp = dpx1 + dpx2 + dpx3;
      }else gradp = dpx1 + dpx2;
      if (gradp > 5.0*pt_min){                 /* EPS_PSHOCK_FLATTEN, flag_shock.c:69-71 */
        o->flag[I3(k,j,i)] |= 4 | 1;
        o->flag[I3(k,j,i+1)] |= 1; o->flag[I3(k,j,i-1)] |= 1;
        o->flag[I3(k,j-1,i)] |= 1; o->flag[I3(k,j+1,i)] |= 1;
        if (dims == 3){ o->flag[I3(k-1,j,i)] |= 1; o->flag[I3(k+1,j,i)] |= 1; }
      }
    }
  }
}

/* =====================================================================
   AdvanceStep, corner-transport upwind (Time_Stepping/ctu_step.c:142-727) with
   the primitive MUSCL-Hancock predictor (States/hancock.c:33-142, MHD/prim_eqn.c:26-89;
   PrimSource vanishes for Cartesian CT) and CTU_CT_Source (ctu_step.c:731-816)
   ===================================================================== */
static void prim_rhs (const Oracle *o, const double *v, const double *dv, Dirs q, double *Adv)
{
  int dims = o->c.dims;
  double tau = 1.0/v[RHO], scrh;
  int nv;
  for (nv = 0; nv < NV; nv++) Adv[nv] = 0.0;
  Adv[RHO] = v[q.vn]*dv[RHO] + v[RHO]*dv[q.vn];
  if (dims == 3) scrh = 0.0 + v[q.bt]*dv[q.bt] + v[q.bb]*dv[q.bb];
  else           scrh = 0.0 + v[q.bt]*dv[q.bt];
  Adv[q.vn] = v[q.vn]*dv[q.vn] + tau*(dv[PRS] + scrh);
  Adv[q.vt] = v[q.vn]*dv[q.vt] - tau*v[q.bn]*dv[q.bt];
  if (dims == 3) Adv[q.vb] = v[q.vn]*dv[q.vb] - tau*v[q.bn]*dv[q.bb];
  Adv[q.bn] = 0.0;
  Adv[q.bt] = v[q.bt]*dv[q.vn] - v[q.bn]*dv[q.vt] + v[q.vn]*dv[q.bt];
  if (dims == 3) Adv[q.bb] = v[q.bb]*dv[q.vn] - v[q.bn]*dv[q.vb] + v[q.vn]*dv[q.bb];
  Adv[PRS] = o->c.gamma*v[PRS]*dv[q.vn] + v[q.vn]*dv[PRS];
}

static void hancock_step (Oracle *o, int beg, int end, Dirs q, double dt, int dir)
{
  int i, nv, dims = o->c.dims;
  double dt_2 = 0.5*dt;
  for (i = beg; i <= end; i++){
    double dv[NV], Adv[NV];
    const double dx = DXA(o,dir,i), d_dl = 1.0/dx;          /* hancock.c:83: d_dl[i] = 1/dx[i] */
    for (nv = 0; nv < NV; nv++) dv[nv] = o->vp[i][nv] - o->vm[i][nv];
    prim_rhs (o, o->v[i], dv, q, Adv);
    for (nv = 0; nv < NV; nv++){
      double scrh;
      if (dims == 2 && (nv == VX3 || nv == BX3)) continue;
      double src = 0.0;                /* PrimSource, prim_eqn.c:289-360 */
      if (nv == q.vn){
        if (o->c.body_force) src += o->gpen[i];
        if (o->phic) src -= (o->ppen[i] - o->ppen[i-1])/(1.0*dx);
      }
      scrh = dt_2*(d_dl*Adv[nv] - src);
      o->vp[i][nv] -= scrh;
      o->vm[i][nv] -= scrh;
    }
  }
  /* CheckPrimStates (check_states.c:18-72): first order where p or rho turned negative */
  for (i = beg; i <= end; i++){
    double *ap = o->vp[i], *am = o->vm[i], *ac = o->v[i];
    int sw = (ap[PRS] < 0.0) || (am[PRS] < 0.0);
    sw = sw || (ap[RHO] < 0.0) || (am[RHO] < 0.0);
    if (sw){
      double bp = ap[q.bn], bm = am[q.bn];
      for (nv = 0; nv < NV; nv++) am[nv] = ap[nv] = ac[nv];
      ap[q.bn] = bp; am[q.bn] = bm;
    }
  }
  for (i = beg; i <= end; i++)
    for (nv = 0; nv < NV; nv++) o->v[i][nv] = 0.5*(o->vp[i][nv] + o->vm[i][nv]);
}

static void prim_to_char (const Oracle *o, double LL[NWAVE][NV], const double *d, double *w, Dirs q)
/* PrimToChar (eigenv.c:1310-1400): only the non-zero entries of the left eigenvectors; CT: w[KDIVB] = 0 */
{
  const int nc = o->c.dims;
  const double *L;
  double wv, wB;
  int k;
  for (k = 0; k < NWAVE; k++) w[k] = 0.0;
  L = LL[KFASTM];
  if (nc == 3){ wv = L[q.vn]*d[q.vn] + L[q.vt]*d[q.vt] + L[q.vb]*d[q.vb]; wB = L[PRS]*d[PRS] + L[q.bt]*d[q.bt] + L[q.bb]*d[q.bb]; }
  else        { wv = L[q.vn]*d[q.vn] + L[q.vt]*d[q.vt];                   wB = L[PRS]*d[PRS] + L[q.bt]*d[q.bt]; }
  w[KFASTM] =  wv + wB;
  w[KFASTP] = -wv + wB;
  L = LL[KENTRP];
  w[KENTRP] = L[RHO]*d[RHO] + L[PRS]*d[PRS];
  w[KDIVB] = 0.0;
  L = LL[KSLOWM];
  if (nc == 3){ wv = L[q.vn]*d[q.vn] + L[q.vt]*d[q.vt] + L[q.vb]*d[q.vb]; wB = L[PRS]*d[PRS] + L[q.bt]*d[q.bt] + L[q.bb]*d[q.bb]; }
  else        { wv = L[q.vn]*d[q.vn] + L[q.vt]*d[q.vt];                   wB = L[PRS]*d[PRS] + L[q.bt]*d[q.bt]; }
  w[KSLOWM] =  wv + wB;
  w[KSLOWP] = -wv + wB;
  if (nc == 3){
    L = LL[KALFVM];
    wv = L[q.vt]*d[q.vt] + L[q.vb]*d[q.vb];
    wB = L[q.bt]*d[q.bt] + L[q.bb]*d[q.bb];
    w[KALFVM] = wv + wB;
    w[KALFVP] = wv - wB;
  }
}

static void char_tracing_step (Oracle *o, int beg, int end, Dirs q, double dt, int dir)
/* CharTracingStep for LINEAR reconstruction (States/char_tracing.c:278-560), CARTESIAN, CHAR_LIMITING NO,
   UNIFORM_CARTESIAN_GRID YES, CHTR_REF_STATE 3 (:268-270), no source terms (PrimSource vanishes with CT and without body
   forces, prim_eqn.c:289-360).  Pinned with 2 components; with 3 the reference's right-eigenvector scratch keeps entries of
   the previous sweep direction (see states_plm_char). */
{
  const int nc = o->c.dims, nw = (nc == 3 ? 8 : 6);
  int i, nv, k;
  for (i = beg; i <= end; i++){
    double (*RR)[NV] = o->Rp[i];
    double LL[NWAVE][NV], lambda[NWAVE], nu[NWAVE], dv[NV], dw[NWAVE];
    double *vc = o->v[i], *vp = o->vp[i], *vm = o->vm[i];
    double dx, dtdx, nu_max, nu_min;
    prim_eigenvectors (o, vc, q, RR, LL, lambda);          /* SoundSpeed2 + PrimEigenvectors, :323-324 */
    dx   = DXA(o,dir,i);                                    /* :346-347 */
    dtdx = dt/dx;
    for (k = 0; k < nw; k++) nu[k] = dtdx*lambda[k];        /* :361 */
    nu_max = MAXV(nu[1], 0.0); nu_min = MINV(nu[0], 0.0);   /* :362 */
    for (nv = NV; nv--;  ) dv[nv] = vp[nv] - vm[nv];        /* :388 */
    prim_to_char (o, LL, dv, dw, q);                        /* :389 */
    for (nv = NV; nv--;  ){                                 /* :409-417, CHTR_REF_STATE 3 */
      if (nc == 2 && (nv == VX3 || nv == BX3)) continue;
      vp[nv] = vc[nv] + 0.5*dv[nv]*(1.0 - nu_max);
      vm[nv] = vc[nv] - 0.5*dv[nv]*(1.0 + nu_min);
    }
    for (k = 0; k < nw; k++){                               /* :441-471 */
      if (nu[k] >= 0.0){
        dw[k] *= 0.5*(nu_max - nu[k]);
        for (nv = 0; nv < NV; nv++){
          if (nc == 2 && (nv == VX3 || nv == BX3)) continue;
          vp[nv] += dw[k]*RR[nv][k];
        }
      }else{
        dw[k] *= 0.5*(nu_min - nu[k]);
        for (nv = 0; nv < NV; nv++){
          if (nc == 2 && (nv == VX3 || nv == BX3)) continue;
          vm[nv] += dw[k]*RR[nv][k];
        }
      }
    }
    for (nv = NV; nv--;  ){                                 /* :477-480: src = 0 */
      if (nc == 2 && (nv == VX3 || nv == BX3)) continue;
      vp[nv] += 0.5*dt*0.0;
      vm[nv] += 0.5*dt*0.0;
    }
  }
  for (i = beg-1; i <= end; i++) o->vp[i][q.bn] = o->vm[i+1][q.bn] = o->bn[i];      /* :531-535 */
  /* CheckPrimStates (check_states.c:18-72): first order where p or rho turned negative */
  for (i = beg; i <= end; i++){
    double *ap = o->vp[i], *am = o->vm[i], *ac = o->v[i];
    int sw = (ap[PRS] < 0.0) || (am[PRS] < 0.0);
    sw = sw || (ap[RHO] < 0.0) || (am[RHO] < 0.0);
    if (sw){
      double bp = ap[q.bn], bm = am[q.bn];
      for (nv = 0; nv < NV; nv++) am[nv] = ap[nv] = ac[nv];
      ap[q.bn] = bp; am[q.bn] = bm;
    }
  }
  for (i = beg; i <= end; i++)                             /* :551-555 */
    for (nv = 0; nv < NV; nv++) o->v[i][nv] = 0.5*(o->vp[i][nv] + o->vm[i][nv]);
}

static void ctu_store_emf (Oracle *o, int dir, int nbeg, int nend, int *idx3)
/* CT_StoreUpwindEMF (ct_emf.c:104-190) for faces nbeg-1 .. nend of the current pencil */
{
  int n, dims = o->c.dims;
  for (n = nbeg-1; n <= nend; n++){
    int id; signed char s;
    idx3[dir] = n;
    id = IDX(o, idx3[2], idx3[1], idx3[0]);
    if      (o->flux[n][RHO] >  EPS_UCT_CONTACT) s = 1;
    else if (o->flux[n][RHO] < -EPS_UCT_CONTACT) s = -1;
    else s = 0;
    if (dir == 0){
      o->ezi[id] = -o->flux[n][BX2];
      if (dims == 3) o->eyi[id] = o->flux[n][BX3];
      o->svx[id] = s;
    }else if (dir == 1){
      o->ezj[id] = o->flux[n][BX1];
      if (dims == 3) o->exj[id] = -o->flux[n][BX3];
      o->svy[id] = s;
    }else{
      o->eyk[id] = -o->flux[n][BX1];
      o->exk[id] =  o->flux[n][BX2];
      o->svz[id] = s;
    }
  }
}

static void ctu_riemann (Oracle *o, Dirs q, int nbeg, int nend)
/* Riemann (nbeg-1 .. nend): stateL = (vp[n], up[n]), stateR = (vm[n+1], um[n+1]) */
{
  int n;
  for (n = nbeg-1; n <= nend; n++){
    o->use_hll = ((o->pflag[n] & 4) || (o->pflag[n+1] & 4));
    if      (o->c.solver == ORC_SOLVER_HLLD)
      riemann_hlld (o, o->vp[n], o->vm[n+1], o->up[n], o->um[n+1], q, o->flux[n], &o->press[n], &o->cmax[n]);
    else if (o->c.solver == ORC_SOLVER_HLLC)
      riemann_hllc (o, o->vp[n], o->vm[n+1], o->up[n], o->um[n+1], q, o->flux[n], &o->press[n], &o->cmax[n]);
    else if (o->c.solver == ORC_SOLVER_TVDLF)
      riemann_tvdlf (o, o->vp[n], o->vm[n+1], o->up[n], o->um[n+1], q, o->flux[n], &o->press[n], &o->cmax[n]);
    else if (o->c.solver == ORC_SOLVER_HLL || o->use_hll)
      riemann_hll  (o, o->vp[n], o->vm[n+1], o->up[n], o->um[n+1], q, o->flux[n], &o->press[n], &o->cmax[n]);
    else
      riemann_roe  (o, o->vp[n], o->vm[n+1], o->up[n], o->um[n+1], q, o->flux[n], &o->press[n], &o->cmax[n]);
  }
}

static void ctu_advance (Oracle *o, double dt)
{
  int dims = o->c.dims, koff = (dims == 3 ? 1 : 0);
  int dir, nv, i, j, k, d, id;
  double dt2 = 0.5*dt;

  /* 2. boundary conditions, shock flags (ctu_step.c:236-244) */
  o->stage = 1;
  boundary (o);
  if (o->c.shock_flattening) flag_shock (o);
  /* 3. Bs0 = Vs, Uc = PrimToCons (Vc) over TOT (:249-262) */
  for (d = 0; d < dims; d++) memcpy (o->Bs0[d], o->Vs[d], sizeof(double)*(size_t)o->tot);
  for (k = 0; k < o->T[2]; k++) for (j = 0; j < o->T[1]; j++) for (i = 0; i < o->T[0]; i++){
    double v[NV], u[NV];
    id = I3(k,j,i);
    for (nv = 0; nv < NV; nv++) v[nv] = o->Vc[nv][id];
    prim_to_cons (o, v, u);
    for (nv = 0; nv < NV; nv++) o->Uc[nv][id] = u[nv];
  }

  /* 4. predictor: normal predictors and normal Riemann problems (:283-420) */
  for (dir = 0; dir < dims; dir++){
    Dirs q = set_vector_indices (dir);
    int lo[3], hi[3], t1, t2, a, b, n, nbeg, nend, ntot = o->T[dir];
#define dt2_dx (dt2/DXA(o,dir,n))                  /* ctu_step.c:310  dt2_dx[i] = dt2/dx[i]; inv_dl[i] = 1/dx[i] */
#define inv_dl (1.0/DXA(o,dir,n))
    for (a = 0; a < 3; a++){ lo[a] = o->beg[a]; hi[a] = o->end[a]; }
    /* transverse +-1 (:290-292), then with CT normal +-1 and transverse +-1 again (:293-297) */
    for (a = 0; a < dims; a++){ if (a != dir){ lo[a] -= 2; hi[a] += 2; } else { lo[a]--; hi[a]++; } }
    nbeg = lo[dir]; nend = hi[dir];
    if (dir == 0){ t1 = 1; t2 = 2; } else if (dir == 1){ t1 = 0; t2 = 2; } else { t1 = 0; t2 = 1; }
    for (b = lo[t2]; b <= hi[t2]; b++) for (a = lo[t1]; a <= hi[t1]; a++){
      int idx3[3];
      idx3[t1] = a; idx3[t2] = b;
      for (n = 0; n < ntot; n++){
        idx3[dir] = n;
        id = IDX(o, idx3[2], idx3[1], idx3[0]);
        for (nv = 0; nv < NV; nv++) o->vn[n][nv] = o->v[n][nv] = o->Vc[nv][id];
        o->gpen[n] = (o->gf[dir] ? o->gf[dir][id] : o->c.grav[dir]);
        if (o->phic) o->ppen[n] = o->phif[dir][id];
        o->bn[n] = o->Vs[dir][id];
        o->pflag[n] = o->flag[id];
      }
      /* 4d. States (nbeg-1 .. nend+1): PLM (incl. face field), Hancock, PrimToCons (plm_states.c:80-312) */
      if (o->c.char_limiting) states_plm_char (o, nbeg-1, nend+1, q);      /* CHAR_LIMITING YES (plm_states.c:448-706) */
      else                    states_plm (o, nbeg-1, nend+1, q.bn);
      if (o->c.ctu == 2) char_tracing_step (o, nbeg-1, nend+1, q, dt, dir);      /* TIME_STEPPING CHARACTERISTIC_TRACING */
      else               hancock_step (o, nbeg-1, nend+1, q, dt, dir);
      for (n = nbeg-1; n <= nend+1; n++){ prim_to_cons (o, o->vp[n], o->up[n]); }
      for (n = nbeg-1; n <= nend+1; n++){ prim_to_cons (o, o->vm[n], o->um[n]); }
      /* 4f. Riemann, EMF, rhs with dt/2 */
      ctu_riemann (o, q, nbeg, nend);
      ctu_store_emf (o, dir, nbeg, nend, idx3);
      if (o->phic) for (n = nbeg-1; n <= nend; n++) o->flux[n][ENG] += o->flux[n][RHO]*o->ppen[n];   /* TotalFlux */
      /* CTU_CT_Source (nbeg-1 .. nend+1), ctu_step.c:731-816 */
      for (n = nbeg-1; n <= nend+1; n++){
        double db = dt2_dx*(o->up[n][q.bn] - o->um[n][q.bn]), scrh;
        const double *v = o->vn[n];
        o->up[n][VX1] += v[BX1]*db; o->um[n][VX1] += v[BX1]*db;
        o->up[n][VX2] += v[BX2]*db; o->um[n][VX2] += v[BX2]*db;
        if (dims == 3){ o->up[n][VX3] += v[BX3]*db; o->um[n][VX3] += v[BX3]*db; }
        o->up[n][q.bt] += v[q.vt]*db; o->um[n][q.bt] += v[q.vt]*db;
        if (dims == 3){ o->up[n][q.bb] += v[q.vb]*db; o->um[n][q.bb] += v[q.vb]*db; }
        if (dims == 3) scrh = v[VX1]*v[BX1] + v[VX2]*v[BX2] + v[VX3]*v[BX3];
        else           scrh = v[VX1]*v[BX1] + v[VX2]*v[BX2];
        o->up[n][ENG] += scrh*db; o->um[n][ENG] += scrh*db;
      }
      for (n = nbeg-1; n <= nend+1; n++){
        idx3[dir] = n;
        id = IDX(o, idx3[2], idx3[1], idx3[0]);
        for (nv = 0; nv < NV; nv++){ o->Up[dir][nv][id] = o->up[n][nv]; o->Um[dir][nv][id] = o->um[n][nv]; }
      }
      /* RightHandSide (nbeg .. nend, dt/2), rhs.c:193-201; 4g. store */
      for (n = nbeg; n <= nend; n++){
        double rhs[NV];
        idx3[dir] = n;
        id = IDX(o, idx3[2], idx3[1], idx3[0]);
        for (nv = 0; nv < NV; nv++) rhs[nv] = -dt2_dx*(o->flux[n][nv] - o->flux[n-1][nv]);
        rhs[q.vn] -= dt2_dx*(o->press[n] - o->press[n-1]);
        if (o->c.body_force){          /* stateC->v is the half-step zone average left by HancockStep (hancock.c:136-141) */
          const double g = (o->gf[dir] ? o->gf[dir][id] : o->c.grav[dir]);
          rhs[q.vn] += dt2*o->v[n][RHO]*g;
          rhs[ENG]  += dt2*0.5*(o->flux[n][RHO] + o->flux[n-1][RHO])*g;
        }
        if (o->phic){
          rhs[q.vn] -= dt2_dx*o->v[n][RHO]*(o->ppen[n] - o->ppen[n-1]);
          rhs[ENG]  -= o->phic[id]*rhs[RHO];
        }
        for (nv = 0; nv < NV; nv++) o->rhs3[dir][nv][id] = rhs[nv];
        o->inv_dt_hyp = MAXV(o->inv_dt_hyp, o->cmax[n]*inv_dl);       /* :416-419 */
      }
    }
  }
  /* emf ranges as left by the predictor sweeps (ct_emf.c:130,152,173 with beg = nbeg-1, end = nend) */
  o->emf_ibeg = o->beg[0]-2; o->emf_iend = o->end[0]+1;
  o->emf_jbeg = o->beg[1]-2; o->emf_jend = o->end[1]+1;
  if (dims == 3){ o->emf_kbeg = o->beg[2]-2; o->emf_kend = o->end[2]+1; }
  else          { o->emf_kbeg = o->emf_kend = 0; }

  /* 5a. Uh = U^n + sum of the half-step rhs over DOM +- 1 (:427-441) */
  for (k = o->beg[2]-koff; k <= o->end[2]+koff; k++)
  for (j = o->beg[1]-1; j <= o->end[1]+1; j++)
  for (i = o->beg[0]-1; i <= o->end[0]+1; i++){
    id = I3(k,j,i);
    for (nv = 0; nv < NV; nv++){
      double dU;
      if (dims == 3) dU = o->rhs3[0][nv][id] + o->rhs3[1][nv][id] + o->rhs3[2][nv][id];
      else           dU = o->rhs3[0][nv][id] + o->rhs3[1][nv][id];
      o->Uh[nv][id] = o->Uc[nv][id] + dU;
    }
  }
  /* 5b. emf; 5d. half-step staggered field and its cell average into Uh (:447-470) */
  ct_compute_emf (o);
  ct_update (o, 0.5*dt);
  for (k = o->beg[2]-koff; k <= o->end[2]+koff; k++)
  for (j = o->beg[1]-1; j <= o->end[1]+1; j++)
  for (i = o->beg[0]-1; i <= o->end[0]+1; i++){
    double b2_old = 0.0, b2_new;          /* CT_AverageMagneticField (d->Vs, Uh), CT_EN_CORRECTION: ct_field_average.c:116-129 */
    if (o->c.en_correction){
      if (dims == 3) b2_old = o->Uh[BX1][I3(k,j,i)]*o->Uh[BX1][I3(k,j,i)] + o->Uh[BX2][I3(k,j,i)]*o->Uh[BX2][I3(k,j,i)]
                            + o->Uh[BX3][I3(k,j,i)]*o->Uh[BX3][I3(k,j,i)];
      else           b2_old = o->Uh[BX1][I3(k,j,i)]*o->Uh[BX1][I3(k,j,i)] + o->Uh[BX2][I3(k,j,i)]*o->Uh[BX2][I3(k,j,i)];
    }
    o->Uh[BX1][I3(k,j,i)] = 0.5*(o->Vs[0][I3(k,j,i)] + o->Vs[0][I3(k,j,i-1)]);
    o->Uh[BX2][I3(k,j,i)] = 0.5*(o->Vs[1][I3(k,j,i)] + o->Vs[1][I3(k,j-1,i)]);
    if (dims == 3) o->Uh[BX3][I3(k,j,i)] = 0.5*(o->Vs[2][I3(k,j,i)] + o->Vs[2][I3(k-1,j,i)]);
    if (o->c.en_correction){
      if (dims == 3) b2_new = o->Uh[BX1][I3(k,j,i)]*o->Uh[BX1][I3(k,j,i)] + o->Uh[BX2][I3(k,j,i)]*o->Uh[BX2][I3(k,j,i)]
                            + o->Uh[BX3][I3(k,j,i)]*o->Uh[BX3][I3(k,j,i)];
      else           b2_new = o->Uh[BX1][I3(k,j,i)]*o->Uh[BX1][I3(k,j,i)] + o->Uh[BX2][I3(k,j,i)]*o->Uh[BX2][I3(k,j,i)];
      o->Uh[ENG][I3(k,j,i)] += 0.5*(b2_new - b2_old);
    }
  }
  /* 5f. Vc = ConsToPrim (Uh) over DOM +- 1: V^{n+1/2} (:486-497) */
  for (k = o->beg[2]-koff; k <= o->end[2]+koff; k++)
  for (j = o->beg[1]-1; j <= o->end[1]+1; j++)
  for (i = o->beg[0]-1; i <= o->end[0]+1; i++){
    double v[NV], u[NV];
    id = I3(k,j,i);
    for (nv = 0; nv < NV; nv++) u[nv] = o->Uh[nv][id];
    o->floor_events += cons_to_prim (o, u, v);
    for (nv = 0; nv < NV; nv++){ o->Uh[nv][id] = u[nv]; o->Vc[nv][id] = v[nv]; }
  }

  /* 6. corrector (:517-640) */
  o->stage = 2;
  for (dir = 0; dir < dims; dir++){
    Dirs q = set_vector_indices (dir);
    int lo[3], hi[3], t1, t2, a, b, n, nbeg = o->beg[dir], nend = o->end[dir];
#undef dt2_dx
#define dtdx (dt/DXA(o,dir,n))
    for (a = 0; a < 3; a++){ lo[a] = o->beg[a]; hi[a] = o->end[a]; }
    for (a = 0; a < dims; a++) if (a != dir){ lo[a]--; hi[a]++; }
    if (dir == 0){ t1 = 1; t2 = 2; } else if (dir == 1){ t1 = 0; t2 = 2; } else { t1 = 0; t2 = 1; }
    for (b = lo[t2]; b <= hi[t2]; b++) for (a = lo[t1]; a <= hi[t1]; a++){
      int idx3[3];
      idx3[t1] = a; idx3[t2] = b;
      for (n = nbeg-1; n <= nend+1; n++){
        idx3[dir] = n;
        id = IDX(o, idx3[2], idx3[1], idx3[0]);
        for (nv = 0; nv < NV; nv++){
          double dU;
          if (dims == 3){
            if      (dir == 0) dU = 0.0 + o->rhs3[1][nv][id] + o->rhs3[2][nv][id];
            else if (dir == 1) dU = o->rhs3[0][nv][id] + 0.0 + o->rhs3[2][nv][id];
            else               dU = o->rhs3[0][nv][id] + o->rhs3[1][nv][id];
          }else{
            if (dir == 0) dU = 0.0 + o->rhs3[1][nv][id];
            else          dU = o->rhs3[0][nv][id] + 0.0;
          }
          o->up[n][nv] = o->Up[dir][nv][id] + dU;
          o->um[n][nv] = o->Um[dir][nv][id] + dU;
          o->v[n][nv]  = o->Vc[nv][id];
        }
        o->pflag[n] = o->flag[id];
      }
      /* normal field: the half-step staggered one (:583-587) */
      for (n = nbeg-2; n <= nend+1; n++){
        idx3[dir] = n;
        id = IDX(o, idx3[2], idx3[1], idx3[0]);
        o->bn[n] = o->Vs[dir][id];
        o->up[n][q.bn] = o->bn[n];
        o->um[n+1][q.bn] = o->bn[n];
      }
      for (n = nbeg-1; n <= nend+1; n++) o->floor_events += cons_to_prim (o, o->um[n], o->vm[n]);
      for (n = nbeg-1; n <= nend+1; n++) o->floor_events += cons_to_prim (o, o->up[n], o->vp[n]);
      ctu_riemann (o, q, nbeg, nend);
      ctu_store_emf (o, dir, nbeg, nend, idx3);
      if (o->phic) for (n = nbeg-1; n <= nend; n++){
        idx3[dir] = n;
        o->ppen[n] = o->phif[dir][IDX(o, idx3[2], idx3[1], idx3[0])];
        o->flux[n][ENG] += o->flux[n][RHO]*o->ppen[n];
      }
      for (n = nbeg; n <= nend; n++){
        double rhs[NV];
        idx3[dir] = n;
        id = IDX(o, idx3[2], idx3[1], idx3[0]);
        for (nv = 0; nv < NV; nv++) rhs[nv] = -dtdx*(o->flux[n][nv] - o->flux[n-1][nv]);
        rhs[q.vn] -= dtdx*(o->press[n] - o->press[n-1]);
        if (o->c.body_force){          /* stateC->v = V^{n+1/2} (ctu_step.c:566-570) */
          const double g = (o->gf[dir] ? o->gf[dir][id] : o->c.grav[dir]);
          rhs[q.vn] += dt*o->v[n][RHO]*g;
          rhs[ENG]  += dt*0.5*(o->flux[n][RHO] + o->flux[n-1][RHO])*g;
        }
        if (o->phic){
          rhs[q.vn] -= dtdx*o->v[n][RHO]*(o->ppen[n] - o->ppen[n-1]);
          rhs[ENG]  -= o->phic[id]*rhs[RHO];
        }
        for (nv = 0; nv < NV; nv++) o->Uc[nv][id] += rhs[nv];
        o->inv_dt_hyp = MAXV(o->inv_dt_hyp, o->cmax[n]*inv_dl);
      }
    }
  }
  o->emf_ibeg = o->beg[0]-1; o->emf_iend = o->end[0];
  o->emf_jbeg = o->beg[1]-1; o->emf_jend = o->end[1];
  if (dims == 3){ o->emf_kbeg = o->beg[2]-1; o->emf_kend = o->end[2]; }
  else          { o->emf_kbeg = o->emf_kend = 0; }
  /* 8. emf; 11. Vs = Bs0 + dt curl E, cell average; 14. ConsToPrim over DOM (:646-700) */
  ct_compute_emf (o);
  ct_update_from (o, o->Bs0, dt);      /* faces outside the emf ranges keep their half-step values, as in the reference */
  ct_average_magnetic_field (o);
  cons_to_prim_3d (o);
}
#undef dtdx
#undef inv_dl

int oracle_advance (Oracle *o, double dt, double *inv_dt_hyp, double *max_mach)
{
  int i, j, k, nv, d, id, dims = o->c.dims;
  double w0, wc;
  o->max_mach = 0.0;                     /* main.c:304 */
  o->inv_dt_hyp = 0.0;                   /* main.c:569 (reset by NextTimeStep) */
  o->floor_events = 0;
  if (o->c.ctu){ ctu_advance (o, dt); goto done; }

  /* ---- stage 1 (rk_step.c:85-139) ---- */
  o->stage = 1;
  boundary (o);
  if (o->c.shock_flattening) flag_shock (o);
  prim_to_cons_3d (o);
  for (nv = 0; nv < NV; nv++)
    for (k = o->beg[2]; k <= o->end[2]; k++) for (j = o->beg[1]; j <= o->end[1]; j++)
    for (i = o->beg[0]; i <= o->end[0]; i++) o->U0[nv][I3(k,j,i)] = o->Uc[nv][I3(k,j,i)];
  for (d = 0; d < dims; d++) memcpy (o->Bs0[d], o->Vs[d], sizeof(double)*(size_t)o->tot);
  update_stage (o, dt);
  ct_average_magnetic_field (o);
  cons_to_prim_3d (o);

  if (o->debug_stop_after == 1) goto done;

  /* ---- stage 2 (rk_step.c:149-186) ---- */
  if (o->c.rk_order == 3){ w0 = 0.75; wc = 0.25; } else { w0 = 0.5; wc = 0.5; }
  o->stage = 2;
  boundary (o);
  update_stage (o, dt);
  for (k = o->beg[2]; k <= o->end[2]; k++) for (j = o->beg[1]; j <= o->end[1]; j++)
  for (i = o->beg[0]; i <= o->end[0]; i++){
    id = I3(k,j,i);
    for (nv = 0; nv < NV; nv++) o->Uc[nv][id] = w0*o->U0[nv][id] + wc*o->Uc[nv][id];
  }
  for (d = 0; d < dims; d++)
    for (id = 0; id < o->tot; id++) o->Vs[d][id] = w0*o->Bs0[d][id] + wc*o->Vs[d][id];
  ct_average_magnetic_field (o);
  cons_to_prim_3d (o);

  /* ---- stage 3 (rk_step.c:204-243) ---- */
  if (o->c.rk_order == 3){
    double one_third = 1.0/3.0;
    o->stage = 3;
    boundary (o);
    update_stage (o, dt);
    for (k = o->beg[2]; k <= o->end[2]; k++) for (j = o->beg[1]; j <= o->end[1]; j++)
    for (i = o->beg[0]; i <= o->end[0]; i++){
      id = I3(k,j,i);
      for (nv = 0; nv < NV; nv++)
        o->Uc[nv][id] = one_third*(o->U0[nv][id] + 2.0*o->Uc[nv][id]);
    }
    for (d = 0; d < dims; d++)
      for (id = 0; id < o->tot; id++)
        o->Vs[d][id] = (o->Bs0[d][id] + 2.0*o->Vs[d][id])/3.0;
    ct_average_magnetic_field (o);
    cons_to_prim_3d (o);
  }

done:
  if (inv_dt_hyp) *inv_dt_hyp = o->inv_dt_hyp;
  if (max_mach)   *max_mach   = o->max_mach;
  return o->floor_events;
}

void oracle_debug_stop_after (Oracle *o, int stage) { o->debug_stop_after = stage; }

double oracle_next_dt (double inv_dt_hyp, double cfl, double cfl_max_var, double dt)
/* main.c:462-465, 532 */
{
  double dt_hyp = 1.0/inv_dt_hyp, dtnext;
  dt_hyp *= cfl;
  dtnext  = dt_hyp;
  dtnext  = MINV(dtnext, cfl_max_var*dt);
  return dtnext;
}

/* =====================================================================
   PPM reconstruction and Roe solver
   ===================================================================== */
#include "mhd_oracle_ppm_roe.inc"

/* ---- taps ---- */
const double *oracle_tap (const Oracle *o, const char *name)
{
  static const char *vn[NV] = {"rho","vx1","vx2","vx3","bx1","bx2","bx3","prs"};
  static const char *un[NV] = {"u_rho","u_mx1","u_mx2","u_mx3","u_bx1","u_bx2","u_bx3","u_eng"};
  int nv;
  for (nv = 0; nv < NV; nv++){
    if (!strcmp(name, vn[nv])) return o->Vc[nv];
    if (!strcmp(name, un[nv])) return o->Uc[nv];
  }
  if (!strcmp(name,"bx1s")) return o->Vs[0];
  if (!strcmp(name,"bx2s")) return o->Vs[1];
  if (!strcmp(name,"bx3s")) return o->Vs[2];
  if (!strcmp(name,"exj")) return o->exj;  if (!strcmp(name,"exk")) return o->exk;
  if (!strcmp(name,"eyi")) return o->eyi;  if (!strcmp(name,"eyk")) return o->eyk;
  if (!strcmp(name,"ezi")) return o->ezi;  if (!strcmp(name,"ezj")) return o->ezj;
  if (!strcmp(name,"ex")) return o->ex;  if (!strcmp(name,"ey")) return o->ey;
  if (!strcmp(name,"ez")) return o->ez;
  if (!strcmp(name,"C_dt")) return o->C_dt;
  return NULL;
}

void oracle_tap_shape (const Oracle *o, int *s3, int *s2, int *s1, int *tot)
{
  *s3 = o->S3; *s2 = o->S2; *s1 = o->S1; *tot = o->tot;
}
