/*
 * mhd_oracle.h -- CPU restatement of the PLUTO 4.3 unsplit RK2/RK3 + CT
 * ideal-MHD step.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; the product (pluto_b200/) never links or calls it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors for this
 * path; the restatement is pinned by bit-comparison against
 * the compiled, unmodified reference (oracle/_ref/pluto_*, built by
 * oracle/ref_build/build_ref.sh) in tests/test_oracle_vs_ref.py and against
 * the fixtures under tests/golden/ that the same binaries generated
 * (tools/make_golden.py).
 */
#ifndef MHD_ORACLE_H
#define MHD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_RECON_PLM = 0, ORC_RECON_PPM = 1 };
enum { ORC_SOLVER_HLLD = 0, ORC_SOLVER_HLL = 1, ORC_SOLVER_ROE = 2, ORC_SOLVER_HLLC = 3, ORC_SOLVER_TVDLF = 4 };
enum { ORC_BC_PERIODIC = 0, ORC_BC_OUTFLOW = 1, ORC_BC_REFLECTIVE = 2, ORC_BC_EQTSYMMETRIC = 3 };
/* LIMITER (plm_states.c:192-236, plm_coeffs.h:72-123): DEFAULT mixes MC / van Leer / minmod */
enum { ORC_LIM_DEFAULT = 0, ORC_LIM_FLAT, ORC_LIM_MINMOD, ORC_LIM_VANALBADA, ORC_LIM_OSPRE,
       ORC_LIM_UMIST, ORC_LIM_VANLEER, ORC_LIM_MC };
/* CT_EMF_AVERAGE (ct_emf.c:241-283) */
enum { ORC_EMF_UCT_CONTACT = 0, ORC_EMF_ARITHMETIC = 1, ORC_EMF_UCT0 = 2, ORC_EMF_UCT_HLL = 3 };

/* Variable order of the 8-slot state vector used by the oracle.  2-D
   (COMPONENTS = 2) runs carry vx3 = Bx3 = 0 in the unused slots, which is
   bit-identical to the reference's 6-variable arithmetic (every extra term
   is an exact +0). */
enum { ORC_RHO = 0, ORC_VX1, ORC_VX2, ORC_VX3, ORC_BX1, ORC_BX2, ORC_BX3, ORC_PRS, ORC_NV };

typedef struct {
  int    dims;          /* 2 or 3  (DIMENSIONS = COMPONENTS)               */
  int    n[3];          /* interior zones NX1, NX2, NX3 (NX3 = 1 in 2-D)   */
  int    recon;         /* ORC_RECON_*                                     */
  int    solver;        /* ORC_SOLVER_*                                    */
  int    rk_order;      /* 2 (RK2) or 3 (RK3)                              */
  int    bc[6];         /* X1_BEG, X1_END, X2_BEG, X2_END, X3_BEG, X3_END  */
  double gamma;         /* g_gamma                                         */
  double dx[3];         /* uniform cell sizes                              */
  double small_dn;      /* g_smallDensity  (1e-12)                         */
  double small_pr;      /* g_smallPressure (1e-12)                         */
  int    limiter;       /* ORC_LIM_*  (PLM only)                           */
  int    emf_average;   /* ORC_EMF_*                                       */
  int    shock_flattening; /* 0 NO, 1 MULTID (flag_shock.c:79-230)            */
  int    ctu;           /* 0: RK2/RK3 (rk_order); 1: corner-transport upwind with the
                           primitive MUSCL-Hancock predictor (TIME_STEPPING HANCOCK,
                           Time_Stepping/ctu_step.c, States/hancock.c); 2: the same step with
                           the characteristic-tracing predictor (TIME_STEPPING CHARACTERISTIC_TRACING,
                           States/char_tracing.c:278-560; 2 components)        */
  int    en_correction; /* CT_EN_CORRECTION YES (MHD/CT/ct_field_average.c:116-129)            */
  int    body_force;    /* BODY_FORCE VECTOR with a uniform acceleration grav[] (MHD/rhs_source.c:214-217,
                           277-280, 342-345; MHD/prim_eqn.c:289-360 in the Hancock predictor)     */
  double grav[3];
  int    char_limiting; /* CHAR_LIMITING YES (States/plm_states.c:448-706 with MHD/eigenv.c:190-560, LINEAR only): slopes limited
                           on the characteristic variables.  PINNED IN 2-D ONLY: the reference's right-eigenvector scratch keeps
                           the entries of the previous sweep direction (only non-zero entries are ever written); in 2-D no stale
                           entry reaches a result, in 3-D the stale Alfven entries of the normal-velocity row do               */
} OracleConfig;

typedef struct Oracle Oracle;

Oracle *oracle_create (const OracleConfig *cfg);
/* static, position-dependent body force: component d of BodyForceVector at every zone centre, ghost zones included,
   g[d][k][j][i] with the extents T3 x T2 x T1 of the reference's Data arrays (NULL for the third component in 2-D).
   Replaces the uniform grav[] of the configuration. */
/* non-uniform Cartesian grid: zone widths grid->dx[d][0 .. T_d-1], ghost zones included (RK path, LINEAR reconstruction) */
void    oracle_set_grid (Oracle *o, const double *dx1, const double *dx2, const double *dx3);
/* UNIFORM_CARTESIAN_GRID NO: the weights PLM_CoefficientsGet returns for direction dir (T entries each) */
void    oracle_set_ppm_coeffs (Oracle *o, int dir, const double *wm1, const double *w0, const double *w1, const double *w2);
void    oracle_set_plm_coeffs (Oracle *o, int dir, const double *cp, const double *cm, const double *wp, const double *wm,
                               const double *dp, const double *dm);
void    oracle_set_body_force (Oracle *o, const double *g1, const double *g2, const double *g3);
/* BODY_FORCE POTENTIAL (rhs.c:162-187, 388-392; rhs_source.c:233-237, 316-320, 358-362; prim_eqn.c:304-307): the potential
   at the zone centres, phic[k][j][i] (T3 x T2 x T1), and at the faces of every direction in the layout of the staggered
   Data arrays (pf1: T3 x T2 x (T1+1) from face -1/2, pf2: T3 x (T2+1) x T1, pf3: (T3+1) x T2 x T1; NULL in 2-D). */
void    oracle_set_body_potential (Oracle *o, const double *phic, const double *pf1, const double *pf2, const double *pf3);
void    oracle_destroy (Oracle *o);
int     oracle_nghost (const Oracle *o);

/* Interior state in the reference's .dbl layout:
   vc[nv][k][j][i]  (nv in oracle order, always 8 slots; n3*n2*n1 each)
   bx1s[k][j][i] with n1+1 faces, bx2s with n2+1, bx3s with n3+1 (3-D). */
void oracle_set_interior (Oracle *o, const double *vc, const double *bx1s,
                          const double *bx2s, const double *bx3s);
void oracle_get_interior (const Oracle *o, double *vc, double *bx1s,
                          double *bx2s, double *bx3s);

/* One AdvanceStep (reference Src/Time_Stepping/rk_step.c:27).  Returns the
   number of ConsToPrim floor events; *inv_dt_hyp is Dts->invDt_hyp as left
   by UpdateStage at stage 1 (update_stage.c:308-312), *max_mach is
   g_maxMach after the step. */
int oracle_advance (Oracle *o, double dt, double *inv_dt_hyp, double *max_mach);

/* NextTimeStep restatement (reference Src/main.c:389-575, hyperbolic part). */
double oracle_next_dt (double inv_dt_hyp, double cfl, double cfl_max_var, double dt);

/* Debug: make oracle_advance return after the given stage (0 = whole step). */
void oracle_debug_stop_after (Oracle *o, int stage);

/* Debug taps: raw pointers into the padded internal arrays, plus the
   padded extents so tests can index them ((k+1)*S2*S1 + (j+1)*S1 + (i+1)). */
const double *oracle_tap (const Oracle *o, const char *name);
void oracle_tap_shape (const Oracle *o, int *s3, int *s2, int *s1, int *tot);

#ifdef __cplusplus
}
#endif
#endif
