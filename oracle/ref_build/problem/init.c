/* ///////////////////////////////////////////////////////////////////// */
/*
  Problem file (PLUTO user API: Init / InitDomain / Analysis /
  UserDefBoundary) used by the parity oracle to drive the UNMODIFIED
  reference sources under /root/reference/Src for the five BASELINE.json
  configurations.  This is test infrastructure: it is only ever linked
  into oracle/_ref/pluto_* (see oracle/ref_build/build_ref.sh).

  One file serves all problems; the problem is selected at run time by
  the user parameter PROBLEM in pluto.ini ([Parameters] block), so that a
  single binary per compile-time scheme (dimensions / reconstruction)
  suffices:

    PROBLEM = 1  Orszag-Tang vortex        (2-D and 3-D variants)
    PROBLEM = 2  MHD blast wave            (parameters P_IN .. RADIUS)
    PROBLEM = 3  MHD rotor (Cartesian)
    PROBLEM = 4  decaying-turbulence box   (synthetic)

  The initial conditions are the classic textbook set-ups (Orszag & Tang
  1979; Balsara & Spicer 1999) with the parameter values the reference
  ships in Test_Problems/MHD/{Orszag_Tang,Blast,Rotor}.

  Analysis() is the full-precision time-step tap of SURVEY.md 8c: the
  reference log prints dt with 5 digits only, so every call appends the
  triple (step, t, dt) as raw doubles to "dt_tap.bin".  Its first call also
  writes the zone widths grid->dx[d] of every direction to "grid_tap.bin"
  (non-uniform grids: the patches of pluto.ini's [Grid] block).
*/
/* ///////////////////////////////////////////////////////////////////// */
#include "pluto.h"

/* ------------------------------------------------------------------
   Deterministic mode table for PROBLEM 4.
   Half-space integer wave vectors with 1 <= |k|^2 <= 4 (16 modes);
   amplitude ~ |k|^-2 times a uniform deviate in (-1,1) per component,
   phases uniform in [0,2pi), all drawn from splitmix64(seed).
   Since k and -k never both appear, the modes are orthogonal on the
   periodic box and the rms values are analytic:
     <v^2>   = sum_m 1/2 |a_m|^2
     <B^2>   = sum_m 1/2 |k_m x c_m|^2        (B = curl A)
   Amplitudes are rescaled so that v_rms = B_rms = 1.
   ------------------------------------------------------------------ */
#define TURB_MAXMODES 64
static int    turb_nm = -1;
static double turb_k[TURB_MAXMODES][3];
static double turb_av[TURB_MAXMODES][3], turb_pv[TURB_MAXMODES][3];
static double turb_aa[TURB_MAXMODES][3], turb_pa[TURB_MAXMODES][3];

static unsigned long long sm64_state;
static unsigned long long SplitMix64(void)
{
  unsigned long long z = (sm64_state += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static double Uniform01(void)
{
  return (double)(SplitMix64() >> 11) * (1.0/9007199254740992.0);
}

static void TurbSetup(void)
{
  int kx, ky, kz, c, m;
  double k2, amp, sv, sb, cx, cy, cz;

  sm64_state = (unsigned long long)(g_inputParam[SEED] + 0.5);
  turb_nm = 0;
  for (kz = -2; kz <= 2; kz++){
  for (ky = -2; ky <= 2; ky++){
  for (kx = -2; kx <= 2; kx++){
    k2 = kx*kx + ky*ky + kz*kz;
    if (k2 < 1 || k2 > 4) continue;
    /* keep one representative of each {k,-k} pair */
    if (kz < 0 || (kz == 0 && ky < 0) || (kz == 0 && ky == 0 && kx < 0)) continue;
    #if DIMENSIONS == 2
    if (kz != 0) continue;
    #endif
    m = turb_nm++;
    turb_k[m][0] = kx; turb_k[m][1] = ky; turb_k[m][2] = kz;
    amp = 1.0/k2;
    for (c = 0; c < 3; c++){
      turb_av[m][c] = amp*(2.0*Uniform01() - 1.0);
      turb_pv[m][c] = 2.0*CONST_PI*Uniform01();
      turb_aa[m][c] = amp*(2.0*Uniform01() - 1.0);
      turb_pa[m][c] = 2.0*CONST_PI*Uniform01();
    }
    #if DIMENSIONS == 2   /* 2-D: in-plane velocity, A = A_z only */
    turb_av[m][2] = 0.0;
    turb_aa[m][0] = turb_aa[m][1] = 0.0;
    #endif
  }}}

  /* The three components of a mode carry independent phases, hence the
     rms of B = curl A is evaluated component by component:
     B_x = d_y A_z - d_z A_y, ... each term an independent cosine.      */
  sv = sb = 0.0;
  for (m = 0; m < turb_nm; m++){
    for (c = 0; c < 3; c++) sv += 0.5*turb_av[m][c]*turb_av[m][c];
    cx = turb_aa[m][0]; cy = turb_aa[m][1]; cz = turb_aa[m][2];
    kx = (int)turb_k[m][0]; ky = (int)turb_k[m][1]; kz = (int)turb_k[m][2];
    sb += 0.5*( (ky*cz)*(ky*cz) + (kz*cy)*(kz*cy)
              + (kz*cx)*(kz*cx) + (kx*cz)*(kx*cz)
              + (kx*cy)*(kx*cy) + (ky*cx)*(ky*cx));
  }
  sv = 1.0/sqrt(sv);
  sb = 1.0/sqrt(sb);
  for (m = 0; m < turb_nm; m++) for (c = 0; c < 3; c++){
    turb_av[m][c] *= sv;
    turb_aa[m][c] *= sb;
  }
}

/* ********************************************************************* */
void Init (double *us, double x1, double x2, double x3)
/*
 *********************************************************************** */
{
  int    problem = (int)(g_inputParam[PROBLEM] + 0.5);
  double x = x1, y = x2, z = x3;

  g_gamma = g_inputParam[GAMMA_EOS];

  us[VX1] = us[VX2] = us[VX3] = 0.0;
  us[BX1] = us[BX2] = us[BX3] = 0.0;
  us[AX1] = us[AX2] = us[AX3] = 0.0;

  if (problem == 1){            /* ---------- Orszag-Tang ---------- */

    us[RHO] = 25./9.;
    us[PRS] = 5.0/3.0;
    #if DIMENSIONS == 2
    us[VX1] = - sin(y);
    us[VX2] =   sin(x);
    us[BX1] = - sin(y);
    us[BX2] =   sin(2.0*x);
    us[AX3] = cos(y) + 0.5*cos(2.0*x);
    #else
    {
      double c0 = 0.8;
      us[VX1] =   0.0;
      us[VX2] = - sin(z);
      us[VX3] =   sin(y);
      us[BX1] = c0*(       sin(y) + sin(z));
      us[BX2] = c0*( - 2.0*sin(2.0*z) + sin(x));
      us[BX3] = c0*(       sin(x) + sin(y));
      us[AX1] = c0*( cos(y) + cos(2.0*z));
      us[AX2] = c0*( cos(z) - cos(x));
      us[AX3] = c0*(-cos(y) + cos(x));
    }
    #endif

  }else if (problem == 2){      /* ---------- Blast wave ---------- */

    double r, theta, phi, B0;
    r = D_EXPAND(x1*x1, + x2*x2, + x3*x3);
    r = sqrt(r);
    us[RHO] = 1.0;
    us[PRS] = g_inputParam[P_OUT];
    if (r <= g_inputParam[RADIUS]) us[PRS] = g_inputParam[P_IN];
    theta = g_inputParam[THETA]*CONST_PI/180.0;
    phi   = g_inputParam[PHI]*CONST_PI/180.0;
    B0    = g_inputParam[BMAG];
    us[BX1] = B0*sin(theta)*cos(phi);
    us[BX2] = B0*sin(theta)*sin(phi);
    us[BX3] = B0*cos(theta);
    us[AX1] = 0.0;
    us[AX2] =  us[BX3]*x1;
    us[AX3] = -us[BX2]*x1 + us[BX1]*x2;

  }else if (problem == 3){      /* ---------- Rotor ---------- */

    double r, r0 = 0.1, r1 = 0.115, f, omega = 20.0;
    double Bx = 5.0/sqrt(4.0*CONST_PI);
    r = sqrt(x1*x1 + x2*x2);
    us[PRS] = 1.0;
    us[BX1] = Bx;
    f = (r1 - r)/(r1 - r0);
    if (r <= r0) {
      us[RHO] = 10.0;
      us[VX1] = -omega*x2;
      us[VX2] =  omega*x1;
    }else if (r < r1) {
      us[RHO] = 1.0 + 9.0*f;
      us[VX1] = -f*omega*x2*r0/r;
      us[VX2] =  f*omega*x1*r0/r;
    }else{
      us[RHO] = 1.0;
    }
    us[AX3] = Bx*x2;

  }else if (problem == 4){      /* ---------- decaying turbulence ---------- */

    int m, c;
    double ph, v[3], a[3];
    if (turb_nm < 0) TurbSetup();
    v[0] = v[1] = v[2] = a[0] = a[1] = a[2] = 0.0;
    for (m = 0; m < turb_nm; m++){
      ph = turb_k[m][0]*x + turb_k[m][1]*y + turb_k[m][2]*z;
      for (c = 0; c < 3; c++){
        v[c] += turb_av[m][c]*cos(ph + turb_pv[m][c]);
        a[c] += turb_aa[m][c]*cos(ph + turb_pa[m][c]);
      }
    }
    us[RHO] = 1.0;
    us[PRS] = 1.0;
    us[VX1] = v[0]; us[VX2] = v[1]; us[VX3] = v[2];
    us[AX1] = a[0]; us[AX2] = a[1]; us[AX3] = a[2];

  }else{
    print ("! Init(): unknown PROBLEM %d\n", problem);
    QUIT_PLUTO(1);
  }
}

/* ********************************************************************* */
void InitDomain (Data *d, Grid *grid)
/*
 *********************************************************************** */
{
}

/* ********************************************************************* */
void Analysis (const Data *d, Grid *grid)
/*
 * Full-precision (step, t, dt) tap.
 *********************************************************************** */
{
  double rec[3];
  FILE  *fp = fopen("dt_tap.bin", g_stepNumber <= 1 ? "wb" : "ab");
  if (fp == NULL) return;
  rec[0] = (double)g_stepNumber;
  rec[1] = g_time;
  rec[2] = g_dt;
  fwrite (rec, sizeof(double), 3, fp);
  fclose (fp);
  if (g_stepNumber <= 1){          /* zone widths of every direction, ghost zones included: grid->dx[d][0 .. np_tot-1] */
    int dir;
    fp = fopen("grid_tap.bin", "wb");
    if (fp == NULL) return;
    for (dir = 0; dir < DIMENSIONS; dir++){
      rec[0] = (double)grid->np_tot[dir];
      fwrite (rec, sizeof(double), 1, fp);
      fwrite (grid->dx[dir], sizeof(double), grid->np_tot[dir], fp);
    }
#if (UNIFORM_CARTESIAN_GRID == NO && RECONSTRUCTION == LINEAR) || (SHOCK_FLATTENING == MULTID && RECONSTRUCTION == PARABOLIC)
    for (dir = 0; dir < DIMENSIONS; dir++){     /* reconstruction weights of every direction: cp, cm, wp, wm, dp, dm (PARABOLIC +
                                                   MULTID: what the minmod fallback of flagged zones takes, ppm_states.c:167-181) */
      PLM_Coeffs c;
      PLM_CoefficientsGet (&c, dir);
      {                                         /* first and last zone are never set (plm_coeffs.c:62-64): written as 0 */
        double *six[6], zero = 0.0;
        int q, m = grid->np_tot[dir];
        six[0] = c.cp; six[1] = c.cm; six[2] = c.wp; six[3] = c.wm; six[4] = c.dp; six[5] = c.dm;
        rec[0] = -6.0;                          /* marker: six arrays of np_tot follow */
        fwrite (rec, sizeof(double), 1, fp);
        for (q = 0; q < 6; q++){
          fwrite (&zero, sizeof(double), 1, fp);
          fwrite (six[q] + 1, sizeof(double), m - 2, fp);
          fwrite (&zero, sizeof(double), 1, fp);
        }
      }
    }
#endif
#if RECONSTRUCTION == PARABOLIC
    for (dir = 0; dir < DIMENSIONS; dir++){     /* interface weights of the parabolic reconstruction, wp[i][-1 .. 2] of every zone
                                                   (PPM_ORDER 4, ppm_states.c:146-150; set for 1 <= i <= np_tot-3, ppm_coeffs.c:348-349,
                                                   analytic on a uniform direction): four arrays of np_tot, zeros where never set */
      PPM_Coeffs c;
      int q, i, m = grid->np_tot[dir];
      PPM_CoefficientsGet (&c, dir);
      rec[0] = -4.0;                            /* marker: four arrays of np_tot follow */
      fwrite (rec, sizeof(double), 1, fp);
      for (q = -1; q <= 2; q++) for (i = 0; i < m; i++){
        double w = (i >= 1 && i <= m - 3 ? c.wp[i][q] : 0.0);
        fwrite (&w, sizeof(double), 1, fp);
      }
    }
#endif
    fclose (fp);
  }
}

/* ********************************************************************* */
void UserDefBoundary (const Data *d, RBox *box, int side, Grid *grid)
/*
 *********************************************************************** */
{
}

#if BODY_FORCE != NO
/* ********************************************************************* */
void BodyForceVector(double *v, double *g, double x1, double x2, double x3)
/*
 * GRAV_MODE 0: uniform acceleration (the gravity of Rayleigh-Taylor-type set-ups).
 * GRAV_MODE 1: static, position-dependent: every component points towards (GRAV > 0: away
 *              from) the coordinate plane x_d = 0, with constant magnitude on either side.
 *********************************************************************** */
{
  g[IDIR] = g_inputParam[GRAV1];
  g[JDIR] = g_inputParam[GRAV2];
  g[KDIR] = g_inputParam[GRAV3];
  if ((int)(g_inputParam[GRAV_MODE] + 0.5) == 1){
    if (x1 < 0.0) g[IDIR] = -g[IDIR];
    if (x2 < 0.0) g[JDIR] = -g[JDIR];
    if (x3 < 0.0) g[KDIR] = -g[KDIR];
  }
}
/* ********************************************************************* */
double BodyForcePotential(double x1, double x2, double x3)
/*
 * Test potential (BODY_FORCE POTENTIAL): a step of height GRAV_d across the plane x_d = 0.013
 * (off every face and zone centre of the test grids), summed over the directions.
 *********************************************************************** */
{
  double phi = 0.0;
  if (x1 < 0.013) phi += g_inputParam[GRAV1];
  if (x2 < 0.013) phi += g_inputParam[GRAV2];
  if (x3 < 0.013) phi += g_inputParam[GRAV3];
  return phi;
}
#endif
