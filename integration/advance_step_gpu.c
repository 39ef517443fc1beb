/* advance_step_gpu.c -- the reference-side binding of libpluto_gpu.so.
 *
 * Drop-in replacement for Src/Time_Stepping/rk_step.c + update_stage.c (TIME_STEPPING RK2 /
 * RK3) or ctu_step.c (TIME_STEPPING HANCOCK; hancock.o stays on the link line, plm_states.o refers to it) in a
 * PLUTO 4.3 build: it defines the one symbol the driver calls,
 *
 *     int AdvanceStep (Data *d, Riemann_Solver *Riemann, timeStep *Dts, Grid *grid)
 *
 * (Src/prototypes.h:4, called from Integrate(), Src/main.c:339-355) and
 * forwards the step to the C ABI of include/pluto_gpu.h.  Everything else of
 * the reference (main loop, pluto.ini parser, Init(), output, restart) is
 * untouched.  Compile it IN PLACE OF rk_step.o/update_stage.o, against the
 * reference's own headers:
 *
 *     gcc -c -O3 -I. -I$PLUTO_DIR/Src -I<repo>/include advance_step_gpu.c
 *     gcc $(OBJ) advance_step_gpu.o -L<repo>/pluto_b200/lib -lpluto_gpu -lm -o pluto
 *
 * Data ownership: the host owns d->Vc / d->Vs.  Two modes:
 *   default             the literal AdvanceStep contract: upload d->Vc, d->Vs, step on the GPU, download the
 *                       result, every call (PCIe-bound) -- WriteData / Analysis / Restart work with no hooks;
 *   PLUTO_GPU_RESIDENT=1  the state stays in HBM: uploaded at the first call (after Startup / RestartFromFile),
 *                       advanced in place by pluto_gpu_advance, and copied back to d->Vc, d->Vs only when the
 *                       host reads it: PlutoGpuSyncHost(d), called from the two one-line hooks of INTEGRATION.md
 *                       (WriteData, Src/write_data.c:92, and CheckForAnalysis, Src/main.c:706).  A hook call is a
 *                       no-op when the host copy is current or the mode is off.
 * PLUTO_GPU_ARITH=fast|exact selects the arithmetic, PLUTO_GPU_DEVICE=<n> the (first) device ordinal (default 0).
 * PLUTO_GPU_NDEV=2|4|8: the domain is cut into that many blocks, one per GPU (round robin over the visible devices), all
 * driven from this single thread through pluto_gpu_multi_* -- the serial reference build uses the GPUs of a box with no MPI;
 * the rank grids are those of Src/Parallel for these counts (slowest index split first).
 */
#include "pluto.h"
#include "pluto_gpu.h"

static PlutoGpu *gpu = NULL;
static PlutoGpuMulti *gpum = NULL;   /* PLUTO_GPU_NDEV > 1: blocks on several devices instead of `gpu` */
static int gpu_resident = 0;         /* PLUTO_GPU_RESIDENT: the state lives in HBM between calls */
static int gpu_host_stale = 0;       /* resident mode: d->Vc, d->Vs are older than the device state */

static void StaggeredBase (const Data *d, double **vs1, double **vs2, double **vs3)
/* contiguous blocks behind the pointer tables (Src/arrays.c:222-330, Src/initialize.c:448-453) */
{
  *vs1 = &d->Vs[BX1s][0][0][-1];
  *vs2 = &d->Vs[BX2s][0][-1][0];
  *vs3 = NULL;
#if DIMENSIONS == 3
  *vs3 = &d->Vs[BX3s][-1][0][0];
#endif
}

/* ********************************************************************* */
void PlutoGpuSyncHost (Data *d)
/*
 * Resident mode: bring d->Vc, d->Vs up to date with the device state (interior and ghost zones) if they are
 * not.  Called by the host wherever it reads the arrays (WriteData, Analysis).
 *********************************************************************** */
{
  double *vs1, *vs2, *vs3;
  if ((gpu == NULL && gpum == NULL) || !gpu_resident || !gpu_host_stale) return;
  StaggeredBase (d, &vs1, &vs2, &vs3);
  if ((gpum ? pluto_gpu_multi_download_data (gpum, d->Vc[0][0][0], vs1, vs2, vs3)
            : pluto_gpu_download_data (gpu, d->Vc[0][0][0], vs1, vs2, vs3)) != 0){
    print ("! PlutoGpuSyncHost: %s\n", pluto_gpu_last_error());
    QUIT_PLUTO(1);
  }
  gpu_host_stale = 0;
  print ("> PlutoGpuSyncHost: state copied from the device at step %ld\n", g_stepNumber);
}

static int BoundaryCode (int type)
{
  if (type == PERIODIC)   return PLUTO_GPU_BC_PERIODIC;
  if (type == OUTFLOW)    return PLUTO_GPU_BC_OUTFLOW;
  if (type == REFLECTIVE) return PLUTO_GPU_BC_REFLECTIVE;
  if (type == EQTSYMMETRIC) return PLUTO_GPU_BC_EQTSYMMETRIC;
  print ("! AdvanceStep(gpu): boundary type %d is not supported by libpluto_gpu\n", type);
  QUIT_PLUTO(1);
  return -1;
}

/* ********************************************************************* */
int AdvanceStep (Data *d, Riemann_Solver *Riemann, timeStep *Dts, Grid *grid)
/*
 *********************************************************************** */
{
  PlutoGpuStepInfo info;
  double *vs1, *vs2, *vs3 = NULL;

#if PHYSICS != MHD || GEOMETRY != CARTESIAN || DIVB_CONTROL != CONSTRAINED_TRANSPORT \
    || EOS != IDEAL || DIMENSIONS != COMPONENTS \
    || (CT_EMF_AVERAGE != UCT_CONTACT && CT_EMF_AVERAGE != ARITHMETIC && CT_EMF_AVERAGE != UCT0 \
        && CT_EMF_AVERAGE != UCT_HLL)
  #error "libpluto_gpu covers ideal MHD, Cartesian, CT with UCT_CONTACT / ARITHMETIC / UCT0 / UCT_HLL, DIMENSIONS == COMPONENTS"
#endif
#if BACKGROUND_FIELD == YES || ENTROPY_SWITCH != NO || RESISTIVITY != NO || VISCOSITY != NO || THERMAL_CONDUCTION != NO \
    || HALL_MHD != NO || AMBIPOLAR_DIFFUSION != NO || ROTATING_FRAME != NO || COOLING != NO || FORCED_TURB != NO || NTRACER != 0 \
    || DIMENSIONAL_SPLITTING != NO || (defined SHEARINGBOX) || (defined FARGO) || (defined PARTICLES)
  #error "libpluto_gpu covers the ideal-MHD step only: no background field, entropy switch, diffusion terms, Hall / ambipolar, rotating frame, cooling, forced turbulence, tracers, dimensional splitting, shearing box, FARGO, particles"
#endif
#if TIME_STEPPING != RK2 && TIME_STEPPING != RK3 && TIME_STEPPING != HANCOCK && TIME_STEPPING != CHARACTERISTIC_TRACING
  #error "libpluto_gpu: TIME_STEPPING must be RK2, RK3, HANCOCK or CHARACTERISTIC_TRACING (EULER is not available on the GPU)"
#endif
#if INTERNAL_BOUNDARY == YES
  #error "libpluto_gpu: INTERNAL_BOUNDARY YES is not available (UserDefBoundary(d, NULL, 0, grid) / InternalBoundaryReset are not called on the GPU)"
#endif
#if UPDATE_VECTOR_POTENTIAL == YES
  #error "libpluto_gpu: UPDATE_VECTOR_POTENTIAL YES is not available (d->Ax1..3 are not advanced on the GPU)"
#endif
#if LIMITER == FOURTH_ORDER_LIM \
    || (SHOCK_FLATTENING != NO && SHOCK_FLATTENING != MULTID)
  #error "libpluto_gpu: FOURTH_ORDER_LIM and SHOCK_FLATTENING other than MULTID are not available on the GPU"
#endif
#if CHAR_LIMITING == YES && (DIMENSIONS != 2 || RECONSTRUCTION != LINEAR || (TIME_STEPPING != RK2 && TIME_STEPPING != RK3 && CT_EN_CORRECTION == YES) \
                             || SHOCK_FLATTENING != NO)
  #error "libpluto_gpu: CHAR_LIMITING YES is available in 2-D with LINEAR reconstruction, without SHOCK_FLATTENING (in 3-D the reference's own result depends on the sweep order: its eigenvector scratch is never cleared)"
#endif

  if (gpu == NULL && gpum == NULL){
    PlutoGpuConfig c;
    int nonuniform = 0;
    int ndev_blocks = getenv ("PLUTO_GPU_NDEV") ? atoi (getenv ("PLUTO_GPU_NDEV")) : 1;
    char *arith = getenv ("PLUTO_GPU_ARITH");
    int idim;
    memset (&c, 0, sizeof (c));
    c.dims = DIMENSIONS;
    c.n[0] = NX1; c.n[1] = NX2; c.n[2] = (DIMENSIONS == 3 ? NX3 : 1);
    c.recon  = (RECONSTRUCTION == PARABOLIC ? PLUTO_GPU_RECON_PARABOLIC : PLUTO_GPU_RECON_LINEAR);
    if      (Riemann == &HLLD_Solver) c.solver = PLUTO_GPU_SOLVER_HLLD;
    else if (Riemann == &HLL_Solver)  c.solver = PLUTO_GPU_SOLVER_HLL;
    else if (Riemann == &Roe_Solver)  c.solver = PLUTO_GPU_SOLVER_ROE;
    else if (Riemann == &HLLC_Solver) c.solver = PLUTO_GPU_SOLVER_HLLC;
    else if (Riemann == &LF_Solver)   c.solver = PLUTO_GPU_SOLVER_TVDLF;
    else{
      print ("! AdvanceStep(gpu): only hlld, hllc, hll, tvdlf and roe are available on the GPU\n");
      QUIT_PLUTO(1);
    }
    c.rk_order = (TIME_STEPPING == RK3 ? 3 : 2);
#if TIME_STEPPING == HANCOCK
  #if DIMENSIONAL_SPLITTING == YES || PRIMITIVE_HANCOCK != YES || CT_EMF_AVERAGE == UCT_HLL || RECONSTRUCTION != LINEAR
    #error "libpluto_gpu, TIME_STEPPING HANCOCK: unsplit, primitive predictor, LINEAR, CT_EMF_AVERAGE UCT_CONTACT / ARITHMETIC / UCT0"
  #endif
    c.time_stepping = PLUTO_GPU_TS_HANCOCK;                /* ctu_step.c */
#elif TIME_STEPPING == CHARACTERISTIC_TRACING
  #if DIMENSIONS != 2 || RECONSTRUCTION != LINEAR || SHOCK_FLATTENING != NO || BODY_FORCE != NO \
      || CT_EN_CORRECTION == YES || CT_EMF_AVERAGE == UCT_HLL || (defined CHTR_REF_STATE && CHTR_REF_STATE != 3)
    #error "libpluto_gpu, TIME_STEPPING CHARACTERISTIC_TRACING: 2-D, LINEAR, no SHOCK_FLATTENING / BODY_FORCE / CT_EN_CORRECTION, CT_EMF_AVERAGE UCT_CONTACT / ARITHMETIC / UCT0 (in 3-D the reference's own result depends on the sweep order: its eigenvector scratch is never cleared)"
  #endif
    c.time_stepping = PLUTO_GPU_TS_CHAR_TRACING;           /* ctu_step.c with char_tracing.c:278-560 as the predictor */
#endif
    /* LIMITER (plm_states.c:192-236) and CT_EMF_AVERAGE (ct_emf.c:241-283) of definitions.h */
    c.limiter = (LIMITER == FLAT_LIM      ? PLUTO_GPU_LIM_FLAT      : LIMITER == MINMOD_LIM ? PLUTO_GPU_LIM_MINMOD :
                 LIMITER == VANALBADA_LIM ? PLUTO_GPU_LIM_VANALBADA : LIMITER == OSPRE_LIM  ? PLUTO_GPU_LIM_OSPRE  :
                 LIMITER == UMIST_LIM     ? PLUTO_GPU_LIM_UMIST     : LIMITER == VANLEER_LIM ? PLUTO_GPU_LIM_VANLEER :
                 LIMITER == MC_LIM        ? PLUTO_GPU_LIM_MC        : PLUTO_GPU_LIM_DEFAULT);
    c.shock_flattening = (SHOCK_FLATTENING == MULTID);     /* flag_shock.c */
    c.en_correction = (CT_EN_CORRECTION == YES);           /* ct_field_average.c:116-129 */
    c.char_limiting = (CHAR_LIMITING == YES);              /* plm_states.c:448-706 */
#if BODY_FORCE & POTENTIAL
    c.body_force |= 2;                                     /* tabulated after the creation */
#endif
#if BODY_FORCE & VECTOR
    c.body_force |= 1;                                     /* tabulated per zone after the creation */
#endif
    c.emf_average = (CT_EMF_AVERAGE == ARITHMETIC ? PLUTO_GPU_EMF_ARITHMETIC :
                     CT_EMF_AVERAGE == UCT0 ? PLUTO_GPU_EMF_UCT0 :
                     CT_EMF_AVERAGE == UCT_HLL ? PLUTO_GPU_EMF_UCT_HLL : PLUTO_GPU_EMF_UCT_CONTACT);
    for (idim = 0; idim < DIMENSIONS; idim++){
      c.bc[2*idim]     = BoundaryCode (grid->lbound[idim]);      /* boundary.c:133-135 */
      c.bc[2*idim + 1] = BoundaryCode (grid->rbound[idim]);
      c.dx[idim] = grid->dx[idim][grid->lbeg[idim]];             /* uniform grid: set_grid.c:400 gives every zone */
      {                                                          /* of a uniform patch the same dx, bit for bit    */
        int ii;
        for (ii = 0; ii < grid->np_tot[idim]; ii++) if (grid->dx[idim][ii] != c.dx[idim]) nonuniform = 1;
      }
    }
    if (nonuniform){
      /* stretched / logarithmic / multi-patch grids (set_grid.c:330-560): the zone widths go to the library after its creation */
#if UNIFORM_CARTESIAN_GRID == NO && (SHOCK_FLATTENING != NO || CHAR_LIMITING == YES)
      print ("! AdvanceStep(gpu): a non-uniform grid with UNIFORM_CARTESIAN_GRID NO is available without SHOCK_FLATTENING and\n"
             "  CHAR_LIMITING on the GPU\n");
      QUIT_PLUTO(1);
#endif
    }
    c.arith    = (arith != NULL && !strcmp (arith, "fast")) ? PLUTO_GPU_ARITH_FAST : PLUTO_GPU_ARITH_EXACT;
    c.device   = getenv ("PLUTO_GPU_DEVICE") ? atoi (getenv ("PLUTO_GPU_DEVICE")) : 0;
    gpu_resident = getenv ("PLUTO_GPU_RESIDENT") != NULL && atoi (getenv ("PLUTO_GPU_RESIDENT")) != 0;
    c.gamma    = g_gamma;
    c.small_dn = g_smallDensity;
    c.small_pr = g_smallPressure;
    if (ndev_blocks > 1){
      /* rank grids as Src/Parallel would choose them for 2 / 4 / 8 processes, slowest index first */
      int g3[3] = {1, 1, 1}, devs[16], nvis = pluto_gpu_device_count (), b;
      if (DIMENSIONS == 3){ if (ndev_blocks >= 2) g3[2] = 2; if (ndev_blocks >= 4) g3[1] = 2; if (ndev_blocks >= 8) g3[0] = 2; }
      else                { if (ndev_blocks >= 2) g3[1] = 2; if (ndev_blocks >= 4) g3[0] = 2; if (ndev_blocks >= 8) g3[1] = 4; }
      if (g3[0]*g3[1]*g3[2] != ndev_blocks || nvis < 1){
        print ("! AdvanceStep(gpu): PLUTO_GPU_NDEV must be 2, 4 or 8 (and a CUDA device must be visible)\n");
        QUIT_PLUTO(1);
      }
      for (b = 0; b < ndev_blocks; b++) devs[b] = (c.device + b) % nvis;
      if (pluto_gpu_multi_create (&c, g3, devs, &gpum) != 0){
        print ("! AdvanceStep(gpu): %s\n", pluto_gpu_last_error());
        QUIT_PLUTO(1);
      }
    }else if (pluto_gpu_create (&c, &gpu) != 0){
      print ("! AdvanceStep(gpu): %s\n", pluto_gpu_last_error());
      QUIT_PLUTO(1);
    }
    if ((gpum ? pluto_gpu_multi_nghost (gpum) : pluto_gpu_nghost (gpu)) != grid->nghost[IDIR]){   /* the host arrays are read with this padding */
      print ("! AdvanceStep(gpu): the library expects %d ghost zones, the grid has %d (Src/get_nghost.c)\n",
             gpum ? pluto_gpu_multi_nghost (gpum) : pluto_gpu_nghost (gpu), grid->nghost[IDIR]);
      QUIT_PLUTO(1);
    }
    if (nonuniform && (gpum ? pluto_gpu_multi_set_grid (gpum, grid->dx[IDIR], grid->dx[JDIR], DIMENSIONS == 3 ? grid->dx[KDIR] : NULL)
                            : pluto_gpu_set_grid (gpu, grid->dx[IDIR], grid->dx[JDIR], DIMENSIONS == 3 ? grid->dx[KDIR] : NULL)) != 0){
      print ("! AdvanceStep(gpu): %s\n", pluto_gpu_last_error());
      QUIT_PLUTO(1);
    }
#if RECONSTRUCTION == PARABOLIC
    if (nonuniform){
      /* interface weights of the parabolic reconstruction as PPM_CoefficientsSet found them for this grid (ppm_coeffs.c:122-139,
         300-480): wp[i][-1 .. 2] of every zone, set for 1 <= i <= np_tot-3 */
      for (idim = 0; idim < DIMENSIONS; idim++){
        PPM_Coeffs qc;
        int np = grid->np_tot[idim], q, ii, rc;
        double *w4 = (double *)calloc ((size_t)4*np, sizeof(double));
        PPM_CoefficientsGet (&qc, idim);
        for (q = -1; q <= 2; q++) for (ii = 1; ii <= np - 3; ii++) w4[(q + 1)*np + ii] = qc.wp[ii][q];
        rc = (gpum ? pluto_gpu_multi_set_ppm_coeffs (gpum, idim, w4, w4 + np, w4 + 2*np, w4 + 3*np)
                   : pluto_gpu_set_ppm_coeffs (gpu, idim, w4, w4 + np, w4 + 2*np, w4 + 3*np));
        free (w4);
        if (rc != 0){
          print ("! AdvanceStep(gpu): %s\n", pluto_gpu_last_error());
          QUIT_PLUTO(1);
        }
      }
    }
#endif
#if (UNIFORM_CARTESIAN_GRID == NO && RECONSTRUCTION == LINEAR) || (SHOCK_FLATTENING == MULTID && RECONSTRUCTION == PARABOLIC)
    /* grid-dependent reconstruction weights: the arrays PLM_CoefficientsSet built for this grid (plm_coeffs.c:30-104); with
       PARABOLIC + MULTID they serve the minmod fallback of the flagged zones (ppm_states.c:167-181) */
    for (idim = 0; idim < DIMENSIONS; idim++){
      PLM_Coeffs pc;
      PLM_CoefficientsGet (&pc, idim);
      if ((gpum ? pluto_gpu_multi_set_plm_coeffs (gpum, idim, pc.cp, pc.cm, pc.wp, pc.wm, pc.dp, pc.dm)
                : pluto_gpu_set_plm_coeffs (gpu, idim, pc.cp, pc.cm, pc.wp, pc.wm, pc.dp, pc.dm)) != 0){
        print ("! AdvanceStep(gpu): %s\n", pluto_gpu_last_error());
        QUIT_PLUTO(1);
      }
    }
#elif UNIFORM_CARTESIAN_GRID == NO
  #error "libpluto_gpu: UNIFORM_CARTESIAN_GRID NO is available with LINEAR reconstruction only"
#endif
#if BODY_FORCE & POTENTIAL
    {                                 /* BodyForcePotential at the zone centres and at the faces of every direction
                                         (rhs.c:162-187), in the layouts of Vc and of the staggered arrays */
      size_t n1 = NX1_TOT, n2 = NX2_TOT, n3 = NX3_TOT, sz[4], off[4], q4;
      double *pt;
      int kk, jj, ii;
      sz[0] = n1*n2*n3; sz[1] = (n1 + 1)*n2*n3; sz[2] = n1*(n2 + 1)*n3; sz[3] = (DIMENSIONS == 3 ? n1*n2*(n3 + 1) : 0);
      off[0] = 0; for (q4 = 1; q4 < 4; q4++) off[q4] = off[q4 - 1] + sz[q4 - 1];
      pt = (double *)malloc ((off[3] + sz[3] + 1)*sizeof(double));
      for (kk = 0; kk < NX3_TOT; kk++) for (jj = 0; jj < NX2_TOT; jj++) for (ii = 0; ii < NX1_TOT; ii++)
        pt[off[0] + ((size_t)kk*n2 + jj)*n1 + ii] = BodyForcePotential (grid->x[IDIR][ii], grid->x[JDIR][jj], grid->x[KDIR][kk]);
      for (kk = 0; kk < NX3_TOT; kk++) for (jj = 0; jj < NX2_TOT; jj++) for (ii = -1; ii < NX1_TOT; ii++)
        pt[off[1] + ((size_t)kk*n2 + jj)*(n1 + 1) + (ii + 1)] =
          BodyForcePotential (ii < 0 ? grid->xl[IDIR][0] : grid->xr[IDIR][ii], grid->x[JDIR][jj], grid->x[KDIR][kk]);
      for (kk = 0; kk < NX3_TOT; kk++) for (jj = -1; jj < NX2_TOT; jj++) for (ii = 0; ii < NX1_TOT; ii++)
        pt[off[2] + ((size_t)kk*(n2 + 1) + (jj + 1))*n1 + ii] =
          BodyForcePotential (grid->x[IDIR][ii], jj < 0 ? grid->xl[JDIR][0] : grid->xr[JDIR][jj], grid->x[KDIR][kk]);
  #if DIMENSIONS == 3
      for (kk = -1; kk < NX3_TOT; kk++) for (jj = 0; jj < NX2_TOT; jj++) for (ii = 0; ii < NX1_TOT; ii++)
        pt[off[3] + ((size_t)(kk + 1)*n2 + jj)*n1 + ii] =
          BodyForcePotential (grid->x[IDIR][ii], grid->x[JDIR][jj], kk < 0 ? grid->xl[KDIR][0] : grid->xr[KDIR][kk]);
  #endif
      if ((gpum ? pluto_gpu_multi_set_body_potential (gpum, pt + off[0], pt + off[1], pt + off[2], DIMENSIONS == 3 ? pt + off[3] : NULL)
                : pluto_gpu_set_body_potential (gpu, pt + off[0], pt + off[1], pt + off[2], DIMENSIONS == 3 ? pt + off[3] : NULL)) != 0){
        print ("! AdvanceStep(gpu): %s\n", pluto_gpu_last_error());
        QUIT_PLUTO(1);
      }
      free (pt);
    }
#endif
#if BODY_FORCE & VECTOR
    {                                 /* BodyForceVector (init.c) tabulated at every zone centre, ghost zones included.
                                         The force must be static and must not depend on the state: every zone is probed
                                         with its own state and with a different one, and a force that answers differently
                                         is refused (rhs_source.c:214-345 would evaluate it with the stage's state) */
      size_t nz = (size_t)NX1_TOT*NX2_TOT*NX3_TOT, id = 0;
      double *gt = (double *)malloc (3*nz*sizeof(double)), g1[3], g2[3], v1[NVAR], v2[NVAR];
      int kk, jj, ii, q;
      for (kk = 0; kk < NX3_TOT; kk++) for (jj = 0; jj < NX2_TOT; jj++) for (ii = 0; ii < NX1_TOT; ii++){
        for (q = 0; q < NVAR; q++){ v1[q] = d->Vc[q][kk][jj][ii]; v2[q] = 1.75*v1[q] + 0.375; }
        g1[0] = g1[1] = g1[2] = 0.0; g2[0] = g2[1] = g2[2] = 0.0;
        BodyForceVector (v1, g1, grid->x[IDIR][ii], grid->x[JDIR][jj], grid->x[KDIR][kk]);
        BodyForceVector (v2, g2, grid->x[IDIR][ii], grid->x[JDIR][jj], grid->x[KDIR][kk]);
        if (g1[0] != g2[0] || g1[1] != g2[1] || g1[2] != g2[2]){
          print ("! AdvanceStep(gpu): BodyForceVector depends on the state (zone %d %d %d); only static forces are available on the GPU\n", ii, jj, kk);
          QUIT_PLUTO(1);
        }
        gt[id] = g1[0]; gt[nz + id] = g1[1]; gt[2*nz + id] = g1[2];
        id++;
      }
      if ((gpum ? pluto_gpu_multi_set_body_force (gpum, gt, gt + nz, DIMENSIONS == 3 ? gt + 2*nz : NULL)
                : pluto_gpu_set_body_force (gpu, gt, gt + nz, DIMENSIONS == 3 ? gt + 2*nz : NULL)) != 0){
        print ("! AdvanceStep(gpu): %s\n", pluto_gpu_last_error());
        QUIT_PLUTO(1);
      }
      free (gt);
    }
#endif
    print ("> AdvanceStep: libpluto_gpu (%s arithmetic), %d ghost zones, %d block(s) from device %d, state %s%s\n",
           c.arith == PLUTO_GPU_ARITH_FAST ? "fast" : "exact", grid->nghost[IDIR], ndev_blocks > 1 ? ndev_blocks : 1, c.device,
           gpu_resident ? "resident in HBM" : "on the host (upload + download per step)", nonuniform ? ", non-uniform grid" : "");
    if (gpu_resident){              /* the state the driver prepared (Startup or RestartFromFile) goes up once */
      StaggeredBase (d, &vs1, &vs2, &vs3);
      if ((gpum ? pluto_gpu_multi_upload_data (gpum, d->Vc[0][0][0], vs1, vs2, vs3)
                : pluto_gpu_upload_data (gpu, d->Vc[0][0][0], vs1, vs2, vs3)) != 0){
        print ("! AdvanceStep(gpu): %s\n", pluto_gpu_last_error());
        QUIT_PLUTO(1);
      }
    }
  }

  StaggeredBase (d, &vs1, &vs2, &vs3);
  if (gpu_resident){
    if ((gpum ? pluto_gpu_multi_advance (gpum, g_dt, &info) : pluto_gpu_advance (gpu, g_dt, &info)) != 0){
      print ("! AdvanceStep(gpu): %s\n", pluto_gpu_last_error());
      QUIT_PLUTO(1);
    }
    gpu_host_stale = 1;
  }else if ((gpum ? pluto_gpu_multi_advance_data (gpum, g_dt, d->Vc[0][0][0], vs1, vs2, vs3, &info)
                  : pluto_gpu_advance_data (gpu, g_dt, d->Vc[0][0][0], vs1, vs2, vs3, &info)) != 0){
    print ("! AdvanceStep(gpu): %s\n", pluto_gpu_last_error());
    QUIT_PLUTO(1);
  }
  if (info.nan_events > 0){
    print ("! AdvanceStep(gpu): %d zones are not finite\n", info.nan_events);
    QUIT_PLUTO(1);                       /* reference: CheckNaN, update_stage.c:192 */
  }
  Dts->invDt_hyp = MAX(Dts->invDt_hyp, info.inv_dt_hyp);   /* update_stage.c:308-312 */
  g_maxMach      = MAX(g_maxMach, info.max_mach);          /* hll_speed.c:105 */
  return 0;
}
