/*
 * pluto_gpu.h -- C ABI of the B200-native unsplit Godunov MHD step.
 *
 * This is the drop-in boundary for ONE path of PLUTO 4.3: the body of
 *
 *     int AdvanceStep (Data *d, Riemann_Solver *Riemann, timeStep *Dts, Grid *grid)
 *
 * (reference Src/prototypes.h:4, Src/Time_Stepping/rk_step.c:27-254, called
 * only from Integrate(), Src/main.c:339-355) together with everything it
 * calls per step: Boundary (Src/boundary.c:41), UpdateStage
 * (Src/Time_Stepping/update_stage.c:37), States (Src/States/plm_states.c:80,
 * ppm_states.c:68), the Riemann solvers (Src/MHD/hlld.c:44, hll.c:30,
 * roe.c:53, hllc.c:42, tvdlf.c:51), RightHandSide (Src/MHD/rhs.c:84), the constrained-transport
 * routines (Src/MHD/CT/ct_emf.c, ct_emf_average.c, ct_update.c,
 * ct_field_average.c, ct_fill_mag_field.c), the mappers
 * (Src/MHD/mappers.c, Src/mappers3D.c) and the CFL reduction feeding
 * NextTimeStep (Src/main.c:389).
 *
 * Plain C types only.  Every function returns 0 on success and a non-zero
 * code on failure (pluto_gpu_last_error() gives the text); the reference's
 * convention for fatal errors is print + QUIT_PLUTO (Src/macros.h:161-170),
 * which the reference-side shim maps these codes to.
 *
 * There is no CPU fallback: without a CUDA device pluto_gpu_create fails.
 */
#ifndef PLUTO_GPU_H
#define PLUTO_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

/* RECONSTRUCTION in definitions.h (Src/pluto.h:305-435): LINEAR / PARABOLIC */
enum { PLUTO_GPU_RECON_LINEAR = 0, PLUTO_GPU_RECON_PARABOLIC = 1 };
/* [Solver] in pluto.ini -> SetSolver (Src/MHD/set_solver.c:40-47) */
enum { PLUTO_GPU_SOLVER_HLLD = 0, PLUTO_GPU_SOLVER_HLL = 1, PLUTO_GPU_SOLVER_ROE = 2,
       PLUTO_GPU_SOLVER_HLLC = 3,      /* Src/MHD/hllc.c:42-236 */
       PLUTO_GPU_SOLVER_TVDLF = 4 };   /* Src/MHD/tvdlf.c:51-135 (Lax-Friedrichs / Rusanov) */
/* [Boundary] in pluto.ini (Src/boundary.c:171-218).  SHARED marks a side
   that abuts another rank's block: it is filled by the halo exchange
   (reference: AL_Exchange_dim, Src/Parallel/al_exchange_dim.c:25) instead
   of a physical condition, exactly as boundary.c:139 skips it. */
enum { PLUTO_GPU_BC_PERIODIC = 0, PLUTO_GPU_BC_OUTFLOW = 1, PLUTO_GPU_BC_REFLECTIVE = 2,
       PLUTO_GPU_BC_SHARED = 3,
       PLUTO_GPU_BC_EQTSYMMETRIC = 4 /* reflective, with the sign table of Src/boundary.c:333-336, 423-427: the normal
                                        velocity and the TRANSVERSE field components change sign */ };
/* arithmetic mode.  EXACT: no FMA contraction, IEEE div/sqrt in the
   reference's operation order (bit-identical to the gcc -O3 x86-64 build
   of the reference).  FAST: same algorithm with FMA contraction and shared
   reciprocals; within the BASELINE.json tolerances, not bit-identical. */
enum { PLUTO_GPU_ARITH_EXACT = 0, PLUTO_GPU_ARITH_FAST = 1 };
/* LIMITER in definitions.h (Src/States/plm_states.c:192-236, plm_coeffs.h:72-123; LINEAR
   reconstruction only).  DEFAULT: MC on density, van Leer on velocity and field, minmod on
   pressure; any other value applies that limiter to every variable. */
enum { PLUTO_GPU_LIM_DEFAULT = 0, PLUTO_GPU_LIM_FLAT, PLUTO_GPU_LIM_MINMOD, PLUTO_GPU_LIM_VANALBADA,
       PLUTO_GPU_LIM_OSPRE, PLUTO_GPU_LIM_UMIST, PLUTO_GPU_LIM_VANLEER, PLUTO_GPU_LIM_MC };
/* CT_EMF_AVERAGE in definitions.h (Src/MHD/CT/ct_emf.c:241-283) */
enum { PLUTO_GPU_EMF_UCT_CONTACT = 0, PLUTO_GPU_EMF_ARITHMETIC = 1, PLUTO_GPU_EMF_UCT0 = 2,
       PLUTO_GPU_EMF_UCT_HLL = 3 /* the reference's default, Src/MHD/CT/ct.h:43-45 */ };
/* TIME_STEPPING in definitions.h.  RK: rk_step.c (RK2 / RK3 by rk_order).  HANCOCK: the unsplit
   corner-transport-upwind step of Src/Time_Stepping/ctu_step.c:142-727 with the primitive
   MUSCL-Hancock predictor (Src/States/hancock.c:33-142) and CTU_CT_Source (ctu_step.c:731-816);
   LINEAR reconstruction, one more ghost zone (Src/get_nghost.c:86-90), one Boundary call per
   step; CT_EMF_AVERAGE UCT_CONTACT, ARITHMETIC or UCT0. */
enum { PLUTO_GPU_TS_RK = 0, PLUTO_GPU_TS_HANCOCK = 1,
       PLUTO_GPU_TS_CHAR_TRACING = 2 };  /* corner transport upwind with the characteristic-tracing predictor
                                            (Src/States/char_tracing.c:278-560), 2-D, LINEAR */

typedef struct {
  int    dims;         /* DIMENSIONS = COMPONENTS: 2 or 3                     */
  int    n[3];         /* interior zones of THIS block (NX1,NX2,NX3; n[2]=1 in 2-D) */
  int    recon;        /* PLUTO_GPU_RECON_*                                   */
  int    solver;       /* PLUTO_GPU_SOLVER_*                                  */
  int    rk_order;     /* TIME_STEPPING: 2 = RK2, 3 = RK3                     */
  int    bc[6];        /* X1_BEG, X1_END, X2_BEG, X2_END, X3_BEG, X3_END      */
  int    arith;        /* PLUTO_GPU_ARITH_*                                   */
  int    device;       /* CUDA device ordinal                                 */
  double gamma;        /* g_gamma          (Src/globals.h:116)                */
  double dx[3];        /* uniform cell sizes grid[d].dx[i]                    */
  double small_dn;     /* g_smallDensity   (Src/globals.h:113)                */
  double small_pr;     /* g_smallPressure                                     */
  int    limiter;      /* PLUTO_GPU_LIM_*  (0 = DEFAULT)                      */
  int    emf_average;  /* PLUTO_GPU_EMF_*  (0 = UCT_CONTACT)                  */
  int    shock_flattening; /* SHOCK_FLATTENING: 0 NO, 1 MULTID (Src/flag_shock.c:79-230;
                              LINEAR reconstruction only)                      */
  int    time_stepping;    /* PLUTO_GPU_TS_* (0 = RK2 / RK3 by rk_order)       */
  int    en_correction;    /* CT_EN_CORRECTION YES: total energy redefined with the face-averaged field
                              (Src/MHD/CT/ct_field_average.c:116-129);
                              CT_EMF_AVERAGE other than UCT_HLL                */
  int    body_force;       /* BODY_FORCE: bit 0 VECTOR, bit 1 POTENTIAL (pluto_gpu_set_body_potential).
                              VECTOR with a UNIFORM acceleration grav[] (what BodyForceVector of
                              init.c returns everywhere): momentum and energy sources of
                              Src/MHD/rhs_source.c:214-217, 277-280, 342-345 and, with HANCOCK, the
                              predictor source of Src/MHD/prim_eqn.c:289-360.  Not with UCT_HLL or
                              SHOCK_FLATTENING.  A static position-dependent force:
                              pluto_gpu_set_body_force                         */
  double grav[3];
  int    char_limiting;    /* CHAR_LIMITING YES (Src/States/plm_states.c:448-706, PrimEigenvectors Src/MHD/eigenv.c:190-470):
                              slopes limited on the characteristic variables.  2-D (DIMENSIONS = COMPONENTS = 2), LINEAR,
                              RK2 / RK3 and the corner-transport-upwind steps, no SHOCK_FLATTENING.  Not in 3-D: the reference's
                              eigenvector scratch keeps entries of the previous sweep direction there, its own result
                              depends on the sweep order                        */
} PlutoGpuConfig;

typedef struct PlutoGpu PlutoGpu;

/* per-step scalars the host needs (reference: Dts->invDt_hyp set at
   update_stage.c:308-312, g_maxMach at hll_speed.c:105) */
typedef struct {
  double inv_dt_hyp;
  double max_mach;
  int    floor_events;     /* ConsToPrim repairs (Src/MHD/mappers.c:130-200)  */
  int    nan_events;       /* zones whose updated state is not finite
                              (reference: CheckNaN, update_stage.c:192)       */
} PlutoGpuStepInfo;

int  pluto_gpu_create   (const PlutoGpuConfig *cfg, PlutoGpu **out);
void pluto_gpu_destroy  (PlutoGpu *h);
const char *pluto_gpu_last_error (void);
int  pluto_gpu_nghost   (const PlutoGpu *h);     /* Src/get_nghost.c:32-50 */
/* BODY_FORCE VECTOR with a STATIC position-dependent force instead of the uniform grav[] of the configuration (which must
   have body_force = 1): component d of BodyForceVector (init.c) at every zone centre, ghost zones included, HOST arrays
   g_d[k][j][i] with the extents T3 x T2 x T1 of the reference's Data arrays (g3 NULL in 2-D). */
/* Non-uniform Cartesian grid (uniform and stretched patches of pluto.ini's [Grid] block, Src/set_grid.c:330-560): the zone
   widths of every direction, dx_d[0 .. T_d-1] = grid->dx[d] with the ghost zones (T_d = n[d] + 2 nghost).  The reference's
   CARTESIAN builds keep the uniform reconstruction weights (UNIFORM_CARTESIAN_GRID YES, Src/States/plm_coeffs.h:23-29); the
   widths enter Src/MHD/rhs.c:195, the inverse time step (Src/Time_Stepping/update_stage.c:229-235), Src/MHD/CT/ct_update.c:91-204
   and the face areas of Src/MHD/CT/ct_fill_mag_field.c:108-114; with the corner-transport-upwind steps also the predictors
   (Src/States/hancock.c:245-296, char_tracing.c:347-362), the transverse correction of ctu_step.c:310, 731-785 and the potential's
   source of Src/MHD/prim_eqn.c:304-307; with SHOCK_FLATTENING MULTID Src/flag_shock.c:143-145.  PARABOLIC reconstruction
   needs pluto_gpu_set_ppm_coeffs as well; dx3 may be NULL in 2-D.
   PlutoGpuConfig.dx is then used by nothing on the path.  Call once after pluto_gpu_create. */
int  pluto_gpu_set_grid (PlutoGpu *h, const double *dx1, const double *dx2, const double *dx3);
/* UNIFORM_CARTESIAN_GRID NO (Src/States/plm_coeffs.h:23-29): grid-dependent weights of the linear reconstruction.  Hand over, for
   every direction, the six arrays of PLM_CoefficientsGet (Src/States/plm_coeffs.c:86-104: cp, cm, wp, wm, dp, dm; T_dir entries each)
   after pluto_gpu_create (and pluto_gpu_set_grid on a non-uniform grid).  RK2 / RK3, LINEAR, plain scheme options. */
int  pluto_gpu_set_plm_coeffs (PlutoGpu *h, int dir, const double *cp, const double *cm, const double *wp, const double *wm,
                               const double *dp, const double *dm);
/* PARABOLIC reconstruction on a non-uniform grid: the interface weights of Src/States/ppm_states.c:146-150 as PPM_CoefficientsGet
   returns them (Src/States/ppm_coeffs.c:586-609; PPM_FindWeights, :300-480, where a direction is not uniform) -- wp[i][-1], wp[i][0],
   wp[i][1], wp[i][2] of every zone i of direction dir, T_dir entries each (set for 1 <= i <= T_dir - 3).  hp = hm = 3 on every
   Cartesian grid (ppm_coeffs.c:544-547).  Required after pluto_gpu_set_grid with PARABOLIC; RK2 / RK3. */
int  pluto_gpu_set_ppm_coeffs (PlutoGpu *h, int dir, const double *wm1, const double *w0, const double *w1, const double *w2);
int  pluto_gpu_set_body_force (PlutoGpu *h, const double *g1, const double *g2, const double *g3);
/* BODY_FORCE POTENTIAL (body_force & 2; Src/MHD/rhs.c:162-187, 388-392, rhs_source.c:233-237, 316-320, 358-362,
   prim_eqn.c:304-307): BodyForcePotential (init.c) at the zone centres, phic[k][j][i] (T3 x T2 x T1), and at the faces of
   every direction in the layout of the staggered Data arrays (pf1: T3 x T2 x (T1+1) starting with face -1/2, pf2:
   T3 x (T2+1) x T1, pf3: (T3+1) x T2 x T1, NULL in 2-D).  HOST arrays; must be called before the first step. */
int  pluto_gpu_set_body_potential (PlutoGpu *h, const double *phic, const double *pf1, const double *pf2, const double *pf3);
int  pluto_gpu_nstages  (const PlutoGpu *h);     /* Boundary calls (= halo exchanges) per step: rk_order, 1 with HANCOCK */

/* ---- state transfer ---------------------------------------------------
   "interior" layout = the reference's .dbl dump layout (Src/bin_io.c:216):
   vc[nv][k][j][i] with nv = rho,vx1,vx2,vx3,Bx1,Bx2,Bx3,prs (always 8
   slots; vx3/Bx3 ignored in 2-D), n3*n2*n1 each; bx1s[k][j][i] with n1+1
   faces, bx2s with n2+1, bx3s with n3+1 (NULL in 2-D).
   "data" layout = the reference's Data arrays including ghost zones
   (Src/structs.h:29-88, Src/arrays.c:222-330): Vc is one block
   [NVAR][T3][T2][T1] (NVAR = 8 in 3-D, 6 in 2-D: rho,vx1,vx2,Bx1,Bx2,prs);
   Vs[d] is the block whose base is &Vs[d][0][0][-1] etc.
   (Src/initialize.c:448-453): T3 x T2 x (T1+1), T3 x (T2+1) x T1,
   (T3+1) x T2 x T1.  Host pointers. */
int pluto_gpu_upload_interior   (PlutoGpu *h, const double *vc, const double *bx1s,
                                 const double *bx2s, const double *bx3s);
int pluto_gpu_download_interior (PlutoGpu *h, double *vc, double *bx1s,
                                 double *bx2s, double *bx3s);
int pluto_gpu_upload_data       (PlutoGpu *h, const double *Vc, const double *Vs1,
                                 const double *Vs2, const double *Vs3);
int pluto_gpu_download_data     (PlutoGpu *h, double *Vc, double *Vs1,
                                 double *Vs2, double *Vs3);

/* ---- the step ---------------------------------------------------------
   pluto_gpu_advance: one AdvanceStep on the device-resident state.
   pluto_gpu_advance_data: the literal AdvanceStep contract on HOST Data
   arrays: upload, step, download (ghost zones come back filled as the
   reference leaves them after the last Boundary call is NOT guaranteed;
   interior zones and interior faces are). */
int pluto_gpu_advance      (PlutoGpu *h, double dt, PlutoGpuStepInfo *info);
int pluto_gpu_advance_data (PlutoGpu *h, double dt, double *Vc, double *Vs1,
                            double *Vs2, double *Vs3, PlutoGpuStepInfo *info);

/* Boundary(d, ALL_DIR, grid) on the current state (reference
   Src/startup.c:286 calls it once before the first output). */
int pluto_gpu_boundary (PlutoGpu *h);

/* NextTimeStep, hyperbolic part (Src/main.c:462-465, 532). Host only. */
double pluto_gpu_next_dt (double inv_dt_hyp, double cfl, double cfl_max_var, double dt);

/* ---- NextTimeStep on the device ------------------------------------------
   The same rule evaluated by a one-thread kernel at the end of every step, dt kept in device
   memory: the host enqueues steps back to back and never waits for the CFL reduction (the
   reference's loop main.c:133-243 with Integrate -> NextTimeStep, without the round trip).
       pluto_gpu_set_dt (h, first_dt)
       repeat:  pluto_gpu_advance_async (h, cfl, cfl_max_var)
       pluto_gpu_sync_results (h, max, infos, dts, &n, &dt_next)   wait; dt used and StepInfo of every
                                                                   step since the last call, the dt the
                                                                   next step will use
   Multi-GPU: per stage as above with pluto_gpu_stage (h, stage, -1.0) (a negative dt keeps the
   device's), then all-reduce(MAX) the two doubles at pluto_gpu_reduction_slots (bit patterns of
   non-negative doubles, so an integer or floating maximum both work) on pluto_gpu_stream(h) and call
   pluto_gpu_next_dt_async.  At most 4096 steps between two synchronisations. */
int pluto_gpu_set_dt          (PlutoGpu *h, double dt);
int pluto_gpu_advance_async   (PlutoGpu *h, double cfl, double cfl_max_var);
int pluto_gpu_next_dt_async   (PlutoGpu *h, double cfl, double cfl_max_var);
int pluto_gpu_reduction_slots (PlutoGpu *h, void **dev_ptr);
int pluto_gpu_sync_results    (PlutoGpu *h, int max_steps, PlutoGpuStepInfo *infos, double *dts,
                               int *n_out, double *dt_next);

/* ---- multi-GPU halo exchange (replaces AL_Exchange_dim) ----------------
   Sides flagged PLUTO_GPU_BC_SHARED abut another rank's block and are filled
   by the caller: the step is split so that the host language can drive the
   exchange (NCCL send/recv, peer copies) dimension by dimension:
       pluto_gpu_step_begin (h)
       for stage in 1..rk_order:
         for dim in 0..dims-1:
            pluto_gpu_halo_pack   (h, stage, dim, send_lo, send_hi)
            <send_lo -> low neighbour, send_hi -> high neighbour along dim>
            pluto_gpu_halo_unpack (h, stage, dim, recv_lo, recv_hi)
            pluto_gpu_boundary_dim (h, stage, dim)     physical sides of dim
         pluto_gpu_stage (h, stage, dt)
       pluto_gpu_step_end (h, &info)     then max-reduce info over the ranks
   Buffers are DEVICE pointers of pluto_gpu_halo_doubles(h, dim) doubles;
   a NULL buffer skips that side.  recv_lo is what the low neighbour packed
   as its send_hi, and vice versa.  All work is enqueued on
   pluto_gpu_stream(h).  The sequential x1 -> x2 -> x3 order fills edges and
   corners exactly as the reference's AL_Exchange_dim loop does
   (Src/Parallel/al_exchange_dim.c:58-88, al_decompose.c:218-229). */
long long pluto_gpu_halo_doubles (const PlutoGpu *h, int dim);
int pluto_gpu_halo_pack    (PlutoGpu *h, int stage, int dim, double *send_lo, double *send_hi);
int pluto_gpu_halo_unpack  (PlutoGpu *h, int stage, int dim, const double *recv_lo, const double *recv_hi);
int pluto_gpu_boundary_dim (PlutoGpu *h, int stage, int dim);
/* All-neighbour variant (one pack launch, one communication group, one unpack
   launch per stage): the caller lists its neighbours by block offset
   (-1/0/+1 per dimension, up to 26 in 3-D; offsets = n_nbr x 3 ints) with one
   send and one receive DEVICE buffer of pluto_gpu_halo_nbr_doubles() doubles
   each.  What a block packs for offset o is what its neighbour unpacks for -o.
   Per stage: pluto_gpu_halo_pack_all -> exchange -> pluto_gpu_halo_unpack_all ->
   pluto_gpu_boundary_dim for every dimension -> pluto_gpu_stage. */
long long pluto_gpu_halo_nbr_doubles (const PlutoGpu *h, const int off[3]);
int pluto_gpu_halo_plan       (PlutoGpu *h, int n_nbr, const int *offsets,
                               double *const *send_bufs, double *const *recv_bufs);
int pluto_gpu_halo_pack_all   (PlutoGpu *h, int stage);
/* The plan of ONE stage (its state buffer) only: a host may give every stage its own send / receive buffers. */
int pluto_gpu_halo_plan_stage (PlutoGpu *h, int stage, int n_nbr, const int *offsets,
                               double *const *send_bufs, double *const *recv_bufs);
/* Ghost zones stored directly into the neighbour GPU's memory over NVLink (replaces the MPI_Sendrecv of
 * Src/Parallel/al_exchange_dim.c:58-88 for ranks on one node): a rank allocates its receive arena with pluto_gpu_ipc_alloc
 * and publishes the 64-byte handle; a neighbouring process maps it (pluto_gpu_ipc_open) and passes the mapped addresses as
 * SEND buffers of pluto_gpu_halo_plan_stage, so the pack launch is the transfer.  pluto_gpu_halo_signal (after the pack
 * launch, same stream) stores the exchange number into one 64-bit counter per peer (addresses inside the peers' arenas);
 * pluto_gpu_halo_wait (before the unpack launch) spins on the device until the `n` counters of this rank, consecutive at
 * `counters`, have reached it.  stream == NULL: the block's own stream. */
int pluto_gpu_ipc_alloc   (int device, size_t bytes, void **ptr, unsigned char handle[64]);
int pluto_gpu_ipc_open    (int device, const unsigned char handle[64], void **ptr);
int pluto_gpu_ipc_close   (void *ptr);
int pluto_gpu_ipc_free    (void *ptr);
int pluto_gpu_halo_signal (PlutoGpu *h, void *stream, int n, unsigned long long *const *peer_counters, unsigned long long value);
int pluto_gpu_halo_wait   (PlutoGpu *h, void *stream, int n, const unsigned long long *counters, unsigned long long value);
int pluto_gpu_halo_unpack_all (PlutoGpu *h, int stage);
/* Overlap of the exchange with computation: a stage can be issued in two parts,
       pluto_gpu_stage_shell    (h, stage, dt)   sweeps, CT and the new state of the zones within
                                                 nghost of a SHARED side (what the neighbours need)
       <record an event on pluto_gpu_stream(h); on a second stream, after that event:
        pluto_gpu_halo_pack_all_on (h, next_stage, stream2) and the exchange>
       pluto_gpu_stage_interior (h, stage)       the new state of all other zones
   so that the ghost zones of the NEXT stage travel (NVLink) while the interior of this one is
   completed; the next stage then waits for the exchange, unpacks and carries on.
   pluto_gpu_stage == shell followed by interior. */
int pluto_gpu_stage_shell    (PlutoGpu *h, int stage, double dt);
int pluto_gpu_stage_interior (PlutoGpu *h, int stage);
int pluto_gpu_halo_pack_all_on (PlutoGpu *h, int stage, void *stream);
int pluto_gpu_step_begin   (PlutoGpu *h);
int pluto_gpu_stage        (PlutoGpu *h, int stage, double dt);
int pluto_gpu_step_end     (PlutoGpu *h, PlutoGpuStepInfo *info);

/* ---- output, restart and run-time diagnostics from the device state -----
   pluto_gpu_write_dbl: data.NNNN.dbl of the reference's "dbl ... single_file" output
   (Src/write_data.c:92-205, Src/bin_io.c:216): for each of rho vx1 vx2 [vx3] Bx1 Bx2 [Bx3]
   prs the interior zones [k][j][i], then Bx1s, Bx2s, [Bx3s] with their extra face, little
   endian doubles; the matching line "nfile t dt nstep single_file little names..." is
   appended to (nfile == 0: starts) <dir>/dbl.out (write_data.c:365-395), so the
   reference's own tools (pyPLUTO, restart) read the files.
   pluto_gpu_read_dbl: the inverse (what RestartFromFile does for d->Vc, d->Vs).
   pluto_gpu_analysis: volume integrals of the state for Analysis()-type diagnostics, summed
   on the device in a fixed order (deterministic): out[0] mass, [1] kinetic, [2] magnetic,
   [3] thermal energy (p/(gamma-1)), [4..6] momentum, [7] max |div B| (staggered field). */
int pluto_gpu_write_dbl (PlutoGpu *h, const char *dir, int nfile, double t, double dt, long nstep);
int pluto_gpu_read_dbl  (PlutoGpu *h, const char *path);
/* Single-precision output of the cell-centred variables from the device state, in the reference's formats and file lists:
 * data.NNNN.flt + flt.out (Src/write_data.c:178-206, Convert_dbl2flt Src/bin_io.c:51) and the legacy-VTK rectilinear-grid
 * file data.NNNN.vtk + vtk.out (Src/write_vtk.c:92-351: header, big-endian node coordinates xl1/xl2/xl3 with n+1 entries
 * each -- grid->xl_glob of the interior, x3 NULL in 2-D -- and one SCALARS block per variable; VTK_VECTOR_DUMP NO,
 * VTK_TIME_INFO NO).  The conversion (and the byte swap) runs on the device; floats cross PCIe. */
int pluto_gpu_write_flt (PlutoGpu *h, const char *dir, int nfile, double t, double dt, long nstep);
int pluto_gpu_write_vtk (PlutoGpu *h, const char *dir, int nfile, double t, double dt, long nstep,
                         const double *xl1, const double *xl2, const double *xl3);
int pluto_gpu_analysis  (PlutoGpu *h, double out[8]);

/* ---- introspection (tests, bench, profiling) -------------------------- */
void     *pluto_gpu_stream        (PlutoGpu *h);   /* cudaStream_t of all launches */
long long pluto_gpu_launch_count  (const PlutoGpu *h);  /* kernels launched so far */
/* per-kernel-class device time: CUDA events recorded on pluto_gpu_stream(h)
   around every launch while enabled; accumulated at pluto_gpu_step_end.
   Classes 0..7: sweep_x1 sweep_x2 sweep_x3 ct_emf ct_update final boundary halo;
   with FAST arithmetic class 0 is the fused x1+x2 sweep ("sweep_x1x2") and class 1 stays empty. */
int pluto_gpu_timing     (PlutoGpu *h, int enable);      /* (re)starts the accumulation */
int pluto_gpu_timing_get (PlutoGpu *h, int cls, const char **name, double *ms, long long *launches);
long long pluto_gpu_device_bytes  (const PlutoGpu *h);
/* raw DEVICE pointer + padded shape of an internal field, for debugging:
   names rho vx1 vx2 vx3 bx1 bx2 bx3 prs bx1s bx2s bx3s.  Element (k,j,i),
   valid from -1, lives at ((k+off[2])*shape[1] + (j+off[1]))*shape[0] + (i+off[0]). */
int pluto_gpu_field (PlutoGpu *h, const char *name, double **dev_ptr,
                     long long shape[3], int off[3]);
/* copy the whole padded array of a field to host memory (shape[0]*shape[1]*shape[2]
   doubles).  Further names: u_rho u_mx1 u_mx2 u_mx3 u_eng, exj exk eyi eyk ezi ezj,
   ex ey ez, cdt; a "1:" / "2:" prefix selects the stage buffers. */
int pluto_gpu_read_field (PlutoGpu *h, const char *name, double *host);

/* ---- several GPUs from ONE host thread ----------------------------------------------------------------------
 * The reference's host is a single-threaded, non-re-entrant C program; without MPI it still gets the GPUs
 * of a box: pluto_gpu_multi_create cuts the domain of cfg (n = zones of the WHOLE domain, bc = the physical conditions;
 * cfg->device is ignored) into grid[0] x grid[1] x grid[2] blocks numbered as Src/Parallel/al_decompose.c does, block b on
 * devices[b] (NULL: device b).  Ghost zones travel by peer stores between the devices inside the pack launch, ordered by
 * events; one host thread issues everything.  The host arrays of upload / download / advance_data are the reference's
 * Data arrays of the whole domain (layouts as pluto_gpu_upload_data); download refreshes the interior zones and faces.
 * Replaces AL_Decompose / AL_Exchange_dim / MPI_Allreduce(MAX) (Src/Parallel/al_decompose.c, al_exchange_dim.c:25-96,
 * Src/main.c:195-199, 415) for a host that is not an MPI program. */
typedef struct PlutoGpuMulti PlutoGpuMulti;
int  pluto_gpu_device_count (void);                          /* CUDA devices visible to this process */
int  pluto_gpu_multi_create (const PlutoGpuConfig *cfg, const int grid[3], const int *devices, PlutoGpuMulti **out);
void pluto_gpu_multi_destroy (PlutoGpuMulti *m);
int  pluto_gpu_multi_nghost (const PlutoGpuMulti *m);
int  pluto_gpu_multi_nblocks (const PlutoGpuMulti *m);
int  pluto_gpu_multi_upload_data (PlutoGpuMulti *m, const double *Vc, const double *Vs1, const double *Vs2, const double *Vs3);
int  pluto_gpu_multi_download_data (PlutoGpuMulti *m, double *Vc, double *Vs1, double *Vs2, double *Vs3);
/* pluto_gpu_set_grid / pluto_gpu_set_plm_coeffs for the blocks: arrays of the WHOLE domain (gn[d] + 2 nghost entries), sliced per block */
int  pluto_gpu_multi_set_grid (PlutoGpuMulti *m, const double *dx1, const double *dx2, const double *dx3);
int  pluto_gpu_multi_set_plm_coeffs (PlutoGpuMulti *m, int dir, const double *cp, const double *cm, const double *wp, const double *wm,
                                     const double *dp, const double *dm);
int  pluto_gpu_multi_set_ppm_coeffs (PlutoGpuMulti *m, int dir, const double *wm1, const double *w0, const double *w1, const double *w2);
/* pluto_gpu_set_body_force / pluto_gpu_set_body_potential for the blocks: HOST arrays of the WHOLE domain in the same layouts
   (ghost zones included; the face arrays with one more entry along their direction), cut into the blocks' pieces */
int  pluto_gpu_multi_set_body_force (PlutoGpuMulti *m, const double *g1, const double *g2, const double *g3);
int  pluto_gpu_multi_set_body_potential (PlutoGpuMulti *m, const double *phic, const double *pf1, const double *pf2, const double *pf3);
int  pluto_gpu_multi_advance (PlutoGpuMulti *m, double dt, PlutoGpuStepInfo *info);
int  pluto_gpu_multi_advance_data (PlutoGpuMulti *m, double dt, double *Vc, double *Vs1, double *Vs2, double *Vs3,
                                   PlutoGpuStepInfo *info);

/* FP64 pipe microbenchmark (independent DFMA chains on every SM): the measured
   denominator of the FP64 roofline, in TFLOP/s (FMA = 2 flops). */
int pluto_gpu_measure_fp64 (int device, double *tflops);
/* Self-test of the arithmetic of the FAST Roe kernels: their branch-free division, reciprocal and square root must be the
 * IEEE results (the Roe solver compares square roots, Src/MHD/roe.c:336-364).  `samples` random and adversarial operand pairs
 * are compared bit for bit with div.rn.f64 / sqrt.rn.f64 on the device; mismatches[0..2] = quotient, reciprocal, root. */
int pluto_gpu_selftest_arith (int device, long long samples, unsigned long long seed, unsigned long long mismatches[3]);

#ifdef __cplusplus
}
#endif
#endif
