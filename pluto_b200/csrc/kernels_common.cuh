// kernels_common.cuh -- geometry of one block in HBM, kernel argument structs
// and the launch interface between the C-ABI host (pluto_gpu.cu) and the
// kernel translation units.
//
// HBM layout.  Every scalar field (8 primitives, 5 conservative
// accumulators, 3 staggered components, 6 face EMFs, 3 edge EMFs, C_dt) is
// a separate array with the SAME padded shape: logical index -1 .. T in
// every active dimension,
//       idx(k,j,i) = (k + off3)*S12 + (j + off2)*S1 + (i + off1),
// S1 = T1 + 2, S12 = S1*(T2 + 2); off = 1 in active dimensions (0 for x3 in
// 2-D).  i is fastest, so warps always run along x1 whatever the sweep
// direction (coalesced 256-byte rows).  Staggered component d at index i
// is the face i+1/2 in direction d (valid from -1), the reference's
// convention (Src/initialize.c:448-453).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct Geom {
  int dims, ng;
  int n[3], T[3], beg[3], end[3], off[3];
  long long S1, S12, tot;
  double dx[3];
};

__host__ __device__ __forceinline__ long long gidx (const Geom &g, int k, int j, int i)
{
  return (long long)(k + g.off[2])*g.S12 + (long long)(j + g.off[1])*g.S1 + (i + g.off[0]);
}

struct PhysPar { double gamma, gmm1, small_dn, small_pr, igmm1; };

// reduction slots (device, unsigned long long each)
enum { RED_CDT = 0, RED_MACH = 1, RED_FLOOR = 2, RED_NAN = 3, RED_ROEFAIL = 4, RED_N = 8 };

struct SweepArgs {
  const double *V[8];        // primitives of the current stage (ghosts filled)
  const double *Bn;          // staggered component normal to the sweep
  double       *U[8];        // conservative accumulators (B slots unused)
  double       *e1, *e2;     // face EMFs of this direction (see sweep_kernels.cuh)
  signed char  *sv;          // sign of the mass flux (UCT_CONTACT upwinding)
  double       *cdt;         // per-zone sum of directional inverse time steps
  unsigned long long *red;   // reduction slots
  Geom    g;
  PhysPar ph;
  const double *dtp;         // device: dt/dx1, dt/dx2, dt/dx3 (read at run time so that a
                             // captured CUDA graph of the step can be replayed with a new dt)
  double  inv_dl;
  int     stage1;            // accumulate C_dt / CFL (g_intStage == 1)
  int     last_dir;          // this sweep completes C_dt -> reduce instead of store
  int     u_from_v;          // x1 sweep: start from U = PrimToCons(V) (rk_step.c:93) instead of
                             // continuing with the U left by the previous stage
  int     chunk_len;         // marching kernels: zones per thread along the sweep
  int     nchunk;
  int     tma;               // fused x1+x2 sweep: stage the ring rows with bulk asynchronous copies (TMA engine, one elected
                             // lane + mbarrier) instead of per-lane cp.async; needs rows of an even number of doubles
  int     plan;              // 1: the launcher picks the chunk count that fills the SMs in whole rounds (plan_chunks),
                             //    chunk_len = the longest chunk the host wants; 0: use chunk_len / nchunk as given
  int     limiter;           // PLUTO_GPU_LIM_* (PLM)
  int     char_lim;          // CHAR_LIMITING YES (2-D, LINEAR): slopes limited on the characteristic variables
  // UCT_HLL only (avg == 3): e1/e2 (e3/e4) receive the fan speeds max(0,-SL), max(0,SR) of the
  // faces instead of the face EMFs, and the limited velocity slopes vp - vm of every
  // reconstructed zone are kept (CT_StoreVelSlopes, ct_stag_slopes.c:5-45)
  int     avg;               // PLUTO_GPU_EMF_*
  double *dvel[3];           // d v_c / d x_dir of this sweep's direction, c = 0..2
  double *dvel2[3];          // fused x1 + x2 sweep: the x2 slopes
  const unsigned char *flag; // SHOCK_FLATTENING MULTID: FLAG_MINMOD 1 | FLAG_HLL 4 per zone (pluto.h:192-194)
  // fused x1 + x2 sweep only: the x2 quantities next to the x1 ones above
  const double *Bn2;         // Bx2s
  double       *e3, *e4;     // ezj, exj
  signed char  *sv2;         // svy
  double  inv_dl2;
  // CT_EN_CORRECTION with EXACT arithmetic: the flux of the NORMAL field component of every face.  It is zero
  // analytically, but hlld.c:200-330 forms it as SL*(Bx* - Bn) with Bx* = (SR Bn - SL Bn)/(SR - SL), a round-off
  // residue that the reference adds to Uc[BXn] and that enters b2_old of ct_field_average.c:116-129
  double *fbn;
  // BODY_FORCE VECTOR (rhs_source.c:214-217, 277-280, 342-345): template flag BF of the sweeps.  Uniform acceleration
  // grav[], or a static per-zone field gf (this sweep's component; gf2: the x2 component of the fused x1+x2 sweep)
  double  grav[3];
  const double *gf, *gf2;
  int     bfv;               // vector part present (BODY_FORCE & VECTOR)
  // BODY_FORCE & POTENTIAL (rhs.c:388-392, rhs_source.c:233-237, 316-320, 358-362): potential at the zone centres and at
  // the faces of this sweep's direction (phif2: x2 faces, fused x1+x2 sweep); NULL without a potential
  const double *phic, *phif, *phif2;
  // x3 sweep of the bench configuration (FAST, 3-D, LINEAR, plain options): the flux difference of the sweep is STORED here
  // (slots RHO, MX1..3, ENG) instead of being added to U -- the sweep then stages no U (38 instead of 48 shared-memory slots per
  // thread: four blocks per SM) and does not depend on the fused x1+x2 sweep; the stage completion forms U + R3.  NULL: U += rhs.
  double *R3[8];
  // Non-uniform Cartesian grid (pluto_gpu_set_grid): gs = 1 and dtx / idl point at per-zone arrays along the sweep direction,
  // dt/dx[n] (rhs.c:195) and 1/dx[n] (inv_dl of update_stage.c:229-235); uniform grid: gs = 0, dtx = dtp + direction (the scalar),
  // idl unused (inv_dl above).  dtx2 / idl2: the x2 direction of the fused x1+x2 sweep.
  const double *dtx, *idl, *dtx2, *idl2;
  int     gs;
  // fused x1+x2 sweep: cell-centred EMFs of every zone it reads (CT_ComputeCenterEMF, ct_emf.c:348-388: Ex1 = vz By - vy Bz,
  // Ex2 = vx Bz - vz Bx, Ex3 = vy Bx - vx By of the stage's input state), stored for ct_emf_kernel, which then loads 12 values
  // per edge triple instead of the 36 primitives it would recompute them from.  NULL: not stored.
  double *Ec[3];
  // RECON_PLMW (UNIFORM_CARTESIAN_GRID NO): cp, cm, wp, wm, dp, dm of the sweep direction, one entry per zone along it
  // (PLM_CoefficientsGet, plm_coeffs.c:86-104); pc2: the x2 direction of the fused x1+x2 sweep
  const double *pc[6], *pc2[6];
  const double *qc, *qc2;            // PARABOLIC on a non-uniform grid: interface weights wp[n][-1 .. 2] along the sweep (qc2: along x2 in the fused sweep), else NULL
};

struct CtArgs {
  const double *V[8];                        // current-stage primitives (cell-centre EMF)
  const double *exj, *exk, *eyi, *eyk, *ezi, *ezj;
  const signed char *svx, *svy, *svz;
  double *ex, *ey, *ez;                      // edge EMFs
  const double *Bs_in[3];                    // staggered field of the current stage
  const double *Bs0[3];                      // staggered field at t^n (stages >= 2)
  double *Bs_out[3];
  Geom   g;
  const double *dtp;                         // device: dt/dx1, dt/dx2, dt/dx3
  double w0, wc;                             // stage weights
  int    combine;                            // 0: none, 1: w0*B0 + wc*B, 2: (B0 + 2 B)/3
  int    avg;                                // PLUTO_GPU_EMF_*
  const double *dvel[3][3];                  // UCT_HLL: dvel[c][d] = d v_c / d x_d
  int    ext;                                // edges / faces of [beg-1-ext, end+ext]: 0 (RK, CTU corrector) or
                                             // 1 (CTU predictor, emf ranges of ctu_step.c:290-297)
  const double *Ec[3];                       // cell-centred EMFs stored by the fused x1+x2 sweep (SweepArgs.Ec), else NULL
  const double *dtx[3];                      // dt/dx of direction d: per zone (gs = 1, non-uniform grid: ct_update.c:91-96 takes
  int    gs;                                 // dt/dx2[j], dt/dx3[k], ...) or the scalar dtp + d (gs = 0)
  double dts;                                // factor on dt/dx: 1, or 1/2 for the half step of the CTU predictor on a non-uniform
                                             // grid (halving is exact: (dt/2)/dx = (dt/dx)/2 bit for bit)
};

// corner-transport-upwind step (TIME_STEPPING HANCOCK, Src/Time_Stepping/ctu_step.c:142-727)
struct CtuArgs {
  const double *V0[8];       // primitives at t^n, ghost zones filled
  const double *Bs0[3];      // staggered field at t^n
  const double *Bsh[3];      // staggered field at t^n + dt/2 (corrector, half-step kernel)
  double       *rhs[3][8];   // half-step right-hand sides of the three normal predictors (ctu_step.c:221)
  double       *U[8];        // conservative accumulators (corrector)
  double       *Vh[8];       // half-step kernel: primitives at t^n + dt/2
  double       *e1, *e2;     // face EMFs of this direction
  signed char  *sv;
  unsigned long long *red;
  const unsigned char *flag;
  Geom    g;
  PhysPar ph;
  const double *dtp;         // device: dt/dx1..3, dt, (dt/2)/dx1..3, dt/2
  double  inv_dl;            // 1/dx of the sweep direction
  int     limiter;
  int     chunk_len, nchunk; // marching sweeps (x2, x3)
  int     en_corr;           // CT_EN_CORRECTION YES: half-step kernel (ct_field_average.c:116-129 on Uh)
  int     bf;                // BODY_FORCE VECTOR: uniform grav[] or per-zone gf of this direction (rhs_source.c:214-345,
  double  grav[3];           // prim_eqn.c:289-360)
  const double *gf;
  const double *phic, *phif; // BODY_FORCE & POTENTIAL: potential at the centres / the faces of this direction (else NULL)
  double *fbn;               // corrector, EXACT + CT_EN_CORRECTION: normal-component flux of the faces (see SweepArgs)
  const double *dtx, *idl;   // dt/dx[n] and 1/dx[n] along the sweep on a non-uniform grid (gs = 1), else dtp + direction (gs = 0)
  const double *dxz;         // ... and dx[n] itself (the potential's source of the predictor, prim_eqn.c:304-307), NULL on a uniform grid
  int     gs;
  int     chtr;              // TIME_STEPPING CHARACTERISTIC_TRACING: predictor by characteristic tracing (2 components)
  int     char_lim;          // CHAR_LIMITING YES (2 components): slopes limited on the characteristic variables (plm_zone_char2)
};

struct FinalArgs {
  const double *U[8];
  const double *Bs[3];
  const double *V0[8];                       // primitives at t^n (stages >= 2)
  double *Vout[8];
  unsigned long long *red;
  Geom    g;
  PhysPar ph;
  double  w0, wc;
  int     combine;                           // 0 none, 1 w0*U0 + wc*U, 2 (U0 + 2U)/3
  int     write_u;                           // 0: U is dead after this stage; 1: store the stage
                                             // result (a later stage continues from it);
                                             // 2: store only ConsToPrim repairs
  double *Uw[8];
  int     box_lo[3], box_n[3];               // zones of this launch, relative to the first interior zone
  // several boxes in ONE launch (the up to six shell slabs of a decomposed block): blockIdx.y selects the box
  int     nbox;                              // 0: the single box above
  int     boxes_lo[6][3], boxes_n[6][3];
  // CT_EN_CORRECTION YES (ct_field_average.c:116-129): the cell-centred conservative field the sweeps WOULD have
  // produced is rebuilt from the face EMFs (= the induction fluxes the sweeps stored) and the stage's input field
  int     en_corr;
  const double *Vin[8];                      // primitives the stage started from (B slots used)
  const double *exj, *exk, *eyi, *eyk, *ezi, *ezj;
  const double *fbn[3];                      // normal-component flux of the x1, x2, x3 faces (EXACT; NULL: zero)
  const double *dtp;                         // device: dt/dx1..3
  const double *dtx[3]; int gs;              // energy correction: dt/dx of the zone along each direction (gs = 1: per-zone arrays, 0: one value)
  // fuse_ct: CT_Update (+ the RK average of the staggered field) evaluated HERE instead of in ct_update_kernel: a zone
  // computes the new field of its six faces from the edge EMFs (the three low ones a second time, bit-identical to the
  // neighbour's), stores the three high ones (and a low one on the face beg-1) into Bs (= Bs_out) and averages them
  int     fuse_ct;
  const double *ex, *ey, *ez;                // edge EMFs
  const double *Bs_in[3], *Bs0[3];           // staggered field of the stage's input and at t^n
  double *Bs_out[3];
  const double *R3[8];                       // flux difference of the x3 sweep kept apart (SweepArgs.R3): u = U + R3; NULL: u = U
};

// Boundary conditions of ONE dimension in one launch: the copy jobs (a field and
// its ghost box on one side) and the div B = 0 fill jobs of both sides.
struct BcField {
  double *q;
  int lo[3], hi[3];          // inclusive destination box
  int sign;                  // reflective: +1 / -1
  int side, type;            // side 0..5, type PLUTO_GPU_BC_*
};
struct BcFill {              // FillMagneticField + CT_AverageNormalMagField
  double *Bc;                // cell-centred normal component (NULL: no averaging)
  int side, type;
};
struct BcArgs {
  BcField f[22];
  BcFill  fill[2];
  double *Bs[3];
  int nf, nfill;
  Geom g;
  const double *dxa[3];      // zone widths of a non-uniform grid (face areas of FillMagneticField), else NULL: g.dx
};

struct FlagArgs {              // FlagShock (flag_shock.c:79-230)
  const double *vx[3], *prs;
  unsigned char *shock;      // pass 1: zone lies in a shock
  unsigned char *flag;       // pass 2: FLAG_HLL | FLAG_MINMOD of the zone itself, FLAG_MINMOD of its neighbours
  const double *dxa[3];      // non-uniform grid: zone widths per direction (flag_shock.c:143-145), else NULL
  Geom g;
};

struct AnalysisArgs {          // volume integrals of the interior state
  const double *V[8];
  const double *Bs[3];
  double *partial;           // [gridDim.x][8] block partial sums (8th: max |div B|)
  Geom g;
  double igmm1;
};

struct CvtArgs {               // interior zones of the 8 primitives as single-precision values, variable after variable
  const double *V[8];
  float *out;                // [live variable][n3][n2][n1]
  Geom g;
  int live[8];               // slot of each variable in `out` (-1: not written)
  int swap;                  // byte-swapped (the big-endian floats of a .vtk file)
};

struct HaloArgs {
  double *q[11];
  int lo[11][3], hi[11][3];  // inclusive box per field
  long long offset[11];      // start of each field inside the buffer
  int nf;
  double *buf;
  Geom g;
};

// all-neighbour halo exchange: one table entry per (neighbour, field), resident in device memory
struct HaloEntry {
  double *field;             // padded field array
  double *buf;               // start of this field's segment inside the neighbour's buffer
  int lo[3], n[3];           // box origin and extents
  long long count;
};

// launchers return the number of kernels launched, or -1 - cudaError of a failed launch (count() in pluto_gpu.cu reports it)
static inline int pg_launch_status (int n = 1)
{
  const cudaError_t e = cudaGetLastError ();
  return e == cudaSuccess ? n : -1 - (int)e;
}

// Function attributes (dynamic shared-memory limit, carve-out) belong to the device the call was made on: a process that drives
// several GPUs (pluto_gpu_multi_*) has to set them once PER DEVICE.  `mask` is the launch site's static record of the devices done.
static inline bool pg_attr_needed (unsigned long long &mask)
{
  int dev = 0;
  cudaGetDevice (&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

// ---- launch interface, one set per arithmetic namespace ----------------------
#define PG_DECLARE_LAUNCHERS(NS)                                                         \
namespace NS {                                                                           \
  int launch_sweep_hlld (int dir, int recon, const SweepArgs &a, cudaStream_t s, bool bf);  \
  int launch_sweep_hll  (int dir, int recon, const SweepArgs &a, cudaStream_t s, bool bf);  \
  int launch_sweep_roe  (int dir, int recon, const SweepArgs &a, cudaStream_t s, bool bf);  \
  int launch_sweep_hllc (int dir, int recon, const SweepArgs &a, cudaStream_t s, bool bf);  \
  int launch_sweep_tvdlf (int dir, int recon, const SweepArgs &a, cudaStream_t s, bool bf); \
  int launch_sweep_xy_hlld (int recon, const SweepArgs &a, cudaStream_t s, bool bf);        \
  int launch_sweep_xy_hll  (int recon, const SweepArgs &a, cudaStream_t s, bool bf);        \
  int launch_sweep_xy_roe  (int recon, const SweepArgs &a, cudaStream_t s, bool bf);        \
  int launch_sweep_xy_hllc (int recon, const SweepArgs &a, cudaStream_t s, bool bf);        \
  int launch_sweep_xy_tvdlf (int recon, const SweepArgs &a, cudaStream_t s, bool bf);       \
  int launch_ctu_sweep_hlld (int dir, int phase, const CtuArgs &a, cudaStream_t s);      \
  int launch_ctu_sweep_hll  (int dir, int phase, const CtuArgs &a, cudaStream_t s);      \
  int launch_ctu_sweep_roe  (int dir, int phase, const CtuArgs &a, cudaStream_t s);      \
  int launch_ctu_sweep_hllc (int dir, int phase, const CtuArgs &a, cudaStream_t s);      \
  int launch_ctu_sweep_tvdlf (int dir, int phase, const CtuArgs &a, cudaStream_t s);     \
  int launch_ctu_half   (const CtuArgs &a, cudaStream_t s);                              \
  int launch_ct_emf     (const CtArgs &a, cudaStream_t s);                               \
  int launch_ct_update  (const CtArgs &a, cudaStream_t s);                               \
  int launch_final      (const FinalArgs &a, cudaStream_t s);                            \
  int launch_bc         (const BcArgs &a, cudaStream_t s);                               \
  int launch_flag_shock (const FlagArgs &a, cudaStream_t s);                             \
  int launch_analysis   (const AnalysisArgs &a, int nblocks, cudaStream_t s);            \
  int launch_cvt_float  (const CvtArgs &a, cudaStream_t s);                              \
  int launch_halo_pack  (const HaloArgs &a, cudaStream_t s);                             \
  int launch_halo_unpack(const HaloArgs &a, cudaStream_t s);                             \
  int launch_halo_table (const HaloEntry *tab, int n, long long maxcount, const Geom &g, bool pack, cudaStream_t s); \
}
PG_DECLARE_LAUNCHERS(pg_exact)
PG_DECLARE_LAUNCHERS(pg_fast)
