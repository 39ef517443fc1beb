// pluto_gpu.cu -- host side of the C ABI declared in include/pluto_gpu.h:
// device-resident state, the RK stage pipeline (the body of the reference's
// AdvanceStep, Src/Time_Stepping/rk_step.c:27-254), boundary orchestration
// (Src/boundary.c:41-315) and host<->device transfer in the reference's own
// array layouts.  No torch types, no CPU fallback.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <unistd.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pluto_gpu.h"
#include "kernels_common.cuh"

static thread_local char g_err[512] = "";

static int fail (const char *fmt, ...)
{
  va_list ap;
  va_start (ap, fmt);
  vsnprintf (g_err, sizeof (g_err), fmt, ap);
  va_end (ap);
  return 1;
}

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
  return fail ("%s:%d: %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString (e_)); } while (0)

enum { NVS = 8 };
enum { PG_MAX_EV = 96, PG_NCLASS = 8 };
enum { HIST_W = 6, HIST_N = 4096 };     // per step: dt used, inv_dt_hyp, max Mach, floors, NaNs, Roe failures
enum { KC_SWEEP_X = 0, KC_SWEEP_Y, KC_SWEEP_Z, KC_CT_EMF, KC_CT_UPDATE, KC_FINAL, KC_BC, KC_HALO };
static const char *kClassName[PG_NCLASS] = {"sweep_x1", "sweep_x2", "sweep_x3", "ct_emf", "ct_update",
                                            "final", "boundary", "halo"};
static const int kCons[5] = {0, 1, 2, 3, 7};       // RHO, MX1, MX2, MX3, ENG

struct PlutoGpu {
  PlutoGpuConfig cfg;
  Geom    g;
  PhysPar ph;
  int     nbuf;                    // state buffers: 2 (RK2) or 3 (RK3)
  cudaStream_t stream;
  void   *pool;
  size_t  pool_bytes;
  double *V[3][NVS];
  double *Bs[3][3];
  double *U[NVS];
  double *exj, *exk, *eyi, *eyk, *ezi, *ezj;
  double *ex, *ey, *ez;
  double *cdt;
  unsigned char *flag, *shock;     // SHOCK_FLATTENING MULTID only: zone flags and the pass-1 shock marks
  double *dvel[3][3];              // UCT_HLL only: limited velocity slopes d v_c / d x_d (own allocation)
  void   *dvel_pool;
  double *scratch;                 // = first array after the state buffers
  size_t  scratch_doubles;
  signed char *sv[3];
  unsigned long long *red;         // device reduction slots
  unsigned long long *red_host;    // pinned mirror
  // all-neighbour halo plan: tables [state buffer][0 pack | 1 unpack] in device memory
  HaloEntry *halo_tab[3][2];
  int     halo_n[3][2];
  long long halo_max[3][2];
  const void *pinned[8];           // host blocks seen by pluto_gpu_advance_data
  int     pinned_by_us[8];         // ... and page-locked here
  int     npinned;
  // device-side NextTimeStep: results of the steps enqueued since the last synchronisation
  double *hist;                    // device ring: HIST_W doubles per step
  double *hist_host;               // pinned mirror
  unsigned long long *hist_count;  // device: steps recorded so far
  long long hist_read;             // steps already handed to the caller
  long long hist_enq;              // steps enqueued so far
  double *dtdev;                   // device: dt/dx[0..2] of the current step
  double *dthost;                  // pinned staging of the same
  cudaGraphExec_t graph;           // captured single-GPU step (all stages), replayed with new dt
  long long graph_launches;        // kernels inside the graph
  int     use_graph;
  int     fuse_xy;                 // FAST: fused x1+x2 sweep kernel (PLUTO_GPU_NO_FUSE_XY=1 disables)
  long long steps_done;
  long long launches;
  int     march_chunk;             // zones per thread along a marching sweep
  int     plan;                    // the sweep launchers choose the chunk count (PLUTO_GPU_NO_PLAN=1 disables)
  int     fuse_ct;                 // RK stages: CT_Update inside final_kernel (PLUTO_GPU_FUSE_CT=1)
  int     shell_w;                 // width of the x1 slabs of the shell (stage completion next to shared sides)
  int     tma;                     // fused sweep: ring rows staged by bulk asynchronous copies (PLUTO_GPU_TMA=1)
  int     ctu;                     // TIME_STEPPING HANCOCK (corner transport upwind)
  int     nstages;                 // Boundary calls per step: rk_order, or 1 with CTU
  double *gfield[3];               // static per-zone body force (pluto_gpu_set_body_force), else NULL
  void   *gfield_pool;
  double *phic, *phif[3];          // body-force potential at centres and faces (pluto_gpu_set_body_potential), else NULL
  void   *phi_pool;
  double *fbn[3];                  // CT_EN_CORRECTION + EXACT: normal-field flux of the faces (own allocation)
  void   *fbn_pool;
  double *R3[NVS];                 // FAST, 3-D, LINEAR, plain options: flux difference of the x3 sweep (own allocation), else NULL
  void   *r3_pool;
  double *plmc[3][6];              // UNIFORM_CARTESIAN_GRID NO: reconstruction weights cp, cm, wp, wm, dp, dm per direction
  double *ppmc[3]; void *ppmc_pool; // PARABOLIC on a non-uniform grid: interface weights wp[n][-1 .. 2] per direction (pluto_gpu_set_ppm_coeffs)
  void   *plmc_pool;               // (pluto_gpu_set_plm_coeffs); plmw: all directions set -> the sweeps run their RECON_PLMW variants
  int     plmw;
  double *Ec[3];                   // FAST + fused x1+x2 sweep + UCT_CONTACT: cell-centred EMFs stored by the sweep (own allocation)
  void   *ec_pool;
  // non-uniform Cartesian grid (pluto_gpu_set_grid): per direction the zone widths dx[n], 1/dx[n] and dt/dx[n] (refreshed with
  // every new dt), n = 0 .. T-1 as in the reference's grid->dx[d]; nu = 0: uniform grid, the scalars of dtdev
  int     nu;
  double *dxa[3], *idxa[3], *dtxa[3];
  void   *grid_pool;
  double *rhs3[3][NVS];            // CTU: half-step right-hand sides of the normal predictors (own allocation)
  void   *ctu_pool;
  // optional per-kernel-class device timing (CUDA events on `stream`)
  int     timing;
  int     nev;                                 // event pairs used in the current step
  cudaEvent_t ev0[PG_MAX_EV], ev1[PG_MAX_EV];
  int     ev_class[PG_MAX_EV];
  double  class_ms[PG_NCLASS];
  long long class_count[PG_NCLASS];
};

const char *pluto_gpu_last_error (void) { return g_err; }

int pluto_gpu_nghost (const PlutoGpu *h) { return h->g.ng; }
int pluto_gpu_nstages (const PlutoGpu *h) { return h->nstages; }

static bool live_var (const PlutoGpu *h, int nv) { return h->g.dims == 3 || (nv != 3 && nv != 6); }

#define DISPATCH(h, call) ((h)->cfg.arith == PLUTO_GPU_ARITH_FAST ? pg_fast::call : pg_exact::call)

static int count (PlutoGpu *h, int r)
{
  if (r < 0) return fail ("kernel launch failed: %s", cudaGetErrorString ((cudaError_t)(-1 - r)));
  h->launches += r;
  return 0;
}

// bracket a launch with events when timing is on
static int tbegin (PlutoGpu *h, int cls)
{
  if (!h->timing || h->nev >= PG_MAX_EV) return -1;
  const int e = h->nev++;
  if (!h->ev0[e]){ cudaEventCreate (&h->ev0[e]); cudaEventCreate (&h->ev1[e]); }
  h->ev_class[e] = cls;
  cudaEventRecord (h->ev0[e], h->stream);
  return e;
}
static void tend (PlutoGpu *h, int e) { if (e >= 0) cudaEventRecord (h->ev1[e], h->stream); }
static void tcollect (PlutoGpu *h)       // after a stream synchronise
{
  for (int e = 0; e < h->nev; e++){
    float ms = 0.f;
    if (cudaEventElapsedTime (&ms, h->ev0[e], h->ev1[e]) == cudaSuccess){
      h->class_ms[h->ev_class[e]] += ms; h->class_count[h->ev_class[e]]++;
    }
  }
  h->nev = 0;
}
#define TIMED(h, cls, expr) do { int te_ = tbegin (h, cls); int rc_ = (expr); tend (h, te_); if (rc_) return 1; } while (0)

// ---------------------------------------------------------------------------
static int create_resources (PlutoGpu *h);

int pluto_gpu_create (const PlutoGpuConfig *cfg, PlutoGpu **out)
{
  *out = NULL;
  if (cfg->dims != 2 && cfg->dims != 3) return fail ("dims must be 2 or 3");
  if (cfg->recon != PLUTO_GPU_RECON_LINEAR && cfg->recon != PLUTO_GPU_RECON_PARABOLIC) return fail ("bad recon");
  if (cfg->solver < 0 || cfg->solver > 4) return fail ("bad solver");
  if (cfg->rk_order != 2 && cfg->rk_order != 3) return fail ("rk_order must be 2 or 3");
  if (cfg->limiter < 0 || cfg->limiter > PLUTO_GPU_LIM_MC) return fail ("bad limiter");
  if (cfg->shock_flattening != 0 && cfg->shock_flattening != 1) return fail ("bad shock_flattening");
  if (cfg->emf_average < 0 || cfg->emf_average > PLUTO_GPU_EMF_UCT_HLL) return fail ("bad emf_average");
  if (cfg->time_stepping < PLUTO_GPU_TS_RK || cfg->time_stepping > PLUTO_GPU_TS_CHAR_TRACING) return fail ("bad time_stepping");
  if (cfg->time_stepping == PLUTO_GPU_TS_CHAR_TRACING){
    if (cfg->dims != 2)
      return fail ("TIME_STEPPING CHARACTERISTIC_TRACING is available in 2-D only (in 3-D the reference's eigenvector scratch, eigenv.c:190-560, "
                   "keeps entries of the previous sweep direction: its result depends on the sweep order and cannot be reproduced)");
    if (cfg->shock_flattening || cfg->body_force || cfg->en_correction)
      return fail ("TIME_STEPPING CHARACTERISTIC_TRACING is available without SHOCK_FLATTENING, BODY_FORCE and CT_EN_CORRECTION");
    // The traced states contain alpha_s = sqrt((cf^2 - a^2)/...) at FIRST order (the fast and the slow family are weighted by their
    // own Courant numbers), and where the transverse field vanishes exactly that difference is pure round-off: the reference's own
    // result moves by 1e-9 with the last bit of the state.  Only its operation order reproduces it, so the re-associated FAST
    // arithmetic is not offered here (measured: dt off by more than 1e-12 within a few steps on blast2d_chtr_mc_roe).
    if (cfg->arith == PLUTO_GPU_ARITH_FAST)
      return fail ("TIME_STEPPING CHARACTERISTIC_TRACING is available with EXACT arithmetic only (its predictor is ill-conditioned where the "
                   "transverse field vanishes: FAST arithmetic cannot stay within 1e-12 of the reference there)");
  }
  if (cfg->en_correction != 0 && cfg->en_correction != 1) return fail ("bad en_correction");
  if (cfg->en_correction && cfg->emf_average == PLUTO_GPU_EMF_UCT_HLL)
    return fail ("CT_EN_CORRECTION YES is available with CT_EMF_AVERAGE UCT_CONTACT / ARITHMETIC / UCT0 "
                 "(the correction is rebuilt from the face EMFs, which UCT_HLL replaces by the fan speeds)");
  if (cfg->body_force < 0 || cfg->body_force > 3) return fail ("bad body_force");
  if (cfg->char_limiting != 0 && cfg->char_limiting != 1) return fail ("bad char_limiting");
  if (cfg->char_limiting){
    if (cfg->dims != 2)
      return fail ("CHAR_LIMITING YES is available in 2-D only (in 3-D the reference's eigenvector scratch, eigenv.c:190-560, keeps "
                   "entries of the previous sweep direction: its result depends on the sweep order and cannot be reproduced)");
    if (cfg->recon != PLUTO_GPU_RECON_LINEAR || cfg->shock_flattening || (cfg->time_stepping != PLUTO_GPU_TS_RK && cfg->en_correction))
      return fail ("CHAR_LIMITING YES is available with LINEAR reconstruction (RK2 / RK3 / HANCOCK / CHARACTERISTIC_TRACING), without "
                   "SHOCK_FLATTENING");
  }
  if (cfg->time_stepping != PLUTO_GPU_TS_RK){
    if (cfg->recon != PLUTO_GPU_RECON_LINEAR) return fail ("TIME_STEPPING HANCOCK needs LINEAR reconstruction (Src/pluto.h: RK only with PARABOLIC)");
    if (cfg->emf_average == PLUTO_GPU_EMF_UCT_HLL)      // the reference refuses the same combination (MHD/CT/ct_emf.c:196-200)
      return fail ("UCT_HLL average not compatible with CTU schemes (stencil too small): use UCT_CONTACT, ARITHMETIC or UCT0");
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount (&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail ("no CUDA device (%s): this library has no CPU path", cudaGetErrorString (e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail ("device %d out of range (%d devices)", cfg->device, ndev);
  CU (cudaSetDevice (cfg->device));

  PlutoGpu *h = (PlutoGpu *)calloc (1, sizeof (PlutoGpu));
  if (!h) return fail ("out of host memory");
  h->cfg = *cfg;
  if (create_resources (h)){                   // nothing is leaked on a failed allocation
    char msg[sizeof (g_err)];
    snprintf (msg, sizeof (msg), "%s", g_err);
    pluto_gpu_destroy (h);
    snprintf (g_err, sizeof (g_err), "%s", msg);
    return 1;
  }
  *out = h;
  return 0;
}

static int create_resources (PlutoGpu *h)
{
  const PlutoGpuConfig *cfg = &h->cfg;
  Geom &g = h->g;
  g.dims = cfg->dims;
  g.ng = (cfg->recon == PLUTO_GPU_RECON_PARABOLIC ? 3 : 2);      // get_nghost.c:32-50
  if (cfg->shock_flattening && g.ng < 3) g.ng = 3;               // get_nghost.c:67-77
  h->ctu = (cfg->time_stepping != PLUTO_GPU_TS_RK);
  if (h->ctu) g.ng++;                                            // CTU + CT, get_nghost.c:86-90
  h->nstages = h->ctu ? 1 : cfg->rk_order;
  for (int d = 0; d < 3; d++){
    if (d < g.dims){
      if (cfg->n[d] < 2*g.ng) return fail ("n[%d] = %d is smaller than 2*nghost", d, cfg->n[d]);
      g.n[d] = cfg->n[d]; g.T[d] = g.n[d] + 2*g.ng; g.beg[d] = g.ng; g.end[d] = g.ng + g.n[d] - 1; g.off[d] = 1;
      g.dx[d] = cfg->dx[d];
    }else{
      g.n[d] = 1; g.T[d] = 1; g.beg[d] = g.end[d] = 0; g.off[d] = 0; g.dx[d] = 1.0;
    }
  }
  g.S1  = g.T[0] + 2;
  g.S12 = g.S1*(g.T[1] + 2);
  g.tot = g.S12*(g.dims == 3 ? g.T[2] + 2 : 1);
  if (g.tot >= (1LL << 31)) return fail ("block of %lld padded zones: the kernels index with 32 bits (< 2^31 zones per block)", g.tot);
  h->ph.gamma = cfg->gamma; h->ph.gmm1 = cfg->gamma - 1.0;
  h->ph.small_dn = cfg->small_dn; h->ph.small_pr = cfg->small_pr;
  h->ph.igmm1 = 1.0/(cfg->gamma - 1.0);
  h->nbuf = (cfg->rk_order == 3 && !h->ctu ? 3 : 2);
  h->march_chunk = 64;

  // one pool: [state buffers][U, face EMFs, edge EMFs, C_dt = scratch][sign bytes]
  const size_t tot = (size_t)g.tot;
  const size_t tot_al = (tot + 31) & ~(size_t)31;             // 256-byte aligned arrays
  const int nstate = h->nbuf*(NVS + 3);
  const int nwork = 5 + 6 + 3 + 1;
  h->pool_bytes = (size_t)(nstate + nwork)*tot_al*sizeof (double) + 3*tot_al;
  if (cudaMalloc (&h->pool, h->pool_bytes) != cudaSuccess){ h->pool = NULL; return fail ("cudaMalloc of %zu bytes failed", h->pool_bytes); }
  CU (cudaMemset (h->pool, 0, h->pool_bytes));
  double *p = (double *)h->pool;
  for (int b = 0; b < h->nbuf; b++){
    for (int nv = 0; nv < NVS; nv++){ h->V[b][nv] = p; p += tot_al; }
    for (int d = 0; d < 3; d++){ h->Bs[b][d] = p; p += tot_al; }
  }
  h->scratch = p;
  h->scratch_doubles = (size_t)nwork*tot_al;
  for (int q = 0; q < 5; q++){ h->U[kCons[q]] = p; p += tot_al; }
  h->exj = p; p += tot_al; h->exk = p; p += tot_al; h->eyi = p; p += tot_al;
  h->eyk = p; p += tot_al; h->ezi = p; p += tot_al; h->ezj = p; p += tot_al;
  h->ex = p; p += tot_al; h->ey = p; p += tot_al; h->ez = p; p += tot_al;
  h->cdt = p; p += tot_al;
  signed char *c = (signed char *)p;
  for (int d = 0; d < 3; d++){ h->sv[d] = c; c += tot_al; }

  if (cfg->shock_flattening){
    CU (cudaMalloc ((void **)&h->flag, 2*(tot_al + 256)));
    CU (cudaMemset (h->flag, 0, 2*(tot_al + 256)));
    h->shock = h->flag + tot_al + 256;
    h->pool_bytes += 2*(tot_al + 256);
  }
  if (cfg->emf_average == PLUTO_GPU_EMF_UCT_HLL){
    const size_t nb = (size_t)g.dims*g.dims*tot_al*sizeof (double);
    if (cudaMalloc (&h->dvel_pool, nb) != cudaSuccess) return fail ("cudaMalloc of %zu bytes (UCT_HLL slopes) failed", nb);
    CU (cudaMemset (h->dvel_pool, 0, nb));
    double *q = (double *)h->dvel_pool;
    for (int c = 0; c < g.dims; c++) for (int d = 0; d < g.dims; d++){ h->dvel[c][d] = q; q += tot_al; }
    h->pool_bytes += nb;
  }
  if (cfg->en_correction && cfg->arith == PLUTO_GPU_ARITH_EXACT){
    const size_t nb = (size_t)g.dims*tot_al*sizeof (double);
    if (cudaMalloc (&h->fbn_pool, nb) != cudaSuccess) return fail ("cudaMalloc of %zu bytes (normal-field fluxes) failed", nb);
    CU (cudaMemset (h->fbn_pool, 0, nb));
    for (int d = 0; d < g.dims; d++) h->fbn[d] = (double *)h->fbn_pool + (size_t)d*tot_al;
    h->pool_bytes += nb;
  }
  // x3 sweep with its flux difference kept apart (SweepArgs.R3; PLUTO_GPU_R3=1): a measured opt-in.  The x3 sweep then stages no U,
  // fits four blocks per SM and is 7.5 % faster, but the stage completion reads 40 B per zone more and loses twice that
  // (profiles/r2r_*: 6.18 against 6.12 ms per step at 256^3)
  if (getenv ("PLUTO_GPU_R3") && atoi (getenv ("PLUTO_GPU_R3")) != 0 && cfg->arith == PLUTO_GPU_ARITH_FAST && g.dims == 3 && cfg->recon == PLUTO_GPU_RECON_LINEAR && !h->ctu && !cfg->shock_flattening
      && cfg->emf_average != PLUTO_GPU_EMF_UCT_HLL && !cfg->body_force && !cfg->en_correction && !cfg->char_limiting
      && getenv ("PLUTO_GPU_NO_FUSE_XY") == NULL){
    const size_t nb = (size_t)5*tot_al*sizeof (double);
    if (cudaMalloc (&h->r3_pool, nb) != cudaSuccess) return fail ("cudaMalloc of %zu bytes (x3 flux differences) failed", nb);
    CU (cudaMemset (h->r3_pool, 0, nb));
    for (int q = 0; q < 5; q++) h->R3[kCons[q]] = (double *)h->r3_pool + (size_t)q*tot_al;
    h->pool_bytes += nb;
  }
  // cell-centred EMFs written by the fused x1+x2 sweep for ct_emf_kernel (SweepArgs.Ec): 12 loads per edge triple instead of 36
  if (cfg->arith == PLUTO_GPU_ARITH_FAST && !h->ctu && cfg->emf_average == PLUTO_GPU_EMF_UCT_CONTACT
      && getenv ("PLUTO_GPU_NO_FUSE_XY") == NULL && getenv ("PLUTO_GPU_NO_EC") == NULL){
    const int ne = (g.dims == 3 ? 3 : 1);
    const size_t nb = (size_t)ne*tot_al*sizeof (double);
    if (cudaMalloc (&h->ec_pool, nb) != cudaSuccess) return fail ("cudaMalloc of %zu bytes (cell-centred EMFs) failed", nb);
    CU (cudaMemset (h->ec_pool, 0, nb));
    if (g.dims == 3) for (int q = 0; q < 3; q++) h->Ec[q] = (double *)h->ec_pool + (size_t)q*tot_al;
    else h->Ec[2] = (double *)h->ec_pool;
    h->pool_bytes += nb;
  }
  if (h->ctu){
    int nlive = 0;
    for (int nv = 0; nv < NVS; nv++) nlive += live_var (h, nv);
    const size_t nb = (size_t)g.dims*nlive*tot_al*sizeof (double);
    if (cudaMalloc (&h->ctu_pool, nb) != cudaSuccess) return fail ("cudaMalloc of %zu bytes (CTU right-hand sides) failed", nb);
    CU (cudaMemset (h->ctu_pool, 0, nb));
    double *q = (double *)h->ctu_pool;
    for (int d = 0; d < g.dims; d++) for (int nv = 0; nv < NVS; nv++) if (live_var (h, nv)){ h->rhs3[d][nv] = q; q += tot_al; }
    h->pool_bytes += nb;
  }
  CU (cudaStreamCreateWithFlags (&h->stream, cudaStreamNonBlocking));
  CU (cudaMalloc ((void **)&h->red, RED_N*sizeof (unsigned long long)));
  CU (cudaMemset (h->red, 0, RED_N*sizeof (unsigned long long)));
  CU (cudaMallocHost ((void **)&h->red_host, RED_N*sizeof (unsigned long long)));
  CU (cudaMalloc ((void **)&h->hist, (size_t)HIST_W*HIST_N*sizeof (double)));
  CU (cudaMallocHost ((void **)&h->hist_host, (size_t)HIST_W*HIST_N*sizeof (double)));
  CU (cudaMalloc ((void **)&h->hist_count, sizeof (unsigned long long)));
  CU (cudaMemset (h->hist_count, 0, sizeof (unsigned long long)));
  CU (cudaMalloc ((void **)&h->dtdev, 8*sizeof (double)));
  CU (cudaMallocHost ((void **)&h->dthost, 8*sizeof (double)));
  h->use_graph = (getenv ("PLUTO_GPU_NO_GRAPH") == NULL);
  h->fuse_xy = (getenv ("PLUTO_GPU_NO_FUSE_XY") == NULL);
  h->plan = (getenv ("PLUTO_GPU_NO_PLAN") == NULL);
  h->fuse_ct = (getenv ("PLUTO_GPU_FUSE_CT") != NULL && atoi (getenv ("PLUTO_GPU_FUSE_CT")) != 0);
  // x1 shell slabs: ng zones are needed; 4 zones = one 32-byte sector per row and array.  (Round 1 used a full warp, 32 zones:
  // coalesced, but 8 x the zones of the slab at 256-byte pieces 4 KB apart -- at 512^3 on 8 GPUs the split stage completion
  // took 8.2 ms per step against 6.7 ms unsplit.)
  h->shell_w = g.ng > 4 ? g.ng : 4;
  if (getenv ("PLUTO_GPU_SHELL_W") && atoi (getenv ("PLUTO_GPU_SHELL_W")) >= g.ng) h->shell_w = atoi (getenv ("PLUTO_GPU_SHELL_W"));
  if (h->shell_w*2 >= g.n[0]) h->shell_w = g.ng;
  // bulk copies need 16-byte aligned rows: an even number of doubles per row (arrays are 256-byte aligned)
  h->tma = (getenv ("PLUTO_GPU_TMA") && atoi (getenv ("PLUTO_GPU_TMA")) != 0 && g.S1 % 2 == 0);
  return 0;
}

void pluto_gpu_destroy (PlutoGpu *h)
{
  if (!h) return;
  cudaSetDevice (h->cfg.device);
  if (h->stream) cudaStreamSynchronize (h->stream);
  if (h->pool) cudaFree (h->pool);
  if (h->dvel_pool) cudaFree (h->dvel_pool);
  if (h->ctu_pool) cudaFree (h->ctu_pool);
  if (h->fbn_pool) cudaFree (h->fbn_pool);
  if (h->r3_pool) cudaFree (h->r3_pool);
  if (h->grid_pool) cudaFree (h->grid_pool);
  if (h->ec_pool) cudaFree (h->ec_pool);
  if (h->plmc_pool) cudaFree (h->plmc_pool);
  if (h->ppmc_pool) cudaFree (h->ppmc_pool);
  if (h->gfield_pool) cudaFree (h->gfield_pool);
  if (h->phi_pool) cudaFree (h->phi_pool);
  if (h->flag) cudaFree (h->flag);
  if (h->red) cudaFree (h->red);
  if (h->red_host) cudaFreeHost (h->red_host);
  if (h->dtdev) cudaFree (h->dtdev);
  if (h->dthost) cudaFreeHost (h->dthost);
  if (h->hist) cudaFree (h->hist);
  if (h->hist_host) cudaFreeHost (h->hist_host);
  if (h->hist_count) cudaFree (h->hist_count);
  for (int b = 0; b < 3; b++) for (int q = 0; q < 2; q++) if (h->halo_tab[b][q]) cudaFree (h->halo_tab[b][q]);
  for (int q = 0; q < h->npinned; q++)
    if (h->pinned_by_us[q] && cudaHostUnregister ((void *)h->pinned[q]) != cudaSuccess) cudaGetLastError ();
  if (h->graph) cudaGraphExecDestroy (h->graph);
  for (int e = 0; e < PG_MAX_EV; e++) if (h->ev0[e]){ cudaEventDestroy (h->ev0[e]); cudaEventDestroy (h->ev1[e]); }
  if (h->stream) cudaStreamDestroy (h->stream);
  free (h);
}

void *pluto_gpu_stream (PlutoGpu *h) { return (void *)h->stream; }

int pluto_gpu_timing (PlutoGpu *h, int enable)
{
  h->timing = enable; h->nev = 0;
  for (int c = 0; c < PG_NCLASS; c++){ h->class_ms[c] = 0.0; h->class_count[c] = 0; }
  return 0;
}

int pluto_gpu_timing_get (PlutoGpu *h, int cls, const char **name, double *ms, long long *launches)
{
  if (cls < 0 || cls >= PG_NCLASS) return 1;
  *name = kClassName[cls]; *ms = h->class_ms[cls]; *launches = h->class_count[cls];
  if (cls == KC_SWEEP_X && h->fuse_xy && h->cfg.arith == PLUTO_GPU_ARITH_FAST && !h->ctu) *name = "sweep_x1x2";
  return 0;
}
long long pluto_gpu_launch_count (const PlutoGpu *h) { return h->launches; }
long long pluto_gpu_device_bytes (const PlutoGpu *h) { return (long long)h->pool_bytes; }

int pluto_gpu_field (PlutoGpu *h, const char *name, double **dev_ptr, long long shape[3], int off[3])
{
  static const char *vn[NVS] = {"rho", "vx1", "vx2", "vx3", "bx1", "bx2", "bx3", "prs"};
  static const char *un[NVS] = {"u_rho", "u_mx1", "u_mx2", "u_mx3", "", "", "", "u_eng"};
  static const char *sn[3] = {"bx1s", "bx2s", "bx3s"};
  *dev_ptr = NULL;
  int buf = 0;                                  // "1:rho" selects state buffer 1
  if (name[0] >= '0' && name[0] <= '2' && name[1] == ':'){ buf = name[0] - '0'; name += 2; }
  if (buf >= h->nbuf) return fail ("state buffer %d does not exist", buf);
  for (int nv = 0; nv < NVS; nv++){
    if (!strcmp (name, vn[nv])) *dev_ptr = h->V[buf][nv];
    if (un[nv][0] && !strcmp (name, un[nv])) *dev_ptr = h->U[nv];
  }
  for (int d = 0; d < 3; d++) if (!strcmp (name, sn[d])) *dev_ptr = h->Bs[buf][d];
  if (!strcmp (name, "ex")) *dev_ptr = h->ex;
  if (!strcmp (name, "ey")) *dev_ptr = h->ey;
  if (!strcmp (name, "ez")) *dev_ptr = h->ez;
  if (!strcmp (name, "exj")) *dev_ptr = h->exj;
  if (!strcmp (name, "exk")) *dev_ptr = h->exk;
  if (!strcmp (name, "eyi")) *dev_ptr = h->eyi;
  if (!strcmp (name, "eyk")) *dev_ptr = h->eyk;
  if (!strcmp (name, "ezi")) *dev_ptr = h->ezi;
  if (!strcmp (name, "ezj")) *dev_ptr = h->ezj;
  if (!strcmp (name, "cdt")) *dev_ptr = h->cdt;
  if (!*dev_ptr) return fail ("unknown field '%s'", name);
  shape[0] = h->g.S1; shape[1] = h->g.T[1] + 2; shape[2] = (h->g.dims == 3 ? h->g.T[2] + 2 : 1);
  for (int d = 0; d < 3; d++) off[d] = h->g.off[d];
  return 0;
}

int pluto_gpu_read_field (PlutoGpu *h, const char *name, double *host)
{
  double *dev; long long shape[3]; int off[3];
  if (pluto_gpu_field (h, name, &dev, shape, off)) return 1;
  CU (cudaSetDevice (h->cfg.device));
  CU (cudaStreamSynchronize (h->stream));
  CU (cudaMemcpy (host, dev, (size_t)(shape[0]*shape[1]*shape[2])*sizeof (double), cudaMemcpyDeviceToHost));
  return 0;
}

// ---------------------------------------------------------------------------
//  host <-> device transfer through the scratch region + pack/unpack kernels
// ---------------------------------------------------------------------------
static void box_set (int lo[3], int hi[3], int l0, int h0, int l1, int h1, int l2, int h2)
{ lo[0] = l0; hi[0] = h0; lo[1] = l1; hi[1] = h1; lo[2] = l2; hi[2] = h2; }

static long long box_count (const int lo[3], const int hi[3])
{ return (long long)(hi[0] - lo[0] + 1)*(hi[1] - lo[1] + 1)*(hi[2] - lo[2] + 1); }

// describe the 11 (3-D) or 8 (2-D) fields of one host layout; returns nf
//   layout 0: interior (.dbl): vc in 8 slots (dead slots skipped but their
//             space kept), staggered with +1 face
//   layout 1: reference Data arrays with ghosts: NVAR live slots only
static int describe_layout (PlutoGpu *h, int layout, int buf, HaloArgs &a,
                            long long seg_off[4], long long seg_len[4])
{
  const Geom &g = h->g;
  const int d3 = (g.dims == 3);
  int nf = 0;
  long long off = 0;
  int lo[3], hi[3];
  if (layout == 0) box_set (lo, hi, g.beg[0], g.end[0], g.beg[1], g.end[1], g.beg[2], g.end[2]);
  else             box_set (lo, hi, 0, g.T[0] - 1, 0, g.T[1] - 1, 0, g.T[2] - 1);
  const long long ncell = box_count (lo, hi);
  seg_off[0] = 0;
  for (int nv = 0; nv < NVS; nv++){
    if (!live_var (h, nv)){ if (layout == 0) off += ncell; continue; }
    a.q[nf] = h->V[buf][nv];
    for (int d = 0; d < 3; d++){ a.lo[nf][d] = lo[d]; a.hi[nf][d] = hi[d]; }
    a.offset[nf] = off; off += ncell; nf++;
  }
  seg_len[0] = off;
  for (int s = 0; s < 3; s++){
    seg_off[1 + s] = off; seg_len[1 + s] = 0;
    if (s == 2 && !d3) continue;
    a.q[nf] = h->Bs[buf][s];
    for (int d = 0; d < 3; d++){ a.lo[nf][d] = lo[d]; a.hi[nf][d] = hi[d]; }
    a.lo[nf][s] -= 1;
    a.offset[nf] = off;
    seg_len[1 + s] = box_count (a.lo[nf], a.hi[nf]);
    off += seg_len[1 + s]; nf++;
  }
  a.nf = nf; a.buf = h->scratch; a.g = g;
  return nf;
}

static int transfer (PlutoGpu *h, int layout, bool upload, double *p0, double *p1, double *p2, double *p3)
{
  CU (cudaSetDevice (h->cfg.device));
  HaloArgs a; memset (&a, 0, sizeof (a));
  long long so[4], sl[4];
  describe_layout (h, layout, 0, a, so, sl);
  double *hp[4] = {p0, p1, p2, p3};
  if ((size_t)(so[3] + sl[3]) > h->scratch_doubles) return fail ("internal: scratch too small");
  if (upload){
    for (int s = 0; s < 4; s++) if (sl[s] > 0){
      if (!hp[s]) return fail ("NULL host pointer for segment %d", s);
      CU (cudaMemcpyAsync (h->scratch + so[s], hp[s], sl[s]*sizeof (double), cudaMemcpyHostToDevice, h->stream));
    }
    if (count (h, DISPATCH (h, launch_halo_unpack) (a, h->stream))) return 1;
    CU (cudaStreamSynchronize (h->stream));
  }else{
    if (count (h, DISPATCH (h, launch_halo_pack) (a, h->stream))) return 1;
    for (int s = 0; s < 4; s++) if (sl[s] > 0){
      if (!hp[s]) return fail ("NULL host pointer for segment %d", s);
      CU (cudaMemcpyAsync (hp[s], h->scratch + so[s], sl[s]*sizeof (double), cudaMemcpyDeviceToHost, h->stream));
    }
    CU (cudaStreamSynchronize (h->stream));
  }
  return 0;
}

int pluto_gpu_upload_interior (PlutoGpu *h, const double *vc, const double *b1, const double *b2, const double *b3)
{ return transfer (h, 0, true, (double *)vc, (double *)b1, (double *)b2, (double *)b3); }
int pluto_gpu_download_interior (PlutoGpu *h, double *vc, double *b1, double *b2, double *b3)
{ return transfer (h, 0, false, vc, b1, b2, b3); }
int pluto_gpu_upload_data (PlutoGpu *h, const double *Vc, const double *s1, const double *s2, const double *s3)
{ return transfer (h, 1, true, (double *)Vc, (double *)s1, (double *)s2, (double *)s3); }
int pluto_gpu_download_data (PlutoGpu *h, double *Vc, double *s1, double *s2, double *s3)
{ return transfer (h, 1, false, Vc, s1, s2, s3); }

// Non-uniform Cartesian grid: the zone widths of every direction, dx_d[0 .. T_d-1] = the reference's grid->dx[d] (ghost zones
// included; set_grid.c builds them from the patches of pluto.ini's [Grid] block).  What changes on the path, with the reference's
// default UNIFORM_CARTESIAN_GRID YES (plm_coeffs.h:23-29: the reconstruction keeps its uniform weights in every CARTESIAN
// build): scrh = dt/dx[i] of the flux difference (rhs.c:195), inv_dl = 1/dx[i] of the inverse time step (update_stage.c:229-235),
// dt/dx2[j], dt/dx3[k], ... of CT_Update (ct_update.c:91-96, 147-152, 202-204) and the face areas of FillMagneticField
// (ct_fill_mag_field.c:108-114).  RK time stepping with LINEAR reconstruction; call after pluto_gpu_create, before the first step.
int pluto_gpu_set_grid (PlutoGpu *h, const double *dx1, const double *dx2, const double *dx3)
{
  CU (cudaSetDevice (h->cfg.device));
  const Geom &g = h->g;
  const double *src[3] = {dx1, dx2, dx3};
  for (int d = 0; d < g.dims; d++){
    if (!src[d]) return fail ("pluto_gpu_set_grid: NULL array for direction %d", d + 1);
    for (int n = 0; n < g.T[d]; n++) if (!(src[d][n] > 0.0)) return fail ("pluto_gpu_set_grid: dx%d[%d] = %g", d + 1, n, src[d][n]);
  }
  // padded rows: the lanes of a sweep that lie beyond the last zone still form an address
  size_t off[4] = {0, 0, 0, 0};
  for (int d = 0; d < 3; d++) off[d + 1] = off[d] + (size_t)(((d < g.dims ? g.T[d] : 1) + 64 + 31) & ~31);
  if (!h->grid_pool){
    const size_t nb = 3*off[3]*sizeof (double);
    if (cudaMalloc (&h->grid_pool, nb) != cudaSuccess){ h->grid_pool = NULL; return fail ("cudaMalloc of %zu bytes (grid) failed", nb); }
    h->pool_bytes += nb;
  }
  double *host = (double *)malloc (3*off[3]*sizeof (double));
  if (!host) return fail ("out of host memory");
  for (size_t q = 0; q < 3*off[3]; q++) host[q] = 1.0;
  for (int d = 0; d < 3; d++){
    h->dxa[d] = (double *)h->grid_pool + off[d];
    h->idxa[d] = (double *)h->grid_pool + off[3] + off[d];
    h->dtxa[d] = (double *)h->grid_pool + 2*off[3] + off[d];
    if (d >= g.dims) continue;
    for (int n = 0; n < g.T[d]; n++){
      host[off[d] + n] = src[d][n];
      host[off[3] + off[d] + n] = 1.0/src[d][n];           // grid->inv_dx (set_geometry.c:244)
    }
  }
  cudaError_t e = cudaMemcpyAsync (h->grid_pool, host, 3*off[3]*sizeof (double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize (h->stream);
  free (host);
  if (e != cudaSuccess) return fail ("pluto_gpu_set_grid: %s", cudaGetErrorString (e));
  h->nu = 1;
  if (h->graph){ cudaGraphExecDestroy (h->graph); h->graph = NULL; }      // the captured step holds the old arguments
  return 0;
}

// UNIFORM_CARTESIAN_GRID NO (Src/States/plm_coeffs.h:23-29): the linear reconstruction takes grid-dependent weights.  The six
// arrays PLM_CoefficientsGet returns for direction dir (plm_coeffs.c:86-104; T_dir entries each, the first and last are never
// used) are handed over as they are, so the host keeps computing them with the reference's own PLM_CoefficientsSet.  Once every
// direction is set the sweeps run their RECON_PLMW variants: dvp = dv[i] wp, dvm = dv[i-1] wm, the limiters of
// plm_coeffs.h:130-152 with cp, cm, vp = v + dv_lim dp, vm = v - dv_lim dm.  RK2 / RK3, LINEAR, plain scheme options.
int pluto_gpu_set_plm_coeffs (PlutoGpu *h, int dir, const double *cp, const double *cm, const double *wp, const double *wm,
                              const double *dp, const double *dm)
{
  CU (cudaSetDevice (h->cfg.device));
  const Geom &g = h->g;
  if (dir < 0 || dir >= g.dims) return fail ("pluto_gpu_set_plm_coeffs: direction %d", dir);
  // PARABOLIC + SHOCK_FLATTENING MULTID: the weights serve the minmod fallback of flagged zones (ppm_states.c:167-181) only
  const bool ppm_flat = (h->cfg.recon == PLUTO_GPU_RECON_PARABOLIC && h->cfg.shock_flattening && !h->ctu);
  if (!ppm_flat && (h->ctu || h->cfg.recon != PLUTO_GPU_RECON_LINEAR || h->cfg.shock_flattening || h->cfg.body_force || h->cfg.en_correction
                    || h->cfg.char_limiting || h->cfg.emf_average == PLUTO_GPU_EMF_UCT_HLL))
    return fail ("pluto_gpu_set_plm_coeffs: grid-dependent reconstruction weights need RK2 / RK3 with LINEAR reconstruction, without "
                 "SHOCK_FLATTENING, BODY_FORCE, CT_EN_CORRECTION, CHAR_LIMITING and UCT_HLL (or PARABOLIC with SHOCK_FLATTENING MULTID)");
  const double *src[6] = {cp, cm, wp, wm, dp, dm};
  for (int q = 0; q < 6; q++) if (!src[q]) return fail ("pluto_gpu_set_plm_coeffs: NULL array");
  int maxT = 1;
  for (int d = 0; d < g.dims; d++) if (g.T[d] > maxT) maxT = g.T[d];
  const size_t row = (size_t)((maxT + 64 + 31) & ~31);
  if (!h->plmc_pool){
    const size_t nb = 18*row*sizeof (double);
    if (cudaMalloc (&h->plmc_pool, nb) != cudaSuccess){ h->plmc_pool = NULL; return fail ("cudaMalloc of %zu bytes (reconstruction weights) failed", nb); }
    CU (cudaMemset (h->plmc_pool, 0, nb));
    h->pool_bytes += nb;
  }
  for (int q = 0; q < 6; q++){
    double *dev = (double *)h->plmc_pool + (size_t)(6*dir + q)*row;
    CU (cudaMemcpy (dev, src[q], (size_t)g.T[dir]*sizeof (double), cudaMemcpyHostToDevice));
    h->plmc[dir][q] = dev;
  }
  h->plmw = (h->cfg.recon == PLUTO_GPU_RECON_LINEAR);
  for (int d = 0; d < g.dims; d++) if (!h->plmc[d][0]) h->plmw = 0;
  if (h->graph){ cudaGraphExecDestroy (h->graph); h->graph = NULL; }
  return 0;
}

// PARABOLIC reconstruction on a non-uniform grid (ppm_states.c:146-150 with the weights of PPM_CoefficientsGet, ppm_coeffs.c:586-609):
// wp[n][-1], wp[n][0], wp[n][1], wp[n][2] of every zone n of direction dir, T_dir entries each.  On the device four consecutive
// doubles per zone, with 32 zones of padding below (the lanes of a sweep that lie outside the row still form an address).
int pluto_gpu_set_ppm_coeffs (PlutoGpu *h, int dir, const double *wm1, const double *w0, const double *w1, const double *w2)
{
  CU (cudaSetDevice (h->cfg.device));
  const Geom &g = h->g;
  if (dir < 0 || dir >= g.dims) return fail ("pluto_gpu_set_ppm_coeffs: direction %d", dir);
  if (h->cfg.recon != PLUTO_GPU_RECON_PARABOLIC) return fail ("pluto_gpu_set_ppm_coeffs: the configuration has no PARABOLIC reconstruction");
  if (!wm1 || !w0 || !w1 || !w2) return fail ("pluto_gpu_set_ppm_coeffs: NULL array");
  int maxT = 1;
  for (int d = 0; d < g.dims; d++) if (g.T[d] > maxT) maxT = g.T[d];
  const size_t row = 4*(size_t)((maxT + 128 + 31) & ~31);
  if (!h->ppmc_pool){
    const size_t nb = 3*row*sizeof (double);
    if (cudaMalloc (&h->ppmc_pool, nb) != cudaSuccess){ h->ppmc_pool = NULL; return fail ("cudaMalloc of %zu bytes (parabolic weights) failed", nb); }
    CU (cudaMemset (h->ppmc_pool, 0, nb));
    h->pool_bytes += nb;
  }
  double *host = (double *)malloc ((size_t)g.T[dir]*4*sizeof (double));
  if (!host) return fail ("out of host memory");
  for (int n = 0; n < g.T[dir]; n++){ host[4*n] = wm1[n]; host[4*n + 1] = w0[n]; host[4*n + 2] = w1[n]; host[4*n + 3] = w2[n]; }
  double *dev = (double *)h->ppmc_pool + (size_t)dir*row + 4*32;
  const cudaError_t e = cudaMemcpy (dev, host, (size_t)g.T[dir]*4*sizeof (double), cudaMemcpyHostToDevice);
  free (host);
  if (e != cudaSuccess) return fail ("cudaMemcpy (parabolic weights): %s", cudaGetErrorString (e));
  h->ppmc[dir] = dev;
  if (h->graph){ cudaGraphExecDestroy (h->graph); h->graph = NULL; }
  return 0;
}

int pluto_gpu_set_body_force (PlutoGpu *h, const double *g1, const double *g2, const double *g3)
{
  CU (cudaSetDevice (h->cfg.device));
  if (!(h->cfg.body_force & 1)) return fail ("pluto_gpu_set_body_force: the configuration has no BODY_FORCE VECTOR (body_force & 1)");
  const Geom &g = h->g;
  const size_t tot_al = ((size_t)g.tot + 31) & ~(size_t)31;
  if (!h->gfield_pool){
    const size_t nb = (size_t)g.dims*tot_al*sizeof (double);
    if (cudaMalloc (&h->gfield_pool, nb) != cudaSuccess){ h->gfield_pool = NULL; return fail ("cudaMalloc of %zu bytes (body force) failed", nb); }
    CU (cudaMemset (h->gfield_pool, 0, nb));
    h->pool_bytes += nb;
  }
  const double *src[3] = {g1, g2, g3};
  const long long ncell = (long long)g.T[0]*g.T[1]*g.T[2];
  if ((size_t)ncell > h->scratch_doubles) return fail ("internal: scratch too small");
  for (int d = 0; d < g.dims; d++){
    if (!src[d]) return fail ("pluto_gpu_set_body_force: NULL array for component %d", d + 1);
    double *dev = (double *)h->gfield_pool + (size_t)d*tot_al;
    HaloArgs a; memset (&a, 0, sizeof (a));
    a.q[0] = dev; a.nf = 1; a.offset[0] = 0; a.buf = h->scratch; a.g = g;
    a.lo[0][0] = a.lo[0][1] = a.lo[0][2] = 0;
    a.hi[0][0] = g.T[0] - 1; a.hi[0][1] = g.T[1] - 1; a.hi[0][2] = g.T[2] - 1;
    CU (cudaMemcpyAsync (h->scratch, src[d], (size_t)ncell*sizeof (double), cudaMemcpyHostToDevice, h->stream));
    if (count (h, DISPATCH (h, launch_halo_unpack) (a, h->stream))) return 1;
    CU (cudaStreamSynchronize (h->stream));
    h->gfield[d] = dev;
  }
  if (h->graph){ cudaGraphExecDestroy (h->graph); h->graph = NULL; }      // the captured step holds the old arguments
  return 0;
}

int pluto_gpu_set_body_potential (PlutoGpu *h, const double *phic, const double *pf1, const double *pf2, const double *pf3)
{
  CU (cudaSetDevice (h->cfg.device));
  if (!(h->cfg.body_force & 2)) return fail ("pluto_gpu_set_body_potential: the configuration has no BODY_FORCE POTENTIAL (body_force & 2)");
  const Geom &g = h->g;
  const size_t tot_al = ((size_t)g.tot + 31) & ~(size_t)31;
  if (!h->phi_pool){
    const size_t nb = (size_t)(1 + g.dims)*tot_al*sizeof (double);
    if (cudaMalloc (&h->phi_pool, nb) != cudaSuccess){ h->phi_pool = NULL; return fail ("cudaMalloc of %zu bytes (potential) failed", nb); }
    CU (cudaMemset (h->phi_pool, 0, nb));
    h->pool_bytes += nb;
  }
  const double *src[4] = {phic, pf1, pf2, pf3};
  for (int q = 0; q < 1 + g.dims; q++){
    if (!src[q]) return fail ("pluto_gpu_set_body_potential: NULL array %d", q);
    double *dev = (double *)h->phi_pool + (size_t)q*tot_al;
    HaloArgs a; memset (&a, 0, sizeof (a));
    a.q[0] = dev; a.nf = 1; a.offset[0] = 0; a.buf = h->scratch; a.g = g;
    a.lo[0][0] = a.lo[0][1] = a.lo[0][2] = 0;
    a.hi[0][0] = g.T[0] - 1; a.hi[0][1] = g.T[1] - 1; a.hi[0][2] = g.T[2] - 1;
    if (q > 0) a.lo[0][q - 1] = -1;                      // faces of direction q-1: one more, starting at -1/2
    const long long cnt = box_count (a.lo[0], a.hi[0]);
    if ((size_t)cnt > h->scratch_doubles) return fail ("internal: scratch too small");
    CU (cudaMemcpyAsync (h->scratch, src[q], (size_t)cnt*sizeof (double), cudaMemcpyHostToDevice, h->stream));
    if (count (h, DISPATCH (h, launch_halo_unpack) (a, h->stream))) return 1;
    CU (cudaStreamSynchronize (h->stream));
    if (q == 0) h->phic = dev; else h->phif[q - 1] = dev;
  }
  if (h->graph){ cudaGraphExecDestroy (h->graph); h->graph = NULL; }
  return 0;
}

// ---------------------------------------------------------------------------
//  Boundary (boundary.c:137-293) on state buffer `buf`, one dimension
// ---------------------------------------------------------------------------
static int boundary_dim (PlutoGpu *h, int buf, int dim)
{
  const Geom &g = h->g;
  BcArgs a; memset (&a, 0, sizeof (a));
  a.g = g;
  for (int s = 0; s < 3; s++) a.Bs[s] = h->Bs[buf][s];
  int nf = 0, nfill = 0;
  for (int hs = 0; hs < 2; hs++){
    const int side = 2*dim + hs;
    const int type = h->cfg.bc[side];
    if (type == PLUTO_GPU_BC_SHARED) continue;                 // boundary.c:139
    const bool eqt = (type == PLUTO_GPU_BC_EQTSYMMETRIC);
    const bool fill = (type == PLUTO_GPU_BC_OUTFLOW || type == PLUTO_GPU_BC_REFLECTIVE || eqt);
    // cell box of this side: ghost layers in `dim`, full extent elsewhere
    int lo[3] = {0, 0, 0}, hi[3] = {g.T[0] - 1, g.T[1] - 1, g.T[2] - 1};
    if (hs == 0) hi[dim] = g.beg[dim] - 1; else lo[dim] = g.end[dim] + 1;
    for (int nv = 0; nv < NVS; nv++){
      if (!live_var (h, nv)) continue;
      // outflow: the cell-centred normal field of the ghost zones is the average of the
      // filled faces (CT_AverageNormalMagField, boundary.c:188-189), written by the fill job
      if (type == PLUTO_GPU_BC_OUTFLOW && nv == 4 + dim) continue;
      BcField &f = a.f[nf++];
      f.q = h->V[buf][nv];
      for (int d = 0; d < 3; d++){ f.lo[d] = lo[d]; f.hi[d] = hi[d]; }
      f.sign = (nv == 1 + dim || nv == 4 + dim) ? -1 : 1;     // FlipSign, boundary.c:318-436
      if (eqt && nv >= 4 && nv <= 6) f.sign = (nv == 4 + dim) ? 1 : -1;        // EQTSYMMETRIC: Bn -> Bn, Bt -> -Bt (:423-427)
      f.side = side; f.type = type;
    }
    for (int s = 0; s < g.dims; s++){
      // staggered boxes (boundary.c:157-164): one more face on the low end of
      // their own direction; the normal component is skipped by outflow and
      // reflective conditions (:175-180, 199-204) and rebuilt from div B = 0
      if (s == dim && type != PLUTO_GPU_BC_PERIODIC) continue;
      BcField &f = a.f[nf++];
      f.q = h->Bs[buf][s];
      for (int d = 0; d < 3; d++){ f.lo[d] = lo[d]; f.hi[d] = hi[d]; }
      // Periodic normal component: the reference fills the low side first, face IBEG-1 <-
      // face IEND, and the high side would then copy it back onto face IEND unchanged
      // (boundary.c:157-164 with the sides in sequence).  Both sides share a launch here,
      // so the high-side box leaves that face out.
      if (!(s == dim && hs == 1)) f.lo[s] -= 1;
      f.sign = eqt ? -1 : 1;                                  // tangential staggered components
      f.side = side; f.type = type;
    }
    if (fill){
      a.fill[nfill].Bc = (type == PLUTO_GPU_BC_OUTFLOW ? h->V[buf][4 + dim] : NULL);
      a.fill[nfill].side = side; a.fill[nfill].type = type;
      nfill++;
    }
  }
  a.nf = nf; a.nfill = nfill;
  for (int d = 0; d < 3; d++) a.dxa[d] = h->nu ? h->dxa[d] : NULL;
  if (nf + nfill == 0) return 0;
  TIMED (h, KC_BC, count (h, DISPATCH (h, launch_bc) (a, h->stream)));
  return 0;
}

// ---------------------------------------------------------------------------
//  one RK stage without its Boundary call: UpdateStage + CT + stage average +
//  ConsToPrim.  in = buffer holding the stage's input state.
// ---------------------------------------------------------------------------
struct StagePlan { int in, out, combine; double w0, wc; };

static StagePlan stage_plan (const PlutoGpu *h, int stage)
{
  StagePlan p; p.w0 = p.wc = 0.0; p.combine = 0; p.in = 0; p.out = 1;
  if (h->ctu){ p.in = 0; p.out = 0; }           // one Boundary call per step, the new state replaces the old one
  else if (h->cfg.rk_order == 2){
    if (stage == 1){ p.in = 0; p.out = 1; }
    else           { p.in = 1; p.out = 0; p.combine = 1; p.w0 = 0.5; p.wc = 0.5; }     // rk_step.c:18-20
  }else{
    if (stage == 1){ p.in = 0; p.out = 1; }
    else if (stage == 2){ p.in = 1; p.out = 2; p.combine = 1; p.w0 = 0.75; p.wc = 0.25; } // rk_step.c:21-23
    else { p.in = 2; p.out = 0; p.combine = 2; }                                         // rk_step.c:226-233
  }
  return p;
}

// Non-uniform grid: dt/dx[n] of every zone and direction from the dt in dtdev[3] (the host's, or the one next_dt_kernel
// left there): the quotient the reference forms zone by zone (rhs.c:195 scrh = dt/dx[i]; ct_update.c:91-96), IEEE division.
__global__ void dtx_fill_kernel (const double *dtdev, const double *dx0, const double *dx1, const double *dx2,
                                 double *o0, double *o1, double *o2, int n0, int n1, int n2)
{
  const double dt = dtdev[3];
  for (int n = blockIdx.x*blockDim.x + threadIdx.x; n < n0 + n1 + n2; n += gridDim.x*blockDim.x){
    if      (n < n0)      o0[n] = __ddiv_rn (dt, dx0[n]);
    else if (n < n0 + n1) o1[n - n0] = __ddiv_rn (dt, dx1[n - n0]);
    else                  o2[n - n0 - n1] = __ddiv_rn (dt, dx2[n - n0 - n1]);
  }
}
static int refresh_dtx (PlutoGpu *h)
{
  if (!h->nu) return 0;
  const Geom &g = h->g;
  dtx_fill_kernel<<<4, 256, 0, h->stream>>>(h->dtdev, h->dxa[0], h->dxa[1], h->dxa[2], h->dtxa[0], h->dtxa[1], h->dtxa[2],
                                             g.T[0], g.T[1], g.dims == 3 ? g.T[2] : 0);
  return count (h, pg_launch_status ());
}

// dt/dx (rhs.c:195, ct_update.c:87-89) goes to the device once per step
static int set_dt (PlutoGpu *h, double dt)
{
  for (int d = 0; d < 3; d++) h->dthost[d] = dt/h->g.dx[d];
  h->dthost[3] = dt;
  for (int d = 0; d < 3; d++) h->dthost[4 + d] = (0.5*dt)/h->g.dx[d];      // CTU half step (ctu_step.c:447-455)
  h->dthost[7] = 0.5*dt;
  CU (cudaMemcpyAsync (h->dtdev, h->dthost, 8*sizeof (double), cudaMemcpyHostToDevice, h->stream));
  return refresh_dtx (h);
}

// part: 0 = the whole stage; 1 = everything up to the stage completion of the SHELL (the
// zones within nghost of a SHARED side, whose new values the neighbouring blocks need);
// 2 = the stage completion of the remaining zones.  1 followed by 2 equals 0.
enum { PART_ALL = 0, PART_SHELL = 1, PART_INTERIOR = 2 };

static int launch_final_boxes (PlutoGpu *h, FinalArgs &f, int part)
{
  const Geom &g = h->g;
  int mlo[3] = {0, 0, 0}, mhi[3] = {0, 0, 0};
  bool split = (part != PART_ALL);
  for (int d = 0; d < g.dims && split; d++){
    // x1 slabs are a full warp wide when the block allows: an ng-wide slab would be read
    // and written in 16-24 byte pieces per row
    const int w = (d == 0 ? h->shell_w : g.ng);
    if (h->cfg.bc[2*d] == PLUTO_GPU_BC_SHARED) mlo[d] = w;
    if (h->cfg.bc[2*d + 1] == PLUTO_GPU_BC_SHARED) mhi[d] = w;
    if (mlo[d] + mhi[d] >= g.n[d]) split = false;            // block too thin: no interior
  }
  if (!split){
    if (part == PART_INTERIOR) return 0;
    for (int d = 0; d < 3; d++){ f.box_lo[d] = 0; f.box_n[d] = g.n[d]; }
    TIMED (h, KC_FINAL, count (h, DISPATCH (h, launch_final) (f, h->stream)));
    return 0;
  }
  // inner box and the (up to six) disjoint slabs around it: x3 slabs span all of x1, x2;
  // x2 slabs the inner x3 range; x1 slabs the inner x2 and x3 ranges
  int ilo[3], in_[3];
  for (int d = 0; d < 3; d++){ ilo[d] = mlo[d]; in_[d] = g.n[d] - mlo[d] - mhi[d]; }
  if (part == PART_INTERIOR){
    for (int d = 0; d < 3; d++){ f.box_lo[d] = ilo[d]; f.box_n[d] = in_[d]; }
    TIMED (h, KC_FINAL, count (h, DISPATCH (h, launch_final) (f, h->stream)));
    return 0;
  }
  // all shell slabs in ONE launch (blockIdx.y = slab)
  f.nbox = 0;
  for (int d = g.dims - 1; d >= 0; d--) for (int hs = 0; hs < 2; hs++){
    const int w = hs ? mhi[d] : mlo[d];
    if (w == 0) continue;
    int *lo = f.boxes_lo[f.nbox], *nn = f.boxes_n[f.nbox];
    for (int q = 0; q < 3; q++){
      if (q > d){ lo[q] = ilo[q]; nn[q] = in_[q]; }      // already covered by the slabs of q
      else      { lo[q] = 0;      nn[q] = g.n[q]; }
    }
    lo[d] = hs ? g.n[d] - w : 0; nn[d] = w;
    f.nbox++;
  }
  if (f.nbox){
    TIMED (h, KC_FINAL, count (h, DISPATCH (h, launch_final) (f, h->stream)));
    f.nbox = 0;
  }
  return 0;
}

static int run_ctu (PlutoGpu *h, int part);

static int run_stage (PlutoGpu *h, int stage, int part = PART_ALL)
{
  const Geom &g = h->g;
  // CTU: the whole step is one "stage" (one Boundary call); shell / interior split the last kernel only
  if (h->ctu) return run_ctu (h, part);
  const StagePlan sp = stage_plan (h, stage);
  FinalArgs f; memset (&f, 0, sizeof (f));
  for (int nv = 0; nv < NVS; nv++){ f.U[nv] = h->U[nv]; f.V0[nv] = h->V[0][nv]; f.Vout[nv] = h->V[sp.out][nv]; }
  for (int d = 0; d < 3; d++) f.Bs[d] = h->Bs[sp.out][d];
  f.red = h->red; f.g = g; f.ph = h->ph; f.w0 = sp.w0; f.wc = sp.wc; f.combine = sp.combine;
  for (int nv = 0; nv < NVS; nv++) f.Uw[nv] = h->U[nv];
  f.en_corr = h->cfg.en_correction; f.dtp = h->dtdev;
  f.gs = h->nu; for (int d = 0; d < 3; d++) f.dtx[d] = h->nu ? h->dtxa[d] : h->dtdev + d;
  for (int nv = 0; nv < NVS; nv++) f.Vin[nv] = h->V[sp.in][nv];
  f.exj = h->exj; f.exk = h->exk; f.eyi = h->eyi; f.eyk = h->eyk; f.ezi = h->ezi; f.ezj = h->ezj;
  for (int d = 0; d < 3; d++) f.fbn[d] = h->fbn[d];
  f.write_u = 0;
  if (stage < h->cfg.rk_order && h->cfg.arith == PLUTO_GPU_ARITH_EXACT) f.write_u = (sp.combine || f.en_corr ? 1 : 2);   // the energy correction stays in Uc
  // CT_Update inside the stage completion (one launch and one pass over the new field less); not with CT_EN_CORRECTION
  // -- and not where the new field overwrites the t^n field it is averaged with (the last RK stage writes buffer 0 in place:
  // a zone would read its low neighbour's face after that neighbour has replaced it)
  f.fuse_ct = (h->fuse_ct && !h->nu && !f.en_corr && !(sp.combine && sp.out == 0));
  f.ex = h->ex; f.ey = h->ey; f.ez = h->ez;
  for (int d = 0; d < 3; d++){ f.Bs_in[d] = h->Bs[sp.in][d]; f.Bs0[d] = h->Bs[0][d]; f.Bs_out[d] = h->Bs[sp.out][d]; }
  if (!f.fuse_ct) for (int nv = 0; nv < NVS; nv++) f.R3[nv] = h->R3[nv];
  if (part == PART_INTERIOR) return launch_final_boxes (h, f, part);

  if (h->flag && stage == 1){
    // FlagShock on the stage-1 state, ghost zones filled (rk_step.c:85-88); the flags hold for the whole step
    FlagArgs fa; memset (&fa, 0, sizeof (fa));
    for (int d = 0; d < 3; d++) fa.vx[d] = h->V[sp.in][1 + d];
    fa.prs = h->V[sp.in][7]; fa.shock = h->shock; fa.flag = h->flag; fa.g = g;
    for (int d = 0; d < 3; d++) fa.dxa[d] = h->nu ? h->dxa[d] : NULL;
    TIMED (h, KC_BC, count (h, DISPATCH (h, launch_flag_shock) (fa, h->stream)));
  }
  SweepArgs s; memset (&s, 0, sizeof (s));
  s.flag = h->flag;
  for (int nv = 0; nv < NVS; nv++){ s.V[nv] = h->V[sp.in][nv]; s.U[nv] = h->U[nv]; }
  s.cdt = h->cdt; s.red = h->red; s.g = g; s.ph = h->ph; s.dtp = h->dtdev;
  if (!f.fuse_ct) for (int nv = 0; nv < NVS; nv++) s.R3[nv] = h->R3[nv];
  s.stage1 = (stage == 1);
  s.limiter = h->cfg.limiter;
  s.char_lim = h->cfg.char_limiting;
  s.avg = h->cfg.emf_average;
  const bool bf = h->cfg.body_force != 0;
  for (int d = 0; d < 3; d++) s.grav[d] = h->cfg.grav[d];
  s.gf2 = h->gfield[1];
  s.bfv = h->cfg.body_force & 1;
  s.phic = h->phic; s.phif2 = h->phif[1];
  if ((h->cfg.body_force & 2) && !h->phic) return fail ("BODY_FORCE POTENTIAL: call pluto_gpu_set_body_potential before the first step");
  // EXACT: later stages continue from the conservative state the previous stage
  // left (as the reference does); FAST: rebuild it from the primitives, which
  // saves reading U in the x1 sweep and differs by round-off only
  s.u_from_v = (stage == 1 || h->cfg.arith == PLUTO_GPU_ARITH_FAST);
  // FAST arithmetic: x1 and x2 are solved by ONE fused kernel (one pass over the
  // primitives, U written once); EXACT keeps one kernel per direction, which continues
  // from the conservative state of the previous stage as the reference does
  const bool fuse_xy = h->fuse_xy && h->cfg.arith == PLUTO_GPU_ARITH_FAST;
  for (int dir = 0; dir < g.dims; dir++){
    s.Bn = h->Bs[sp.in][dir];
    s.inv_dl = 1.0/g.dx[dir];              // set_geometry.c (inv_dx)
    s.gs = h->nu;
    s.dtx = h->nu ? h->dtxa[dir] : h->dtdev + dir;
    s.idl = h->idxa[dir];
    s.last_dir = (dir == g.dims - 1);
    s.sv = h->sv[dir];
    s.fbn = h->fbn[dir];
    s.gf = h->gfield[dir];
    s.phif = h->phif[dir];
    if (dir == 0){ s.e1 = h->ezi; s.e2 = h->eyi; }
    else if (dir == 1){ s.e1 = h->ezj; s.e2 = h->exj; }
    else { s.e1 = h->eyk; s.e2 = h->exk; }
    for (int c = 0; c < 3; c++) s.dvel[c] = h->dvel[c][dir];
    const int recon = h->plmw ? 2 /* RECON_PLMW */ : h->cfg.recon;
    if (h->flag && h->cfg.recon == PLUTO_GPU_RECON_PARABOLIC && !h->plmc[dir][0])
      return fail ("SHOCK_FLATTENING MULTID with PARABOLIC reconstruction: hand over the weights of PLM_CoefficientsGet first "
                   "(pluto_gpu_set_plm_coeffs; ppm_states.c:167-181 takes them for the zones it flattens)");
    for (int q = 0; q < 6; q++){ s.pc[q] = h->plmc[dir][q]; s.pc2[q] = h->plmc[1][q]; }
    if (h->nu && h->cfg.recon == PLUTO_GPU_RECON_PARABOLIC){
      for (int d = 0; d < g.dims; d++) if (!h->ppmc[d])
        return fail ("PARABOLIC reconstruction on a non-uniform grid: hand over the weights of PPM_CoefficientsGet first (pluto_gpu_set_ppm_coeffs)");
    }
    s.qc = h->ppmc[dir]; s.qc2 = h->ppmc[1];
    if (dir > 0 || fuse_xy){
      // zones per thread along a marching sweep: long enough to amortise the
      // extra face per chunk, short enough to fill the 148 SMs (2-D grids have
      // few pencils)
      const int mdir = (dir == 0 ? 1 : dir);
      const int td = (mdir == 1 ? 2 : 1);
      long long npen = (long long)(g.n[0] + 2)*(g.dims == 3 ? g.n[td] + 2 : 1);
      if (dir == 0){                       // fused: 32 lanes per x1 segment of (30 or 29) zones
        const int stride = 32 - (recon == PLUTO_GPU_RECON_PARABOLIC ? 2 : 1) - 1;
        npen = (long long)((g.n[0] + stride - 1)/stride)*32*(g.dims == 3 ? g.n[2] + 2 : 1);
      }
      const long long want = (228000 + npen - 1)/npen;               // chunks for ~4 waves of threads
      long long len = (g.n[mdir] + want - 1)/want;
      if (len < 4) len = 4;
      if (len > h->march_chunk) len = h->march_chunk;
      if (len > g.n[mdir]) len = g.n[mdir];
      s.chunk_len = (int)len;
      s.nchunk = (g.n[mdir] + s.chunk_len - 1)/s.chunk_len;
      // default: the launcher chooses the chunk count that fills the SMs in whole rounds (plan_chunks,
      // sweep_kernels.cuh); the figures above are what PLUTO_GPU_NO_PLAN=1 falls back to
      s.plan = h->plan;
      if (s.plan){ s.chunk_len = (g.n[mdir] < 128 ? g.n[mdir] : 128); s.nchunk = (g.n[mdir] + s.chunk_len - 1)/s.chunk_len; }
    }
    int r;
    if (dir == 0 && fuse_xy){
      s.Bn2 = h->Bs[sp.in][1]; s.e3 = h->ezj; s.e4 = h->exj; s.sv2 = h->sv[1];
      for (int c = 0; c < 3; c++) s.dvel2[c] = h->dvel[c][1];
      s.inv_dl2 = 1.0/g.dx[1];
      s.dtx2 = h->nu ? h->dtxa[1] : h->dtdev + 1;
      s.idl2 = h->idxa[1];
      for (int q = 0; q < 3; q++) s.Ec[q] = h->Ec[q];
      s.last_dir = (g.dims == 2);
      s.tma = h->tma;
      const int te = tbegin (h, KC_SWEEP_X);
      if      (h->cfg.solver == PLUTO_GPU_SOLVER_HLLD) r = pg_fast::launch_sweep_xy_hlld (recon, s, h->stream, bf);
      else if (h->cfg.solver == PLUTO_GPU_SOLVER_HLL)  r = pg_fast::launch_sweep_xy_hll  (recon, s, h->stream, bf);
      else if (h->cfg.solver == PLUTO_GPU_SOLVER_HLLC)  r = pg_fast::launch_sweep_xy_hllc  (recon, s, h->stream, bf);
      else if (h->cfg.solver == PLUTO_GPU_SOLVER_TVDLF) r = pg_fast::launch_sweep_xy_tvdlf (recon, s, h->stream, bf);
      else                                             r = pg_fast::launch_sweep_xy_roe  (recon, s, h->stream, bf);
      tend (h, te);
      if (count (h, r)) return 1;
      dir = 1;                             // x2 is done
      continue;
    }
    const int te = tbegin (h, KC_SWEEP_X + dir);
    if      (h->cfg.solver == PLUTO_GPU_SOLVER_HLLD) r = DISPATCH (h, launch_sweep_hlld) (dir, recon, s, h->stream, bf);
    else if (h->cfg.solver == PLUTO_GPU_SOLVER_HLL)  r = DISPATCH (h, launch_sweep_hll)  (dir, recon, s, h->stream, bf);
    else if (h->cfg.solver == PLUTO_GPU_SOLVER_HLLC)  r = DISPATCH (h, launch_sweep_hllc)  (dir, recon, s, h->stream, bf);
    else if (h->cfg.solver == PLUTO_GPU_SOLVER_TVDLF) r = DISPATCH (h, launch_sweep_tvdlf) (dir, recon, s, h->stream, bf);
    else                                             r = DISPATCH (h, launch_sweep_roe)  (dir, recon, s, h->stream, bf);
    tend (h, te);
    if (count (h, r)) return 1;
  }

  CtArgs c; memset (&c, 0, sizeof (c));
  for (int nv = 0; nv < NVS; nv++) c.V[nv] = h->V[sp.in][nv];
  c.exj = h->exj; c.exk = h->exk; c.eyi = h->eyi; c.eyk = h->eyk; c.ezi = h->ezi; c.ezj = h->ezj;
  c.svx = h->sv[0]; c.svy = h->sv[1]; c.svz = h->sv[2];
  c.ex = h->ex; c.ey = h->ey; c.ez = h->ez;
  for (int d = 0; d < 3; d++){
    c.Bs_in[d] = h->Bs[sp.in][d]; c.Bs0[d] = h->Bs[0][d]; c.Bs_out[d] = h->Bs[sp.out][d];
  }
  c.g = g; c.w0 = sp.w0; c.wc = sp.wc; c.combine = sp.combine; c.dtp = h->dtdev;
  c.gs = h->nu; c.dts = 1.0;
  for (int d = 0; d < 3; d++) c.dtx[d] = h->nu && d < g.dims ? h->dtxa[d] : h->dtdev + d;
  if (fuse_xy) for (int q = 0; q < 3; q++) c.Ec[q] = h->Ec[q];
  c.avg = h->cfg.emf_average;
  for (int q = 0; q < 3; q++) for (int d = 0; d < 3; d++) c.dvel[q][d] = h->dvel[q][d];
  TIMED (h, KC_CT_EMF, count (h, DISPATCH (h, launch_ct_emf) (c, h->stream)));
  if (!f.fuse_ct) TIMED (h, KC_CT_UPDATE, count (h, DISPATCH (h, launch_ct_update) (c, h->stream)));

  return launch_final_boxes (h, f, part);
}

// ---------------------------------------------------------------------------
//  corner transport upwind: the body of AdvanceStep in ctu_step.c:142-727 after its Boundary
//  call.  Buffer 0 holds t^n (and receives t^n + dt), buffer 1 the half-step state.
// ---------------------------------------------------------------------------
static int march_chunks (const PlutoGpu *h, int dir, int nzones, int ext, int *chunk_len)
{
  const Geom &g = h->g;
  const int td = (dir == 1 ? 2 : 1);
  const long long npen = (long long)(g.n[0] + 2*ext)*(g.dims == 3 ? g.n[td] + 2*ext : 1);
  const long long want = (228000 + npen - 1)/npen;               // chunks for ~4 waves of threads
  long long len = (nzones + want - 1)/want;
  if (len < 4) len = 4;
  if (len > h->march_chunk) len = h->march_chunk;
  if (len > nzones) len = nzones;
  *chunk_len = (int)len;
  return (nzones + (int)len - 1)/(int)len;
}

static int run_ctu (PlutoGpu *h, int part)
{
  const Geom &g = h->g;
  FinalArgs f; memset (&f, 0, sizeof (f));
  for (int nv = 0; nv < NVS; nv++){ f.U[nv] = h->U[nv]; f.V0[nv] = h->V[0][nv]; f.Vout[nv] = h->V[0][nv]; f.Uw[nv] = h->U[nv]; }
  for (int d = 0; d < 3; d++) f.Bs[d] = h->Bs[0][d];
  f.red = h->red; f.g = g; f.ph = h->ph; f.combine = 0; f.write_u = 0;
  f.en_corr = h->cfg.en_correction; f.dtp = h->dtdev;           // Uc[B] = U^n[B] + the corrector's induction right-hand sides
  f.gs = h->nu; for (int d = 0; d < 3; d++) f.dtx[d] = h->nu ? h->dtxa[d] : h->dtdev + d;
  for (int nv = 0; nv < NVS; nv++) f.Vin[nv] = h->V[0][nv];
  f.exj = h->exj; f.exk = h->exk; f.eyi = h->eyi; f.eyk = h->eyk; f.ezi = h->ezi; f.ezj = h->ezj;
  for (int d = 0; d < 3; d++) f.fbn[d] = h->fbn[d];
  // multi-GPU overlap: the new state of the zones next to SHARED sides first (PART_SHELL), so that the ONE
  // exchange of the next step travels while the remaining zones are mapped to primitives (PART_INTERIOR)
  if (part == PART_INTERIOR) return launch_final_boxes (h, f, part);
  if (h->flag){                          // FlagShock on V^n (ctu_step.c:240-244)
    FlagArgs fa; memset (&fa, 0, sizeof (fa));
    for (int d = 0; d < 3; d++) fa.vx[d] = h->V[0][1 + d];
    fa.prs = h->V[0][7]; fa.shock = h->shock; fa.flag = h->flag; fa.g = g;
    for (int d = 0; d < 3; d++) fa.dxa[d] = h->nu ? h->dxa[d] : NULL;
    TIMED (h, KC_BC, count (h, DISPATCH (h, launch_flag_shock) (fa, h->stream)));
  }
  CtuArgs s; memset (&s, 0, sizeof (s));
  for (int nv = 0; nv < NVS; nv++){ s.V0[nv] = h->V[0][nv]; s.U[nv] = h->U[nv]; s.Vh[nv] = h->V[1][nv]; }
  for (int d = 0; d < 3; d++){
    s.Bs0[d] = h->Bs[0][d]; s.Bsh[d] = h->Bs[1][d];
    for (int nv = 0; nv < NVS; nv++) s.rhs[d][nv] = h->rhs3[d][nv];
  }
  s.red = h->red; s.flag = h->flag; s.g = g; s.ph = h->ph; s.dtp = h->dtdev; s.limiter = h->cfg.limiter;
  s.chtr = (h->cfg.time_stepping == PLUTO_GPU_TS_CHAR_TRACING);
  s.char_lim = h->cfg.char_limiting;
  s.en_corr = h->cfg.en_correction;
  s.bf = h->cfg.body_force & 1;
  for (int d = 0; d < 3; d++) s.grav[d] = h->cfg.grav[d];
  s.phic = h->phic;
  if ((h->cfg.body_force & 2) && !h->phic) return fail ("BODY_FORCE POTENTIAL: call pluto_gpu_set_body_potential before the first step");

  CtArgs c; memset (&c, 0, sizeof (c));
  c.exj = h->exj; c.exk = h->exk; c.eyi = h->eyi; c.eyk = h->eyk; c.ezi = h->ezi; c.ezj = h->ezj;
  c.svx = h->sv[0]; c.svy = h->sv[1]; c.svz = h->sv[2];
  c.ex = h->ex; c.ey = h->ey; c.ez = h->ez;
  c.g = g;

  for (int phase = 0; phase < 2; phase++){
    for (int dir = 0; dir < g.dims; dir++){
      s.inv_dl = 1.0/g.dx[dir];
      s.gs = h->nu; s.dtx = h->nu ? h->dtxa[dir] : h->dtdev + dir; s.idl = h->idxa[dir]; s.dxz = h->nu ? h->dxa[dir] : NULL;
      s.sv = h->sv[dir];
      s.fbn = h->fbn[dir];
      s.gf = h->gfield[dir];
      s.phif = h->phif[dir];
      if (dir == 0){ s.e1 = h->ezi; s.e2 = h->eyi; }
      else if (dir == 1){ s.e1 = h->ezj; s.e2 = h->exj; }
      else { s.e1 = h->eyk; s.e2 = h->exk; }
      if (dir > 0) s.nchunk = march_chunks (h, dir, g.n[dir] + (phase == 0 ? 2 : 0), phase == 0 ? 2 : 1, &s.chunk_len);
      int r;
      const int te = tbegin (h, KC_SWEEP_X + dir);
      if      (h->cfg.solver == PLUTO_GPU_SOLVER_HLLD) r = DISPATCH (h, launch_ctu_sweep_hlld) (dir, phase, s, h->stream);
      else if (h->cfg.solver == PLUTO_GPU_SOLVER_HLL)  r = DISPATCH (h, launch_ctu_sweep_hll)  (dir, phase, s, h->stream);
      else if (h->cfg.solver == PLUTO_GPU_SOLVER_HLLC)  r = DISPATCH (h, launch_ctu_sweep_hllc)  (dir, phase, s, h->stream);
      else if (h->cfg.solver == PLUTO_GPU_SOLVER_TVDLF) r = DISPATCH (h, launch_ctu_sweep_tvdlf) (dir, phase, s, h->stream);
      else                                             r = DISPATCH (h, launch_ctu_sweep_roe)  (dir, phase, s, h->stream);
      tend (h, te);
      if (count (h, r)) return 1;
    }
    if (phase == 0){
      // emf of the predictor fluxes on the extended ranges (CT_EMF_IntegrateToCorner returns at once in the
      // predictor, ct_emf_average.c:90-92: UCT_CONTACT reduces to the arithmetic average), B^{n+1/2} on the
      // faces of DOM+-1, V^{n+1/2} on DOM+-1 (ctu_step.c:447-497)
      for (int nv = 0; nv < NVS; nv++) c.V[nv] = h->V[0][nv];
      for (int d = 0; d < 3; d++){ c.Bs_in[d] = h->Bs[0][d]; c.Bs0[d] = h->Bs[0][d]; c.Bs_out[d] = h->Bs[1][d]; }
      c.avg = (h->cfg.emf_average == PLUTO_GPU_EMF_UCT_CONTACT ? PLUTO_GPU_EMF_ARITHMETIC : h->cfg.emf_average);
      c.ext = 1; c.combine = 0; c.dtp = h->dtdev + 4;
      c.gs = h->nu; c.dts = (h->nu ? 0.5 : 1.0);           // non-uniform grid: half of the per-zone dt/dx (exact)
      for (int d = 0; d < 3; d++) c.dtx[d] = h->nu && d < g.dims ? h->dtxa[d] : c.dtp + d;
      TIMED (h, KC_CT_EMF, count (h, DISPATCH (h, launch_ct_emf) (c, h->stream)));
      TIMED (h, KC_CT_UPDATE, count (h, DISPATCH (h, launch_ct_update) (c, h->stream)));
      TIMED (h, KC_FINAL, count (h, DISPATCH (h, launch_ctu_half) (s, h->stream)));
    }
  }
  // emf of the corrector fluxes (cell-centred part from V^{n+1/2}), B^{n+1} = B^n + dt curl E, cell average,
  // ConsToPrim (ctu_step.c:646-700)
  for (int nv = 0; nv < NVS; nv++) c.V[nv] = h->V[1][nv];
  for (int d = 0; d < 3; d++){ c.Bs_in[d] = h->Bs[0][d]; c.Bs0[d] = h->Bs[0][d]; c.Bs_out[d] = h->Bs[0][d]; }
  c.avg = h->cfg.emf_average; c.ext = 0; c.combine = 0; c.dtp = h->dtdev;
  c.gs = h->nu; c.dts = 1.0;
  for (int d = 0; d < 3; d++) c.dtx[d] = h->nu && d < g.dims ? h->dtxa[d] : c.dtp + d;
  TIMED (h, KC_CT_EMF, count (h, DISPATCH (h, launch_ct_emf) (c, h->stream)));
  TIMED (h, KC_CT_UPDATE, count (h, DISPATCH (h, launch_ct_update) (c, h->stream)));

  return launch_final_boxes (h, f, part);
}

// ---------------------------------------------------------------------------
int pluto_gpu_step_begin (PlutoGpu *h)
{
  CU (cudaSetDevice (h->cfg.device));
  CU (cudaMemsetAsync (h->red, 0, RED_N*sizeof (unsigned long long), h->stream));
  return 0;
}

static int stage_in_buf (const PlutoGpu *h, int stage) { return stage_plan (h, stage).in; }

int pluto_gpu_boundary_dim (PlutoGpu *h, int stage, int dim)
{
  CU (cudaSetDevice (h->cfg.device));
  if (stage < 1 || stage > h->nstages) return fail ("stage %d out of range", stage);
  return boundary_dim (h, stage_in_buf (h, stage), dim);
}

int pluto_gpu_stage (PlutoGpu *h, int stage, double dt)
{
  CU (cudaSetDevice (h->cfg.device));
  if (stage < 1 || stage > h->nstages) return fail ("stage %d out of range", stage);
  if (stage == 1 && dt >= 0.0 && set_dt (h, dt)) return 1;      // one dt per step (dt < 0: the device's own)
  return run_stage (h, stage);
}

int pluto_gpu_stage_shell (PlutoGpu *h, int stage, double dt)
{
  CU (cudaSetDevice (h->cfg.device));
  if (stage < 1 || stage > h->nstages) return fail ("stage %d out of range", stage);
  if (stage == 1 && dt >= 0.0 && set_dt (h, dt)) return 1;
  return run_stage (h, stage, PART_SHELL);
}

int pluto_gpu_stage_interior (PlutoGpu *h, int stage)
{
  CU (cudaSetDevice (h->cfg.device));
  if (stage < 1 || stage > h->nstages) return fail ("stage %d out of range", stage);
  return run_stage (h, stage, PART_INTERIOR);
}

int pluto_gpu_step_end (PlutoGpu *h, PlutoGpuStepInfo *info)
{
  CU (cudaSetDevice (h->cfg.device));
  CU (cudaMemcpyAsync (h->red_host, h->red, RED_N*sizeof (unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  CU (cudaStreamSynchronize (h->stream));
  if (h->timing) tcollect (h);
  double cd, mach;
  memcpy (&cd, &h->red_host[RED_CDT], sizeof (double));
  memcpy (&mach, &h->red_host[RED_MACH], sizeof (double));
  if (info){
    info->inv_dt_hyp = h->ctu ? cd : cd/(double)h->g.dims;     // update_stage.c:308-312; ctu_step.c:416-419
    info->max_mach = mach;
    info->floor_events = (int)h->red_host[RED_FLOOR];
    info->nan_events = (int)h->red_host[RED_NAN];
  }
  if (h->red_host[RED_ROEFAIL])
    return fail ("Roe_Solver: a2 < 0 at %llu interfaces (the reference aborts, roe.c:300-306)",
                 h->red_host[RED_ROEFAIL]);
  return 0;
}

static int enqueue_step (PlutoGpu *h)
{
  CU (cudaMemsetAsync (h->red, 0, RED_N*sizeof (unsigned long long), h->stream));
  for (int stage = 1; stage <= h->nstages; stage++){
    const int in = stage_in_buf (h, stage);
    for (int d = 0; d < h->g.dims; d++) if (boundary_dim (h, in, d)) return 1;
    if (run_stage (h, stage)) return 1;
  }
  return 0;
}

int pluto_gpu_advance (PlutoGpu *h, double dt, PlutoGpuStepInfo *info)
{
  CU (cudaSetDevice (h->cfg.device));
  if (set_dt (h, dt)) return 1;
  // The ~45-130 launches of a step are captured once into a CUDA graph (dt lives
  // in device memory, so the graph is replayed unchanged); the first step and
  // steps with per-kernel timing enabled are launched directly.
  if (h->use_graph && !h->timing && h->steps_done >= 1){
    if (!h->graph){
      cudaGraph_t gr = NULL;
      const long long l0 = h->launches;
      CU (cudaStreamBeginCapture (h->stream, cudaStreamCaptureModeThreadLocal));
      const int rc = enqueue_step (h);
      cudaError_t e = cudaStreamEndCapture (h->stream, &gr);
      if (rc || e != cudaSuccess){ if (gr) cudaGraphDestroy (gr); return rc ? 1 : fail ("graph capture: %s", cudaGetErrorString (e)); }
      h->graph_launches = h->launches - l0;
      h->launches = l0;
      e = cudaGraphInstantiate (&h->graph, gr, 0);
      cudaGraphDestroy (gr);
      if (e != cudaSuccess) return fail ("cudaGraphInstantiate: %s", cudaGetErrorString (e));
    }
    CU (cudaGraphLaunch (h->graph, h->stream));
    h->launches += h->graph_launches;
  }else{
    if (enqueue_step (h)) return 1;
  }
  h->steps_done++;
  return pluto_gpu_step_end (h, info);
}

int pluto_gpu_boundary (PlutoGpu *h)
{
  CU (cudaSetDevice (h->cfg.device));
  for (int d = 0; d < h->g.dims; d++) if (boundary_dim (h, 0, d)) return 1;
  CU (cudaStreamSynchronize (h->stream));
  return 0;
}

// The reference allocates Data once (Src/initialize.c:444-500) with malloc: page-lock
// the four blocks the first time they are seen so that every later step moves them
// by DMA at full PCIe rate instead of through the driver's pageable staging copy.
static void pin_once (PlutoGpu *h, const void *p, size_t bytes)
{
  if (!p || h->npinned >= 8) return;
  for (int q = 0; q < h->npinned; q++) if (h->pinned[q] == p) return;
  cudaPointerAttributes at;
  int ours = 0;
  if (cudaPointerGetAttributes (&at, p) == cudaSuccess && at.type == cudaMemoryTypeUnregistered){
    if (cudaHostRegister ((void *)p, bytes, cudaHostRegisterDefault) == cudaSuccess) ours = 1;
    else cudaGetLastError ();
  }
  h->pinned[h->npinned] = p; h->pinned_by_us[h->npinned++] = ours;
}

int pluto_gpu_advance_data (PlutoGpu *h, double dt, double *Vc, double *s1, double *s2, double *s3,
                            PlutoGpuStepInfo *info)
{
  CU (cudaSetDevice (h->cfg.device));
  {
    HaloArgs a; memset (&a, 0, sizeof (a));
    long long so[4], sl[4];
    describe_layout (h, 1, 0, a, so, sl);
    double *hp[4] = {Vc, s1, s2, s3};
    for (int q = 0; q < 4; q++) if (sl[q] > 0) pin_once (h, hp[q], (size_t)sl[q]*sizeof (double));
  }
  if (pluto_gpu_upload_data (h, Vc, s1, s2, s3)) return 1;
  if (pluto_gpu_advance (h, dt, info)) return 1;
  return pluto_gpu_download_data (h, Vc, s1, s2, s3);
}

// ---------------------------------------------------------------------------
//  NextTimeStep on the device (Src/main.c:462-465, 532): one thread turns the CFL
//  reduction of the step just enqueued into the next dt, in the arithmetic of
//  pluto_gpu_next_dt, records the step in the history ring and stores dt/dx for the
//  kernels of the next step.  The host never has to wait for a step to enqueue the next.
// ---------------------------------------------------------------------------
__global__ void next_dt_kernel (const unsigned long long *red, double *dtdev, double *hist,
                                unsigned long long *hist_count, double dx0, double dx1, double dx2,
                                int dims, double cfl, double cfl_max_var)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double dt = dtdev[3];
  const double inv = __ddiv_rn (__longlong_as_double ((long long)red[RED_CDT]), (double)dims);   // update_stage.c:308-312
  double dtnext = __dmul_rn (__ddiv_rn (1.0, inv), cfl);
  const double cap = __dmul_rn (cfl_max_var, dt);
  if (!(dtnext <= cap)) dtnext = cap;
  const unsigned long long n = *hist_count;
  double *hrow = hist + (n % HIST_N)*HIST_W;
  hrow[0] = dt; hrow[1] = inv; hrow[2] = __longlong_as_double ((long long)red[RED_MACH]);
  hrow[3] = (double)red[RED_FLOOR]; hrow[4] = (double)red[RED_NAN]; hrow[5] = (double)red[RED_ROEFAIL];
  *hist_count = n + 1;
  dtdev[0] = __ddiv_rn (dtnext, dx0); dtdev[1] = __ddiv_rn (dtnext, dx1); dtdev[2] = __ddiv_rn (dtnext, dx2);
  dtdev[3] = dtnext;
  const double dth = __dmul_rn (0.5, dtnext);
  dtdev[4] = __ddiv_rn (dth, dx0); dtdev[5] = __ddiv_rn (dth, dx1); dtdev[6] = __ddiv_rn (dth, dx2);
  dtdev[7] = dth;
}

int pluto_gpu_set_dt (PlutoGpu *h, double dt)
{
  CU (cudaSetDevice (h->cfg.device));
  return set_dt (h, dt);
}

int pluto_gpu_reduction_slots (PlutoGpu *h, void **dev_ptr)
{
  *dev_ptr = (void *)h->red;       // [RED_CDT], [RED_MACH]: bit patterns of non-negative doubles
  return 0;
}

int pluto_gpu_next_dt_async (PlutoGpu *h, double cfl, double cfl_max_var)
{
  CU (cudaSetDevice (h->cfg.device));
  if (h->hist_enq - h->hist_read >= HIST_N) return fail ("more than %d steps enqueued without pluto_gpu_sync_results", HIST_N);
  next_dt_kernel<<<1, 32, 0, h->stream>>>(h->red, h->dtdev, h->hist, h->hist_count, h->g.dx[0], h->g.dx[1], h->g.dx[2],
                                          h->ctu ? 1 : h->g.dims, cfl, cfl_max_var);
  if (count (h, pg_launch_status ())) return 1;
  if (refresh_dtx (h)) return 1;
  h->hist_enq++;
  return 0;
}

int pluto_gpu_advance_async (PlutoGpu *h, double cfl, double cfl_max_var)
{
  CU (cudaSetDevice (h->cfg.device));
  if (h->use_graph && !h->timing && h->steps_done >= 1){
    if (!h->graph){
      cudaGraph_t gr = NULL;
      const long long l0 = h->launches;
      CU (cudaStreamBeginCapture (h->stream, cudaStreamCaptureModeThreadLocal));
      const int rc = enqueue_step (h);
      cudaError_t e = cudaStreamEndCapture (h->stream, &gr);
      if (rc || e != cudaSuccess){ if (gr) cudaGraphDestroy (gr); return rc ? 1 : fail ("graph capture: %s", cudaGetErrorString (e)); }
      h->graph_launches = h->launches - l0;
      h->launches = l0;
      e = cudaGraphInstantiate (&h->graph, gr, 0);
      cudaGraphDestroy (gr);
      if (e != cudaSuccess) return fail ("cudaGraphInstantiate: %s", cudaGetErrorString (e));
    }
    CU (cudaGraphLaunch (h->graph, h->stream));
    h->launches += h->graph_launches;
  }else{
    if (enqueue_step (h)) return 1;
  }
  h->steps_done++;
  return pluto_gpu_next_dt_async (h, cfl, cfl_max_var);
}

int pluto_gpu_sync_results (PlutoGpu *h, int max_steps, PlutoGpuStepInfo *infos, double *dts, int *n_out, double *dt_next)
{
  CU (cudaSetDevice (h->cfg.device));
  const long long n = h->hist_enq - h->hist_read;
  if (n > 0) CU (cudaMemcpyAsync (h->hist_host, h->hist, (size_t)HIST_W*HIST_N*sizeof (double), cudaMemcpyDeviceToHost, h->stream));
  CU (cudaMemcpyAsync (h->dthost, h->dtdev, 8*sizeof (double), cudaMemcpyDeviceToHost, h->stream));
  CU (cudaStreamSynchronize (h->stream));
  if (h->timing) tcollect (h);
  if (dt_next) *dt_next = h->dthost[3];
  int m = 0; unsigned long long roe = 0;
  for (long long q = 0; q < n; q++){
    const double *r = h->hist_host + ((h->hist_read + q) % HIST_N)*HIST_W;
    roe += (unsigned long long)r[5];
    if (m < max_steps){
      if (dts) dts[m] = r[0];
      if (infos){ infos[m].inv_dt_hyp = r[1]; infos[m].max_mach = r[2]; infos[m].floor_events = (int)r[3]; infos[m].nan_events = (int)r[4]; }
      m++;
    }
  }
  h->hist_read = h->hist_enq;
  if (n_out) *n_out = m;
  if (roe) return fail ("Roe_Solver: a2 < 0 at %llu interfaces (the reference aborts, roe.c:300-306)", roe);
  return 0;
}

// ---------------------------------------------------------------------------
//  output / restart in the reference's .dbl format, diagnostics
// ---------------------------------------------------------------------------
static size_t interior_doubles (const PlutoGpu *h, size_t seg[4])
{
  const Geom &g = h->g;
  const size_t n1 = g.n[0], n2 = g.n[1], n3 = (g.dims == 3 ? g.n[2] : 1);
  seg[0] = (size_t)NVS*n1*n2*n3;                      // 8 slots (dead ones keep their space)
  seg[1] = (n1 + 1)*n2*n3; seg[2] = n1*(n2 + 1)*n3; seg[3] = (g.dims == 3 ? n1*n2*(n3 + 1) : 0);
  return seg[0] + seg[1] + seg[2] + seg[3];
}

int pluto_gpu_write_dbl (PlutoGpu *h, const char *dir, int nfile, double t, double dt, long nstep)
{
  size_t seg[4];
  const size_t tot = interior_doubles (h, seg);
  double *buf = NULL;
  if (cudaMallocHost ((void **)&buf, tot*sizeof (double)) != cudaSuccess) return fail ("pinned staging of %zu bytes failed", tot*sizeof (double));
  double *p[4] = {buf, buf + seg[0], buf + seg[0] + seg[1], seg[3] ? buf + seg[0] + seg[1] + seg[2] : NULL};
  if (pluto_gpu_download_interior (h, p[0], p[1], p[2], p[3])){ cudaFreeHost (buf); return 1; }
  char path[1024];
  snprintf (path, sizeof (path), "%s/data.%04d.dbl", dir, nfile);
  FILE *f = fopen (path, "wb");
  if (!f){ cudaFreeHost (buf); return fail ("cannot open %s", path); }
  const size_t ncell = seg[0]/NVS;
  size_t nw = 0, want = 0;
  for (int nv = 0; nv < NVS; nv++){
    if (!live_var (h, nv)) continue;
    nw += fwrite (buf + (size_t)nv*ncell, sizeof (double), ncell, f); want += ncell;
  }
  for (int s = 1; s < 4; s++) if (seg[s]){ nw += fwrite (p[s], sizeof (double), seg[s], f); want += seg[s]; }
  fclose (f);
  cudaFreeHost (buf);
  if (nw != want) return fail ("short write to %s", path);
  // dbl.out (write_data.c:365-395)
  snprintf (path, sizeof (path), "%s/dbl.out", dir);
  // as the reference does: file 0 starts the list; file n goes on line n, after skipping the n lines before it, and
  // whatever followed (a run restarted from an earlier file, a re-run into the same directory) is overwritten -- the
  // reference's restart and pyPLUTO read line n as file n
  if (nfile == 0) f = fopen (path, "w");
  else{
    f = fopen (path, "r+");
    if (!f) f = fopen (path, "w");                    // no list yet (the reference would fail here)
    else{
      char sline[512];
      for (int q = 0; q < nfile; q++) if (!fgets (sline, sizeof (sline), f)) break;
      fseek (f, ftell (f), SEEK_SET);
    }
  }
  if (!f) return fail ("cannot open %s", path);
  fprintf (f, "%d %12.6e %12.6e %ld single_file little ", nfile, t, dt, nstep);
  fprintf (f, h->g.dims == 3 ? "rho vx1 vx2 vx3 Bx1 Bx2 Bx3 prs Bx1s Bx2s Bx3s \n" : "rho vx1 vx2 Bx1 Bx2 prs Bx1s Bx2s \n");
  const long end = ftell (f);
  fclose (f);
  if (end > 0 && truncate (path, end) != 0) return fail ("cannot truncate %s", path);   // drop the stale lines of an earlier run
  return 0;
}

// one line of a "<ext>.out" list in the reference's format and position (write_data.c:365-395)
static int write_out_line (PlutoGpu *h, const char *dir, const char *ext, int nfile, double t, double dt, long nstep, bool staggered)
{
  char path[1024];
  snprintf (path, sizeof (path), "%s/%s.out", dir, ext);
  FILE *f;
  if (nfile == 0) f = fopen (path, "w");
  else{
    f = fopen (path, "r+");
    if (!f) f = fopen (path, "w");
    else{
      char sline[512];
      for (int q = 0; q < nfile; q++) if (!fgets (sline, sizeof (sline), f)) break;
      fseek (f, ftell (f), SEEK_SET);
    }
  }
  if (!f) return fail ("cannot open %s", path);
  fprintf (f, "%d %12.6e %12.6e %ld single_file little ", nfile, t, dt, nstep);
  fprintf (f, h->g.dims == 3 ? "rho vx1 vx2 vx3 Bx1 Bx2 Bx3 prs " : "rho vx1 vx2 Bx1 Bx2 prs ");
  if (staggered) fprintf (f, h->g.dims == 3 ? "Bx1s Bx2s Bx3s " : "Bx1s Bx2s ");
  fprintf (f, "\n");
  const long end = ftell (f);
  fclose (f);
  if (end > 0 && truncate (path, end) != 0) return fail ("cannot truncate %s", path);
  return 0;
}

// interior primitives as floats in pinned host memory: converted on the device (half the bytes cross PCIe)
static int stage_floats (PlutoGpu *h, bool swap, float **host, size_t *nz_out, int *nlive_out)
{
  CU (cudaSetDevice (h->cfg.device));
  const Geom &g = h->g;
  const size_t nz = (size_t)g.n[0]*g.n[1]*(g.dims == 3 ? g.n[2] : 1);
  CvtArgs a; memset (&a, 0, sizeof (a));
  int nlive = 0;
  for (int nv = 0; nv < NVS; nv++){ a.V[nv] = h->V[0][nv]; a.live[nv] = live_var (h, nv) ? nlive++ : -1; }
  if ((size_t)nlive*nz*sizeof (float) > h->scratch_doubles*sizeof (double)) return fail ("internal: scratch too small");
  a.out = (float *)h->scratch;                          // the work arrays are free between steps
  a.g = g; a.swap = swap ? 1 : 0;
  if (count (h, pg_exact::launch_cvt_float (a, h->stream))) return 1;
  if (cudaMallocHost ((void **)host, (size_t)nlive*nz*sizeof (float)) != cudaSuccess) return fail ("pinned staging of %zu bytes failed", (size_t)nlive*nz*sizeof (float));
  cudaError_t ce = cudaMemcpyAsync (*host, a.out, (size_t)nlive*nz*sizeof (float), cudaMemcpyDeviceToHost, h->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize (h->stream);
  if (ce != cudaSuccess){ cudaFreeHost (*host); *host = NULL; return fail ("stage_floats: %s", cudaGetErrorString (ce)); }
  *nz_out = nz; *nlive_out = nlive;
  return 0;
}

// data.NNNN.flt + flt.out: single-precision cell-centred variables, single_file (write_data.c:178-206)
int pluto_gpu_write_flt (PlutoGpu *h, const char *dir, int nfile, double t, double dt, long nstep)
{
  float *buf = NULL; size_t nz = 0; int nlive = 0;
  if (stage_floats (h, false, &buf, &nz, &nlive)) return 1;
  char path[1024];
  snprintf (path, sizeof (path), "%s/data.%04d.flt", dir, nfile);
  FILE *f = fopen (path, "wb");
  if (!f){ cudaFreeHost (buf); return fail ("cannot open %s", path); }
  const size_t nw = fwrite (buf, sizeof (float), (size_t)nlive*nz, f);
  fclose (f);
  cudaFreeHost (buf);
  if (nw != (size_t)nlive*nz) return fail ("short write to %s", path);
  return write_out_line (h, dir, "flt", nfile, t, dt, nstep, false);
}

// data.NNNN.vtk + vtk.out: legacy VTK, BINARY, RECTILINEAR_GRID with the node coordinates xl (n+1 per direction; x3 may be
// NULL in 2-D), every cell-centred variable as a SCALARS block of big-endian floats (write_vtk.c:92-351, write_data.c:234-262;
// VTK_VECTOR_DUMP NO and VTK_TIME_INFO NO, the reference's defaults)
int pluto_gpu_write_vtk (PlutoGpu *h, const char *dir, int nfile, double t, double dt, long nstep,
                         const double *xl1, const double *xl2, const double *xl3)
{
  const Geom &g = h->g;
  float *buf = NULL; size_t nz = 0; int nlive = 0;
  if (stage_floats (h, true, &buf, &nz, &nlive)) return 1;
  char path[1024];
  snprintf (path, sizeof (path), "%s/data.%04d.vtk", dir, nfile);
  FILE *f = fopen (path, "wb");
  if (!f){ cudaFreeHost (buf); return fail ("cannot open %s", path); }
  const int np[3] = {g.n[0] + 1, g.n[1] + 1, g.dims == 3 ? g.n[2] + 1 : 1};
  const double *xl[3] = {xl1, xl2, xl3};
  fprintf (f, "# vtk DataFile Version 2.0\nPLUTO 4.3 VTK Data\nBINARY\nDATASET RECTILINEAR_GRID\n");
  fprintf (f, "DIMENSIONS %d %d %d\n", np[0], np[1], np[2]);
  const char *cname[3] = {"X_COORDINATES", "\nY_COORDINATES", "\nZ_COORDINATES"};
  for (int d = 0; d < 3; d++){
    fprintf (f, "%s %d float\n", cname[d], np[d]);
    for (int i = 0; i < np[d]; i++){
      const float x = (d < g.dims && xl[d]) ? (float)xl[d][i] : 0.0f;
      unsigned u; memcpy (&u, &x, 4);
      u = (u >> 24) | ((u >> 8) & 0xff00u) | ((u << 8) & 0xff0000u) | (u << 24);
      fwrite (&u, 4, 1, f);
    }
  }
  fprintf (f, "\nCELL_DATA %zu\n", nz);
  static const char *names[NVS] = {"rho", "vx1", "vx2", "vx3", "Bx1", "Bx2", "Bx3", "prs"};
  size_t nw = 0; int q = 0;
  for (int nv = 0; nv < NVS; nv++){
    if (!live_var (h, nv)) continue;
    fprintf (f, "\nSCALARS %s float\nLOOKUP_TABLE default\n", names[nv]);
    nw += fwrite (buf + (size_t)q*nz, sizeof (float), nz, f);
    q++;
  }
  fclose (f);
  cudaFreeHost (buf);
  if (nw != (size_t)nlive*nz) return fail ("short write to %s", path);
  return write_out_line (h, dir, "vtk", nfile, t, dt, nstep, false);
}

int pluto_gpu_read_dbl (PlutoGpu *h, const char *path)
{
  size_t seg[4];
  const size_t tot = interior_doubles (h, seg);
  double *buf = (double *)calloc (tot, sizeof (double));
  if (!buf) return fail ("out of host memory");
  double *p[4] = {buf, buf + seg[0], buf + seg[0] + seg[1], seg[3] ? buf + seg[0] + seg[1] + seg[2] : NULL};
  FILE *f = fopen (path, "rb");
  if (!f){ free (buf); return fail ("cannot open %s", path); }
  const size_t ncell = seg[0]/NVS;
  size_t nr = 0, want = 0;
  for (int nv = 0; nv < NVS; nv++){
    if (!live_var (h, nv)) continue;
    nr += fread (buf + (size_t)nv*ncell, sizeof (double), ncell, f); want += ncell;
  }
  for (int s = 1; s < 4; s++) if (seg[s]){ nr += fread (p[s], sizeof (double), seg[s], f); want += seg[s]; }
  fclose (f);
  int rc = 0;
  if (nr != want) rc = fail ("%s: expected %zu doubles, read %zu (grid or dimensions differ)", path, want, nr);
  else rc = pluto_gpu_upload_interior (h, p[0], p[1], p[2], p[3]);
  free (buf);
  return rc;
}

int pluto_gpu_analysis (PlutoGpu *h, double out[8])
{
  CU (cudaSetDevice (h->cfg.device));
  int nb = 592;                                         // 148 SMs x 4 blocks
  if ((size_t)nb*8 > h->scratch_doubles) nb = (int)(h->scratch_doubles/8);
  if (nb < 1) return fail ("internal: scratch too small");
  AnalysisArgs a; memset (&a, 0, sizeof (a));
  for (int nv = 0; nv < NVS; nv++) a.V[nv] = h->V[0][nv];
  for (int d = 0; d < 3; d++) a.Bs[d] = h->Bs[0][d];
  a.partial = h->scratch;                               // the work arrays are free between steps
  a.g = h->g; a.igmm1 = h->ph.igmm1;
  if (count (h, pg_exact::launch_analysis (a, nb, h->stream))) return 1;
  double *host = (double *)malloc ((size_t)nb*8*sizeof (double));
  if (!host) return fail ("out of host memory");
  cudaError_t ce = cudaMemcpyAsync (host, a.partial, (size_t)nb*8*sizeof (double), cudaMemcpyDeviceToHost, h->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize (h->stream);
  if (ce != cudaSuccess){ free (host); return fail ("pluto_gpu_analysis: %s", cudaGetErrorString (ce)); }
  double vol = 1.0;
  if (h->nu){ free (host); return fail ("pluto_gpu_analysis: volume integrals on a non-uniform grid are not available"); }
  for (int d = 0; d < h->g.dims; d++) vol *= h->g.dx[d];
  for (int q = 0; q < 8; q++) out[q] = 0.0;
  for (int b = 0; b < nb; b++){
    for (int q = 0; q < 7; q++) out[q] += host[(size_t)b*8 + q];
    if (host[(size_t)b*8 + 7] > out[7]) out[7] = host[(size_t)b*8 + 7];
  }
  for (int q = 0; q < 7; q++) out[q] *= vol;
  free (host);
  return 0;
}

double pluto_gpu_next_dt (double inv_dt_hyp, double cfl, double cfl_max_var, double dt)
{
  double dt_hyp = 1.0/inv_dt_hyp, dtnext;                      // main.c:462-465
  dt_hyp *= cfl;
  dtnext  = dt_hyp;
  if (!(dtnext <= cfl_max_var*dt)) dtnext = cfl_max_var*dt;    // main.c:532 (MIN)
  return dtnext;
}

// ---------------------------------------------------------------------------
//  halo exchange support.  For dimension `dim` every field's box spans the
//  FULL extent (ghosts included) of the other dimensions, so the sequential
//  x1 -> x2 -> x3 exchange fills edges and corners (al_decompose.c:218-229).
//  Cell-centred fields move ng layers; the staggered component normal to
//  `dim` moves ng+1 layers towards the high neighbour (the shared face is
//  owned by the low block, al_decompose.c:243-252) and ng the other way.
// ---------------------------------------------------------------------------
static void halo_describe (PlutoGpu *h, int buf, int dim, int hs, bool send, HaloArgs &a)
{
  const Geom &g = h->g;
  memset (&a, 0, sizeof (a));
  int nf = 0; long long off = 0;
  for (int fidx = 0; fidx < NVS + 3; fidx++){
    const bool stag = fidx >= NVS;
    const int s = fidx - NVS;
    if (!stag && !live_var (h, fidx)) continue;
    if (stag && s >= g.dims) continue;
    int lo[3] = {0, 0, 0}, hi[3] = {g.T[0] - 1, g.T[1] - 1, g.T[2] - 1};
    if (stag) lo[s] = -1;
    const bool normal = stag && s == dim;
    if (send){
      if (hs == 0){ lo[dim] = g.beg[dim]; hi[dim] = g.beg[dim] + g.ng - 1; }          // to the low neighbour
      else        { lo[dim] = g.end[dim] - g.ng + 1 - (normal ? 1 : 0); hi[dim] = g.end[dim]; }
    }else{
      if (hs == 0){ lo[dim] = (normal ? -1 : 0); hi[dim] = g.beg[dim] - 1; }           // my low ghosts
      else        { lo[dim] = g.end[dim] + 1; hi[dim] = g.T[dim] - 1; }
    }
    a.q[nf] = stag ? h->Bs[buf][s] : h->V[buf][fidx];
    for (int d = 0; d < 3; d++){ a.lo[nf][d] = lo[d]; a.hi[nf][d] = hi[d]; }
    a.offset[nf] = off; off += box_count (lo, hi); nf++;
  }
  a.nf = nf; a.g = g;
}

long long pluto_gpu_halo_doubles (const PlutoGpu *h, int dim)
{
  HaloArgs a;
  halo_describe ((PlutoGpu *)h, 0, dim, 1, true, a);          // the larger of the two directions
  return a.offset[a.nf - 1] + box_count (a.lo[a.nf - 1], a.hi[a.nf - 1]);
}

int pluto_gpu_halo_pack (PlutoGpu *h, int stage, int dim, double *send_lo, double *send_hi)
{
  CU (cudaSetDevice (h->cfg.device));
  const int buf = stage_in_buf (h, stage);
  HaloArgs a;
  if (send_lo){ halo_describe (h, buf, dim, 0, true, a); a.buf = send_lo; TIMED (h, KC_HALO, count (h, DISPATCH (h, launch_halo_pack) (a, h->stream))); }
  if (send_hi){ halo_describe (h, buf, dim, 1, true, a); a.buf = send_hi; TIMED (h, KC_HALO, count (h, DISPATCH (h, launch_halo_pack) (a, h->stream))); }
  return 0;
}

int pluto_gpu_halo_unpack (PlutoGpu *h, int stage, int dim, const double *recv_lo, const double *recv_hi)
{
  CU (cudaSetDevice (h->cfg.device));
  const int buf = stage_in_buf (h, stage);
  HaloArgs a;
  if (recv_lo){ halo_describe (h, buf, dim, 0, false, a); a.buf = (double *)recv_lo; TIMED (h, KC_HALO, count (h, DISPATCH (h, launch_halo_unpack) (a, h->stream))); }
  if (recv_hi){ halo_describe (h, buf, dim, 1, false, a); a.buf = (double *)recv_hi; TIMED (h, KC_HALO, count (h, DISPATCH (h, launch_halo_unpack) (a, h->stream))); }
  return 0;
}


// ---------------------------------------------------------------------------
//  FP64 pipe microbenchmark: 8 independent DFMA chains per thread, enough warps
//  to saturate the pipe.  Gives the MEASURED denominator of the FP64 roofline
//  (SURVEY.md 8d asks for it instead of the nominal 37.2 TFLOP/s).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_chain_kernel (double *out, double a, double b, int iters)
{
  double x0 = threadIdx.x*1e-9, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0;
  double x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
  for (int i = 0; i < iters; i++){
    x0 = fma (x0, a, b); x1 = fma (x1, a, b); x2 = fma (x2, a, b); x3 = fma (x3, a, b);
    x4 = fma (x4, a, b); x5 = fma (x5, a, b); x6 = fma (x6, a, b); x7 = fma (x7, a, b);
  }
  out[(size_t)blockIdx.x*blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int pluto_gpu_measure_fp64 (int device, double *tflops)
{
  CU (cudaSetDevice (device));
  cudaDeviceProp prop;
  CU (cudaGetDeviceProperties (&prop, device));
  const int nb = prop.multiProcessorCount*8, tpb = 256, iters = 20000;
  double *out;
  CU (cudaMalloc ((void **)&out, (size_t)nb*tpb*sizeof (double)));
  cudaEvent_t e0, e1;
  CU (cudaEventCreate (&e0)); CU (cudaEventCreate (&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++){
    CU (cudaEventRecord (e0, 0));
    dfma_chain_kernel<<<nb, tpb>>>(out, 0.999999, 1e-7, iters);
    CU (cudaEventRecord (e1, 0));
    CU (cudaEventSynchronize (e1));
    float ms = 0.f;
    CU (cudaEventElapsedTime (&ms, e0, e1));
    const double tf = 2.0*8.0*(double)iters*(double)nb*tpb/(ms*1e-3)/1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy (e0); cudaEventDestroy (e1); cudaFree (out);
  *tflops = best;
  return 0;
}

// ---------------------------------------------------------------------------
//  Several blocks -- one per GPU -- driven from ONE host thread: what the reference's single-threaded, non-re-entrant C
//  host needs to use the 8 GPUs of a box.  Replaces Src/Parallel for a host that is NOT an MPI program: the
//  domain is cut into grid[0] x grid[1] x grid[2] blocks (rank = c1 + p1 (c2 + p2 c3), as al_decompose.c numbers them), every
//  block lives on its own device and stream, and the ghost zones travel by peer stores: a block's pack launch writes into
//  its neighbours' receive buffers (cudaDeviceEnablePeerAccess; same process, no IPC), an event per block and exchange orders
//  pack -> unpack across the streams, a second one unpack -> next pack (the buffers are single).  The host arrays are the
//  reference's Data arrays of the WHOLE domain.
// ---------------------------------------------------------------------------
enum { PGM_MAX_BLOCKS = 16, PGM_MAX_NBR = 26 };
struct PlutoGpuMulti {
  int nb, dims, ng, grid[3], gn[3], ln[3];
  PlutoGpu *blk[PGM_MAX_BLOCKS];
  int nnbr[PGM_MAX_BLOCKS], nbr[PGM_MAX_BLOCKS][PGM_MAX_NBR], noff[PGM_MAX_BLOCKS][PGM_MAX_NBR][3];
  double *recv[PGM_MAX_BLOCKS][PGM_MAX_NBR];           // receive buffer of block b for its neighbour q (on b's device)
  cudaEvent_t ev_pack[PGM_MAX_BLOCKS], ev_unpack[PGM_MAX_BLOCKS];
  double *hV[PGM_MAX_BLOCKS], *hS[PGM_MAX_BLOCKS][3];  // pinned staging of a block's Data arrays
};

static void pgm_coords (const PlutoGpuMulti *m, int r, int c[3])
{ c[0] = r % m->grid[0]; c[1] = (r/m->grid[0]) % m->grid[1]; c[2] = r/(m->grid[0]*m->grid[1]); }

void pluto_gpu_multi_destroy (PlutoGpuMulti *m)
{
  if (!m) return;
  for (int b = 0; b < m->nb; b++){
    if (m->blk[b]) cudaSetDevice (m->blk[b]->cfg.device);
    for (int q = 0; q < PGM_MAX_NBR; q++) if (m->recv[b][q]) cudaFree (m->recv[b][q]);
    if (m->ev_pack[b]) cudaEventDestroy (m->ev_pack[b]);
    if (m->ev_unpack[b]) cudaEventDestroy (m->ev_unpack[b]);
    if (m->hV[b]) cudaFreeHost (m->hV[b]);
    for (int d = 0; d < 3; d++) if (m->hS[b][d]) cudaFreeHost (m->hS[b][d]);
    if (m->blk[b]) pluto_gpu_destroy (m->blk[b]);
  }
  free (m);
}

int pluto_gpu_multi_create (const PlutoGpuConfig *cfg, const int grid[3], const int *devices, PlutoGpuMulti **out)
{
  *out = NULL;
  const int dims = cfg->dims;
  const int nb = grid[0]*grid[1]*(dims == 3 ? grid[2] : 1);
  if (nb < 1 || nb > PGM_MAX_BLOCKS) return fail ("pluto_gpu_multi_create: %d blocks (1 .. %d)", nb, PGM_MAX_BLOCKS);
  if (dims == 2 && grid[2] != 1) return fail ("pluto_gpu_multi_create: grid[2] must be 1 in 2-D");
  for (int d = 0; d < dims; d++) if (cfg->n[d] % grid[d]) return fail ("pluto_gpu_multi_create: n[%d] = %d is not divisible by %d blocks", d, cfg->n[d], grid[d]);
  PlutoGpuMulti *m = (PlutoGpuMulti *)calloc (1, sizeof (PlutoGpuMulti));
  if (!m) return fail ("out of host memory");
  m->nb = nb; m->dims = dims;
  for (int d = 0; d < 3; d++){ m->grid[d] = (d < dims ? grid[d] : 1); m->gn[d] = (d < dims ? cfg->n[d] : 1); m->ln[d] = m->gn[d]/m->grid[d]; }
  int rc = 0;
  for (int b = 0; b < nb && !rc; b++){
    int c[3]; pgm_coords (m, b, c);
    PlutoGpuConfig bc = *cfg;
    bc.device = devices ? devices[b] : b;
    for (int d = 0; d < 3; d++) bc.n[d] = m->ln[d];
    // a side is SHARED where another block abuts (boundary.c:139), across a periodic boundary too
    for (int d = 0; d < dims; d++) if (m->grid[d] > 1){
      const bool per = cfg->bc[2*d] == PLUTO_GPU_BC_PERIODIC;
      if (c[d] > 0 || per) bc.bc[2*d] = PLUTO_GPU_BC_SHARED;
      if (c[d] < m->grid[d] - 1 || per) bc.bc[2*d + 1] = PLUTO_GPU_BC_SHARED;
    }
    rc = pluto_gpu_create (&bc, &m->blk[b]);
  }
  if (rc){ char msg[sizeof (g_err)]; snprintf (msg, sizeof (msg), "%s", g_err); pluto_gpu_multi_destroy (m); snprintf (g_err, sizeof (g_err), "%s", msg); return 1; }
  m->ng = m->blk[0]->g.ng;
  // peer access between every pair of distinct devices (stores into a neighbour's buffer, events across devices)
  for (int a = 0; a < nb; a++) for (int b = 0; b < nb; b++){
    const int da = m->blk[a]->cfg.device, db = m->blk[b]->cfg.device;
    if (da == db) continue;
#ifndef PG_EMU
    int can = 0;
    cudaDeviceCanAccessPeer (&can, da, db);
    if (!can){ pluto_gpu_multi_destroy (m); return fail ("device %d cannot access device %d directly", da, db); }
    cudaSetDevice (da);
    const cudaError_t e = cudaDeviceEnablePeerAccess (db, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled){ pluto_gpu_multi_destroy (m); return fail ("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString (e)); }
    cudaGetLastError ();
#endif
  }
  // neighbours of every block (faces, edges, corners), their receive buffers, events, staging
  for (int b = 0; b < nb; b++){
    int c[3]; pgm_coords (m, b, c);
    PlutoGpu *h = m->blk[b];
    cudaSetDevice (h->cfg.device);
    int n = 0;
    for (int o2 = (dims == 3 && m->grid[2] > 1 ? -1 : 0); o2 <= (dims == 3 && m->grid[2] > 1 ? 1 : 0); o2++)
    for (int o1 = (m->grid[1] > 1 ? -1 : 0); o1 <= (m->grid[1] > 1 ? 1 : 0); o1++)
    for (int o0 = (m->grid[0] > 1 ? -1 : 0); o0 <= (m->grid[0] > 1 ? 1 : 0); o0++){
      const int o[3] = {o0, o1, o2};
      if (!o0 && !o1 && !o2) continue;
      int cc[3]; bool ok = true;
      for (int d = 0; d < 3; d++){
        cc[d] = c[d] + o[d];
        if (cc[d] < 0 || cc[d] >= m->grid[d]){
          if (d < dims && cfg->bc[2*d] == PLUTO_GPU_BC_PERIODIC) cc[d] = (cc[d] + m->grid[d]) % m->grid[d];
          else ok = false;
        }
      }
      if (!ok) continue;
      m->nbr[b][n] = cc[0] + m->grid[0]*(cc[1] + m->grid[1]*cc[2]);
      for (int d = 0; d < 3; d++) m->noff[b][n][d] = o[d];
      const long long cnt = pluto_gpu_halo_nbr_doubles (h, o);
      if (cudaMalloc ((void **)&m->recv[b][n], (size_t)(cnt > 0 ? cnt : 1)*sizeof (double)) != cudaSuccess){ pluto_gpu_multi_destroy (m); return fail ("cudaMalloc of a halo buffer failed"); }
      n++;
    }
    m->nnbr[b] = n;
    cudaEventCreateWithFlags (&m->ev_pack[b], cudaEventDisableTiming);
    cudaEventCreateWithFlags (&m->ev_unpack[b], cudaEventDisableTiming);
    const Geom &g = h->g;
    const size_t T1 = g.T[0], T2 = g.T[1], T3 = g.T[2];
    const int nvar = (dims == 3 ? NVS : 6);
    if (cudaMallocHost ((void **)&m->hV[b], nvar*T1*T2*T3*sizeof (double)) != cudaSuccess ||
        cudaMallocHost ((void **)&m->hS[b][0], (T1 + 1)*T2*T3*sizeof (double)) != cudaSuccess ||
        cudaMallocHost ((void **)&m->hS[b][1], T1*(T2 + 1)*T3*sizeof (double)) != cudaSuccess ||
        (dims == 3 && cudaMallocHost ((void **)&m->hS[b][2], T1*T2*(T3 + 1)*sizeof (double)) != cudaSuccess)){
      pluto_gpu_multi_destroy (m); return fail ("pinned staging for block %d failed", b);
    }
  }
  // plans: block b sends to neighbour q by writing into q's receive buffer for the opposite offset
  for (int b = 0; b < nb; b++){
    double *send[PGM_MAX_NBR], *recv[PGM_MAX_NBR]; int offs[3*PGM_MAX_NBR];
    for (int q = 0; q < m->nnbr[b]; q++){
      const int p = m->nbr[b][q];
      int qp = -1;
      for (int t = 0; t < m->nnbr[p]; t++)
        if (m->nbr[p][t] == b && m->noff[p][t][0] == -m->noff[b][q][0] && m->noff[p][t][1] == -m->noff[b][q][1] && m->noff[p][t][2] == -m->noff[b][q][2]) qp = t;
      if (qp < 0){ pluto_gpu_multi_destroy (m); return fail ("internal: neighbour tables are not symmetric"); }
      send[q] = m->recv[p][qp]; recv[q] = m->recv[b][q];
      for (int d = 0; d < 3; d++) offs[3*q + d] = m->noff[b][q][d];
    }
    if (m->nnbr[b] && pluto_gpu_halo_plan (m->blk[b], m->nnbr[b], offs, send, recv)){ pluto_gpu_multi_destroy (m); return 1; }
  }
  *out = m;
  return 0;
}

// non-uniform grid / grid-dependent reconstruction weights of the WHOLE domain: every block takes its slice (its own zones and the
// ghost entries on either side are contiguous in the global arrays)
int pluto_gpu_multi_set_grid (PlutoGpuMulti *m, const double *dx1, const double *dx2, const double *dx3)
{
  const double *src[3] = {dx1, dx2, dx3};
  for (int b = 0; b < m->nb; b++){
    int c[3]; pgm_coords (m, b, c);
    const double *p[3] = {NULL, NULL, NULL};
    for (int d = 0; d < m->dims; d++){
      if (!src[d]) return fail ("pluto_gpu_multi_set_grid: NULL array for direction %d", d + 1);
      p[d] = src[d] + (size_t)c[d]*m->ln[d];
    }
    if (pluto_gpu_set_grid (m->blk[b], p[0], p[1], p[2])) return 1;
  }
  return 0;
}
int pluto_gpu_multi_set_plm_coeffs (PlutoGpuMulti *m, int dir, const double *cp, const double *cm, const double *wp, const double *wm,
                                    const double *dp, const double *dm)
{
  if (dir < 0 || dir >= m->dims) return fail ("pluto_gpu_multi_set_plm_coeffs: direction %d", dir);
  for (int b = 0; b < m->nb; b++){
    int c[3]; pgm_coords (m, b, c);
    const size_t off = (size_t)c[dir]*m->ln[dir];
    if (pluto_gpu_set_plm_coeffs (m->blk[b], dir, cp + off, cm + off, wp + off, wm + off, dp + off, dm + off)) return 1;
  }
  return 0;
}

int pluto_gpu_multi_set_ppm_coeffs (PlutoGpuMulti *m, int dir, const double *wm1, const double *w0, const double *w1, const double *w2)
{
  if (dir < 0 || dir >= m->dims) return fail ("pluto_gpu_multi_set_ppm_coeffs: direction %d", dir);
  for (int b = 0; b < m->nb; b++){
    int c[3]; pgm_coords (m, b, c);
    const size_t off = (size_t)c[dir]*m->ln[dir];
    if (pluto_gpu_set_ppm_coeffs (m->blk[b], dir, wm1 + off, w0 + off, w1 + off, w2 + off)) return 1;
  }
  return 0;
}

int pluto_gpu_device_count (void)
{
  int n = 0;
  return cudaGetDeviceCount (&n) == cudaSuccess ? n : 0;
}

int pluto_gpu_multi_nghost (const PlutoGpuMulti *m) { return m->ng; }
int pluto_gpu_multi_nblocks (const PlutoGpuMulti *m) { return m->nb; }

// copy between the global Data arrays (host) and one block's staging arrays: rows along x1
static void pgm_rows (const PlutoGpuMulti *m, int b, bool to_block, int stag, double *glob, double *loc, bool interior_only)
{
  int c[3]; pgm_coords (m, b, c);
  const int ng = m->ng, dims = m->dims;
  int Tg[3], Tl[3], off[3];
  for (int d = 0; d < 3; d++){
    Tg[d] = (d < dims ? m->gn[d] + 2*ng : 1); Tl[d] = (d < dims ? m->ln[d] + 2*ng : 1); off[d] = c[d]*m->ln[d];
  }
  const int eg[3] = {Tg[0] + (stag == 0), Tg[1] + (stag == 1), Tg[2] + (stag == 2)};
  const int el[3] = {Tl[0] + (stag == 0), Tl[1] + (stag == 1), Tl[2] + (stag == 2)};
  int lo[3] = {0, 0, 0}, hi[3] = {el[0] - 1, el[1] - 1, el[2] - 1};
  if (interior_only) for (int d = 0; d < dims; d++){
    // interior zones; a staggered component keeps the face below its first zone (array position = face index + 1)
    lo[d] = ng; hi[d] = ng + m->ln[d] - 1 + (stag == d ? 1 : 0);
  }
  const size_t nrow = (size_t)(hi[0] - lo[0] + 1);
  for (int k = lo[2]; k <= hi[2]; k++) for (int j = lo[1]; j <= hi[1]; j++){
    double *pl = loc + ((size_t)k*el[1] + j)*el[0] + lo[0];
    double *pg = glob + ((size_t)(k + off[2])*eg[1] + (j + off[1]))*eg[0] + lo[0] + off[0];
    if (to_block) memcpy (pl, pg, nrow*sizeof (double)); else memcpy (pg, pl, nrow*sizeof (double));
  }
}

// BODY_FORCE with several blocks: the arrays of the whole domain (layouts of pluto_gpu_set_body_force / _potential), cut into
// the blocks' pieces -- ghost zones included, which overlap the neighbours' interiors exactly as the state's do
static int pgm_sliced (PlutoGpuMulti *m, int narr, const double *const *glob, const int *stag, const char *what,
                       int (*set) (PlutoGpu *, const double *const *))
{
  size_t totl = 1;
  for (int d = 0; d < m->dims; d++) totl *= (size_t)(m->ln[d] + 2*m->ng + 1);
  double *tmp = (double *)malloc ((size_t)narr*totl*sizeof (double));
  if (!tmp) return fail ("out of host memory");
  int rc = 0;
  for (int b = 0; b < m->nb && !rc; b++){
    const double *loc[4] = {NULL, NULL, NULL, NULL};
    for (int q = 0; q < narr && !rc; q++){
      if (!glob[q]){ rc = fail ("%s: NULL array %d", what, q); break; }
      pgm_rows (m, b, true, stag[q], (double *)glob[q], tmp + (size_t)q*totl, false);
      loc[q] = tmp + (size_t)q*totl;
    }
    if (!rc) rc = set (m->blk[b], loc);
  }
  free (tmp);
  return rc;
}
int pluto_gpu_multi_set_body_force (PlutoGpuMulti *m, const double *g1, const double *g2, const double *g3)
{
  const double *glob[3] = {g1, g2, g3};
  const int stag[3] = {-1, -1, -1};
  return pgm_sliced (m, m->dims, glob, stag, "pluto_gpu_multi_set_body_force",
                     [] (PlutoGpu *h, const double *const *l){ return pluto_gpu_set_body_force (h, l[0], l[1], l[2]); });
}
int pluto_gpu_multi_set_body_potential (PlutoGpuMulti *m, const double *phic, const double *pf1, const double *pf2, const double *pf3)
{
  const double *glob[4] = {phic, pf1, pf2, pf3};
  const int stag[4] = {-1, 0, 1, 2};
  return pgm_sliced (m, 1 + m->dims, glob, stag, "pluto_gpu_multi_set_body_potential",
                     [] (PlutoGpu *h, const double *const *l){ return pluto_gpu_set_body_potential (h, l[0], l[1], l[2], l[3]); });
}

int pluto_gpu_multi_upload_data (PlutoGpuMulti *m, const double *Vc, const double *s1, const double *s2, const double *s3)
{
  const int nvar = (m->dims == 3 ? NVS : 6), ng = m->ng;
  size_t totg = 1, totl = 1;
  for (int d = 0; d < m->dims; d++){ totg *= (size_t)(m->gn[d] + 2*ng); totl *= (size_t)(m->ln[d] + 2*ng); }
  const double *S[3] = {s1, s2, s3};
  for (int b = 0; b < m->nb; b++){
    for (int nv = 0; nv < nvar; nv++) pgm_rows (m, b, true, -1, (double *)Vc + (size_t)nv*totg, m->hV[b] + (size_t)nv*totl, false);
    for (int d = 0; d < m->dims; d++) pgm_rows (m, b, true, d, (double *)S[d], m->hS[b][d], false);
    if (pluto_gpu_upload_data (m->blk[b], m->hV[b], m->hS[b][0], m->hS[b][1], m->dims == 3 ? m->hS[b][2] : NULL)) return 1;
  }
  return 0;
}

int pluto_gpu_multi_download_data (PlutoGpuMulti *m, double *Vc, double *s1, double *s2, double *s3)
{
  const int nvar = (m->dims == 3 ? NVS : 6), ng = m->ng;
  size_t totg = 1, totl = 1;
  for (int d = 0; d < m->dims; d++){ totg *= (size_t)(m->gn[d] + 2*ng); totl *= (size_t)(m->ln[d] + 2*ng); }
  double *S[3] = {s1, s2, s3};
  for (int b = 0; b < m->nb; b++){
    if (pluto_gpu_download_data (m->blk[b], m->hV[b], m->hS[b][0], m->hS[b][1], m->dims == 3 ? m->hS[b][2] : NULL)) return 1;
    for (int nv = 0; nv < nvar; nv++) pgm_rows (m, b, false, -1, Vc + (size_t)nv*totg, m->hV[b] + (size_t)nv*totl, true);
    for (int d = 0; d < m->dims; d++) pgm_rows (m, b, false, d, S[d], m->hS[b][d], true);
  }
  return 0;
}

int pluto_gpu_multi_advance (PlutoGpuMulti *m, double dt, PlutoGpuStepInfo *info)
{
  const int nst = m->blk[0]->nstages;
  for (int b = 0; b < m->nb; b++) if (pluto_gpu_step_begin (m->blk[b])) return 1;
  for (int stage = 1; stage <= nst; stage++){
    for (int b = 0; b < m->nb; b++){
      PlutoGpu *h = m->blk[b];
      if (!m->nnbr[b]) continue;
      CU (cudaSetDevice (h->cfg.device));
      // the neighbours have emptied the buffers this block is about to fill (previous exchange)
      for (int q = 0; q < m->nnbr[b]; q++) CU (cudaStreamWaitEvent (h->stream, m->ev_unpack[m->nbr[b][q]], 0));
      if (pluto_gpu_halo_pack_all (h, stage)) return 1;
      CU (cudaEventRecord (m->ev_pack[b], h->stream));
    }
    for (int b = 0; b < m->nb; b++){
      PlutoGpu *h = m->blk[b];
      CU (cudaSetDevice (h->cfg.device));
      if (m->nnbr[b]){
        for (int q = 0; q < m->nnbr[b]; q++) CU (cudaStreamWaitEvent (h->stream, m->ev_pack[m->nbr[b][q]], 0));
        if (pluto_gpu_halo_unpack_all (h, stage)) return 1;
        CU (cudaEventRecord (m->ev_unpack[b], h->stream));
      }
      for (int d = 0; d < m->dims; d++) if (pluto_gpu_boundary_dim (h, stage, d)) return 1;
      if (pluto_gpu_stage (h, stage, dt)) return 1;
    }
  }
  PlutoGpuStepInfo tot; memset (&tot, 0, sizeof (tot));
  int rc = 0;
  for (int b = 0; b < m->nb; b++){            // MPI_Allreduce(MAX) of invDt_hyp and g_maxMach (main.c:195-199, 415)
    PlutoGpuStepInfo i; memset (&i, 0, sizeof (i));
    if (pluto_gpu_step_end (m->blk[b], &i)) rc = 1;
    if (i.inv_dt_hyp > tot.inv_dt_hyp) tot.inv_dt_hyp = i.inv_dt_hyp;
    if (i.max_mach > tot.max_mach) tot.max_mach = i.max_mach;
    tot.floor_events += i.floor_events; tot.nan_events += i.nan_events;
  }
  if (info) *info = tot;
  return rc;
}

int pluto_gpu_multi_advance_data (PlutoGpuMulti *m, double dt, double *Vc, double *s1, double *s2, double *s3, PlutoGpuStepInfo *info)
{
  if (pluto_gpu_multi_upload_data (m, Vc, s1, s2, s3)) return 1;
  if (pluto_gpu_multi_advance (m, dt, info)) return 1;
  return pluto_gpu_multi_download_data (m, Vc, s1, s2, s3);
}

namespace pg_fast { int launch_arith_selftest (unsigned long long seed, int nblocks, int n, unsigned long long *bad, cudaStream_t s); }

// The FAST Roe kernels use a branch-free correctly rounded division / reciprocal / square root (mhd_device.cuh): compare them
// with div.rn.f64 / sqrt.rn.f64 on `samples` operand pairs; mismatches[0..2] = quotient, reciprocal, root.
extern "C" int pluto_gpu_selftest_arith (int device, long long samples, unsigned long long seed, unsigned long long mismatches[3])
{
  CU (cudaSetDevice (device));
  unsigned long long *bad;
  CU (cudaMalloc ((void **)&bad, 3*sizeof (unsigned long long)));
  CU (cudaMemset (bad, 0, 3*sizeof (unsigned long long)));
  const int nblocks = 148*8, per = (int)((samples + (long long)nblocks*256 - 1)/((long long)nblocks*256));
  if (pg_fast::launch_arith_selftest (seed, nblocks, per < 1 ? 1 : per, bad, 0) < 0){ cudaFree (bad); return fail ("selftest launch failed"); }
  cudaError_t ce = cudaMemcpy (mismatches, bad, 3*sizeof (unsigned long long), cudaMemcpyDeviceToHost);
  cudaFree (bad);
  if (ce != cudaSuccess) return fail ("pluto_gpu_selftest_arith: %s", cudaGetErrorString (ce));
  return 0;
}

// ---------------------------------------------------------------------------
//  All-neighbour exchange: instead of three sequential per-dimension swaps, every
//  face, edge and corner neighbour (up to 26 in 3-D) gets its own buffer and all
//  transfers of a stage are independent -> ONE pack launch, ONE communication
//  group, ONE unpack launch per stage.  Boxes along a dimension with offset o:
//    o = -1: send [beg, beg+ng-1]        receive [0 (-1 for the normal staggered comp), beg-1]
//    o = +1: send [end-ng+1 (-1), end]   receive [end+1, T-1]
//    o =  0: [beg (-1 for a component staggered in that dimension whose low side is a physical boundary), end]
//  i.e. the same layers AL_Exchange_dim moves (al_decompose.c:218-252), with the
//  edge/corner pieces addressed directly instead of relayed through ghost zones.
// ---------------------------------------------------------------------------
static void nbr_box (const PlutoGpu *h, const int off[3], int stag, bool send, int lo[3], int hi[3])
{
  const Geom &g = h->g;
  for (int d = 0; d < 3; d++){
    const int st = (stag == d);
    if (d >= g.dims){ lo[d] = hi[d] = 0; continue; }
    // o = 0: the face beg-1 of a component staggered in d belongs to the low neighbour when that side is SHARED;
    // it then arrives with the o[d] = -1 piece of the diagonal neighbour (the owner's value), not with this one --
    // two senders would otherwise write it, with different bits wherever the blocks' own copies of a shared face
    // are not identical (initial conditions evaluated at y = 0 and y = 2 pi)
    if (off[d] == 0){ lo[d] = g.beg[d] - (st && h->cfg.bc[2*d] != PLUTO_GPU_BC_SHARED ? 1 : 0); hi[d] = g.end[d]; }
    else if (send){
      if (off[d] < 0){ lo[d] = g.beg[d]; hi[d] = g.beg[d] + g.ng - 1; }
      else           { lo[d] = g.end[d] - g.ng + 1 - st; hi[d] = g.end[d]; }
    }else{
      if (off[d] < 0){ lo[d] = -st; hi[d] = g.beg[d] - 1; }
      else           { lo[d] = g.end[d] + 1; hi[d] = g.T[d] - 1; }
    }
  }
}

long long pluto_gpu_halo_nbr_doubles (const PlutoGpu *h, const int off[3])
{
  long long tot = 0, tot_r = 0;
  for (int f = 0; f < NVS + 3; f++){
    const int stag = f >= NVS ? f - NVS : -1;
    if (stag < 0 && !live_var (h, f)) continue;
    if (stag >= h->g.dims) continue;
    int lo[3], hi[3];
    nbr_box (h, off, stag, true, lo, hi);  tot   += box_count (lo, hi);
    nbr_box (h, off, stag, false, lo, hi); tot_r += box_count (lo, hi);
  }
  return tot > tot_r ? tot : tot_r;
}

static int halo_plan_buffers (PlutoGpu *h, int b0, int b1, int n_nbr, const int *offsets, double *const *send_bufs,
                              double *const *recv_bufs);

int pluto_gpu_halo_plan (PlutoGpu *h, int n_nbr, const int *offsets, double *const *send_bufs,
                         double *const *recv_bufs)
{ return halo_plan_buffers (h, 0, h->nbuf, n_nbr, offsets, send_bufs, recv_bufs); }

// the same for ONE state buffer (= one stage of the step): lets a host give every stage its own buffers, e.g. receive
// areas in a peer GPU's memory that alternate from stage to stage (pluto_gpu_ipc_*, pluto_b200/parallel.py PeerExchanger)
int pluto_gpu_halo_plan_stage (PlutoGpu *h, int stage, int n_nbr, const int *offsets, double *const *send_bufs,
                               double *const *recv_bufs)
{
  const int b = stage_in_buf (h, stage);
  return halo_plan_buffers (h, b, b + 1, n_nbr, offsets, send_bufs, recv_bufs);
}

static int halo_plan_buffers (PlutoGpu *h, int b0, int b1, int n_nbr, const int *offsets, double *const *send_bufs,
                              double *const *recv_bufs)
{
  CU (cudaSetDevice (h->cfg.device));
  for (int b = b0; b < b1; b++) for (int dirn = 0; dirn < 2; dirn++){
    HaloEntry *tab = (HaloEntry *)calloc ((size_t)n_nbr*(NVS + 3) + 1, sizeof (HaloEntry));
    int ne = 0; long long mx = 0;
    for (int q = 0; q < n_nbr; q++){
      const int *off = offsets + 3*q;
      double *buf = dirn == 0 ? send_bufs[q] : recv_bufs[q];
      long long pos = 0;
      for (int f = 0; f < NVS + 3; f++){
        const int stag = f >= NVS ? f - NVS : -1;
        if (stag < 0 && !live_var (h, f)) continue;
        if (stag >= h->g.dims) continue;
        int lo[3], hi[3];
        nbr_box (h, off, stag, dirn == 0, lo, hi);
        HaloEntry &e = tab[ne++];
        e.field = stag >= 0 ? h->Bs[b][stag] : h->V[b][f];
        e.buf = buf + pos;
        for (int d = 0; d < 3; d++){ e.lo[d] = lo[d]; e.n[d] = hi[d] - lo[d] + 1; }
        e.count = box_count (lo, hi);
        pos += e.count;
        if (e.count > mx) mx = e.count;
      }
    }
    if (h->halo_tab[b][dirn]) cudaFree (h->halo_tab[b][dirn]);
    h->halo_tab[b][dirn] = NULL;
    if (ne > 0){
      cudaError_t ce = cudaMalloc ((void **)&h->halo_tab[b][dirn], (size_t)ne*sizeof (HaloEntry));
      if (ce == cudaSuccess) ce = cudaMemcpy (h->halo_tab[b][dirn], tab, (size_t)ne*sizeof (HaloEntry), cudaMemcpyHostToDevice);
      if (ce != cudaSuccess){ free (tab); return fail ("pluto_gpu_halo_plan: %s", cudaGetErrorString (ce)); }
    }
    h->halo_n[b][dirn] = ne; h->halo_max[b][dirn] = mx;
    free (tab);
  }
  return 0;
}

int pluto_gpu_halo_pack_all (PlutoGpu *h, int stage)
{
  CU (cudaSetDevice (h->cfg.device));
  const int b = stage_in_buf (h, stage);
  TIMED (h, KC_HALO, count (h, DISPATCH (h, launch_halo_table) (h->halo_tab[b][0], h->halo_n[b][0], h->halo_max[b][0], h->g, true, h->stream)));
  return 0;
}

int pluto_gpu_halo_pack_all_on (PlutoGpu *h, int stage, void *stream)
{
  CU (cudaSetDevice (h->cfg.device));
  const int b = stage_in_buf (h, stage);
  return count (h, DISPATCH (h, launch_halo_table) (h->halo_tab[b][0], h->halo_n[b][0], h->halo_max[b][0], h->g, true, (cudaStream_t)stream));
}

int pluto_gpu_halo_unpack_all (PlutoGpu *h, int stage)
{
  CU (cudaSetDevice (h->cfg.device));
  const int b = stage_in_buf (h, stage);
  TIMED (h, KC_HALO, count (h, DISPATCH (h, launch_halo_table) (h->halo_tab[b][1], h->halo_n[b][1], h->halo_max[b][1], h->g, false, h->stream)));
  return 0;
}

// ---------------------------------------------------------------------------
//  Ghost zones written straight into the neighbour GPU's memory (NVLink peer stores), no communication library on the data
//  path: a rank allocates its receive arena with pluto_gpu_ipc_alloc and publishes the 64-byte IPC handle; a neighbour
//  (another process on the same node) maps it with pluto_gpu_ipc_open and hands the mapped addresses to
//  pluto_gpu_halo_plan_stage as its SEND buffers -- the pack launch then IS the transfer.  Arrival is signalled by a counter
//  per (receiver, sender) pair in the same arena: pluto_gpu_halo_signal (after the pack launch, same stream) stores the
//  exchange number into the peers' counters, pluto_gpu_halo_wait (before the unpack launch) spins until every counter of
//  this rank has reached it.  All of it is stream-ordered device work: nothing waits on the host.
//  Replaces AL_Exchange_dim's MPI_Sendrecv (Src/Parallel/al_exchange_dim.c:58-88).
// ---------------------------------------------------------------------------
struct PeerFlags { unsigned long long *p[32]; };

__global__ void halo_signal_kernel (PeerFlags f, int n, unsigned long long value)
{
  const int q = threadIdx.x;
  if (q < n){
    __threadfence_system ();                      // the pack launch before this one has completed; order the counter after it
    *(volatile unsigned long long *)f.p[q] = value;
    __threadfence_system ();
  }
}

__global__ void halo_wait_kernel (const unsigned long long *flags, int n, unsigned long long value, unsigned long long *red)
{
  const int q = threadIdx.x;
  if (q < n){
    const volatile unsigned long long *p = flags + q;
    const long long t0 = clock64 ();
    while (*p < value){
      if (clock64 () - t0 > 20000000000LL){       // ~10 s: a neighbour died -- report instead of hanging the GPU
        atomicAdd (red + RED_NAN, 1ull);
        break;
      }
    }
    __threadfence_system ();
  }
}

int pluto_gpu_ipc_alloc (int device, size_t bytes, void **ptr, unsigned char handle[64])
{
#ifdef PG_EMU
  (void)device; (void)bytes; (void)ptr; (void)handle;
  return fail ("pluto_gpu_ipc_alloc: not available under the kernel interpreter");
#else
  CU (cudaSetDevice (device));
  static_assert (sizeof (cudaIpcMemHandle_t) == 64, "IPC handle size");
  CU (cudaMalloc (ptr, bytes));
  CU (cudaMemset (*ptr, 0, bytes));
  cudaIpcMemHandle_t hd;
  CU (cudaIpcGetMemHandle (&hd, *ptr));
  memcpy (handle, &hd, 64);
  return 0;
#endif
}

int pluto_gpu_ipc_open (int device, const unsigned char handle[64], void **ptr)
{
#ifdef PG_EMU
  (void)device; (void)handle; (void)ptr;
  return fail ("pluto_gpu_ipc_open: not available under the kernel interpreter");
#else
  CU (cudaSetDevice (device));
  cudaIpcMemHandle_t hd;
  memcpy (&hd, handle, 64);
  CU (cudaIpcOpenMemHandle (ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  return 0;
#endif
}

int pluto_gpu_ipc_close (void *ptr)
{
#ifndef PG_EMU
  if (ptr) CU (cudaIpcCloseMemHandle (ptr));
#endif
  return 0;
}

int pluto_gpu_ipc_free (void *ptr)
{
  if (ptr) CU (cudaFree (ptr));
  return 0;
}

int pluto_gpu_halo_signal (PlutoGpu *h, void *stream, int n, unsigned long long *const *peer_counters, unsigned long long value)
{
  if (n < 0 || n > 32) return fail ("pluto_gpu_halo_signal: %d neighbours", n);
  if (n == 0) return 0;
  CU (cudaSetDevice (h->cfg.device));
  PeerFlags f; memset (&f, 0, sizeof (f));
  for (int q = 0; q < n; q++) f.p[q] = peer_counters[q];
  halo_signal_kernel<<<1, 32, 0, stream ? (cudaStream_t)stream : h->stream>>>(f, n, value);
  return count (h, pg_launch_status ());
}

int pluto_gpu_halo_wait (PlutoGpu *h, void *stream, int n, const unsigned long long *counters, unsigned long long value)
{
  if (n <= 0) return 0;
  CU (cudaSetDevice (h->cfg.device));
  halo_wait_kernel<<<1, 32, 0, stream ? (cudaStream_t)stream : h->stream>>>(counters, n, value, h->red);
  return count (h, pg_launch_status ());
}
