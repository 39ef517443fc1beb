// emu_runtime.cpp (TEST INFRASTRUCTURE ONLY) -- the lock-step interpreter behind
// tests/emu/include/cuda_runtime.h.  A launch runs its blocks one after the other; the
// threads of a block are fibres (ucontext) scheduled round-robin.  A fibre runs until it
// finishes or reaches a collective: warp collectives release a warp once all of its
// unfinished lanes have arrived, __syncthreads releases the block likewise.  Everything is
// sequential and deterministic; a collective that not all lanes reach is reported as a
// deadlock (on the GPU it would be undefined behaviour).
#include <ucontext.h>
#include <vector>
#include "include/cuda_runtime.h"

uint3 threadIdx, blockIdx;
dim3  blockDim, gridDim;

namespace pg_emu {

enum { RUN = 0, WAIT_WARP = 1, WAIT_BLOCK = 2, DONE = 3 };
static const size_t kStack = 512*1024;

struct Copy { double *dst; const double *src; };
struct Fibre { ucontext_t ctx; char *stack; int state; unsigned long long ncoll;
               std::vector<std::vector<Copy> > groups; std::vector<Copy> open; };
struct WarpBox { uint64_t slot[2][32]; unsigned mask[2]; unsigned long long gen[2]; };

static ucontext_t g_main;
static std::vector<Fibre> g_fib;
static std::vector<WarpBox> g_warp;
static std::vector<char> g_smem;
static int g_cur = -1;
static const std::function<void ()> *g_body = nullptr;

static void fibre_entry ()
{
  (*g_body) ();
  g_fib[g_cur].state = DONE;
  swapcontext (&g_fib[g_cur].ctx, &g_main);
}

static void yield_as (int state)
{
  Fibre &f = g_fib[g_cur];
  f.state = state;
  swapcontext (&f.ctx, &g_main);
}

int lane_id () { return g_cur & 31; }
void *dyn_smem () { return g_smem.data (); }

uint64_t exchange (uint64_t mine, int src_lane)
{
  const int t = g_cur, w = t >> 5, l = t & 31;
  Fibre &f = g_fib[t];
  const unsigned long long k = ++f.ncoll;
  const int par = (int)(k & 1);
  WarpBox &b = g_warp[w];
  if (b.gen[par] != k){ b.gen[par] = k; b.mask[par] = 0; }
  b.slot[par][l] = mine; b.mask[par] |= 1u << l;
  yield_as (WAIT_WARP);
  return g_warp[w].slot[par][src_lane];
}

unsigned ballot (int pred)
{
  const int t = g_cur, w = t >> 5;
  const int par = (int)((g_fib[t].ncoll + 1) & 1);
  exchange (pred ? 1 : 0, t & 31);
  const WarpBox &b = g_warp[w];
  unsigned m = 0;
  for (int l = 0; l < 32; l++) if ((b.mask[par] >> l & 1u) && b.slot[par][l]) m |= 1u << l;
  return m;
}

void barrier_block () { yield_as (WAIT_BLOCK); }

void async_copy8 (double *dst, const double *src) { Copy c = {dst, src}; g_fib[g_cur].open.push_back (c); }
void async_commit () { Fibre &f = g_fib[g_cur]; f.groups.push_back (f.open); f.open.clear (); }
void async_wait (int keep)
{
  Fibre &f = g_fib[g_cur];
  while ((int)f.groups.size () > keep){
    for (const Copy &c : f.groups.front ()) *c.dst = *c.src;
    f.groups.erase (f.groups.begin ());
  }
}

static void run_block (int nthr)
{
  const int nwarp = (nthr + 31)/32;
  if ((int)g_fib.size () < nthr){
    const size_t old = g_fib.size ();
    g_fib.resize (nthr);
    for (size_t q = old; q < g_fib.size (); q++) g_fib[q].stack = (char *)malloc (kStack);
  }
  g_warp.assign (nwarp, WarpBox ());
  for (int t = 0; t < nthr; t++){
    Fibre &f = g_fib[t];
    getcontext (&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = kStack; f.ctx.uc_link = &g_main;
    makecontext (&f.ctx, fibre_entry, 0);
    f.state = RUN; f.ncoll = 0; f.groups.clear (); f.open.clear ();
  }
  // PG_EMU_REVERSE=1 schedules the runnable threads (and the blocks) in descending order: a kernel whose result
  // depends on the order in which the lanes of a warp run between two collectives has a race
  static const bool reverse = getenv ("PG_EMU_REVERSE") != nullptr && getenv ("PG_EMU_REVERSE")[0] == '1';
  for (;;){
    bool progressed = false;
    for (int q = 0; q < nthr; q++){
      const int t = reverse ? nthr - 1 - q : q;
      if (g_fib[t].state != RUN) continue;
      {
      g_cur = t; threadIdx.x = (unsigned)t; threadIdx.y = threadIdx.z = 0;
      swapcontext (&g_main, &g_fib[t].ctx);
      progressed = true;
      }
    }
    int ndone = 0, nblock = 0;
    for (int t = 0; t < nthr; t++){ ndone += g_fib[t].state == DONE; nblock += g_fib[t].state == WAIT_BLOCK; }
    if (ndone == nthr) break;
    bool released = false;
    for (int w = 0; w < nwarp; w++){
      const int lo = w*32, hi = lo + 32 < nthr ? lo + 32 : nthr;
      int nw = 0, nother = 0;
      unsigned long long k = 0; bool same = true;
      for (int t = lo; t < hi; t++){
        if (g_fib[t].state == WAIT_WARP){ if (nw && g_fib[t].ncoll != k) same = false; k = g_fib[t].ncoll; nw++; }
        else if (g_fib[t].state != DONE) nother++;
      }
      if (nw && !nother){
        if (!same){ fprintf (stderr, "pg_emu: lanes of warp %d wait at different collectives\n", w); abort (); }
        for (int t = lo; t < hi; t++) if (g_fib[t].state == WAIT_WARP) g_fib[t].state = RUN;
        released = true;
      }
    }
    if (nblock && nblock + ndone == nthr){
      for (int t = 0; t < nthr; t++) if (g_fib[t].state == WAIT_BLOCK) g_fib[t].state = RUN;
      released = true;
    }
    if (!progressed && !released){
      fprintf (stderr, "pg_emu: deadlock in block (%u,%u): a collective was not reached by every lane\n", blockIdx.x, blockIdx.y);
      abort ();
    }
  }
}

void launch_impl (dim3 grid, dim3 block, size_t smem, const std::function<void ()> &body)
{
  if (block.y != 1 || block.z != 1){ fprintf (stderr, "pg_emu: one-dimensional blocks only\n"); abort (); }
  gridDim = grid; blockDim = block;
  g_body = &body;
  const bool reverse = getenv ("PG_EMU_REVERSE") != nullptr && getenv ("PG_EMU_REVERSE")[0] == '1';
  for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned qx = 0; qx < grid.x; qx++){
    const unsigned bx = reverse ? grid.x - 1 - qx : qx;
    blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
    g_smem.assign (smem + 64, 0);
    run_block ((int)block.x);
  }
  g_body = nullptr;
}

} // namespace pg_emu
