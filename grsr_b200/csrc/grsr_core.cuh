// grsr_core.cuh -- device code of the reversal-distance hot path (sm_100a).
//
// One thread block owns one permutation pair (gamma, pi) and keeps that pair's whole breakpoint
// graph in shared memory; a launch processes many pairs (each block walks a contiguous chunk of the
// pair list).  What it reproduces, bit for bit, is the reference GRIMM 2.01 pipeline
// invdist_noncircular_G (reference G/uniinvdist.c:150-214):
//
//   reference stage (file:line)                           here
//   ---------------------------------------------------   -----------------------------------------
//   double_perms            G/graph_edit.c:46-157          phase A: grey edges written straight from
//   buildG_grey             G/mcrdist.c:116-145              the signed positions of gamma's genes in pi
//   num_breakpoints         G/uniinvdist.c:124-142         phase A (adjacency count)
//   classify_cycle_path     G/graph_edit.c:207-457         phase B (pointer jumping) + phase C
//   classify_connected_c.   G/graph_components.c:257-336   orientation bit carried by phase B / S6
//   form_connected_comp.    G/graph_components.c:50-242    slow path S1-S5 (extent queries + union-find)
//   calc_num_hurdles_and_f. G/graph_components.c:465-727   slow path S7-S9
//
// The sequential walks and the stack sweep of the reference are replaced by data-parallel
// formulations (DESIGN.md section 3 gives the equivalence arguments); the numbers produced are the
// same.  All arithmetic is 32-bit integer; vertex ids fit 15 bits in shared memory (2n+2 <= 32767).
//
// The file is also compiled by g++ against tests/emu/cuda_emu.hpp (a SIMT emulator used only by
// the CPU test-suite to debug kernel logic); nothing in the product selects that build.
#pragma once
#include <stdint.h>
#include <type_traits>
#ifndef GRSR_EMU
#include <cooperative_groups.h>
#endif

#ifdef GRSR_EMU
#define GRSR_DYN_SMEM(name) unsigned char *name = emu::dyn_smem()
// The emulator runs one thread at a time, which would let late threads of a jumping round see
// words already updated in that round (the most favourable schedule).  This extra barrier between
// the gathers and the stores forces the least favourable one -- everybody reads before anybody
// writes -- so that the round count is tested against the guaranteed doubling only.
#define GRSR_EMU_READ_BEFORE_WRITE() __syncthreads()
#define GRSR_EMU_READ_BEFORE_WRITE_TM() TM::sync()
#else
#define GRSR_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define GRSR_EMU_READ_BEFORE_WRITE() ((void)0)
#define GRSR_EMU_READ_BEFORE_WRITE_TM() ((void)0)
#endif

namespace grsr {

#ifdef GRSR_EMU
__device__ inline unsigned cluster_rank() { return emu::cluster_rank(); }
__device__ inline unsigned cluster_size() { return emu::cluster_size(); }
__device__ inline void cluster_sync() { emu::cluster_barrier(); }
#else
__device__ inline unsigned cluster_rank() { return cooperative_groups::this_cluster().block_rank(); }
__device__ inline unsigned cluster_size() { return cooperative_groups::this_cluster().num_blocks(); }
__device__ inline void cluster_sync() { cooperative_groups::this_cluster().sync(); }
#endif

typedef uint16_t vid_t;
static const unsigned kFull = 0xFFFFFFFFu;
static const int kMaxSmemGenes = 14000;   // 2n+2 <= 32767 and ~10 B per vertex within 227 KB smem

// Jumping word of vertex v:  [31:17] label = smallest vertex seen on v's sigma-orbit window,
//                            [16:2]  next = end of the window (sigma^k(v)), i.e. bits [16:0] & ~3
//                                    are the byte offset of W[next],
//                            [1]     1 if an oriented grey edge was seen on that window,
//                            [0]     1 if v is an end of an adjacency (then sigma(v) = v).
#define GRSR_W_NEXT(w) (((w) >> 2) & 0x7FFFu)
#define GRSR_W_ORBIT 2u
#define GRSR_W_ADJ 1u
#define GRSR_W_MAKE(label, next, flags) (((unsigned)(label) << 17) | ((unsigned)(next) << 2) | (flags))

struct Stats { int d, b, c, h, f; };

// ---- phase timers (diagnostic build only: -DGRSR_PHASE_PROFILE, scripts/prof_phases.py) ------------
// Thread 0 of every block charges the cycles since the previous mark to the phase that was running.
// Phase ids: see scripts/prof_phases.py.  In the product build the marks compile to nothing.
#if defined(GRSR_PHASE_PROFILE) && !defined(GRSR_EMU)
__device__ unsigned long long grsr_prof_cycles[64];
__device__ unsigned long long grsr_prof_counts[64];
struct PhaseState { long long last; int cur; int base; };
__device__ inline PhaseState &phase_state() { __shared__ PhaseState ps; return ps; }
__device__ inline void phase_init() {
  if (threadIdx.x == 0) { PhaseState &ps = phase_state(); ps.last = clock64(); ps.cur = 0; ps.base = 0; }
}
__device__ inline void phase_mark(int id) {
  if (threadIdx.x == 0) {
    PhaseState &ps = phase_state();
    const long long t = clock64();
    atomicAdd(&grsr_prof_cycles[ps.cur], (unsigned long long)(t - ps.last));
    atomicAdd(&grsr_prof_counts[id], 1ull);
    ps.last = t; ps.cur = id;
  }
}
__device__ inline void phase_base(int b) { if (threadIdx.x == 0) phase_state().base = b; }
__device__ inline void phase_mark_rel(int k) { if (threadIdx.x == 0) phase_mark(phase_state().base + k); }
#define GRSR_PHASE_INIT() phase_init()
#define GRSR_PHASE(id) phase_mark(id)
#define GRSR_PHASE_BASE(b) phase_base(b)
#define GRSR_PHASE_REL(k) phase_mark_rel(k)
#else
#define GRSR_PHASE_INIT() ((void)0)
#define GRSR_PHASE(id) ((void)0)
#define GRSR_PHASE_BASE(b) ((void)0)
#define GRSR_PHASE_REL(k) ((void)0)
#endif

// Where the result of pair p = (gamma row gi, pi row pj) goes.  stride 1: out[p] = d; stride 5:
// out + 5p = {d,b,c,h,f}; stride -G (matrix mode): the symmetric scatter of setinvmatrix
// (G/uniinvdist.c:84-88) done by the kernel itself, out[gi*G + pj] = out[pj*G + gi] = d -- `out` may
// then be a matrix in ANOTHER device's memory (peer mapping over NVLink): 8 bytes per pair.
__device__ inline void put_result(int32_t *__restrict__ out, int stride, long long p, int gi, int pj,
                                  int d, int b, int c, int h, int f) {
  if (stride == 1) out[p] = d;
  else if (stride == 5) { int32_t *o = out + 5 * p; o[0] = d; o[1] = b; o[2] = c; o[3] = h; o[4] = f; }
  else {
    const long long G = -stride;
    out[(long long)gi * G + pj] = d;
    out[(long long)pj * G + gi] = d;
  }
}

template <class VT> __host__ __device__ constexpr unsigned none_of() { return (unsigned)(VT)(~(VT)0); }

// Workspace of one pair.  S = 2n+2 vertices.  VT = vertex id type: uint16_t when the graph lives in
// shared memory (2n+2 <= 32767), uint32_t for the global-memory build used beyond that.
template <class VT>
struct WorkT {
  void *W;                  // jumping words: uint32_t[Sp] (shared-memory build, Sp = slots of the
                            //   block) or uint64_t[S] (global-memory build)
  uint32_t *P;              // [S] per-cycle scratch of the slow path (parent / reach / counters);
                            //   aliases W in the shared-memory build
  VT *A1;                   // [S]  slow path: per-root flags/links, then compacted lists
  VT *A2;                   // [S]  slow path: cycle id per vertex, then compacted lists
  VT *ORB;                  // [S]  sigma-orbit label per vertex (scenario kernels only)
  VT *BLK;                  // [2*NB] per-32-vertex block: min cycle id, max cycle reach
  unsigned long long *red;  // [34] reduction scratch (always shared memory)
  unsigned long long *xch;  // cluster build: global exchange area of the cluster (kXchWords), else null
  unsigned epoch;           // cluster build: number of team collectives issued so far (buffer parity)
  int S, Sp, NB;
};
typedef WorkT<uint16_t> Work;

__host__ __device__ inline int padded_size(int n, int slots) {
  int s = 2 * n + 2;
  return ((s + slots - 1) / slots) * slots;
}
__host__ __device__ inline int num_vblocks(int n) { return (2 * n + 2 + 31) >> 5; }

// slots = threads per block x words per thread (a multiple of 32); only W is padded to it
__host__ __device__ inline size_t work_bytes(int n, int slots, bool keep_orbits) {
  size_t sp = (size_t)padded_size(n, slots), sv = (size_t)((2 * n + 2 + 7) & ~7);
  size_t b = sp * 4 + sv * 2 + sv * 2 + (keep_orbits ? sv * 2 : 0);
  b += (((size_t)num_vblocks(n) * 2 * sizeof(vid_t)) + 15) & ~(size_t)15;
  b += 34 * sizeof(unsigned long long);
  return b;
}

__device__ inline void work_carve(Work &wk, unsigned char *base, int n, int slots, bool keep_orbits) {
  int sp = padded_size(n, slots), sv = (2 * n + 2 + 7) & ~7;
  wk.S = 2 * n + 2; wk.Sp = sp; wk.NB = num_vblocks(n);
  wk.W = (void *)base; wk.P = (uint32_t *)base;  base += (size_t)sp * 4;
  wk.A1 = (vid_t *)base;    base += (size_t)sv * 2;
  wk.A2 = (vid_t *)base;    base += (size_t)sv * 2;
  wk.ORB = keep_orbits ? (vid_t *)base : (vid_t *)0;
  if (keep_orbits) base += (size_t)sv * 2;
  wk.BLK = (vid_t *)base;   base += (((size_t)wk.NB * 2 * sizeof(vid_t)) + 15) & ~(size_t)15;
  wk.red = (unsigned long long *)base;
  wk.xch = (unsigned long long *)0; wk.epoch = 0;
}

// ---- block-wide reductions (every thread of the block must call; result returned to all) ----------

__device__ inline unsigned long long blk_sum_u64(unsigned long long v, unsigned long long *red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(kFull, v, o);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = (lane < nw) ? red[lane] : 0ull;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(kFull, v, o);
    if (lane == 0) red[32] = v;
  }
  __syncthreads();
  v = red[32];
  __syncthreads();
  return v;
}

__device__ inline unsigned blk_max_u32(unsigned v, unsigned long long *red) {
  v = __reduce_max_sync(kFull, v);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = (lane < nw) ? (unsigned)red[lane] : 0u;
    v = __reduce_max_sync(kFull, v);
    if (lane == 0) red[32] = v;
  }
  __syncthreads();
  v = (unsigned)red[32];
  __syncthreads();
  return v;
}

// Ordered compaction: out[0..total) = value(i) for the i in [0,N) with f(i, value) true, in
// increasing i.  Each warp owns a contiguous chunk, counts with ballots, then writes.
template <class VT, class F>
__device__ inline int blk_compact(int N, F f, VT *out, unsigned long long *red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int chunk = (((N + nw - 1) / nw) + 31) & ~31;
  int lo = wid * chunk, hi = min(N, lo + chunk);
  int cnt = 0;
  for (int base = lo; base < hi; base += 32) {
    int i = base + lane; unsigned val = 0;
    bool ok = (i < hi) && f(i, val);
    cnt += __popc(__ballot_sync(kFull, ok));
  }
  uint32_t *cnts = (uint32_t *)red;
  if (lane == 0) cnts[wid] = (uint32_t)cnt;
  __syncthreads();
  int pos = 0, total = 0;
  for (int w = 0; w < nw; w++) { int cw = (int)cnts[w]; if (w < wid) pos += cw; total += cw; }
  for (int base = lo; base < hi; base += 32) {
    int i = base + lane; unsigned val = 0;
    bool ok = (i < hi) && f(i, val);
    unsigned m = __ballot_sync(kFull, ok);
    if (ok) out[pos + __popc(m & ((1u << lane) - 1u))] = (VT)val;
    pos += __popc(m);
  }
  __syncthreads();
  return total;
}

// --------------------------------------------------------------------------------------------------
// Teams: the set of threads that works on one pair.  TeamBlock = one thread block (every kernel of
// the shared-memory build and the plain global-memory build).  TeamCluster = a thread-block cluster
// (up to 16 SMs) working on one pair of the global-memory build: same code, team-wide thread ids,
// cluster barriers, and collectives that combine per-block partial results through a small global
// exchange area (double-buffered by the collective count, so one cluster barrier per collective).
// --------------------------------------------------------------------------------------------------
struct TeamBlock {
  static __device__ inline int tid() { return threadIdx.x; }
  static __device__ inline int size() { return blockDim.x; }
  static __device__ inline void sync() { __syncthreads(); }
  template <class WK> static __device__ inline int any(int pred, WK &) { return __syncthreads_or(pred); }
  template <class WK> static __device__ inline unsigned long long sum(unsigned long long v, WK &wk) {
    return blk_sum_u64(v, wk.red);
  }
  template <class WK> static __device__ inline unsigned max32(unsigned v, WK &wk) { return blk_max_u32(v, wk.red); }
  template <class VT, class F, class WK>
  static __device__ inline int compact(int N, F f, VT *out, WK &wk) { return blk_compact(N, f, out, wk.red); }
};

static const int kMaxClusterBlocks = 16;
static const int kXchWords = 2 * kMaxClusterBlocks + 2 * kMaxClusterBlocks * 32 + 8;  // u64 words

struct TeamCluster {
  static __device__ inline int tid() { return (int)(cluster_rank() * blockDim.x + threadIdx.x); }
  static __device__ inline int size() { return (int)(cluster_size() * blockDim.x); }
  // barrier.cluster arrive.release / wait.acquire: orders the global-memory accesses of the cluster
  static __device__ inline void sync() { cluster_sync(); }
  // slot of this collective in the exchange area (two alternating buffers of kMaxClusterBlocks)
  template <class WK> static __device__ inline volatile unsigned long long *slot(WK &wk) {
    volatile unsigned long long *p = wk.xch + (wk.epoch & 1u) * kMaxClusterBlocks;
    wk.epoch++;
    return p;
  }
  // every lane fetches one block's partial (one L2 round trip), the warp combines them
  static __device__ inline unsigned long long part(volatile unsigned long long *x, unsigned long long dflt) {
    const unsigned lane = threadIdx.x & 31;
    return (lane < cluster_size()) ? x[lane] : dflt;
  }
  template <class WK> static __device__ inline int any(int pred, WK &wk) {
    volatile unsigned long long *x = slot(wk);
    int b = __syncthreads_or(pred);
    if (threadIdx.x == 0) x[cluster_rank()] = (unsigned long long)b;
    sync();
    return __any_sync(kFull, part(x, 0ull) != 0ull);
  }
  template <class WK> static __device__ inline unsigned long long sum(unsigned long long v, WK &wk) {
    volatile unsigned long long *x = slot(wk);
    v = blk_sum_u64(v, wk.red);
    if (threadIdx.x == 0) x[cluster_rank()] = v;
    sync();
    unsigned long long r = part(x, 0ull);
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(kFull, r, o);
    return r;
  }
  template <class WK> static __device__ inline unsigned max32(unsigned v, WK &wk) {
    volatile unsigned long long *x = slot(wk);
    v = blk_max_u32(v, wk.red);
    if (threadIdx.x == 0) x[cluster_rank()] = v;
    sync();
    return __reduce_max_sync(kFull, (unsigned)part(x, 0ull));
  }
  // ordered compaction over the whole team: every warp of the cluster owns a contiguous chunk
  template <class VT, class F, class WK>
  static __device__ inline int compact(int N, F f, VT *out, WK &wk) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int wid = (int)cluster_rank() * wpb + (threadIdx.x >> 5), nw = (int)cluster_size() * wpb;
    volatile unsigned *cnts = (volatile unsigned *)(wk.xch + 2 * kMaxClusterBlocks) +
                              (wk.epoch & 1u) * (kMaxClusterBlocks * 32);
    wk.epoch++;
    int chunk = (((N + nw - 1) / nw) + 31) & ~31;
    int lo = wid * chunk, hi = min(N, lo + chunk);
    int cnt = 0;
    for (int base = lo; base < hi; base += 32) {
      int i = base + lane; unsigned val = 0;
      bool ok = (i < hi) && f(i, val);
      cnt += __popc(__ballot_sync(kFull, ok));
    }
    if (lane == 0) cnts[wid] = (unsigned)cnt;
    sync();
    // counts of the warps before mine and of all warps: lanes fetch strided, the warp adds up
    unsigned before = 0, all = 0;
    for (int w = lane; w < nw; w += 32) { unsigned cw = cnts[w]; all += cw; if (w < wid) before += cw; }
    unsigned long long both = ((unsigned long long)before << 32) | all;
    for (int o = 16; o > 0; o >>= 1) both += __shfl_xor_sync(kFull, both, o);
    int pos = (int)(both >> 32);
    const int total = (int)(both & 0xFFFFFFFFull);
    for (int base = lo; base < hi; base += 32) {
      int i = base + lane; unsigned val = 0;
      bool ok = (i < hi) && f(i, val);
      unsigned m = __ballot_sync(kFull, ok);
      if (ok) out[pos + __popc(m & ((1u << lane) - 1u))] = (VT)val;
      pos += __popc(m);
    }
    sync();
    return total;
  }
};

__device__ inline int ceil_log2(int x) { int r = 0; while ((1 << r) < x) r++; return r; }

// Follow parent words up to the representative (union-find forest kept in W[root]).
__device__ inline unsigned uf_find(const uint32_t *W, unsigned x) {
  unsigned p;
  while ((p = W[x]) != x) x = p;
  return x;
}

// Ordered compaction of RUN STARTS (TeamBlock only): among the i in [0,N) with f(i, value) true, taken
// in increasing i, keeps value(i) whenever it differs from the value of the previous kept i -- the
// run-length encoding of the filtered sequence in ONE counting and ONE writing sweep (instead of
// compacting the sequence and then its run starts).  The counting sweep leaves value(i) (or NONE) in
// cache[i], so that the writing sweep does not evaluate f again; out may alias whatever f reads, but
// not cache.  Each warp owns a contiguous chunk; a chunk's first run merges into the previous non-empty
// chunk's last run when their values are equal (warp 0 sorts that out with shuffles).
// values must be < 2^21 - 1 (packed per-warp records).  Returns the number of runs.
template <class VT, class F>
__device__ inline int blk_compact_runs(int N, F f, VT *cache, VT *out, unsigned long long *red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const unsigned kNoVal = 0x1FFFFFu;
  constexpr unsigned kNone = none_of<VT>();
  const int chunk = (((N + nw - 1) / nw) + 31) & ~31;
  const int lo = wid * chunk, hi = min(N, lo + chunk);
  unsigned cnt = 0, first = kNoVal, carry = kNoVal;
  for (int base = lo; base < hi; base += 32) {
    const int i = base + lane; unsigned val = 0;
    const bool ok = (i < hi) && f(i, val);
    if (i < hi) cache[i] = (VT)(ok ? val : kNone);
    const unsigned m = __ballot_sync(kFull, ok);
    if (m == 0u) continue;
    const unsigned below = m & lt;
    const unsigned pv = __shfl_sync(kFull, val, below ? 31 - __clz((int)below) : 0);
    const unsigned prev = below ? pv : carry;
    cnt += __popc(__ballot_sync(kFull, ok && val != prev));
    if (first == kNoVal) first = __shfl_sync(kFull, val, __ffs((int)m) - 1);
    carry = __shfl_sync(kFull, val, 31 - __clz((int)m));
  }
  if (lane == 0) red[wid] = ((unsigned long long)cnt << 42) | ((unsigned long long)first << 21) | carry;
  __syncthreads();
  if (wid == 0) {
    // lane w holds chunk w's record: start position of its runs and the value just before its chunk
    const unsigned long long r = (lane < nw) ? red[lane] : ((unsigned long long)kNoVal << 21) | kNoVal;
    unsigned c = (unsigned)(r >> 42);
    const unsigned fv = (unsigned)(r >> 21) & kNoVal, lv = (unsigned)r & kNoVal;
    const unsigned nonempty = __ballot_sync(kFull, fv != kNoVal);
    const unsigned before_m = nonempty & lt;
    const unsigned plv = __shfl_sync(kFull, lv, before_m ? 31 - __clz((int)before_m) : 0);
    const unsigned before = before_m ? plv : kNoVal;
    if (fv != kNoVal && fv == before) c--;              // first run continues the previous chunk's last run
    unsigned incl = c;
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += y; }
    red[lane] = ((unsigned long long)(incl - c) << 32) | before;
    if (lane == 31) red[32] = incl;
  }
  __syncthreads();
  unsigned pos = (unsigned)(red[wid] >> 32);
  carry = (unsigned)red[wid];
  const unsigned total = (unsigned)red[32];
  for (int base = lo; base < hi; base += 32) {
    const int i = base + lane;
    const unsigned val = (i < hi) ? (unsigned)cache[i] : kNone;
    const bool ok = val != kNone;
    const unsigned m = __ballot_sync(kFull, ok);
    if (m == 0u) continue;
    const unsigned below = m & lt;
    const unsigned pv = __shfl_sync(kFull, val, below ? 31 - __clz((int)below) : 0);
    const unsigned prev = below ? pv : carry;
    const bool start = ok && val != prev;
    const unsigned sm = __ballot_sync(kFull, start);
    if (start) out[pos + __popc(sm & lt)] = (VT)val;
    pos += __popc(sm);
    carry = __shfl_sync(kFull, val, 31 - __clz((int)m));
  }
  __syncthreads();
  return (int)total;
}

// --------------------------------------------------------------------------------------------------
// Hurdles and fortress for a graph that has at least one unoriented cycle.
// Entry state: A2[v] = cycle id (leftmost vertex) or NONE for adjacency vertices;
//              A1[r] = 1 if cycle r (r = its leftmost vertex, always even) owns an oriented grey
//              edge, else 0;
//              P[r] = r and P[r+1] = 0 for every cycle r (the union-find forest and the reach, "S1",
//              written by the caller together with A2 / A1; P may alias the jumping words, which are
//              dead by now).
// Returns (h << 1) | f.
// --------------------------------------------------------------------------------------------------
#ifndef GRSR_SMALL_EXTENT
#define GRSR_SMALL_EXTENT 48   // S4: extents up to this many vertices are scanned by one lane
#endif
// kComponentsOnly: stop after S6 -- on return A2[v] = cycle id, P[c] = component (root cycle) of cycle c,
// P[comp + 1] = 1 if the component is oriented; h and f are not computed (returns 0).  Used by the
// unsigned-distance kernel, which needs the components of sign_2strips (G/unsigned.c:958-1040).
template <class TM, class VT, bool kComponentsOnly = false>
__device__ inline unsigned hurdles_slow_path(WorkT<VT> &wk, int n) {
  const int tid = TM::tid(), nt = TM::size(), lane = tid & 31;
  const int S = wk.S;
  constexpr unsigned kNone = none_of<VT>();
  constexpr unsigned kNoLink = kNone & ~1u;
  uint32_t *W = wk.P; VT *A1 = wk.A1, *A2 = wk.A2, *BLK = wk.BLK;
  VT *bmin = BLK, *bmax = BLK + wk.NB;

  // S2: reach of every cycle = max vertex carrying its id (reference: right[], G/graph_edit.c:392).
  //     Vertices are visited from the right, so the first sweep already holds the reach of every long
  //     cycle and later vertices only compare; the atomic is issued by the few that still improve it.
  GRSR_PHASE(21);
  for (int v = S - 1 - tid; v >= 0; v -= nt) {
    const unsigned c = A2[v];
    if (c != kNone && W[c + 1] < (uint32_t)v) atomicMax(&W[c + 1], (uint32_t)v);
  }
  TM::sync();

  // S3: per block of 32 vertices: smallest cycle id and largest reach seen in it (one warp per block).
  GRSR_PHASE(22);
  for (int B = tid >> 5; B < wk.NB; B += nt >> 5) {
    const int v = (B << 5) + lane;
    unsigned mn = kNone, mx = 0;
    if (v < S) {
      const unsigned c = A2[v];
      if (c != kNone) { mn = c; mx = W[c + 1]; }
    }
    mn = __reduce_min_sync(kFull, mn);
    mx = __reduce_max_sync(kFull, mx);
    if (lane == 0) { bmin[B] = (VT)mn; bmax[B] = (VT)mx; }
  }
  TM::sync();

  // S4: for every cycle P = (l, r): among the vertices strictly inside (l, r), the smallest cycle id
  //     (a cycle that starts left of P and reaches into it) and the largest reach (a cycle that
  //     reaches into P and ends right of it).  Those two neighbours are enough to recover the
  //     components of the reference's stack sweep (G/graph_components.c:101-138).
  //     A short extent is scanned by the cycle's own lane (four vertices per round, so that the
  //     dependent id -> reach loads overlap); long ones are handed to the whole warp.
  GRSR_PHASE(23);
  for (int base = tid & ~31; base <= n; base += nt) {
    int k = base + lane;
    unsigned r0 = 0, R0 = 0; bool isroot = false, big = false;
    if (k <= n) { r0 = 2u * k; isroot = (A2[r0] == r0); if (isroot) R0 = W[r0 + 1]; }
    if (isroot) {
      big = (R0 - r0) > (unsigned)GRSR_SMALL_EXTENT;
      if (!big) {
        unsigned mn = kNone, mx = 0;
        unsigned v = r0 + 2;                           // r0 + 1 is P's own vertex (black edge r0, r0+1)
        for (; v + 4 <= R0; v += 4) {
          const unsigned c0 = A2[v], c1 = A2[v + 1], c2 = A2[v + 2], c3 = A2[v + 3];
          const unsigned x0 = (c0 != kNone) ? W[c0 + 1] : 0u, x1 = (c1 != kNone) ? W[c1 + 1] : 0u;
          const unsigned x2 = (c2 != kNone) ? W[c2 + 1] : 0u, x3 = (c3 != kNone) ? W[c3 + 1] : 0u;
          mn = min(mn, min(min(c0, c1), min(c2, c3)));
          mx = max(mx, max(max(x0, x1), max(x2, x3)));
        }
        for (; v < R0; v++) {
          const unsigned c = A2[v];
          if (c != kNone) { mn = min(mn, c); mx = max(mx, W[c + 1]); }
        }
        unsigned lk1 = (mn < r0) ? mn : kNoLink;
        unsigned lk2 = (mx > R0) ? (unsigned)A2[mx] : kNone;
        A1[r0] = (VT)((A1[r0] & 1u) | lk1);
        A1[r0 + 1] = (VT)lk2;
      }
    }
    unsigned todo = __ballot_sync(kFull, isroot && big);
    while (todo) {
      int src = __ffs((int)todo) - 1; todo &= todo - 1;
      unsigned rr = __shfl_sync(kFull, r0, src), RR = __shfl_sync(kFull, R0, src);
      unsigned mn = kNone, mx = 0;
      int lo = (int)rr + 1, hi = (int)RR - 1;
      if (lo <= hi) {
        int bl = lo >> 5, bh = hi >> 5;
        int v = (bl << 5) + lane;
        if (v >= lo && v <= hi) {
          unsigned c = A2[v];
          if (c != kNone) { mn = min(mn, c); mx = max(mx, W[c + 1]); }
        }
        if (bh > bl) {
          v = (bh << 5) + lane;
          if (v <= hi) {
            unsigned c = A2[v];
            if (c != kNone) { mn = min(mn, c); mx = max(mx, W[c + 1]); }
          }
        }
        for (int B = bl + 1 + lane; B < bh; B += 32) {
          mn = min(mn, (unsigned)bmin[B]); mx = max(mx, (unsigned)bmax[B]);
        }
      }
      mn = __reduce_min_sync(kFull, mn);
      mx = __reduce_max_sync(kFull, mx);
      if (lane == src) {
        unsigned lk1 = (mn < rr) ? mn : kNoLink;
        unsigned lk2 = (mx > RR) ? (unsigned)A2[mx] : kNone;
        A1[rr] = (VT)((A1[rr] & 1u) | lk1);
        A1[rr + 1] = (VT)lk2;
      }
    }
  }
  TM::sync();

  // S5: connected components of the cycles under those links: hook larger root under smaller,
  //     compress, repeat until no link joins two trees.  The reach (W[r+1]) is not needed any more:
  //     the first sweep clears it, S6 reuses the word as the component's orientation flag.
  GRSR_PHASE(24);
  for (bool first_sweep = true;; first_sweep = false) {
    int changed = 0;
    for (int k = tid; k <= n; k += nt) {
      unsigned r = 2u * k;
      if (A2[r] != r) continue;
      if (first_sweep) W[r + 1] = 0;
      unsigned me = uf_find(W, r);
      unsigned l1 = A1[r] & kNoLink, l2 = A1[r + 1];
      if (l1 != kNoLink) {
        unsigned o = uf_find(W, l1);
        if (o != me) { atomicMin(&W[max(o, me)], min(o, me)); changed = 1; }
      }
      if (l2 != kNone) {
        me = uf_find(W, r);
        unsigned o = uf_find(W, l2);
        if (o != me) { atomicMin(&W[max(o, me)], min(o, me)); changed = 1; }
      }
    }
    if (!TM::any(changed, wk)) break;
    for (int k = tid; k <= n; k += nt) {
      unsigned r = 2u * k;
      if (A2[r] == r) W[r] = uf_find(W, r);
    }
    TM::sync();
  }

  // S6: a component is oriented iff one of its cycles is (G/graph_components.c:309-336).
  //     W[c+1] of the component root c becomes that flag.
  GRSR_PHASE(25);
  for (int k = tid; k <= n; k += nt) {
    unsigned r = 2u * k;
    if (A2[r] == r && (A1[r] & 1u)) W[W[r] + 1] = 1;
  }
  TM::sync();

  if (kComponentsOnly) return 0;

  // S7 + S8: the vertices of unoriented components, in vertex order, as component ids (the filtered
  //     scan of G/graph_components.c:566-585); runs of equal ids = the reference's "blocks", one
  //     entry per run.
  GRSR_PHASE(26);
  auto unoriented_comp = [&](int v, unsigned &val) -> bool {
    unsigned c = A2[v];
    if (c == kNone) return false;
    unsigned comp = W[c];
    if (W[comp + 1]) return false;
    val = comp; return true;
  };
  int nb;
  const VT *run;
  if constexpr (std::is_same<TM, TeamBlock>::value) {
    // A1's flags and links are dead by now: it caches the sweep's values; the runs replace the cycle ids
    nb = blk_compact_runs(S, unoriented_comp, A1, A2, wk.red);
    run = A2;
    if (nb == 0) return 0;
  } else {
    int m = TM::compact(S, unoriented_comp, A1, wk);
    if (m == 0) return 0;
    GRSR_PHASE(27);
    nb = TM::compact(m, [&](int t, unsigned &val) -> bool {
      unsigned c = A1[t];
      if (t > 0 && A1[t - 1] == c) return false;
      val = c; return true;
    }, A2, wk);
    run = A2;
  }

  // S9: per component: number of runs (W[c]) and 1 + index of its last run (W[c+1]).
  GRSR_PHASE(28);
  for (int t = tid; t < nb; t += nt) { unsigned c = run[t]; W[c] = 0; W[c + 1] = 0; }
  TM::sync();
  for (int t = tid; t < nb; t += nt) {
    unsigned c = run[t];
    atomicAdd(&W[c], 1u);
    atomicMax(&W[c + 1], (uint32_t)(t + 1));
  }
  TM::sync();
  const unsigned first = run[0], last = run[nb - 1];
  // 1 + index of the last-but-one run of the final component (needed for its "right" neighbour)
  unsigned pl = 0;
  for (int t = tid; t < nb - 1; t += nt) if (run[t] == last) pl = max(pl, (unsigned)(t + 1));
  const unsigned prevlast = TM::max32(pl, wk);

  // Hurdles (one run), greatest hurdle (first == last with two runs), super-hurdles
  // (G/graph_components.c:603-624, 640-656, 674-694).  left/right are the reference's
  // transition-overwritten neighbours: left = run before the component's last run, right = run
  // after its last run that has a successor.
  unsigned hcount = 0, scount = 0;
  for (int t = tid; t < nb; t += nt) {
    unsigned c = run[t];
    if (W[c + 1] != (uint32_t)(t + 1)) continue;          // only the thread at the last run of c
    unsigned blocks = W[c];
    bool great = (first == last && c == first && blocks == 2);
    bool hurdle = (blocks == 1) || great;
    if (!hurdle) continue;
    hcount++;
    unsigned left = (t >= 1) ? (unsigned)run[t - 1] : kNone;
    unsigned right = (t < nb - 1) ? (unsigned)run[t + 1]
                                  : (prevlast > 0 ? (unsigned)run[prevlast] : kNone);
    if (left != kNone && left == right) {
      bool nb_great = (first == last && left == first && W[left] == 2);
      if (W[left] == 2 && !nb_great) scount++;
    }
  }
  if (tid == 0 && first != last) {
    // a hurdle at one end protecting a non-hurdle at the other end (:640-656)
    bool hf = (W[first] == 1), hl = (W[last] == 1);
    unsigned right_first = run[W[first + 1]];             // last run of `first` is not the final run
    unsigned left_last = run[nb - 2];
    if (right_first == last && hf && !hl) scount++;
    else if (left_last == first && hl && !hf) scount++;
  }
  unsigned long long tot = TM::sum(((unsigned long long)hcount << 32) | scount, wk);
  unsigned h = (unsigned)(tot >> 32), s = (unsigned)(tot & 0xFFFFFFFFu);
  unsigned f = (h == s && (s & 1u)) ? 1u : 0u;
  return (h << 1) | f;
}

// --------------------------------------------------------------------------------------------------
// Cheap sufficient test run before the slow path: is every unoriented cycle certainly part of an
// oriented component?  For an unoriented cycle Q with leftmost vertex q, the vertices between q+1
// and Q's next vertex lie strictly inside Q's extent; if one of them belongs to an ORIENTED cycle P
// that starts left of q, P has a vertex inside Q's extent and one outside it, so P and Q share a
// component (the interleaving relation behind G/graph_components.c:101-138) and that component is
// oriented.  If this holds for every unoriented cycle there is no unoriented component: h = f = 0.
// On random permutations (a few giant oriented cycles) it settles practically every case with one
// 32-vertex probe per unoriented cycle.  word(v) = label << 2 | oriented << 1 | adjacency.
// Returns the verdict to all threads; "false" only means "not settled, run the slow path".
// --------------------------------------------------------------------------------------------------
template <class TM, class WK, class WordFn>
__device__ inline bool unoriented_cycles_all_covered(WK &wk, int n, int S, WordFn word) {
  const int tid = TM::tid(), nt = TM::size(), lane = tid & 31;
  int unresolved = 0;
  for (int base = tid & ~31; base <= n && !unresolved; base += nt) {
    const int k = base + lane;
    const unsigned q = 2u * k;
    bool isq = false;
    if (k <= n) {
      const unsigned w0 = word(q), w1 = word(q + 1);
      isq = !(w0 & 1u) && (min(w0 >> 2, w1 >> 2) == q) && !(w0 & 2u);
    }
    unsigned todo = __ballot_sync(kFull, isq);
    while (todo) {
      const int src = __ffs((int)todo) - 1; todo &= todo - 1;
      const unsigned qq = __shfl_sync(kFull, q, src);
      bool hit = false;
      for (int it = 0; it < 8; it++) {
        const unsigned v = qq + 2u + 32u * it + lane;
        int st = 2;                                   // past the last vertex: stop
        if (v < (unsigned)S) {
          st = 0;
          const unsigned w = word(v);
          if (!(w & 1u)) {
            const unsigned cy = min(w >> 2, word(v ^ 1u) >> 2);
            if (cy == qq) st = 2;                     // Q's own next vertex: the probe ends here
            else if (cy < qq && (w & 2u)) st = 1;     // an oriented cycle that starts left of Q
          }
        }
        const unsigned mh = __ballot_sync(kFull, st == 1), ms = __ballot_sync(kFull, st == 2);
        if (mh | ms) { hit = mh && (!ms || __ffs((int)mh) < __ffs((int)ms)); break; }
      }
      if (!hit) { unresolved = 1; break; }            // one unsettled cycle decides: stop probing
    }
  }
  return !TM::any(unresolved, wk);
}

// --------------------------------------------------------------------------------------------------
// dist_core: full invdist_noncircular_G for the pair (gamma, pi) described by ends(), which tells
// where the two ends of gamma's t-th gene sit in pi.  Vertices are the doubled positions of pi,
// 0..2n+1, exactly the
// reference's numbering: gene at position p owns vertices 2p+1 (left end) and 2p+2 (right end) when
// it reads forward, swapped when it reads backward; black edges join (2k, 2k+1).  Grey edge t
// (t = 0..n) joins the right end of gamma's gene t-1 (vertex 0 for t = 0) to the left end of
// gamma's gene t (vertex 2n+1 for t = n) -- which is what perm / invperm / greyEdges of the
// reference encode (G/graph_edit.c:69-145, G/mcrdist.c:116-145).
//
// ends(t, L, R), t in [0, n): the vertices of the left and right end of gamma's t-th gene.
// NT = threads per block, EPT = jumping words kept in registers per thread; NT * EPT == wk.Sp.
// kKeepOrbits additionally leaves in wk.ORB[v] the sigma-orbit label of every vertex
// (sigma(v) = grey[v^1]; kNone for the two ends of an adjacency); two breakpoint vertices x, y have
// an even graphdist (G/graph_edit.c:1165-1238) exactly when ORB[x] == ORB[y].
// All threads of the block must call; all receive the result.
// --------------------------------------------------------------------------------------------------
// kComponentsOnly (unsigned-distance kernel): b and c as usual; st.h = 1 means "the component analysis
// ran (up to S6, see hurdles_slow_path) because some unoriented cycle may lie in an unoriented
// component", st.h = 0 means every component is oriented; st.f and st.d are not meaningful.
template <int NT, int EPT, bool kKeepOrbits, class EndsFn, bool kComponentsOnly = false>
__device__ inline Stats dist_core(Work &wk, int n, EndsFn ends) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int S = wk.S;
  constexpr unsigned kNone = none_of<uint16_t>();
  uint32_t *W = (uint32_t *)wk.W;

  // phase A: one lane per grey edge t = 0..n; each warp owns a contiguous run of edges so that the
  // right end of gene t-1 arrives by shuffle.  x = right end of gene t-1 (vertex 0 for t = 0),
  // y = left end of gene t (vertex 2n+1 for t = n).  sigma(x^1) = y and sigma(y^1) = x seed the
  // jumping words; the edge is an adjacency when x and y are the two ends of one black edge
  // (G/mcrdist.c:135-140) and oriented when x and y have the same parity
  // (G/graph_components.c:324-329).
  unsigned nadj = 0;
  GRSR_PHASE_REL(0);
  {
    const int per = ((((n + 1) + (NT >> 5) - 1) / (NT >> 5)) + 31) & ~31;
    const int t0 = wid * per;
    const int t1 = min(n + 1, t0 + per);
    unsigned carry = 0;
    if (t0 > 0 && t0 <= n) { unsigned l; ends(t0 - 1, l, carry); }
    for (int base = t0; base < t1; base += 32) {
      const int t = base + lane;
      unsigned y = (unsigned)S - 1u, r = 0;
      if (t < n) ends(t, y, r);
      unsigned x = __shfl_up_sync(kFull, r, 1);
      if (lane == 0) x = carry;
      carry = __shfl_sync(kFull, r, 31);
      if (t <= n) {
        const unsigned xb = x ^ 1u, yb = y ^ 1u;
        const unsigned fl = ((((x ^ y) & 1u) == 0u) ? GRSR_W_ORBIT : 0u) | ((xb == y) ? GRSR_W_ADJ : 0u);
        nadj += (xb == y);
        W[xb] = GRSR_W_MAKE(xb, y, fl);
        W[yb] = GRSR_W_MAKE(yb, x, fl);
      }
    }
  }
  for (int v = S + tid; v < NT * EPT; v += NT) W[v] = GRSR_W_MAKE(v, v, GRSR_W_ADJ);  // padding
  __syncthreads();

  // phase B: pointer jumping along sigma with min-label and or-flag, in place.  Each word is one
  // atomic 32-bit (label, next, flags) snapshot, so after round r every word covers >= 2^r
  // successors; stop after ceil(log2(n+1)) rounds or as soon as a whole round changed no label and
  // no flag.  A thread keeps its EPT words in registers; per round it gathers the words of its
  // successors (in batches, loads before stores), merges and publishes.
  GRSR_PHASE_REL(1);
  {
    constexpr int UB = (EPT < 8) ? EPT : (EPT % 8 == 0 ? 8 : (EPT % 6 == 0 ? 6 : 5));   // gathers per batch; divides EPT
    static_assert(EPT % UB == 0, "words per thread must split into whole gather batches");
    uint32_t R[EPT];
#pragma unroll
    for (int j = 0; j < EPT; j++) R[j] = W[tid + j * NT];
    const int rounds = ceil_log2(n + 1);
    const char *Wb = (const char *)W;
    for (int r = 0; r < rounds; r++) {
      uint32_t chg = 0;
#pragma unroll
      for (int j0 = 0; j0 < EPT; j0 += UB) {
        uint32_t U[UB];
#pragma unroll
        for (int j = 0; j < UB; j++) U[j] = *(const uint32_t *)(Wb + (R[j0 + j] & 0x1FFFCu));
        GRSR_EMU_READ_BEFORE_WRITE();
#pragma unroll
        for (int j = 0; j < UB; j++) {
          const uint32_t w = R[j0 + j], wu = U[j];
          const uint32_t nw = (min(w, wu) & 0xFFFE0000u) | (wu & 0x1FFFFu) | (w & GRSR_W_ORBIT);
          chg |= (nw ^ w);
          R[j0 + j] = nw;
          W[tid + (j0 + j) * NT] = nw;
        }
      }
      if (!__syncthreads_or((chg & 0xFFFE0002u) != 0)) break;
    }
  }
  __syncthreads();

  // phase C: per black edge k = (2k, 2k+1): adjacency flag from the word; cycle id = min of the two
  // orbit labels; the edge whose left end is its own cycle id is the cycle's root -> count cycles
  // and the roots whose orbit saw no oriented grey edge.
  unsigned ncyc = 0, nun = 0;
  GRSR_PHASE_REL(2);
  for (int k = tid; k <= n; k += NT) {
    const unsigned a = 2u * k;
    const uint2 ww = *(const uint2 *)(W + a);
    const bool adj = (ww.x & GRSR_W_ADJ) != 0;
    const unsigned la = ww.x >> 17, lb = ww.y >> 17;
    if (kKeepOrbits) ((uint32_t *)wk.ORB)[k] = adj ? 0xFFFFFFFFu : (la | (lb << 16));
    const bool root = !adj && (min(la, lb) == a);
    ncyc += root;
    nun += (root && !(ww.x & GRSR_W_ORBIT));
  }
  unsigned long long tot =
      blk_sum_u64(((unsigned long long)nadj << 42) | ((unsigned long long)ncyc << 21) | nun, wk.red);
  Stats st;
  st.b = (n + 1) - (int)(tot >> 42);
  st.c = (int)((tot >> 21) & 0x1FFFFFu);
  st.h = 0; st.f = 0;
  if ((tot & 0x1FFFFFu) != 0 &&
      !unoriented_cycles_all_covered<TeamBlock>(wk, n, S, [&](unsigned v) -> unsigned {
        const uint32_t w = W[v];
        return ((w >> 17) << 2) | (w & 3u);
      })) {
    // some unoriented cycle is not obviously covered: the component analysis is needed.  Cycle
    // id per vertex and the orientation flag of every cycle first.
    GRSR_PHASE_REL(3);
    uint32_t *A2w = (uint32_t *)wk.A2;
    for (int k = tid; k <= n; k += NT) {
      const unsigned a = 2u * k;
      const uint32_t w0 = W[a], w1 = W[a + 1];
      const unsigned cy = (w0 & GRSR_W_ADJ) ? kNone : min(w0 >> 17, w1 >> 17);
      A2w[k] = cy | (cy << 16);
      wk.A1[a] = (uint16_t)((w0 >> 1) & 1u);
      // S1 of the component analysis: forest and reach of a cycle live in the two words of its
      // leftmost black edge (the jumping words of this edge have just been read by this thread)
      if (cy == a) { W[a] = a; W[a + 1] = 0; }
    }
    __syncthreads();
    unsigned hf = hurdles_slow_path<TeamBlock, uint16_t, kComponentsOnly>(wk, n);
    st.h = (int)(hf >> 1); st.f = (int)(hf & 1u);
    if (kComponentsOnly) st.h = 1;
  }
  GRSR_PHASE(0);
  st.d = st.b - st.c + st.h + st.f;
  return st;
}

// --------------------------------------------------------------------------------------------------
// dist_core_large: the same pipeline for graphs that do not fit shared memory (2n+2 > 32767 or the
// block's shared-memory budget): 32-bit vertex ids, 64-bit jumping words
//   [63:33] label | [32:2] next | [1] oriented edge seen | [0] adjacency
// in a global-memory (L2-resident) workspace, one block per pair, words re-read every round
// instead of living in registers.  Phases and results are those of dist_core.
// --------------------------------------------------------------------------------------------------
#define GRSR_L_MAKE(label, next, flags) \
  (((unsigned long long)(label) << 33) | ((unsigned long long)(next) << 2) | (unsigned long long)(flags))
#define GRSR_L_NEXT(w) ((unsigned)(((w) >> 2) & 0x7FFFFFFFull))
#define GRSR_L_LABEL(w) ((unsigned)((w) >> 33))

template <class TM, bool kKeepOrbits, class EndsFn>
__device__ inline Stats dist_core_large(WorkT<uint32_t> &wk, int n, EndsFn ends) {
  const int tid = TM::tid(), nt = TM::size(), lane = tid & 31, wid = tid >> 5;
  const int S = wk.S;
  constexpr unsigned kNone = none_of<uint32_t>();
  unsigned long long *W = (unsigned long long *)wk.W;

  unsigned nadj = 0;
  {
    const int nwarps = nt >> 5;
    const int per = ((((n + 1) + nwarps - 1) / nwarps) + 31) & ~31;
    const int t0 = wid * per;
    const int t1 = min(n + 1, t0 + per);
    unsigned carry = 0;
    if (t0 > 0 && t0 <= n) { unsigned l; ends(t0 - 1, l, carry); }
    for (int base = t0; base < t1; base += 32) {
      const int t = base + lane;
      unsigned y = (unsigned)S - 1u, r = 0;
      if (t < n) ends(t, y, r);
      unsigned x = __shfl_up_sync(kFull, r, 1);
      if (lane == 0) x = carry;
      carry = __shfl_sync(kFull, r, 31);
      if (t <= n) {
        const unsigned xb = x ^ 1u, yb = y ^ 1u;
        const unsigned fl = ((((x ^ y) & 1u) == 0u) ? GRSR_W_ORBIT : 0u) | ((xb == y) ? GRSR_W_ADJ : 0u);
        nadj += (xb == y);
        W[xb] = GRSR_L_MAKE(xb, y, fl);
        W[yb] = GRSR_L_MAKE(yb, x, fl);
      }
    }
  }
  TM::sync();

  {
    const int rounds = ceil_log2(n + 1);
    for (int r = 0; r < rounds; r++) {
      unsigned long long chg = 0;
      for (int vb = 0; vb < S; vb += 4 * nt) {
        const int v0 = vb + tid;
        unsigned long long w[4], u[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { int v = v0 + j * nt; w[j] = (v < S) ? W[v] : 0ull; }
#pragma unroll
        for (int j = 0; j < 4; j++) { int v = v0 + j * nt; u[j] = (v < S) ? W[GRSR_L_NEXT(w[j])] : 0ull; }
        GRSR_EMU_READ_BEFORE_WRITE_TM();
#pragma unroll
        for (int j = 0; j < 4; j++) {
          int v = v0 + j * nt;
          if (v < S) {
            const unsigned long long m = (w[j] < u[j]) ? w[j] : u[j];
            const unsigned long long nw = (m & 0xFFFFFFFE00000000ull) | (u[j] & 0x1FFFFFFFFull) | (w[j] & 2ull);
            if (nw != w[j]) { W[v] = nw; chg |= (nw ^ w[j]); }
          }
        }
      }
      if (!TM::any((chg & 0xFFFFFFFE00000002ull) != 0ull, wk)) break;
    }
  }
  TM::sync();

  unsigned ncyc = 0, nun = 0;
  for (int k = tid; k <= n; k += nt) {
    const unsigned a = 2u * k;
    const unsigned long long w0 = W[a], w1 = W[a + 1];
    const bool adj = (w0 & GRSR_W_ADJ) != 0;
    const unsigned la = GRSR_L_LABEL(w0), lb = GRSR_L_LABEL(w1);
    if (kKeepOrbits) { wk.ORB[a] = adj ? kNone : la; wk.ORB[a + 1] = adj ? kNone : lb; }
    const bool root = !adj && (min(la, lb) == a);
    ncyc += root;
    nun += (root && !(w0 & GRSR_W_ORBIT));
  }
  unsigned long long tot =
      TM::sum(((unsigned long long)nadj << 42) | ((unsigned long long)ncyc << 21) | nun, wk);
  Stats st;
  st.b = (n + 1) - (int)(tot >> 42);
  st.c = (int)((tot >> 21) & 0x1FFFFFu);
  st.h = 0; st.f = 0;
  if ((tot & 0x1FFFFFu) != 0 &&
      !unoriented_cycles_all_covered<TM>(wk, n, S, [&](unsigned v) -> unsigned {
        const unsigned long long w = W[v];
        return (GRSR_L_LABEL(w) << 2) | (unsigned)(w & 3ull);
      })) {
    for (int k = tid; k <= n; k += nt) {
      const unsigned a = 2u * k;
      const unsigned long long w0 = W[a], w1 = W[a + 1];
      const unsigned cy = (w0 & GRSR_W_ADJ) ? kNone : min(GRSR_L_LABEL(w0), GRSR_L_LABEL(w1));
      wk.A2[a] = cy; wk.A2[a + 1] = cy;
      wk.A1[a] = (unsigned)((w0 >> 1) & 1ull);
      if (cy == a) { wk.P[a] = a; wk.P[a + 1] = 0; }
    }
    TM::sync();
    unsigned hf = hurdles_slow_path<TM>(wk, n);
    st.h = (int)(hf >> 1); st.f = (int)(hf & 1u);
  }
  st.d = st.b - st.c + st.h + st.f;
  return st;
}

// global-memory workspace of the large build, per block
__host__ __device__ inline size_t large_work_bytes(int n, bool keep_orbits) {
  size_t sv = (size_t)((2 * n + 2 + 7) & ~7);
  size_t b = sv * 8 + sv * 4 /*P*/ + sv * 4 + sv * 4 + (keep_orbits ? sv * 4 : 0);
  b += (((size_t)num_vblocks(n) * 2 * 4) + 15) & ~(size_t)15;
  return (b + 255) & ~(size_t)255;
}

__device__ inline void large_work_carve(WorkT<uint32_t> &wk, unsigned char *base, int n, bool keep_orbits,
                                        unsigned long long *red_smem) {
  size_t sv = (size_t)((2 * n + 2 + 7) & ~7);
  wk.S = 2 * n + 2; wk.Sp = (int)sv; wk.NB = num_vblocks(n);
  wk.W = (void *)base;        base += sv * 8;
  wk.P = (uint32_t *)base;    base += sv * 4;
  wk.A1 = (uint32_t *)base;   base += sv * 4;
  wk.A2 = (uint32_t *)base;   base += sv * 4;
  wk.ORB = keep_orbits ? (uint32_t *)base : (uint32_t *)0;
  if (keep_orbits) base += sv * 4;
  wk.BLK = (uint32_t *)base;
  wk.red = red_smem;
  wk.xch = (unsigned long long *)0; wk.epoch = 0;
}

// Ends of a gene given its signed position q = +-(p+1) in pi: a forward gene at position p owns
// vertices 2p+1 (left) and 2p+2 (right); a backward one has them swapped.
__device__ inline void ends_from_signed_pos(int q, unsigned &L, unsigned &R) {
  const unsigned a = (unsigned)(q > 0 ? q : -q), neg = (q < 0) ? 1u : 0u;
  L = 2u * a - 1u + neg;
  R = 2u * a - neg;
}

// ends() for a pair of rows of the genome table: gamma streamed from global memory, pi staged in
// shared memory as tail[a-1] = the vertex where gene +a begins in pi (2p+1 if pi holds +a at
// position p, 2p+2 if it holds -a); the other end of the gene is the other vertex of {2p+1, 2p+2}.
struct EndsFromRows {
  const int32_t *gamma; const uint16_t *tail;
  __device__ inline void operator()(int t, unsigned &L, unsigned &R) const {
    const int g = gamma[t];
    const unsigned v = tail[(g > 0 ? g : -g) - 1];
    const unsigned h = ((v - 1u) ^ 1u) + 1u;
    L = (g > 0) ? v : h;
    R = (g > 0) ? h : v;
  }
};

// ---- kernels --------------------------------------------------------------------------------------

// Signed inverse of every genome: sinv[row][|g|-1] = sign(g) * (pos+1).
__global__ void signed_inverse_kernel(const int32_t *__restrict__ genomes, int32_t *__restrict__ sinv,
                                      long long total, int n) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    long long row = e / n; int pos = (int)(e - row * n);
    int g = genomes[e];
    int a = g > 0 ? g : -g;
    sinv[row * n + (a - 1)] = g > 0 ? (pos + 1) : -(pos + 1);
  }
}

__host__ __device__ inline size_t stage_bytes(int n) { return (((size_t)n * 2) + 15) & ~(size_t)15; }

// Launch shape of the one-pair-per-block kernels: threads per block (NT) and jumping words per
// thread (EPT) with NT * EPT >= 2n+2, both compile-time in the kernels.  About 16 words per thread
// keeps the per-round loop in registers while several blocks stay resident per SM -- the shape for
// throughput (many more pairs than resident blocks).  When all pairs of a call are resident at
// once the pair count is the only parallelism and latency per pair matters: then 8 words per
// thread (twice the threads per pair) is faster (target_ept = 8).  Only the combinations listed in
// GRSR_DISPATCH_SHAPE exist.
__host__ inline bool shape_exists(int t, int e) {
  if (t == 32) return e == 1 || e == 2 || e == 4 || e == 8 || e == 16;
  if (t == 64) return e == 8 || e == 16;
  if (t == 128 || t == 256) return e == 8 || e == 16 || e == 32;
  if (t == 512) return e == 8 || e == 16 || e == 32 || e == 64;
  if (t == 1024) return e == 10 || e == 12 || e == 16;
  return false;
}

__host__ inline void pick_shape(int n, int forced_threads, int *threads, int *ept, int target_ept = 16) {
  const int S = 2 * n + 2;
  int t = 32, e = 1;
  while (t < 512 && t * target_ept < S) t *= 2;
  while (e < 64 && t * e < S) e *= 2;
  if (!shape_exists(t, e)) {                 // e.g. (64, 4): fall back to the throughput shape
    t = 32; e = 1;
    while (t < 512 && t * 16 < S) t *= 2;
    while (e < 64 && t * e < S) e *= 2;
  }
  // 8192 < S <= 16384 (config 4: n = 5000): one pair's graph takes so much shared memory that at most
  // one or two blocks fit per SM whatever their size, so the block is made as large as it can be
  // (1024 threads: twice the warps to hide the shared-memory latency of every pass) with just enough
  // words per thread (10 / 12 / 16: the jumping array is not padded to the next power of two)
  if (S > 8192 && S <= 16384) { t = 1024; e = (S <= 10240) ? 10 : (S <= 12288 ? 12 : 16); }
  if (forced_threads > 0) {
    int fe = 1;
    while (fe < 64 && forced_threads * fe < S) fe *= 2;
    if (shape_exists(forced_threads, fe) && forced_threads * fe >= S) { t = forced_threads; e = fe; }
  }
  *threads = t; *ept = e;
}

#define GRSR_SHAPE_CASE(T, E, ...) \
  if (nt_ == T && ept_ == E) { constexpr int kNt = T; constexpr int kEpt = E; __VA_ARGS__; } else
#define GRSR_DISPATCH_SHAPE(threads, ept, ...)                                        \
  do {                                                                                \
    const int nt_ = (threads), ept_ = (ept);                                          \
    GRSR_SHAPE_CASE(32, 1, __VA_ARGS__) GRSR_SHAPE_CASE(32, 2, __VA_ARGS__)           \
    GRSR_SHAPE_CASE(32, 4, __VA_ARGS__) GRSR_SHAPE_CASE(32, 8, __VA_ARGS__)           \
    GRSR_SHAPE_CASE(32, 16, __VA_ARGS__) GRSR_SHAPE_CASE(64, 8, __VA_ARGS__)          \
    GRSR_SHAPE_CASE(64, 16, __VA_ARGS__)                                              \
    GRSR_SHAPE_CASE(128, 8, __VA_ARGS__) GRSR_SHAPE_CASE(128, 16, __VA_ARGS__)        \
    GRSR_SHAPE_CASE(128, 32, __VA_ARGS__) GRSR_SHAPE_CASE(256, 8, __VA_ARGS__)        \
    GRSR_SHAPE_CASE(256, 16, __VA_ARGS__) GRSR_SHAPE_CASE(256, 32, __VA_ARGS__)       \
    GRSR_SHAPE_CASE(512, 8, __VA_ARGS__) GRSR_SHAPE_CASE(512, 16, __VA_ARGS__)        \
    GRSR_SHAPE_CASE(512, 32, __VA_ARGS__) GRSR_SHAPE_CASE(512, 64, __VA_ARGS__)       \
    GRSR_SHAPE_CASE(1024, 10, __VA_ARGS__) GRSR_SHAPE_CASE(1024, 12, __VA_ARGS__)     \
    GRSR_SHAPE_CASE(1024, 16, __VA_ARGS__) {}                                         \
  } while (0)

// registers per thread are capped so that the shared-memory-limited number of blocks fits
#define GRSR_MIN_BLOCKS(NT, EPT) ((EPT) > 16 || (NT) > 512 ? 1 : ((NT) >= 512 ? 3 : ((NT) == 256 ? 6 : ((NT) == 128 ? 10 : 16))))

// Batched distance: pair p = (gamma row pairs[2p], pi row pairs[2p+1]) as in setinvmatrix
// (reference G/uniinvdist.c:80-90).  stride 1: out[p] = d; stride 5: out[5p..] = {d,b,c,h,f}.
// Block b owns the contiguous pairs [b*chunk, (b+1)*chunk); pi stays staged in shared memory while
// consecutive pairs share their pi row.
template <int NT, int EPT>
__global__ void __launch_bounds__(NT, GRSR_MIN_BLOCKS(NT, EPT))
dist_pairs_kernel(const int32_t *__restrict__ genomes, const int32_t *__restrict__ sinv, int n,
                  const int32_t *__restrict__ pairs, long long npairs, int32_t *__restrict__ out,
                  int stride) {
  GRSR_DYN_SMEM(smem_raw);
  Work wk;
  work_carve(wk, smem_raw, n, NT * EPT, false);
  GRSR_PHASE_INIT(); GRSR_PHASE_BASE(8);
  uint16_t *stage = (uint16_t *)(smem_raw + work_bytes(n, NT * EPT, false));
  const long long chunk = (npairs + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * chunk;
  const long long p1 = (p0 + chunk < npairs) ? p0 + chunk : npairs;
  int staged = -1;
  for (long long p = p0; p < p1; p++) {
    const int gi = pairs[2 * p], pj = pairs[2 * p + 1];
    if (pj != staged) {
      const int32_t *sp = sinv + (long long)pj * n;
      for (int i = threadIdx.x; i < n; i += NT) {
        const int s = sp[i];
        stage[i] = (uint16_t)(s > 0 ? 2 * s - 1 : -2 * s);   // vertex where gene +(i+1) begins in pi
      }
      staged = pj;
      __syncthreads();
    }
    EndsFromRows ends; ends.gamma = genomes + (long long)gi * n; ends.tail = stage;
    Stats st = dist_core<NT, EPT, false>(wk, n, ends);
    if (threadIdx.x == 0) put_result(out, stride, p, gi, pj, st.d, st.b, st.c, st.h, st.f);
  }
}

// ends() for the large build: gamma and the signed inverse of pi both read from global memory
struct EndsFromRowsGlobal {
  const int32_t *gamma; const int32_t *sinv_pi;
  __device__ inline void operator()(int t, unsigned &L, unsigned &R) const {
    const int g = gamma[t];
    const int s = sinv_pi[(g > 0 ? g : -g) - 1];
    ends_from_signed_pos(g > 0 ? s : -s, L, R);
  }
};

// Batched distance for num_genes beyond the shared-memory build: one block per pair, workspace
// scratch + blockIdx.x * large_work_bytes(n) in global memory.
__global__ void __launch_bounds__(1024)
dist_pairs_large_kernel(const int32_t *__restrict__ genomes, const int32_t *__restrict__ sinv, int n,
                        const int32_t *__restrict__ pairs, long long npairs, int32_t *__restrict__ out,
                        int stride, unsigned char *scratch) {
  __shared__ unsigned long long red[34];
  WorkT<uint32_t> wk;
  large_work_carve(wk, scratch + (size_t)blockIdx.x * large_work_bytes(n, false), n, false, red);
  for (long long p = blockIdx.x; p < npairs; p += gridDim.x) {
    EndsFromRowsGlobal ends;
    ends.gamma = genomes + (long long)pairs[2 * p] * n;
    ends.sinv_pi = sinv + (long long)pairs[2 * p + 1] * n;
    Stats st = dist_core_large<TeamBlock, false>(wk, n, ends);
    if (threadIdx.x == 0)
      put_result(out, stride, p, pairs[2 * p], pairs[2 * p + 1], st.d, st.b, st.c, st.h, st.f);
  }
}

} // namespace grsr
