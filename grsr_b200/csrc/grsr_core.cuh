// grsr_core.cuh -- device code of the reversal-distance hot path (sm_100a).
//
// One thread block owns one permutation pair (gamma, pi) and keeps that pair's whole breakpoint
// graph in shared memory; many pairs are processed per launch (grid-stride over the pair list).
// What it reproduces, bit for bit, is the reference GRIMM 2.01 pipeline invdist_noncircular_G
// (reference G/uniinvdist.c:150-214):
//
//   reference stage (file:line)                           here
//   ---------------------------------------------------   -----------------------------------------
//   double_perms            G/graph_edit.c:46-157          phase 1 (signed relative permutation rho)
//   buildG_grey             G/mcrdist.c:116-145            phase 2 (grey[v] = where[perm[v]^1])
//   classify_cycle_path     G/graph_edit.c:207-457         phase 3 (pointer jumping) + phase 4
//   form_connected_comp.    G/graph_components.c:50-242    slow path S1-S5 (extent queries + union-find)
//   classify_connected_c.   G/graph_components.c:257-336   phase 4b / S6
//   calc_num_hurdles_and_f. G/graph_components.c:465-727   slow path S7-S9
//   num_breakpoints         G/uniinvdist.c:124-142         phase 2 (adjacency count)
//
// The sequential walks and the stack sweep of the reference are replaced by data-parallel
// formulations (see DESIGN.md section 3 for the equivalence arguments); the numbers produced are
// the same.  All arithmetic is 32-bit integer; vertex ids are stored as 16-bit in shared memory
// (2n+2 <= 65535).
//
// The file is also compiled by g++ against tests/emu/cuda_emu.hpp (a SIMT emulator used only by
// the CPU test-suite to debug kernel logic); nothing in the product selects that build.
#pragma once
#include <stdint.h>

#ifdef GRSR_EMU
#define GRSR_DYN_SMEM(name) unsigned char *name = emu::dyn_smem()
#else
#define GRSR_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace grsr {

typedef uint16_t vid_t;
static const unsigned kNone = 0xFFFFu;
static const unsigned kFull = 0xFFFFFFFFu;
static const int kMaxSmemGenes = 14000;   // 2n+2 vertices * ~10 B must fit 227 KB of shared memory

struct Stats { int d, b, c, h, f; };

// Shared-memory workspace of one pair.  S = 2n+2 vertices, padded to Sp (multiple of 8).
struct Work {
  uint32_t *W;              // [Sp] pointer-jumping words (label<<16 | next); later per-root scratch
  vid_t *A1;                // [Sp] perm -> grey partner -> per-root links -> compacted lists
  vid_t *A2;                // [Sp] value->position -> cycle id -> compacted lists
  vid_t *ORB;               // [Sp] sigma-orbit label per vertex (scenario kernels only)
  vid_t *BLK;               // [2*NB] per-32-vertex block: min cycle id, max cycle reach
  unsigned long long *red;  // [34] reduction scratch
  int S, Sp, NB;
};

__host__ __device__ inline int padded_size(int n) { return (2 * n + 2 + 7) & ~7; }
__host__ __device__ inline int num_vblocks(int n) { return (2 * n + 2 + 31) >> 5; }

__host__ __device__ inline size_t work_bytes(int n, bool keep_orbits) {
  size_t sp = (size_t)padded_size(n);
  size_t b = sp * 4 + sp * 2 + sp * 2 + (keep_orbits ? sp * 2 : 0);
  b += (((size_t)num_vblocks(n) * 2 * sizeof(vid_t)) + 15) & ~(size_t)15;
  b += 34 * sizeof(unsigned long long);
  return b;
}

__device__ inline void work_carve(Work &wk, unsigned char *base, int n, bool keep_orbits) {
  int sp = padded_size(n);
  wk.S = 2 * n + 2; wk.Sp = sp; wk.NB = num_vblocks(n);
  wk.W = (uint32_t *)base;  base += (size_t)sp * 4;
  wk.A1 = (vid_t *)base;    base += (size_t)sp * 2;
  wk.A2 = (vid_t *)base;    base += (size_t)sp * 2;
  wk.ORB = keep_orbits ? (vid_t *)base : (vid_t *)0;
  if (keep_orbits) base += (size_t)sp * 2;
  wk.BLK = (vid_t *)base;   base += (((size_t)wk.NB * 2 * sizeof(vid_t)) + 15) & ~(size_t)15;
  wk.red = (unsigned long long *)base;
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
template <class F>
__device__ inline int blk_compact(int N, F f, vid_t *out, unsigned long long *red) {
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
    if (ok) out[pos + __popc(m & ((1u << lane) - 1u))] = (vid_t)val;
    pos += __popc(m);
  }
  __syncthreads();
  return total;
}

__device__ inline int ceil_log2(int x) { int r = 0; while ((1 << r) < x) r++; return r; }

// Follow parent words up to the representative (union-find forest kept in W[root]).
__device__ inline unsigned uf_find(const uint32_t *W, unsigned x) {
  unsigned p;
  while ((p = W[x]) != x) x = p;
  return x;
}

// --------------------------------------------------------------------------------------------------
// Hurdles and fortress for a graph that has at least one unoriented cycle.
// Entry state: A2[v] = cycle id (leftmost vertex) or kNone for adjacency vertices;
//              W[r] = 1 if cycle r (r = its leftmost vertex, always even) owns an oriented grey
//              edge, else 0; W[odd] = 0.  A1 is free.
// Returns (h << 1) | f.
// --------------------------------------------------------------------------------------------------
__device__ inline unsigned hurdles_slow_path(Work &wk, int n) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  const int S = wk.S;
  uint32_t *W = wk.W; vid_t *A1 = wk.A1, *A2 = wk.A2, *BLK = wk.BLK;
  vid_t *bmin = BLK, *bmax = BLK + wk.NB;

  // S1: park the orientation flag in A1[r], start the union-find forest (W[r] = r) and the
  //     reach (W[r+1] = rightmost vertex of cycle r, filled by S2).
  for (int k = tid; k <= n; k += nt) {
    unsigned a = 2u * k;
    if (A2[a] == a) { A1[a] = (vid_t)W[a]; W[a] = a; W[a + 1] = 0; }
  }
  __syncthreads();

  // S2: reach of every cycle = max vertex carrying its id (reference: right[], G/graph_edit.c:392).
  //     Lanes hold consecutive vertices, so the highest lane of each equal-id group carries the max.
  for (int base = tid & ~31; base < S; base += nt) {
    int v = base + lane;
    unsigned c = (v < S) ? A2[v] : kNone;
    unsigned grp = __match_any_sync(kFull, c);
    if (c != kNone && lane == 31 - __clz((int)grp)) atomicMax(&W[c + 1], (uint32_t)v);
  }
  __syncthreads();

  // S3: per block of 32 vertices: smallest cycle id and largest reach seen in it.
  for (int B = tid; B < wk.NB; B += nt) {
    unsigned mn = kNone, mx = 0;
    for (int j = 0; j < 32; j++) {
      int v = (B << 5) + ((j + B) & 31);       // rotated start: conflict-free across lanes
      if (v < S) {
        unsigned c = A2[v];
        if (c != kNone) { mn = min(mn, c); mx = max(mx, W[c + 1]); }
      }
    }
    bmin[B] = (vid_t)mn; bmax[B] = (vid_t)mx;
  }
  __syncthreads();

  // S4: for every cycle P = (l, r): among the vertices strictly inside (l, r), the smallest cycle id
  //     (a cycle that starts left of P and reaches into it) and the largest reach (a cycle that
  //     reaches into P and ends right of it).  Those two neighbours are enough to recover the
  //     components of the reference's stack sweep (G/graph_components.c:101-138).
  for (int base = tid & ~31; base <= n; base += nt) {
    int k = base + lane;
    unsigned r0 = 0, R0 = 0; bool isroot = false;
    if (k <= n) { r0 = 2u * k; isroot = (A2[r0] == r0); if (isroot) R0 = W[r0 + 1]; }
    unsigned todo = __ballot_sync(kFull, isroot);
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
        unsigned lk1 = (mn < rr) ? mn : 0xFFFEu;
        unsigned lk2 = (mx > RR) ? (unsigned)A2[mx] : kNone;
        A1[rr] = (vid_t)((A1[rr] & 1u) | lk1);
        A1[rr + 1] = (vid_t)lk2;
      }
    }
  }
  __syncthreads();

  // S5: connected components of the cycles under those links: hook larger root under smaller,
  //     compress, repeat until no link joins two trees.
  for (;;) {
    int changed = 0;
    for (int k = tid; k <= n; k += nt) {
      unsigned r = 2u * k;
      if (A2[r] != r) continue;
      unsigned me = uf_find(W, r);
      unsigned l1 = A1[r] & 0xFFFEu, l2 = A1[r + 1];
      if (l1 != 0xFFFEu) {
        unsigned o = uf_find(W, l1);
        if (o != me) { atomicMin(&W[max(o, me)], min(o, me)); changed = 1; }
      }
      if (l2 != kNone) {
        me = uf_find(W, r);
        unsigned o = uf_find(W, l2);
        if (o != me) { atomicMin(&W[max(o, me)], min(o, me)); changed = 1; }
      }
    }
    if (!__syncthreads_or(changed)) break;
    for (int k = tid; k <= n; k += nt) {
      unsigned r = 2u * k;
      if (A2[r] == r) W[r] = uf_find(W, r);
    }
    __syncthreads();
  }

  // S6: a component is oriented iff one of its cycles is (G/graph_components.c:309-336).
  //     W[c+1] of the component root c becomes that flag (the reach is no longer needed).
  for (int k = tid; k <= n; k += nt) {
    unsigned r = 2u * k;
    if (A2[r] == r) W[r + 1] = 0;
  }
  __syncthreads();
  for (int k = tid; k <= n; k += nt) {
    unsigned r = 2u * k;
    if (A2[r] == r && (A1[r] & 1u)) W[W[r] + 1] = 1;
  }
  __syncthreads();

  // S7: the vertices of unoriented components, in vertex order, as a list of component ids
  //     (the filtered scan of G/graph_components.c:566-585).
  int m = blk_compact(S, [&](int v, unsigned &val) -> bool {
    unsigned c = A2[v];
    if (c == kNone) return false;
    unsigned comp = W[c];
    if (W[comp + 1]) return false;
    val = comp; return true;
  }, A1, wk.red);
  if (m == 0) return 0;

  // S8: runs of equal ids = the reference's "blocks"; keep one entry per run.
  int nb = blk_compact(m, [&](int t, unsigned &val) -> bool {
    unsigned c = A1[t];
    if (t > 0 && A1[t - 1] == c) return false;
    val = c; return true;
  }, A2, wk.red);
  const vid_t *run = A2;

  // S9: per component: number of runs (W[c]) and 1 + index of its last run (W[c+1]).
  for (int t = tid; t < nb; t += nt) { unsigned c = run[t]; W[c] = 0; W[c + 1] = 0; }
  __syncthreads();
  for (int t = tid; t < nb; t += nt) {
    unsigned c = run[t];
    atomicAdd(&W[c], 1u);
    atomicMax(&W[c + 1], (uint32_t)(t + 1));
  }
  __syncthreads();
  const unsigned first = run[0], last = run[nb - 1];
  // 1 + index of the last-but-one run of the final component (needed for its "right" neighbour)
  unsigned pl = 0;
  for (int t = tid; t < nb - 1; t += nt) if (run[t] == last) pl = max(pl, (unsigned)(t + 1));
  const unsigned prevlast = blk_max_u32(pl, wk.red);

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
  unsigned long long tot = blk_sum_u64(((unsigned long long)hcount << 32) | scount, wk.red);
  unsigned h = (unsigned)(tot >> 32), s = (unsigned)(tot & 0xFFFFFFFFu);
  unsigned f = (h == s && (s & 1u)) ? 1u : 0u;
  return (h << 1) | f;
}

// --------------------------------------------------------------------------------------------------
// dist_core: full invdist_noncircular_G for the pair described by rho, where
//   rho(i) in +-(1..n) = signed position in gamma of the gene at position i of pi
// (perm[2i+1], perm[2i+2] of the reference follow from it, G/graph_edit.c:69-116).
// kKeepOrbits additionally leaves in wk.ORB[v] the sigma-orbit label of every vertex
// (sigma(v) = grey[v^1]; kNone for the two ends of an adjacency); two breakpoint vertices x, y have
// an even graphdist (G/graph_edit.c:1165-1238) exactly when ORB[x] == ORB[y].
// All threads of the block must call; all receive the result.
// --------------------------------------------------------------------------------------------------
template <bool kKeepOrbits, class RhoFn>
__device__ inline Stats dist_core(Work &wk, int n, RhoFn rho) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int S = wk.S;
  uint32_t *W = wk.W; vid_t *A1 = wk.A1, *A2 = wk.A2;
  uint32_t *A1w = (uint32_t *)A1, *A2w = (uint32_t *)A2;

  // phase 1: doubled relative permutation (A1 = perm) and its inverse (A2 = where)
  for (int i = tid; i < n; i += nt) {
    int r = rho(i);
    unsigned a = (r > 0) ? (unsigned)(2 * r - 1) : (unsigned)(-2 * r);
    unsigned b = (r > 0) ? (unsigned)(2 * r) : (unsigned)(-2 * r - 1);
    unsigned p = 2u * i + 1u;
    A1[p] = (vid_t)a; A1[p + 1] = (vid_t)b;
    A2[a] = (vid_t)p; A2[b] = (vid_t)(p + 1);
  }
  if (tid == 0) { A1[0] = 0; A2[0] = 0; A1[S - 1] = (vid_t)(S - 1); A2[S - 1] = (vid_t)(S - 1); }
  __syncthreads();

  // phase 2: grey partner of v = where[perm[v] ^ 1]; an edge whose partner is its black neighbour
  // is an adjacency (G/mcrdist.c:125-141).  sigma(v) = grey[v ^ 1] seeds the jumping words.
  unsigned nadj = 0;
  for (int k = tid; k <= n; k += nt) {
    unsigned a = 2u * k;
    uint32_t pr = A1w[k];
    unsigned ga = A2[(pr & 0xFFFFu) ^ 1u], gb = A2[(pr >> 16) ^ 1u];
    if (ga == a + 1) {
      A1w[k] = 0xFFFFFFFFu;
      W[a] = (a << 16) | a; W[a + 1] = ((a + 1) << 16) | (a + 1);
      nadj++;
    } else {
      A1w[k] = ga | (gb << 16);
      W[a] = (a << 16) | gb; W[a + 1] = ((a + 1) << 16) | ga;
    }
  }
  __syncthreads();

  // phase 3: pointer jumping along sigma with min-label, in place.  Each word is one atomic
  // 32-bit (label, next) snapshot, so after round r every label covers >= 2^r successors; stop
  // after ceil(log2(n+1)) rounds or as soon as a whole round changed no label.
  const int rounds = ceil_log2(n + 1);
  for (int r = 0; r < rounds; r++) {
    int changed = 0;
    for (int v = tid; v < S; v += nt) {
      uint32_t w = W[v];
      unsigned u = w & 0xFFFFu;
      if (u == (unsigned)v) continue;
      uint32_t wu = W[u];
      unsigned lab = min(w >> 16, wu >> 16);
      uint32_t nw = (lab << 16) | (wu & 0xFFFFu);
      if (nw != w) { W[v] = nw; changed |= (lab != (w >> 16)); }
    }
    if (!__syncthreads_or(changed)) break;
  }

  // phase 4a: cycle id = min of the two orbit labels of a black edge's ends (kNone for an
  // adjacency); a black edge whose left end is its own cycle id is a cycle root -> count cycles.
  unsigned ncyc = 0;
  for (int k = tid; k <= n; k += nt) {
    unsigned a = 2u * k;
    unsigned la = W[a] >> 16, lb = W[a + 1] >> 16;
    const bool adj = (A1w[k] == 0xFFFFFFFFu);
    if (kKeepOrbits) ((uint32_t *)wk.ORB)[k] = adj ? 0xFFFFFFFFu : (la | (lb << 16));
    unsigned cy = adj ? kNone : min(la, lb);
    A2w[k] = cy | (cy << 16);
    ncyc += (cy == a);
    W[a] = 0; W[a + 1] = 0;
  }
  __syncthreads();

  // phase 4b: a grey edge (i, j) is oriented iff j - i is even (G/graph_components.c:324-329);
  // flag the cycle that owns one.
  for (int k = tid; k <= n; k += nt) {
    uint32_t g = A1w[k];
    if (g != 0xFFFFFFFFu) {
      unsigned ga = g & 0xFFFFu, gb = g >> 16;
      if (!(ga & 1u) || (gb & 1u)) W[A2[2 * k]] = 1;
    }
  }
  __syncthreads();

  // phase 4c: unoriented cycles left?  If none every component is oriented: h = f = 0.
  unsigned nun = 0;
  for (int k = tid; k <= n; k += nt) {
    unsigned a = 2u * k;
    if (A2[a] == a && W[a] == 0) nun++;
  }
  unsigned long long tot =
      blk_sum_u64(((unsigned long long)nadj << 42) | ((unsigned long long)ncyc << 21) | nun, wk.red);
  Stats st;
  st.b = (n + 1) - (int)(tot >> 42);
  st.c = (int)((tot >> 21) & 0x1FFFFFu);
  st.h = 0; st.f = 0;
  if ((tot & 0x1FFFFFu) != 0) {
    unsigned hf = hurdles_slow_path(wk, n);
    st.h = (int)(hf >> 1); st.f = (int)(hf & 1u);
  }
  st.d = st.b - st.c + st.h + st.f;
  return st;
}

// rho for a pair of rows of the genome table: pi in global memory, the signed inverse of gamma
// staged in shared memory (sinv[|g|-1] = +-(position+1) of gene |g| in gamma).
struct RhoFromRows {
  const int32_t *pi; const int32_t *sinv_sm;
  __device__ inline int operator()(int i) const {
    int g = pi[i];
    int s = sinv_sm[(g > 0 ? g : -g) - 1];
    return g > 0 ? s : -s;
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

// Batched distance: pair p = (gamma row pairs[2p], pi row pairs[2p+1]) as in setinvmatrix
// (reference G/uniinvdist.c:80-90).  stride 1: out[p] = d; stride 5: out[5p..] = {d,b,c,h,f}.
__global__ void __launch_bounds__(256)
dist_pairs_kernel(const int32_t *__restrict__ genomes, const int32_t *__restrict__ sinv, int n,
                  const int32_t *__restrict__ pairs, long long npairs, int32_t *__restrict__ out,
                  int stride) {
  GRSR_DYN_SMEM(smem_raw);
  Work wk;
  work_carve(wk, smem_raw, n, false);
  int32_t *stage = (int32_t *)wk.W;
  for (long long p = blockIdx.x; p < npairs; p += gridDim.x) {
    const int32_t *sg = sinv + (long long)pairs[2 * p] * n;
    const int32_t *pi = genomes + (long long)pairs[2 * p + 1] * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) stage[i] = sg[i];
    __syncthreads();
    RhoFromRows rho; rho.pi = pi; rho.sinv_sm = stage;
    Stats st = dist_core<false>(wk, n, rho);
    if (threadIdx.x == 0) {
      if (stride == 1) out[p] = st.d;
      else { int32_t *o = out + 5 * p; o[0] = st.d; o[1] = st.b; o[2] = st.c; o[3] = st.h; o[4] = st.f; }
    }
    __syncthreads();
  }
}

} // namespace grsr
