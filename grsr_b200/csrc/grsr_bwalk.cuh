// grsr_bwalk.cuh -- block-level version of the walks of grsr_walk.cuh, for the kernels that keep ONE
// pair per thread block (the scenario engine): dist_core_bwalk computes b and c -- and, on request,
// the orbit label of every vertex -- with O(n) work instead of dist_core's O(n log n) pointer
// jumping over all vertices.  Unoriented cycles get dist_core's "covered by an oriented cycle" test;
// when that does not settle them the cycle ids go straight to hurdles_slow_path (components, hurdles,
// fortress).  Only when an orbit without a node is long does it report "not settled"; the caller then
// runs dist_core, which recomputes everything.  Same reference stages as dist_core phases A-C (double_perms / buildG_grey /
// classify_cycle_path / num_breakpoints, G/graph_edit.c:46-157,207-457, G/mcrdist.c:116-145,
// G/uniinvdist.c:124-142) and the early exit of G/graph_components.c:522-527.
//
// tau(u) = grey[u] ^ 1 ("follow the grey edge, cross the black edge").  Its orbits are, as vertex
// sets, the orbits of dist_core's sigma(v) = grey[v ^ 1] (tau = X sigma X with X = ^1, and the
// mirror image of a sigma-orbit is the other sigma-orbit of the same cycle), so the labels left in
// wk.ORB are the ones dist_core<keep orbits> leaves.
//
// Every vertex that is a multiple of 8 is a node.  A thread out of work takes the next node from a
// block-wide cursor; if the node is unclaimed it claims it and walks along tau, claiming every
// vertex it reaches, until it arrives at a claimed vertex -- always the start of a walk, because a
// segment can only be entered through its first vertex.  Starts and steps alternate between block
// barriers, so a claim is never raced.  One word per walk (segment minimum, walk it ran into,
// "crossed an oriented grey edge"), pointer jumping over the walk words, roots = walks whose own
// segment holds the orbit minimum; vertices never claimed lie on orbits without a node and are
// resolved by a direct walk from each of them.
#pragma once
#include "grsr_core.cuh"

namespace grsr {

#ifdef GRSR_BW_STATS
__device__ unsigned long long g_bw_stats[8];
#define GRSR_BW_COUNT(i) do { if (threadIdx.x == 0) atomicAdd(&g_bw_stats[i], 1ull); } while (0)
#else
#define GRSR_BW_COUNT(i) ((void)0)
#endif

static const unsigned kBwVis = 0x8000u;
static const int kBwShift = 3;               // node stride 8
static const int kBwGroup = 8;               // tau steps between two block barriers
static const int kBwMinGenes = 48;           // below this the aliased arrays would not fit
static const int kBwLeftoverBudget = 1024;   // steps a thread may spend on orbits without a node

// Arrays carved from the jumping-word region of the Work (4 * Sp >= 4 * S bytes, dead until
// dist_core's phase A rewrites it): H u16[S+1 -> 8], MAP u16[K], WS u16[K], SEG u16[K], NW u32[K],
// 2 counters.  With K = ceil(S / 8): 2S + 16 + 10K + 16 <= 3.25 S + 58 <= 4S for S >= 78.
struct BwMem {
  uint16_t *H;      // [S + 1] tau(u) | claimed bit; entry S = a claimed pad for threads without a walk
  uint16_t *MAP;    // [K] node -> walk id (0xFFFF: no walk starts there)
  uint16_t *WS;     // [K] walk id -> start vertex
  uint16_t *SEG;    // [K] walk id -> smallest vertex of its segment
  uint32_t *NW;     // [K] walk words: label:15 | next walk:15 | oriented:1 | unused:1
  unsigned *ctr;    // [0] node cursor, [1] walks started, [2] unoriented cycles listed, [3] spare
  int K;
};

__device__ inline void bw_carve(BwMem &m, void *region, int S) {
  unsigned char *b = (unsigned char *)region;
  m.K = (S + (1 << kBwShift) - 1) >> kBwShift;
  const size_t kb = (((size_t)m.K * 2) + 15) & ~(size_t)15;
  m.H = (uint16_t *)b;    b += (((size_t)(S + 1) * 2) + 15) & ~(size_t)15;
  m.MAP = (uint16_t *)b;  b += kb;
  m.WS = (uint16_t *)b;   b += kb;
  m.SEG = (uint16_t *)b;  b += kb;
  m.NW = (uint32_t *)b;   b += (size_t)m.K * 4;
  m.ctr = (unsigned *)b;
}

__host__ __device__ inline bool bw_fits(int n, int slots) {
  const int S = 2 * n + 2, K = (S + 7) >> 3;
  const size_t kb = (((size_t)K * 2) + 15) & ~(size_t)15;
  const size_t need = ((((size_t)(S + 1) * 2) + 15) & ~(size_t)15) + 3 * kb + (size_t)K * 4 + 16;
  return n >= kBwMinGenes && need <= (size_t)padded_size(n, slots) * 4;
}

// All threads of the block call; all receive the result.  settled == false: run dist_core instead
// (nothing this function wrote needs to be kept).  kKeepOrbits: wk.ORB as after dist_core<.., true>.
template <int NT, bool kKeepOrbits, class EndsFn>
__device__ inline Stats dist_core_bwalk(Work &wk, int n, EndsFn ends, bool &settled) {
  const int tid = threadIdx.x;
  const int S = wk.S;
  constexpr unsigned kNone = none_of<uint16_t>();
  BwMem m;
  bw_carve(m, wk.W, S);
  uint16_t *H = m.H;
  const int K = m.K;
  const unsigned mask = (1u << kBwShift) - 1u;

  // phase A: one thread per grey edge t = 0..n: x = right end of gene t-1 (vertex 0 first), y = left
  // end of gene t (vertex 2n+1 last); tau(x) = y ^ 1, tau(y) = x ^ 1; an adjacency (y == x ^ 1) makes
  // both ends fixed points of tau, born claimed.
  unsigned nadj = 0;
  for (int t = tid; t <= n; t += NT) {
    unsigned x = 0, y = (unsigned)S - 1u, l, r;
    if (t > 0) { ends(t - 1, l, r); x = r; }
    if (t < n) { ends(t, l, r); y = l; }
    const bool adj = (y == (x ^ 1u));
    nadj += adj ? 1u : 0u;
    const unsigned vis = adj ? kBwVis : 0u;
    H[x] = (uint16_t)((y ^ 1u) | vis);
    H[y] = (uint16_t)((x ^ 1u) | vis);
  }
  for (int i = tid; i < K; i += NT) m.MAP[i] = 0xFFFFu;
  if (tid == 0) { H[S] = (uint16_t)kBwVis; m.ctr[0] = 0; m.ctr[1] = 0; m.ctr[2] = 0; m.ctr[3] = 0; }
  __syncthreads();

  // walks
  {
    int myid = -1;
    bool ended = true;
    unsigned x = (unsigned)S, nx = (unsigned)S, mn = 0, orr = 0, e = kBwVis;
    for (;;) {
      if (ended && myid >= 0) {                       // close the walk that has ended
        m.NW[myid] = GRSR_W_MAKE(mn, m.MAP[nx >> kBwShift], (orr & 1u) ? GRSR_W_ORBIT : 0u);
        m.SEG[myid] = (uint16_t)mn;
        myid = -1; x = (unsigned)S; nx = (unsigned)S; orr = 0;
      }
      if (ended) {                                    // take nodes from the cursor until one is unclaimed
        for (;;) {
          const unsigned i = atomicAdd(&m.ctr[0], 1u);
          if (i >= (unsigned)K) break;
          const unsigned u0 = i << kBwShift;
          const unsigned e0 = H[u0];
          if (e0 & kBwVis) continue;
          myid = (int)atomicAdd(&m.ctr[1], 1u);
          H[u0] = (uint16_t)(e0 | kBwVis);
          wk.A2[u0] = (uint16_t)myid;
          m.MAP[i] = (uint16_t)myid;
          x = u0; nx = e0; mn = u0; orr = 0; ended = false;
          break;
        }
      }
      __syncthreads();
#pragma unroll
      for (int s = 0; s < kBwGroup; s++) {            // a step of an ended / idle thread changes nothing
        const unsigned u = nx;
        orr |= x ^ u;
        e = H[u];
        if (!(e & kBwVis)) { H[u] = (uint16_t)(e | kBwVis); wk.A2[u] = (uint16_t)myid; mn = min(mn, u); x = u; nx = e; }
      }
      ended = (e & kBwVis) != 0u;
      // go on while somebody walks or nodes are left (thread 0 reads the cursor before the barrier,
      // when nobody moves it)
      const int more = (!ended) || (tid == 0 && m.ctr[0] < (unsigned)K);
      if (!__syncthreads_or(more)) break;
    }
    if (myid >= 0) {
      m.NW[myid] = GRSR_W_MAKE(mn, m.MAP[nx >> kBwShift], (orr & 1u) ? GRSR_W_ORBIT : 0u);
      m.SEG[myid] = (uint16_t)mn;
    }
  }
  __syncthreads();
  const int W = (int)m.ctr[1];

  // jumping over the walk words (in place; a word is one 32-bit snapshot, as in dist_core phase B)
  {
    constexpr int kPer = 8;                            // K / NT <= 8 for every launch shape
    const int rounds = ceil_log2(W);
    const char *Nb = (const char *)m.NW;
    for (int r = 0; r < rounds; r++) {
      uint32_t w[kPer], u[kPer];
      uint32_t chg = 0;
#pragma unroll
      for (int j = 0; j < kPer; j++) {
        const int id = tid + j * NT;
        if (id < W) { w[j] = m.NW[id]; u[j] = *(const uint32_t *)(Nb + (w[j] & 0x1FFFCu)); }
      }
      GRSR_EMU_READ_BEFORE_WRITE();
#pragma unroll
      for (int j = 0; j < kPer; j++) {
        const int id = tid + j * NT;
        if (id < W) {
          const uint32_t nw = (min(w[j], u[j]) & 0xFFFE0000u) | (u[j] & 0x1FFFFu) | (w[j] & GRSR_W_ORBIT);
          if (nw != w[j]) { m.NW[id] = nw; chg |= (nw ^ w[j]); }
        }
      }
      if (!__syncthreads_or((chg & 0xFFFE0002u) != 0u)) break;
    }
  }
  __syncthreads();

  // roots of the walked orbits; orbit labels of the claimed vertices.  UQ (one entry per cycle at
  // most) lists the leftmost vertices of the unoriented cycles.
  // Per cycle (= even-labelled root) wk.A1[leftmost vertex] = "owns an oriented grey edge", the entry
  // state of hurdles_slow_path; the list uses the odd entries of A1 (UQ(k) = A1[2k+1]).  wk.ORB is
  // not touched unless kKeepOrbits: the scenario search keeps using G0's labels across trial evaluations.
  uint16_t *UQ = wk.A1 + 1;
  unsigned ncyc = 0, bad = 0;
  for (int id = tid; id < W; id += NT) {
    const uint32_t w = m.NW[id];
    const unsigned label = w >> 17;
    if (m.SEG[id] == label) {                          // this segment holds the orbit minimum
      if (!(label & 1u)) {                             // the orbit of a cycle's leftmost vertex
        ncyc++;
        wk.A1[label] = (uint16_t)((w >> 1) & 1u);
        if (!(w & GRSR_W_ORBIT)) UQ[2u * atomicAdd(&m.ctr[2], 1u)] = (uint16_t)label;   // unoriented cycle
      }
    }
  }
  // every claimed vertex carries the id of its walk in A2 (written as it was claimed): turn it into
  // the cycle id (A2, the entry state of hurdles_slow_path) and the orbit label (ORB) -- one
  // independent lookup per vertex; adjacency ends become NONE, unclaimed vertices are done below
  for (int v = tid; v < S; v += NT) {
    const unsigned e0 = H[v];
    if (!(e0 & kBwVis)) continue;
    if ((e0 & 0x7FFFu) == (unsigned)v) {
      wk.A2[v] = (uint16_t)kNone;
      if (kKeepOrbits) wk.ORB[v] = (uint16_t)kNone;
    } else {
      const unsigned label = m.NW[wk.A2[v]] >> 17;
      wk.A2[v] = (uint16_t)(label & ~1u);
      if (kKeepOrbits) wk.ORB[v] = (uint16_t)label;
    }
  }
  // orbits without a node: every unclaimed vertex walks until it meets a smaller one; the one that
  // gets round is the root (and labels its orbit)
  {
    int budget = kBwLeftoverBudget;
    for (int v = tid; v < S; v += NT) {
      const unsigned e0 = H[v];
      if (e0 & kBwVis) continue;
      unsigned x1 = e0, orr = (unsigned)v ^ e0;
      bool root = true;
      while (x1 != (unsigned)v) {
        if (x1 < (unsigned)v) { root = false; break; }
        if (--budget < 0) { root = false; bad = 1; break; }
        const unsigned nx1 = H[x1] & 0x7FFFu;
        orr |= x1 ^ nx1;
        x1 = nx1;
      }
      if (root) {
        if (!(v & 1)) {
          ncyc++;
          wk.A1[v] = (uint16_t)(orr & 1u);
          if (!(orr & 1u)) UQ[2u * atomicAdd(&m.ctr[2], 1u)] = (uint16_t)v;
        }
        unsigned u = (unsigned)v;
        do {
          wk.A2[u] = (uint16_t)(v & ~1);
          if (kKeepOrbits) wk.ORB[u] = (uint16_t)v;
          u = H[u] & 0x7FFFu;
        } while (u != (unsigned)v);
      }
    }
  }
  const unsigned long long tot =
      blk_sum_u64(((unsigned long long)nadj << 42) | ((unsigned long long)ncyc << 21) | bad, wk.red);
  Stats st;
  st.b = (n + 1) - (int)(tot >> 42);
  st.c = (int)((tot >> 21) & 0x1FFFFFu);
  st.h = 0; st.f = 0;
  st.d = st.b - st.c;
  settled = ((tot & 0x1FFFFFu) == 0);
  // Unoriented cycles: the "covered by an oriented cycle" test of dist_core
  // (unoriented_cycles_all_covered), one warp per cycle: the vertices right of its leftmost vertex q
  // up to the cycle's own next vertex lie inside its extent; one of them on an oriented cycle that
  // starts left of q puts q's component among the oriented ones.  All covered => h = f = 0.
  const int nq = (int)m.ctr[2];                       // complete: blk_sum_u64 went through barriers
  if (settled && nq > 0) {
    const int lane = tid & 31, wid = tid >> 5, nwarps = NT >> 5;
    int unresolved = 0;
    for (int idx = wid; idx < nq && !unresolved; idx += nwarps) {
      const unsigned qq = UQ[2 * idx];
      bool hit = false;
      for (int it = 0; it < 8; it++) {
        const unsigned v = qq + 2u + 32u * it + lane;
        int stt = 2;                                  // past the last vertex: stop
        if (v < (unsigned)S) {
          stt = 0;
          const unsigned cy = wk.A2[v];
          if (cy != kNone) {                          // not an adjacency end
            const unsigned ori = wk.A1[cy] & 1u;
            if (cy == qq) stt = 2;                    // the cycle's own next vertex: the probe ends here
            else if (cy < qq && ori) stt = 1;         // an oriented cycle that starts left of q
          }
        }
        const unsigned mh = __ballot_sync(kFull, stt == 1), ms = __ballot_sync(kFull, stt == 2);
        if (mh | ms) { hit = mh && (!ms || __ffs((int)mh) < __ffs((int)ms)); break; }
      }
      if (!hit) unresolved = 1;
    }
    const bool covered = !__syncthreads_or(unresolved);
    GRSR_BW_COUNT(covered ? 2 : 3);
    if (!covered) {
      // the component analysis is needed: the cycle id of every vertex (A2) and the flags in A1 are
      // what hurdles_slow_path starts from -- dist_core's phases A-C are not repeated
      __syncthreads();
      const unsigned hf = hurdles_slow_path<TeamBlock>(wk, n);
      st.h = (int)(hf >> 1); st.f = (int)(hf & 1u);
      st.d = st.b - st.c + st.h + st.f;
    }
  } else {
    GRSR_BW_COUNT(settled ? 0 : 1);
  }
  return st;
}

} // namespace grsr
