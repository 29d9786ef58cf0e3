// grsr_walk.cuh -- warp-per-pair reversal distance: the first tier of the batched distance path.
//
// dist_core (grsr_core.cuh) labels the cycles of the breakpoint graph by pointer jumping over all
// 2n+2 vertices: O(n log n) work per pair.  Here ONE WARP owns a pair and does O(n) work:
//
//   phase A  the same grey-edge construction as dist_core (reference double_perms + buildG_grey,
//            G/graph_edit.c:46-157, G/mcrdist.c:116-145; adjacencies counted for num_breakpoints,
//            G/uniinvdist.c:124-142), but stored as a 16-bit successor table
//              H[u+1] = tau(u) = grey[u] ^ 1      ("follow the grey edge, cross the black edge")
//            so that the two entries of a gene's two ends share one 32-bit word: one store per gene.
//            Each lane owns 8 consecutive genes of a 256-gene chunk (two 16-byte loads), so the
//            neighbouring genes' ends are in its own registers.  An alternating cycle of the graph is
//            two tau-orbits (one per direction); the orbit of its leftmost vertex q (always even) has
//            label q, the other one label q+1.
//   walks    one vertex in every 2^shift is a NODE (WalkGrid: a grid skewed so that 32 consecutive nodes
//            lie in 32 banks); free nodes are handed to the lanes that are out of work.  A walk CLAIMS
//            its start (bit 15 of the H entry), then follows tau, claiming every vertex it reaches,
//            until it arrives at a claimed vertex -- which is always the start of a walk (tau is a
//            bijection: a segment can only be entered through its first vertex).  It leaves one word:
//            smallest vertex of the segment, id of the walk it ran into, "an oriented grey edge was
//            crossed".  A lane whose walk ended takes the next free node, which splits whatever long
//            stretch that node lies on: the load balances itself.  Starts happen in a separate
//            section of the loop, so a claim is never raced.  This replaces the sequential walk of
//            classify_cycle_path (G/graph_edit.c:207-457) with ~1.5 * 2n/32 steps per lane and ~85
//            walks per pair.  A step carries the shared-memory ADDRESS of the entry under examination
//            (7 instructions; the dependent chain is load -> test || add -> load).  Measured and
//            rejected: claims on even vertices only -- fewer store wavefronts,
//            but the second predicate costs more than they save.
//   jumping  pointer jumping over the walk words only (registers, like dist_core's phase B); the
//            walk whose own segment holds the orbit minimum is the orbit's root.
//   rest     EVEN vertices never claimed lie on orbits without a node (short cycles): the lane that finds
//            one in its part of H walks its orbit until it meets a smaller vertex; the one that gets
//            round is the root.  (Odd ones are not tried: an orbit with an odd minimum is the twin of the
//            orbit that carries the cycle's label, and counts nothing.)
//   result   c = number of roots with an even label; when no root is unoriented every component is
//            oriented and d = b - c (the early exit of G/graph_components.c:522-527).  Unoriented
//            cycles first get the cheap "covered by an oriented cycle" test of dist_core; a pair it
//            does not settle (hurdles possible), or whose leftover orbits are long, is DEFERRED: its
//            index is appended to a list that the block-per-pair kernel (dist_deferred_kernel, the
//            second tier, with the full component / hurdle / fortress analysis) processes next.
//
// Results are those of dist_core, bit for bit; only which tier computes them differs.
#pragma once
#include "grsr_core.cuh"
#include <type_traits>

#ifdef GRSR_EMU
#define GRSR_EMU_READ_BEFORE_WRITE_WARP() __syncwarp()
#else
#define GRSR_EMU_READ_BEFORE_WRITE_WARP() ((void)0)
#endif

// Emulator-only diagnostics (scripts/walk_occupancy.py): how many lanes take each tau step, how many
// start sections a pair needs.  Never defined in a library build.
#ifdef GRSR_WALK_STATS
extern "C" unsigned long long g_walk_stats[64];
#define GRSR_WALK_STAT(i, v) (g_walk_stats[i] += (v))
#else
#define GRSR_WALK_STAT(i, v) ((void)0)
#endif

namespace grsr {

static const unsigned kVis = 0x8000u;        // H entry: vertex claimed by a walk (or an adjacency end)
static const int kWalkMaxNodes = 256;        // nodes (possible walk starts) per pair
static const int kWalkMaxWalks = 128;        // walks per pair (4 words per lane in the jumping phase; ids fit a byte)
#ifndef GRSR_WALK_GROUP
#define GRSR_WALK_GROUP 12
#endif
static const int kWalkGroup = GRSR_WALK_GROUP;          // tau steps between two votes (same-box A/B: 6 / 8 / 12 / 16 / 24 -> 4.68 / 4.64 / 4.59 / 4.60 / 4.69 ms)
static const int kWalkIdle = 8;              // default: lanes that must be out of work before new walks start
static const int kWalkGuide4 = 6;            // guided self-scheduling: a claim takes remaining / (kWalkGuide4 / 4 x takers)
static const int kWalkMaxUnoriented = 30;    // unoriented cycles probed per pair before deferring
static const int kWalkLeftoverBudget = 1536; // steps a lane may spend on orbits without a node

// Node stride 2^shift for n genes: the finest one with at most kWalkMaxNodes nodes.
__host__ __device__ inline int walk_shift(int n) {
  const int S = 2 * n + 2;
  int sh = 1;
  while (((S + (1 << sh) - 1) >> sh) > kWalkMaxNodes) sh++;
  return sh;
}

// Node k sits at vertex (k << shift) + skew(k), skew even and below the stride, chosen so that the H
// entries of any 32 consecutive nodes lie in 32 different shared-memory banks (the warp examines 32
// nodes per load; on the plain grid they would share 32 / 2^(shift-1) banks).  Entry of vertex x: 16-bit
// index x + 1, i.e. word x / 2 for even x; node stride = 2^(shift-1) words.
struct WalkGrid {
  int shift, s1;
  unsigned m1;
  __device__ inline explicit WalkGrid(int sh) : shift(sh), s1(sh < 6 ? 6 - sh : 0) {
    const unsigned words = 1u << (sh - 1);
    m1 = (words < 32u ? words : 32u) - 1u;
  }
  __device__ inline unsigned vertex(unsigned k) const { return (k << shift) + 2u * ((k >> s1) & m1); }
  __device__ inline bool is_node(unsigned x) const { return x == vertex(x >> shift); }
};

// kernel argument: node stride and start threshold in one int
__host__ __device__ inline int walk_shape_arg(int shift, int idle_min) { return shift | (idle_min << 8); }

// 32-bit words of H: n+2 used (entries 0 .. 2n+2), padded so that 32 lanes x 16-byte loads cover it in whole rounds
__host__ __device__ inline int walk_hwords(int n) { return (n + 2 + 127) & ~127; }

// shared memory of one warp: tail u16[n+1] | H u16[2 * hwords] | walk words u32[128] | MAP u8[256] | misc
__host__ __device__ inline size_t walk_tail_bytes(int n) { return (((size_t)(n + 1) * 2) + 15) & ~(size_t)15; }
__host__ __device__ inline size_t walk_private_bytes(int n) {
  return (size_t)walk_hwords(n) * 4 + (size_t)kWalkMaxWalks * 4 + (size_t)kWalkMaxNodes + 128;
}
__host__ __device__ inline size_t walk_warp_bytes(int n) { return walk_tail_bytes(n) + walk_private_bytes(n); }

// The cycle a vertex lies on (leftmost vertex, i.e. orbit label with bit 0 cleared) and whether that
// cycle owns an oriented grey edge: walk to the next start of a walk and read its converged word,
// or get round an orbit that has none.
__device__ inline unsigned walk_cycle_of(const uint16_t *H, const uint32_t *NW, const uint8_t *MAP,
                                         unsigned v, const WalkGrid &grid, unsigned &oriented) {
  unsigned x = v, mn = v, orr = 0;
  for (;;) {
    if (!(x & 1u) && grid.is_node(x)) {
      const unsigned id = MAP[x >> grid.shift];
      if (id != 0xFFu) {
        const uint32_t w = NW[id];
        oriented = ((orr & 1u) | ((w >> 1) & 1u));
        return (w >> 17) & ~1u;
      }
    }
    const unsigned nx = H[x + 1] & 0x7FFFu;
    orr |= x ^ nx;
    x = nx;
    if (x == v) { oriented = orr & 1u; return mn & ~1u; }
    mn = min(mn, x);
  }
}

// Ends of gamma's gene g in pi (vertices): tail[a-1] = where gene +a begins, MINUS ONE (2p if pi holds
// +a at position p, else 2p+1), so that the sign of g is one XOR away.  Positions past the last gene
// hold the sentinel gene n+1, whose tail entry is the frame vertex 2n+1 (its "left end").
__device__ inline void walk_ends(const uint16_t *tail, int g, unsigned &L, unsigned &R) {
  const unsigned t = (unsigned)tail[(g > 0 ? g : -g) - 1] ^ ((unsigned)g >> 31);
  L = t + 1u;
  R = (t ^ 1u) + 1u;
}
// the staged form of one entry of the signed inverse row
__device__ inline uint16_t walk_tail_entry(int s) { return (uint16_t)(s > 0 ? 2 * s - 2 : -2 * s - 1); }

// (hi16 << 16) | lo16 of two 16-bit values
#ifdef GRSR_EMU
__device__ inline unsigned walk_pack(unsigned lo16, unsigned hi16) { return (lo16 & 0xFFFFu) | (hi16 << 16); }
__device__ inline unsigned walk_rotl(unsigned w, unsigned by) { by &= 31u; return by ? (w << by) | (w >> (32u - by)) : w; }
#else
__device__ inline unsigned walk_pack(unsigned lo16, unsigned hi16) { return __byte_perm(lo16, hi16, 0x5410); }
__device__ inline unsigned walk_rotl(unsigned w, unsigned by) { return __funnelshift_l(w, w, by); }   // by mod 32
#endif

// The tau steps carry the shared-memory ADDRESS of the entry under examination instead of the vertex:
// the claim goes to the address the entry was loaded from, the next address is one add away, and both
// the segment minimum (addresses grow with the vertex) and the orientation parity (bit 1 of the address
// = bit 0 of the vertex, H being 4-byte aligned) can be kept on addresses.
#ifdef GRSR_EMU
typedef uintptr_t walk_addr_t;
__device__ inline walk_addr_t walk_addr_of(const void *p) { return (uintptr_t)p; }
__device__ inline unsigned walk_lds16(walk_addr_t a) { return *(const uint16_t *)a; }
__device__ inline void walk_sts16(walk_addr_t a, unsigned v) { *(uint16_t *)a = (uint16_t)v; }
__device__ inline void walk_sts32(walk_addr_t a, unsigned v) { *(uint32_t *)a = v; }
#else
typedef unsigned walk_addr_t;
__device__ inline walk_addr_t walk_addr_of(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ inline unsigned walk_lds16(walk_addr_t a) {
  unsigned v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ inline void walk_sts16(walk_addr_t a, unsigned v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ inline void walk_sts32(walk_addr_t a, unsigned v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
#endif

// Genes t0 .. t0+7 of a row (n+1 past the end): two 16-byte loads when the row allows it.
__device__ inline void walk_load8(const int32_t *gamma, int t0, int n, bool vec, int (&G)[8]) {
#ifndef GRSR_EMU
  if (vec && t0 + 8 <= n) {
    const int4 a = __ldg((const int4 *)(gamma + t0)), b = __ldg((const int4 *)(gamma + t0 + 4));
    G[0] = a.x; G[1] = a.y; G[2] = a.z; G[3] = a.w; G[4] = b.x; G[5] = b.y; G[6] = b.z; G[7] = b.w;
    return;
  }
#else
  (void)vec;
#endif
#pragma unroll
  for (int k = 0; k < 8; k++) G[k] = (t0 + k < n) ? gamma[t0 + k] : n + 1;
}

// Pointer jumping over the first 32 * NJ walk words (ids >= W are words that point at themselves),
// then the roots: a walk whose own segment holds the minimum of its orbit.  Counts the cycles (roots
// with an even label) and lists the unoriented ones in UL.
template <int NJ>
__device__ inline void walk_jump(uint32_t *NW, uint32_t *UL, int W, int lane, unsigned &ncyc, unsigned &nun) {
  uint32_t Rg[NJ], seg[NJ];
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    const int id = lane + 32 * j;
    if (id < W) Rg[j] = NW[id];
    else { Rg[j] = GRSR_W_MAKE(0x7FFFu, id, GRSR_W_ADJ); NW[id] = Rg[j]; }
    seg[j] = Rg[j] >> 17;
  }
  __syncwarp();
  const int rounds = ceil_log2(W);
  const char *Nb = (const char *)NW;
  for (int r = 0; r < rounds; r++) {
    uint32_t chg = 0, U[NJ];
#pragma unroll
    for (int j = 0; j < NJ; j++) U[j] = *(const uint32_t *)(Nb + (Rg[j] & 0x1FFFCu));
    GRSR_EMU_READ_BEFORE_WRITE_WARP();
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      const uint32_t w = Rg[j], wu = U[j];
      const uint32_t nw = (min(w, wu) & 0xFFFE0000u) | (wu & 0x1FFFFu) | (w & GRSR_W_ORBIT);
      chg |= (nw ^ w);
      Rg[j] = nw;
      NW[lane + 32 * j] = nw;
    }
    __syncwarp();
    if (!__any_sync(kFull, (chg & 0xFFFE0002u) != 0u)) break;
  }
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    const uint32_t w = Rg[j];
    const unsigned label = w >> 17;
    if (!(w & GRSR_W_ADJ) && seg[j] == label && !(label & 1u)) {   // root of the orbit of a cycle's leftmost vertex
      ncyc++;
      if (!(w & GRSR_W_ORBIT)) {
        nun++;
        const unsigned k = atomicAdd(&UL[0], 1u);
        if (k < (unsigned)kWalkMaxUnoriented) UL[1 + k] = label;
      }
    }
  }
}

// Per-warp working set in shared memory.
struct WalkMem {
  const uint16_t *tail;     // [n+1] staged pi row (vertex where gene +a begins) + the sentinel entry
  uint32_t *Hw;             // successor table, 32-bit view ([walk_hwords])
  uint32_t *NW;             // [kWalkMaxWalks] walk words
  uint8_t *MAP;             // [kWalkMaxNodes] node -> walk id (0xFF: no walk started there)
  uint32_t *UL;             // [32] unoriented-cycle list and counters
};

__device__ inline void walk_carve_private(WalkMem &wm, unsigned char *base, int n) {
  wm.Hw = (uint32_t *)base;   base += (size_t)walk_hwords(n) * 4;
  wm.NW = (uint32_t *)base;   base += (size_t)kWalkMaxWalks * 4;
  wm.MAP = (uint8_t *)base;   base += (size_t)kWalkMaxNodes;
  wm.UL = (uint32_t *)base;
}

// pads of H (never rewritten): entry 0, entry S+1 and the words up to the 4-word multiple
__device__ inline void walk_init_private(WalkMem &wm, int n, int lane) {
  const int hwords = walk_hwords(n);
  for (int w = (n + 1) + lane; w < hwords; w += 32) wm.Hw[w] = 0x80008000u;
  if (lane == 0) ((uint16_t *)wm.Hw)[0] = (uint16_t)kVis;
}

// One pair by one warp (all 32 lanes call): gamma = row gi of the genome table, pi = the row staged
// in wm.tail.  stride 1: out[p] = d; stride 5: {d,b,c,h,f}; a deferred pair is appended to defer_list.
__device__ inline bool walk_pair(const WalkMem &wm, const int32_t *__restrict__ genomes, int n, int gi,
                                 int pj, long long p, int32_t *__restrict__ out, int stride, int shape,
                                 int32_t *defer_list, unsigned *defer_count) {
  const int shift = shape & 0xFF, idle_min = (shape >> 8) & 0xFF;   // node stride 2^shift; lanes out of work before new walks start
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int S = 2 * n + 2;
  const int hwords = walk_hwords(n);
  const uint16_t *tail = wm.tail;
  uint32_t *Hw = wm.Hw, *NW = wm.NW, *UL = wm.UL;
  uint8_t *MAP = wm.MAP;
  uint16_t *H = (uint16_t *)Hw;
  const WalkGrid grid(shift);
  const int K = (S + (1 << shift) - 1) >> shift;     // nodes 0 .. K-1 (the last one may fall past the last vertex)
  const bool vec = ((n & 3) == 0) && ((((size_t)genomes) & 15) == 0);

  if (lane == 0) UL[0] = 0;
  {
    uint32_t *M32 = (uint32_t *)MAP;                  // no walk starts anywhere yet
    for (int j = lane; j < (K + 3) / 4; j += 32) M32[j] = 0xFFFFFFFFu;
  }
  __syncwarp();

  // ---- phase A: lane owns genes cb + 8*lane .. +7 of gamma; L/R = left/right end in pi ----------
  const int32_t *gamma = genomes + (long long)gi * n;
  unsigned nadj = 0;
  {
    const walk_addr_t hw1 = walk_addr_of(Hw) + 1u;
    int G[8], Gn[8];
    walk_load8(gamma, 8 * lane, n, vec, G);
    unsigned carry = 0;                               // right end of the gene before this chunk (vertex 0 first)
    for (int cb = 0; cb < n; cb += 256) {
      const int tb = cb + 8 * lane;
      walk_load8(gamma, tb + 256, n, vec, Gn);
      unsigned L[8], R[8], Lx = 0, Rx;
#pragma unroll
      for (int k = 0; k < 8; k++) walk_ends(tail, G[k], L[k], R[k]);
      if (lane == 0) walk_ends(tail, Gn[0], Lx, Rx);    // the first gene of the next chunk: lane 31's right neighbour
      unsigned pr = __shfl_up_sync(kFull, R[7], 1);
      if (lane == 0) pr = carry;
      const unsigned nl0 = __shfl_sync(kFull, Lx, 0);
      unsigned nl = __shfl_down_sync(kFull, L[0], 1);
      if (lane == 31) nl = nl0;
      carry = __shfl_sync(kFull, R[7], 31);
      // the chunk's genes; a chunk that lies wholly inside the row has no sentinel genes to mask
      auto genes = [&](auto all_real) {
        constexpr bool kAllReal = decltype(all_real)::value;
        // m[k] = kVis when grey edge (gene k-1, gene k) of this lane's 8 genes is an adjacency: R_{k-1} ^ 1 == L_k
        unsigned m[9];
#pragma unroll
        for (int k = 0; k < 9; k++) {
          const unsigned Rp = (k == 0) ? pr : R[(k + 7) & 7], Lk = (k == 8) ? nl : L[k & 7];
          m[k] = ((Rp ^ Lk ^ 1u) == 0u) ? kVis : 0u;
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const unsigned Lt = L[k], Rt = R[k];
          const unsigned ea = (((k == 0) ? pr : R[(k + 7) & 7]) ^ 1u) | m[k];       // tau(L_t) = R_{t-1} ^ 1
          const unsigned eb = (((k == 7) ? nl : L[(k + 1) & 7]) ^ 1u) | m[k + 1];   // tau(R_t) = L_{t+1} ^ 1
          const bool real = kAllReal || (tb + k < n);
          nadj += real ? m[k] : 0u;                                     // in units of kVis
          // word p+1 = (entry of 2p+1, entry of 2p+2): L_t's entry is the low half iff L_t is odd
          const unsigned hl = walk_pack(eb, ea);
          const unsigned word = walk_rotl(hl, Lt << 4);
          if (real) walk_sts32(hw1 + Lt + Rt, word);                    // L_t + R_t + 1 = 4(p+1): word p+1
        }
      };
      if (cb + 256 <= n) genes(std::true_type()); else genes(std::false_type());
#pragma unroll
      for (int k = 0; k < 8; k++) G[k] = Gn[k];
    }
    if (lane == 0) {                                  // the two frame vertices and grey edge n
      unsigned L0, R0, Ll, Rl;
      walk_ends(tail, gamma[0], L0, R0);
      walk_ends(tail, gamma[n - 1], Ll, Rl);
      H[1] = (uint16_t)((L0 ^ 1u) | ((L0 == 1u) ? kVis : 0u));         // tau(0) = L_0 ^ 1
      const bool adjn = ((Rl ^ 1u) == (unsigned)S - 1u);
      H[S] = (uint16_t)((Rl ^ 1u) | (adjn ? kVis : 0u));              // tau(2n+1) = R_{n-1} ^ 1
      nadj += adjn ? kVis : 0u;
    }
  }
  __syncwarp();

  // ---- walks -------------------------------------------------------------------------------------
  // Lane state: `at` = address of the entry under examination (the vertex the walk steps to next), e =
  // that entry.  A lane without a walk looks at the pad entry H[0] (claimed): a step from there, like a
  // step of a walk that has ended, changes nothing, so the steps need no "am I walking" test.
  int W = 0;                                          // walks started so far (warp-uniform)
  {
    // Free nodes are handed out in node order from one cursor (pc) by the whole warp: 32 nodes are
    // examined per shared-memory load, the r-th lane out of work takes the r-th free one.
    int pc = 0, myid = -1;                            // pc: first node not examined yet (warp-uniform)
    bool ended = true;
    const walk_addr_t h0 = walk_addr_of(H), h1 = h0 + 2u;   // entry of vertex u: h1 + 2u
    walk_addr_t at = h0, mn = 0, org = 0, orr = 0;    // org: the walk's first vertex; orr: OR of (vertex ^ first vertex)
    unsigned e = kVis;
    uint32_t *SCR = UL;                               // 32 words of scratch (UL is not in use yet)
    for (;;) {
      // start section: close the walks that have ended, hand free nodes to the lanes out of work
      if (ended && myid >= 0) {
        const unsigned hit = (unsigned)((at - h0) >> 1) - 1u;          // the claimed vertex the walk ran into
        NW[myid] = GRSR_W_MAKE((unsigned)((mn - h0) >> 1) - 1u, MAP[hit >> shift], (orr & 2u) ? GRSR_W_ORBIT : 0u);
        myid = -1;
        at = h0;
      }
      const unsigned needm = __ballot_sync(kFull, ended);
      const int myrank = __popc(needm & lt_mask);
      int need = min(__popc(needm), kWalkMaxWalks - W), served = 0;
      while (need > served && pc < K) {
        const int node = pc + lane;
        const unsigned xn = grid.vertex((unsigned)node);
        const unsigned en = (node < K && xn < (unsigned)S) ? (unsigned)H[xn + 1] : kVis;
        const bool isfree = !(en & kVis);
        const unsigned fm = __ballot_sync(kFull, isfree);
        const int nf = __popc(fm);
        const int take = min(need - served, nf);
        if (take > 0) {
          const int fr = __popc(fm & lt_mask);
          if (isfree && fr < take) SCR[fr] = en | ((unsigned)lane << 16);
          __syncwarp();
          const int slot = myrank - served;
          if (ended && slot >= 0 && slot < take) {
            const unsigned v = SCR[slot];
            const int mynode = pc + (int)(v >> 16);
            const unsigned x = grid.vertex((unsigned)mynode), e0 = v & 0xFFFFu;
            myid = W + myrank;
            H[x + 1] = (uint16_t)(e0 | kVis);
            MAP[mynode] = (uint8_t)myid;
            mn = org = h1 + 2u * x;
            at = h1 + 2u * e0;
            orr = org ^ at;
            ended = false;
          }
          const int last = (int)(SCR[take - 1] >> 16);  // lane of the last free node handed out
          __syncwarp();
          pc += (nf > take) ? last + 1 : 32;
          served += take;
        } else {
          pc += 32;
        }
      }
      W += served;
      if (lane == 0) GRSR_WALK_STAT(44, 1);
      __syncwarp();
      if (!__any_sync(kFull, !ended)) break;
      const bool pool = (pc < K) && (W < kWalkMaxWalks);  // lanes out of work can still be served
      e = walk_lds16(at);                             // after every start of this section has claimed its node
      for (;;) {
#pragma unroll
        for (int s = 0; s < kWalkGroup; s++) {
#ifdef GRSR_WALK_STATS
          { const int act = __popc(__ballot_sync(kFull, !(e & kVis))); if (lane == 0) { GRSR_WALK_STAT(act, 1); GRSR_WALK_STAT(40 + (pool ? 0 : 1), 1); GRSR_WALK_STAT(42 + (pool ? 0 : 1), act); } }
#endif
          // unclaimed: claim it and examine tau of it.  The next address does not wait for the test, and
          // the load the next step depends on is issued before the claim.  (One level of predication:
          // a nested test makes the compiler branch, and a branch per step costs 8 %.)
          const walk_addr_t nx = h1 + 2u * e;
          if (!(e & kVis)) {
            const unsigned claimed = e | kVis;
            const walk_addr_t cur = at;
            mn = at < mn ? at : mn;
            at = nx;
            e = walk_lds16(at);
            walk_sts16(cur, claimed);
            orr |= at ^ org;
          }
        }
        ended = (e & kVis) != 0u;
        const unsigned walking = __ballot_sync(kFull, !ended);
        if (walking == 0u || (pool && 32 - __popc(walking) >= idle_min)) break;
      }
      __syncwarp();                                   // nobody examines nodes while others still claim
    }
    if (lane == 0) { UL[0] = 0; GRSR_WALK_STAT(45, 1); GRSR_WALK_STAT(46, W); }   // the scratch becomes the unoriented-cycle list
  }
  __syncwarp();

  // ---- jumping over the walk words -----------------------------------------------------------------
  unsigned ncyc = 0, nun = 0;
  walk_jump<4>(NW, UL, W, lane, ncyc, nun);

  // ---- orbits without a node: the scan of H finds the unclaimed vertices, the lane that finds one ----
  // walks its orbit at once (H is only read from here on): until it meets a smaller vertex, or gets
  // round -- then it is the root of that orbit.
  int defer = 0;
  {
    const uint4 *H4 = (const uint4 *)Hw;
    const int hq = hwords >> 2;
    int budget = kWalkLeftoverBudget;
    for (int q = lane; q < hq; q += 32) {                // hq is a multiple of 32: whole rounds
      const uint4 x = H4[q];
      const bool some = ((x.x & x.y & x.z & x.w) & 0x80000000u) == 0u;   // an even vertex (high halves) unclaimed
      if (__any_sync(kFull, some)) {
        if (some) {
          // bit k of um <-> entry 8q+k (vertex 8q+k-1) unclaimed.  Only EVEN vertices are tried: the
          // orbit of a cycle's leftmost vertex has an even minimum, its twin an odd one that counts nothing.
          const unsigned y0 = ~x.x & 0x80000000u, y1 = ~x.y & 0x80000000u;
          const unsigned y2 = ~x.z & 0x80000000u, y3 = ~x.w & 0x80000000u;
          unsigned um = (y0 >> 30) | (y1 >> 28) | (y2 >> 26) | (y3 >> 24);
          while (um) {
            const int k = __ffs((int)um) - 1;
            um &= um - 1u;
            const unsigned u0 = (unsigned)(8 * q + k) - 1u;       // vertex of entry 8q+k
            unsigned x1 = H[u0 + 1] & 0x7FFFu, orr = u0 ^ x1;
            bool root = true;
            while (x1 != u0) {
              if (x1 < u0) { root = false; break; }                // a smaller vertex owns this orbit
              if (--budget < 0) { root = false; defer = 1; break; }
              const unsigned nx = H[x1 + 1] & 0x7FFFu;
              orr |= x1 ^ nx;
              x1 = nx;
            }
            if (root) {
              ncyc++;
              if (!(orr & 1u)) {
                nun++;
                const unsigned kk = atomicAdd(&UL[0], 1u);
                if (kk < (unsigned)kWalkMaxUnoriented) UL[1 + kk] = u0;
              }
            }
          }
        }
      }
    }
  }
  nadj = __reduce_add_sync(kFull, nadj) >> 15;        // counted in units of kVis
  ncyc = __reduce_add_sync(kFull, ncyc);
  nun = __reduce_add_sync(kFull, nun);
  defer = __any_sync(kFull, defer);
  __syncwarp();
  const int b = (n + 1) - (int)nadj, c = (int)ncyc;

  // ---- unoriented cycles: the "covered by an oriented cycle" test of dist_core -------------------
  if (!defer && nun != 0u) {
    if (nun > (unsigned)kWalkMaxUnoriented) defer = 1;
    for (unsigned k = 0; k < nun && !defer; k++) {
      const unsigned qq = UL[1 + k];
      bool hit = false;
      // the first vertices right of q nearly always decide (q sits inside a long oriented cycle):
      // 4 of them first, then 32 at a time, 260 in all
      unsigned off = 0;
      for (int it = 0; it < 9; it++) {
        const unsigned width = (it == 0) ? 4u : 32u;
        const unsigned v = qq + 2u + off + lane;
        off += width;
        int st = 2;                                   // past the last vertex: stop
        if ((unsigned)lane >= width) st = 0;          // not probing in this round
        else if (v < (unsigned)S) {
          st = 0;
          const unsigned ent = H[v + 1];
          if ((ent & 0x7FFFu) != v) {                 // not an adjacency end
            unsigned ori;
            const unsigned cy = walk_cycle_of(H, NW, MAP, v, grid, ori);
            if (cy == qq) st = 2;                     // Q's own next vertex: the probe ends here
            else if (cy < qq && ori) st = 1;          // an oriented cycle that starts left of Q
          }
        }
        const unsigned mh = __ballot_sync(kFull, st == 1), ms = __ballot_sync(kFull, st == 2);
        if (mh | ms) { hit = mh && (!ms || __ffs((int)mh) < __ffs((int)ms)); break; }
      }
      if (!hit) defer = 1;
    }
  }

  if (lane == 0) {
    if (defer) defer_list[atomicAdd(defer_count, 1u)] = (int32_t)p;
    else put_result(out, stride, p, gi, pj, b - c, b, c, 0, 0);
  }
  __syncwarp();
  return defer != 0;
}

// A warp that had to defer every one of its first kWalkGiveUp pairs stops trying: on a batch of
// hurdle-rich pairs the first tier then costs a few per cent instead of its full time on top of the
// second tier's.  The rest of its pairs go to the deferred list unseen.
static const int kWalkGiveUp = 8;
struct WalkQuota {
  int done, deferred;
  __device__ inline bool gave_up() const { return done >= kWalkGiveUp && deferred == done; }
  __device__ inline void count(bool was_deferred) { done++; deferred += was_deferred ? 1 : 0; }
};
__device__ inline void walk_defer_unseen(long long p, int32_t *defer_list, unsigned *defer_count) {
  if ((threadIdx.x & 31) == 0) defer_list[atomicAdd(defer_count, 1u)] = (int32_t)p;
}

// Verdict of walk_order_kernel: runs of equal pi rows are long (at most one change per 64 pairs).
__device__ inline bool walk_list_sorted(const unsigned *changes, long long npairs) {
  return (unsigned long long)*changes * 64ull <= (unsigned long long)npairs;
}

// Work distribution of both kernels: guided self-scheduling over one zero-initialised global counter of
// pairs.  A taker (a warp, or a block of the shared-row kernel) claims the next `chunk` pairs, where
// chunk = what is left / (2 x takers), at least `lo`: large chunks while there is plenty of work (few
// claims, few block barriers), single pairs per warp at the end (nobody waits long for the slowest).
// guide4 = 4 x the divisor per taker (6: claim two thirds of an even share of what is left; measured best of 4..32)
__device__ inline long long walk_claim(unsigned *next, long long npairs, long long takers, int lo, int guide4,
                                       long long *count) {
  const long long seen = (long long)*(volatile unsigned *)next;
  long long c = 4 * (npairs - seen) / (guide4 * takers);
  if (c < lo) c = lo;
  const long long start = (long long)atomicAdd(next, (unsigned)c);
  *count = c;
  return start;
}

// One warp, one pair at a time, every warp with its own staged pi row: a warp draws chunks of
// consecutive pairs from the ticket; tail[] stays staged while consecutive pairs share their pi row.
// Serves pair lists in any order.
__global__ void __launch_bounds__(256)
dist_walk_kernel(const int32_t *__restrict__ genomes, const int32_t *__restrict__ sinv, int n,
                 const int32_t *__restrict__ pairs, long long npairs, int32_t *__restrict__ out,
                 int stride, int shape, int32_t *defer_list, unsigned *defer_count,
                 const unsigned *__restrict__ order_flag, unsigned *ticket) {
  if (order_flag && walk_list_sorted(order_flag, npairs)) return;   // pi-sorted list: the shared-row kernel runs
  GRSR_DYN_SMEM(smem_raw);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  unsigned char *base = smem_raw + (size_t)wid * walk_warp_bytes(n);
  uint16_t *tail = (uint16_t *)base;
  WalkMem wm;
  wm.tail = tail;
  walk_carve_private(wm, base + walk_tail_bytes(n), n);
  walk_init_private(wm, n, lane);
  if (lane == 0) tail[n] = (uint16_t)(2 * n);         // the sentinel gene's entry (vertex 2n+1, minus one)
  __syncwarp();

  int staged = -1;
  WalkQuota quota = {0, 0};
  for (;;) {
    long long p0 = 0, chunk = 0;
    if (lane == 0) p0 = walk_claim(ticket, npairs, (long long)gridDim.x * wpb, 1, kWalkGuide4, &chunk);
    p0 = __shfl_sync(kFull, p0, 0);
    chunk = __shfl_sync(kFull, chunk, 0);
    if (p0 >= npairs) break;
    const long long p1 = (p0 + chunk < npairs) ? p0 + chunk : npairs;
    for (long long p = p0; p < p1; p++) {
      if (quota.gave_up()) { walk_defer_unseen(p, defer_list, defer_count); continue; }
      const int gi = pairs[2 * p], pj = pairs[2 * p + 1];
      if (pj != staged) {
        const int32_t *sp = sinv + (long long)pj * n;
        for (int i = lane; i < n; i += 32) {
          const int s = sp[i];
          tail[i] = walk_tail_entry(s);                  // vertex where gene +(i+1) begins in pi, minus one
        }
        staged = pj;
        __syncwarp();
      }
      quota.count(walk_pair(wm, genomes, n, gi, pj, p, out, stride, shape, defer_list, defer_count));
    }
  }
}

// The same for pair lists sorted by pi row (the column-by-column triangle of a distance matrix): the
// warps of a block share ONE staged pi row, which leaves room for more resident warps per SM.  A block
// draws chunks of consecutive pairs from the ticket and goes through a chunk one run of equal pi rows
// at a time; inside a run its warps take pairs from a shared-memory counter, one at a time.
__global__ void __launch_bounds__(1024)
dist_walk_shared_kernel(const int32_t *__restrict__ genomes, const int32_t *__restrict__ sinv, int n,
                        const int32_t *__restrict__ pairs, long long npairs, int32_t *__restrict__ out,
                        int stride, int shape, int32_t *defer_list, unsigned *defer_count,
                        const unsigned *__restrict__ order_flag, unsigned *ticket) {
  if (order_flag && !walk_list_sorted(order_flag, npairs)) return;  // not sorted: the private-row kernel runs
  GRSR_DYN_SMEM(smem_raw);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  unsigned *s_ctl = (unsigned *)smem_raw;             // first 16 bytes: [0] run length, [1] next pair of the run, [2] chunk start, [3] chunk length
  uint16_t *tail = (uint16_t *)(smem_raw + 16);
  WalkMem wm;
  wm.tail = tail;
  walk_carve_private(wm, smem_raw + 16 + walk_tail_bytes(n) + (size_t)wid * walk_private_bytes(n), n);
  walk_init_private(wm, n, lane);
  if (threadIdx.x == 0) tail[n] = (uint16_t)(2 * n);

  WalkQuota quota = {0, 0};
  int staged = -1;
  for (;;) {
    __syncthreads();                                   // everybody is done with the previous chunk
    if (threadIdx.x == 0) {
      long long c;
      const int g4 = (shape >> 16) & 0xFF;
      const long long st = walk_claim(ticket, npairs, gridDim.x, wpb, g4 ? g4 : kWalkGuide4, &c);
      s_ctl[2] = (unsigned)(st < npairs ? st : npairs);
      s_ctl[3] = (unsigned)c;
    }
    __syncthreads();
    const long long b0 = (long long)s_ctl[2];
    if (b0 >= npairs) break;
    const long long b1 = (b0 + (long long)s_ctl[3] < npairs) ? b0 + (long long)s_ctl[3] : npairs;
    for (long long p = b0; p < b1;) {
      const int pj = pairs[2 * p + 1];
      const unsigned left = (unsigned)(b1 - p);
      __syncthreads();                                 // everybody is done with the previous run
      if (threadIdx.x == 0) { s_ctl[0] = left; s_ctl[1] = 0; }
      __syncthreads();
      for (unsigned qb = 0; qb < left; qb += blockDim.x) {
        const unsigned q = qb + threadIdx.x;
        const bool diff = q < left && pairs[2 * (p + q) + 1] != pj;
        if (diff) atomicMin(&s_ctl[0], q);
        if (__syncthreads_or(diff)) break;
      }
      if (pj != staged) {
        const int32_t *sp = sinv + (long long)pj * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const int s = sp[i];
          tail[i] = walk_tail_entry(s);
        }
        staged = pj;
      }
      __syncthreads();
      const unsigned run = s_ctl[0];
      for (;;) {
        unsigned q = 0;
        if (lane == 0) q = atomicAdd(&s_ctl[1], 1u);
        q = __shfl_sync(kFull, q, 0);
        if (q >= run) break;
        if (quota.gave_up()) { walk_defer_unseen(p + q, defer_list, defer_count); continue; }
        quota.count(walk_pair(wm, genomes, n, pairs[2 * (p + q)], pj, p + q, out, stride, shape,
                              defer_list, defer_count));
      }
      p += run;
    }
  }
}

// Decides which of the two kernels serves a list: counts the positions where the pi row changes
// (*changes must be zero at launch).  Runs of equal pi rows are "long" when there is at most one
// change per 64 pairs on average (walk_list_sorted).
__global__ void walk_order_kernel(const int32_t *__restrict__ pairs, long long npairs, unsigned *changes) {
  unsigned c = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q + 1 < npairs; q += stride)
    c += (pairs[2 * q + 1] != pairs[2 * q + 3]) ? 1u : 0u;
  c = __reduce_add_sync(kFull, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(changes, c);
}

// Second tier: the pairs the walk kernel deferred, through dist_core (block per pair).
template <int NT, int EPT>
__global__ void __launch_bounds__(NT, GRSR_MIN_BLOCKS(NT, EPT))
dist_deferred_kernel(const int32_t *__restrict__ genomes, const int32_t *__restrict__ sinv, int n,
                     const int32_t *__restrict__ pairs, const int32_t *__restrict__ defer_list,
                     const unsigned *__restrict__ defer_count, int32_t *__restrict__ out, int stride) {
  GRSR_DYN_SMEM(smem_raw);
  Work wk;
  work_carve(wk, smem_raw, n, NT * EPT, false);
  GRSR_PHASE_INIT(); GRSR_PHASE_BASE(8);
  uint16_t *stage = (uint16_t *)(smem_raw + work_bytes(n, NT * EPT, false));
  const long long total = (long long)*defer_count;
  int staged = -1;
  for (long long q = blockIdx.x; q < total; q += gridDim.x) {
    const long long p = defer_list[q];
    const int gi = pairs[2 * p], pj = pairs[2 * p + 1];
    if (pj != staged) {
      __syncthreads();
      const int32_t *sp = sinv + (long long)pj * n;
      for (int i = threadIdx.x; i < n; i += NT) {
        const int s = sp[i];
        stage[i] = (uint16_t)(s > 0 ? 2 * s - 1 : -2 * s);
      }
      staged = pj;
      __syncthreads();
    }
    EndsFromRows ends; ends.gamma = genomes + (long long)gi * n; ends.tail = stage;
    Stats st = dist_core<NT, EPT, false>(wk, n, ends);
    if (threadIdx.x == 0) put_result(out, stride, p, gi, pj, st.d, st.b, st.c, st.h, st.f);
  }
}

} // namespace grsr
