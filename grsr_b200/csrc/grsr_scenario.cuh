// grsr_scenario.cuh -- device code of the sorting-by-reversals scenario search (sm_100a).
//
// One thread block owns one (source, destination) pair for the whole scenario and keeps the
// current permutation, the destination's signed inverse and the breakpoint graph on chip (shared
// memory) -- or, for genomes too long for that, in an L2-resident global workspace; blocks take
// pairs from a global ticket counter, so a launch sorts many pairs concurrently.
// It reproduces, step for step, the reference's default -s engine:
//
//   print_scenario_2   G/opt_scenario.c:227-290   outer loop: rebuild G0 = graph(gamma = dest,
//                                                 pi = current), stop at d == 0
//   try_optrev         G/opt_scenario.c:527-576   candidates (i, j) in (j - i, i) order whose two
//                                                 cut points are breakpoints
//   optrev_try_op      G/opt_scenario.c:951-986   if G0 has no hurdle: keep (i, j) only when
//                                                 graphdist(2i+1, 2j+2) is even; then accept the
//                                                 first whose recomputed distance is d - 1
//
// "graphdist even" is evaluated in O(1): vertices x and y are an even number of edges apart
// (walking black edge first) exactly when they lie on the same orbit of sigma(v) = grey[v ^ 1],
// and dist_core<keep orbits> leaves every vertex's orbit label in ORB.  The candidate scan is a
// parallel arg-min over the breakpoint list instead of the reference's O(n^2) double loop; rejected
// candidates are resumed from in key order, so the accepted reversal is the reference's.
#pragma once
#include "grsr_core.cuh"

namespace grsr {

__device__ inline unsigned long long blk_min_u64(unsigned long long v, unsigned long long *red) {
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long x = __shfl_down_sync(kFull, v, o);
    v = (x < v) ? x : v;
  }
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = (lane < nw) ? red[lane] : ~0ull;
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long x = __shfl_down_sync(kFull, v, o);
      v = (x < v) ? x : v;
    }
    if (lane == 0) red[32] = v;
  }
  __syncthreads();
  v = red[32];
  __syncthreads();
  return v;
}

template <class WK> __device__ inline unsigned long long team_min64(TeamBlock, unsigned long long v, WK &wk) {
  return blk_min_u64(v, wk.red);
}
template <class WK> __device__ inline unsigned long long team_min64(TeamCluster, unsigned long long v, WK &wk) {
  volatile unsigned long long *x = TeamCluster::slot(wk);
  v = blk_min_u64(v, wk.red);
  if (threadIdx.x == 0) x[cluster_rank()] = v;
  TeamCluster::sync();
  unsigned long long r = TeamCluster::part(x, ~0ull);
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long y = __shfl_xor_sync(kFull, r, o);
    r = (y < r) ? y : r;
  }
  return r;
}

// ends() of G0 = graph(gamma = destination, pi = current) (G/opt_scenario.c:229): the signed
// position in the current genome of the destination's t-th gene is the inverse of rho0.
template <class SV>
struct EndsG0 {
  const SV *inv0;
  __device__ inline void operator()(int t, unsigned &L, unsigned &R) const {
    ends_from_signed_pos((int)inv0[t], L, R);
  }
};

// ends() of the trial evaluation invdist_noncircular(trial, dest) (G/opt_scenario.c:980), gamma =
// trial = current with [i..j] reversed and negated (G/scenario.c:475-486), pi = destination:
// signed position in the destination of trial's t-th gene.
template <class SV>
struct EndsTrial {
  const SV *rho0; int i, j;
  __device__ inline void operator()(int t, unsigned &L, unsigned &R) const {
    const int q = (t >= i && t <= j) ? -(int)rho0[i + j - t] : (int)rho0[t];
    ends_from_signed_pos(q, L, R);
  }
};

// The pair ticket may live in ANOTHER device's memory (several GPUs drawing pairs from one queue over
// NVLink peer mapping): system-scope atomic.
#ifdef GRSR_EMU
__device__ inline unsigned ticket_next(unsigned int *ticket) { return atomicAdd(ticket, 1u); }
#else
__device__ inline unsigned ticket_next(unsigned int *ticket) { return atomicAdd_system(ticket, 1u); }
#endif

// How a launch maps tickets to pairs: pair = first + ticket * stride.  One queue shared by all
// devices: (0, 1); device k of g without peer access: (k, g) with a private ticket.
struct PairQueue { unsigned int *ticket; int first, stride; };

#define GRSR_SCEN_OK        0
#define GRSR_SCEN_OVERFLOW  1   /* more steps than max_steps */
#define GRSR_SCEN_STUCK     2   /* no candidate reduces the distance (reference: S_FAIL) */

// Policy of the shared-memory build: 16-bit ids and values, register-resident jumping.
template <int NT, int EPT>
struct ScenSmall {
  typedef uint16_t vid;
  typedef int16_t sval;
  typedef WorkT<uint16_t> work;
  typedef TeamBlock team;
  template <bool kKeep, class E>
  static __device__ inline Stats dist(work &wk, int n, E e) { return dist_core<NT, EPT, kKeep>(wk, n, e); }
};

// Policy of the global-memory build: 32-bit ids and values; TM = one block or a whole cluster per pair.
template <class TM>
struct ScenLarge {
  typedef uint32_t vid;
  typedef int32_t sval;
  typedef WorkT<uint32_t> work;
  typedef TM team;
  template <bool kKeep, class E>
  static __device__ inline Stats dist(work &wk, int n, E e) { return dist_core_large<TM, kKeep>(wk, n, e); }
};

// The scenario of every pair this block draws from the ticket.  CUR, SDST, INV0, RHO0: n values
// each; BPL: n+1 ids; they and wk live wherever the policy's kernel put them.
// steps: npairs x max_steps (start, end) pairs; nsteps/status: per pair; evals: per pair number of
// full trial evaluations (diagnostic, may be null); ticket: zero-initialised work counter.
template <class Pol>
__device__ inline void scenario_body(typename Pol::work &wk, typename Pol::sval *CUR,
                                     typename Pol::sval *SDST, typename Pol::sval *INV0,
                                     typename Pol::sval *RHO0, typename Pol::vid *BPL,
                                     const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                                     int npairs, int n, int32_t *__restrict__ steps,
                                     int32_t *__restrict__ nsteps, int max_steps,
                                     int32_t *__restrict__ status, unsigned long long *__restrict__ evals,
                                     PairQueue queue, volatile int *s_pair, volatile unsigned *bound,
                                     int resume = 0) {
  typedef typename Pol::sval SV;
  typedef typename Pol::vid VT;
  typedef typename Pol::team TM;
  constexpr unsigned kNone = none_of<VT>();
  const unsigned long long kNoKey = ~0ull;
  const VT *ORB = wk.ORB;
  const int tid = TM::tid(), nt = TM::size();
  GRSR_PHASE_INIT();

  for (;;) {
    if (tid == 0) {
      const long long q = (long long)queue.first + (long long)ticket_next(queue.ticket) * queue.stride;
      *s_pair = (q < npairs) ? (int)q : npairs;
    }
    TM::sync();
    const int p = *s_pair;
    TM::sync();
    if (p >= npairs) break;

    const int32_t *ps = src + (long long)p * n, *pd = dst + (long long)p * n;
    for (int i = tid; i < n; i += nt) {
      CUR[i] = (SV)ps[i];
      int g = pd[i];
      SDST[(g > 0 ? g : -g) - 1] = (SV)(g > 0 ? (i + 1) : -(i + 1));
    }
    TM::sync();

    // resume: steps [0, nsteps[p]) were written (and evals[p] evaluations made) by the stepwise path;
    // a pair it ended (step limit reached, search stuck) is left as it is
    if (resume && status[p] != GRSR_SCEN_OK) continue;
    int ns = resume ? nsteps[p] : 0, code = GRSR_SCEN_OK;
    unsigned long long nevals = (resume && evals) ? evals[p] : 0ull;
    for (;;) {
      // rho0 (signed position in the destination of the gene at position i of the current genome)
      // and its inverse: G0 and every trial of this step are written from these two arrays
      GRSR_PHASE(1);
      for (int i = tid; i < n; i += nt) {
        int g = CUR[i];
        int s = SDST[(g > 0 ? g : -g) - 1];
        int r = g > 0 ? s : -s;
        RHO0[i] = (SV)r;
        INV0[(r > 0 ? r : -r) - 1] = (SV)(r > 0 ? (i + 1) : -(i + 1));
      }
      TM::sync();
      EndsG0<SV> eg; eg.inv0 = INV0;
      GRSR_PHASE_BASE(2);
      const Stats g0 = Pol::template dist<true>(wk, n, eg);   // G/opt_scenario.c:229
      GRSR_PHASE(6);
      if (g0.d == 0) break;
      if (ns >= max_steps) { code = GRSR_SCEN_OVERFLOW; break; }

      // black edges that are breakpoints, ascending: edge k = (2k, 2k+1) sits left of position k.
      // A reversal (i, j) cuts edges i and j+1 (the tests of G/opt_scenario.c:538-561).
      const int nbp = TM::compact(n + 1, [&](int k, unsigned &val) -> bool {
        if (ORB[2 * k] == kNone) return false;
        val = (unsigned)k; return true;
      }, BPL, wk);
      const bool filter = (g0.h == 0);                        // G/opt_scenario.c:951

      int fi = -1, fj = -1;
      if (!filter) {
        // Hurdles present: no filter, every pair of breakpoint edges (e1 < e2) is a candidate
        // (i = e1, j = e2 - 1) and gets a full evaluation, in (j - i, i) order, until one yields
        // d - 1 (G/opt_scenario.c:951, 527-576).  One diagonal j - i = g at a time: its candidates are
        // listed in order (INV0 is free once G0 is built) and evaluated one after another; the next
        // diagonal is the smallest gap > g that occurs between two breakpoint edges.
        SV *CL = INV0;
        for (int g = 0; fi < 0;) {
          GRSR_PHASE(7);
          unsigned long long mg = kNoKey;
          for (int t = tid; t < nbp; t += nt) {
            const int need = (int)BPL[t] + 1 + g;
            int lo = t + 1, hi = nbp;
            while (lo < hi) { int mid = (lo + hi) >> 1; if ((int)BPL[mid] < need) lo = mid + 1; else hi = mid; }
            if (lo < nbp) { const unsigned long long gg = (unsigned long long)((int)BPL[lo] - (int)BPL[t] - 1); mg = gg < mg ? gg : mg; }
          }
          mg = team_min64(TM(), mg, wk);
          if (mg == kNoKey) break;                            // no candidate left at all
          g = (int)mg;
          const int nc = TM::compact(nbp, [&](int t, unsigned &val) -> bool {
            const int want = (int)BPL[t] + 1 + g;
            int lo = t + 1, hi = nbp;
            while (lo < hi) { int mid = (lo + hi) >> 1; if ((int)BPL[mid] < want) lo = mid + 1; else hi = mid; }
            if (lo >= nbp || (int)BPL[lo] != want) return false;
            val = (unsigned)BPL[t]; return true;
          }, CL, wk);
          for (int q = 0; q < nc; q++) {
            const int i = (int)CL[q], j = i + g;
            EndsTrial<SV> et; et.rho0 = RHO0; et.i = i; et.j = j;
            GRSR_PHASE_BASE(8);
            const Stats t1 = Pol::template dist<false>(wk, n, et);  // G/opt_scenario.c:974-986
            nevals++;
            if (t1.d == g0.d - 1) { fi = i; fj = j; break; }
          }
          g++;
        }
      }
      unsigned long long lastkey = kNoKey;
      // hurdle-free graph: the smallest reversal length found so far in the current search is shared
      // by the block: a lane whose orbit partner is far away stops scanning as soon as somebody has a
      // shorter candidate
      while (filter) {
        // smallest key (j - i, i) greater than lastkey among admissible candidates
        GRSR_PHASE(7);
        if (tid == 0) *bound = 0xFFFFFFFFu;
        TM::sync();
        unsigned long long best = kNoKey;
        for (int t = tid; t < nbp; t += nt) {
          int e1 = (int)BPL[t];
          int s = t + 1;
          if (lastkey != kNoKey) {
            int g = (int)(lastkey >> 32), i0 = (int)(lastkey & 0xFFFFFFFFull);
            int lb = e1 + 1 + g + (e1 > i0 ? 0 : 1);          // smallest admissible right edge
            int lo = t + 1, hi = nbp;
            while (lo < hi) { int mid = (lo + hi) >> 1; if ((int)BPL[mid] < lb) lo = mid + 1; else hi = mid; }
            s = lo;
          }
          int e2 = -1;
          const unsigned want = ORB[2 * e1 + 1];
          for (; s < nbp; s++) {
            const int c = (int)BPL[s];
            if ((unsigned)(c - e1 - 1) > *bound) break;       // cannot beat a candidate already found
            if (ORB[2 * c] == want) { e2 = c; break; }
          }
          if (e2 >= 0) {
            const unsigned len0 = (unsigned)(e2 - e1 - 1);
            unsigned long long key = ((unsigned long long)len0 << 32) | (unsigned)e1;
            best = (key < best) ? key : best;
            atomicMin((unsigned *)bound, len0);
          }
        }
        best = team_min64(TM(), best, wk);
        if (best == kNoKey) break;
        const int i = (int)(best & 0xFFFFFFFFull);
        const int j = i + (int)(best >> 32);
        EndsTrial<SV> et; et.rho0 = RHO0; et.i = i; et.j = j;
        GRSR_PHASE_BASE(8);
        const Stats t1 = Pol::template dist<false>(wk, n, et);  // G/opt_scenario.c:974-986
        nevals++;
        if (t1.d == g0.d - 1) { fi = i; fj = j; break; }
        lastkey = best;
      }
      if (fi < 0) { code = GRSR_SCEN_STUCK; break; }

      // current <- trial (G/opt_scenario.c:283-285)
      GRSR_PHASE(13);
      for (int k = tid; 2 * k <= fj - fi; k += nt) {
        int a = fi + k, b = fj - k;
        SV x = CUR[a], y = CUR[b];
        CUR[a] = (SV)(-y); CUR[b] = (SV)(-x);
      }
      if (tid == 0) {
        int32_t *o = steps + 2 * ((long long)p * max_steps + ns);
        o[0] = fi; o[1] = fj;
      }
      ns++;
      TM::sync();
    }
    if (tid == 0) {
      nsteps[p] = ns; status[p] = code;
      if (evals) evals[p] = nevals;
    }
    TM::sync();
  }
}

__host__ __device__ inline size_t scen_arr_bytes(int n) { return (((size_t)n + 2) * 2 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t scen_extra_bytes(int n) { return 5 * scen_arr_bytes(n); }

template <int NT, int EPT>
__global__ void __launch_bounds__(NT, GRSR_MIN_BLOCKS(NT, EPT))
scenario_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int npairs, int n,
                int32_t *__restrict__ steps, int32_t *__restrict__ nsteps, int max_steps,
                int32_t *__restrict__ status, unsigned long long *__restrict__ evals,
                PairQueue queue, int resume) {
  GRSR_DYN_SMEM(smem_raw);
  Work wk;
  work_carve(wk, smem_raw, n, NT * EPT, true);
  GRSR_PHASE_INIT(); GRSR_PHASE_BASE(8);
  unsigned char *extra = smem_raw + work_bytes(n, NT * EPT, true);
  int16_t *CUR = (int16_t *)extra;
  int16_t *SDST = (int16_t *)(extra + scen_arr_bytes(n));
  int16_t *INV0 = (int16_t *)(extra + 2 * scen_arr_bytes(n));
  int16_t *RHO0 = (int16_t *)(extra + 3 * scen_arr_bytes(n));
  uint16_t *BPL = (uint16_t *)(extra + 4 * scen_arr_bytes(n));
  __shared__ int s_pair;
  scenario_body<ScenSmall<NT, EPT> >(wk, CUR, SDST, INV0, RHO0, BPL, src, dst, npairs, n, steps, nsteps,
                                     max_steps, status, evals, queue, &s_pair,
                                     (volatile unsigned *)&wk.red[33], resume);
}

__host__ __device__ inline size_t scen_large_arr_bytes(int n) { return (((size_t)n + 2) * 4 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t scen_large_bytes(int n) {
  return ((large_work_bytes(n, true) + 5 * scen_large_arr_bytes(n)) + 255) & ~(size_t)255;
}

// global-memory build: scratch + blockIdx.x * scen_large_bytes(n) is this block's workspace
__global__ void __launch_bounds__(1024)
scenario_large_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int npairs, int n,
                      int32_t *__restrict__ steps, int32_t *__restrict__ nsteps, int max_steps,
                      int32_t *__restrict__ status, unsigned long long *__restrict__ evals,
                      PairQueue queue, unsigned char *scratch) {
  __shared__ unsigned long long red[34];
  __shared__ int s_pair;
  unsigned char *base = scratch + (size_t)blockIdx.x * scen_large_bytes(n);
  WorkT<uint32_t> wk;
  large_work_carve(wk, base, n, true, red);
  unsigned char *extra = base + large_work_bytes(n, true);
  int32_t *CUR = (int32_t *)extra;
  int32_t *SDST = (int32_t *)(extra + scen_large_arr_bytes(n));
  int32_t *INV0 = (int32_t *)(extra + 2 * scen_large_arr_bytes(n));
  int32_t *RHO0 = (int32_t *)(extra + 3 * scen_large_arr_bytes(n));
  uint32_t *BPL = (uint32_t *)(extra + 4 * scen_large_arr_bytes(n));
  scenario_body<ScenLarge<TeamBlock> >(wk, CUR, SDST, INV0, RHO0, BPL, src, dst, npairs, n, steps, nsteps,
                                       max_steps, status, evals, queue, &s_pair,
                                       (volatile unsigned *)&red[33]);
}

// All candidate reversals of ONE (source, destination) pair scored in parallel -- the inner loop
// of print_rev_stats_t1 (reference G/testrev.c:156-215) and the evaluation step of optrev_try_op
// (G/opt_scenario.c:974-986): out[c] = invdist_noncircular(trial_c, dest) where trial_c is the
// source with positions cands[2c]..cands[2c+1] reversed and negated.  One block per candidate
// (grid-stride); rho0 of the pair is rebuilt once per block in shared memory.
template <int NT, int EPT>
__global__ void __launch_bounds__(NT, GRSR_MIN_BLOCKS(NT, EPT))
trial_batch_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int n,
                   const int32_t *__restrict__ cands, long long ncand, int32_t *__restrict__ out,
                   const int32_t *__restrict__ ncand_dev) {
  GRSR_DYN_SMEM(smem_raw);
  Work wk;
  work_carve(wk, smem_raw, n, NT * EPT, false);
  GRSR_PHASE_INIT(); GRSR_PHASE_BASE(8);
  unsigned char *extra = smem_raw + work_bytes(n, NT * EPT, false);
  int16_t *SDST = (int16_t *)extra;
  int16_t *RHO0 = (int16_t *)(extra + scen_arr_bytes(n));
  const int tid = threadIdx.x;
  if (ncand_dev) ncand = *ncand_dev;            // candidate count produced on the device
  if ((long long)blockIdx.x >= ncand) return;
  for (int i = tid; i < n; i += NT) {
    int g = dst[i];
    SDST[(g > 0 ? g : -g) - 1] = (int16_t)(g > 0 ? (i + 1) : -(i + 1));
  }
  __syncthreads();
  for (int i = tid; i < n; i += NT) {
    int g = src[i];
    int s = SDST[(g > 0 ? g : -g) - 1];
    RHO0[i] = (int16_t)(g > 0 ? s : -s);
  }
  __syncthreads();
  for (long long c = blockIdx.x; c < ncand; c += gridDim.x) {
    EndsTrial<int16_t> et; et.rho0 = RHO0; et.i = cands[2 * c]; et.j = cands[2 * c + 1];
    const Stats st = dist_core<NT, EPT, false>(wk, n, et);
    if (tid == 0) out[c] = st.d;
  }
}

// cluster build: one thread-block cluster per pair, workspace scratch + cluster * scen_large_bytes(n),
// exchange area xch_all + cluster * kXchWords (its last words hold the pair ticket and the search
// bound).  Dynamic shared memory: 34 u64 of per-block reduction scratch.
__global__ void __launch_bounds__(1024)
scenario_cluster_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int npairs, int n,
                        int32_t *__restrict__ steps, int32_t *__restrict__ nsteps, int max_steps,
                        int32_t *__restrict__ status, unsigned long long *__restrict__ evals,
                        PairQueue queue, unsigned char *scratch, unsigned long long *xch_all) {
  GRSR_DYN_SMEM(smem_raw);
  unsigned long long *red = (unsigned long long *)smem_raw;
  const unsigned cid = blockIdx.x / cluster_size();
  unsigned char *base = scratch + (size_t)cid * scen_large_bytes(n);
  unsigned long long *xch = xch_all + (size_t)cid * kXchWords;
  WorkT<uint32_t> wk;
  large_work_carve(wk, base, n, true, red);
  wk.xch = xch;
  unsigned char *extra = base + large_work_bytes(n, true);
  int32_t *CUR = (int32_t *)extra;
  int32_t *SDST = (int32_t *)(extra + scen_large_arr_bytes(n));
  int32_t *INV0 = (int32_t *)(extra + 2 * scen_large_arr_bytes(n));
  int32_t *RHO0 = (int32_t *)(extra + 3 * scen_large_arr_bytes(n));
  uint32_t *BPL = (uint32_t *)(extra + 4 * scen_large_arr_bytes(n));
  scenario_body<ScenLarge<TeamCluster> >(wk, CUR, SDST, INV0, RHO0, BPL, src, dst, npairs, n, steps, nsteps,
                                         max_steps, status, evals, queue,
                                         (volatile int *)(xch + kXchWords - 4),
                                         (volatile unsigned *)(xch + kXchWords - 2));
}

// cluster build of the batched distance (few, very long pairs)
__global__ void __launch_bounds__(1024)
dist_pairs_cluster_kernel(const int32_t *__restrict__ genomes, const int32_t *__restrict__ sinv, int n,
                          const int32_t *__restrict__ pairs, long long npairs, int32_t *__restrict__ out,
                          int stride, unsigned char *scratch, unsigned long long *xch_all) {
  GRSR_DYN_SMEM(smem_raw);
  unsigned long long *red = (unsigned long long *)smem_raw;
  const unsigned cs = cluster_size();
  const unsigned cid = blockIdx.x / cs, nclusters = gridDim.x / cs;
  WorkT<uint32_t> wk;
  large_work_carve(wk, scratch + (size_t)cid * large_work_bytes(n, false), n, false, red);
  wk.xch = xch_all + (size_t)cid * kXchWords;
  for (long long p = cid; p < npairs; p += nclusters) {
    EndsFromRowsGlobal ends;
    ends.gamma = genomes + (long long)pairs[2 * p] * n;
    ends.sinv_pi = sinv + (long long)pairs[2 * p + 1] * n;
    Stats st = dist_core_large<TeamCluster, false>(wk, n, ends);
    if (TeamCluster::tid() == 0)
      put_result(out, stride, p, pairs[2 * p], pairs[2 * p + 1], st.d, st.b, st.c, st.h, st.f);
  }
}

} // namespace grsr
