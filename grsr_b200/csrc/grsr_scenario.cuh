// grsr_scenario.cuh -- device code of the sorting-by-reversals scenario search (sm_100a).
//
// One thread block owns one (source, destination) pair for the whole scenario and keeps the
// current permutation, the destination's signed inverse and the breakpoint graph in shared memory;
// blocks take pairs from a global ticket counter, so a launch sorts many pairs concurrently.
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
// and dist_core<true> leaves every vertex's orbit label in ORB.  The candidate scan is a parallel
// arg-min over the breakpoint list instead of the reference's O(n^2) double loop; rejected
// candidates are resumed from in key order, so the accepted reversal is the reference's.
#pragma once
#include "grsr_core.cuh"

namespace grsr {

static const unsigned kNoKey = 0xFFFFFFFFu;
static const int kKeyShift = 14;             // key = (j - i) << 14 | i ; n <= kMaxSmemGenes < 2^14

__host__ __device__ inline size_t scen_arr_bytes(int n) { return (((size_t)n + 2) * 2 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t scen_extra_bytes(int n) { return 5 * scen_arr_bytes(n); }

__device__ inline unsigned blk_min_u32(unsigned v, unsigned long long *red) {
  v = __reduce_min_sync(kFull, v);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = (lane < nw) ? (unsigned)red[lane] : kNoKey;
    v = __reduce_min_sync(kFull, v);
    if (lane == 0) red[32] = v;
  }
  __syncthreads();
  v = (unsigned)red[32];
  __syncthreads();
  return v;
}

// rho0(i): signed position in the destination of the gene at position i of the current genome
struct RhoCurrent {
  const int16_t *cur; const int16_t *sdst;
  __device__ inline int operator()(int i) const {
    int g = cur[i];
    int s = sdst[(g > 0 ? g : -g) - 1];
    return g > 0 ? s : -s;
  }
};

// ends() of G0 = graph(gamma = destination, pi = current) (G/opt_scenario.c:229): the signed
// position in the current genome of the destination's t-th gene is the inverse of rho0.
struct EndsG0 {
  const int16_t *inv0;
  __device__ inline void operator()(int t, unsigned &L, unsigned &R) const {
    ends_from_signed_pos(inv0[t], L, R);
  }
};

// ends() of the trial evaluation invdist_noncircular(trial, dest) (G/opt_scenario.c:980), gamma =
// trial = current with [i..j] reversed and negated (G/scenario.c:475-486), pi = destination:
// signed position in the destination of trial's t-th gene.
struct EndsTrial {
  const int16_t *rho0; int i, j;
  __device__ inline void operator()(int t, unsigned &L, unsigned &R) const {
    const int q = (t >= i && t <= j) ? -(int)rho0[i + j - t] : (int)rho0[t];
    ends_from_signed_pos(q, L, R);
  }
};

#define GRSR_SCEN_OK        0
#define GRSR_SCEN_OVERFLOW  1   /* more steps than max_steps */
#define GRSR_SCEN_STUCK     2   /* no candidate reduces the distance (reference: S_FAIL) */

// steps: npairs x max_steps (start, end) pairs; nsteps/status: per pair; evals: per pair number of
// full trial evaluations (diagnostic, may be null); ticket: zero-initialised work counter.
template <int NT, int EPT>
__global__ void __launch_bounds__(NT, GRSR_MIN_BLOCKS(NT, EPT))
scenario_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int npairs, int n,
                int32_t *__restrict__ steps, int32_t *__restrict__ nsteps, int max_steps,
                int32_t *__restrict__ status, unsigned long long *__restrict__ evals,
                unsigned int *ticket) {
  GRSR_DYN_SMEM(smem_raw);
  Work wk;
  work_carve(wk, smem_raw, n, NT * EPT, true);
  unsigned char *extra = smem_raw + work_bytes(n, NT * EPT, true);
  int16_t *CUR = (int16_t *)extra;
  int16_t *SDST = (int16_t *)(extra + scen_arr_bytes(n));
  int16_t *INV0 = (int16_t *)(extra + 2 * scen_arr_bytes(n));
  int16_t *RHO0 = (int16_t *)(extra + 3 * scen_arr_bytes(n));
  vid_t *BPL = (vid_t *)(extra + 4 * scen_arr_bytes(n));
  const vid_t *ORB = wk.ORB;
  const int tid = threadIdx.x, nt = NT;
  __shared__ int s_pair;

  for (;;) {
    if (tid == 0) s_pair = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int p = s_pair;
    __syncthreads();
    if (p >= npairs) break;

    const int32_t *ps = src + (long long)p * n, *pd = dst + (long long)p * n;
    for (int i = tid; i < n; i += nt) {
      CUR[i] = (int16_t)ps[i];
      int g = pd[i];
      SDST[(g > 0 ? g : -g) - 1] = (int16_t)(g > 0 ? (i + 1) : -(i + 1));
    }
    __syncthreads();

    int ns = 0, code = GRSR_SCEN_OK;
    unsigned long long nevals = 0;
    RhoCurrent rho0; rho0.cur = CUR; rho0.sdst = SDST;
    for (;;) {
      // rho0 and its inverse: G0 and every trial of this step are written from these two arrays
      for (int i = tid; i < n; i += nt) {
        int r = rho0(i);
        RHO0[i] = (int16_t)r;
        INV0[(r > 0 ? r : -r) - 1] = (int16_t)(r > 0 ? (i + 1) : -(i + 1));
      }
      __syncthreads();
      EndsG0 eg; eg.inv0 = INV0;
      const Stats g0 = dist_core<NT, EPT, true>(wk, n, eg);   // G/opt_scenario.c:229
      if (g0.d == 0) break;
      if (ns >= max_steps) { code = GRSR_SCEN_OVERFLOW; break; }

      // black edges that are breakpoints, ascending: edge k = (2k, 2k+1) sits left of position k.
      // A reversal (i, j) cuts edges i and j+1 (the tests of G/opt_scenario.c:538-561).
      const int nbp = blk_compact(n + 1, [&](int k, unsigned &val) -> bool {
        if (ORB[2 * k] == kNone) return false;
        val = (unsigned)k; return true;
      }, BPL, wk.red);
      const bool filter = (g0.h == 0);                        // G/opt_scenario.c:951

      unsigned lastkey = kNoKey;
      int fi = -1, fj = -1;
      for (;;) {
        // smallest key (j - i, i) greater than lastkey among admissible candidates
        unsigned best = kNoKey;
        for (int t = tid; t < nbp; t += nt) {
          int e1 = BPL[t];
          int s = t + 1;
          if (lastkey != kNoKey) {
            int g = (int)(lastkey >> kKeyShift), i0 = (int)(lastkey & ((1u << kKeyShift) - 1u));
            int lb = e1 + 1 + g + (e1 > i0 ? 0 : 1);          // smallest admissible right edge
            int lo = t + 1, hi = nbp;
            while (lo < hi) { int mid = (lo + hi) >> 1; if ((int)BPL[mid] < lb) lo = mid + 1; else hi = mid; }
            s = lo;
          }
          if (filter) {
            unsigned want = ORB[2 * e1 + 1];
            while (s < nbp && ORB[2 * (int)BPL[s]] != want) s++;
          }
          if (s < nbp) {
            int e2 = BPL[s];
            best = min(best, ((unsigned)(e2 - e1 - 1) << kKeyShift) | (unsigned)e1);
          }
        }
        best = blk_min_u32(best, wk.red);
        if (best == kNoKey) break;
        const int i = (int)(best & ((1u << kKeyShift) - 1u));
        const int j = i + (int)(best >> kKeyShift);
        EndsTrial et; et.rho0 = RHO0; et.i = i; et.j = j;
        const Stats t1 = dist_core<NT, EPT, false>(wk, n, et);  // G/opt_scenario.c:974-986
        nevals++;
        if (t1.d == g0.d - 1) { fi = i; fj = j; break; }
        lastkey = best;
      }
      if (fi < 0) { code = GRSR_SCEN_STUCK; break; }

      // current <- trial (G/opt_scenario.c:283-285)
      for (int k = tid; 2 * k <= fj - fi; k += nt) {
        int a = fi + k, b = fj - k;
        int16_t x = CUR[a], y = CUR[b];
        CUR[a] = (int16_t)(-y); CUR[b] = (int16_t)(-x);
      }
      if (tid == 0) {
        int32_t *o = steps + 2 * ((long long)p * max_steps + ns);
        o[0] = fi; o[1] = fj;
      }
      ns++;
      __syncthreads();
    }
    if (tid == 0) {
      nsteps[p] = ns; status[p] = code;
      if (evals) evals[p] = nevals;
    }
    __syncthreads();
  }
}

} // namespace grsr
