// grsr_stepwise.cuh -- candidate-parallel scenario steps for the case "few pairs, hurdles present".
//
// While the breakpoint graph has hurdles the reference switches its filter off and gives EVERY
// breakpoint pair a full distance evaluation, in (j - i, i) order, until one yields d - 1
// (G/opt_scenario.c:951, 527-576): hundreds to thousands of independent evaluations per step.  The
// persistent scenario kernel evaluates them one after another inside the pair's block, which is
// the right thing when thousands of pairs keep the chip busy.  With only a few pairs in flight this
// file spreads the evaluations of ONE step over the whole GPU instead:
//
//   step_prepare_kernel   one block: G0 of the current genome, then the next batch of candidates in
//                         key order -- whole diagonals j - i = const, i.e. every admissible candidate
//                         with key in (lastkey, bound] -- written to global memory
//   trial_batch_kernel    (grsr_scenario.cuh) one block per candidate: d(trial, destination)
//   step_select_kernel    one block: the successful candidate with the SMALLEST key is the
//                         reference's choice; apply it, or move lastkey to the end of the batch
//
// Once h + f = 0 every accepted step keeps it 0 (d' = d - 1 forces the cycle count up by one and
// h' + f' = 0), so the host hands the pair over to the persistent kernel for the hurdle-free tail.
#pragma once
#include "grsr_scenario.cuh"

namespace grsr {

struct StepState {
  int32_t d, h, f, nbp;            // G0 of the current genome (written by prepare)
  int32_t ncand;                   // candidates in this batch
  int32_t found;                   // select: 1 = a step was applied, 0 = batch exhausted without success
  int32_t nsteps;                  // steps emitted so far
  int32_t status;                  // GRSR_SCEN_*
  unsigned long long lastkey;      // every candidate with key <= lastkey of this step has been rejected
  unsigned long long batch_end;    // largest key of the current batch
  unsigned long long evals;        // full trial evaluations made so far (whole batches are scored)
  int32_t target;                  // candidates wanted in the next batch of a hurdle step (0: kStepFirstBatch)
  int32_t failed;                  // candidates scored without success in the earlier batches of this step
};

static const int kStepFirstBatch = 256;   // first batch of a hurdle step; x4 after every batch without success

// One block.  cur/dst: the pair (global, int32).  cands: capacity cap pairs (i, j).
template <int NT, int EPT>
__global__ void __launch_bounds__(NT, GRSR_MIN_BLOCKS(NT, EPT))
step_prepare_kernel(const int32_t *__restrict__ cur, const int32_t *__restrict__ dst, int n,
                    StepState *st, int32_t *__restrict__ cands, int cap) {
  GRSR_DYN_SMEM(smem_raw);
  Work wk;
  work_carve(wk, smem_raw, n, NT * EPT, true);
  GRSR_PHASE_INIT(); GRSR_PHASE_BASE(8);
  unsigned char *extra = smem_raw + work_bytes(n, NT * EPT, true);
  int16_t *CUR = (int16_t *)extra;
  int16_t *SDST = (int16_t *)(extra + scen_arr_bytes(n));
  int16_t *INV0 = (int16_t *)(extra + 2 * scen_arr_bytes(n));
  uint16_t *BPL = (uint16_t *)(extra + 4 * scen_arr_bytes(n));
  const uint16_t *ORB = wk.ORB;
  const int tid = threadIdx.x;
  const unsigned long long kNoKey = ~0ull;
  volatile unsigned *shared_u = (volatile unsigned *)&wk.red[33];

  for (int i = tid; i < n; i += NT) {
    CUR[i] = (int16_t)cur[i];
    int g = dst[i];
    SDST[(g > 0 ? g : -g) - 1] = (int16_t)(g > 0 ? (i + 1) : -(i + 1));
  }
  __syncthreads();
  for (int i = tid; i < n; i += NT) {
    int g = CUR[i];
    int s = SDST[(g > 0 ? g : -g) - 1];
    int r = g > 0 ? s : -s;
    INV0[(r > 0 ? r : -r) - 1] = (int16_t)(r > 0 ? (i + 1) : -(i + 1));
  }
  __syncthreads();
  EndsG0<int16_t> eg; eg.inv0 = INV0;
  const Stats g0 = dist_core<NT, EPT, true>(wk, n, eg);
  if (g0.d == 0) {
    if (tid == 0) { st->d = 0; st->h = 0; st->f = 0; st->nbp = 0; st->ncand = 0; }
    return;
  }
  const int nbp = blk_compact(n + 1, [&](int k, unsigned &val) -> bool {
    if (ORB[2 * k] == 0xFFFFu) return false;
    val = (unsigned)k; return true;
  }, BPL, wk.red);
  const unsigned long long lastkey = st->lastkey;
  const int g_last = (lastkey == kNoKey) ? -1 : (int)(lastkey >> 32);
  // low word 0xFFFFFFFF = "the whole diagonal g_last has been scored" (batch_end of a hurdle batch):
  // it must compare greater than every position, so it stays unsigned in a 64-bit integer
  const long long i_last = (lastkey == kNoKey) ? -1ll : (long long)(lastkey & 0xFFFFFFFFull);
  __syncthreads();

  int ncand = 0;
  unsigned long long batch_end = lastkey;
  if (g0.h == 0) {
    // hurdle-free graph: the reference's parity filter applies; one candidate per round, the
    // smallest admissible key above lastkey (same search as the persistent kernel)
    if (tid == 0) *shared_u = 0xFFFFFFFFu;
    __syncthreads();
    unsigned long long best = kNoKey;
    for (int t = tid; t < nbp; t += NT) {
      const int e1 = (int)BPL[t];
      int s = t + 1;
      if (lastkey != kNoKey) {
        const int lb = e1 + 1 + g_last + (e1 > i_last ? 0 : 1);
        int lo = t + 1, hi = nbp;
        while (lo < hi) { int mid = (lo + hi) >> 1; if ((int)BPL[mid] < lb) lo = mid + 1; else hi = mid; }
        s = lo;
      }
      const unsigned want = ORB[2 * e1 + 1];
      int e2 = -1;
      for (; s < nbp; s++) {
        const int c = (int)BPL[s];
        if ((unsigned)(c - e1 - 1) > *shared_u) break;
        if (ORB[2 * c] == want) { e2 = c; break; }
      }
      if (e2 >= 0) {
        const unsigned len0 = (unsigned)(e2 - e1 - 1);
        const unsigned long long key = ((unsigned long long)len0 << 32) | (unsigned)e1;
        best = (key < best) ? key : best;
        atomicMin((unsigned *)&wk.red[33], len0);
      }
    }
    best = blk_min_u64(best, wk.red);
    if (best != kNoKey) {
      if (tid == 0) { cands[0] = (int)(best & 0xFFFFFFFFull); cands[1] = cands[0] + (int)(best >> 32); }
      ncand = 1; batch_end = best;
    }
  } else {
    // hurdles: every pair of breakpoint edges (e1 < e2) is a candidate (i = e1, j = e2 - 1), key
    // (e2 - e1 - 1, e1).  Diagonals are appended in key order, smallest first, until the batch holds
    // about `target` candidates; a diagonal that does not fit is cut after its first entries in e1 order
    // (ordered compaction), so a batch is always "every candidate with key in (lastkey, batch_end]".
    // target starts small -- the first success is usually among the first candidates, and scoring a
    // whole diagonal of n entries for it wastes thousands of evaluations -- and is quadrupled by
    // step_select_kernel after a batch without success.  Empty diagonals are skipped by a block-wide
    // minimum of every edge's next gap.
    int target = st->target;
    if (target < 1) target = kStepFirstBatch;
    if (target > cap) target = cap;
    uint16_t *LST = wk.A1;                           // free once G0 is built: one diagonal's e1 list
    int g = (lastkey == kNoKey) ? 0 : g_last;        // diagonal to resume in
    bool partial = (lastkey != kNoKey);              // first diagonal: only entries with e1 > i_last
    for (;;) {
      // smallest gap >= g that occurs (for the resumed diagonal: with e1 > i_last)
      unsigned long long mg = kNoKey;
      for (int t = tid; t < nbp; t += NT) {
        const int e1 = (int)BPL[t];
        const int need = e1 + 1 + g + ((partial && e1 <= i_last) ? 1 : 0);
        int lo = t + 1, hi = nbp;
        while (lo < hi) { int mid = (lo + hi) >> 1; if ((int)BPL[mid] < need) lo = mid + 1; else hi = mid; }
        if (lo < nbp) { unsigned long long gg = (unsigned long long)((int)BPL[lo] - e1 - 1); mg = gg < mg ? gg : mg; }
      }
      mg = blk_min_u64(mg, wk.red);
      if (mg == kNoKey) break;                       // no candidate left at all
      const bool resumed = partial && ((int)mg == g);
      g = (int)mg;
      // the entries of diagonal g (after i_last if it is the resumed one), ascending e1
      const int cnt = blk_compact(nbp, [&](int t, unsigned &val) -> bool {
        const int e1 = (int)BPL[t];
        if (resumed && e1 <= i_last) return false;
        const int want = e1 + 1 + g;
        int lo = t + 1, hi = nbp;
        while (lo < hi) { int mid = (lo + hi) >> 1; if ((int)BPL[mid] < want) lo = mid + 1; else hi = mid; }
        if (lo >= nbp || (int)BPL[lo] != want) return false;
        val = (unsigned)e1; return true;
      }, LST, wk.red);
      const int room = min(target, cap) - ncand;
      const int take = cnt < room ? cnt : room;
      for (int q = tid; q < take; q += NT) {
        const int e1 = (int)LST[q];
        cands[2 * (ncand + q)] = e1; cands[2 * (ncand + q) + 1] = e1 + g;
      }
      ncand += take;
      batch_end = (take == cnt) ? (((unsigned long long)(unsigned)g << 32) | 0xFFFFFFFFull)   // the whole diagonal is in
                                : (((unsigned long long)(unsigned)g << 32) | (unsigned)LST[take - 1]);
      __syncthreads();
      partial = false;
      g++;
      if (ncand >= target || ncand >= cap || g > n) break;
    }
  }
  if (tid == 0) {
    st->d = g0.d; st->h = g0.h; st->f = g0.f; st->nbp = nbp; st->ncand = ncand; st->batch_end = batch_end;
  }
}

// One block.  d2[c] = distance of candidate c.  Applies the successful candidate with the smallest
// key to cur (global) and appends it to steps, or records that the batch is exhausted.
__global__ void __launch_bounds__(256)
step_select_kernel(int32_t *cur, int n, StepState *st, const int32_t *__restrict__ cands,
                   const int32_t *__restrict__ d2, int32_t *__restrict__ steps, int max_steps) {
  __shared__ unsigned long long red[34];
  const int tid = threadIdx.x;
  const int ncand = st->ncand, d = st->d;
  if (d == 0) return;
  if (tid == 0) st->evals += (unsigned long long)ncand;
  unsigned long long best = ~0ull;
  for (int c = tid; c < ncand; c += blockDim.x) {
    if (d2[c] == d - 1) {
      const unsigned i = (unsigned)cands[2 * c], j = (unsigned)cands[2 * c + 1];
      const unsigned long long key = ((unsigned long long)(j - i) << 32) | i;
      best = key < best ? key : best;
    }
  }
  best = blk_min_u64(best, red);
  if (best == ~0ull) {
    if (tid == 0) {
      st->found = 0;
      if (ncand == 0) st->status = GRSR_SCEN_STUCK;          // nothing left to try
      st->lastkey = st->batch_end;
      st->failed += ncand;
      const int t0 = st->target < 1 ? kStepFirstBatch : st->target;
      st->target = t0 < (1 << 24) ? 4 * t0 : t0;             // nothing in this batch: a bigger one next
    }
    return;
  }
  const int fi = (int)(best & 0xFFFFFFFFull), fj = fi + (int)(best >> 32);
  if (st->nsteps >= max_steps) { if (tid == 0) { st->status = GRSR_SCEN_OVERFLOW; st->found = 0; } return; }
  // how many candidates the reference's sequential scan would have scored for this step: those of the
  // failed batches plus the ones of this batch up to the winner.  The next step starts with a batch of
  // one and a half times that (consecutive steps of a hurdle-rich scenario need similar numbers).
  unsigned long long upto = 0;
  for (int c = tid; c < ncand; c += blockDim.x) {
    const unsigned i = (unsigned)cands[2 * c], j = (unsigned)cands[2 * c + 1];
    upto += ((((unsigned long long)(j - i) << 32) | i) <= best) ? 1ull : 0ull;
  }
  upto = blk_sum_u64(upto, red);
  for (int k = tid; 2 * k <= fj - fi; k += blockDim.x) {
    const int a = fi + k, b = fj - k;
    const int32_t x = cur[a], y = cur[b];
    cur[a] = -y; cur[b] = -x;
  }
  if (tid == 0) {
    steps[2 * st->nsteps] = fi; steps[2 * st->nsteps + 1] = fj;
    st->nsteps = st->nsteps + 1;
    st->found = 1;
    st->lastkey = ~0ull;
    const long long needed = (long long)st->failed + (long long)upto;
    const long long next = needed + needed / 2;
    st->target = (int32_t)(next < kStepFirstBatch ? kStepFirstBatch : (next > (1 << 24) ? (1 << 24) : next));
    st->failed = 0;
  }
}

} // namespace grsr
