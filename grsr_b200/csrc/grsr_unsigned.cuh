// grsr_unsigned.cuh -- device half of the exact unsigned reversal distance (`grimm -u`, SURVEY.md
// section 8f row N2): the 2^k sign assignments of the singletons, each scored as the reference scores
// it, spread over the whole GPU (one block per assignment, grid-stride).
//
//   loop_unsigned_perm  G/unsigned.c:836-869   depth-first over the singletons, toggling without
//                                              toggling back: the leaves are visited in reflected
//                                              Gray-code order -- leaf t has singleton j flipped iff
//                                              bit (k-1-j) of t ^ (t >> 1) is set
//   sign_2strips        G/unsigned.c:885-1095  graph of (g1, g2 with these signs); in every UNORIENTED
//                                              component the first 2-strip (in position order) whose
//                                              ends lie on one cycle is re-signed; distance of the result
//   best_d / best_g2    G/unsigned.c:1058-1066 strict "<": the FIRST leaf reaching the minimum wins --
//                                              here a 64-bit atomicMin over (distance << 32 | leaf)
//
// Vertices are the doubled positions of pi = g2, as everywhere else (dist_core): the 2-strip at positions
// (r, r+1) owns vertices 2r+1 .. 2r+4 (G/unsigned.c:1007-1020).
#pragma once
#include "grsr_core.cuh"

namespace grsr {

__host__ __device__ inline size_t unsigned_extra_bytes(int n) {
  return stage_bytes(n) + ((((size_t)n * 2) + 15) & ~(size_t)15);     // tail u16[n] | flags (doubleton flips) u16[n]
}

// gamma = g1 (global), pi0 = g2 with canonical strip signs (global), sidx[p] = index of the singleton at
// position p or -1, doub[i] = position of 2-strip i.  Leaves [t0, t0 + count) are scored; best receives
// min (d << 32 | t).  emit != null (count == 1): the fully signed g2 of leaf t0 is written to emit[n].
template <int NT, int EPT>
__global__ void __launch_bounds__(NT, GRSR_MIN_BLOCKS(NT, EPT))
unsigned_signs_kernel(const int32_t *__restrict__ g1, const int32_t *__restrict__ pi0,
                      const int32_t *__restrict__ sidx, const int32_t *__restrict__ doub, int n, int k, int m,
                      unsigned long long t0, unsigned long long count, unsigned long long *best,
                      int32_t *emit) {
  GRSR_DYN_SMEM(smem_raw);
  Work wk;
  work_carve(wk, smem_raw, n, NT * EPT, false);
  GRSR_PHASE_INIT(); GRSR_PHASE_BASE(8);
  uint16_t *tail = (uint16_t *)(smem_raw + work_bytes(n, NT * EPT, false));
  uint16_t *FL = (uint16_t *)((unsigned char *)tail + stage_bytes(n));       // FL[p] = 1: sign at position p flipped by sign_2strips
  const int tid = threadIdx.x;
  uint32_t *W = (uint32_t *)wk.W;
  constexpr unsigned kNone = none_of<uint16_t>();

  for (unsigned long long q = blockIdx.x; q < count; q += gridDim.x) {
    const unsigned long long t = t0 + q;
    const unsigned long long gray = t ^ (t >> 1);
    // pi of this leaf: tail[a-1] = vertex where gene +a begins (2p+1 if pi holds +a at p, else 2p+2)
    for (int p = tid; p < n; p += NT) {
      int x = pi0[p];
      const int s = sidx[p];
      if (s >= 0 && s < k && ((gray >> (k - 1 - s)) & 1ull)) x = -x;
      tail[(x > 0 ? x : -x) - 1] = (uint16_t)(x > 0 ? 2 * p + 1 : 2 * p + 2);
      FL[p] = 0;
    }
    __syncthreads();
    EndsFromRows ends; ends.gamma = g1; ends.tail = tail;
    // graph of the leaf (invdist_noncircular_G, G/unsigned.c:939): components only
    Stats st = dist_core<NT, EPT, false, EndsFromRows, true>(wk, n, ends);
    int d = st.b - st.c;                                   // every component oriented: h = f = 0, nothing to re-sign
    if (st.h) {
      // W[c] = component of cycle c, W[comp+1] = 1 if oriented (hurdles_slow_path, components only).
      // Mark every unoriented component "no 2-strip yet", then let the 2-strips bid with their index.
      for (int kk = tid; kk <= n; kk += NT) {
        const unsigned r = 2u * kk;
        if (wk.A2[r] == r && W[r] == r) W[r + 1] = W[r + 1] ? 0xFFFFFFFFu : 0xFFFFFFFEu;
      }
      __syncthreads();
      int any = 0;
      for (int i = tid; i < m; i += NT) {
        const unsigned v = (unsigned)doub[i];
        const unsigned c1 = wk.A2[2 * v + 1];
        if (c1 != kNone && c1 == wk.A2[2 * v + 4]) {       // cc >= 0 and both ends on one cycle (:1013-1020)
          const unsigned comp = W[c1];
          if (W[comp + 1] != 0xFFFFFFFFu) { atomicMin(&W[comp + 1], (uint32_t)i); any = 1; }
        }
      }
      any = __syncthreads_or(any);
      if (any) {
        for (int i = tid; i < m; i += NT) {
          const unsigned v = (unsigned)doub[i];
          const unsigned c1 = wk.A2[2 * v + 1];
          if (c1 != kNone && c1 == wk.A2[2 * v + 4] && W[W[c1] + 1] == (uint32_t)i) { FL[v] = 1; FL[v + 1] = 1; }
        }
        __syncthreads();
        for (int p = tid; p < n; p += NT) {
          if (FL[p]) {                                      // the gene at p changes sign: its two ends swap
            int x = pi0[p];
            const unsigned a = (unsigned)(x > 0 ? x : -x);
            tail[a - 1] = (uint16_t)(((tail[a - 1] - 1u) ^ 1u) + 1u);
          }
        }
        __syncthreads();
      }
      // all signs set: the distance (invdist_noncircular, G/unsigned.c:1044)
      st = dist_core<NT, EPT, false>(wk, n, ends);
      d = st.d;
    }
    if (tid == 0 && best) atomicMin(best, ((unsigned long long)(unsigned)d << 32) | (t & 0xFFFFFFFFull));
    if (emit) {
      for (int p = tid; p < n; p += NT) {
        int x = pi0[p];
        const int s = sidx[p];
        if (s >= 0 && s < k && ((gray >> (k - 1 - s)) & 1ull)) x = -x;
        emit[p] = FL[p] ? -x : x;
      }
    }
    __syncthreads();
  }
}

} // namespace grsr
