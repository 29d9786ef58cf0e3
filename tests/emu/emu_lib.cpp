// emu_lib.cpp -- TEST INFRASTRUCTURE ONLY: runs the kernels of grsr_core.cuh / grsr_scenario.cuh on
// the CPU through the fiber SIMT emulator (cuda_emu.hpp) so kernel logic can be checked where no GPU
// exists.  Built by tests/emu/Makefile into tests/emu/libgrsr_emu.so; loaded only by tests/.
#define GRSR_EMU 1
#include "cuda_emu.hpp"
#ifdef GRSR_WALK_STATS
extern "C" { unsigned long long g_walk_stats[64]; }
#endif
#include "../../grsr_b200/csrc/grsr_core.cuh"
#include "../../grsr_b200/csrc/grsr_scenario.cuh"
#include "../../grsr_b200/csrc/grsr_stepwise.cuh"
#include "../../grsr_b200/csrc/grsr_walk.cuh"
#include "../../grsr_b200/csrc/grsr_unsigned.cuh"
#include "../../grsr_b200/csrc/grsr_unsigned_host.hpp"
#include <vector>

extern "C" int emu_dist_pairs(const int32_t *genomes, int G, int n, const int32_t *pairs,
                              long long npairs, int32_t *out, int stride, int threads, int grid) {
  std::vector<int32_t> sinv((size_t)G * n);
  long long total = (long long)G * n;
  int32_t *sp = sinv.data();
  emu::launch(2, 64, 0, [&] { grsr::signed_inverse_kernel(genomes, sp, total, n); });
  int t, e;
  grsr::pick_shape(n, threads, &t, &e);
  size_t smem = grsr::work_bytes(n, t * e, false) + grsr::stage_bytes(n);
  GRSR_DISPATCH_SHAPE(t, e, emu::launch((unsigned)grid, (unsigned)t, smem, [&] {
    grsr::dist_pairs_kernel<kNt, kEpt>(genomes, sp, n, pairs, npairs, out, stride);
  }));
  return 0;
}

extern "C" int emu_scenario(const int32_t *src, const int32_t *dst, int npairs, int n, int32_t *steps,
                            int32_t *nsteps, int max_steps, int32_t *status,
                            unsigned long long *evals, int threads, int grid) {
  unsigned int ticket = 0;
  int t, e;
  grsr::pick_shape(n, threads, &t, &e);
  size_t smem = grsr::work_bytes(n, t * e, true) + grsr::scen_extra_bytes(n);
  GRSR_DISPATCH_SHAPE(t, e, emu::launch((unsigned)grid, (unsigned)t, smem, [&] {
    grsr::scenario_kernel<kNt, kEpt>(src, dst, npairs, n, steps, nsteps, max_steps, status, evals, grsr::PairQueue{&ticket, 0, 1}, 0);
  }));
  return 0;
}

// the global-memory ("large") build, run at small n to check its logic
extern "C" int emu_dist_pairs_large(const int32_t *genomes, int G, int n, const int32_t *pairs,
                                    long long npairs, int32_t *out, int stride, int threads, int grid) {
  std::vector<int32_t> sinv((size_t)G * n);
  long long total = (long long)G * n;
  int32_t *sp = sinv.data();
  emu::launch(2, 64, 0, [&] { grsr::signed_inverse_kernel(genomes, sp, total, n); });
  std::vector<unsigned char> scratch(grsr::large_work_bytes(n, false) * (size_t)grid + 64);
  unsigned char *sc = scratch.data();
  emu::launch((unsigned)grid, (unsigned)threads, 0, [&] {
    grsr::dist_pairs_large_kernel(genomes, sp, n, pairs, npairs, out, stride, sc);
  });
  return 0;
}

extern "C" int emu_scenario_large(const int32_t *src, const int32_t *dst, int npairs, int n, int32_t *steps,
                                  int32_t *nsteps, int max_steps, int32_t *status,
                                  unsigned long long *evals, int threads, int grid) {
  unsigned int ticket = 0;
  std::vector<unsigned char> scratch(grsr::scen_large_bytes(n) * (size_t)grid + 64);
  unsigned char *sc = scratch.data();
  emu::launch((unsigned)grid, (unsigned)threads, 0, [&] {
    grsr::scenario_large_kernel(src, dst, npairs, n, steps, nsteps, max_steps, status, evals, grsr::PairQueue{&ticket, 0, 1}, sc);
  });
  return 0;
}

// the cluster build (several blocks per pair), run at small n
extern "C" int emu_dist_pairs_cluster(const int32_t *genomes, int G, int n, const int32_t *pairs,
                                      long long npairs, int32_t *out, int stride, int threads,
                                      int nclusters, int cs) {
  std::vector<int32_t> sinv((size_t)G * n);
  long long total = (long long)G * n;
  int32_t *sp = sinv.data();
  emu::launch(2, 64, 0, [&] { grsr::signed_inverse_kernel(genomes, sp, total, n); });
  std::vector<unsigned char> scratch(grsr::large_work_bytes(n, false) * (size_t)nclusters + 64);
  std::vector<unsigned long long> xch((size_t)grsr::kXchWords * nclusters, 0xCDCDCDCDCDCDCDCDull);
  unsigned char *sc = scratch.data();
  unsigned long long *xp = xch.data();
  emu::launch_clusters((unsigned)nclusters, (unsigned)cs, (unsigned)threads, 34 * 8, [&] {
    grsr::dist_pairs_cluster_kernel(genomes, sp, n, pairs, npairs, out, stride, sc, xp);
  });
  return 0;
}

extern "C" int emu_scenario_cluster(const int32_t *src, const int32_t *dst, int npairs, int n, int32_t *steps,
                                    int32_t *nsteps, int max_steps, int32_t *status,
                                    unsigned long long *evals, int threads, int nclusters, int cs) {
  unsigned int ticket = 0;
  std::vector<unsigned char> scratch(grsr::scen_large_bytes(n) * (size_t)nclusters + 64);
  std::vector<unsigned long long> xch((size_t)grsr::kXchWords * nclusters, 0xCDCDCDCDCDCDCDCDull);
  unsigned char *sc = scratch.data();
  unsigned long long *xp = xch.data();
  emu::launch_clusters((unsigned)nclusters, (unsigned)cs, (unsigned)threads, 34 * 8, [&] {
    grsr::scenario_cluster_kernel(src, dst, npairs, n, steps, nsteps, max_steps, status, evals, grsr::PairQueue{&ticket, 0, 1}, sc, xp);
  });
  return 0;
}

extern "C" int emu_trial_distances(const int32_t *src, const int32_t *dst, int n, const int32_t *cands,
                                   long long ncand, int32_t *out, int threads, int grid) {
  int t, e;
  grsr::pick_shape(n, threads, &t, &e);
  size_t smem = grsr::work_bytes(n, t * e, false) + 2 * grsr::scen_arr_bytes(n);
  GRSR_DISPATCH_SHAPE(t, e, emu::launch((unsigned)grid, (unsigned)t, smem, [&] {
    grsr::trial_batch_kernel<kNt, kEpt>(src, dst, n, cands, ncand, out, (const int32_t *)nullptr);
  }));
  return 0;
}

// candidate-parallel steps while hurdles exist, then the persistent kernel for the rest -- the same
// host loop as scenario_on_device() in grsr_capi.cu, with every launch going through the emulator
static long long g_stepwise_evals = 0;   // candidates scored by the candidate-parallel phase of the last call
extern "C" long long emu_last_stepwise_evals() { return g_stepwise_evals; }

extern "C" int emu_scenario_stepwise(const int32_t *src, const int32_t *dst, int npairs, int n,
                                     int32_t *steps, int32_t *nsteps, int max_steps, int32_t *status,
                                     int threads, int cap, int *stepwise_steps) {
  int t, e;
  grsr::pick_shape(n, threads, &t, &e);
  size_t smem_prep = grsr::work_bytes(n, t * e, true) + grsr::scen_extra_bytes(n);
  size_t smem_trial = grsr::work_bytes(n, t * e, false) + 2 * grsr::scen_arr_bytes(n);
  std::vector<int32_t> cur((size_t)npairs * n), cands((size_t)cap * 2), d2((size_t)cap);
  memcpy(cur.data(), src, sizeof(int32_t) * (size_t)npairs * n);
  *stepwise_steps = 0;
  g_stepwise_evals = 0;
  for (int p = 0; p < npairs; p++) {
    grsr::StepState st;
    memset(&st, 0, sizeof st);
    st.lastkey = ~0ull;
    int32_t *cp = cur.data() + (size_t)p * n;
    const int32_t *dp = dst + (size_t)p * n;
    int32_t *sp = steps + 2 * (size_t)p * max_steps;
    for (;;) {
      grsr::StepState *stp = &st;
      int32_t *cd = cands.data(), *dd = d2.data();
      GRSR_DISPATCH_SHAPE(t, e, emu::launch(1, (unsigned)t, smem_prep, [&] {
        grsr::step_prepare_kernel<kNt, kEpt>(cp, dp, n, stp, cd, cap);
      }));
      GRSR_DISPATCH_SHAPE(t, e, emu::launch(3, (unsigned)t, smem_trial, [&] {
        grsr::trial_batch_kernel<kNt, kEpt>(cp, dp, n, cd, 0, dd, &stp->ncand);
      }));
      emu::launch(1, 64, 0, [&] { grsr::step_select_kernel(cp, n, stp, cd, dd, sp, max_steps); });
      if (st.d == 0 || st.status != 0) break;
      if (st.h == 0) break;                      // hurdle-free from here on: persistent kernel
    }
    nsteps[p] = st.nsteps; status[p] = st.status;
    *stepwise_steps += st.nsteps;
    g_stepwise_evals += (long long)st.evals;
  }
  unsigned int ticket = 0;
  size_t smem = smem_prep;
  const int32_t *cs = cur.data();
  std::vector<int32_t> st2((size_t)npairs, 0);
  int32_t *st2p = st2.data();
  GRSR_DISPATCH_SHAPE(t, e, emu::launch(2, (unsigned)t, smem, [&] {
    grsr::scenario_kernel<kNt, kEpt>(cs, dst, npairs, n, steps, nsteps, max_steps, st2p,
                                     (unsigned long long *)nullptr, grsr::PairQueue{&ticket, 0, 1}, 1);
  }));
  for (int p = 0; p < npairs; p++) if (status[p] == 0) status[p] = st2[p];
  return 0;
}

// two-tier distance path: warp-per-pair walk kernels (shared = -1: walk_order_kernel decides between
// the private-row and the shared-row kernel, as grsr_dist_pairs_dev does; 0 / 1 force one), then the
// block-per-pair kernel on the pairs they deferred; *deferred = how many pairs took the second tier
extern "C" int emu_dist_pairs_walk(const int32_t *genomes, int G, int n, const int32_t *pairs,
                                   long long npairs, int32_t *out, int stride, int wpb, int grid,
                                   int *deferred, int shared) {
  std::vector<int32_t> sinv((size_t)G * n);
  long long total = (long long)G * n;
  int32_t *sp = sinv.data();
  emu::launch(2, 64, 0, [&] { grsr::signed_inverse_kernel(genomes, sp, total, n); });
  int shift = grsr::walk_shape_arg(grsr::walk_shift(n), grsr::kWalkIdle);
  std::vector<int32_t> dl((size_t)npairs + 1, -1);
  unsigned dc = 0, flag = 0, tk1 = 0, tk2 = 0;
  int32_t *dlp = dl.data();
  unsigned *dcp = &dc, *fp = (shared < 0) ? &flag : nullptr, *t1 = &tk1, *t2 = &tk2;
  if (fp) emu::launch(3, 64, 0, [&] { grsr::walk_order_kernel(pairs, npairs, fp); });
  if (shared != 0) {
    size_t smem = 16 + grsr::walk_tail_bytes(n) + grsr::walk_private_bytes(n) * (size_t)wpb;
    emu::launch((unsigned)grid, (unsigned)(32 * wpb), smem, [&] {
      grsr::dist_walk_shared_kernel(genomes, sp, n, pairs, npairs, out, stride, shift, dlp, dcp, fp, t1);
    });
  }
  if (shared != 1) {
    size_t smem = grsr::walk_warp_bytes(n) * (size_t)wpb;
    emu::launch((unsigned)grid, (unsigned)(32 * wpb), smem, [&] {
      grsr::dist_walk_kernel(genomes, sp, n, pairs, npairs, out, stride, shift, dlp, dcp, fp, t2);
    });
  }
  *deferred = (int)dc;
  int t, e;
  grsr::pick_shape(n, 0, &t, &e);
  size_t smem2 = grsr::work_bytes(n, t * e, false) + grsr::stage_bytes(n);
  GRSR_DISPATCH_SHAPE(t, e, emu::launch(3, (unsigned)t, smem2, [&] {
    grsr::dist_deferred_kernel<kNt, kEpt>(genomes, sp, n, pairs, dlp, dcp, out, stride);
  }));
  return shared < 0 ? (int)((unsigned long long)flag * 64ull <= (unsigned long long)npairs) : 0;
}

// The walk kernel's node grid (grsr_walk.cuh WalkGrid), for the bank / mapping properties test.
extern "C" int emu_walk_grid_vertex(int shift, int k) { return (int)grsr::WalkGrid(shift).vertex((unsigned)k); }
extern "C" int emu_walk_grid_is_node(int shift, int x) { return grsr::WalkGrid(shift).is_node((unsigned)x) ? 1 : 0; }

// grsr_unsigned_dist as the library computes it (host preprocessing of grsr_unsigned_host.hpp + the
// sign-enumeration kernel), with the kernel run through the emulator.  info7 = {dist, approx, passes,
// num_sing[0..1], num_doub[0..1]}.
extern "C" int emu_unsigned_dist(const int32_t *g1, const int32_t *g2, int n, int circular, int32_t *out_g1,
                                 int32_t *out_g2, int32_t *info7, int threads, int grid) {
  int t, e;
  grsr::pick_shape(n, threads, &t, &e);
  const size_t smem = grsr::work_bytes(n, t * e, false) + grsr::unsigned_extra_bytes(n);
  auto eval = [&](const int32_t *gamma, grsr::UnsignedPass &P, int k, int &best_d, std::vector<int32_t> &best_pi) -> int {
    const int m = (int)P.doubletons.size();
    std::vector<int32_t> sidx((size_t)n, -1), emit((size_t)n, 0);
    for (size_t j = 0; j < P.singletons.size(); j++) sidx[(size_t)P.singletons[j]] = (int32_t)j;
    const unsigned long long leaves = 1ull << k;
    unsigned long long best = ~0ull, *bp = &best;
    const int32_t *pi0 = P.pi0.data(), *sx = sidx.data(), *db = P.doubletons.data();
    GRSR_DISPATCH_SHAPE(t, e, emu::launch((unsigned)grid, (unsigned)t, smem, [&] {
      grsr::unsigned_signs_kernel<kNt, kEpt>(gamma, pi0, sx, db, n, k, m, 0ull, leaves, bp, (int32_t *)nullptr);
    }));
    const int d = (int)(best >> 32);
    if (d < best_d) {
      const unsigned long long tt = best & 0xFFFFFFFFull;
      int32_t *ep = emit.data();
      GRSR_DISPATCH_SHAPE(t, e, emu::launch(1, (unsigned)t, smem, [&] {
        grsr::unsigned_signs_kernel<kNt, kEpt>(gamma, pi0, sx, db, n, k, m, tt, 1ull, (unsigned long long *)nullptr, ep);
      }));
      best_pi = emit;
      best_d = d;
    }
    return 0;
  };
  grsr::UnsignedResult R;
  int rc = grsr::unsigned_distance(g1, g2, n, circular != 0, eval, R);
  if (rc) return rc;
  info7[0] = R.dist; info7[1] = R.approx; info7[2] = R.passes;
  info7[3] = R.num_sing[0]; info7[4] = R.num_sing[1]; info7[5] = R.num_doub[0]; info7[6] = R.num_doub[1];
  memcpy(out_g1, R.g1.data(), sizeof(int32_t) * (size_t)n);
  memcpy(out_g2, R.g2.data(), sizeof(int32_t) * (size_t)n);
  return 0;
}
