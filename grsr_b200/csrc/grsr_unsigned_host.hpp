// grsr_unsigned_host.hpp -- host half of the exact UNSIGNED reversal distance (`grimm -u`, SURVEY.md
// section 8f row N2): the O(n) preprocessing of reference G/unsigned.c around the part that costs
// 2^k distance evaluations.
//
//   unsigned_mcdist          G/unsigned.c:649-780   driver: circular handling, 1 or 2 passes, best signs
//   find_3strip / test_strip G/unsigned.c:393-569   circular genomes: a common strip of 3 genes fixes the rotation
//   assign_canon_signs       G/unsigned.c:99-388    strips of >= 2 genes get their canonical signs; the
//                                                   singletons (sign unknown) and 2-strips are listed
//   loop_unsigned_perm       G/unsigned.c:836-869   every sign assignment of the singletons   } on the GPU:
//   sign_2strips             G/unsigned.c:885-1095  2-strips of unoriented components re-signed } grsr_unsigned.cuh
//
// Plain C++ (no CUDA): the library (grsr_capi.cu) and the CPU test build (tests/emu) both include it
// and supply `eval`, which scores all 2^k assignments of one pass.  No distance arithmetic happens here.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace grsr {

static const int kMaxUnsignedSingletons = 22;   // MAX_UNSIGNED_SING (G/mcstructs.h:55): 2^22 assignments at most

// circular_align_position (G/circ_align.c:178-188)
inline void ua_reverse_unsigned(int32_t *a, int lo, int hi) { for (; lo < hi; lo++, hi--) { int32_t t = a[lo]; a[lo] = a[hi]; a[hi] = t; } }
inline void ua_reverse_signed(int32_t *a, int lo, int hi) {
  for (; lo < hi; lo++, hi--) { int32_t t = a[lo]; a[lo] = -a[hi]; a[hi] = -t; }
  if (lo == hi) a[lo] = -a[lo];
}
inline void ua_align_position(int32_t *genes, int start, int n) {
  if (start < n) {
    ua_reverse_unsigned(genes, 0, start - 1);
    ua_reverse_unsigned(genes, start, n - 1);
    ua_reverse_unsigned(genes, 0, n - 1);
  } else {
    ua_reverse_signed(genes, 0, start - n - 1);
    ua_reverse_signed(genes, start - n, n - 1);
  }
}

// circular_align_1(g, n, align_on, align_front = TRUE) (G/circ_align.c:92-156)
inline void ua_align_on_gene(int32_t *genes, int n, int align_on) {
  int pos, sign = 1;
  for (pos = 0; pos < n; pos++) {
    if (genes[pos] == align_on) { sign = 1; break; }
    if (genes[pos] == -align_on) { sign = -1; break; }
  }
  if (pos == n) return;
  if (sign > 0) ua_align_position(genes, pos, n);
  else ua_align_position(genes, pos + 1 + n, n);
}

// test_strip (G/unsigned.c:393-431)
inline int ua_test_strip(const int32_t *g1, const int32_t *g2, int p1, int p2, int d1, int d2, int n, int max_len) {
  int len;
  for (len = 1; len < max_len; len++) {
    p1 += d1; if (p1 == -1) p1 = n - 1; else if (p1 == n) p1 = 0;
    p2 += d2; if (p2 == -1) p2 = n - 1; else if (p2 == n) p2 = 0;
    if (g1[p1] != g2[p2] && g1[p1] != -g2[p2]) break;
  }
  return len;
}

// find_3strip (G/unsigned.c:444-569)
inline void ua_find_3strip(const int32_t *g1, const int32_t *g2, int n, int *g1pos, int *g2pos) {
  *g1pos = *g2pos = -1;
  if (n < 4) return;
  std::vector<int> ip((size_t)n + 1);
  for (int i = 0; i < n; i++) ip[(size_t)abs(g1[i])] = i;
  for (int i = 0; i < n; i++) {
    const int cur = abs(g2[i]);
    if (ua_test_strip(g1, g2, ip[(size_t)cur], i, 1, 1, n, 3) == 3) { *g1pos = ip[(size_t)cur]; *g2pos = i; break; }
    if (ua_test_strip(g1, g2, ip[(size_t)cur], i, -1, 1, n, 3) == 3) { *g1pos = ip[(size_t)cur] + n; *g2pos = i; break; }
  }
  if (*g2pos == 0) {
    if (*g1pos < n) {
      int len = ua_test_strip(g1, g2, *g1pos, *g2pos, -1, -1, n, n - 2);
      if (len == n - 2) len = 1;
      len--;
      *g1pos -= len; if (*g1pos < 0) *g1pos += n;
      *g2pos -= len; if (*g2pos < 0) *g2pos += n;
    } else {
      // The reference extends the strip with test_strip(genes1, genes2, *g1pos, ...) where *g1pos is
      // still offset by n (:541-545): its walk reads genes1[n + ...], past the end of the array, so
      // the comparison fails on whatever lies there and the strip is not extended.  That outcome
      // (extension length 0) is what is reproduced here, without the out-of-bounds read.
      int len = 1;
      len--;
      *g1pos += len; if (*g1pos >= 2 * n) *g1pos -= n;
      *g2pos -= len; if (*g2pos < 0) *g2pos += n;
    }
  }
  if (*g1pos >= n) { if (++*g1pos == 2 * n) *g1pos = n; }
}

struct UnsignedPass {
  std::vector<int32_t> pi0;        // g2 with canonical signs on every strip of >= 2 genes (loop_g2)
  std::vector<int32_t> singletons; // positions in g2, increasing
  std::vector<int32_t> doubletons; // positions r of the 2-strips (r, r+1), increasing
};

// assign_canon_signs for unichromosomal genomes (G/unsigned.c:99-388 with num_chromosomes == 0, so
// lowcap = n + 1: no gene is a cap)
inline void ua_assign_canon_signs(const int32_t *g1, const int32_t *g2, int n, UnsignedPass &P) {
  std::vector<int> ip((size_t)n + 1);
  for (int i = 0; i < n; i++) ip[(size_t)abs(g1[i])] = i;
  P.pi0.assign(g2, g2 + n);
  P.singletons.clear(); P.doubletons.clear();
  int32_t *genes2 = P.pi0.data();
  int i, i1 = 0, i2 = n;
  for (i = i1; i < i2; i++) {                       // identical strip at the start (:186-200)
    if (ip[(size_t)abs(genes2[i])] == i) genes2[i] = g1[i]; else break;
  }
  i1 = i;
  for (i = i2 - 1; i > i1; i--) {                   // ... and at the end (:203-217)
    if (ip[(size_t)abs(genes2[i])] == i) genes2[i] = g1[i]; else break;
  }
  i2 = i + 1;
  int strip_len = 0, strip_dir = 0, cur_dir = 0, last = 0;
  for (i = i1; i < i2; i++) {
    const int cur_gene = abs(genes2[i]);
    const int pos1 = ip[(size_t)cur_gene];
    bool done = true;
    cur_dir = pos1 - last;
    if (strip_len == 1 && (cur_dir == 1 || cur_dir == -1)) {
      strip_dir = cur_dir;
      genes2[i - 1] = strip_dir * g1[last];         // canonical sign of the strip's first gene
    }
    if (strip_dir == cur_dir && strip_len >= 1) {
      genes2[i] = strip_dir * g1[pos1];
      done = false;
    }
    if (done) {
      if (strip_len == 1) P.singletons.push_back(i - 1);
      else if (strip_len == 2) P.doubletons.push_back(i - 2);
      strip_len = 1;
      // the direction of the strip that starts here is not known yet; the reference leaves the old
      // strip_dir in place (:296-303 only overwrites it when strip_len == 1 and |cur_dir| == 1)
    } else {
      strip_len++;
    }
    last = pos1;
  }
  if (strip_len == 1) P.singletons.push_back(i - 1);
  else if (strip_len == 2) P.doubletons.push_back(i - 2);
}

struct UnsignedResult {
  int dist;
  int approx;                      // dist_error: more than kMaxUnsignedSingletons singletons, the rest kept as read
  int passes;                      // 1, or 2 for circular genomes without a common 3-strip
  int num_sing[2], num_doub[2];    // what -v prints per pass ("Number of Singletons", "Number of 2-strips")
  std::vector<int32_t> g1, g2;     // the genomes with the optimal signs (and circular shifts): cappedp1 / cappedp2
};

// unsigned_mcdist for num_chromosomes == 0 (G/unsigned.c:649-780).
// eval(g1, pass, k, best_d, best_pi): scores the 2^k sign assignments of the first k singletons of the
// pass in the reference's enumeration order and, if the smallest distance found is < best_d, updates
// best_d and best_pi (the fully signed g2 of the FIRST assignment that reaches it); returns 0 or an
// error code.
template <class Eval>
inline int unsigned_distance(const int32_t *g1_in, const int32_t *g2_in, int n, bool circular, Eval eval,
                             UnsignedResult &R) {
  std::vector<int32_t> g1(g1_in, g1_in + n), g2(g2_in, g2_in + n);
  int tries = 1;
  if (circular) {
    int p1, p2;
    ua_find_3strip(g1.data(), g2.data(), n, &p1, &p2);
    if (p1 >= 0) { ua_align_position(g1.data(), p1, n); ua_align_position(g2.data(), p2, n); }
    else { ua_align_on_gene(g2.data(), n, g1[0]); tries = 2; }
  }
  int best_d = 2 * n + 1;
  std::vector<int32_t> best_pi(g2);
  R.passes = 0; R.approx = 0;
  R.num_sing[0] = R.num_sing[1] = R.num_doub[0] = R.num_doub[1] = 0;
  for (; tries > 0; tries--) {
    UnsignedPass P;
    ua_assign_canon_signs(g1.data(), g2.data(), n, P);
    R.num_sing[R.passes] = (int)P.singletons.size();
    R.num_doub[R.passes] = (int)P.doubletons.size();
    R.passes++;
    int k = (int)P.singletons.size();
    R.approx = 0;                                    // the reference resets the flag in every pass (:729)
    if (k > kMaxUnsignedSingletons) { R.approx = 1; k = kMaxUnsignedSingletons; }
    int rc = eval(g1.data(), P, k, best_d, best_pi);
    if (rc) return rc;
    if (tries == 2) ua_align_position(g2.data(), n + 1, n);   // flip g2 around its first gene (:752)
  }
  R.dist = best_d;
  R.g1 = g1;
  R.g2 = best_pi;
  return 0;
}

} // namespace grsr
