/* grimm_oracle.c -- TEST INFRASTRUCTURE ONLY (see grimm_oracle.h).
 *
 * Sequential restatement of the reference algorithm.  Each routine cites the reference lines whose
 * behaviour it reproduces; the code itself is written from that behaviour, with its own data
 * layout (one workspace struct, no aliasing of scratch arrays).  Vertex numbering follows the
 * reference: size = 2n+2 vertices, black edges join (2k, 2k+1), k = 0..n.
 */
#include <stdlib.h>
#include <string.h>
#include "grimm_oracle.h"

#define NONE (-1)

typedef struct {
  int n, size;
  int *perm;      /* position -> value of the doubled relative permutation                 */
  int *where;     /* value -> position                                                      */
  int *grey;      /* grey-edge partner, NONE for an adjacency                               */
  int *cycle;     /* leftmost vertex of the vertex's cycle, NONE for adjacency vertices     */
  int *rightmost; /* indexed by leftmost vertex: largest vertex of that cycle               */
  int *link;      /* union forest over cycle roots built by the sweep                       */
  int *cc;        /* component id per vertex, NONE for adjacency vertices                   */
  int *st_root, *st_reach;      /* the sweep's stack                                        */
  int *chain;                   /* linked lists used to spread component ids               */
  /* per component */
  int ncomp;
  int *c_root, *c_oriented, *c_blocks, *c_left, *c_right, *c_flags;
  int ncycles, nadj;
} ws_t;

static void *xcalloc(size_t k, size_t sz) {
  void *p = calloc(k ? k : 1, sz);
  if (!p) abort();
  return p;
}

static void ws_init(ws_t *w, int n) {
  int size = 2 * n + 2, m = n + 2;
  w->n = n; w->size = size;
  w->perm = xcalloc(size, sizeof(int));      w->where = xcalloc(size, sizeof(int));
  w->grey = xcalloc(size, sizeof(int));      w->cycle = xcalloc(size, sizeof(int));
  w->rightmost = xcalloc(size, sizeof(int)); w->link = xcalloc(size, sizeof(int));
  w->cc = xcalloc(size, sizeof(int));        w->st_root = xcalloc(size, sizeof(int));
  w->st_reach = xcalloc(size, sizeof(int));  w->chain = xcalloc(size, sizeof(int));
  w->c_root = xcalloc(m, sizeof(int));       w->c_oriented = xcalloc(m, sizeof(int));
  w->c_blocks = xcalloc(m, sizeof(int));     w->c_left = xcalloc(m, sizeof(int));
  w->c_right = xcalloc(m, sizeof(int));      w->c_flags = xcalloc(m, sizeof(int));
}

static void ws_free(ws_t *w) {
  free(w->perm); free(w->where); free(w->grey); free(w->cycle); free(w->rightmost);
  free(w->link); free(w->cc); free(w->st_root); free(w->st_reach); free(w->chain);
  free(w->c_root); free(w->c_oriented); free(w->c_blocks); free(w->c_left);
  free(w->c_right); free(w->c_flags);
}

/* a3: double_perms (G/graph_edit.c:46-157), offset = 0.
 * slot[] plays the part of perm1 (doubled inverse of gamma); perm = perm1 o perm2. */
static void build_perm(ws_t *w, const int *gamma, const int *pi) {
  int n = w->n, size = w->size, i;
  int *slot = w->where; /* borrowed: overwritten by build_grey afterwards */
  for (i = 0; i < n; i++) {
    int g = gamma[i], lo = 2 * i + 1;
    if (g > 0) { slot[2 * g - 1] = lo;     slot[2 * g] = lo + 1; }
    else       { slot[-2 * g] = lo;        slot[-2 * g - 1] = lo + 1; }
  }
  for (i = 0; i < n; i++) {
    int g = pi[i], first, second;
    if (g > 0) { first = 2 * g - 1; second = 2 * g; }
    else       { first = -2 * g;    second = -2 * g - 1; }
    w->perm[2 * i + 1] = slot[first];
    w->perm[2 * i + 2] = slot[second];
  }
  w->perm[0] = 0;
  w->perm[size - 1] = size - 1;
}

/* a4: buildG_grey, num_chromosomes == 0 (G/mcrdist.c:116-145). */
static void build_grey(ws_t *w) {
  int size = w->size, v;
  const int *perm = w->perm;
  int *where = w->where, *grey = w->grey;
  for (v = 0; v < size; v++) { where[perm[v]] = v; grey[v] = NONE; }
  if (where[1] != 1) grey[0] = where[1];
  for (v = 1; v < size - 1; v += 2) {
    int val = perm[v], pa, pb;
    if (val < perm[v + 1]) { pa = where[val - 1]; pb = where[val + 2]; }
    else                   { pa = where[val + 1]; pb = where[val - 2]; }
    if (pa != v - 1) grey[v] = pa;
    if (pb != v + 2) grey[v + 1] = pb;
  }
  if (where[size - 2] != size - 2) grey[size - 1] = where[size - 2];
}

/* a5: classify_cycle_path (G/graph_edit.c:207-457) for a graph that has cycles only. */
static void label_cycles(ws_t *w) {
  int size = w->size, v;
  int *cycle = w->cycle, *grey = w->grey;
  w->ncycles = 0; w->nadj = 0;
  for (v = 0; v < size; v++) {
    cycle[v] = NONE; w->rightmost[v] = NONE;
    if (grey[v] == NONE) w->nadj++;
  }
  w->nadj /= 2;
  for (v = 0; v < size; v++) {
    int u, far;
    if (grey[v] == NONE || cycle[v] != NONE) continue;
    u = v; far = v; cycle[v] = v;
    do {
      u ^= 1;                    /* black edge (G/graph_edit.c:346) */
      cycle[u] = v; if (u > far) far = u;
      u = grey[u];               /* grey edge */
      cycle[u] = v; if (u > far) far = u;
    } while (u != v);
    w->rightmost[v] = far;
    w->ncycles++;
  }
}

/* a6: form_connected_components (G/graph_components.c:50-242).
 * Left-to-right sweep with a stack of open components; component ids are handed out in the order
 * the components close (G/graph_components.c:131-136). */
static void sweep_components(ws_t *w) {
  int size = w->size, v, top = -1, k;
  int *link = w->link, *cycle = w->cycle;
  w->ncomp = 0;
  for (v = 0; v < size; v++) link[v] = cycle[v];
  for (v = 0; v < size; v++) {
    int reach;
    if (w->grey[v] == NONE) continue;
    if (link[v] == v) {                       /* leftmost vertex of a cycle: open it */
      top++;
      w->st_root[top] = v;
      w->st_reach[top] = w->rightmost[v];
      continue;
    }
    reach = v;
    while (w->st_root[top] > link[v]) {       /* everything opened after my cycle joins it */
      link[w->st_root[top]] = link[v];
      if (w->st_reach[top] > reach) reach = w->st_reach[top];
      top--;
    }
    if (w->st_reach[top] < reach) w->st_reach[top] = reach;
    if (w->st_reach[top] <= v) {              /* nothing of it lies to the right: close it */
      w->c_root[w->ncomp++] = w->st_root[top];
      top--;
    }
  }
  /* spread ids: forest -> chains hanging off the component root (G/graph_components.c:154-181) */
  for (v = 0; v < size; v++) w->chain[v] = NONE;
  for (v = 0; v < size; v++) {
    if (w->grey[v] == NONE) { w->cc[v] = NONE; continue; }
    if (link[v] != v) { w->chain[v] = w->chain[link[v]]; w->chain[link[v]] = v; }
  }
  for (k = 0; k < w->ncomp; k++)
    for (v = w->c_root[k]; v != NONE; v = w->chain[v]) w->cc[v] = k;
}

/* a7: classify_connected_components (G/graph_components.c:257-336): a component is oriented iff
 * one of its grey edges joins two vertices an even distance apart. */
static void orient_components(ws_t *w) {
  int v, k;
  for (k = 0; k < w->ncomp; k++) w->c_oriented[k] = 0;
  for (v = 0; v < w->size; v++) {
    int u = w->grey[v];
    if (v < u && ((u - v) % 2 == 0)) w->c_oriented[w->cc[v]] = 1;
  }
}

#define F_HURDLE 1
#define F_GREAT  2
#define F_SUPER  4

/* a8: calc_num_hurdles_and_fortress(G, C_U) (G/graph_components.c:465-727). */
static void count_hurdles(ws_t *w, int *h_out, int *f_out, int *unor_out) {
  int k, v, unor = 0, first = NONE, last = NONE;
  int hurdles = 0, supers = 0;
  *h_out = 0; *f_out = 0; *unor_out = 0;
  if (w->ncomp == 0) return;
  for (k = 0; k < w->ncomp; k++) if (!w->c_oriented[k]) unor++;
  *unor_out = unor;
  if (unor == 0) return;
  for (k = 0; k < w->ncomp; k++) {
    w->c_blocks[k] = 0; w->c_flags[k] = 0; w->c_left[k] = NONE; w->c_right[k] = NONE;
  }
  /* runs of one unoriented component along the vertex line (:566-585); left/right are rewritten
   * at every change of run, so they end up describing the neighbourhood of the LAST change */
  for (v = 0; v < w->size; v++) {
    k = w->cc[v];
    if (k == NONE || w->c_oriented[k] || k == last) continue;
    if (last == NONE) first = k;
    else { w->c_right[last] = k; w->c_left[k] = last; }
    last = k;
    w->c_blocks[k]++;
  }
  for (k = 0; k < w->ncomp; k++)
    if (w->c_blocks[k] == 1) { w->c_flags[k] |= F_HURDLE; hurdles++; }          /* :603-615 */
  if (first == last && w->c_blocks[first] == 2) {                                /* :617-624 */
    w->c_flags[first] |= F_HURDLE | F_GREAT; hurdles++;
  }
  if (first != NONE && first < last) {                                           /* :640-656 */
    if (w->c_right[first] == last && (w->c_flags[first] & F_HURDLE) &&
        !(w->c_flags[last] & F_HURDLE)) {
      w->c_flags[first] |= F_SUPER; supers++;
    } else if (w->c_left[last] == first && (w->c_flags[last] & F_HURDLE) &&
               !(w->c_flags[first] & F_HURDLE)) {
      w->c_flags[last] |= F_SUPER; supers++;
    }
  }
  for (k = 0; k < w->ncomp; k++) {                                               /* :674-694 */
    int nb;
    if (!(w->c_flags[k] & (F_HURDLE | F_GREAT | F_SUPER))) continue;
    nb = w->c_left[k];
    if (nb == NONE || nb != w->c_right[k]) continue;
    if (w->c_blocks[nb] == 2 && !(w->c_flags[nb] & F_GREAT)) {
      w->c_flags[k] |= F_SUPER; supers++;
    }
  }
  *h_out = hurdles;
  if (hurdles == supers && (supers % 2) == 1) *f_out = 1;                        /* :707-709 */
}

/* a9: num_breakpoints (G/uniinvdist.c:124-142). */
static int count_breakpoints(const ws_t *w) {
  const int *perm = w->perm;
  int size = w->size, i, b = (perm[1] != 1);
  for (i = 2; i < size - 1; i += 2) {
    int x = perm[i], y = perm[i + 1];
    int x1 = (x == size ? 1 : x + 1), y1 = (y == size ? 1 : y + 1);
    if (y != x1 && x != y1) b++;
  }
  return b;
}

/* a10: invdist_noncircular_G (G/uniinvdist.c:150-214); the workspace keeps the graph. */
static int distance_ws(ws_t *w, const int *gamma, const int *pi, orc_stats_t *st) {
  int h, f, unor, b, d;
  build_perm(w, gamma, pi);
  build_grey(w);
  label_cycles(w);
  sweep_components(w);
  orient_components(w);
  count_hurdles(w, &h, &f, &unor);
  b = count_breakpoints(w);
  d = b - w->ncycles + h + f;
  if (st) {
    st->d = d; st->b = b; st->c = w->ncycles; st->h = h; st->f = f;
    st->comp_u = unor; st->comp_o = w->ncomp - unor;
  }
  return d;
}

int orc_distance(const int *gamma, const int *pi, int n, orc_stats_t *st) {
  ws_t w; int d;
  ws_init(&w, n);
  d = distance_ws(&w, gamma, pi, st);
  ws_free(&w);
  return d;
}

void orc_dist_pairs(const int *genomes, int n, const int *pairs, long npairs, int *out, int stride) {
  ws_t w; long p; orc_stats_t st;
  ws_init(&w, n);
  for (p = 0; p < npairs; p++) {
    distance_ws(&w, genomes + (long)pairs[2 * p] * n, genomes + (long)pairs[2 * p + 1] * n, &st);
    if (stride == 1) out[p] = st.d;
    else { int *o = out + 5 * p; o[0] = st.d; o[1] = st.b; o[2] = st.c; o[3] = st.h; o[4] = st.f; }
  }
  ws_free(&w);
}

int orc_graph(const int *gamma, const int *pi, int n, int *perm, int *grey, int *cycle, int *cc) {
  ws_t w; int nc; size_t bytes = sizeof(int) * (size_t)(2 * n + 2);
  ws_init(&w, n);
  distance_ws(&w, gamma, pi, NULL);
  memcpy(perm, w.perm, bytes); memcpy(grey, w.grey, bytes);
  memcpy(cycle, w.cycle, bytes); memcpy(cc, w.cc, bytes);
  nc = w.ncomp;
  ws_free(&w);
  return nc;
}

/* a17: graphdist (G/graph_edit.c:1165-1238) on a graph of cycles: number of edges walked from x,
 * black edge first, until y is met; -1 if they are in different cycles. */
static int walk_length(const ws_t *w, int x, int y) {
  int v = x, steps = 0;
  if (w->cycle[x] != w->cycle[y]) return -1;
  if (x == y) return 0;
  if (w->cycle[x] == NONE) return ((x ^ 1) == y) ? 1 : -1;
  for (;;) {
    v ^= 1; steps++;
    if (v == y) return steps;
    v = w->grey[v]; steps++;
    if (v == y) return steps;
    if (v == x) abort();
  }
}

/* a13: successor table of the destination (G/scenario.c:95-171, macro G/scenario.h:121).
 * succ is indexed by gene + n + 1 so that negative genes have their own slots. */
static void build_successors(const int *dst, int n, int *succ) {
  int i;
  for (i = 0; i < n - 1; i++) {
    int g = dst[i], h = dst[i + 1];
    succ[g + n + 1] = h;
    succ[-h + n + 1] = -g;
  }
  succ[dst[n - 1] + n + 1] = n + 1;
  succ[-dst[0] + n + 1] = -(n + 1);
}

/* a14-a16: print_scenario_2 / try_optrev / optrev_try_op (G/opt_scenario.c:227-290, 527-576,
 * 951-986). */
static int scenario_run(const int *src, const int *dst, int n, int *steps, int maxsteps, long *evals,
                        int prefix);

int orc_scenario(const int *src, const int *dst, int n, int *steps, int maxsteps, long *evals) {
  return scenario_run(src, dst, n, steps, maxsteps, evals, 0);
}

/* the first maxsteps steps of the same scenario (all of it when it is shorter): returns their number */
int orc_scenario_prefix(const int *src, const int *dst, int n, int *steps, int maxsteps, long *evals) {
  return scenario_run(src, dst, n, steps, maxsteps, evals, 1);
}

static int scenario_run(const int *src, const int *dst, int n, int *steps, int maxsteps, long *evals,
                        int prefix) {
  ws_t g0, wt;
  int *cur = xcalloc(n, sizeof(int)), *trial = xcalloc(n, sizeof(int));
  int *succ = xcalloc(2 * n + 3, sizeof(int));
  int nsteps = 0, status = 0;
  ws_init(&g0, n); ws_init(&wt, n);
  memcpy(cur, src, sizeof(int) * (size_t)n);
  build_successors(dst, n, succ);
  for (;;) {
    orc_stats_t st0;
    int d = distance_ws(&g0, dst, cur, &st0);      /* gamma = destination, pi = current (:229) */
    int len, i, j, k, found = 0;
    long nev = 0;
    if (d == 0) break;
    if (prefix && nsteps >= maxsteps) break;
    for (len = 0; len < n && !found; len++) {
      for (i = 0; i + len < n; i++) {
        int *t;
        j = i + len;
        if (i == 0) { if (cur[0] == dst[0]) continue; }                        /* :538-548 */
        else if (succ[cur[i - 1] + n + 1] == cur[i]) continue;
        if (j == n - 1) { if (cur[j] == dst[j]) continue; }                    /* :551-561 */
        else if (succ[cur[j] + n + 1] == cur[j + 1]) continue;
        if (st0.h == 0 && walk_length(&g0, 2 * i + 1, 2 * j + 2) % 2 != 0)     /* :951-971 */
          continue;
        for (k = 0; k < n; k++) trial[k] = cur[k];                             /* G/scenario.c:475 */
        for (k = i; k <= j; k++) trial[k] = -cur[i + j - k];
        nev++;
        if (distance_ws(&wt, trial, dst, NULL) == d - 1) {  /* gamma = trial, pi = dest (:980) */
          found = 1;
          if (nsteps >= maxsteps) { status = -1; goto done; }
          steps[2 * nsteps] = i; steps[2 * nsteps + 1] = j;
          if (evals) evals[nsteps] = nev;
          nsteps++;
          t = cur; cur = trial; trial = t;
          break;
        }
      }
    }
    if (!found) { status = -2; break; }
  }
done:
  ws_free(&g0); ws_free(&wt); free(cur); free(trial); free(succ);
  return status < 0 ? status : nsteps;
}

void orc_trial_distances(const int *src, const int *dst, int n, const int *cands, long ncand, int *out) {
  ws_t w; long c; int k;
  int *trial = xcalloc(n, sizeof(int));
  ws_init(&w, n);
  for (c = 0; c < ncand; c++) {
    int i = cands[2 * c], j = cands[2 * c + 1];
    for (k = 0; k < n; k++) trial[k] = src[k];
    for (k = i; k <= j; k++) trial[k] = -src[i + j - k];        /* G/scenario.c:475-486 */
    out[c] = distance_ws(&w, trial, dst, NULL);                 /* gamma = trial, pi = dest */
  }
  ws_free(&w); free(trial);
}

static void flip_range(int *a, int lo, int hi, int negate) {
  for (; lo < hi; lo++, hi--) {
    int t = a[lo];
    a[lo] = negate ? -a[hi] : a[hi];
    a[hi] = negate ? -t : t;
  }
  if (lo == hi && negate) a[lo] = -a[lo];
}

/* a12: circular_align(..., align_on = 0, align_front = TRUE) (G/circ_align.c:50-188). */
void orc_circular_align(int *genomes, int G, int n) {
  int a = genomes[0], gi;
  for (gi = 1; gi < G; gi++) {
    int *g = genomes + (long)gi * n, p, sign = 1;
    for (p = 0; p < n; p++) {
      if (g[p] == a) { sign = 1; break; }
      if (g[p] == -a) { sign = -1; break; }
    }
    if (sign > 0) {
      if (p == 0) continue;
      flip_range(g, 0, p - 1, 0); flip_range(g, p, n - 1, 0); flip_range(g, 0, n - 1, 0);
    } else {
      flip_range(g, 0, p, 1); flip_range(g, p + 1, n - 1, 1);
    }
  }
}
