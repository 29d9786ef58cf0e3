/* grimm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded restatement of the reference GRIMM 2.01 algorithm for the one path this
 * repository accelerates (unichromosomal signed reversal distance and the default -s scenario
 * search).  It exists to CHECK the CUDA path; the product (grsr_b200/, include/) never includes,
 * links or calls it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py compares every function below with the
 * unmodified reference compiled in place (oracle/_ref/libgrimmref.so, oracle/_ref/grimm) on
 * exhaustive small n, random, hurdle- and fortress-rich inputs, and tests/golden/ holds
 * reference-generated fixtures (example matrices and scenarios, -F tables, published docx cases)
 * that are checked on machines where the reference sources are absent.
 */
#ifndef GRIMM_ORACLE_H
#define GRIMM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int d;       /* reversal distance b - c + h + f            (G/uniinvdist.c:186) */
  int b;       /* breakpoints                                (G/uniinvdist.c:124-142) */
  int c;       /* cycles, adjacencies excluded               (G/graph_edit.c:432) */
  int h;       /* hurdles (greatest hurdle included)         (G/graph_components.c:603-624) */
  int f;       /* fortress 0/1                               (G/graph_components.c:707-709) */
  int comp_u;  /* unoriented components                      (G/uniinvdist.c:209) */
  int comp_o;  /* oriented components                        (G/uniinvdist.c:210) */
} orc_stats_t;

/* invdist_noncircular_v(g1=gamma, g2=pi, offset 0)  (G/uniinvdist.c:235-246).  Returns d. */
int orc_distance(const int *gamma, const int *pi, int n, orc_stats_t *st /* may be NULL */);

/* Same for a list of index pairs into a row-major genome table: pairs[2p]=gamma row, pairs[2p+1]=pi
 * row (setinvmatrix order, G/uniinvdist.c:80-90).  stride 1: out[p]=d; stride 5: {d,b,c,h,f}. */
void orc_dist_pairs(const int *genomes, int n, const int *pairs, long npairs, int *out, int stride);

/* Stage dump of the breakpoint graph for one pair; every array has 2n+2 entries.
 *   perm  (G/graph_edit.c:141-145)   grey (G/mcrdist.c:116-145, -1 = adjacency)
 *   cycle (G/graph_edit.c:296-437, leftmost vertex of the cycle, -1 for adjacency vertices)
 *   cc    (G/graph_components.c:154-181, component id in closing order, -1 for adjacency vertices)
 * Returns the number of components. */
int orc_graph(const int *gamma, const int *pi, int n, int *perm, int *grey, int *cycle, int *cc);

/* print_scenario_2(g1=src, g2=dst, scenario_type 4) minus the printing  (G/opt_scenario.c:73-301,
 * 472-585, 919-987): steps[2k]=start position, steps[2k+1]=end position of step k+1.
 * Returns the number of steps, -1 if it exceeds maxsteps, -2 if the search fails
 * ("ERROR: reversals ended prematurely", G/opt_scenario.c:260).
 * If evals != NULL it receives the number of full distance evaluations made per step. */
int orc_scenario(const int *src, const int *dst, int n, int *steps, int maxsteps, long *evals);

/* The first maxsteps steps of that scenario (the whole of it when it is shorter); returns their
 * number, -2 if the search fails. */
int orc_scenario_prefix(const int *src, const int *dst, int n, int *steps, int maxsteps, long *evals);

/* The inner loop of print_rev_stats_t1 (G/testrev.c:156-215): out[c] = invdist_noncircular(trial_c,
 * dst) with trial_c = src after copy_and_reverse_genes(cands[2c], cands[2c+1]). */
void orc_trial_distances(const int *src, const int *dst, int n, const int *cands, long ncand, int *out);

/* circular_align(genomes, G, n, 0, TRUE)  (G/circ_align.c:50-71): in place, row-major table. */
void orc_circular_align(int *genomes, int G, int n);

#ifdef __cplusplus
}
#endif
#endif
