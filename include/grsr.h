/* grsr.h -- C ABI of libgrsr_cuda.so, the B200 (sm_100a) drop-in for GRIMM 2.01's unichromosomal
 * reversal-distance / sorting-by-reversals path.
 *
 * The reference has no library interface: GRSR calls the `grimm` executable
 * (reference scr/reptA.sh:61) whose main() (G/mcmain.c:195) calls the C functions below directly.
 * Each entry point here names the reference function it stands in for; a maintainer of GRIMM binds
 * them exactly where those functions are called today.
 *
 * Conventions
 *   - plain C, caller-owned buffers, no global state except the CUDA context and a per-thread
 *     error string; every function is thread-compatible.
 *   - a genome is `num_genes` signed 32-bit gene ids, a signed permutation of 1..num_genes, exactly
 *     the contents of `struct genome_struct.genes` (G/mcstructs.h:66-71) AFTER delete_caps
 *     (G/mcread_input.c:541) and, for -C, after circular_align (G/circ_align.c:50).  A genome table
 *     is row-major: genomes[g * num_genes + i].
 *   - return value 0 = success, negative = error code below; grsr_last_error() explains.
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails with
 *     GRSR_ERR_NODEVICE.
 */
#ifndef GRSR_H
#define GRSR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRSR_OK             0
#define GRSR_ERR_INVALID   (-1)  /* bad argument, or a genome that is not a signed permutation */
#define GRSR_ERR_CUDA      (-2)  /* CUDA runtime error */
#define GRSR_ERR_NODEVICE  (-3)  /* no CUDA device / requested device missing */
#define GRSR_ERR_TOOLARGE  (-4)  /* num_genes beyond grsr_max_genes() */
#define GRSR_ERR_NOMEM     (-5)
#define GRSR_ERR_SEARCH    (-6)  /* scenario search found no distance-reducing reversal
                                    (reference prints "ERROR: reversals ended prematurely",
                                    G/opt_scenario.c:260) */

/* graphstats_t fields used by the unichromosomal path (G/mcstructs.h:261-296, filled at
 * G/uniinvdist.c:202-208) */
typedef struct {
  int32_t d;  /* reversal distance  b - c + h + f */
  int32_t b;  /* br: breakpoints */
  int32_t c;  /* c4: cycles, adjacencies excluded */
  int32_t h;  /* hurdles */
  int32_t f;  /* fortress (0/1) */
} grsr_stats_t;

/* one reversal of a scenario: positions are 0-based, inclusive (rearrangement_t.s/.e,
 * G/opt_scenario.c:569-571) */
typedef struct {
  int32_t start;
  int32_t end;
} grsr_step_t;

int grsr_abi_version(void);
const char *grsr_last_error(void);
int grsr_device_count(void);           /* number of CUDA devices, 0 if none, never fails */
int grsr_max_genes(void);              /* largest num_genes accepted (global-memory build) */
int grsr_max_genes_on_chip(int scenario); /* largest num_genes served by the shared-memory build
                                          (scenario != 0: by the scenario kernels) */
void grsr_release(void);               /* frees the device buffers the host-buffer entry points keep
                                          between calls */

/* ---- host-buffer entry points (what GRIMM's own code would call) ------------------------------ */

/* Distance of every listed pair.  pairs[2p] = row used as gamma (g1), pairs[2p+1] = row used as pi
 * (g2) -- the argument order of invdist_noncircular(g1, g2, 0, num_genes, distmem)
 * (G/uniinvdist.c:224).  stride 1: out[p] = d.  stride 5: out + 5p is a grsr_stats_t, i.e.
 * invdist_noncircular_v (G/uniinvdist.c:235).  The pair list is split into `ngpus` contiguous
 * slices run on devices first_device .. first_device+ngpus-1; no inter-GPU communication. */
int grsr_dist_pairs(const int32_t *genomes, int num_genomes, int num_genes,
                    const int32_t *pairs, int64_t npairs,
                    int32_t *out, int stride, int first_device, int ngpus);

/* setinvmatrix(distmatrix, genomes, num_genes, num_genomes, distmem, FALSE)  (G/uniinvdist.c:75-92):
 * distmatrix is num_genomes x num_genomes row-major; diagonal 0, [i][j] = [j][i] = d(g_i, g_j), i<j.
 * The kernels write both triangle halves of the matrix themselves (on ngpus > 1 devices into the
 * first device's copy, through NVLink peer mappings); the host receives one finished matrix. */
int grsr_dist_matrix(const int32_t *genomes, int num_genomes, int num_genes,
                     int32_t *distmatrix, int first_device, int ngpus);

/* One shard of the same matrix for multi-process use (one process per GPU): the upper triangle is
 * enumerated column by column -- (0,1), (0,2), (1,2), (0,3), ... i.e. pair (i,j), i<j, has index
 * j(j-1)/2 + i -- as a flat list of P = G(G-1)/2 pairs; shard s of k owns the contiguous range
 * [P*s/k, P*(s+1)/k).  out receives that range's distances (its length is returned in *count). */
int grsr_dist_matrix_shard(const int32_t *genomes, int num_genomes, int num_genes,
                           int shard, int num_shards, int32_t *out, int64_t *count, int device);

/* The loop of do_fn_distmatrix_v (G/mcmain.c:1235-1242): all ordered pairs (i, j) including the
 * diagonal, statsmatrix[i * num_genomes + j] = stats of (gamma = g_i, pi = g_j). */
int grsr_stats_matrix(const int32_t *genomes, int num_genomes, int num_genes,
                      grsr_stats_t *statsmatrix, int first_device, int ngpus);

/* print_scenario_2(g1 = src, g2 = dst, num_genes, 0, scenario_type 4) minus the printing
 * (G/opt_scenario.c:73-301): for each of npairs independent (src, dst) pairs the optimal sequence
 * of reversals GRIMM would print.  src/dst: npairs x num_genes row-major.  steps: npairs x
 * max_steps entries, pair p's steps at steps + p*max_steps; nsteps[p] = number of steps
 * (= the reversal distance).  max_steps >= num_genes + 1 is always enough. */
int grsr_scenario_batch(const int32_t *src, const int32_t *dst, int npairs, int num_genes,
                        grsr_step_t *steps, int32_t *nsteps, int max_steps,
                        int first_device, int ngpus);

/* The first max_steps steps of the same scenarios (the whole scenario where it is shorter): what
 * print_scenario_2 has printed when it is stopped after max_steps "Step k:" lines.  A scenario is a
 * sequential chain of steps, each costing up to O(n^2) full distance evaluations when hurdles are
 * present (G/opt_scenario.c:951-986), so callers that sample, preview or time-box a batch bound the
 * work per pair this way.  complete[p] (may be NULL) = 1 if pair p reached its destination;
 * evals[p] (may be NULL) = number of full trial evaluations made for pair p.
 *
 * Both scenario calls shard over ngpus devices dynamically: all devices draw pairs from one queue
 * (a ticket in the first device's memory, NVLink peer atomics), because per-pair work differs by
 * orders of magnitude. */
int grsr_scenario_prefix(const int32_t *src, const int32_t *dst, int npairs, int num_genes,
                         grsr_step_t *steps, int32_t *nsteps, int max_steps, int32_t *complete,
                         int64_t *evals, int first_device, int ngpus);

/* Every listed candidate reversal of one pair scored at once -- the loop of print_rev_stats_t1
 * (G/testrev.c:156-215, `grimm -t 1`) and the evaluation of optrev_try_op (G/opt_scenario.c:974-986):
 * out[c] = invdist_noncircular(trial_c, dst) where trial_c = src with positions cands[2c] ..
 * cands[2c+1] (0-based, inclusive) reversed and negated (copy_and_reverse_genes, G/scenario.c:475).
 * num_genes <= grsr_max_genes_on_chip(0). */
int grsr_trial_distances(const int32_t *src, const int32_t *dst, int num_genes,
                         const int32_t *cands, int64_t ncand, int32_t *out, int device);

/* unsigned_mcdist_nomem(g1, g2, num_genes, 0 chromosomes, circular, verbose, &dist_error)
 * (G/unsigned.c:786-830, called at G/mcmain.c:810 for `grimm -u -g i,j` and, through
 * set_unsigneddist_matrix, G/unsigned.c:1107-1131, for the -u matrix): the exact reversal distance of
 * two UNSIGNED genomes -- the minimum of the signed distance over all signings -- by the reference's
 * algorithm: strips of >= 2 genes signed canonically, every sign assignment of the singletons
 * enumerated (at most 2^22, info->approx = 1 beyond that, as in the reference), 2-strips of
 * unoriented components re-signed per assignment (G/unsigned.c:885-1095).  The 2^k evaluations run on
 * the GPU.  The signs of g1 / g2 on input are ignored where the reference ignores them.
 * signed_g1 / signed_g2 (may be NULL) receive the genomes with the optimal signs (for circular
 * genomes also rotated / flipped the way the reference leaves them): the reference overwrites its
 * input with these, so a following scenario sorts signed_g1 into signed_g2.  For linear genomes
 * signed_g1 is g1 unchanged.  num_genes must fit the on-chip kernels. */
typedef struct {
  int32_t dist;         /* the unsigned reversal distance */
  int32_t approx;       /* 1: more than 22 singletons, answer approximated (reference: dist_error) */
  int32_t passes;       /* 1, or 2 for circular genomes without a common strip of 3 genes */
  int32_t num_sing[2];  /* per pass: what -v prints as "Number of Singletons" */
  int32_t num_doub[2];  /* per pass: "Number of 2-strips" */
} grsr_unsigned_info_t;
int grsr_unsigned_dist(const int32_t *g1, const int32_t *g2, int num_genes, int circular,
                       int32_t *signed_g1, int32_t *signed_g2, grsr_unsigned_info_t *info, int device);

/* ---- device-resident entry points (pointers are CUDA device pointers on the CURRENT device;
 *      stream is a cudaStream_t; no validation, no copies, asynchronous).  Up to
 *      grsr_max_genes_on_chip() genes a call takes its scratch from the device's stream-ordered
 *      memory pool, so calls on different streams are independent.  Beyond that the kernels use a
 *      per-device workspace the library owns: keep at most one such call in flight per device. --- */

/* signed inverse of every genome: d_sinv[g*n + |x|-1] = sign(x) * (position of x in genome g + 1) */
int grsr_signed_inverse_dev(const int32_t *d_genomes, int num_genomes, int num_genes,
                            int32_t *d_sinv, void *stream);

/* the batched distance kernel on resident data */
int grsr_dist_pairs_dev(const int32_t *d_genomes, const int32_t *d_sinv, int num_genes,
                        const int32_t *d_pairs, int64_t npairs, int32_t *d_out, int stride,
                        void *stream);

/* the same for a list the caller knows to be sorted by pi row (pairs[2p+1] non-decreasing in long
 * runs, e.g. the output of grsr_triangle_pairs_dev): skips the order test on the device.  On any
 * other list the results are still correct, only slower. */
int grsr_dist_pairs_sorted_dev(const int32_t *d_genomes, const int32_t *d_sinv, int num_genes,
                               const int32_t *d_pairs, int64_t npairs, int32_t *d_out, int stride,
                               void *stream);

/* fills d_pairs[2q], d_pairs[2q+1] with pair first+q of the column-by-column upper triangle */
int grsr_triangle_pairs_dev(int num_genomes, int64_t first, int64_t count, int32_t *d_pairs,
                            void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GRSR_H */
