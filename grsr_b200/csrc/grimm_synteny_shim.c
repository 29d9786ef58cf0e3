/* grimm_synteny_shim.c -- link-time drop-in for GRIMM-Synteny.
 *
 * GRIMM-Synteny links byte-identical copies of GRIMM's distance files and calls
 *     mcdist_noncircular(&genome_list[i], &genome_list[j], n, num_chromosomes, distmem, &stats)
 * for every species pair of every synteny block with num_chromosomes == 0
 * (reference GRIMM-synteny/GRIMM_SYNTENY-2.02/gscomp.c:1479-1484) and for the macro-rearrangement
 * matrix with num_chromosomes = max_chr (gscomp.c:2839-2844); scratch comes from gs_grimm_alloc
 * (gs_grimm.c:36-84).  With num_chromosomes == 0 the reference function is nothing but
 * invdist_noncircular_v (mcrdist.c:1873-1876) -- the path libgrsr_cuda.so replaces.
 *
 * This file is compiled INSIDE the GRIMM-Synteny tree (it includes that tree's own mcstructs.h and
 * mcrdist.h) and linked with
 *     -Wl,--wrap=mcdist_noncircular  grimm_synteny_shim.o  -lgrsr_cuda
 * so that every call of mcdist_noncircular from the other objects lands in
 * __wrap_mcdist_noncircular below, with no source change in GRIMM-Synteny:
 *   - num_chromosomes == 0 : the pair goes to the GPU (grsr_dist_pairs, stride 5) and the
 *     unichromosomal fields of graphstats_t are filled as uniinvdist.c:202-208 fills them;
 *   - num_chromosomes  > 0 : multichromosomal genomes are outside the path; the call is handed to
 *     the reference's own function (__real_mcdist_noncircular), untouched.
 * Errors follow the reference's convention: message on stderr, exit(-1) (e_malloc.c:43-58).
 *
 * gs_grsr_micro_matrix() is the batched form of the gs_micro_1block loop (one GPU call per block
 * instead of one per species pair); INTEGRATION.md section 2.6 shows the patch that uses it.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mcstructs.h"
#include "mcrdist.h"
#include "grsr.h"

int __real_mcdist_noncircular(struct genome_struct *g1, struct genome_struct *g2, int num_genes,
                              int num_chromosomes, distmem_t *distmem, graphstats_t *graphstats);

static int shim_device(void)
{
  const char *s = getenv("GRSR_DEVICE");
  return (s && *s) ? atoi(s) : 0;
}

static void shim_fail(const char *what)
{
  fprintf(stderr, "ERROR: %s: %s\n", what, grsr_last_error());
  exit(-1);
}

/* uniinvdist.c:202-208; the two component counts (:209-210) are not produced by the GPU path and no
 * caller in GRIMM-Synteny reads them: they are set to -1 */
static void fill_stats(graphstats_t *gs, const grsr_stats_t *st)
{
  gs->br = st->b; gs->c4 = st->c; gs->h = st->h; gs->f = st->f; gs->d = st->d;
  gs->num_components_u = -1; gs->num_components_o = -1;
}

int __wrap_mcdist_noncircular(struct genome_struct *g1, struct genome_struct *g2, int num_genes,
                              int num_chromosomes, distmem_t *distmem, graphstats_t *graphstats)
{
  int32_t *table, pair[2] = {0, 1};
  grsr_stats_t st;
  if (num_chromosomes != 0)
    return __real_mcdist_noncircular(g1, g2, num_genes, num_chromosomes, distmem, graphstats);
  (void)distmem;                       /* the reference's scratch is not used on the GPU path */
  if (num_genes <= 0) {                /* a block without anchors: b = c = 0 (the reference returns 0 too) */
    if (graphstats) { memset(&st, 0, sizeof st); fill_stats(graphstats, &st); }
    return 0;
  }
  table = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)num_genes);
  if (!table) { fprintf(stderr, "ERROR: Could not allocate the genome table\n"); exit(-1); }
  memcpy(table, g1->genes, sizeof(int32_t) * (size_t)num_genes);                 /* gamma */
  memcpy(table + num_genes, g2->genes, sizeof(int32_t) * (size_t)num_genes);     /* pi */
  if (grsr_dist_pairs(table, 2, num_genes, pair, 1, (int32_t *)&st, 5, shim_device(), 1) != GRSR_OK)
    shim_fail("mcdist_noncircular (GPU)");
  free(table);
  if (graphstats) fill_stats(graphstats, &st);
  return st.d;
}

/* The double loop of gs_micro_1block (gscomp.c:1476-1489) in one call: statsmatrix[i][j], j < i, for
 * the nspecies genomes of genome_list (num_genes anchors each), including the "# reuses" kludge the
 * caller stores in field n (gscomp.c:1486-1488). */
void gs_grsr_micro_matrix(struct genome_struct *genome_list, int nspecies, int num_genes,
                          graphstats_t **statsmatrix)
{
  int i, j, npairs = nspecies * (nspecies - 1) / 2, p = 0;
  int32_t *table, *pairs;
  grsr_stats_t *out;
  if (npairs <= 0) return;
  if (num_genes <= 0) {
    grsr_stats_t z; memset(&z, 0, sizeof z);
    for (i = 1; i < nspecies; i++) for (j = 0; j < i; j++) { fill_stats(&statsmatrix[i][j], &z); statsmatrix[i][j].n = 0; }
    return;
  }
  table = (int32_t *)malloc(sizeof(int32_t) * (size_t)nspecies * num_genes);
  pairs = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)npairs);
  out = (grsr_stats_t *)malloc(sizeof(grsr_stats_t) * (size_t)npairs);
  if (!table || !pairs || !out) { fprintf(stderr, "ERROR: Could not allocate the genome table\n"); exit(-1); }
  for (i = 0; i < nspecies; i++)
    memcpy(table + (size_t)i * num_genes, genome_list[i].genes, sizeof(int32_t) * (size_t)num_genes);
  for (i = 1; i < nspecies; i++) for (j = 0; j < i; j++) { pairs[2 * p] = i; pairs[2 * p + 1] = j; p++; }
  if (grsr_dist_pairs(table, nspecies, num_genes, pairs, npairs, (int32_t *)out, 5, shim_device(), 1) != GRSR_OK)
    shim_fail("gs_grsr_micro_matrix (GPU)");
  p = 0;
  for (i = 1; i < nspecies; i++)
    for (j = 0; j < i; j++, p++) {
      fill_stats(&statsmatrix[i][j], &out[p]);
      statsmatrix[i][j].n = 2 * statsmatrix[i][j].d - statsmatrix[i][j].br;
    }
  free(out); free(pairs); free(table);
}
