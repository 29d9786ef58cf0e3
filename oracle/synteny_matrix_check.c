/* synteny_matrix_check.c -- TEST INFRASTRUCTURE ONLY.  Checks the batched helper of the GRIMM-Synteny
 * shim (gs_grsr_micro_matrix, grsr_b200/csrc/grimm_synteny_shim.c) against the reference's own
 * mcdist_noncircular (reached as __real_mcdist_noncircular under -Wl,--wrap) and against the
 * one-pair wrap, on seeded genomes.  Built by oracle/Makefile into oracle/_ref/synteny_matrix_check
 * with the reference's objects; run by tests/test_gpu_synteny.py.  Prints "OK <pairs>" or a mismatch. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mcstructs.h"
#include "mcrdist.h"

FILE *outfile;
int __real_mcdist_noncircular(struct genome_struct *, struct genome_struct *, int, int, distmem_t *, graphstats_t *);
void gs_grsr_micro_matrix(struct genome_struct *genome_list, int nspecies, int num_genes, graphstats_t **statsmatrix);

static unsigned rnd_state = 12345u;
static unsigned rnd(void) { rnd_state = rnd_state * 1664525u + 1013904223u; return rnd_state >> 8; }

int main(void)
{
  int sizes[4] = {1, 7, 40, 333}, si, checked = 0;
  outfile = stdout;
  for (si = 0; si < 4; si++) {
    const int n = sizes[si], ns = 6;
    struct genome_struct gl[6];
    graphstats_t *rows[6], one, ref;
    distmem_t dm;
    int i, j, k;
    mcdist_allocmem(n, 0, &dm);
    for (i = 0; i < ns; i++) {
      memset(&gl[i], 0, sizeof gl[i]);
      gl[i].genes = (int *)malloc(sizeof(int) * (size_t)n);
      rows[i] = (graphstats_t *)calloc((size_t)ns, sizeof(graphstats_t));
      for (k = 0; k < n; k++) gl[i].genes[k] = k + 1;
      if (i > 0)                                   /* shuffle, random signs; genome 1 in sorted blocks of 3 (hurdles) */
        for (k = n - 1; k > 0; k--) {
          int r = (int)(rnd() % (unsigned)(k + 1)), t = gl[i].genes[k];
          gl[i].genes[k] = gl[i].genes[r]; gl[i].genes[r] = t;
        }
      if (i > 1) for (k = 0; k < n; k++) if (rnd() & 1) gl[i].genes[k] = -gl[i].genes[k];
    }
    if (n >= 6) { int t = gl[0].genes[1]; gl[0].genes[1] = gl[0].genes[2]; gl[0].genes[2] = t; }
    gs_grsr_micro_matrix(gl, ns, n, rows);
    for (i = 1; i < ns; i++)
      for (j = 0; j < i; j++) {
        mcdist_noncircular(&gl[i], &gl[j], n, 0, &dm, &one);            /* the wrap: one pair on the GPU */
        __real_mcdist_noncircular(&gl[i], &gl[j], n, 0, &dm, &ref);     /* the reference's arithmetic */
        if (rows[i][j].d != ref.d || rows[i][j].br != ref.br || rows[i][j].c4 != ref.c4 || rows[i][j].h != ref.h ||
            rows[i][j].f != ref.f || rows[i][j].n != 2 * ref.d - ref.br ||
            one.d != ref.d || one.br != ref.br || one.c4 != ref.c4 || one.h != ref.h || one.f != ref.f) {
          printf("MISMATCH n=%d pair (%d,%d): batched d=%d b=%d c=%d h=%d f=%d, wrap d=%d, reference d=%d b=%d c=%d h=%d f=%d\n",
                 n, i, j, rows[i][j].d, rows[i][j].br, rows[i][j].c4, rows[i][j].h, rows[i][j].f, one.d,
                 ref.d, ref.br, ref.c4, ref.h, ref.f);
          return 1;
        }
        checked++;
      }
    mcdist_freemem(&dm);
  }
  printf("OK %d\n", checked);
  return 0;
}
