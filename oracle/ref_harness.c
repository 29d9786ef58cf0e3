/* ref_harness.c -- TEST INFRASTRUCTURE ONLY (checker / CPU baseline), never linked into the product.
 *
 * A thin caller of the UNMODIFIED reference GRIMM 2.01 functions, compiled together with the
 * reference's own object files (everything except mcmain.o) into oracle/_ref/libgrimmref.so by
 * oracle/Makefile.  No reference arithmetic is restated here: every number returned comes out of
 * the reference's invdist_noncircular_v (G/uniinvdist.c:235) or print_scenario_2
 * (G/opt_scenario.c:73).  The reference parser's MAX_NUM_GENES limit (G/mcstructs.h:42) is
 * not involved because genomes arrive as arrays.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "mcstructs.h"
#include "uniinvdist.h"
#include "mcrdist.h"
#include "scenario.h"
#include "opt_scenario.h"
#include "unsigned.h"

FILE *outfile; /* the global the reference prints to (defined in G/mcmain.c:66, which we leave out) */

static void stats_one(const int *gamma, const int *pi, int n, distmem_t *dm, int *out5)
{
  struct genome_struct g1, g2;
  graphstats_t gs;
  memset(&g1, 0, sizeof g1);
  memset(&g2, 0, sizeof g2);
  g1.genes = (int *)gamma;
  g2.genes = (int *)pi;
  invdist_noncircular_v(&g1, &g2, 0, n, dm, &gs);
  out5[0] = gs.d; out5[1] = gs.br; out5[2] = gs.c4; out5[3] = gs.h; out5[4] = gs.f;
}

/* one pair: out5 = {d, b, c, h, f} */
int ref_dist_stats(const int *gamma, const int *pi, int n, int *out5)
{
  distmem_t dm;
  mcdist_allocmem(n, 0, &dm);
  stats_one(gamma, pi, n, &dm, out5);
  mcdist_freemem(&dm);
  return out5[0];
}

typedef struct {
  const int *genomes; int n;
  const int *pairs; long lo, hi;
  int *out; int stride;
} job_t;

static void *worker(void *arg)
{
  job_t *j = (job_t *)arg;
  distmem_t dm;
  int tmp[5];
  long p;
  mcdist_allocmem(j->n, 0, &dm);
  for (p = j->lo; p < j->hi; p++) {
    const int *ga = j->genomes + (long)j->pairs[2*p] * j->n;
    const int *pi = j->genomes + (long)j->pairs[2*p+1] * j->n;
    stats_one(ga, pi, j->n, &dm, tmp);
    if (j->stride == 1) j->out[p] = tmp[0];
    else memcpy(j->out + 5*p, tmp, sizeof tmp);
  }
  mcdist_freemem(&dm);
  return NULL;
}

/* pairs[2p] = gamma index, pairs[2p+1] = pi index (setinvmatrix order, G/uniinvdist.c:80-90).
 * stride 1 -> out[p] = d ; stride 5 -> out[5p..] = {d,b,c,h,f}.  nthreads host threads, each with
 * its own distmem_t exactly like one reference process would have. */
void ref_dist_pairs(const int *genomes, int n, const int *pairs, long npairs,
                    int *out, int stride, int nthreads)
{
  int t;
  if (nthreads < 1) nthreads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  job_t *jobs = (job_t *)malloc(sizeof(job_t) * nthreads);
  for (t = 0; t < nthreads; t++) {
    jobs[t].genomes = genomes; jobs[t].n = n; jobs[t].pairs = pairs;
    jobs[t].lo = npairs * t / nthreads; jobs[t].hi = npairs * (t + 1) / nthreads;
    jobs[t].out = out; jobs[t].stride = stride;
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

/* Run the reference's default -s engine (scenario_type 4, G/mcmain.c:240,948) from src to dst and
 * return the step list parsed back out of the text it prints: steps[2k]=start, steps[2k+1]=end.
 * Returns the number of steps, or -1 if more than maxsteps.  Not thread-safe (global outfile). */
int ref_scenario(const int *src, const int *dst, int n, int *steps, int maxsteps)
{
  struct genome_struct g1, g2;
  char *buf = NULL; size_t len = 0;
  FILE *saved = outfile;
  int nsteps = 0;
  char *p;
  memset(&g1, 0, sizeof g1); memset(&g2, 0, sizeof g2);
  g1.genes = (int *)src; g2.genes = (int *)dst;
  outfile = open_memstream(&buf, &len);
  print_scenario_2(&g1, &g2, n, 0, 4);
  fclose(outfile);
  outfile = saved;
  for (p = buf; (p = strstr(p, "Step ")) != NULL; p += 5) {
    int k, s, sv, e, ev;
    if (sscanf(p, "Step %d: Chrom. 0, gene %d [%d] through chrom. 0, gene %d [%d]",
               &k, &s, &sv, &e, &ev) == 5) {
      if (nsteps >= maxsteps) { free(buf); return -1; }
      steps[2*nsteps] = s; steps[2*nsteps+1] = e; nsteps++;
    }
  }
  free(buf);
  return nsteps;
}

/* The first `cap` steps of the reference's scenario (all of it when it is shorter).  print_scenario_2
 * has no step limit, so it runs in a forked child that prints into a pipe; the parent parses the
 * "Step k:" lines as they arrive and kills the child once it has `cap` of them.  This is how the
 * n = 100000 pair of config 5 (days of CPU for the whole scenario) and hurdle-rich n = 5000 pairs
 * (minutes per pair) get reference step prefixes.  Returns the number of steps read. */
#include <unistd.h>
#include <signal.h>
#include <sys/wait.h>
int ref_scenario_prefix(const int *src, const int *dst, int n, int *steps, int cap)
{
  int fd[2], nsteps = 0;
  pid_t pid;
  FILE *in;
  char *line = NULL; size_t lcap = 0;
  if (pipe(fd) != 0) return -1;
  pid = fork();
  if (pid < 0) return -1;
  if (pid == 0) {
    struct genome_struct g1, g2;
    close(fd[0]);
    memset(&g1, 0, sizeof g1); memset(&g2, 0, sizeof g2);
    g1.genes = (int *)src; g2.genes = (int *)dst;
    outfile = fdopen(fd[1], "w");
    print_scenario_2(&g1, &g2, n, 0, 4);
    fflush(outfile);
    _exit(0);
  }
  close(fd[1]);
  in = fdopen(fd[0], "r");
  while (nsteps < cap && getline(&line, &lcap, in) > 0) {
    int k, s, sv, e, ev;
    if (strncmp(line, "Step ", 5) == 0 &&
        sscanf(line, "Step %d: Chrom. 0, gene %d [%d] through chrom. 0, gene %d [%d]",
               &k, &s, &sv, &e, &ev) == 5) {
      steps[2*nsteps] = s; steps[2*nsteps+1] = e; nsteps++;
    }
  }
  kill(pid, SIGKILL);
  fclose(in);
  waitpid(pid, NULL, 0);
  free(line);
  return nsteps;
}

/* unsigned_mcdist_nomem (G/unsigned.c:786-830) on copies of the two genomes: returns the distance,
 * out_g1 / out_g2 = the genomes as the reference leaves them (optimal signs, circular shifts),
 * *approx = dist_error.  Not thread-safe (the reference prints through the global outfile). */
int ref_unsigned_dist(const int *g1, const int *g2, int n, int circular, int *out_g1, int *out_g2, int *approx)
{
  struct genome_struct a, b;
  int d, err = 0;
  FILE *saved = outfile;
  memset(&a, 0, sizeof a); memset(&b, 0, sizeof b);
  memcpy(out_g1, g1, sizeof(int) * (size_t)n);
  memcpy(out_g2, g2, sizeof(int) * (size_t)n);
  a.genes = out_g1; b.genes = out_g2;
  outfile = stderr;
  d = unsigned_mcdist_nomem(&a, &b, n, 0, circular, 0, &err);
  outfile = saved;
  *approx = err;
  return d;
}
