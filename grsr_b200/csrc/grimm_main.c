/* grimm_main.c -- host program with GRIMM 2.01's command line for the unichromosomal path
 * (reference G/mcmain.c:195-996), calling the B200 kernels through the C ABI of include/grsr.h.
 *
 *   grimm -f datafile [-o outfile] -L|-C [-g i,j] [-m] [-d] [-s] [-S 1..7] [-t 1] [-u] [-v]
 *
 * What is kept byte for byte (GRSR's scr/reptA.sh:61-62,453-454 greps this text):
 *   input format          mcread_data, check_genes, delete_caps   G/mcread_input.c:83-554
 *   -C alignment          circular_align                          G/circ_align.c:50-188
 *   report header         G/mcmain.c:655-686
 *   pairwise report       G/mcmain.c:798-953, opt_printstep G/opt_scenario.c:1001-1060,
 *                         print_genome_multichrom_3 G/write_data.c:131-221
 *   matrix reports        do_fn_distmatrix[_v], print_full_distmatrix, print_matrix_offset
 *                         G/mcmain.c:1009-1283
 *   error texts + exit(-1)
 * What is NOT here: every number in the report comes from libgrsr_cuda.so (grsr_dist_pairs,
 * grsr_dist_matrix, grsr_stats_matrix, grsr_scenario_batch, grsr_trial_distances, grsr_unsigned_dist
 * for -u: unsigned_mcdist_nomem G/mcmain.c:808-815, set_unsigneddist_matrix :1155-1164); this file holds
 * no distance or scenario arithmetic.  -t 1 (print_rev_stats_t1, G/testrev.c:86-230) scores every
 * breakpoint-to-breakpoint reversal in one batched call.  Options outside the accelerated path
 * (multichromosomal mode, -c -z -t 2..4 -U -W -T -F -X -D -M) are refused with a message
 * instead of being silently mis-served.
 *
 * Resident mode (not in the reference; it exists because of how GRSR calls the program): scr/reptA.sh:61
 * starts one `grimm` per genome pair on ~30-block genomes, and creating a CUDA context costs a process
 * 0.4-1 s on a B200 box (profiles/r2_startup.txt) against 0.7 ms for the whole reference run.
 *     grimm --server /path/to/socket &        one process that keeps the context (and the kernels) warm
 *     GRSR_SERVER=/path/to/socket grimm ...   every other invocation: sends its argv, working directory,
 *                                             stdout and stderr to that process and returns its exit status
 * Without GRSR_SERVER, or when nobody listens there, the invocation runs in its own process as usual.
 */
#include <ctype.h>
#include <errno.h>
#include <setjmp.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include "grsr.h"

/* In resident mode a request must not end the process: every exit() of the report code below lands in
 * grimm_exit, which returns to the server loop instead. */
static jmp_buf server_jmp;
static int serving = 0;
static void grimm_exit(int status)
{
  if (serving) longjmp(server_jmp, (status & 255) + 1);
  exit(status);
}
#define exit(status) grimm_exit(status)

#define MAX_STR_LEN 1024 /* G/mcstructs.h:47 */
#define GNAME_WIDTH 20   /* G/mcmain.c:1008 */

static FILE *outfile;

static void usage(const char *prog)
{
  fprintf(stderr, "\nGRIMM 2.01 command line, grsr-b200 build (CUDA sm_100a, unichromosomal path)\n");
  fprintf(stderr, "\nUsage: %s -f datafile [-o outfile] [other options]\n\n", prog);
  fprintf(stderr, "   -f datafile: name of file containing all the genomes\n");
  fprintf(stderr, "   -o outfile: output file name.  Output goes to STDOUT unless -o used.\n");
  fprintf(stderr, "   -v: Verbose output.\n");
  fprintf(stderr, "       In pairwise mode, outputs all the graph parameters.\n");
  fprintf(stderr, "       In matrix mode, outputs matrices for all graph parameters.\n");
  fprintf(stderr, "\nGenome type (one is required by this build):\n");
  fprintf(stderr, "   -L: unichromosomal (directed) linear\n");
  fprintf(stderr, "   -C: unichromosomal circular\n");
  fprintf(stderr, "\nGenome selection:\n");
  fprintf(stderr, "   -g i,j: compare genomes i and j (counting from 1)\n");
  fprintf(stderr, "      Rearrangement scenario will go from genome i to genome j.\n");
  fprintf(stderr, "      For files with 3 or more genomes, unless -g is specified, a distance\n");
  fprintf(stderr, "      matrix is printed and options -d, -s are not available.\n");
  fprintf(stderr, "   -m: force matrix output even for 2 genomes\n");
  fprintf(stderr, "\nFunction: (default -d -s)\n");
  fprintf(stderr, "   -d: distance\n");
  fprintf(stderr, "   -s: display an optimal scenario\n");
  fprintf(stderr, "   -S #: same, #=2,...,7 (all equivalent for unichromosomal genomes)\n");
  fprintf(stderr, "   -t 1: test all 1-step reversals, list the ones reducing distance by 1,\n");
  fprintf(stderr, "       and give statistics\n");
  fprintf(stderr, "\nFormat of datafile: each entry has the form\n");
  fprintf(stderr, "   >human\n   5 -2 3 -1 $\n   >mouse\n   ...\n");
  fprintf(stderr, "Genes are separated by any whitespace; \"#\" starts a comment.\n\n");
}

static void die_usage(const char *prog)
{
  usage(prog);
  exit(-1);
}

static void *xmalloc(size_t bytes, const char *what)
{
  void *p = malloc(bytes ? bytes : 1);
  if (!p) { /* style of G/e_malloc.c:40-61 */
    fprintf(stderr, "ERROR: Could not allocate memory for %s\n", what);
    exit(-1);
  }
  return p;
}

/* ------------------------------------------------------------------------------------------------
 * input: the whole file is read once, then scanned twice like mcread_data's two passes
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
  const char *buf;
  long len, pos;
} src_t;

static int src_getc(src_t *s) { return s->pos < s->len ? (unsigned char)s->buf[s->pos++] : EOF; }

/* discard_line (G/mcread_input.c:150-169): stops after '\n' or '\r'; refuses over-long lines */
static void skip_line(src_t *s)
{
  int c, numchars = 0;
  while ((c = src_getc(s)) != EOF && c != '\n' && c != '\r') {
    numchars++;
    if (numchars >= MAX_STR_LEN) {
      fflush(outfile);
      fprintf(stderr, "ERROR: \tBuffer overflow reading input.\n");
      fprintf(stderr, "\tMCDIST is compiled with a maximum ");
      fprintf(stderr, "read buffer length of %d.\n", MAX_STR_LEN);
      exit(-1);
    }
  }
}

typedef struct {
  int num_genomes;
  int num_genes;       /* real genes per genome (caps never materialised) */
  int num_chromosomes; /* max chromosomes per genome, as counted by the reference */
  int32_t *genes;      /* num_genomes x num_genes, row-major */
  char **names;
} input_t;

static void check_genome(const int32_t *genes, int n, int num_chromosomes, int at_genome)
{
  /* check_genes (G/mcread_input.c:83-135) runs on the capped genome of n + 2*chromosomes genes.
   * One chromosome: cap n+1, the genes, cap n+2 (G/mcread_input.c:423-425,496-500) -- checked
   * literally in that order.  More chromosomes are refused later anyway; there the caps are
   * treated as already seen. */
  int capped = n + 2 * num_chromosomes, i, total = n + 2;
  char *seen = (char *)xmalloc((size_t)capped + 1, "checkArr");
  memset(seen, 0, (size_t)capped + 1);
  if (num_chromosomes != 1) {
    for (i = n + 1; i <= capped; i++) seen[i] = 1;
    total = n;
  }
  for (i = 0; i < total; i++) {
    long g0, g1;
    if (num_chromosomes == 1) g0 = (i == 0) ? n + 1 : (i == n + 1 ? n + 2 : genes[i - 1]);
    else g0 = genes[i];
    g1 = g0 > 0 ? g0 : -g0;
    if (g1 == 0) {
      fflush(outfile);
      fprintf(stderr, "ERROR: Reading Input Genome %d\n", at_genome);
      fprintf(stderr, "ERROR: Input gene == 0\n");
      exit(-1);
    }
    if (g1 > capped) {
      fflush(outfile);
      fprintf(stderr, "ERROR: Reading Input Genome %d\n", at_genome);
      fprintf(stderr, "ERROR: Input gene (%ld) larger than number of genes %d\n", g0, capped);
      exit(-1);
    }
    if (seen[g1]) {
      fflush(outfile);
      fprintf(stderr, "ERROR: Reading Input Genome %d\n", at_genome);
      fprintf(stderr, "ERROR: Duplicate gene (%ld)\n", g1);
      exit(-1);
    }
    seen[g1] = 1;
  }
  free(seen);
}

static void read_input(FILE *fp, input_t *in)
{
  char *buf;
  long cap = 1 << 16, len = 0;
  size_t got;
  src_t s;
  int c, pass;
  int num_genomes = 0, num_genes = -1, num_chrom = 1;
  int gnum = 0, gnumc = 0, cnum = 0, in_seg = 0, seg_len = 0;

  buf = (char *)xmalloc((size_t)cap, "input buffer");
  while ((got = fread(buf + len, 1, (size_t)(cap - len), fp)) > 0) {
    len += (long)got;
    if (len == cap) {
      cap *= 2;
      buf = (char *)realloc(buf, (size_t)cap);
      if (!buf) { fprintf(stderr, "ERROR: Could not allocate memory for input buffer\n"); exit(-1); }
    }
  }
  fclose(fp);

  /* pass 1: count genomes, genes, chromosomes (G/mcread_input.c:231-336) */
  s.buf = buf; s.len = len; s.pos = 0;
  for (;;) {
    while ((c = src_getc(&s)) != EOF && isspace(c)) ;
    if (c == '#') { skip_line(&s); continue; }
    if (c == '>' || c == EOF) {
      if (num_genomes == 1) { num_genes = gnum; num_chrom = cnum; }
      if (num_genomes > 0 && num_genes != gnum) {
        fflush(outfile);
        fprintf(stderr, "ERROR: mcread_data() inconsistent num of genes\n");
        fprintf(stderr, "ERROR: Expecting %d genes, but ", num_genes);
        fprintf(stderr, "Genome %d has %d genes\n", num_genomes, gnum);
        exit(-1);
      }
      if (num_chrom < cnum) num_chrom = cnum;
      if (c == EOF) break;
      skip_line(&s);
      num_genomes++;
      gnum = gnumc = cnum = 0;
      continue;
    }
    if (num_genomes > 0) {
      if (c == ';' || c == '$') {
        gnumc = 0;
        if (in_seg) { fprintf(stderr, "Genome %d has unterminated [\n", num_genomes); exit(-1); }
        continue;
      }
      if (c == '(' || c == ')' || c == '{' || c == '}') continue;
      if (c == '[') {
        if (in_seg) { fprintf(stderr, "Genome %d has unbalanced [..[\n", num_genomes); exit(-1); }
        in_seg++; seg_len = 0;
        continue;
      }
      if (c == ']') {
        if (!in_seg) { fprintf(stderr, "Genome %d has unmatched ]\n", num_genomes); exit(-1); }
        else if (seg_len == 0) { fprintf(stderr, "Genome %d has empty []\n", num_genomes); exit(-1); }
        else in_seg--;
      }
      if (c == '-' || isdigit(c)) {
        gnum++; gnumc++;
        if (gnumc == 1) cnum++;
        while (s.pos < s.len && isdigit((unsigned char)s.buf[s.pos])) s.pos++;
        if (in_seg) seg_len++;
      }
      continue; /* any other character inside a genome is skipped, as the reference does */
    }
    fflush(outfile);
    fprintf(stderr, "ERROR: Improper file format.  Bad character %c at position %d.\n", c, (int)s.pos);
    exit(-1);
  }

  in->num_genomes = num_genomes;
  in->num_genes = num_genes;
  in->num_chromosomes = num_chrom;
  in->genes = NULL; in->names = NULL;
  if (num_genomes == 0 || num_genes + 2 * num_chrom == 0) { free(buf); return; }

  in->genes = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)num_genomes * (size_t)(num_genes > 0 ? num_genes : 1), "genes");
  in->names = (char **)xmalloc(sizeof(char *) * (size_t)num_genomes, "genome_list");

  /* pass 2: store and validate (G/mcread_input.c:401-524) */
  s.pos = 0;
  {
    int genome = -1;
    gnum = 0;
    for (pass = 0;; pass++) {
      while ((c = src_getc(&s)) != EOF && isspace(c)) ;
      if (c == '#') { skip_line(&s); continue; }
      if (c == '>' || c == EOF) {
        if (genome >= 0)
          check_genome(in->genes + (size_t)genome * (size_t)num_genes, num_genes, num_chrom, genome + 1);
        if (c == EOF) break;
        genome++;
        { /* fgets(buf, MAX_STR_LEN) then cut at the first '\n' or '\r' (G/mcread_input.c:440-450) */
          char *nm = (char *)xmalloc(MAX_STR_LEN, "gnamePtr");
          int k = 0, ch;
          while (k < MAX_STR_LEN - 1 && (ch = src_getc(&s)) != EOF) {
            nm[k++] = (char)ch;
            if (ch == '\n') break;
          }
          nm[k] = '\0';
          for (k = 0; nm[k]; k++) if (nm[k] == '\n' || nm[k] == '\r') { nm[k] = '\0'; break; }
          in->names[genome] = nm;
        }
        gnum = 0;
        continue;
      }
      if (genome >= 0) {
        if (c == '-' || isdigit(c)) { /* fscanf("%d") after pushing the lead character back */
          long v = 0; int neg = (c == '-'), any = 0;
          if (!neg) { v = c - '0'; any = 1; }
          while (s.pos < s.len && isdigit((unsigned char)s.buf[s.pos])) {
            v = v * 10 + (s.buf[s.pos++] - '0'); any = 1;
            if (v > 2000000000L) v = 2000000000L;
          }
          if (!any) v = 0;
          if (gnum < num_genes) in->genes[(size_t)genome * (size_t)num_genes + gnum] = (int32_t)(neg ? -v : v);
          gnum++;
        }
        continue;
      }
      fflush(outfile);
      fprintf(stderr, "ERROR: Pass 2, Improper file format.  Bad character %c at position %d.\n", c, (int)s.pos);
      exit(-1);
    }
  }
  free(buf);
}

/* circular_align(genomes, G, n, 0, TRUE) (G/circ_align.c:50-156): every genome after the first is
 * rotated, or flipped, so that it starts with the first gene of genome 1 */
static void reverse_unsigned(int32_t *a, int lo, int hi)
{
  for (; lo < hi; lo++, hi--) { int32_t t = a[lo]; a[lo] = a[hi]; a[hi] = t; }
}

static void reverse_signed(int32_t *a, int lo, int hi)
{
  for (; lo <= hi; lo++, hi--) { int32_t t = a[lo]; a[lo] = -a[hi]; a[hi] = -t; }
}

static void circular_align(int32_t *genes, int G, int n)
{
  int32_t first = genes[0];
  int g, pos;
  for (g = 1; g < G; g++) {
    int32_t *row = genes + (size_t)g * (size_t)n;
    int sign = 1;
    for (pos = 0; pos < n; pos++) {
      if (row[pos] == first) { sign = 1; break; }
      if (row[pos] == -first) { sign = -1; break; }
    }
    if (sign > 0) {
      if (pos == 0) continue;
      reverse_unsigned(row, 0, pos - 1);
      reverse_unsigned(row, pos, n - 1);
      reverse_unsigned(row, 0, n - 1);
    } else {
      reverse_signed(row, 0, pos);
      reverse_signed(row, pos + 1, n - 1);
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * printing
 * ---------------------------------------------------------------------------------------------- */

static int gene_width(int n) { return n >= 1000 ? 7 : (n >= 100 ? 4 : 3); } /* G/write_data.c:108-113 */

static void print_row(const int32_t *genes, int n)
{ /* print_genome_multichrom_3(..., width -1, multiline 0, showcaps 2, showname 0), no caps present */
  int w = gene_width(n), i;
  for (i = 0; i < n; i++) fprintf(outfile, "%*d ", w, genes[i]);
  if (n > 0) fprintf(outfile, "\n");
  fflush(outfile);
}

static void print_name(const char *name)
{ /* truncate or pad to GNAME_WIDTH, then a tab (G/mcmain.c:1031-1040) */
  int k, pad = 0;
  for (k = 0; k < GNAME_WIDTH; k++) {
    if (!pad) { if (name[k] != '\0') fputc(name[k], outfile); else pad = 1; }
    if (pad) fputc(' ', outfile);
  }
  fputc('\t', outfile);
}

static int num_digits(int v)
{ /* G/matrixmisc.c:39-54 */
  int k = 1;
  if (v < 0) { k++; v = -v; }
  while (v >= 10) { v /= 10; k++; }
  return k;
}

/* the reference computes the column width from max(0, negative entries) (G/mcmain.c:1015-1024),
 * which is 0 for every matrix it can print: the width is always num_digits(0) = 1 */
static int entry_width(const int32_t *m, int stride, int G)
{
  int i, j, mx = 0;
  for (i = 0; i < G; i++)
    for (j = 0; j < i; j++)
      if (mx > m[((size_t)j * G + i) * stride]) mx = m[((size_t)j * G + i) * stride];
  return num_digits(mx);
}

static void print_matrix_body(const int32_t *m, int stride, int G, char **names, int width)
{
  int i, j;
  for (i = 0; i < G; i++) {
    print_name(names[i]);
    for (j = 0; j < G; j++) fprintf(outfile, "%*d ", width, m[((size_t)i * G + j) * stride]);
    fprintf(outfile, "\n");
  }
  fprintf(outfile, "\n");
}

static void gpu_fail(int rc)
{
  fflush(outfile);
  fprintf(stderr, "ERROR: %s\n", grsr_last_error());
  (void)rc;
  exit(-1);
}

static int first_device(void)
{
  const char *s = getenv("GRSR_DEVICE");
  return (s && *s) ? atoi(s) : 0;
}

static int num_gpus(void)
{
  const char *s = getenv("GRSR_NGPUS");
  int k = (s && *s) ? atoi(s) : 1;
  return k < 1 ? 1 : k;
}

/* -S 1: print_scenario (G/scenario.c:190-433), the legacy engine.  Per step the candidates are the
 * reversals [i..j] whose two ends are breakpoints (:270-281, :313-321), i-major / j-minor; the first
 * one whose trial permutation is closer to the destination wins (d2 < d, :343).  Every candidate's
 * distance comes from grsr_trial_distances, a chunk of the ordered list at a time, until one
 * succeeds.  Rows are printed with print_genes (" %4d" per gene, G/scenario.c:451-463). */
static void print_genes_legacy(FILE *out, const int32_t *genes, int n)
{
  int k;
  for (k = 0; k < n; k++) fprintf(out, " %4d", genes[k]);
  fprintf(out, "\n");
}

static void legacy_scenario(FILE *out, const int32_t *g1, const int32_t *g2, int n, int d)
{
  const long chunk = 4096;
  int32_t *cur = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)n, "current_genome");
  int32_t *succ = (int32_t *)xmalloc(sizeof(int32_t) * (2 * (size_t)n + 1), "successors");
  int32_t *cands = (int32_t *)xmalloc(sizeof(int32_t) * 2 * (size_t)chunk, "candidates");
  int32_t *d2 = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)chunk, "trial distances");
  int i, j, t = 1;
  memcpy(cur, g1, sizeof(int32_t) * (size_t)n);
  for (i = 0; i < 2 * n + 1; i++) succ[i] = 0;
  for (i = 0; i + 1 < n; i++) { succ[g2[i] + n] = g2[i + 1]; succ[-g2[i + 1] + n] = -g2[i]; }
  fprintf(out, "\n======================================================================\n\n");
  fprintf(out, "An optimal sequence of reversals:\n");
  fprintf(out, "Step 0: (Source)\n");
  print_genes_legacy(out, cur, n);
  while (d > 0) {
    int found = 0, start = 0, end = 0, newd = d;
    long filled = 0, c;
    i = 0; j = -1;                                   /* cursor over the ordered candidate list */
    for (;;) {
      /* next candidate after (i, j) */
      int have = 0;
      for (; i < n; i++, j = -1) {
        int is_start = (i == 0) ? (cur[0] != g2[0]) : (succ[cur[i - 1] + n] != cur[i]);
        if (!is_start) continue;
        if (j < i) j = i - 1;
        for (j = j + 1; j < n; j++) {
          int is_end = (j == n - 1) ? (cur[j] != g2[j]) : (succ[cur[j] + n] != cur[j + 1]);
          if (is_end) { have = 1; break; }
        }
        if (have) break;
      }
      if (have) { cands[2 * filled] = i; cands[2 * filled + 1] = j; filled++; }
      if (filled == chunk || (!have && filled > 0)) {
        int rc = grsr_trial_distances(cur, g2, n, cands, filled, d2, first_device());
        if (rc) gpu_fail(rc);
        for (c = 0; c < filled; c++) {
          if (d2[c] - d > 1 || d2[c] - d < -1) {     /* the reference's self-check (:323-333) */
            int32_t *trial = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)n, "trial_genome");
            int a = cands[2 * c], b = cands[2 * c + 1], x, y;
            memcpy(trial, cur, sizeof(int32_t) * (size_t)n);
            for (x = a, y = b; x <= y; x++, y--) { int32_t tmp = trial[x]; trial[x] = -trial[y]; trial[y] = -tmp; }
            fprintf(out, "\nERROR: Potential distance caculation error:\n");
            fprintf(out, "Source:\n\t");
            print_genes_legacy(out, cur, n);
            fprintf(out, "Reversal from gene %d to gene %d:\n\t", a, b);
            print_genes_legacy(out, trial, n);
            free(trial);
          }
          if (d2[c] < d) { found = 1; start = cands[2 * c]; end = cands[2 * c + 1]; newd = d2[c]; break; }
        }
        filled = 0;
        if (found) break;
      }
      if (!have) break;
    }
    if (!found) {
      fprintf(out, "ERROR: reversals ended prematurely\n");
      break;
    }
    d = newd;
    {
      int startvalue = cur[start], endvalue = cur[end], x, y;
      for (x = start, y = end; x <= y; x++, y--) { int32_t tmp = cur[x]; cur[x] = -cur[y]; cur[y] = -tmp; }
      fprintf(out, "Step %d: ", t);
      fprintf(out, "Chrom. %d, gene %d [%d] through chrom. %d, gene %d [%d]: ", 0, start, startvalue, 0, end, endvalue);
      fprintf(out, "%s", "Reversal");
      if (d == 0) fprintf(out, " (Destination)");
      fprintf(out, "\n");
      print_genes_legacy(out, cur, n);
    }
    t++;
  }
  free(cur); free(succ); free(cands); free(d2);
}

static int grimm_main(int argc, char *argv[])
{
  int c, errflg = 0, i;
  const char *inputfname = NULL, *outputfname = NULL;
  struct stat statbuf;
  FILE *input;
  int unichromosome = 0, circular = 0, verbose = 0, unsigned_dist = 0;
  int fn_distance = 0, fn_scenario = 0, scenario_type = 4, fn_any = 0, fn_matrix = 0, fn_revstats = 0;
  int gindex1 = -1, gindex2 = -1;
  char refused = 0;
  input_t in;
  int G, n;

  outfile = stdout;

  while ((c = getopt(argc, argv, "f:o:CLdczsmt:uvM:D:g:T:S:X:F:U:W:")) != -1) {
    switch (c) {
    case 'f': inputfname = optarg; break;
    case 'o': outputfname = optarg; break;
    case 'L': unichromosome = 1; circular = 0; break;
    case 'C': unichromosome = 1; circular = 1; break;
    case 'd': fn_distance = 1; fn_any = 1; break;
    case 's': fn_scenario = 1; fn_any = 1; break;
    case 'S':
      fn_scenario = 1; fn_any = 1;
      if (sscanf(optarg, "%d", &scenario_type) != 1 || scenario_type <= 0 || scenario_type > 7) {
        fprintf(stderr, "invalid option for -S\n");
        errflg++;
      }
      break;
    case 'm': fn_matrix = 1; break;
    case 'v': verbose = 1; break;
    case 'g':
      if (sscanf(optarg, "%d,%d", &gindex1, &gindex2) != 2 || gindex1 <= 0 || gindex2 <= 0) {
        fprintf(stderr, "-g needs two genome numbers\n");
        errflg++;
      } else { gindex1--; gindex2--; }
      break;
    case 't':
      if (sscanf(optarg, "%d", &fn_revstats) != 1 || fn_revstats < 1 || fn_revstats > 4) {
        fprintf(stderr, "ERROR: -t #   must be between 1 and 4\n");
        errflg++;
      }
      fn_any = 1;
      break;
    case 'u': unsigned_dist = 1; break;
    case 'c': case 'z': case 'M': case 'D': case 'T': case 'X': case 'F':
    case 'U': case 'W':
      refused = (char)c;
      break;
    case '?':
      fprintf(stderr, "Unrecognized option: -%c\n", optopt);
      errflg++;
      break;
    default:
      break;
    }
  }
  if (errflg) die_usage(argv[0]);
  if (refused) {
    fprintf(stderr, "ERROR: option -%c is outside the path this build accelerates "
                    "(unichromosomal -L/-C with -g -m -d -s -S -t 1 -u -v); use the original GRIMM for it\n", refused);
    die_usage(argv[0]);
  }

  if (inputfname == NULL) {
    fprintf(stderr, "ERROR: input filename required\n");
    die_usage(argv[0]);
  }
  if (stat(inputfname, &statbuf)) {
    fprintf(stderr, "ERROR: Input file (%s): ", inputfname);
    perror("");
    die_usage(argv[0]);
  }
  if (statbuf.st_mode & S_IFDIR) {
    fprintf(stderr, "ERROR: Input file (%s) is a directory.\n", inputfname);
    die_usage(argv[0]);
  }
  input = fopen(inputfname, "r");
  if (input == NULL) {
    fprintf(stderr, "ERROR: Could not open input file (%s): ", inputfname);
    perror("");
    die_usage(argv[0]);
  }
  if (outputfname != NULL) {
    if (!stat(outputfname, &statbuf)) {
      fprintf(stderr, "ERROR: Output file (%s) already exists.\n", outputfname);
      die_usage(argv[0]);
    }
    outfile = fopen(outputfname, "a+");
    if (outfile == NULL) {
      fprintf(stderr, "ERROR: Could not open output file (%s): ", outputfname);
      perror("");
      die_usage(argv[0]);
    }
  }

  read_input(input, &in);
  if (in.num_genomes == 0 || in.num_genes + 2 * in.num_chromosomes == 0) {
    fprintf(stderr, "No data found in file.\n");
    die_usage(argv[0]);
  }
  G = in.num_genomes; n = in.num_genes;

  if (fn_matrix && gindex1 != -1) {
    fprintf(stderr, "The -g and -m options may not be used together.\n");
    die_usage(argv[0]);
  }
  if ((G > 2 && gindex1 == -1) || fn_matrix) {
    if (fn_any) {
      fprintf(stderr, "A distance matrix is displayed when there are multiple genomes or\n");
      fprintf(stderr, "the -m option is given.  The -d -c -s -t -T options are not available;\n");
      fprintf(stderr, "you must specify two genomes with the -g #,# option.\n");
    }
    fn_scenario = fn_distance = 0;
    fn_revstats = 0;
    fn_any = fn_matrix = 1;
  }
  if (gindex1 == -1) { gindex1 = 0; gindex2 = 1; }
  if (gindex1 >= G || gindex2 >= G) {
    fprintf(stderr, "ERROR: specified %d genomes, but requested -g %d,%d\n", G, gindex1 + 1, gindex2 + 1);
    die_usage(argv[0]);
  }
  if (unichromosome && in.num_chromosomes > 1) {
    fprintf(stderr, "ERROR: -C and -L are only for unichromosomal genomes, but have %d chromosomes\n",
            in.num_chromosomes);
    die_usage(argv[0]);
  }
  if (!unichromosome) {
    fprintf(stderr, "ERROR: multichromosomal mode is outside the path this build accelerates; "
                    "give -L (linear) or -C (circular), or use the original GRIMM\n");
    die_usage(argv[0]);
  }
  if (n < 1) { /* genomes without genes: the reference goes on with zero-length arrays; refuse */
    fprintf(stderr, "No data found in file.\n");
    die_usage(argv[0]);
  }

  if (circular) circular_align(in.genes, G, n);

  fprintf(outfile, "Running:                      \t%s\n", argv[0]);
  fprintf(outfile, "Input:                        \t%s\n", inputfname);
  fprintf(outfile, "Genome:                       \t%s\n", circular ? "Circular" : "Linear");
  fprintf(outfile, "Signs:                        \t%s\n", unsigned_dist ? "Unsigned" : "Signed");
  fprintf(outfile, "Number of Genomes:            \t%d\n", G);
  fprintf(outfile, "Number of Genes:              \t%d\n", n);
  fprintf(outfile, "Number of Chromosomes:        \t%d", 0);
  fprintf(outfile, " (unichromosomal)\n");
  fflush(outfile);

  if (fn_matrix && unsigned_dist) {
    /* do_fn_distmatrix with set_unsigneddist_matrix (G/mcmain.c:1155-1164, G/unsigned.c:1107-1131): every
     * pair through the unsigned distance; -v is ignored here as in the reference (:1133) */
    int32_t *dm = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)G * (size_t)G, "distmatrix");
    int j, dist_error = 0;
    fprintf(outfile, "\nDistance Matrix:\n");
    for (i = 0; i < G; i++) {
      dm[(size_t)i * G + i] = 0;
      for (j = i + 1; j < G; j++) {
        grsr_unsigned_info_t info;
        int rc = grsr_unsigned_dist(in.genes + (size_t)i * n, in.genes + (size_t)j * n, n, circular, NULL, NULL,
                                    &info, first_device());
        if (rc) gpu_fail(rc);
        dm[(size_t)i * G + j] = dm[(size_t)j * G + i] = info.dist;
        if (info.approx) dist_error = 1;
      }
    }
    if (dist_error) {
      fprintf(outfile, "WARNING: These unsigned permutations are too complex;\n");
      fprintf(outfile, "some answers are approximated.\n");
    }
    fprintf(outfile, "\n   %d\n", G);
    print_matrix_body(dm, 1, G, in.names, entry_width(dm, 1, G));
    free(dm);
  } else if (fn_matrix) {
    int rc;
    if (verbose) { /* do_fn_distmatrix_v, G/mcmain.c:1206-1283 */
      static const char *titles[5] = {"Distance", "Number of Breakpoints", "Number of Cycles",
                                      "Number of Hurdles", "Number of Fortresses"};
      grsr_stats_t *sm = (grsr_stats_t *)xmalloc(sizeof(grsr_stats_t) * (size_t)G * (size_t)G, "statsmatrix");
      int f;
      rc = grsr_stats_matrix(in.genes, G, n, sm, first_device(), num_gpus());
      if (rc) gpu_fail(rc);
      for (f = 0; f < 5; f++) {
        const int32_t *base = (const int32_t *)sm + f;
        fprintf(outfile, "\n");
        fprintf(outfile, "%s Matrix:\n", titles[f]);
        fprintf(outfile, "   %d\n", G);
        print_matrix_body(base, 5, G, in.names, entry_width(base, 5, G));
      }
      free(sm);
    } else { /* do_fn_distmatrix + print_full_distmatrix, G/mcmain.c:1122-1191, 1009-1056 */
      int32_t *dm = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)G * (size_t)G, "distmatrix");
      fprintf(outfile, "\nDistance Matrix:\n");
      rc = grsr_dist_matrix(in.genes, G, n, dm, first_device(), num_gpus());
      if (rc) gpu_fail(rc);
      fprintf(outfile, "\n   %d\n", G);
      print_matrix_body(dm, 1, G, in.names, entry_width(dm, 1, G));
      free(dm);
    }
  } else {
    grsr_stats_t st;
    int32_t pair[2];
    int rc;
    const int32_t *g1 = in.genes + (size_t)gindex1 * (size_t)n;
    const int32_t *g2 = in.genes + (size_t)gindex2 * (size_t)n;

    if (!fn_any) { fn_distance = 1; fn_scenario = 1; }
    if (fn_revstats > 1) {
      fprintf(stderr, "ERROR: -t %d is outside the path this build accelerates (only -t 1 is "
                      "unichromosomal); use the original GRIMM\n", fn_revstats);
      die_usage(argv[0]);
    }
    pair[0] = gindex1; pair[1] = gindex2;
    if (unsigned_dist) {
      /* unsigned_mcdist_nomem (G/mcmain.c:808-815): the two genomes are overwritten with their optimal
       * signs (circular: also shifts), everything after this works on the signed genomes */
      grsr_unsigned_info_t info;
      int32_t *s1 = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)n, "signed genome");
      int32_t *s2 = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)n, "signed genome");
      int q;
      rc = grsr_unsigned_dist(g1, g2, n, circular, s1, s2, &info, first_device());
      if (rc) gpu_fail(rc);
      if (circular) memcpy(in.genes + (size_t)gindex1 * (size_t)n, s1, sizeof(int32_t) * (size_t)n);
      memcpy(in.genes + (size_t)gindex2 * (size_t)n, s2, sizeof(int32_t) * (size_t)n);
      free(s1); free(s2);
      if (verbose)
        for (q = 0; q < info.passes; q++) {           /* assign_canon_signs, once per pass (G/unsigned.c:377-381) */
          fprintf(outfile, "Number of Singletons:         \t%d\n", info.num_sing[q]);
          fprintf(outfile, "Number of 2-strips:           \t%d\n", info.num_doub[q]);
        }
      if (info.approx)
        fprintf(outfile, "WARNING: These unsigned permutations are too complex; answers are approximated.\n");
      st.d = info.dist; st.b = st.c = st.h = st.f = 0;
    } else {
      rc = grsr_dist_pairs(in.genes, G, n, pair, 1, (int32_t *)&st, 5, first_device(), 1);
      if (rc) gpu_fail(rc);
    }

    if (verbose && !unsigned_dist) {
      fprintf(outfile, "Number of Breakpoints:         \t%d\n", st.b);
      fprintf(outfile, "Number of Cycles:              \t%d\n", st.c);
      fprintf(outfile, "Number of Hurdles:             \t%d\n", st.h);
      fprintf(outfile, "Fortress:                      \t%d\n", st.f);
    }
    if (fn_distance) fprintf(outfile, "Reversal Distance:            \t%d\n", st.d);

    if (fn_scenario && scenario_type == 1) {
      legacy_scenario(outfile, g1, g2, n, st.d);
    } else if (fn_scenario) {
      int max_steps = n + 1, ns = 0, k, j;
      grsr_step_t *steps = (grsr_step_t *)xmalloc(sizeof(grsr_step_t) * (size_t)max_steps, "steps");
      int32_t *cur = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)n, "current_genome");
      fprintf(outfile, "\n======================================================================\n\n");
      fprintf(outfile, "An optimal sequence of reversals:\n");
      fprintf(outfile, "Step 0: (Source)\n");
      print_row(g1, n);
      rc = grsr_scenario_batch(g1, g2, 1, n, steps, &ns, max_steps, first_device(), 1);
      if (rc) {
        if (rc == GRSR_ERR_SEARCH) fprintf(outfile, "ERROR: reversals ended prematurely\n");
        gpu_fail(rc);
      }
      memcpy(cur, g1, sizeof(int32_t) * (size_t)n);
      for (k = 0; k < ns; k++) { /* opt_printstep, G/opt_scenario.c:1001-1060 */
        int s = steps[k].start, e = steps[k].end;
        int startvalue = cur[s], endvalue = cur[e];
        for (i = s, j = e; i <= j; i++, j--) { int32_t t = cur[i]; cur[i] = -cur[j]; cur[j] = -t; }
        fprintf(outfile, "Step %d: ", k + 1);
        fprintf(outfile, "Chrom. %d, gene %d [%d] through chrom. %d, gene %d [%d]: ", 0, s, startvalue, 0, e, endvalue);
        fprintf(outfile, "%s", "Reversal");
        if (ns - k <= 1) fprintf(outfile, " (Destination)");
        fprintf(outfile, "\n");
        print_row(cur, n);
      }
      free(steps); free(cur);
    }
  }

  if (!fn_matrix && fn_revstats == 1) { /* print_rev_stats_t1, G/testrev.c:86-230 */
    const int32_t *g1 = in.genes + (size_t)gindex1 * (size_t)n;
    const int32_t *g2 = in.genes + (size_t)gindex2 * (size_t)n;
    /* SUCCESSOR table of the destination (G/scenario.c:137-143): succ[g + n] for g in -n..n */
    int32_t *succ = (int32_t *)xmalloc(sizeof(int32_t) * (2 * (size_t)n + 1), "successors");
    char *is_start = (char *)xmalloc((size_t)n, "starts"), *is_end = (char *)xmalloc((size_t)n, "ends");
    int32_t *cands, *d2, pair[2];
    grsr_stats_t st;
    long ncand = 0, c = 0;
    int num_bkpt = 0, j, rc, d;
    int num_down1 = 0, num_same = 0, num_up1 = 0, num_error = 0, num_tillfound = 0, foundany = 0;
    fprintf(outfile, "\n======================================================================\n\n");
    fprintf(outfile, "Test all single step reversals:\n\n");
    for (i = 0; i < 2 * n + 1; i++) succ[i] = 0;
    for (i = 0; i + 1 < n; i++) { succ[g2[i] + n] = g2[i + 1]; succ[-g2[i + 1] + n] = -g2[i]; }
    for (i = 0; i < n; i++) {
      is_start[i] = (char)(i == 0 || succ[g1[i - 1] + n] != g1[i]);
      is_end[i] = (char)(i == n - 1 || succ[g1[i] + n] != g1[i + 1]);
      num_bkpt += is_start[i];
    }
    num_bkpt++; /* the end counts as a breakpoint for unichromosomal genomes (:127) */
    fprintf(outfile, "Breakpoint positions (%d):", num_bkpt);
    for (i = 0; i < n; i++) if (is_start[i]) fprintf(outfile, " %d", i);
    fprintf(outfile, " %d", n);
    fprintf(outfile, "\n\n");
    for (i = 0; i < n; i++) if (is_start[i]) for (j = i; j < n; j++) if (is_end[j]) ncand++;
    cands = (int32_t *)xmalloc(sizeof(int32_t) * 2 * (size_t)ncand, "candidates");
    d2 = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)ncand, "trial distances");
    for (i = 0; i < n; i++) if (is_start[i]) for (j = i; j < n; j++) if (is_end[j]) {
      cands[2 * c] = i; cands[2 * c + 1] = j; c++;
    }
    pair[0] = gindex1; pair[1] = gindex2;
    rc = grsr_dist_pairs(in.genes, G, n, pair, 1, (int32_t *)&st, 5, first_device(), 1);
    if (rc) gpu_fail(rc);
    d = st.d;
    rc = grsr_trial_distances(g1, g2, n, cands, ncand, d2, first_device());
    if (rc) gpu_fail(rc);
    c = 0;
    for (i = 0; i < n; i++) {
      int foundany_i = 0;
      if (!is_start[i]) continue;
      for (j = i; j < n; j++) {
        if (!is_end[j]) continue;
        if (!foundany) num_tillfound++;
        if (d2[c] == d - 1) {
          num_down1++;
          if (!foundany_i) { fprintf(outfile, "%d: ", i); foundany = foundany_i = 1; }
          fprintf(outfile, " %d", j);
        } else if (d2[c] == d) num_same++;
        else if (d2[c] == d + 1) num_up1++;
        else {
          num_error++;
          fprintf(outfile, "\nERROR: Potential distance caculation error;\n");
          fprintf(outfile, " Reversal in positions %d..%d changed distance from %d to %d\n", i, j, d, d2[c]);
        }
        c++;
      }
      if (foundany_i) fprintf(outfile, "\n");
    }
    fprintf(outfile, "\nNumber of reversals changing distance by\n   -1 is %d;  0 is %d;  +1 is %d;  other is %d;  total: %d\n",
            num_down1, num_same, num_up1, num_error, num_down1 + num_same + num_up1 + num_error);
    fprintf(outfile, "First success after %d trials\n", num_tillfound);
    free(succ); free(is_start); free(is_end); free(cands); free(d2);
  }

  for (i = 0; i < G; i++) free(in.names[i]);
  free(in.names);
  free(in.genes);
  if (outfile != stdout) fclose(outfile);
  return 0;
}

#undef exit

/* ---- resident mode ------------------------------------------------------------------------------------- */

static int send_all(int fd, const void *buf, size_t len)
{
  const char *p = (const char *)buf;
  while (len) { ssize_t k = write(fd, p, len); if (k <= 0) { if (errno == EINTR) continue; return -1; } p += k; len -= (size_t)k; }
  return 0;
}
static int recv_all(int fd, void *buf, size_t len)
{
  char *p = (char *)buf;
  while (len) { ssize_t k = read(fd, p, len); if (k <= 0) { if (k < 0 && errno == EINTR) continue; return -1; } p += k; len -= (size_t)k; }
  return 0;
}

/* client: hand this invocation to the resident process; -1 = nobody there (run locally) */
static int run_remote(const char *path, int argc, char **argv)
{
  struct sockaddr_un addr;
  struct msghdr msg;
  struct iovec iov;
  char cbuf[CMSG_SPACE(2 * sizeof(int))], cwd[4096];
  struct cmsghdr *cm;
  int fd, i, hdr[2], fds[2] = {1, 2}, status = 255;
  if (strlen(path) >= sizeof addr.sun_path) return -1;
  fd = socket(AF_UNIX, SOCK_STREAM, 0);
  if (fd < 0) return -1;
  memset(&addr, 0, sizeof addr);
  addr.sun_family = AF_UNIX;
  strcpy(addr.sun_path, path);
  if (connect(fd, (struct sockaddr *)&addr, sizeof addr) != 0) { close(fd); return -1; }
  if (!getcwd(cwd, sizeof cwd)) strcpy(cwd, "/");
  fflush(stdout); fflush(stderr);
  hdr[0] = argc; hdr[1] = (int)strlen(cwd);
  memset(&msg, 0, sizeof msg);
  iov.iov_base = hdr; iov.iov_len = sizeof hdr;
  msg.msg_iov = &iov; msg.msg_iovlen = 1;
  msg.msg_control = cbuf; msg.msg_controllen = sizeof cbuf;
  cm = CMSG_FIRSTHDR(&msg);
  cm->cmsg_level = SOL_SOCKET; cm->cmsg_type = SCM_RIGHTS; cm->cmsg_len = CMSG_LEN(2 * sizeof(int));
  memcpy(CMSG_DATA(cm), fds, sizeof fds);                       /* our stdout and stderr travel with the request */
  if (sendmsg(fd, &msg, 0) != (ssize_t)sizeof hdr) { close(fd); return -1; }
  if (send_all(fd, cwd, (size_t)hdr[1])) { close(fd); return -1; }
  for (i = 0; i < argc; i++) {
    int len = (int)strlen(argv[i]);
    if (send_all(fd, &len, sizeof len) || send_all(fd, argv[i], (size_t)len)) { close(fd); return -1; }
  }
  if (recv_all(fd, &status, sizeof status)) status = 255;
  close(fd);
  return status;
}

/* one request on connection fd: receive argv / cwd / the client's stdout and stderr, run the report
 * with them, send the exit status back */
static void serve_one(int fd)
{
  struct msghdr msg;
  struct iovec iov;
  char cbuf[CMSG_SPACE(2 * sizeof(int))], *cwd = NULL, **argv = NULL;
  struct cmsghdr *cm;
  int hdr[2], fds[2] = {-1, -1}, i, status = 255, argc = 0, save1, save2, jr;
  memset(&msg, 0, sizeof msg);
  iov.iov_base = hdr; iov.iov_len = sizeof hdr;
  msg.msg_iov = &iov; msg.msg_iovlen = 1;
  msg.msg_control = cbuf; msg.msg_controllen = sizeof cbuf;
  if (recvmsg(fd, &msg, 0) != (ssize_t)sizeof hdr) return;
  cm = CMSG_FIRSTHDR(&msg);
  if (cm && cm->cmsg_level == SOL_SOCKET && cm->cmsg_type == SCM_RIGHTS && cm->cmsg_len == CMSG_LEN(2 * sizeof(int)))
    memcpy(fds, CMSG_DATA(cm), sizeof fds);
  if (fds[0] < 0 || hdr[0] < 1 || hdr[0] > 4096 || hdr[1] < 0 || hdr[1] > 65536) goto done;
  argc = hdr[0];
  cwd = (char *)calloc((size_t)hdr[1] + 1, 1);
  argv = (char **)calloc((size_t)argc + 1, sizeof(char *));
  if (!cwd || !argv || recv_all(fd, cwd, (size_t)hdr[1])) goto done;
  for (i = 0; i < argc; i++) {
    int len;
    if (recv_all(fd, &len, sizeof len) || len < 0 || len > 1 << 20) goto done;
    argv[i] = (char *)calloc((size_t)len + 1, 1);
    if (!argv[i] || recv_all(fd, argv[i], (size_t)len)) goto done;
  }
  fflush(stdout); fflush(stderr);
  save1 = dup(1); save2 = dup(2);
  dup2(fds[0], 1); dup2(fds[1], 2);
  if (chdir(cwd) != 0) { fprintf(stderr, "ERROR: cannot enter %s\n", cwd); status = 255; }
  else {
    optind = 0;                                       /* glibc: restart getopt from scratch */
    serving = 1;
    jr = setjmp(server_jmp);
    if (jr == 0) status = grimm_main(argc, argv);
    else status = jr - 1;
    serving = 0;
  }
  fflush(stdout); fflush(stderr);
  dup2(save1, 1); dup2(save2, 2);
  close(save1); close(save2);
done:
  if (fds[0] >= 0) close(fds[0]);
  if (fds[1] >= 0) close(fds[1]);
  send_all(fd, &status, sizeof status);
  if (argv) { for (i = 0; i < argc; i++) free(argv[i]); free(argv); }
  free(cwd);
}

static int run_server(const char *path)
{
  struct sockaddr_un addr;
  int ls, one[6] = {1, 2, 3, 1, -2, 3}, pair[2] = {0, 1};
  grsr_stats_t st;
  if (strlen(path) >= sizeof addr.sun_path) { fprintf(stderr, "ERROR: socket path too long\n"); return 255; }
  /* create the CUDA context and load the kernels before the first request arrives */
  if (grsr_dist_pairs(one, 2, 3, pair, 1, (int32_t *)&st, 5, first_device(), 1) != GRSR_OK) {
    fprintf(stderr, "ERROR: %s\n", grsr_last_error());
    return 255;
  }
  signal(SIGPIPE, SIG_IGN);
  ls = socket(AF_UNIX, SOCK_STREAM, 0);
  memset(&addr, 0, sizeof addr);
  addr.sun_family = AF_UNIX;
  strcpy(addr.sun_path, path);
  unlink(path);
  if (ls < 0 || bind(ls, (struct sockaddr *)&addr, sizeof addr) != 0 || listen(ls, 64) != 0) {
    fprintf(stderr, "ERROR: cannot listen on %s: ", path);
    perror("");
    return 255;
  }
  fprintf(stderr, "grimm: resident on %s (GRSR_SERVER=%s grimm ... to use it)\n", path, path);
  for (;;) {
    int fd = accept(ls, NULL, NULL);
    if (fd < 0) { if (errno == EINTR) continue; break; }
    serve_one(fd);
    close(fd);
  }
  return 0;
}

int main(int argc, char *argv[])
{
  const char *srv = getenv("GRSR_SERVER");
  /* The CUDA driver initialises every visible GPU: 5.5 s on an 8-GPU B200 box against 0.5-1 s with one
   * (profiles/r2_startup.txt).  Unless the caller asked for several devices or set the variable, only
   * the device this process is going to use stays visible; it is then device 0. */
  if (!getenv("CUDA_VISIBLE_DEVICES") && num_gpus() == 1) {
    char one[16];
    snprintf(one, sizeof one, "%d", first_device());
    setenv("CUDA_VISIBLE_DEVICES", one, 1);
    setenv("GRSR_DEVICE", "0", 1);
  }
  if (argc == 3 && strcmp(argv[1], "--server") == 0) return run_server(argv[2]);
  if (srv && *srv) {
    int status = run_remote(srv, argc, argv);
    if (status >= 0) return status;
  }
  return grimm_main(argc, argv);
}
