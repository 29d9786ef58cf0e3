/* dele_transp.c -- host-side restatement of GRSR's scr/deleTransp.java:
 * the step that runs immediately before `grimm` for every genome pair (reference scr/reptA.sh:58,
 * scr/deleTransp.sh:13).  It spawns one JVM per genome pair in the reference; once the distance /
 * scenario call takes milliseconds that JVM is what the pair loop waits for.
 *
 *     deleTransp <n> <g0_1 .. g0_n> $ <g1_1 .. g1_n> $        (the argv of `java deleTransp`)
 *
 * Same observable behaviour as the Java program:
 *   files   BefMerge.txt AftMerge.txt BefFilter.txt AftFilter.txt (block-id maps read back by
 *           scr/getleftBlk.sh, getrightBlk.sh, getTBIBlk.sh) and mgr_macro2.txt (grimm's input),
 *           all in the current directory
 *   stdout  the transposition / block-interchange summary reptA.sh stores in tbi.txt
 * Plain host C on purpose: relabelling and list surgery on a few dozen blocks per pair; there is
 * nothing here for a GPU.  Every function cites the Java lines it restates; the Java program's
 * index quirks (ArrayList.remove(int) vs remove(Object), "id > 0" meaning found, stale positions
 * after a wrap-around removal) are kept, because parity with its output is the contract.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int *v; int n, cap; } list_t;

static void die(const char *msg) { fprintf(stderr, "deleTransp: %s\n", msg); exit(1); }

static void l_init(list_t *l) { l->v = NULL; l->n = 0; l->cap = 0; }
static void l_add(list_t *l, int x)
{
  if (l->n == l->cap) {
    l->cap = l->cap ? 2 * l->cap : 16;
    l->v = (int *)realloc(l->v, sizeof(int) * (size_t)l->cap);
    if (!l->v) die("out of memory");
  }
  l->v[l->n++] = x;
}
static int l_index_of(const list_t *l, int x) { int i; for (i = 0; i < l->n; i++) if (l->v[i] == x) return i; return -1; }
static int l_get(const list_t *l, int i)
{
  if (i < 0 || i >= l->n) die("index out of bounds (java.lang.IndexOutOfBoundsException in the reference)");
  return l->v[i];
}
/* ArrayList.remove(int index) */
static void l_remove_at(list_t *l, int i)
{
  if (i < 0 || i >= l->n) die("index out of bounds (java.lang.IndexOutOfBoundsException in the reference)");
  memmove(l->v + i, l->v + i + 1, sizeof(int) * (size_t)(l->n - i - 1));
  l->n--;
}
/* ArrayList.remove(Object o): first occurrence, nothing if absent */
static void l_remove_value(list_t *l, int x) { int i = l_index_of(l, x); if (i >= 0) l_remove_at(l, i); }
static int iabs(int x) { return x < 0 ? -x : x; }

/* A scaffold as the Java program passes it around: block tokens as STRINGS (inOrder compares text). */
typedef struct { char **tok; int n; } toks_t;

static toks_t toks_from_list(const list_t *l)
{
  toks_t t; int i; char buf[32];
  t.n = l->n; t.tok = (char **)malloc(sizeof(char *) * (size_t)(l->n > 0 ? l->n : 1));
  for (i = 0; i < l->n; i++) { snprintf(buf, sizeof buf, "%d", l->v[i]); t.tok[i] = strdup(buf); }
  return t;
}
static void toks_free(toks_t *t) { int i; for (i = 0; i < t->n; i++) free(t->tok[i]); free(t->tok); }
static int toks_index_of(const toks_t *t, const char *s) { int i; for (i = 0; i < t->n; i++) if (strcmp(t->tok[i], s) == 0) return i; return -1; }

/* inOrder (deleTransp.java:51-92): source becomes 1..n, the destination is rewritten in the source's
 * numbering; <filename>.txt receives "old=new" per source block. */
static void in_order(const toks_t *s, const toks_t *d, const char *filename, list_t *o0, list_t *o1)
{
  char path[64], buf[64];
  FILE *f;
  int h;
  snprintf(path, sizeof path, "%s.txt", filename);
  f = fopen(path, "w");
  if (!f) die("cannot write a block map file");
  o0->n = 0; o1->n = 0;
  for (h = 0; h < s->n; h++) {
    const char *old = d->tok[h];
    int id, nb;
    l_add(o0, h + 1);
    fprintf(f, "%s=%d\n", s->tok[h], h + 1);
    id = toks_index_of(s, old);                                   /* :73 */
    if (id >= 0) nb = id + 1;
    else {
      snprintf(buf, sizeof buf, "-%s", old);
      id = toks_index_of(s, buf);                                 /* :77  d = 3, s = -3 */
      if (id >= 0) nb = -(id + 1);
      else {                                                      /* :81  d = -3, s = 3 */
        size_t k, m = 0;
        for (k = 0; old[k]; k++) if (old[k] != '-') buf[m++] = old[k];
        buf[m] = 0;
        id = toks_index_of(s, buf);
        nb = -(id + 1);
      }
    }
    l_add(o1, nb);
  }
  fclose(f);
}

/* merge2Blks (deleTransp.java:94-228): both scaffolds rotated (or reverse-complemented) to start with
 * block 1, then every block that follows its predecessor in BOTH scaffolds is merged into it. */
static void merge2blks(const list_t pair[2], const char *filename, list_t g[2])
{
  const int blk = pair[0].n;
  int j, k, c1;
  char path[64];
  FILE *f;
  int cursor = 0;                       /* the Iterator over g[0] */
  for (j = 0; j < 2; j++) {
    int one = l_index_of(&pair[j], 1);
    g[j].n = 0;
    if (one == 0) {
      for (k = 0; k < blk; k++) l_add(&g[j], pair[j].v[k]);
    } else if (one > 0) {
      c1 = 0;
      for (k = 0; k < blk; k++) {
        if (one + k < blk) l_add(&g[j], pair[j].v[one + k]);
        else { l_add(&g[j], pair[j].v[c1]); c1++; }
      }
    } else {                            /* -1 present: reverse complement, starting from it (:120-131) */
      int mone = l_index_of(&pair[j], -1);
      c1 = 0;
      for (k = 0; k < blk; k++) {
        if (mone - k >= 0) l_add(&g[j], -l_get(&pair[j], mone - k));
        else { l_add(&g[j], -l_get(&pair[j], blk - 1 - c1)); c1++; }
      }
    }
  }
  while (cursor < g[0].n) {             /* :141 */
    int p[2], count[2] = {0, 0}, all_eq = 1, next0, j1;
    const int cblk = g[0].v[cursor++];  /* it.next() */
    for (j1 = 0; j1 < 2; j1++) {
      p[j1] = l_index_of(&g[j1], cblk);
      if (p[j1] < 0) p[j1] = -l_index_of(&g[j1], -cblk);
    }
    if (cursor < g[0].n) next0 = l_get(&g[0], p[0] + 1);
    else break;
    if (p[1] >= 0 && p[1] < g[1].n - 1 && l_get(&g[1], p[1] + 1) != next0) all_eq = 0;
    else if (p[1] == g[1].n - 1 && l_get(&g[1], 0) != next0) all_eq = 0;
    else if (p[1] < 0 && l_get(&g[1], -p[1] - 1) != -next0) all_eq = 0;
    while (all_eq) {
      /* it.next(); it.remove(): the block after cblk leaves g[0] */
      if (cursor >= g[0].n) die("iterator exhausted (java.util.NoSuchElementException in the reference)");
      l_remove_at(&g[0], cursor);
      if (p[1] >= 0) {
        if (p[1] < g[1].n - 1) l_remove_at(&g[1], p[1] + 1);
        else if (p[1] == g[1].n - 1) l_remove_at(&g[1], 0);
      }
      if (p[1] < 0) { l_remove_at(&g[1], -p[1] - count[1] - 1); count[1]++; }
      if (cursor >= g[0].n) break;
      next0 = l_get(&g[0], p[0] + 1);
      if (p[1] >= 0 && p[1] < g[1].n - 1 && l_get(&g[1], p[1] + 1) != next0) all_eq = 0;
      else if (p[1] == g[1].n - 1 && l_get(&g[1], 0) != next0) all_eq = 0;
      else if (p[1] < 0 && l_get(&g[1], -p[1] - count[1] - 1) != -next0) all_eq = 0;
    }
  }
  snprintf(path, sizeof path, "%s.txt", filename);
  f = fopen(path, "w");
  if (!f) die("cannot write a block map file");
  for (k = 0; k < g[0].n; k++) {        /* "first:last=first" of every merged block (:213-224) */
    int end = (k != g[0].n - 1) ? g[0].v[k + 1] - 1 : l_get(&pair[0], blk - 1);
    fprintf(f, "%d:%d=%d\n", g[0].v[k], end, g[0].v[k]);
  }
  fclose(f);
}

static void print_named(const char *name, const list_t *l)
{
  int i;
  if (l->n == 0) return;
  printf("%s", name);
  for (i = 0; i < l->n; i++) printf(" %d", l->v[i]);
  printf("\n");
}

/* filterTBI (deleTransp.java:230-382): blocks that moved by a transposition or a block interchange
 * are taken out of both scaffolds and reported on stdout. */
static void filter_tbi(list_t gp[2])
{
  const int blk = gp[0].n;
  list_t L[4], tp, itp, bi, ibi, hibi;
  int i, c, tno, k2;
  for (c = 0; c < 4; c++) l_init(&L[c]);            /* 0,2 deletion side; 1,3 insertion side */
  l_init(&tp); l_init(&itp); l_init(&bi); l_init(&ibi); l_init(&hibi);
  for (i = 0; i <= blk - 2; i++)
    if (gp[1].v[i] == gp[1].v[i + 1] - 2) l_add(&L[0], gp[1].v[i] + 1);
  for (i = 0; i <= blk - 3; i++) {
    if (gp[1].v[i] == gp[1].v[i + 2] - 1) l_add(&L[1], gp[1].v[i + 1]);
    if (gp[1].v[i + 1] != gp[1].v[i] + 1 && gp[1].v[i + 2] == gp[1].v[i] + 2) {
      l_add(&L[2], gp[1].v[i] + 1);
      l_add(&L[3], gp[1].v[i + 1]);
    }
  }
  if (blk >= 2 && gp[1].v[0] == gp[0].v[0]) {        /* the last block (:253-268) */
    if (gp[1].v[blk - 1] == gp[0].v[blk - 2]) l_add(&L[0], gp[0].v[blk - 1]);
    if (gp[1].v[blk - 2] == gp[0].v[blk - 1]) l_add(&L[1], gp[1].v[blk - 1]);
    if (gp[1].v[blk - 1] != gp[0].v[blk - 1] && gp[1].v[blk - 2] == gp[0].v[blk - 2]) {
      l_add(&L[2], gp[0].v[blk - 1]);
      l_add(&L[3], gp[1].v[blk - 1]);
    }
  }
  tno = L[1].n;
  for (i = 0; i < tno; i++) {                        /* transpositions (:278-289) */
    const int x = L[1].v[i];
    if (l_index_of(&L[0], x) >= 0) {
      l_add(&tp, iabs(x)); l_remove_value(&gp[1], x); l_remove_value(&gp[0], iabs(x));
    } else if (l_index_of(&L[0], -x) >= 0) {
      l_add(&itp, iabs(x)); l_remove_value(&gp[1], x); l_remove_value(&gp[0], iabs(x));
    }
  }
  /* block interchanges: two iterators in step over L[2] / L[3] (:292-348); k2 = both cursors */
  k2 = 0;
  while (k2 < L[2].n && k2 < L[3].n) {
    const int c2 = L[2].v[k2], c3 = L[3].v[k2];
    int inv = 0, id;
    k2++;
    if (iabs(c2) == iabs(c3)) { k2--; l_remove_at(&L[2], k2); l_remove_at(&L[3], k2); continue; }
    id = l_index_of(&L[2], c3);
    if (id < 0) { id = l_index_of(&L[2], -c3); if (id > 0) inv = 1; }
    if (id > 0) {                                    /* "found" means index > 0 in the reference */
      const int partner = l_get(&L[3], id);
      if (partner == c2 || partner == -c2) {
        list_t *dst = (partner == c2) ? (inv ? &hibi : &bi) : (inv ? &ibi : &hibi);
        l_add(dst, iabs(partner)); l_add(dst, iabs(c3));
        l_remove_value(&gp[1], partner); l_remove_value(&gp[1], c3);
        l_remove_value(&gp[0], iabs(partner)); l_remove_value(&gp[0], iabs(c3));
        k2--; l_remove_at(&L[2], k2); l_remove_at(&L[3], k2);
      }
    }
  }
  if (tp.n > 1) {                                    /* adjacent transposed blocks count once (:351-364) */
    int prev = tp.v[0];
    i = 1;
    while (i < tp.n) {
      if (iabs(tp.v[i] - prev) == 1) l_remove_at(&tp, i);
      else { prev = tp.v[i]; i++; }
    }
  }
  printf("%d\n", tp.n + itp.n + bi.n / 2 + ibi.n / 2 + hibi.n / 2);
  printf("%d %d %d %d %d\n", tp.n, itp.n, bi.n / 2, ibi.n / 2, hibi.n / 2);
  print_named("T", &tp); print_named("IT", &itp); print_named("B", &bi);
  print_named("IB", &ibi); print_named("HIB", &hibi);
}

static void write_row(FILE *f, const list_t *l)
{
  int i;
  if (l->n == 0) fprintf(f, " ");                    /* "" + " $" in the reference */
  for (i = 0; i < l->n; i++) fprintf(f, "%d ", l->v[i]);
  fprintf(f, "$\n");
}

int main(int argc, char **argv)
{
  int n, i;
  toks_t s, d;
  list_t a[2], m[2], b[2];
  FILE *f;
  if (argc < 2) die("usage: deleTransp <n> <g0 blocks> $ <g1 blocks> $");
  n = atoi(argv[1]);
  if (n < 1 || argc < 2 * n + 3) die("not enough blocks on the command line");
  argv++;                               /* argv[k] is now the Java program's args[k] */
  s.n = d.n = n;
  s.tok = (char **)malloc(sizeof(char *) * (size_t)n);
  d.tok = (char **)malloc(sizeof(char *) * (size_t)n);
  for (i = 1; i <= n; i++) { s.tok[i - 1] = strdup(argv[i]); d.tok[i - 1] = strdup(argv[i + n + 1]); }   /* :20-23 */
  for (i = 0; i < 2; i++) { l_init(&a[i]); l_init(&m[i]); l_init(&b[i]); }

  in_order(&s, &d, "BefMerge", &a[0], &a[1]);                  /* :27 */
  merge2blks(a, "AftMerge", m);                                /* :28 */
  toks_free(&s); toks_free(&d);
  s = toks_from_list(&m[0]); d = toks_from_list(&m[1]);
  in_order(&s, &d, "BefFilter", &b[0], &b[1]);                 /* :32 */
  filter_tbi(b);                                               /* :34 */
  toks_free(&s); toks_free(&d);
  s = toks_from_list(&b[0]); d = toks_from_list(&b[1]);
  in_order(&s, &d, "AftFilter", &a[0], &a[1]);                 /* :38 */
  f = fopen("mgr_macro2.txt", "w");                            /* :41-48 */
  if (!f) die("cannot write mgr_macro2.txt");
  fprintf(f, ">genome1\n"); write_row(f, &a[0]);
  fprintf(f, ">genome2\n"); write_row(f, &a[1]);
  fclose(f);
  return 0;
}
