/* loaders.c - plain-text input loaders (see include/libecp_b200_io.h).
 *
 * Replaces loadBS / loadGeometry / loadECP of the reference's example program (example/ex1.c:11-123).  The files are
 * read into memory and consumed field by field with strtol / strtod: an integer field ends where the digits end, like
 * scanf's %d, which is what makes the INDEXED reading of the shipped ECP file reproduce the reference's values
 * ("-2.6739" -> index -2, then ".6739").
 */
#include "../../include/libecp_b200_io.h"

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  char *buf;
  size_t pos, len;
} Text;

static int text_open(const char *path, Text *t) {
  FILE *fp = fopen(path, "rb");
  if (!fp) return -1;
  fseek(fp, 0, SEEK_END);
  long n = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  t->buf = malloc((size_t)(n > 0 ? n : 0) + 1);
  t->len = fread(t->buf, 1, (size_t)(n > 0 ? n : 0), fp);
  t->buf[t->len] = 0;
  t->pos = 0;
  fclose(fp);
  return 0;
}
static void skip_space(Text *t) {
  while (t->pos < t->len && isspace((unsigned char)t->buf[t->pos])) t->pos++;
}
static int get_int(Text *t, int *v) {
  skip_space(t);
  char *end;
  const long x = strtol(t->buf + t->pos, &end, 10);
  if (end == t->buf + t->pos) return -2;
  t->pos = (size_t)(end - t->buf);
  *v = (int)x;
  return 0;
}
static int get_double(Text *t, double *v) {
  skip_space(t);
  char *end;
  const double x = strtod(t->buf + t->pos, &end);
  if (end == t->buf + t->pos) return -2;
  t->pos = (size_t)(end - t->buf);
  *v = x;
  return 0;
}
static void skip_line(Text *t) {
  while (t->pos < t->len && t->buf[t->pos] != '\n') t->pos++;
  if (t->pos < t->len) t->pos++;
}
static void skip_word(Text *t) {
  skip_space(t);
  while (t->pos < t->len && !isspace((unsigned char)t->buf[t->pos])) t->pos++;
}

void libecp_io_free(void *p) { free(p); }

int libecp_io_ao_dim(int nrShells, const int *lBS) {
  int dim = 0;
  for (int s = 0; s < nrShells; s++) dim += (lBS[s] + 1) * (lBS[s] + 2) / 2;
  return dim;
}

int libecp_io_load_xyz(const char *path, int *nrAtoms, double **geometry) {
  Text t;
  *geometry = NULL;
  *nrAtoms = 0;
  if (text_open(path, &t)) return -1;
  int n = 0, rc = get_int(&t, &n);
  if (rc || n < 0) {
    free(t.buf);
    return -2;
  }
  skip_line(&t); /* rest of the count line */
  skip_line(&t); /* comment line */
  double *g = malloc((size_t)(3 * n + 1) * sizeof(double));
  for (int i = 0; i < n && !rc; i++) {
    skip_word(&t); /* element symbol */
    for (int k = 0; k < 3 && !rc; k++) rc = get_double(&t, &g[3 * i + k]);
  }
  free(t.buf);
  if (rc) {
    free(g);
    return -2;
  }
  *nrAtoms = n;
  *geometry = g;
  return 0;
}

/* shared walk over "<header ints> ; per shell <l> <K> ; per primitive <fields>": pass 0 counts, pass 1 fills */
typedef struct {
  int *shells, *l, *K;
  double *f[3]; /* up to three real fields per primitive, in the order they are stored */
} Blocks;

static int walk(Text *t, int nrAtoms, int headerInts, int indexed, int nreal, const int *order, Blocks *b, int *nsh_,
                int *nprim_) {
  int nsh = 0, nprim = 0;
  t->pos = 0;
  for (int i = 0; i < nrAtoms; i++) {
    int hdr[3] = {0, 0, 0};
    for (int k = 0; k < headerInts; k++)
      if (get_int(t, &hdr[k])) return -2;
    const int ns = hdr[headerInts - 1];
    if (ns < 0) return -2;
    if (b) b->shells[i] = ns;
    for (int s = 0; s < ns; s++, nsh++) {
      int l, K;
      if (get_int(t, &l) || get_int(t, &K) || K < 0 || l < 0) return -2;
      if (b) {
        b->l[nsh] = l;
        b->K[nsh] = K;
      }
      for (int p = 0; p < K; p++, nprim++) {
        int idx;
        if (indexed && get_int(t, &idx)) return -2;
        for (int k = 0; k < nreal; k++) {
          double x;
          if (get_double(t, &x)) return -2;
          if (b) b->f[order[k]][nprim] = x;
        }
      }
    }
  }
  *nsh_ = nsh;
  *nprim_ = nprim;
  return 0;
}

int libecp_io_load_bs(const char *path, int nrAtoms, int **shellsBS, int **lBS, int **KBS, double **aBS, double **dBS,
                      int *nrShells) {
  Text t;
  *shellsBS = *lBS = *KBS = NULL;
  *aBS = *dBS = NULL;
  if (text_open(path, &t)) return -1;
  static const int order[2] = {0, 1}; /* <index> <exponent> <coefficient> */
  int nsh = 0, nprim = 0;
  int rc = walk(&t, nrAtoms, 2, 1, 2, order, NULL, &nsh, &nprim);
  if (!rc) {
    Blocks b;
    b.shells = calloc((size_t)nrAtoms + 1, sizeof(int));
    b.l = calloc((size_t)nsh + 1, sizeof(int));
    b.K = calloc((size_t)nsh + 1, sizeof(int));
    b.f[0] = calloc((size_t)nprim + 1, sizeof(double));
    b.f[1] = calloc((size_t)nprim + 1, sizeof(double));
    b.f[2] = NULL;
    rc = walk(&t, nrAtoms, 2, 1, 2, order, &b, &nsh, &nprim);
    *shellsBS = b.shells;
    *lBS = b.l;
    *KBS = b.K;
    *aBS = b.f[0];
    *dBS = b.f[1];
    if (nrShells) *nrShells = nsh;
  }
  free(t.buf);
  return rc;
}

int libecp_io_load_ecp(const char *path, int nrAtoms, int format, int **shellsECP, int **lECP, int **KECP, double **aECP,
                       double **dECP, double **nECP) {
  Text t;
  *shellsECP = *lECP = *KECP = NULL;
  *aECP = *dECP = *nECP = NULL;
  if (format != LIBECP_IO_ECP_INDEXED && format != LIBECP_IO_ECP_SHIPPED) return -2;
  if (text_open(path, &t)) return -1;
  /* storage order of the real fields: f[0] = a (exponent), f[1] = d (coefficient), f[2] = n (power) */
  static const int orderIndexed[3] = {0, 1, 2}; /* <index> <a> <d> <n> */
  static const int orderShipped[3] = {1, 0, 2}; /* <d> <a> <n>         */
  const int indexed = format == LIBECP_IO_ECP_INDEXED;
  const int *order = indexed ? orderIndexed : orderShipped;
  int nsh = 0, nprim = 0;
  int rc = walk(&t, nrAtoms, 3, indexed, 3, order, NULL, &nsh, &nprim);
  if (!rc) {
    Blocks b;
    b.shells = calloc((size_t)nrAtoms + 1, sizeof(int));
    b.l = calloc((size_t)nsh + 1, sizeof(int));
    b.K = calloc((size_t)nsh + 1, sizeof(int));
    for (int k = 0; k < 3; k++) b.f[k] = calloc((size_t)nprim + 1, sizeof(double));
    rc = walk(&t, nrAtoms, 3, indexed, 3, order, &b, &nsh, &nprim);
    *shellsECP = b.shells;
    *lECP = b.l;
    *KECP = b.K;
    *aECP = b.f[0];
    *dECP = b.f[1];
    *nECP = b.f[2];
  }
  free(t.buf);
  return rc;
}
