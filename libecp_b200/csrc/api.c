/* api.c - the drop-in C API (include/libecp.h, getIntegrals.h, dimensions.h) on top of the host
 * builder and the CUDA layer.  Mirrors the reference's entry points:
 *   libECP_init            src/libecp.c:53-201
 *   calculateECPIntegrals  src/libecp.c:212-404   (callbacks replayed on the host in the reference's order)
 *   libECP_free            src/libecp.c:407-444
 *   getIntegrals           src/getIntegrals.c:45-95
 *   cartesianShellOrder[Index]  src/dimensions.c:17-57
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/dimensions.h"
#include "../../include/getIntegrals.h"
#include "../../include/libecp_b200.h"
#include "builder.h"
#include "tables.h"

struct _libECPHandle {
  EcpTables *tab;
  EcpDev *dev;
  struct BuildWorker *worker;
  EcpBatchBuf *bb, *bb2; /* two batch buffers: batch i+1 is built on a host thread while the GPU works on batch i */
  const double *geometry;
  int nrAtoms, empty;
  int rank, world;
  long long maxTriples;
  double *hostBlocks;
  size_t hostBlocksCap;
  struct StreamState *stream; /* host consumer: rows that are final are downloaded while the pass goes on (below) */
  double triPerPair; /* executed triples per tested shell pair, measured on the batches run so far (device enumeration) */
  /* derivative runs: the expanded shell list the tables borrow (libECP_init) */
  int *xShells, *xL, *xK;
  double *xD, *xA;
  libecp_b200_stats_t stats;
};

struct BuildWorker;
static void worker_free(struct BuildWorker *w);
static int g_host_threads = 0;
static int g_device = -1;
static int g_tables_only = 0;
static char g_apierr[256] = "";

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void free_expanded(libECPHandle *h) {
  free(h->xShells); free(h->xL); free(h->xK); free(h->xD); free(h->xA);
}

void libecp_b200_set_device(int device) { g_device = device; }
/* host threads of the batch builder and of the host-side += (OpenMP).  Launchers such as torchrun export
 * OMP_NUM_THREADS=1; a multi-GPU caller gives each rank its share of the cores explicitly. */
void libecp_b200_set_host_threads(int n) {
  g_host_threads = n;
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
void libecp_b200_set_tables_only(int on) { g_tables_only = on; }
void libecp_b200_set_serial_kernels(libECPHandle *h, int on) {
  if (h && h->dev) ecpdev_set_serial(h->dev, on);
}
const char *libecp_b200_last_error(void) { return g_apierr[0] ? g_apierr : ecpdev_last_error(); }

double libecp_b200_fp64_peak(int device, int iters) { return ecpdev_fp64_peak_probe(device, iters); }

libECPHandle *libECP_init(int nrAtoms, double *geometry, int *shellsECP, int *lECP, int *KECP, double *nECP,
                          double *dECP, double *aECP, int *shellsBS, int *lBS, int *KBS, double *dBS, double *aBS,
                          int n, int lmax, int *shellOrdering, int largeGridOrder, double tolerance, double accuracy) {
  g_apierr[0] = 0; /* lmax is only read together with shellOrdering (reference src/libecp.c:152-166) */
  if (n < 0 || n > 2) {
    snprintf(g_apierr, sizeof(g_apierr), "libecp_b200: derivative order n=%d not supported (n = 0, 1 or 2)", n);
    return NULL;
  }
  libECPHandle *h = calloc(1, sizeof(*h));
  if (!h) {
    snprintf(g_apierr, sizeof(g_apierr), "libecp_b200: out of memory");
    return NULL;
  }
  h->nrAtoms = nrAtoms;
  h->geometry = geometry;
  h->world = 1;
  h->maxTriples = 3000000; /* ~0.9 ms of fixed device time per batch (57 launches, joins): fewer, larger batches */
  {
    const char *e = getenv("LIBECP_B200_BATCH_TRIPLES");
    if (e && atoll(e) > 0) h->maxTriples = atoll(e);
  }
  EcpBuildOpts opts = {shellOrdering, lmax, NULL, 0, NULL, NULL};
  if (n >= 1) {
    /* Derivatives (scope row f1; reference src/libecp.c:203-210,246-250,322-330): a derivative block is an ordinary
     * block between shells shifted in angular momentum - l + k with the coefficients d zeta^k (src/type1.c:239-246,
     * src/type2.c:263-269,459-462) or l - k with d - screened as the unshifted shell (src/type2.c:251).  The handle
     * therefore runs on an expanded shell list: every shell is followed by its shifted copies (ecp_deriv_copy: n = 1:
     * l+1, l-1; n = 2: l+1, l+2, l with d zeta, l-1, l-2; copies below l = 0 are left out, they come last); the builder
     * pairs them as the reference's shift table prescribes.  All sizes (tables, Bessel depth, classes) follow from the
     * expanded list exactly as the reference derives them from maxLBS + n.
     * n = 2, shifts (+1,0) and (0,+1) (src/libecp.c:362-369): the terms of a second derivative that raise one Cartesian
     * exponent and lower another - momentum l, coefficients d zeta.  The reference evaluates chi / gamma at l + 1 and
     * shifts only their lower-degree part (the arrays are cumulative over the momenta); here the copy "l with d zeta"
     * gives the same block directly. */
    const int ncopy = ecp_deriv_ncopies(n);
    int nsh = 0, nprim = 0;
    for (int i = 0; i < nrAtoms; i++)
      for (int j = 0; j < shellsBS[i]; j++) nprim += KBS[nsh++];
    h->xShells = malloc((nrAtoms + 1) * sizeof(int));
    h->xL = malloc((ncopy * nsh + 1) * sizeof(int));
    h->xK = malloc((ncopy * nsh + 1) * sizeof(int));
    h->xD = malloc((ncopy * nprim + 1) * sizeof(double));
    h->xA = malloc((ncopy * nprim + 1) * sizeof(double));
    int *par = malloc((ncopy * nsh + 1) * sizeof(int)), *vsh = malloc((ncopy * nsh + 1) * sizeof(int));
    int *vloc = malloc((ncopy * nsh + 1) * sizeof(int));
    int s = 0, p = 0, xs = 0, xp = 0;
    for (int i = 0; i < nrAtoms; i++) {
      int cnt = 0;
      for (int j = 0; j < shellsBS[i]; j++, s++) {
        const int l = lBS[s], K = KBS[s], x0 = xs;
        for (int c = 0; c < ncopy; c++) {
          int dl, zp;
          ecp_deriv_copy(n, c, &dl, &zp);
          if (l + dl < 0) break; /* the lowered copies come last */
          h->xL[xs] = l + dl;
          h->xK[xs] = K;
          par[xs] = x0;
          vsh[xs] = c; /* 0 = the caller's shell, > 0 = a copy */
          vloc[xs] = j;
          for (int k = 0; k < K; k++, xp++) {
            double dd = dBS[p + k];
            for (int q = 0; q < zp; q++) dd *= aBS[p + k]; /* da *= zeta  (src/type2.c:267-269) */
            h->xA[xp] = aBS[p + k];
            h->xD[xp] = dd;
          }
          xs++;
          cnt++;
        }
        p += K;
      }
      h->xShells[i] = cnt;
    }
    opts.screenParent = par;
    opts.deriv = n;
    opts.virtShift = vsh;
    opts.virtLocal = vloc;
    h->tab = ecp_tables_build(nrAtoms, geometry, shellsECP, lECP, KECP, nECP, dECP, aECP, h->xShells, h->xL, h->xK, h->xD,
                              h->xA, largeGridOrder, tolerance, accuracy, &opts);
    free(par);
    free(vsh);
    free(vloc);
  } else
    h->tab = ecp_tables_build(nrAtoms, geometry, shellsECP, lECP, KECP, nECP, dECP, aECP, shellsBS, lBS, KBS, dBS, aBS,
                              largeGridOrder, tolerance, accuracy, &opts);
  if (!h->tab) {
    snprintf(g_apierr, sizeof(g_apierr), "libecp_b200: %s", ecp_tables_last_error());
    free_expanded(h);
    free(h);
    return NULL;
  }
  if (h->tab->v.nTypes == 0) {
    h->empty = 1;
    return h;
  }
  if (g_tables_only) { /* test hook: tables + builder without a device; every compute entry point fails */
    h->bb = ecp_batch_new(h->tab);
    h->bb2 = ecp_batch_new(h->tab);
    ecp_batch_share_scratch(h->bb2, h->bb);
    return h;
  }
  int dev = g_device;
  if (dev < 0) {
    const char *e = getenv("LIBECP_B200_DEVICE");
    dev = e ? atoi(e) : 0;
  }
  h->dev = ecpdev_create(&h->tab->v, dev);
  if (!h->dev) { /* no CPU fallback: fail loudly */
    fprintf(stderr, "libecp_b200: cannot create device context: %s\n", ecpdev_last_error());
    ecp_tables_free(h->tab);
    free_expanded(h);
    free(h);
    return NULL;
  }
  h->bb = ecp_batch_new(h->tab);
  h->bb2 = ecp_batch_new(h->tab);
  ecp_batch_share_scratch(h->bb2, h->bb);
  return h;
}

void libECP_free(libECPHandle *h) {
  if (!h) return;
  worker_free(h->worker);
  if (h->dev) ecpdev_destroy(h->dev);
  if (h->bb) ecp_batch_free(h->bb);
  if (h->bb2) ecp_batch_free(h->bb2);
  if (h->tab) ecp_tables_free(h->tab);
  free(h->hostBlocks);
  free_expanded(h);
  free(h);
}

/* device scratch and result buffers of freed handles are parked for the next handle (ecp_cuda.cu); give them back */
void libecp_b200_release_cache(void) { ecpdev_release_cache(); }

void libecp_b200_set_shard(libECPHandle *h, int rank, int world) {
  h->rank = rank;
  h->world = world < 1 ? 1 : world;
}

static void add_stats(libECPHandle *h, const EcpBatchBuf *bb, const EcpDevStats *st, double msBuild) {
  libecp_b200_stats_t *s = &h->stats;
  const EcpBatch *b = &bb->b;
  s->nominal_triples += bb->nominal;
  s->executed_triples += b->nTriples;
  s->shell_slots += b->nSSlots;
  s->atom_slots += b->nASlots;
  s->prim_pairs += b->nPairs;
  s->fast_quadratures += b->clsWork[h->tab->v.nClasses];
  s->fast_failed += st->nFastFail;
  s->fallback_items += st->nFallbackItems;
  s->type1_fallback_pairs += st->nType1Fail;
  s->stale_centre_events += st->nStaleCentre;
  s->kernel_launches += st->launches;
  s->batches += 1;
  s->h2d_bytes += st->h2dBytes;
  s->d2h_bytes += st->d2hBytes;
  s->tables_h2d_bytes = ecpdev_table_bytes(h->dev);
  s->ms_build += msBuild;
  s->ms_tables += st->ms_tables;
  s->ms_fastT += st->ms_fastT;
  s->ms_fallback += st->ms_fallback;
  s->ms_link += st->ms_link;
  s->ms_type1 += st->ms_type1;
  s->ms_chi += st->ms_chi;
  s->ms_shift += st->ms_shift;
  s->ms_device_total += st->ms_total;
}

/* the batch builder as a pipeline stage: batch i+1 is built on a helper host thread (OpenMP team inside) while the
 * calling thread drives the GPU through batch i.  The helper lives as long as the handle (its OpenMP team is reused). */
typedef struct {
  libECPHandle *h;
  EcpBatchBuf *bb;
  int *centre;
  int keepCanon, took;
  int flags, slot, prefetch; /* prefetch: start the H2D of the finished batch into input set `slot` */
  int devEnum;               /* screening only; the device enumerates the triples */
  long long maxTriples;      /* size target of this batch */
  double ms;
} BuildJob;
static void build_job(BuildJob *j) {
  const double t0 = now_ms();
  if (j->h->dev) ecpdev_bind_thread(j->h->dev);
  /* matrix runs: the host only screens, the device enumerates the triples (ecp_enum.cuh); runs that deliver callback
   * blocks keep the host enumeration (the replay needs the canonical list) */
  if (j->devEnum) {
    j->bb->triPerPair = j->h->triPerPair;
    j->took = ecp_batch_build_slots(j->h->tab, j->h->geometry, j->centre, j->maxTriples, j->h->rank, j->h->world, j->bb);
  } else
    j->took = ecp_batch_build(j->h->tab, j->h->geometry, j->centre, j->maxTriples, j->h->rank, j->h->world, j->keepCanon,
                              (j->flags & 2) != 0, j->bb);
  j->ms = now_ms() - t0;
  if (j->prefetch && j->took > 0 && j->h->dev) ecpdev_prefetch_batch(j->h->dev, &j->bb->b, j->flags, j->slot);
}
struct BuildWorker {
  pthread_t th;
  pthread_mutex_t mu;
  pthread_cond_t cv;
  BuildJob *job; /* posted job, NULL when idle */
  int done, quit, started;
};
static void *worker_main(void *p) {
  struct BuildWorker *w = p;
#ifdef _OPENMP
  if (g_host_threads > 0) omp_set_num_threads(g_host_threads); /* the ICV is per host thread */
#endif
  pthread_mutex_lock(&w->mu);
  for (;;) {
    while (!w->job && !w->quit) pthread_cond_wait(&w->cv, &w->mu);
    if (w->quit) break;
    BuildJob *j = w->job;
    pthread_mutex_unlock(&w->mu);
    build_job(j);
    pthread_mutex_lock(&w->mu);
    w->job = NULL;
    w->done = 1;
    pthread_cond_broadcast(&w->cv);
  }
  pthread_mutex_unlock(&w->mu);
  return NULL;
}
static struct BuildWorker *worker_new(void) {
  struct BuildWorker *w = calloc(1, sizeof(*w));
  pthread_mutex_init(&w->mu, NULL);
  pthread_cond_init(&w->cv, NULL);
  w->started = pthread_create(&w->th, NULL, worker_main, w) == 0;
  return w;
}
static void worker_free(struct BuildWorker *w) {
  if (!w) return;
  if (w->started) {
    pthread_mutex_lock(&w->mu);
    w->quit = 1;
    pthread_cond_broadcast(&w->cv);
    pthread_mutex_unlock(&w->mu);
    pthread_join(w->th, NULL);
  }
  pthread_mutex_destroy(&w->mu);
  pthread_cond_destroy(&w->cv);
  free(w);
}
static void worker_post(struct BuildWorker *w, BuildJob *j) {
  pthread_mutex_lock(&w->mu);
  w->done = 0;
  w->job = j;
  pthread_cond_broadcast(&w->cv);
  pthread_mutex_unlock(&w->mu);
}
static void worker_wait(struct BuildWorker *w) {
  pthread_mutex_lock(&w->mu);
  while (!w->done) pthread_cond_wait(&w->cv, &w->mu);
  pthread_mutex_unlock(&w->mu);
}

/* batch size targets of a pass (see run_all): a small first batch, a half-size second one, then full size */
static long long pass_batch_size_enum(const libECPHandle *h, int i, int centre) {
  /* device-enumerated batches: the host part of a batch is the screening of its centres (a fraction of a millisecond
   * per 100 centres), so only the very first batch is kept small; everything after it is full size at any world size -
   * a rank of 8 then runs its pass in two batches instead of three or four (about 0.9 ms of fixed device time each) */
  const char *e = getenv("LIBECP_B200_BATCH_TRIPLES");
  const long long full = (e && atoll(e) > 0) ? h->maxTriples : 2 * h->maxTriples; /* 6 M: 83.3 -> 81.0 ms per config-5 pass */
  const long long first = full / 6 < 500000 ? full / 6 : 500000;
  /* streamed host consumer: rows leave the device batch by batch, the rows the LAST batch finishes are the tail of the
   * call - third-size batches over the last 45 % of the centres (about 0.7 ms of fixed cost each) */
  if (h->stream && i > 0 && h->world == 1 && centre > (h->nrAtoms * 11) / 20 && !getenv("LIBECP_B200_STREAM_EVEN")) return full / 3;
  return i == 0 ? first : full;
}
static long long pass_batch_size(const libECPHandle *h, int i) {
  long long full = h->maxTriples;
  if (h->world > 1 && !getenv("LIBECP_B200_BATCH_TRIPLES")) {
    /* a nearly empty batch was measured at ~0.9 ms of device time (the longest fallback item and type-1 pair of a
     * batch are serial chains of a few hundred microseconds, plus ~57 launches); large batches hide most of it, six
     * 0.5 M batches per rank at 8 GPUs less so.  A rank's pass is cut into about three batches (ramp 1/6, 1/2, 1). */
    full = 4 * h->maxTriples / h->world;
    if (full > h->maxTriples) full = h->maxTriples;
    if (full < 500000) full = 500000;
  }
  return i == 0 ? full / 6 : (i == 1 ? full / 2 : full);
}

/* ---- streamed download of the host consumer ----
 * The AO rows of an atom are final as soon as the pass is beyond the last centre that can reach the atom
 * (ecp_atom_last_centre: the builder's atom-level prune).  After every batch the rows that have become final are handed
 * to a helper thread, which packs their non-zero runs, moves them on the download stream and adds them into the caller's
 * matrix (src/getIntegrals.c:36-42) while the next batches compute; only the rows the last batch finishes are left as a
 * tail.  How early rows become final depends on the spatial order of the atoms (a lattice-ordered nanocrystal: 60 % of
 * the upper triangle before the last batch); in the worst case everything is downloaded after the pass, as without
 * streaming.  One download at a time: a batch that ends while the previous download is still running keeps its rows
 * for the next one. */
typedef struct StreamState {
  libECPHandle *h;
  double *I;
  int rowdim;
  int *lastC;               /* per atom */
  unsigned char *sent;      /* per atom */
  unsigned char *ownedRows; /* per AO row, NULL = all (unsharded) */
  unsigned char *mask;      /* rows of the running download */
  pthread_t th;
  int running, done, rc, jobs;
  long long moved;
} StreamState;
static void *stream_thread(void *p) {
  StreamState *s = p;
#ifdef _OPENMP
  if (g_host_threads > 0) omp_set_num_threads(g_host_threads > 5 ? g_host_threads - 3 : g_host_threads); /* the builder's team works beside it */
#endif
  long long moved = 0;
  const int rc = ecpdev_matrix_add_to_host(s->h->dev, s->I, s->rowdim, s->mask, &moved, 1);
  s->moved += moved;
  if (rc && !s->rc) s->rc = rc;
  __atomic_store_n(&s->done, 1, __ATOMIC_RELEASE);
  return NULL;
}
static void stream_join(StreamState *s) {
  if (!s->running) return;
  pthread_join(s->th, NULL);
  s->running = 0;
}
/* rows of the atoms that are final once every centre < centreEnd is done and that have not been sent: into s->mask;
 * returns the number of upper-triangle elements they hold */
static long long stream_collect(StreamState *s, int centreEnd) {
  const EcpTables *t = s->h->tab;
  const EcpHostTables *v = &t->v;
  const int n = v->nAO;
  long long elems = 0;
  memset(s->mask, 0, (size_t)n + 1);
  for (int X = 0; X < v->nrAtoms; X++) {
    if (s->sent[X] || s->lastC[X] < 0 || s->lastC[X] >= centreEnd) continue;
    const int s0 = t->atomFirstShell[X], s1 = t->atomFirstShell[X + 1];
    if (s0 == s1) continue;
    const int r0 = v->shellAO[s0], r1 = (s1 < v->nrShells) ? v->shellAO[s1] : n;
    for (int r = r0; r < r1; r++)
      if (!s->ownedRows || s->ownedRows[r]) {
        s->mask[r] = 1;
        elems += n - r;
      }
  }
  return elems;
}
static void stream_mark_sent(StreamState *s, int centreEnd) {
  for (int X = 0; X < s->h->tab->v.nrAtoms; X++)
    if (s->lastC[X] >= 0 && s->lastC[X] < centreEnd) s->sent[X] = 1;
}
/* after a batch (not the last one of the pass): centres < centreEnd are done */
static void stream_after_batch(StreamState *s, int centreEnd) {
  if (s->running) {
    if (!__atomic_load_n(&s->done, __ATOMIC_ACQUIRE)) return; /* still busy: these rows go with the next download */
    stream_join(s);
  }
  if (s->rc) return;
  const long long elems = stream_collect(s, centreEnd);
  {
    const char *e = getenv("LIBECP_B200_STREAM_MIN_BYTES"); /* tests stream small matrices */
    const long long minBytes = e ? atoll(e) : (16LL << 20);
    if (elems == 0 || elems * 8 < minBytes) return; /* not worth a download of its own yet */
  }
  stream_mark_sent(s, centreEnd);
  s->done = 0;
  if (pthread_create(&s->th, NULL, stream_thread, s) == 0) {
    s->running = 1;
    s->jobs++;
  } else {
    stream_thread(s);
  }
}

/* drive all batches; flags as ecpdev_run_batch; cb may be NULL */
static int run_all(libECPHandle *h, int flags, ECPCallback cb, void *args) {
  int result = 0, centre = 0;
  memset(&h->stats, 0, sizeof(h->stats));
  if (h->empty) return 0;
  if (!h->dev) {
    fprintf(stderr, "libecp_b200: handle has no device context (tables-only); there is no CPU compute path\n");
    return -1;
  }
  ecpdev_invalidate_prefetch(h->dev);
  EcpBatchBuf *bufs[2] = {h->bb, h->bb2};
  /* Batch size: the builder works one batch ahead of the GPU, so the first batch of a pass is built with the GPU idle
   * and is kept small; a rank of a sharded run owns 1/world of the triples and takes smaller batches so that its pass
   * still has enough of them to pipeline (below ~0.5 M triples the per-batch launches and kernel tails start to show). */
  BuildJob job = {h, bufs[0], &centre, cb != NULL, 0, flags, 0, 0, 0, pass_batch_size(h, 0), 0.0};
  {
    const char *e = getenv("LIBECP_B200_ENUM"); /* =host: the host builder enumerates the triples of matrix runs too */
    job.devEnum = (flags == 1) && !h->tab->deriv && !(e && !strcmp(e, "host"));
    if (job.devEnum) job.maxTriples = pass_batch_size_enum(h, 0, 0);
  }
  build_job(&job); /* first batch: nothing to overlap with */
  if (getenv("LIBECP_B200_TRACE")) fprintf(stderr, "[libecp_b200] first batch built in %.1f ms\n", job.ms);
  if (!h->worker) h->worker = worker_new();
  const int threaded = h->worker->started && !getenv("LIBECP_B200_NO_PIPELINE");
  for (int i = 0; job.took > 0; i++) {
    EcpBatchBuf *cur = bufs[i & 1];
    const double msBuild = job.ms;
    const double t0 = now_ms();
    const int centreEnd = centre; /* centres < centreEnd are done when this batch is (the helper thread moves the cursor on) */
    /* next batch on the helper thread (it advances the centre cursor; nobody else reads it meanwhile) */
    job.bb = bufs[(i + 1) & 1];
    job.slot = (i + 1) & 1;
    job.maxTriples = job.devEnum ? pass_batch_size_enum(h, i + 1, centre) : pass_batch_size(h, i + 1); /* ramp: the GPU must not wait for a full-size build behind a small batch */
    job.prefetch = threaded;
    if (threaded) worker_post(h->worker, &job);
    EcpBatch *b = &cur->b;
    if ((flags & 2) && (size_t)b->outTotal > h->hostBlocksCap) {
      free(h->hostBlocks);
      h->hostBlocksCap = (size_t)b->outTotal * 5 / 4 + 1024;
      h->hostBlocks = malloc(h->hostBlocksCap * sizeof(double));
    }
    EcpDevStats st;
    const int rc = ecpdev_run_batch(h->dev, b, flags, i & 1, (flags & 2) ? h->hostBlocks : NULL, &st);
    if (threaded)
      worker_wait(h->worker);
    else
      build_job(&job);
    if (rc) {
      fprintf(stderr, "libecp_b200: device failure: %s\n", ecpdev_last_error());
      return -rc;
    }
    if (b->devEnum && b->pairCand > 0 && b->nTriples > 0) /* feedback for the size estimate of the next batches */
      h->triPerPair = (double)b->nTriples / (double)b->pairCand;
    add_stats(h, cur, &st, msBuild);
    if (getenv("LIBECP_B200_TRACE"))
      fprintf(stderr, "[libecp_b200] rank %d batch: %d triples build %.1f ms (overlapped) run_batch+join wall %.1f ms device %.1f ms\n",
              h->rank, b->nTriples, msBuild, now_ms() - t0, st.ms_total);
    if (st.err1 && result == 0) result = 1; /* src/libecp.h:23-27 */
    if (st.err2 && result == 0) result = 2;
    if (result) break;
    if (h->stream && job.took > 0) stream_after_batch(h->stream, centreEnd);
    if (cb) { /* replay in the reference's loop order, type 1 then type 2 (src/libecp.c:332-373) */
      typedef void (*CallSite)(int, int, int, int, int, int, int, int, int, double *, void *);
      CallSite call = (CallSite)cb;
      const EcpBatchBuf *bb = cur;
      for (int k = 0; k < bb->nCanon; k++) {
        const int sa = bb->cnShA[k], sb = bb->cnShB[k]; /* blocks of a derivative run have the shifted sizes */
        const int mixed = h->tab->deriv == 2 && sa + sb == 1 && sa >= 0 && sb >= 0; /* momentum unchanged (src/libecp.c:362-369) */
        const int nb = mixed ? IJK_DIM(bb->cnLa[k]) * IJK_DIM(bb->cnLb[k]) : IJK_DIM(bb->cnLa[k] + sa) * IJK_DIM(bb->cnLb[k] + sb);
        double *blk = h->hostBlocks + bb->cnOut[k];
        call(bb->cnA[k], bb->cnS1[k], bb->cnLa[k], sa, bb->cnB[k], bb->cnS2[k], bb->cnLb[k], sb, bb->cnC[k], blk, args);
        call(bb->cnA[k], bb->cnS1[k], bb->cnLa[k], sa, bb->cnB[k], bb->cnS2[k], bb->cnLb[k], sb, bb->cnC[k], blk + nb, args);
      }
    }
  }
  return result;
}

int calculateECPIntegrals(libECPHandle *h, ECPCallback cb, void *args) { return run_all(h, 2, cb, args); }

/* AO rows of the shells whose shell-pair rows this rank owns (NULL for an unsharded handle); caller frees */
static unsigned char *owned_rows(const libECPHandle *h) {
  if (h->world <= 1) return NULL;
  const EcpHostTables *v = &h->tab->v;
  unsigned char *owned = calloc(v->nAO + 1, 1);
  for (int s = 0; s < v->nrShells; s++)
    if (ecp_pair_owner(h->tab, s, s, h->world) == h->rank)
      for (int k = 0; k < IJK_DIM(v->shellL[s]); k++) owned[v->shellAO[s] + k] = 1;
  return owned;
}

int libecp_b200_integrals_device(libECPHandle *h, void **devMatrix, int *nAO) {
  if (nAO) *nAO = h->tab->v.nAO;
  if (h->empty) {
    if (devMatrix) *devMatrix = NULL;
    return 0;
  }
  if (!h->dev) return -1;
  if (h->tab->deriv) { /* the matrix consumer is the n = 0 one-call interface (reference src/getIntegrals.c:78-82) */
    snprintf(g_apierr, sizeof(g_apierr), "libecp_b200: derivative handles deliver callback blocks only (calculateECPIntegrals)");
    return -1;
  }
  unsigned char *owned = owned_rows(h);
  int rc = ecpdev_matrix_begin(h->dev, owned, (long long)h->world * 1000003LL + h->rank);
  free(owned);
  if (rc) return -rc;
  rc = run_all(h, 1, NULL, NULL);
  if (devMatrix) *devMatrix = ecpdev_matrix_ptr(h->dev);
  return rc;
}

/* spherical-harmonic (pure 5d / 7f ...) form of the result: dimension, device-resident matrix, host accumulation */
int libecp_b200_spherical_dim(libECPHandle *h) {
  int n = 0;
  for (int s = 0; s < h->tab->v.nrShells; s++) n += 2 * h->tab->v.shellL[s] + 1;
  return n;
}
int libecp_b200_spherical_device(libECPHandle *h, void **devS, int *nSph) {
  if (nSph) *nSph = libecp_b200_spherical_dim(h);
  if (h->empty) {
    if (devS) *devS = NULL;
    return 0;
  }
  void *dm = NULL;
  const int rc = libecp_b200_integrals_device(h, &dm, NULL);
  if (rc < 0) return rc;
  if (ecpdev_spherical(h->dev, devS, nSph)) return -1;
  return rc;
}
int libecp_b200_spherical_host(libECPHandle *h, int rowdim, double *S) {
  if (h->empty) return 0;
  void *dm = NULL;
  const int rc = libecp_b200_integrals_device(h, &dm, NULL);
  if (rc < 0) return rc;
  if (ecpdev_spherical_add_to_host(h->dev, S, rowdim)) return -1;
  return rc;
}

void *libecp_b200_matrix_ptr(libECPHandle *h) { return (h && h->dev) ? ecpdev_matrix_ptr(h->dev) : NULL; }

int libecp_b200_pair_owner(libECPHandle *h, int a, int b, int world) { return ecp_pair_owner(h->tab, a, b, world); }
/* AO rows (ascending) whose shell-pair rows `rank` of `world` owns; returns the count (cap may be 0 to size) */
long long libecp_b200_owned_rows(libECPHandle *h, int rank, int world, int *rows, long long cap) {
  const EcpHostTables *v = &h->tab->v;
  long long n = 0;
  for (int s = 0; s < v->nrShells; s++)
    if (ecp_pair_owner(h->tab, s, s, world) == rank)
      for (int k = 0; k < IJK_DIM(v->shellL[s]); k++, n++)
        if (rows && n < cap) rows[n] = v->shellAO[s] + k;
  return n;
}
int libecp_b200_pack_rows(libECPHandle *h, const int *rows, long long nrows, void *devPacked, long long cap, long long *elems) {
  if (h->empty) {
    if (elems) *elems = 0;
    return 0;
  }
  if (!h->dev) return -1;
  return ecpdev_matrix_rows(h->dev, 0, rows, nrows, devPacked, cap, elems) ? -1 : 0;
}
int libecp_b200_unpack_rows(libECPHandle *h, const int *rows, long long nrows, const void *devPacked, long long cap) {
  if (h->empty) return 0;
  if (!h->dev) return -1;
  return ecpdev_matrix_rows(h->dev, 1, rows, nrows, (void *)devPacked, cap, NULL) ? -1 : 0;
}

/* ---- C-ABI collective (include/libecp_b200.h): NCCL all-gather of the shards of a device-resident result ---- */
int libecp_b200_comm_unique_id(void *id128) { return ecpdev_comm_unique_id(id128) ? -1 : 0; }
static int comm_layout(libECPHandle *h, int world) {
  int **rows = calloc(world, sizeof(int *));
  long long *nrows = calloc(world, sizeof(long long));
  for (int r = 0; r < world; r++) {
    nrows[r] = libecp_b200_owned_rows(h, r, world, NULL, 0);
    rows[r] = malloc((size_t)(nrows[r] + 1) * sizeof(int));
    libecp_b200_owned_rows(h, r, world, rows[r], nrows[r]);
  }
  const int rc = ecpdev_allgather_layout(h->dev, (const int *const *)rows, nrows);
  for (int r = 0; r < world; r++) free(rows[r]);
  free(rows);
  free(nrows);
  return rc ? -1 : 0;
}
int libecp_b200_comm_init(libECPHandle *h, int rank, int world, const void *id128) {
  if (!h || h->empty || !h->dev || world < 1 || rank < 0 || rank >= world) return -1;
  if (ecpdev_comm_init(h->dev, rank, world, id128, NULL)) return -1;
  libecp_b200_set_shard(h, rank, world);
  return comm_layout(h, world);
}
int libecp_b200_comm_attach(libECPHandle *h, void *ncclComm, int rank, int world) {
  if (!h || h->empty || !h->dev || !ncclComm || world < 1 || rank < 0 || rank >= world) return -1;
  if (ecpdev_comm_init(h->dev, rank, world, NULL, ncclComm)) return -1;
  libecp_b200_set_shard(h, rank, world);
  return comm_layout(h, world);
}
int libecp_b200_allgather(libECPHandle *h, long long *bytesReceived) {
  if (!h || h->empty || !h->dev) return -1;
  const double t0 = now_ms();
  const int rc = ecpdev_allgather(h->dev, bytesReceived) ? -1 : 0;
  if (getenv("LIBECP_B200_TRACE")) fprintf(stderr, "[libecp_b200] rank %d allgather issued in %.2f ms\n", h->rank, now_ms() - t0);
  return rc;
}
int libecp_b200_device_sync(libECPHandle *h) {
  const double t0 = now_ms();
  const int rc = (h && h->dev) ? (ecpdev_sync(h->dev) ? -1 : 0) : 0;
  if (h && getenv("LIBECP_B200_TRACE")) fprintf(stderr, "[libecp_b200] rank %d device_sync %.2f ms\n", h->rank, now_ms() - t0);
  return rc;
}
void libecp_b200_comm_free(libECPHandle *h) {
  if (h && h->dev) ecpdev_comm_destroy(h->dev);
}

/* Host consumer (getIntegrals): the pass is cut into `panels` row panels - the rows this rank owns are dealt once more
 * (libecp_b200_set_shard machinery: panel p of P inside rank r of W owns what rank p * W + r of W * P would) - and run
 * one after the other; as soon as a panel's pass has returned its rows are final, and a helper thread packs them,
 * moves them over PCIe on the download stream and adds them into the caller's matrix (src/getIntegrals.c:36-42) while
 * the next panel computes.  Only the last panel's transfer (1 / panels of the 1.44 GB of config 5) is left as a tail.
 * Every panel screens all centres again and recomputes their tables (~3 ms per config-5 pass), so few panels. */
typedef struct {
  libECPHandle *h;
  double *I;
  int rowdim, rank, world, rc;
  long long moved;
  pthread_t th;
} PanelJob;
static void *panel_download(void *p) {
  PanelJob *j = p;
  unsigned char *owned = NULL;
  if (j->world > 1) {
    const EcpHostTables *v = &j->h->tab->v;
    owned = calloc(v->nAO + 1, 1);
    for (int s = 0; s < v->nrShells; s++)
      if (ecp_pair_owner(j->h->tab, s, s, j->world) == j->rank)
        for (int k = 0; k < IJK_DIM(v->shellL[s]); k++) owned[v->shellAO[s] + k] = 1;
  }
#ifdef _OPENMP
  if (g_host_threads > 0) omp_set_num_threads(g_host_threads);
#endif
  j->rc = ecpdev_matrix_add_to_host(j->h->dev, j->I, j->rowdim, owned, &j->moved, 1);
  free(owned);
  return NULL;
}
static int host_panels(const libECPHandle *h) {
  const char *e = getenv("LIBECP_B200_HOST_PANELS");
  if (e && atoi(e) > 0) return atoi(e) > 16 ? 16 : atoi(e);
  /* Dense download (LIBECP_B200_D2H=dense): worth it when the transfer is long compared with a panel's fixed costs -
   * upper triangle above ~256 MB.  The default sparse download moves a third of the bytes and is bound by the host +=
   * (configuration 5: 14 ms after an 81 ms pass); an extra panel costs 5-8 ms (per-centre tables, screening, smaller
   * batches, the download competing with the builder for host cores) and saves half of that tail at best: one panel
   * (measured 95.2 / 96.3 / 100.6 ms for 1 / 2 / 3 panels, profiles/r2/README.md). */
  const char *m = getenv("LIBECP_B200_D2H");
  if (!(m && !strcmp(m, "dense"))) return 1;
  const double bytes = 4.0 * (double)h->tab->v.nAO * h->tab->v.nAO / h->world;
  return bytes > 256e6 ? 3 : 1;
}
int libecp_b200_integrals_host(libECPHandle *h, int rowdim, double *I) {
  const double tCall = now_ms();
  const int P = (h->empty || !h->dev || h->tab->deriv) ? 1 : host_panels(h);
  if (P <= 1) {
    void *dm = NULL;
    StreamState st;
    memset(&st, 0, sizeof(st));
    {
      const char *e = getenv("LIBECP_B200_STREAM_D2H");
      if (!h->empty && h->dev && !h->tab->deriv && !(e && !strcmp(e, "0"))) {
        const int nat = h->tab->v.nrAtoms, n = h->tab->v.nAO;
        st.h = h;
        st.I = I;
        st.rowdim = rowdim;
        st.lastC = malloc((size_t)(nat + 1) * sizeof(int));
        st.sent = calloc((size_t)nat + 1, 1);
        st.mask = calloc((size_t)n + 1, 1);
        st.ownedRows = owned_rows(h);
        ecp_atom_last_centre(h->tab, h->geometry, st.lastC);
        h->stream = &st;
      }
    }
    const int rc = libecp_b200_integrals_device(h, &dm, NULL);
    h->stream = NULL;
    if (st.h) {
      stream_join(&st);
      long long rest = 0;
      int rc2 = st.rc;
      if (rc >= 0 && !rc2) {
        const double tA = now_ms();
        rest = stream_collect(&st, h->nrAtoms + 1); /* everything that has not been sent (atoms no centre reaches stay zero) */
        if (rest) {
          long long moved = 0;
          rc2 = ecpdev_matrix_add_to_host(h->dev, I, rowdim, st.mask, &moved, 0);
          st.moved += moved;
        }
        if (getenv("LIBECP_B200_TRACE"))
          fprintf(stderr, "[libecp_b200] rank %d/%d integrals_host: %d streamed downloads, tail %.1f MB dense-equivalent in %.1f ms (whole call %.1f ms)\n",
                  h->rank, h->world, st.jobs, rest * 8 / 1e6, now_ms() - tA, now_ms() - tCall);
      }
      free(st.lastC);
      free(st.sent);
      free(st.mask);
      free(st.ownedRows);
      if (rc < 0) return rc;
      h->stats.d2h_bytes += st.moved;
      if (rc2) return -rc2;
      return rc;
    }
    if (rc < 0 || h->empty) return rc;
    long long moved = 0;
    unsigned char *owned = owned_rows(h); /* only the AO rows of the shells this rank owns can be non-zero */
    const double tA = now_ms();
    const int rc2 = ecpdev_matrix_add_to_host(h->dev, I, rowdim, owned, &moved, 0);
    if (getenv("LIBECP_B200_TRACE"))
      fprintf(stderr, "[libecp_b200] rank %d/%d integrals_host: add_to_host %.1f ms (whole call so far %.1f ms)\n", h->rank, h->world,
              now_ms() - tA, now_ms() - tCall);
    free(owned);
    h->stats.d2h_bytes += moved;
    if (rc2) return -rc2;
    return rc;
  }
  /* clear the rows of the whole rank once; the panels then accumulate without clearing */
  unsigned char *owned = owned_rows(h);
  int rc = ecpdev_matrix_begin(h->dev, owned, (long long)h->world * 1000003LL + h->rank);
  free(owned);
  if (rc) return -rc;
  const int rank0 = h->rank, world0 = h->world;
  libecp_b200_stats_t tot;
  memset(&tot, 0, sizeof(tot));
  PanelJob jobs[16];
  int result = 0, started = 0;
  for (int p = 0; p < P; p++) {
    h->rank = p * world0 + rank0; /* deal % (world0 P) = p world0 + rank0  =>  deal % world0 = rank0: a subset of the rank's rows */
    h->world = world0 * P;
    const int r = run_all(h, 1, NULL, NULL);
    { /* statistics of the call = sum over the panels */
      const libecp_b200_stats_t *s = &h->stats;
      long long *a = (long long *)&tot;
      const long long *c = (const long long *)s;
      const int nll = (int)((const char *)&s->ms_build - (const char *)s) / (int)sizeof(long long);
      for (int k = 0; k < nll; k++) a[k] += c[k];
      tot.tables_h2d_bytes = s->tables_h2d_bytes;
      tot.nominal_triples = s->nominal_triples; /* every panel walks the same loop domain */
      double *ad = &tot.ms_build;
      const double *cd = &s->ms_build;
      for (int k = 0; k < 9; k++) ad[k] += cd[k];
    }
    if (r < 0) {
      result = r;
      break;
    }
    if (r && !result) result = r;
    PanelJob *j = &jobs[p];
    j->h = h;
    j->I = I;
    j->rowdim = rowdim;
    j->rank = h->rank;
    j->world = h->world;
    j->rc = 0;
    j->moved = 0;
    if (p > 0) pthread_join(jobs[p - 1].th, NULL); /* one download at a time: they share the PCIe link and the host cores */
    if (pthread_create(&j->th, NULL, panel_download, j)) {
      panel_download(j);
      j->th = 0;
    }
    started = p + 1;
  }
  h->rank = rank0;
  h->world = world0;
  for (int p = (started > 1 ? started - 1 : 0); p < started; p++)
    if (jobs[p].th) pthread_join(jobs[p].th, NULL);
  h->stats = tot;
  for (int p = 0; p < started; p++) {
    h->stats.d2h_bytes += jobs[p].moved;
    if (jobs[p].rc && result >= 0) result = -jobs[p].rc;
  }
  if (getenv("LIBECP_B200_TRACE"))
    fprintf(stderr, "[libecp_b200] rank %d/%d integrals_host: %d panels, whole call %.1f ms\n", h->rank, h->world, P, now_ms() - tCall);
  return result;
}

int getIntegrals(int nrAtoms, double *geometry, int *shellsECP, int *KECP, int *lECP, double *nECP, double *dECP,
                 double *aECP, int *shellsBS, int *lBS, int *KBS, double *dBS, double *aBS, int largeGridOrder,
                 double tolerance, double accuracy, int rowdim, double *I) {
  libECPHandle *h = libECP_init(nrAtoms, geometry, shellsECP, lECP, KECP, nECP, dECP, aECP, shellsBS, lBS, KBS, dBS,
                                aBS, 0, -1, NULL, largeGridOrder, tolerance, accuracy);
  if (NULL == h) {
    printf("error initializing libECP\n"); /* src/getIntegrals.c:84-87 */
    return 1;
  }
  libecp_b200_integrals_host(h, rowdim, I); /* the reference ignores the integrate rc as well (src/getIntegrals.c:88) */
  libECP_free(h);
  return 0;
}

void libecp_b200_get_stats(libECPHandle *h, libecp_b200_stats_t *out) { *out = h->stats; }

int libecp_b200_screening(libECPHandle *h, int centre, int *end_l, int *start, int *end, int *skip) {
  const EcpTables *t = h->tab;
  if (h->empty || centre < 0 || centre >= h->nrAtoms || t->atomType[centre] < 0) return -1;
  const EcpType *T = &t->types[t->atomType[centre]];
  for (int l = 0; l < T->L; l++) end_l[l] = T->endl[l];
  for (int s = 0; s < t->v.nrShells; s++) {
    const double *rX = h->geometry + 3 * t->shellAtom[s], *rC = h->geometry + 3 * centre;
    const double x = rC[0] - rX[0], y = rC[1] - rX[1], z = rC[2] - rX[2];
    ecp_shell_window(t, T->endLast, t->shellRadius[s], __builtin_sqrt(x * x + y * y + z * z), &start[s], &end[s], &skip[s]);
  }
  return 0;
}

int libecp_b200_host_table(libECPHandle *h, const char *name, const double **ptr) {
  const EcpTables *t = h->tab;
  const EcpHostTables *v = &t->v;
  if (h->empty) return 0;
  const int cdT = C_DIM(v->tmDim);
#define RET(p, n) \
  do {            \
    *ptr = (p);   \
    return (n);   \
  } while (0)
  if (!strcmp(name, "fac")) RET(t->fac, v->nfac);
  if (!strcmp(name, "dfac")) RET(t->dfac, v->nfac);
  if (!strcmp(name, "poly2sph")) RET(t->poly2sph, cdT * L_DIM(v->tmDim));
  if (!strcmp(name, "omega")) RET(t->omega, v->nomega);
  if (!strcmp(name, "small_x")) RET(t->small_x, ECP_SMALL_ORDER);
  if (!strcmp(name, "small_w")) RET(t->small_w, ECP_SMALL_ORDER);
  if (!strcmp(name, "large_x")) RET(t->large_x, v->largeOrder);
  if (!strcmp(name, "large_w")) RET(t->large_w, v->largeOrder);
  if (!strcmp(name, "bessel")) RET(t->besselK, (v->besselLMax + 1) * 1601);
  if (!strcmp(name, "besselC")) RET(t->besselC, v->besselLMax + 1);
  if (!strcmp(name, "shellRadius")) RET(t->shellRadius, v->nrShells);
  if (!strcmp(name, "besselT")) RET(t->besselT, 1601 * v->besselStride);
  if (!strcmp(name, "small_rs")) RET(t->small_rs, ECP_SMALL_SLOTS);
  if (!strcmp(name, "small_ws")) RET(t->small_ws, ECP_SMALL_SLOTS);
  if (!strcmp(name, "large_xs")) RET(t->large_xs, v->largeSlots);
  if (!strcmp(name, "large_ws")) RET(t->large_ws, v->largeSlots);
  if (!strcmp(name, "typeUtab")) RET(t->typeUtab, v->nTypes * v->maxLECP * v->nU * ECP_SMALL_SLOTS);
  if (!strcmp(name, "typeUL")) RET(t->typeUL, v->nTypes * ECP_SMALL_SLOTS);
  if (!strcmp(name, "cart2sph")) RET(t->cart2sph, v->ncart2sph);
#undef RET
  *ptr = NULL;
  return 0;
}

int libecp_b200_host_itable(libECPHandle *h, const char *name, int *out, int cap) {
  const EcpTables *t = h->tab;
  const EcpHostTables *v = &t->v;
  int n = 0;
  if (h->empty) return 0;
#define PUT(x)              \
  do {                      \
    if (n < cap) out[n] = (x); \
    n++;                    \
  } while (0)
  if (!strcmp(name, "small_oidx"))
    for (int i = 0; i < ECP_SMALL_SLOTS; i++) PUT(t->small_oidx[i]);
  else if (!strcmp(name, "large_oidx"))
    for (int i = 0; i < v->largeSlots; i++) PUT(t->large_oidx[i]);
  else if (!strcmp(name, "small_meta")) {
    for (int i = 0; i < ECP_SMALL_LEVELS; i++) PUT(v->small_levPairs[i]);
    for (int i = 0; i < ECP_SMALL_LEVELS; i++) PUT(v->small_levJ[i]);
    for (int i = 0; i < ECP_SMALL_LEVELS; i++) PUT(v->small_levN[i]);
    for (int i = 0; i <= ECP_SMALL_LEVELS; i++) PUT(v->small_levSlot[i]);
  } else if (!strcmp(name, "dims")) {
    PUT(v->maxLECP); PUT(v->maxLBS); PUT(v->maxAlpha); PUT(v->maxLambda); PUT(v->tmDim); PUT(v->besselLMax);
    PUT(v->besselStride); PUT(v->largeOrder); PUT(v->largeSlots); PUT(v->largeLevels); PUT(v->nU); PUT(v->nTypes);
    PUT(v->nClasses); PUT(v->maxQPerL); PUT(v->nrShells); PUT(v->nAO);
  } else if (!strcmp(name, "ijk"))
    for (int i = 0; i < 3 * C_DIM(v->tmDim); i++) PUT(t->ijk[i]);
  else if (!strcmp(name, "ijkIndex"))
    for (int i = 0; i < v->ijkDim * v->ijkDim * v->ijkDim; i++) PUT(t->ijkIndex[i]);
  else if (!strcmp(name, "atomType"))
    for (int i = 0; i < h->nrAtoms; i++) PUT(t->atomType[i]);
  else if (!strcmp(name, "lastCentre")) { /* streamed download: last centre that can reach each atom (builder.c) */
    int *lc = malloc((size_t)(h->nrAtoms + 1) * sizeof(int));
    ecp_atom_last_centre(t, h->geometry, lc);
    for (int i = 0; i < h->nrAtoms; i++) PUT(lc[i]);
    free(lc);
  }
#undef PUT
  return n;
}

/* executed-triple list of the whole job in the reference's loop order (host only; tests):
 * rows of 7 ints (A, s1, la, B, s2, lb, C); returns the count (fills at most cap rows) */
long long libecp_b200_triple_list(libECPHandle *h, int *out, long long cap) {
  long long n = 0;
  int centre = 0;
  if (h->empty) return 0;
  EcpBatchBuf *bb = ecp_batch_new(h->tab);
  while (ecp_batch_build(h->tab, h->geometry, &centre, 1 << 20, h->rank, h->world, 1, 1, bb) > 0)
    for (int k = 0; k < bb->nCanon; k++, n++)
      if (n < cap) {
        int *r = out + 7 * n;
        r[0] = bb->cnA[k]; r[1] = bb->cnS1[k]; r[2] = bb->cnLa[k]; r[3] = bb->cnB[k];
        r[4] = bb->cnS2[k]; r[5] = bb->cnLb[k]; r[6] = bb->cnC[k];
      }
  ecp_batch_free(bb);
  return n;
}

/* callback keys of the whole job in call order, one row of 9 ints per executed (shifted) triple - every row stands for
 * the type-1 and the type-2 callback (A, s1, la, shifta, B, s2, lb, shiftb, C); host only; tests */
long long libecp_b200_callback_keys(libECPHandle *h, int *out, long long cap) {
  long long n = 0;
  int centre = 0;
  if (h->empty) return 0;
  EcpBatchBuf *bb = ecp_batch_new(h->tab);
  while (ecp_batch_build(h->tab, h->geometry, &centre, 1 << 20, h->rank, h->world, 1, 1, bb) > 0)
    for (int k = 0; k < bb->nCanon; k++, n++)
      if (n < cap) {
        int *r = out + 9 * n;
        r[0] = bb->cnA[k]; r[1] = bb->cnS1[k]; r[2] = bb->cnLa[k]; r[3] = bb->cnShA[k]; r[4] = bb->cnB[k];
        r[5] = bb->cnS2[k]; r[6] = bb->cnLb[k]; r[7] = bb->cnShB[k]; r[8] = bb->cnC[k];
      }
  ecp_batch_free(bb);
  return n;
}

/* host batch builder alone (no device work): wall ms of building every batch of one pass, with the batch sizes and
 * the two alternating buffers of a real pass; for tuning / tests */
double libecp_b200_build_only(libECPHandle *h, long long *triples, int *batches) {
  long long n = 0;
  int centre = 0, nb = 0;
  if (triples) *triples = 0;
  if (batches) *batches = 0;
  if (h->empty) return 0.0;
  EcpBatchBuf *bufs[2] = {h->bb, h->bb2};
  const double t0 = now_ms();
  const char *e = getenv("LIBECP_B200_ENUM");
  if (!h->tab->deriv && !(e && !strcmp(e, "host"))) { /* what a matrix run leaves to the host: screening and slot layout */
    while (ecp_batch_build_slots(h->tab, h->geometry, &centre, pass_batch_size_enum(h, nb, centre), h->rank, h->world, bufs[nb & 1]) > 0) {
      n += bufs[nb & 1]->b.pairCand; /* shell pairs the device will test (it counts the triples itself) */
      nb++;
    }
  } else
    while (ecp_batch_build(h->tab, h->geometry, &centre, pass_batch_size(h, nb), h->rank, h->world, 0, 0, bufs[nb & 1]) > 0) {
      n += bufs[nb & 1]->b.nTriples;
      nb++;
    }
  const double ms = now_ms() - t0;
  if (triples) *triples = n;
  if (batches) *batches = nb;
  return ms;
}

int libecp_b200_debug_unit(libECPHandle *h, const char *what, int n, const double *in, long long nin, const int *ipar, int npar,
                           double *out, long long nout) {
  if (!h || h->empty || !h->dev) return -1;
  return ecpdev_unit(h->dev, what, n, in, nin, ipar, npar, out, nout) ? -1 : 0;
}
int libecp_b200_debug_fetch(libECPHandle *h, const char *what, double *dst, long long n) {
  if (h->empty || !h->dev) return -1;
  return ecpdev_debug_fetch(h->dev, what, dst, n);
}

/* libint Cartesian ordering (reference src/dimensions.c:17-57) */
int *cartesianShellOrder(const int am) {
  int *t = calloc(3 * C_DIM(am), sizeof(int));
  for (int l = 0; l <= am; l++) {
    int c = 0;
    for (int i = 0; i <= l; i++)
      for (int j = 0; j <= i; j++, c++) {
        int *e = t + CIJK_INDEX(l, c);
        e[0] = l - i;
        e[1] = i - j;
        e[2] = j;
      }
  }
  return t;
}
int *cartesianShellOrderIndex(const int am, int *ijk) {
  const int dim = am + 1;
  int *t = calloc(dim * dim * dim, sizeof(int));
  for (int l = 0; l <= am; l++)
    for (int c = 0; c < IJK_DIM(l); c++) {
      const int *e = ijk + CIJK_INDEX(l, c);
      t[e[0] * dim * dim + e[1] * dim + e[2]] = C_INDEX(l, c);
    }
  return t;
}
