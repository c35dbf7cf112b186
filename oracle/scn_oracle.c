/*
 * ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called
 * from the product (detection_3d_b200/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Plain-C restatement of the CPU algorithm of the SparseConvNet backbone path of
 * zhupan007/Detection_3D.  Every function cites the reference file:line it
 * follows (paths relative to /root/reference/SparseConvNet/sparseconvnet/SCN/).
 *
 * Third-party piece restated here: google::dense_hash_map (sparsehash 2.0.x,
 * un-vendored and un-pinned in the reference, absent from this image).  Its
 * published algorithm (densehashtable.h) is restated in dhm_* below; it decides
 * iteration order and therefore rulebook order and coarse-grid numbering.
 *
 * Pinning: this file is checked (tests/test_oracle_vs_ref.py, run in the build
 * container; tests/golden/ fixtures for the GPU box) against the reference's own
 * sources compiled from /root/reference with oracle/shim/google/dense_hash_map
 * (oracle/_ref).  The reference ships NO golden vectors for this path
 * (SURVEY.md section 4), so parity is pinned to reference-code-plus-restated-
 * sparsehash, not to a sparsehash binary.
 */
#include <limits.h>

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t Int; /* Metadata/32bits.h:11 */

/* ------------------------------------------------------------------ hash */
/* IntArrayHash<3>, Metadata/32bits.h:57-66.  `Int hash` wraps mod 2^32; the
 * result converts to size_t with sign extension (irrelevant for < 2^31 buckets). */
uint64_t oracle_point_hash(Int x, Int y, Int z) {
  uint32_t h = 16777619u;
  h *= 2166136261u; h ^= (uint32_t)x;
  h *= 2166136261u; h ^= (uint32_t)y;
  h *= 2166136261u; h ^= (uint32_t)z;
  return (uint64_t)(int64_t)(int32_t)h;
}

/* --------------------------------------------------- dense_hash_map model */
/* sparsehash densehashtable.h: HT_DEFAULT_STARTING_BUCKETS 32, HT_MIN_BUCKETS 4,
 * HT_OCCUPANCY_PCT 50, quadratic probing JUMP_(key,n)=n, grow by rebuilding in
 * ascending bucket order.  Empty key = (INT_MIN,INT_MIN,INT_MIN)
 * (Metadata/Metadata.cpp:17-23). */
typedef struct {
  Int *k;   /* 3 ints per bucket */
  Int *v;
  size_t nb, n;
} DHM;

static size_t dhm_min_buckets(size_t num_elts, size_t wanted) {
  size_t sz = 4;
  while (sz < wanted || num_elts >= (size_t)(sz * 0.5f)) sz *= 2;
  return sz;
}
static void dhm_alloc(DHM *m, size_t nb) {
  m->nb = nb;
  m->k = (Int *)malloc(sizeof(Int) * 3 * nb);
  m->v = (Int *)malloc(sizeof(Int) * nb);
  for (size_t i = 0; i < 3 * nb; i++) m->k[i] = INT_MIN;
}
static void dhm_init(DHM *m) { m->n = 0; dhm_alloc(m, 32); }
static void dhm_free(DHM *m) { free(m->k); free(m->v); m->k = NULL; m->v = NULL; }
static int dhm_is_empty(const DHM *m, size_t b) { return m->k[3 * b] == INT_MIN && m->k[3 * b + 1] == INT_MIN && m->k[3 * b + 2] == INT_MIN; }
/* copy_from(): re-insert every element of `old` in ascending bucket order */
static void dhm_rebuild(DHM *m, size_t nb) {
  DHM old = *m;
  dhm_alloc(m, nb);
  size_t mask = nb - 1;
  for (size_t i = 0; i < old.nb; i++) {
    if (dhm_is_empty(&old, i)) continue;
    size_t probes = 0, b = oracle_point_hash(old.k[3 * i], old.k[3 * i + 1], old.k[3 * i + 2]) & mask;
    while (!dhm_is_empty(m, b)) { ++probes; b = (b + probes) & mask; }
    memcpy(m->k + 3 * b, old.k + 3 * i, 3 * sizeof(Int));
    m->v[b] = old.v[i];
  }
  free(old.k); free(old.v);
}
/* copy constructor: min_buckets(size, 32) buckets, elements in bucket order.
 * Happens when std::vector<SparseGrid> reallocates (IOLayersRules.h:83-84). */
static void dhm_copy_construct(DHM *m) { dhm_rebuild(m, dhm_min_buckets(m->n, 32)); }
/* find_position(); returns bucket or -1, *ins = insert position */
static long dhm_find(const DHM *m, const Int *p, size_t *ins) {
  size_t mask = m->nb - 1, probes = 0, b = oracle_point_hash(p[0], p[1], p[2]) & mask;
  for (;;) {
    if (dhm_is_empty(m, b)) { if (ins) *ins = b; return -1; }
    if (m->k[3 * b] == p[0] && m->k[3 * b + 1] == p[1] && m->k[3 * b + 2] == p[2]) return (long)b;
    ++probes; b = (b + probes) & mask;
  }
}
/* find_or_insert / insert for a key known to be absent: resize_delta(1) then place */
static void dhm_insert_new(DHM *m, const Int *p, Int val) {
  if (!(m->nb >= 4 && m->n + 1 <= (size_t)(m->nb * 0.5f))) {
    size_t needed = dhm_min_buckets(m->n + 1, 0);
    if (needed > m->nb) dhm_rebuild(m, dhm_min_buckets(m->n + 1, m->nb));
  }
  size_t ins;
  dhm_find(m, p, &ins);
  memcpy(m->k + 3 * ins, p, 3 * sizeof(Int));
  m->v[ins] = val;
  m->n++;
}

/* ------------------------------------------------------------- metadata */
typedef struct { DHM mp; Int ctr; } SparseGrid; /* Metadata/Metadata.h:28-33 */
typedef struct {                                  /* std::vector<SparseGrid> */
  SparseGrid *g;
  size_t size, cap;
} SparseGrids;

/* std::vector::resize with libstdc++ growth (_M_default_append): reallocation
 * copy-constructs the existing SparseGrids (sparsehash 2.0.x has no move ctor). */
static void sgs_resize(SparseGrids *s, size_t n) {
  if (n <= s->size) { /* the reference never shrinks a populated vector */ return; }
  if (n > s->cap) {
    size_t add = n - s->size, newcap = s->size + (s->size > add ? s->size : add);
    SparseGrid *ng = (SparseGrid *)malloc(sizeof(SparseGrid) * newcap);
    for (size_t i = 0; i < s->size; i++) { ng[i] = s->g[i]; dhm_copy_construct(&ng[i].mp); }
    free(s->g);
    s->g = ng; s->cap = newcap;
  }
  for (size_t i = s->size; i < n; i++) { dhm_init(&s->g[i].mp); s->g[i].ctr = 0; }
  s->size = n;
}
static void sgs_clear(SparseGrids *s) {
  for (size_t i = 0; i < s->size; i++) dhm_free(&s->g[i].mp);
  free(s->g); s->g = NULL; s->size = s->cap = 0;
}

typedef struct { int nlists; Int **list; long *len; } RuleBook; /* vector<vector<Int>> */
static void rb_init(RuleBook *rb, int nlists) {
  rb->nlists = nlists;
  rb->list = (Int **)calloc(nlists, sizeof(Int *));
  rb->len = (long *)calloc(2 * (size_t)nlists, sizeof(long)); /* len, cap */
}
static void rb_push(RuleBook *rb, int i, Int a) {
  long *len = &rb->len[i], *cap = &rb->len[rb->nlists + i];
  if (*len == *cap) { *cap = *cap ? *cap * 2 : 16; rb->list[i] = (Int *)realloc(rb->list[i], sizeof(Int) * (size_t)*cap); }
  rb->list[i][(*len)++] = a;
}
static void rb_free(RuleBook *rb) {
  for (int i = 0; i < rb->nlists; i++) free(rb->list[i]);
  free(rb->list); free(rb->len); rb->list = NULL; rb->len = NULL; rb->nlists = 0;
}

#define MAXG 32
#define MAXRB 64
typedef struct {
  int ngrids;
  long gsize[MAXG][3];
  SparseGrids grids[MAXG];
  long nActive[MAXG];
  RuleBook inputRules;            /* Metadata.h:56 */
  int nsub; long subkey[MAXRB][6]; RuleBook sub[MAXRB];   /* submanifoldRuleBooks */
  int nrb;  long rbkey[MAXRB][9];  RuleBook rb[MAXRB];    /* ruleBooks */
  int ns2d; long s2dkey[MAXG][3]; RuleBook s2d[MAXG];      /* sparseToDenseRuleBooks */
} OMeta;

OMeta *omd_create(void) { return (OMeta *)calloc(1, sizeof(OMeta)); }
void omd_destroy(OMeta *m) {
  for (int i = 0; i < m->ngrids; i++) sgs_clear(&m->grids[i]);
  if (m->inputRules.nlists) rb_free(&m->inputRules);
  for (int i = 0; i < m->nsub; i++) rb_free(&m->sub[i]);
  for (int i = 0; i < m->ns2d; i++) rb_free(&m->s2d[i]);
  for (int i = 0; i < m->nrb; i++) rb_free(&m->rb[i]);
  free(m);
}
static int omd_grid(OMeta *m, const long *sz) { /* grids[...] operator[] semantics */
  for (int i = 0; i < m->ngrids; i++)
    if (m->gsize[i][0] == sz[0] && m->gsize[i][1] == sz[1] && m->gsize[i][2] == sz[2]) return i;
  int i = m->ngrids++;
  memcpy(m->gsize[i], sz, 3 * sizeof(long));
  return i;
}
long omd_nactive(OMeta *m, const long *sz) { return m->nActive[omd_grid(m, sz)]; }
long omd_batch_size(OMeta *m, const long *sz) { return (long)m->grids[omd_grid(m, sz)].size; }

/* inputLayerRules, Metadata/IOLayersRules.h:18-125 (modes 1..4; mode 0 at :28-58).
 * rules[0] = {mode, maxActive, nInputRows, nOutputRows}; rules[1] = nOut x (1+maxActive). */
long omd_input_layer(OMeta *m, const long *spatial, const int64_t *coords, long nrows, long ncols,
                     long batchSize, long mode) {
  int gi = omd_grid(m, spatial);
  SparseGrids *S = &m->grids[gi];
  long nActive = 0;
  sgs_resize(S, (size_t)batchSize);
  rb_init(&m->inputRules, mode == 0 ? 1 : 2);
  RuleBook *R = &m->inputRules;
  Int p[3];
  if (mode == 0) {
    rb_push(R, 0, (Int)mode); rb_push(R, 0, 1); rb_push(R, 0, (Int)nrows); rb_push(R, 0, (Int)nrows);
    if (ncols == 3) sgs_resize(S, 1);
    for (long i = 0; i < nrows; i++) {
      const int64_t *c = coords + i * ncols;
      p[0] = (Int)c[0]; p[1] = (Int)c[1]; p[2] = (Int)c[2];
      size_t idx = 0;
      if (ncols == 4) { idx = (size_t)c[3]; if (idx + 1 >= S->size) sgs_resize(S, idx + 1); }
      long b = dhm_find(&S->g[idx].mp, p, NULL);
      if (b < 0) dhm_insert_new(&S->g[idx].mp, p, (Int)i); else S->g[idx].mp.v[b] = (Int)i;
    }
    m->nActive[gi] = nrows;
    return nrows;
  }
  /* outputRows: vector<vector<Int>> -> CSR built in two passes over a row->voxel map */
  Int *vox = (Int *)malloc(sizeof(Int) * (size_t)(nrows ? nrows : 1));
  if (ncols == 3) sgs_resize(S, 1);
  for (long i = 0; i < nrows; i++) {
    const int64_t *c = coords + i * ncols;
    p[0] = (Int)c[0]; p[1] = (Int)c[1]; p[2] = (Int)c[2];
    size_t idx = 0;
    if (ncols == 4) { idx = (size_t)c[3]; if (idx + 1 >= S->size) sgs_resize(S, idx + 1); }
    DHM *mp = &S->g[idx].mp;
    long b = dhm_find(mp, p, NULL);
    if (b < 0) { dhm_insert_new(mp, p, (Int)nActive); vox[i] = (Int)nActive++; }
    else vox[i] = mp->v[b];
  }
  Int *cnt = (Int *)calloc((size_t)nActive + 1, sizeof(Int));
  for (long i = 0; i < nrows; i++) cnt[vox[i]]++;
  Int maxActive = 0;
  for (long i = 0; i < nActive; i++) if (cnt[i] > maxActive) maxActive = cnt[i];
  rb_push(R, 0, (Int)mode); rb_push(R, 0, 1); rb_push(R, 0, (Int)nrows); rb_push(R, 0, (Int)nActive);
  if (mode == 1 || mode == 2) { /* :100-110 -- mode 1 keeps FIRST, mode 2 keeps LAST */
    Int *pick = (Int *)malloc(sizeof(Int) * (size_t)(nActive ? nActive : 1));
    for (long i = 0; i < nActive; i++) pick[i] = -1;
    for (long i = 0; i < nrows; i++) if (mode == 2 || pick[vox[i]] < 0) pick[vox[i]] = (Int)i;
    for (long i = 0; i < nActive; i++) { rb_push(R, 1, 1); rb_push(R, 1, pick[i]); }
    free(pick);
  } else { /* modes 3, 4, :111-124 */
    R->list[0][1] = maxActive;
    long w = 1 + maxActive;
    long tot = nActive * w;
    R->list[1] = (Int *)calloc((size_t)(tot ? tot : 1), sizeof(Int));
    R->len[1] = tot; R->len[R->nlists + 1] = tot;
    for (long i = 0; i < nrows; i++) { Int *row = R->list[1] + vox[i] * w; row[1 + row[0]++] = (Int)i; }
  }
  free(cnt); free(vox);
  m->nActive[gi] = nActive;
  return nActive;
}

/* Metadata::getSpatialLocations, Metadata/Metadata.cpp:147-168: out[nActive][4] int64 */
void omd_spatial_locations(OMeta *m, const long *sz, int64_t *out) {
  SparseGrids *S = &m->grids[omd_grid(m, sz)];
  for (size_t i = 0; i < S->size; i++) {
    DHM *mp = &S->g[i].mp;
    for (size_t b = 0; b < mp->nb; b++) {
      if (dhm_is_empty(mp, b)) continue;
      int64_t *o = out + 4 * (int64_t)(mp->v[b] + S->g[i].ctr);
      o[0] = mp->k[3 * b]; o[1] = mp->k[3 * b + 1]; o[2] = mp->k[3 * b + 2]; o[3] = (int64_t)i;
    }
  }
}
/* ids (+ctr) of batch item b in hash-iteration (ascending bucket) order */
long omd_iteration_order(OMeta *m, const long *sz, long b, Int *out) {
  SparseGrids *S = &m->grids[omd_grid(m, sz)];
  if ((size_t)b >= S->size) return 0;
  DHM *mp = &S->g[b].mp;
  long n = 0;
  for (size_t i = 0; i < mp->nb; i++) if (!dhm_is_empty(mp, i)) { if (out) out[n] = mp->v[i] + S->g[b].ctr; n++; }
  return n;
}

/* getSubmanifoldRuleBook -> SubmanifoldConvolution_SgsToRules(_OMP) -> _SgToRules,
 * Metadata/Metadata.cpp:429-443, SubmanifoldConvolutionRules.h:11-87.
 * Region enumeration last-dimension-fastest: RectangularRegions.h:56-71. */
RuleBook *omd_submanifold_rules(OMeta *m, const long *sz, const long *f) {
  for (int i = 0; i < m->nsub; i++)
    if (!memcmp(m->subkey[i], sz, 3 * sizeof(long)) && !memcmp(m->subkey[i] + 3, f, 3 * sizeof(long))) return &m->sub[i];
  int ri = m->nsub++;
  memcpy(m->subkey[ri], sz, 3 * sizeof(long)); memcpy(m->subkey[ri] + 3, f, 3 * sizeof(long));
  RuleBook *R = &m->sub[ri];
  rb_init(R, (int)(f[0] * f[1] * f[2]));
  SparseGrids *S = &m->grids[omd_grid(m, sz)];
  for (size_t gi = 0; gi < S->size; gi++) { /* batch items in order (:59-87 concatenation) */
    DHM *mp = &S->g[gi].mp; Int ctr = S->g[gi].ctr;
    for (size_t b = 0; b < mp->nb; b++) {
      if (dhm_is_empty(mp, b)) continue;
      const Int *o = mp->k + 3 * b;
      Int lb[3], q[3];
      for (int d = 0; d < 3; d++) lb[d] = o[d] - (Int)(f[d] / 2);
      int off = 0;
      for (q[0] = lb[0]; q[0] < lb[0] + f[0]; q[0]++)
        for (q[1] = lb[1]; q[1] < lb[1] + f[1]; q[1]++)
          for (q[2] = lb[2]; q[2] < lb[2] + f[2]; q[2]++, off++) {
            long nb = dhm_find(mp, q, NULL);
            if (nb >= 0) { rb_push(R, off, mp->v[nb] + ctr); rb_push(R, off, mp->v[b] + ctr); }
          }
    }
  }
  return R;
}

/* getRuleBook -> Convolution_InputSgsToRulesAndOutputSgs_OMP ->
 * Convolution_InputSgToRulesAndOutputSg, Metadata/Metadata.cpp:484-510,
 * ConvolutionRules.h:11-34,61-105; OutputRegionCalculator RectangularRegions.h:109-119,
 * InputRegionCalculator :95-105, offset() :30-38. */
RuleBook *omd_conv_rules(OMeta *m, const long *inS, const long *outS, const long *f, const long *s) {
  for (int i = 0; i < m->nrb; i++)
    if (!memcmp(m->rbkey[i], inS, 3 * sizeof(long)) && !memcmp(m->rbkey[i] + 3, f, 3 * sizeof(long)) &&
        !memcmp(m->rbkey[i] + 6, s, 3 * sizeof(long))) return &m->rb[i];
  int ri = m->nrb++;
  memcpy(m->rbkey[ri], inS, 3 * sizeof(long)); memcpy(m->rbkey[ri] + 3, f, 3 * sizeof(long)); memcpy(m->rbkey[ri] + 6, s, 3 * sizeof(long));
  RuleBook *R = &m->rb[ri];
  rb_init(R, (int)(f[0] * f[1] * f[2]));
  int gin = omd_grid(m, inS), gout = omd_grid(m, outS);
  SparseGrids *I = &m->grids[gin], *O = &m->grids[gout];
  sgs_clear(O);
  sgs_resize(O, I->size);
  long out_n = 0;
  for (size_t gi = 0; gi < I->size; gi++) {
    DHM *mp = &I->g[gi].mp; DHM *op = &O->g[gi].mp;
    Int ictr = I->g[gi].ctr;
    Int local = 0; /* oSG.ctr counts from 0 per item; offset added when merging (:79-103) */
    for (size_t b = 0; b < mp->nb; b++) {
      if (dhm_is_empty(mp, b)) continue;
      const Int *p = mp->k + 3 * b;
      Int lb[3], ub[3], j[3];
      for (int d = 0; d < 3; d++) {
        long lo = (p[d] - f[d] + s[d]) / s[d]; /* C++ long division truncates toward zero */
        lb[d] = (Int)(lo > 0 ? lo : 0);
        long hi = p[d] / s[d];
        ub[d] = (Int)(hi < outS[d] - 1 ? hi : outS[d] - 1);
      }
      if (lb[0] > ub[0] || lb[1] > ub[1] || lb[2] > ub[2]) continue;
      for (j[0] = lb[0]; j[0] <= ub[0]; j[0]++)
        for (j[1] = lb[1]; j[1] <= ub[1]; j[1]++)
          for (j[2] = lb[2]; j[2] <= ub[2]; j[2]++) {
            long off = 0, mul = 1;
            for (int d = 2; d >= 0; d--) { off += mul * (p[d] - j[d] * s[d]); mul *= f[d]; }
            long ob = dhm_find(op, j, NULL);
            Int oid;
            if (ob < 0) { oid = local++; dhm_insert_new(op, j, oid); } else oid = op->v[ob];
            rb_push(R, (int)off, mp->v[b] + ictr);
            rb_push(R, (int)off, oid + (Int)out_n);
          }
    }
    O->g[gi].ctr = (Int)out_n;
    out_n += local;
  }
  m->nActive[gout] = out_n;
  return R;
}

/* getSparseToDenseRuleBook -> SparseToDense_InputSgsToRulesAndOutputSgs, Metadata/Metadata.cpp:469-483,
 * ConvolutionRules.h:109-151: one list per batch item of (row = id + ctr, offset = RectangularRegion([0, size-1]).offset(point),
 * last dimension fastest, RectangularRegions.h:30-38) in hash-iteration order. */
RuleBook *omd_sparse_to_dense_rules(OMeta *m, const long *sz) {
  for (int i = 0; i < m->ns2d; i++)
    if (!memcmp(m->s2dkey[i], sz, 3 * sizeof(long))) return &m->s2d[i];
  int ri = m->ns2d++;
  memcpy(m->s2dkey[ri], sz, 3 * sizeof(long));
  RuleBook *R = &m->s2d[ri];
  SparseGrids *S = &m->grids[omd_grid(m, sz)];
  rb_init(R, (int)S->size);
  for (size_t gi = 0; gi < S->size; gi++) {
    DHM *mp = &S->g[gi].mp; Int ctr = S->g[gi].ctr;
    for (size_t b = 0; b < mp->nb; b++) {
      if (dhm_is_empty(mp, b)) continue;
      const Int *p = mp->k + 3 * b;
      rb_push(R, (int)gi, mp->v[b] + ctr);
      rb_push(R, (int)gi, (Int)(((long)p[0] * sz[1] + p[1]) * sz[2] + p[2]));
    }
  }
  return R;
}
/* SparseToDense_ForwardPass / _BackwardPass, CPU/SparseToDense.cpp:7-31: rules = (row, offset) pairs of ONE batch item,
 * out / d_out point at that item's [nPlanes][spatialVolume] block (caller zero-fills, :46). */
void o_sparse_to_dense_forward(const float *in, float *out, long nPlanes, long spatialVolume, const Int *rules, long nHot) {
  for (long s = 0; s < nHot; s++) {
    const float *i = in + (long)rules[2 * s] * nPlanes;
    float *o = out + rules[2 * s + 1];
    for (long c = 0; c < nPlanes; c++) o[c * spatialVolume] = i[c];
  }
}
void o_sparse_to_dense_backward(float *d_in, const float *d_out, long nPlanes, long spatialVolume, const Int *rules, long nHot) {
  for (long s = 0; s < nHot; s++) {
    float *di = d_in + (long)rules[2 * s] * nPlanes;
    const float *o = d_out + rules[2 * s + 1];
    for (long c = 0; c < nPlanes; c++) di[c] = o[c * spatialVolume];
  }
}

RuleBook *omd_input_rules(OMeta *m) { return &m->inputRules; }
long orb_nlists(RuleBook *rb) { return rb->nlists; }
long orb_list_size(RuleBook *rb, long i) { return rb->len[i]; }
void orb_list_copy(RuleBook *rb, long i, Int *dst) { memcpy(dst, rb->list[i], sizeof(Int) * (size_t)rb->len[i]); }

/* ----------------------------------------------------------- compute path */
/* InputLayer_ForwardPass, CPU/IOLayers.cpp:11-29 (output pre-zeroed by caller :66-69) */
void o_input_layer_forward(const float *in, float *out, long nRows, long maxActive, long nPlanes,
                           const Int *rules, int average) {
  for (long row = 0; row < nRows; row++) {
    const Int *r = rules + row * (1 + maxActive);
    Int nA = r[0];
    float mult = (average && nA > 0) ? 1.0f / nA : 1.0f;
    float *o = out + row * nPlanes;
    for (long c = 0; c < nPlanes; c++) o[c] = 0.f;
    for (Int i = 1; i <= nA; i++) {
      const float *f = in + (long)r[i] * nPlanes;
      for (long c = 0; c < nPlanes; c++) o[c] += mult * f[c];
    }
  }
}
/* InputLayer_BackwardPass, CPU/IOLayers.cpp:30-47 */
void o_input_layer_backward(float *d_in, const float *d_out, long nInRows, long nRows, long maxActive,
                            long nPlanes, const Int *rules, int average) {
  memset(d_in, 0, sizeof(float) * (size_t)(nInRows * nPlanes));
  for (long row = 0; row < nRows; row++) {
    const Int *r = rules + row * (1 + maxActive);
    Int nA = r[0];
    float mult = (average && nA > 0) ? 1.0f / nA : 1.0f;
    const float *o = d_out + row * nPlanes;
    for (Int i = 1; i <= nA; i++) {
      float *f = d_in + (long)r[i] * nPlanes;
      for (long c = 0; c < nPlanes; c++) f[c] += mult * o[c];
    }
  }
}

/* One rule list of cpu_{Submanifold,}Convolution_updateOutput / cpu_Deconvolution_updateOutput:
 * rule_index_select (CPU/Convolution.cpp:8-24) -> at::matmul (:74) -> rule_index_add_ (:28-43).
 * src_col/dst_col pick the pair column: conv (0,1) CPU/Convolution.cpp:72-75,
 * deconv (1,0) CPU/Deconvolution.cpp:33-36.  groups == 1 (all shipped configs).
 * out must be pre-zeroed (or bias-filled) by the caller (:55-58). */
void o_conv_list_forward(const float *in, float *out, const float *w, long nIn, long nOut,
                         const Int *rules, long nRules, int src_col, int dst_col) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < nRules; i++) {
    const float *a = in + (long)rules[2 * i + src_col] * nIn;
    float *o = out + (long)rules[2 * i + dst_col] * nOut; /* a dst row occurs at most once per list */
    for (long c = 0; c < nIn; c++) {
      float av = a[c];
      const float *wr = w + c * nOut;
      for (long q = 0; q < nOut; q++) o[q] += av * wr[q];
    }
  }
}
/* One rule list of cpu_*_backward (CPU/Convolution.cpp:96-113, CPU/Deconvolution.cpp:60-76):
 * dW[k] = gather(in)^T @ gather(dOut)   (matmul_out overwrites dw)
 * dIn  += scatter(gather(dOut) @ W[k]^T)                                        */
void o_conv_list_backward(const float *in, float *d_in, const float *d_out, const float *w, float *dw,
                          long nIn, long nOut, const Int *rules, long nRules, int src_col, int dst_col) {
  memset(dw, 0, sizeof(float) * (size_t)(nIn * nOut));
  for (long i = 0; i < nRules; i++) {
    const float *a = in + (long)rules[2 * i + src_col] * nIn;
    const float *g = d_out + (long)rules[2 * i + dst_col] * nOut;
    float *da = d_in + (long)rules[2 * i + src_col] * nIn;
    for (long c = 0; c < nIn; c++) {
      const float *wr = w + c * nOut;
      float *dwr = dw + c * nOut;
      float acc = 0.f, av = a[c];
      for (long q = 0; q < nOut; q++) { acc += g[q] * wr[q]; dwr[q] += av * g[q]; }
      da[c] += acc;
    }
  }
}

/* BatchNormalization_ForwardPass, CPU/BatchNormalization.cpp:12-62 (stride == nPlanes) */
void o_bn_forward(const float *in, float *out, long nPlanes, long nActive, float *saveMean,
                  float *saveInvStd, float *runningMean, float *runningVar, const float *weight,
                  const float *bias, float eps, float momentum, int train, float leakiness) {
  if (train) {
    memset(saveMean, 0, sizeof(float) * (size_t)nPlanes);
    memset(saveInvStd, 0, sizeof(float) * (size_t)nPlanes);
    for (long r = 0; r < nActive; r++)
      for (long c = 0; c < nPlanes; c++) { float v = in[r * nPlanes + c]; saveMean[c] += v; saveInvStd[c] += v * v; }
    for (long c = 0; c < nPlanes; c++) {
      saveMean[c] /= nActive;
      runningMean[c] = momentum * runningMean[c] + (1 - momentum) * saveMean[c];
      saveInvStd[c] -= saveMean[c] * saveMean[c] * nActive;
      runningVar[c] = momentum * runningVar[c] + (1 - momentum) * saveInvStd[c] / (nActive - 1);
      saveInvStd[c] = powf(saveInvStd[c] / nActive + eps, -0.5f);
    }
  } else {
    for (long c = 0; c < nPlanes; c++) { saveMean[c] = runningMean[c]; saveInvStd[c] = powf(runningVar[c] + eps, -0.5f); }
  }
  float *w = (float *)malloc(sizeof(float) * (size_t)nPlanes), *b = (float *)malloc(sizeof(float) * (size_t)nPlanes);
  for (long c = 0; c < nPlanes; c++) {
    w[c] = saveInvStd[c] * (weight ? weight[c] : 1);
    b[c] = -saveMean[c] * w[c] + (bias ? bias[c] : 0);
  }
  for (long r = 0; r < nActive; r++)
    for (long c = 0; c < nPlanes; c++) {
      float o = in[r * nPlanes + c] * w[c] + b[c];
      out[r * nPlanes + c] = o * ((o > 0) ? 1.f : leakiness);
    }
  free(w); free(b);
}
/* BatchNormalization_BackwardPass, CPU/BatchNormalization.cpp:64-107 (d_out rewritten in place) */
void o_bn_backward(const float *in, float *d_in, const float *out, float *d_out, long nPlanes, long nActive,
                   const float *saveMean, const float *saveInvStd, const float *weight, float *d_weight,
                   float *d_bias, float leakiness) {
  float *gm = (float *)calloc((size_t)nPlanes, sizeof(float)), *dp = (float *)calloc((size_t)nPlanes, sizeof(float)),
        *k = (float *)calloc((size_t)nPlanes, sizeof(float));
  for (long r = 0; r < nActive; r++)
    for (long c = 0; c < nPlanes; c++) {
      float d = d_out[r * nPlanes + c] * ((out[r * nPlanes + c] > 0) ? 1.f : leakiness);
      d_out[r * nPlanes + c] = d;
      gm[c] += d;
      dp[c] += (in[r * nPlanes + c] - saveMean[c]) * d;
    }
  for (long c = 0; c < nPlanes; c++) {
    if (d_bias) d_bias[c] = gm[c];
    gm[c] /= nActive;
    k[c] = dp[c] * saveInvStd[c] * saveInvStd[c] / nActive;
  }
  for (long r = 0; r < nActive; r++)
    for (long c = 0; c < nPlanes; c++)
      d_in[r * nPlanes + c] = (d_out[r * nPlanes + c] - gm[c] - (in[r * nPlanes + c] - saveMean[c]) * k[c]) *
                              saveInvStd[c] * (weight ? weight[c] : 1);
  if (d_weight) for (long c = 0; c < nPlanes; c++) d_weight[c] = dp[c] * saveInvStd[c];
  free(gm); free(dp); free(k);
}

/* ----------------------------------------------------------- ROIAlignRotated3D (maskrcnn_benchmark/csrc/cuda/ROIAlignRotated3D_cuda.cu)
 * Restatement of the reference's CUDA kernels as plain loops, T = float as in the only instantiation the reference uses:
 * bilinear_interpolate :16-86, RoIAlignRotated3DForward :89-172, bilinear_interpolate_gradient :176-232,
 * RoIAlignRotated3DBackwardFeature :235-346.  Checked against the reference's own kernels (compiled unmodified into
 * oracle/_ref/libroialign3d_ref.so) by the GPU parity tests. */
static float roi3d_interp(const float *d, int H, int W, int Z, float y, float x, float z) {
  if (y < -1.0f || y > H || x < -1.0f || x > W || z < -1.0f) return 0.f; /* :27 (`zsize > zsize` never rejects) */
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  if (z <= 0) z = 0;
  int yl = (int)y, xl = (int)x, zl = (int)z, yh, xh, zh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
  if (zl >= Z - 1) { zh = zl = Z - 1; z = (float)zl; } else zh = zl + 1;
  float ly = y - yl, lx = x - xl, lz = z - zl, hy = 1.f - ly, hx = 1.f - lx, hz = 1.f - lz;
  float v1 = d[(yl * W + xl) * Z + zl], v2 = d[(yl * W + xh) * Z + zl], v3 = d[(yh * W + xl) * Z + zl], v4 = d[(yh * W + xh) * Z + zl];
  float v5 = d[(yl * W + xl) * Z + zh], v6 = d[(yl * W + xh) * Z + zh], v7 = d[(yh * W + xl) * Z + zh], v8 = d[(yh * W + xh) * Z + zh];
  float w1 = hy * hx * hz, w2 = hy * lx * hz, w3 = ly * hx * hz, w4 = ly * lx * hz, w5 = hy * hx * lz, w6 = hy * lx * lz, w7 = ly * hx * lz, w8 = ly * lx * lz;
  return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4 + w5 * v5 + w6 * v6 + w7 * v7 + w8 * v8;
}
/* backward == 0: out[n][C][PH][PW][PZ] from input[B][C][H][W][Z];  backward != 0: d_input (zeroed by the caller) += from top_diff = out */
void o_roi_align_rotated_3d(const float *input, float *d_input, long C, long H, long W, long Z, const float *rois, long n_rois, float scale, long PH, long PW,
                            long PZ, long sampling, float *out, int backward) {
  for (long n = 0; n < n_rois; n++) {
    const float *r = rois + n * 8;
    int bi = (int)r[0];
    float cw = r[1] * scale, ch = r[2] * scale, cz = r[3] * scale, rw = r[4] * scale, rh = r[5] * scale, rz = r[6] * scale;
    float theta = (float)(r[7] * 3.14159265358979323846 / 180.0);
    rw = rw > 1.f ? rw : 1.f; rh = rh > 1.f ? rh : 1.f; rz = rz > 1.f ? rz : 1.f;
    float bh = rh / (float)PH, bw = rw / (float)PW, bz = rz / (float)PZ;
    int gh = sampling > 0 ? (int)sampling : (int)ceilf(rh / PH), gw = sampling > 0 ? (int)sampling : (int)ceilf(rw / PW), gz = sampling > 0 ? (int)sampling : (int)ceilf(rz / PZ);
    float sh = -rh / 2.0f, sw = -rw / 2.0f, sz = -rz / 2.0f, cosT = cosf(theta), sinT = sinf(theta), count = (float)(gh * gw * gz);
    for (long c = 0; c < C; c++) {
      const float *d = input ? input + ((long)bi * C + c) * H * W * Z : NULL;
      float *dd = d_input ? d_input + ((long)bi * C + c) * H * W * Z : NULL;
      for (long ph = 0; ph < PH; ph++) for (long pw = 0; pw < PW; pw++) for (long pz = 0; pz < PZ; pz++) {
        float *o = out + (((n * C + c) * PH + ph) * PW + pw) * PZ + pz;
        float val = 0.f;
        for (int iy = 0; iy < gh; iy++) {
          float yy = sh + ph * bh + (iy + .5f) * bh / (float)gh;
          for (int ix = 0; ix < gw; ix++) {
            float xx = sw + pw * bw + (ix + .5f) * bw / (float)gw;
            for (int iz = 0; iz < gz; iz++) {
              float zz = sz + pz * bz + (iz + .5f) * bz / (float)gz;
              float x = xx * cosT + yy * sinT + cw, y = yy * cosT - xx * sinT + ch, z = zz + cz;
              if (!backward) { val += roi3d_interp(d, (int)H, (int)W, (int)Z, y, x, z); continue; }
              if (y < -1.0f || y > H || x < -1.0f || x > W || z < -1.0f || z > Z) continue; /* :184 (the gradient variant does test z > zsize) */
              if (y <= 0) y = 0;
              if (x <= 0) x = 0;
              if (z <= 0) z = 0;
              int yl = (int)y, xl = (int)x, zl = (int)z, yh, xh, zh;
              if (yl >= H - 1) { yh = yl = (int)H - 1; y = (float)yl; } else yh = yl + 1;
              if (xl >= W - 1) { xh = xl = (int)W - 1; x = (float)xl; } else xh = xl + 1;
              if (zl >= Z - 1) { zh = zl = (int)Z - 1; z = (float)zl; } else zh = zl + 1;
              float ly = y - yl, lx = x - xl, lz = z - zl, hy = 1.f - ly, hx = 1.f - lx, hz = 1.f - lz, t = *o;
              dd[(yl * W + xl) * Z + zl] += t * (hy * hx * hz) / count; dd[(yl * W + xh) * Z + zl] += t * (hy * lx * hz) / count;
              dd[(yh * W + xl) * Z + zl] += t * (ly * hx * hz) / count; dd[(yh * W + xh) * Z + zl] += t * (ly * lx * hz) / count;
              dd[(yl * W + xl) * Z + zh] += t * (hy * hx * lz) / count; dd[(yl * W + xh) * Z + zh] += t * (hy * lx * lz) / count;
              dd[(yh * W + xl) * Z + zh] += t * (ly * hx * lz) / count; dd[(yh * W + xh) * Z + zh] += t * (ly * lx * lz) / count;
            }
          }
        }
        if (!backward) *o = val / count;
      }
    }
  }
}
