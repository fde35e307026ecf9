/*
 * oracle/bidi_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the reference's per-timestep hierarchy update
 * (222464/BIDInet, BIDInet/source/{sdr,neo}); each function cites the reference
 * file:line it follows.  It keeps the reference's serial structure and its fp32
 * expression order (build with -O2 -ffp-contract=off, no -march: the x86-64
 * reference build has no FMA), but stores the state structure-of-arrays in the
 * serialised order of include/bidinet_b200.h, so a state array is one memcpy.
 *
 * Pinning: the reference holds no tests or golden vectors for this path (SURVEY
 * section 4), so this file is pinned against the reference ITSELF compiled here
 * (oracle/_ref/libbidiref.so, tests/test_oracle_vs_ref.py: bit-exact state after
 * every step for all four variants) and against tests/golden/ fixtures generated
 * from it by tests/golden/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * legs may call this; the product never does.
 */
#include "bidi_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ======================= std::mt19937 and libstdc++ distributions ============== */
/* The reference draws every weight and every exploration sample from one
 * caller-owned std::mt19937 (RunnerMain.cpp:36).  libstdc++ 13 semantics:
 *   generate_canonical<float,24>: ONE engine draw, float(draw) / 2^32, clamped
 *     below 1 (bits/random.tcc);
 *   uniform_real_distribution<float>(a,b): canonical*(b-a)+a;
 *   normal_distribution<float>: Marsaglia polar, pair cache inside the object. */
typedef struct { uint32_t mt[624]; int idx; } mt19937;

static void mt_seed(mt19937* g, uint32_t s) {
  g->mt[0] = s;
  for (int i = 1; i < 624; i++) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}
static uint32_t mt_next(mt19937* g) {
  if (g->idx >= 624) {
    for (int k = 0; k < 624; k++) {
      uint32_t y = (g->mt[k] & 0x80000000u) | (g->mt[(k + 1) % 624] & 0x7fffffffu);
      g->mt[k] = g->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}
static float canonical(mt19937* g) {
  float r = (float)mt_next(g) / 4294967296.0f;
  if (r >= 1.0f) r = nextafterf(1.0f, 0.0f);
  return r;
}
static float uniform_real(mt19937* g, float a, float b) { return canonical(g) * (b - a) + a; }
typedef struct { int saved_available; float saved; } normal_state;
static float normal_real(mt19937* g, normal_state* st, float mean, float stddev) {
  float ret;
  if (st->saved_available) { st->saved_available = 0; ret = st->saved; }
  else {
    float x, y, r2;
    do {
      x = 2.0f * canonical(g) - 1.0f;
      y = 2.0f * canonical(g) - 1.0f;
      r2 = x * x + y * y;
    } while (r2 > 1.0f || r2 == 0.0f);
    float mult = sqrtf(-2.0f * logf(r2) / r2);
    st->saved = x * mult; st->saved_available = 1;
    ret = y * mult;
  }
  return ret * stddev + mean;
}

static float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); } /* sdr/RSDR.h:55-57 */
/* std::max / std::min semantics (first argument wins ties, so signed zeros match) */
static float smax(float a, float b) { return (a < b) ? b : a; }
static float smin(float a, float b) { return (b < a) ? b : a; }
static float clamp01(float x) { return smin(1.0f, smax(0.0f, x)); }

/* ======================= storage ============================================== */
/* A clipped, per-unit connection list in compressed-row form: unit i owns
 * entries off[i]..off[i+1]-1, in the reference's dx-outer / dy-inner order. */
typedef struct { int n_units; int* off; int* idx; float* w; float* tr; } conns;

typedef struct {
  int vw, vh, hw, hh, nv, nh;
  float *vis_input, *vis_recon;
  float *thr, *state, *state_prev, *act, *recon, *spike, *spike_prev;
  conns ff, rec, lat; /* lat.w == NULL for the k-winners coder (index-only list) */
} coder;

/* All columns of one node layer, pooled so every array is contiguous. */
typedef struct {
  int n, cells;
  int na;       /* action entries per node: 3 (neo::Column) or 2*(3 + recurrent inputs) (deep::SDRRL: actions, then anti-actions) */
  float *a_err, *cell_as, *cell_ae, *avg_surprise; /* deep::SDRRL only: Action::_error [node][na]; Cell::_actionState/_actionError [node][cell]; [node] */
  int* in_off;  /* [n+1] prefix sums of nIn */
  float *inputs, *recon_err;            /* [sum nIn] */
  float* ff_w;                          /* [node][cell][input] at cells*in_off[node] */
  float* lat_w;                         /* [node][cell][cell] */
  float *thr, *act, *spike, *spike_prev, *state; /* [node][cell] */
  float *q_w, *q_tr;                    /* [node][cell] */
  float *a_state, *a_expl;              /* [node][3] */
  float *a_w, *a_tr;                    /* [node][3][cell] */
  float* prev_value;                    /* [node] */
} columns;

typedef struct {
  int n;
  float *state, *state_prev, *act, *act_prev, *baseline;
  float* local_reward; /* deep::CSRL only */
  conns fb, pred;
  columns col;         /* neo::Column per node, or deep::SDRRL per node (deep::CSRL) */
} pnodes;

/* sdr::QPRSDR (sdr/QPRSDR.h:10-75): a sparse Q network over the hierarchy's grids; only
 * connections on a path from an action input are kept (sdr/QPRSDR.cpp:72-81). */
typedef struct {
  int n;
  float *state, *err, *bias_w, *bias_tr;
  conns ff; /* kept connections; idx = input cell (layer 0) or node of the layer below */
} qlayer;
typedef struct {
  int half, n_nodes, n_top; /* actions; action nodes = 2*half; nodes of the top layer */
  qlayer ql[BIDI_MAX_LAYERS];
  float *qc_w, *qc_tr;
  float *a_pred, *a_derive, *a_expl, *a_err;
  int *a_input, *act_index; /* node -> input cell; input cell -> action node or -1 */
  float prev_value[1];
} qnet;

typedef struct {
  bidi_config cfg;
  bidi_layer_desc ld[BIDI_MAX_LAYERS];
  int L, n_in;
  qnet q;
  coder cod[BIDI_MAX_LAYERS];
  pnodes pn[BIDI_MAX_LAYERS + 1]; /* pn[L] = input prediction nodes */
  float* prediction;              /* k-winners only */
  float* reward_offsets;          /* deep::CSRL only: _lastLayerRewardOffsets [units of the top layer] */
  mt19937 gen;
} oracle;

static float* fzeros(size_t n) { return (float*)calloc(n ? n : 1, sizeof(float)); }

/* Window of `radius` around (cx,cy) clipped to a tw x th grid, dx outer / dy inner
 * (sdr/RSDR.cpp:48-62).  Two passes: count, then fill. */
static void build_conns(conns* c, int n_units, int uw, int tw, int th, int radius,
                        float ratio_x, float ratio_y, int identity, int exclude_centre, int weighted, int traced) {
  c->n_units = n_units;
  c->off = (int*)calloc((size_t)n_units + 1, sizeof(int));
  for (int pass = 0; pass < 2; pass++) {
    int k = 0;
    for (int i = 0; i < n_units; i++) {
      int ux = i % uw, uy = i / uw;
      int cx = identity ? ux : (int)roundf((float)ux * ratio_x); /* sdr/RSDR.cpp:40-41 */
      int cy = identity ? uy : (int)roundf((float)uy * ratio_y);
      if (pass == 0) c->off[i] = k;
      if (radius >= 0)
        for (int dx = -radius; dx <= radius; dx++)
          for (int dy = -radius; dy <= radius; dy++) {
            if (exclude_centre && dx == 0 && dy == 0) continue;
            int tx = cx + dx, ty = cy + dy;
            if (tx >= 0 && tx < tw && ty >= 0 && ty < th) {
              if (pass == 1) c->idx[k] = tx + ty * tw;
              k++;
            }
          }
    }
    if (pass == 0) {
      c->off[n_units] = k;
      c->idx = (int*)calloc((size_t)k + 1, sizeof(int));
      c->w = weighted ? fzeros((size_t)k) : NULL;
      c->tr = traced ? fzeros((size_t)k) : NULL;
    }
  }
}
static void free_conns(conns* c) { free(c->off); free(c->idx); free(c->w); free(c->tr); }

static void alloc_columns(columns* c, int n, int cells, const int* n_in) {
  c->n = n; c->cells = cells; c->na = 3;
  c->in_off = (int*)calloc((size_t)n + 1, sizeof(int));
  for (int i = 0; i < n; i++) c->in_off[i + 1] = c->in_off[i] + n_in[i];
  size_t tot = (size_t)c->in_off[n], nc = (size_t)n * cells;
  c->inputs = fzeros(tot); c->recon_err = fzeros(tot);
  c->ff_w = fzeros(tot * cells); c->lat_w = fzeros(nc * cells);
  c->thr = fzeros(nc); c->act = fzeros(nc); c->spike = fzeros(nc); c->spike_prev = fzeros(nc); c->state = fzeros(nc);
  c->q_w = fzeros(nc); c->q_tr = fzeros(nc);
  c->a_state = fzeros((size_t)n * 3); c->a_expl = fzeros((size_t)n * 3);
  c->a_w = fzeros(nc * 3); c->a_tr = fzeros(nc * 3);
  c->prev_value = fzeros((size_t)n);
}
static void free_columns(columns* c) {
  free(c->in_off); free(c->inputs); free(c->recon_err); free(c->ff_w); free(c->lat_w);
  free(c->thr); free(c->act); free(c->spike); free(c->spike_prev); free(c->state);
  free(c->q_w); free(c->q_tr); free(c->a_state); free(c->a_expl); free(c->a_w); free(c->a_tr); free(c->prev_value);
}

/* ======================= createRandom ========================================= */
/* neo/Column.cpp:7-44: per cell nIn ff draws, nCells lateral draws, one q draw;
 * then per action nCells draws. */
static void column_create_random(oracle* o, columns* c, int node, float thr) {
  const bidi_config* g = &o->cfg;
  int nin = c->in_off[node + 1] - c->in_off[node], cells = c->cells;
  float* ff = c->ff_w + (size_t)cells * c->in_off[node];
  for (int i = 0; i < cells; i++) {
    c->thr[node * cells + i] = thr;
    for (int j = 0; j < nin; j++) ff[i * nin + j] = uniform_real(&o->gen, g->init_min_w, g->init_max_w);
    for (int j = 0; j < cells; j++) c->lat_w[((size_t)node * cells + i) * cells + j] = uniform_real(&o->gen, g->init_min_inh, g->init_max_inh);
    c->q_w[node * cells + i] = uniform_real(&o->gen, g->init_min_w, g->init_max_w);
  }
  for (int a = 0; a < 3; a++)
    for (int k = 0; k < cells; k++) c->a_w[((size_t)node * 3 + a) * cells + k] = uniform_real(&o->gen, g->init_min_w, g->init_max_w);
}

/* sdr/QPRSDR.cpp:8-97, after _prsdr.createRandom has consumed its draws. */
static void q_create_random(oracle* o) {
  const bidi_config* g = &o->cfg;
  qnet* q = &o->q;
  const int L = o->L;
  q->half = g->q_n_actions; q->n_nodes = 2 * q->half;
  q->act_index = (int*)calloc((size_t)o->n_in, sizeof(int));
  for (int i = 0; i < o->n_in; i++) q->act_index[i] = -1;
  for (int i = 0; i < q->half; i++) q->act_index[g->q_action_idx[i]] = i; /* :13-20 (anti indices are never looked up) */
  q->n_top = o->ld[L - 1].width * o->ld[L - 1].height;
  q->qc_w = fzeros(q->n_top); q->qc_tr = fzeros(q->n_top);
  for (int i = 0; i < q->n_top; i++) q->qc_w[i] = uniform_real(&o->gen, g->init_min_w, g->init_max_w) / (float)q->n_top; /* :30-31 */
  int wp = g->in_width, hp = g->in_height;
  for (int l = 0; l < L; l++) {
    const bidi_layer_desc* d = &o->ld[l];
    qlayer* ql = &q->ql[l];
    ql->n = d->width * d->height;
    ql->state = fzeros(ql->n); ql->err = fzeros(ql->n); ql->bias_w = fzeros(ql->n); ql->bias_tr = fzeros(ql->n);
    const float rx = (float)wp / (float)d->width, ry = (float)hp / (float)d->height;
    /* worst case: every in-bounds slot kept */
    const int span = 2 * d->r_ff + 1;
    ql->ff.n_units = ql->n;
    ql->ff.off = (int*)calloc((size_t)ql->n + 1, sizeof(int));
    ql->ff.idx = (int*)calloc((size_t)ql->n * span * span + 1, sizeof(int));
    ql->ff.w = fzeros((size_t)ql->n * span * span); ql->ff.tr = fzeros((size_t)ql->n * span * span);
    int k = 0;
    for (int qi = 0; qi < ql->n; qi++) {
      ql->ff.off[qi] = k;
      ql->bias_w[qi] = uniform_real(&o->gen, g->init_min_w, g->init_max_w); /* :50 */
      const int hx = qi % d->width, hy = qi / d->width;
      const int cx = (int)roundf((float)hx * rx), cy = (int)roundf((float)hy * ry);
      for (int dx = -d->r_ff; dx <= d->r_ff; dx++)
        for (int dy = -d->r_ff; dy <= d->r_ff; dy++) {
          const int tx = cx + dx, ty = cy + dy;
          if (tx >= 0 && tx < wp && ty >= 0 && ty < hp) {
            const int t = tx + ty * wp;
            const float w = uniform_real(&o->gen, g->init_min_w, g->init_max_w); /* drawn whether kept or not, :70 */
            const int keep = l == 0 ? q->act_index[t] != -1 : q->ql[l - 1].ff.off[t + 1] > q->ql[l - 1].ff.off[t];
            if (keep) { ql->ff.idx[k] = t; ql->ff.w[k] = w; k++; }
          }
        }
    }
    ql->ff.off[ql->n] = k;
    wp = d->width; hp = d->height;
  }
  q->a_pred = fzeros(q->n_nodes); q->a_derive = fzeros(q->n_nodes); q->a_expl = fzeros(q->n_nodes); q->a_err = fzeros(q->n_nodes);
  q->a_input = (int*)calloc((size_t)q->n_nodes + 1, sizeof(int));
  for (int i = 0; i < q->half; i++) { q->a_input[i] = g->q_action_idx[i]; q->a_input[q->half + i] = g->q_anti_idx[i]; }
  q->prev_value[0] = 0.0f; /* _prevValue is not initialised by the reference (sdr/QPRSDR.h:66); 0 here and in the harness */
}

static void csrl_create(oracle* o);
void* orc_create(const bidi_config* cfg, const bidi_layer_desc* layers, unsigned seed) {
  if (cfg->n_layers < 1 || cfg->n_layers > BIDI_MAX_LAYERS) return NULL;
  oracle* o = (oracle*)calloc(1, sizeof(oracle));
  o->cfg = *cfg; o->L = cfg->n_layers; o->n_in = cfg->in_width * cfg->in_height;
  memcpy(o->ld, layers, sizeof(bidi_layer_desc) * (size_t)o->L);
  mt_seed(&o->gen, seed);
  if (cfg->variant == BIDI_CSRL) { csrl_create(o); return o; }
  const int V = cfg->variant, L = o->L;
  const int kw = V == BIDI_KWINNERS, agent = V == BIDI_NEO_AGENT;
  const int traced = V == BIDI_NEO_HIERARCHY || V == BIDI_NEO_AGENT;
  const float wlo = cfg->init_min_w, whi = cfg->init_max_w, ilo = cfg->init_min_inh, ihi = cfg->init_max_inh;

  /* structure first (no draws), then the draws in the reference's order */
  int vw = cfg->in_width, vh = cfg->in_height;
  for (int l = 0; l < L; l++) {
    const bidi_layer_desc* d = &o->ld[l];
    coder* c = &o->cod[l];
    c->vw = vw; c->vh = vh; c->hw = d->width; c->hh = d->height; c->nv = vw * vh; c->nh = d->width * d->height;
    c->vis_input = fzeros(c->nv); c->vis_recon = fzeros(c->nv);
    c->thr = fzeros(c->nh); c->state = fzeros(c->nh); c->state_prev = fzeros(c->nh); c->act = fzeros(c->nh);
    c->recon = fzeros(c->nh); c->spike = fzeros(c->nh); c->spike_prev = fzeros(c->nh);
    for (int i = 0; i < c->nh; i++) c->thr[i] = cfg->init_threshold;
    float rx = (float)vw / (float)d->width, ry = (float)vh / (float)d->height; /* sdr/RSDR.cpp:33-34 */
    build_conns(&c->ff, c->nh, c->hw, vw, vh, d->r_ff, rx, ry, 0, 0, 1, traced);
    build_conns(&c->rec, c->nh, c->hw, c->hw, c->hh, d->r_rec, 1, 1, 1, 1, 1, traced); /* sdr/RSDR.cpp:92-116: -1 => none */
    build_conns(&c->lat, c->nh, c->hw, c->hw, c->hh, d->r_lat, 1, 1, 1, 1, !kw, 0);
    pnodes* p = &o->pn[l];
    p->n = c->nh;
    p->state = fzeros(p->n); p->state_prev = fzeros(p->n); p->act = fzeros(p->n); p->act_prev = fzeros(p->n); p->baseline = fzeros(p->n);
    if (l < L - 1) { /* sdr/PredictiveRSDR.cpp:31-34,46-69 */
      float fx = (float)o->ld[l + 1].width / (float)d->width, fy = (float)o->ld[l + 1].height / (float)d->height;
      build_conns(&p->fb, p->n, d->width, o->ld[l + 1].width, o->ld[l + 1].height, d->r_fb, fx, fy, 0, 0, 1, 0);
    } else build_conns(&p->fb, p->n, d->width, 1, 1, -1, 1, 1, 1, 0, 1, 0);
    build_conns(&p->pred, p->n, d->width, d->width, d->height, d->r_pred, 1, 1, 1, 0, 1, 0); /* centre included, :72-91 */
    if (agent) { /* neo/Agent.cpp:90: nIn = |pred| + 2|fb| */
      int* nin = (int*)calloc((size_t)p->n, sizeof(int));
      for (int i = 0; i < p->n; i++) nin[i] = (p->pred.off[i + 1] - p->pred.off[i]) + 2 * (p->fb.off[i + 1] - p->fb.off[i]);
      alloc_columns(&p->col, p->n, d->col_cells, nin);
      free(nin);
    }
    vw = d->width; vh = d->height;
  }
  if (kw) o->prediction = fzeros(o->n_in);
  else { /* input prediction nodes, sdr/IPredictiveRSDR.cpp:96-135 */
    pnodes* p = &o->pn[L];
    p->n = o->n_in;
    p->state = fzeros(p->n); p->state_prev = fzeros(p->n); p->act = fzeros(p->n); p->act_prev = fzeros(p->n); p->baseline = fzeros(p->n);
    float fx = (float)o->ld[0].width / (float)cfg->in_width, fy = (float)o->ld[0].height / (float)cfg->in_height;
    build_conns(&p->fb, p->n, cfg->in_width, o->ld[0].width, o->ld[0].height, cfg->r_infb, fx, fy, 0, 0, 1, 0);
    build_conns(&p->pred, p->n, cfg->in_width, 1, 1, -1, 1, 1, 1, 0, 1, 0);
    if (agent) { /* neo/Agent.cpp:137: nIn = 2|fb| */
      int* nin = (int*)calloc((size_t)p->n, sizeof(int));
      for (int i = 0; i < p->n; i++) nin[i] = 2 * (p->fb.off[i + 1] - p->fb.off[i]);
      alloc_columns(&p->col, p->n, cfg->col_cells, nin);
      free(nin);
    }
  }

  /* draws: SURVEY App. C */
  for (int l = 0; l < L; l++) {
    coder* c = &o->cod[l];
    for (int hi = 0; hi < c->nh; hi++) {
      for (int k = c->ff.off[hi]; k < c->ff.off[hi + 1]; k++) c->ff.w[k] = uniform_real(&o->gen, wlo, whi);
      if (kw) { /* sdr/RSDR.cpp: ff, (lateral index-only), recurrent */
        for (int k = c->rec.off[hi]; k < c->rec.off[hi + 1]; k++) c->rec.w[k] = uniform_real(&o->gen, wlo, whi);
      } else { /* sdr/IRSDR.cpp:50-121: ff, recurrent, lateral(inhibitionDist) */
        for (int k = c->rec.off[hi]; k < c->rec.off[hi + 1]; k++) c->rec.w[k] = uniform_real(&o->gen, wlo, whi);
        for (int k = c->lat.off[hi]; k < c->lat.off[hi + 1]; k++) c->lat.w[k] = uniform_real(&o->gen, ilo, ihi);
      }
    }
    pnodes* p = &o->pn[l];
    for (int pi = 0; pi < p->n; pi++) {
      (void)uniform_real(&o->gen, wlo, whi); /* p._bias: drawn, never used (sdr/PredictiveRSDR.cpp:39) */
      for (int k = p->fb.off[pi]; k < p->fb.off[pi + 1]; k++) p->fb.w[k] = uniform_real(&o->gen, wlo, whi);
      for (int k = p->pred.off[pi]; k < p->pred.off[pi + 1]; k++) p->pred.w[k] = uniform_real(&o->gen, wlo, whi);
      if (agent) column_create_random(o, &p->col, pi, cfg->init_threshold);
    }
  }
  if (!kw) {
    pnodes* p = &o->pn[L];
    for (int pi = 0; pi < p->n; pi++) {
      (void)uniform_real(&o->gen, wlo, whi);
      for (int k = p->fb.off[pi]; k < p->fb.off[pi + 1]; k++) p->fb.w[k] = uniform_real(&o->gen, wlo, whi);
      if (agent) column_create_random(o, &p->col, pi, cfg->init_threshold);
    }
  }
  if (V == BIDI_QPRSDR) q_create_random(o);
  return o;
}

void orc_destroy(void* h) {
  oracle* o = (oracle*)h;
  if (!o) return;
  if (o->cfg.variant == BIDI_QPRSDR) {
    qnet* q = &o->q;
    for (int l = 0; l < o->L; l++) { qlayer* ql = &q->ql[l]; free(ql->state); free(ql->err); free(ql->bias_w); free(ql->bias_tr); free_conns(&ql->ff); }
    free(q->qc_w); free(q->qc_tr); free(q->a_pred); free(q->a_derive); free(q->a_expl); free(q->a_err); free(q->a_input); free(q->act_index);
  }
  for (int l = 0; l <= o->L; l++) {
    if (l < o->L) {
      coder* c = &o->cod[l];
      free(c->vis_input); free(c->vis_recon); free(c->thr); free(c->state); free(c->state_prev); free(c->act);
      free(c->recon); free(c->spike); free(c->spike_prev);
      free_conns(&c->ff); free_conns(&c->rec); free_conns(&c->lat);
    }
    pnodes* p = &o->pn[l];
    if (!p->n) continue;
    free(p->state); free(p->state_prev); free(p->act); free(p->act_prev); free(p->baseline);
    free_conns(&p->fb); free_conns(&p->pred);
    if (o->cfg.variant == BIDI_NEO_AGENT) free_columns(&p->col);
  }
  free(o->prediction);
  free(o);
}

/* ======================= k-winners coder: sdr::RSDR =========================== */
/* sdr/RSDR.cpp:148-163 (and the second half of activate, :135-145) */
static void rsdr_inhibit(const coder* c, float sparsity, const float* activations, float* states) {
  for (int hi = 0; hi < c->nh; hi++) {
    int n = c->lat.off[hi + 1] - c->lat.off[hi];
    float num_active = sparsity * (float)n;
    float inhibition = 0.0f;
    for (int k = c->lat.off[hi]; k < c->lat.off[hi + 1]; k++)
      inhibition += activations[c->lat.idx[k]] >= activations[hi] ? 1.0f : 0.0f;
    states[hi] = inhibition < num_active ? 1.0f : 0.0f;
  }
}
/* sdr/RSDR.cpp:121-146 */
static void rsdr_activate(coder* c, float sparsity) {
  for (int hi = 0; hi < c->nh; hi++) {
    float sum = -c->thr[hi];
    for (int k = c->ff.off[hi]; k < c->ff.off[hi + 1]; k++) sum += c->vis_input[c->ff.idx[k]] * c->ff.w[k];
    for (int k = c->rec.off[hi]; k < c->rec.off[hi + 1]; k++) sum += c->state_prev[c->rec.idx[k]] * c->rec.w[k];
    c->act[hi] = sum;
  }
  rsdr_inhibit(c, sparsity, c->act, c->state);
}
/* scatter reconstruction with arbitrary per-unit drive `s` and scale `mult`:
 * sdr/RSDR.cpp:165-194, sdr/IRSDR.cpp:218-254, neo/SparseCoder.cpp:242-259 */
static void coder_reconstruct(coder* c, const float* s, int use_mult, float mult) {
  memset(c->vis_recon, 0, sizeof(float) * (size_t)c->nv);
  memset(c->recon, 0, sizeof(float) * (size_t)c->nh);
  for (int hi = 0; hi < c->nh; hi++) {
    for (int k = c->ff.off[hi]; k < c->ff.off[hi + 1]; k++)
      c->vis_recon[c->ff.idx[k]] += use_mult ? c->ff.w[k] * s[hi] * mult : c->ff.w[k] * s[hi];
    for (int k = c->rec.off[hi]; k < c->rec.off[hi + 1]; k++)
      c->recon[c->rec.idx[k]] += use_mult ? c->rec.w[k] * s[hi] * mult : c->rec.w[k] * s[hi];
  }
}
/* sdr/RSDR.cpp:196-212 */
static void rsdr_reconstruct_ff(const coder* c, const float* states, float* recon) {
  memset(recon, 0, sizeof(float) * (size_t)c->nv);
  for (int hi = 0; hi < c->nh; hi++)
    for (int k = c->ff.off[hi]; k < c->ff.off[hi + 1]; k++) recon[c->ff.idx[k]] += c->ff.w[k] * states[hi];
}
/* sdr/RSDR.cpp:241-266 */
static void rsdr_learn(coder* c, const float* att, float lr_ff, float lr_rec, float lr_thr, float sparsity) {
  float* ve = (float*)malloc(sizeof(float) * (size_t)c->nv);
  float* he = (float*)malloc(sizeof(float) * (size_t)c->nh);
  for (int vi = 0; vi < c->nv; vi++) ve[vi] = c->vis_input[vi] - c->vis_recon[vi];
  for (int hi = 0; hi < c->nh; hi++) he[hi] = c->state_prev[hi] - c->recon[hi];
  for (int hi = 0; hi < c->nh; hi++) {
    float learn = c->state[hi];
    if (learn > 0.0f) {
      for (int k = c->ff.off[hi]; k < c->ff.off[hi + 1]; k++) c->ff.w[k] += lr_ff * att[hi] * learn * ve[c->ff.idx[k]];
      for (int k = c->rec.off[hi]; k < c->rec.off[hi + 1]; k++) c->rec.w[k] += lr_rec * att[hi] * learn * he[c->rec.idx[k]];
    }
    c->thr[hi] += lr_thr * (c->state[hi] - sparsity);
  }
  free(ve); free(he);
}

/* ======================= iterative coders ===================================== */
/* One leaky-integrate-and-fire sweep over the hidden units: sdr/IRSDR.cpp:138-168
 * (== :177-209, == neo/SparseCoder.cpp:129-158). */
/* start: per-unit initial excitation (activateNoise, neo/SparseCoder.cpp:200) or NULL for 0 */
static void icoder_sweep(coder* c, const float* ve, const float* he, float leak, const float* start, float scale) {
  for (int hi = 0; hi < c->nh; hi++) {
    float excitation = start ? start[hi] * scale : 0.0f;
    for (int k = c->ff.off[hi]; k < c->ff.off[hi + 1]; k++) excitation += ve[c->ff.idx[k]] * c->ff.w[k];
    for (int k = c->rec.off[hi]; k < c->rec.off[hi + 1]; k++) excitation += he[c->rec.idx[k]] * c->rec.w[k];
    float inhibition = 0.0f;
    for (int k = c->lat.off[hi]; k < c->lat.off[hi + 1]; k++) inhibition += c->spike_prev[c->lat.idx[k]] * c->lat.w[k];
    c->act[hi] = (1.0f - leak) * c->act[hi] + excitation - inhibition;
    if (c->act[hi] > c->thr[hi]) { c->act[hi] = 0.0f; c->spike[hi] = 1.0f; }
    else c->spike[hi] = 0.0f;
  }
}
static void icoder_errors(const coder* c, float* ve, float* he) {
  for (int vi = 0; vi < c->nv; vi++) ve[vi] = c->vis_input[vi] - c->vis_recon[vi];
  for (int hi = 0; hi < c->nh; hi++) he[hi] = c->state_prev[hi] - c->recon[hi];
}
/* sdr/IRSDR.cpp:125-216 */
static void irsdr_activate(coder* c, int settle, int measure, float leak) {
  float* ve = (float*)malloc(sizeof(float) * (size_t)c->nv);
  float* he = (float*)malloc(sizeof(float) * (size_t)c->nh);
  for (int hi = 0; hi < c->nh; hi++) { c->act[hi] = 0.0f; c->state[hi] = 0.0f; }
  for (int it = 0; it < settle; it++) {
    icoder_errors(c, ve, he);
    icoder_sweep(c, ve, he, leak, NULL, 0.0f);
    memcpy(c->spike_prev, c->spike, sizeof(float) * (size_t)c->nh);
    coder_reconstruct(c, c->spike, 0, 0.0f);
  }
  float measure_inv = 1.0f / (float)measure;
  for (int it = 0; it < measure; it++) {
    icoder_errors(c, ve, he);
    icoder_sweep(c, ve, he, leak, NULL, 0.0f);
    for (int hi = 0; hi < c->nh; hi++) c->state[hi] += measure_inv * c->spike[hi];
    memcpy(c->spike_prev, c->spike, sizeof(float) * (size_t)c->nh);
    coder_reconstruct(c, c->spike, 0, 0.0f);
  }
  coder_reconstruct(c, c->state, 0, 0.0f); /* :215 reconstructFromStates */
  free(ve); free(he);
}
/* neo/SparseCoder.cpp:116-176 */
/* normals != NULL: SparseCoder::activateNoise (neo/SparseCoder.cpp:178-240), normals[it][hi] = noiseDist(generator) */
static void sparsecoder_activate(coder* c, int iter, float leak, const float* normals, float noise) {
  float* ve = (float*)malloc(sizeof(float) * (size_t)c->nv);
  float* he = (float*)malloc(sizeof(float) * (size_t)c->nh);
  for (int hi = 0; hi < c->nh; hi++) { c->act[hi] = 0.0f; c->state[hi] = 0.0f; }
  float counter = 0.0f;
  for (int it = 0; it < iter; it++) {
    icoder_errors(c, ve, he);
    icoder_sweep(c, ve, he, leak, normals ? normals + (size_t)it * c->nh : NULL, noise);
    for (int hi = 0; hi < c->nh; hi++) c->state[hi] += c->spike[hi];
    memcpy(c->spike_prev, c->spike, sizeof(float) * (size_t)c->nh);
    counter += 1.0f;
    coder_reconstruct(c, c->state, 1, 1.0f / counter);
  }
  float mult = 1.0f / counter;
  for (int hi = 0; hi < c->nh; hi++) c->state[hi] *= mult;
  free(ve); free(he);
}
/* lateral + threshold tail shared by both learn flavours: sdr/IRSDR.cpp:313-317 */
static void icoder_learn_tail(coder* c, int hi, float lr_lat, float lr_thr, float sparsity) {
  for (int k = c->lat.off[hi]; k < c->lat.off[hi + 1]; k++)
    c->lat.w[k] = smax(0.0f, c->lat.w[k] + lr_lat * (c->state[hi] * c->state[c->lat.idx[k]] - sparsity * sparsity));
  c->thr[hi] = smax(0.0f, c->thr[hi] + (c->state[hi] - sparsity) * lr_thr);
}
/* sdr/IRSDR.cpp:287-319 (7-argument learn) */
static void irsdr_learn(coder* c, float lr_ff, float lr_rec, float lr_lat, float lr_thr, float sparsity, float decay, float maxd) {
  float* ve = (float*)malloc(sizeof(float) * (size_t)c->nv);
  float* he = (float*)malloc(sizeof(float) * (size_t)c->nh);
  icoder_errors(c, ve, he);
  for (int hi = 0; hi < c->nh; hi++) {
    float learn = c->state[hi];
    for (int k = c->ff.off[hi]; k < c->ff.off[hi + 1]; k++) {
      float delta = lr_ff * learn * ve[c->ff.idx[k]] - decay * c->ff.w[k];
      c->ff.w[k] += smin(maxd, smax(-maxd, delta));
    }
    for (int k = c->rec.off[hi]; k < c->rec.off[hi + 1]; k++) {
      float delta = lr_rec * learn * he[c->rec.idx[k]] - decay * c->rec.w[k];
      c->rec.w[k] += smin(maxd, smax(-maxd, delta));
    }
    icoder_learn_tail(c, hi, lr_lat, lr_thr, sparsity);
  }
  free(ve); free(he);
}
/* neo/SparseCoder.cpp:326-361 (reward-modulated, eligibility traces) */
static void sparsecoder_learn(coder* c, const float* rewards, float lambda, float lr_ff, float lr_rec, float lr_lat,
                              float lr_thr, float sparsity, float decay, float maxd) {
  float* ve = (float*)malloc(sizeof(float) * (size_t)c->nv);
  float* he = (float*)malloc(sizeof(float) * (size_t)c->nh);
  icoder_errors(c, ve, he);
  for (int hi = 0; hi < c->nh; hi++) {
    float learn = c->state[hi];
    for (int k = c->ff.off[hi]; k < c->ff.off[hi + 1]; k++) {
      float delta = lr_ff * rewards[hi] * c->ff.tr[k] - decay * c->ff.w[k];
      c->ff.w[k] += smin(maxd, smax(-maxd, delta));
      c->ff.tr[k] = lambda * c->ff.tr[k] + learn * ve[c->ff.idx[k]];
    }
    for (int k = c->rec.off[hi]; k < c->rec.off[hi + 1]; k++) {
      float delta = lr_rec * rewards[hi] * c->rec.tr[k] - decay * c->rec.w[k];
      c->rec.w[k] += smin(maxd, smax(-maxd, delta));
      c->rec.tr[k] = lambda * c->rec.tr[k] + learn * he[c->rec.idx[k]];
    }
    icoder_learn_tail(c, hi, lr_lat, lr_thr, sparsity);
  }
  free(ve); free(he);
}

/* ======================= neo::Column ========================================== */
typedef struct {
  float sparsity, gamma, leak, a_ff, a_lat, a_thr, a_q, a_act, gamma_lambda, expl_std, expl_break;
  int iter;
} col_params;

/* neo/Column.cpp:46-169.  u3/v3: the three (u,v) exploration pairs of this visit. */
static void column_sim_step(columns* c, int node, float reward, const col_params* p, const float* u3, const float* v3) {
  const int cells = c->cells, nin = c->in_off[node + 1] - c->in_off[node];
  float* in = c->inputs + c->in_off[node];
  float* rerr = c->recon_err + c->in_off[node];
  float* ff = c->ff_w + (size_t)cells * c->in_off[node];
  float* lat = c->lat_w + (size_t)node * cells * cells;
  float *thr = c->thr + node * cells, *act = c->act + node * cells, *spike = c->spike + node * cells;
  float *spike_prev = c->spike_prev + node * cells, *state = c->state + node * cells;
  float *q_w = c->q_w + node * cells, *q_tr = c->q_tr + node * cells;
  for (int i = 0; i < cells; i++) { act[i] = 0.0f; state[i] = 0.0f; }
  float counter = 0.0f;
  for (int it = 0; it < p->iter; it++) {
    for (int i = 0; i < cells; i++) {
      float excitation = 0.0f;
      for (int j = 0; j < nin; j++) excitation += ff[i * nin + j] * rerr[j];
      float inhibition = 0.0f;
      for (int j = 0; j < cells; j++) inhibition += lat[i * cells + j] * spike_prev[j];
      float activation = (1.0f - p->leak) * act[i] + excitation - inhibition;
      if (activation > thr[i]) { activation = 0.0f; spike[i] = 1.0f; }
      else spike[i] = 0.0f;
      state[i] += spike[i];
      act[i] = activation;
    }
    for (int i = 0; i < cells; i++) spike_prev[i] = spike[i];
    counter += 1.0f;
    float multiplier = 1.0f / counter;
    for (int i = 0; i < nin; i++) {
      float recon = 0.0f;
      for (int j = 0; j < cells; j++) recon += spike[j] * multiplier * ff[j * nin + i];
      rerr[i] = in[i] - recon;
    }
  }
  float multiplier = 1.0f / counter;
  for (int j = 0; j < cells; j++) state[j] *= multiplier;
  float q = 0.0f;
  for (int k = 0; k < cells; k++) q += q_w[k] * state[k];
  for (int a = 0; a < 3; a++) {
    float sum = 0.0f;
    const float* aw = c->a_w + ((size_t)node * 3 + a) * cells;
    for (int k = 0; k < cells; k++) sum += aw[k] * state[k];
    c->a_state[node * 3 + a] = sigmoidf(sum);
  }
  for (int a = 0; a < 3; a++) { /* :126-131 */
    if (u3[a] < p->expl_break) c->a_expl[node * 3 + a] = v3[a];
    else c->a_expl[node * 3 + a] = clamp01(c->a_state[node * 3 + a] + v3[a]);
  }
  float td_error = reward + p->gamma * q - c->prev_value[node];
  float q_alpha_td = p->a_q * td_error, act_alpha_td = p->a_act * td_error;
  for (int k = 0; k < cells; k++) {
    q_w[k] += q_alpha_td * q_tr[k];
    q_tr[k] = q_tr[k] * p->gamma_lambda + state[k];
  }
  for (int a = 0; a < 3; a++) {
    float* aw = c->a_w + ((size_t)node * 3 + a) * cells;
    float* at = c->a_tr + ((size_t)node * 3 + a) * cells;
    for (int i = 0; i < cells; i++) {
      aw[i] += act_alpha_td * at[i];
      at[i] = at[i] * p->gamma_lambda + (c->a_expl[node * 3 + a] - c->a_state[node * 3 + a]) * state[i];
    }
  }
  float sparsity_squared = p->sparsity * p->sparsity;
  for (int i = 0; i < cells; i++) {
    if (state[i] > 0.0f)
      for (int j = 0; j < nin; j++) ff[i * nin + j] += p->a_ff * state[i] * rerr[j];
    for (int j = 0; j < cells; j++)
      lat[i * cells + j] = smax(0.0f, lat[i * cells + j] + p->a_lat * (state[i] * state[j] - sparsity_squared));
    thr[i] += p->a_thr * (state[i] - p->sparsity);
  }
  c->prev_value[node] = q;
}

static col_params layer_col_params(const bidi_layer_desc* d) {
  col_params p = { d->col_sparsity, d->col_gamma, d->col_leak, d->col_alpha_ff, d->col_alpha_lat, d->col_alpha_threshold,
                   d->col_alpha_q, d->col_alpha_action, d->col_gamma_lambda, d->col_expl_stddev, d->col_expl_break, d->col_iter };
  return p;
}
static col_params input_col_params(const bidi_config* g) {
  col_params p = { g->col_sparsity, g->col_gamma, g->col_leak, g->col_alpha_ff, g->col_alpha_lat, g->col_alpha_threshold,
                   g->col_alpha_q, g->col_alpha_action, g->col_gamma_lambda, g->col_expl_stddev, g->col_expl_break, g->col_iter };
  return p;
}

int orc_num_columns(void* h) {
  oracle* o = (oracle*)h;
  if (o->cfg.variant != BIDI_NEO_AGENT && o->cfg.variant != BIDI_CSRL) return 0;
  int n = 0;
  for (int l = 0; l <= o->L; l++) n += o->pn[l].n;
  return n;
}

/* Exploration draws in column visit order (neo/Agent.cpp:155-156,212; neo/Column.cpp:126-131). */
static void draw_noise(const oracle* o, mt19937* g, float* u, float* v) {
  size_t k = 0;
  for (int l = o->L - 1; l >= -1; l--) {
    int n = l >= 0 ? o->pn[l].n : o->pn[o->L].n;
    float std = l >= 0 ? o->ld[l].col_expl_stddev : o->cfg.col_expl_stddev;
    float brk = l >= 0 ? o->ld[l].col_expl_break : o->cfg.col_expl_break;
    for (int pi = 0; pi < n; pi++) {
      normal_state ns = { 0, 0.0f }; /* pertDist is constructed per visit, :48 */
      for (int a = 0; a < 3; a++, k++) {
        u[k] = uniform_real(g, 0.0f, 1.0f);
        v[k] = u[k] < brk ? uniform_real(g, 0.0f, 1.0f) : normal_real(g, &ns, 0.0f, std);
      }
    }
  }
}
static void q_draw_noise(const oracle* o, mt19937* g, float* u, float* v);
static int csrl_num_noise(const oracle* o);
static void csrl_draw_noise(const oracle* o, mt19937* g, float* u, float* v);
static void step_csrl(oracle* o, float reward, const float* u, const float* v, int learn);
int orc_peek_noise(void* h, float* u, float* v) {
  oracle* o = (oracle*)h;
  mt19937 g = o->gen;
  if (o->cfg.variant == BIDI_QPRSDR) { q_draw_noise(o, &g, u, v); return 0; }
  if (o->cfg.variant == BIDI_CSRL) { csrl_draw_noise(o, &g, u, v); return 0; }
  if (o->cfg.variant != BIDI_NEO_AGENT) return -1;
  draw_noise(o, &g, u, v);
  return 0;
}

/* ======================= hierarchy steps ====================================== */
/* delta rule + activation of one prediction node; the part the four classes share
 * (sdr/PredictiveRSDR.cpp:136-158, sdr/IPredictiveRSDR.cpp:169-190). */
static float node_learn_and_activate(pnodes* p, int pi, int learn, float err, float lr_fb, float lr_pred,
                                     const pnodes* above, const float* coder_state, const float* coder_state_prev) {
  if (learn) {
    for (int k = p->fb.off[pi]; k < p->fb.off[pi + 1]; k++) p->fb.w[k] += lr_fb * err * above->state_prev[p->fb.idx[k]];
    for (int k = p->pred.off[pi]; k < p->pred.off[pi + 1]; k++) p->pred.w[k] += lr_pred * err * coder_state_prev[p->pred.idx[k]];
  }
  float activation = 0.0f;
  for (int k = p->fb.off[pi]; k < p->fb.off[pi + 1]; k++) activation += p->fb.w[k] * above->state[p->fb.idx[k]];
  for (int k = p->pred.off[pi]; k < p->pred.off[pi + 1]; k++) activation += p->pred.w[k] * coder_state[p->pred.idx[k]];
  return activation;
}

/* sdr/PredictiveRSDR.cpp:100-194 */
static void step_kwinners(oracle* o, int learn) {
  const int L = o->L;
  float* att[BIDI_MAX_LAYERS];
  for (int l = 0; l < L; l++) {
    coder* c = &o->cod[l];
    rsdr_activate(c, o->ld[l].sparsity);
    coder_reconstruct(c, c->state, 0, 0.0f);
    if (l < L - 1) memcpy(o->cod[l + 1].vis_input, c->state, sizeof(float) * (size_t)c->nh);
  }
  for (int l = L - 1; l >= 0; l--) {
    const bidi_layer_desc* d = &o->ld[l];
    coder* c = &o->cod[l];
    pnodes* p = &o->pn[l];
    att[l] = fzeros(p->n);
    float* pact = fzeros(p->n);
    for (int pi = 0; pi < p->n; pi++) {
      float err = 0.0f;
      if (learn) {
        err = c->state[pi] - p->state_prev[pi];
        float surprise = err * err;
        att[l][pi] = sigmoidf(d->attention_factor * (surprise - p->baseline[pi]));
        p->baseline[pi] = (1.0f - d->avg_surprise_decay) * p->baseline[pi] + d->avg_surprise_decay * surprise;
      }
      pact[pi] = p->act[pi] = node_learn_and_activate(p, pi, learn, err, d->learn_fb, d->learn_pred,
                                                      l < L - 1 ? &o->pn[l + 1] : NULL, c->state, c->state_prev);
    }
    rsdr_inhibit(c, d->sparsity, pact, p->state);
    free(pact);
  }
  for (int l = 0; l < L; l++) {
    const bidi_layer_desc* d = &o->ld[l];
    coder* c = &o->cod[l];
    pnodes* p = &o->pn[l];
    if (learn) rsdr_learn(c, att[l], d->learn_ff, d->learn_rec, d->learn_threshold, d->sparsity);
    memcpy(c->state_prev, c->state, sizeof(float) * (size_t)c->nh); /* stepEnd, sdr/RSDR.cpp:297-300 */
    memcpy(p->state_prev, p->state, sizeof(float) * (size_t)p->n);
    memcpy(p->act_prev, p->act, sizeof(float) * (size_t)p->n);
    free(att[l]);
  }
  rsdr_reconstruct_ff(&o->cod[0], o->pn[0].state, o->prediction);
}

/* Input prediction nodes: sdr/IPredictiveRSDR.cpp:198-218, neo/Agent.cpp:212-247 */
static void input_nodes(oracle* o, int learn, float reward, const float* u, const float* v, size_t* noise_k) {
  pnodes* p = &o->pn[o->L];
  const pnodes* p0 = &o->pn[0];
  const coder* c0 = &o->cod[0];
  const int agent = o->cfg.variant == BIDI_NEO_AGENT;
  col_params cp = input_col_params(&o->cfg);
  for (int pi = 0; pi < p->n; pi++) {
    float gate = 1.0f;
    if (agent) {
      float* in = p->col.inputs + p->col.in_off[pi];
      int ci = 0;
      for (int k = p->fb.off[pi]; k < p->fb.off[pi + 1]; k++) {
        in[ci++] = p0->col.a_expl[p->fb.idx[k] * 3 + 2]; /* _signal */
        in[ci++] = p0->state[p->fb.idx[k]];
      }
      column_sim_step(&p->col, pi, reward, &cp, u + *noise_k, v + *noise_k);
      *noise_k += 3;
      gate = p->col.a_expl[pi * 3 + 1]; /* _learnPrediction */
    }
    if (learn) {
      float err = agent ? gate * (c0->vis_input[pi] - p->state_prev[pi]) : c0->vis_input[pi] - p->state_prev[pi];
      for (int k = p->fb.off[pi]; k < p->fb.off[pi + 1]; k++) p->fb.w[k] += o->cfg.learn_infb * err * p0->state_prev[p->fb.idx[k]];
    }
    float activation = 0.0f;
    for (int k = p->fb.off[pi]; k < p->fb.off[pi + 1]; k++) activation += p->fb.w[k] * p0->state[p->fb.idx[k]];
    p->act[pi] = activation;
    p->state[pi] = activation;
  }
}

/* sdr/IPredictiveRSDR.cpp:138-240 */
static void step_iterative(oracle* o, int learn) {
  const int L = o->L;
  for (int l = 0; l < L; l++) {
    coder* c = &o->cod[l];
    irsdr_activate(c, o->ld[l].iter_settle, o->ld[l].iter_measure, o->ld[l].leak);
    if (l < L - 1) memcpy(o->cod[l + 1].vis_input, c->state, sizeof(float) * (size_t)c->nh);
  }
  for (int l = L - 1; l >= 0; l--) {
    const bidi_layer_desc* d = &o->ld[l];
    coder* c = &o->cod[l];
    pnodes* p = &o->pn[l];
    for (int pi = 0; pi < p->n; pi++) {
      float err = 0.0f;
      if (learn) {
        err = c->state[pi] - p->state_prev[pi];
        float error2 = err * err;
        /* rewards[l][pi] = sigmoid(sens*(baseline-error2)) is computed and never used, :165 */
        p->baseline[pi] = (1.0f - d->baseline_decay) * p->baseline[pi] + d->baseline_decay * error2;
      }
      p->act[pi] = node_learn_and_activate(p, pi, learn, err, d->learn_fb, d->learn_pred,
                                           l < L - 1 ? &o->pn[l + 1] : NULL, c->state, c->state_prev);
      p->state[pi] = clamp01(p->act[pi]);
    }
  }
  size_t nk = 0;
  input_nodes(o, learn, 0.0f, NULL, NULL, &nk);
  for (int l = 0; l <= L; l++) {
    pnodes* p = &o->pn[l];
    if (l < L) {
      const bidi_layer_desc* d = &o->ld[l];
      coder* c = &o->cod[l];
      if (learn) irsdr_learn(c, d->learn_ff, d->learn_rec, d->learn_lat, d->learn_threshold, d->sparsity, d->weight_decay, d->max_weight_delta);
      memcpy(c->state_prev, c->state, sizeof(float) * (size_t)c->nh);
    }
    memcpy(p->state_prev, p->state, sizeof(float) * (size_t)p->n);
    memcpy(p->act_prev, p->act, sizeof(float) * (size_t)p->n);
  }
}

/* neo/PredictiveHierarchy.cpp:137-245 and neo/Agent.cpp:141-284 */
/* gen_normals != NULL: simStepGenerate (neo/PredictiveHierarchy.cpp:247-315) = the same pass without any learning,
 * the coders driven by noise */
static void step_neo_ex(oracle* o, float reward, const float* u, const float* v, int learn, const float* gen_normals, float gen_noise) {
  const int L = o->L;
  const int agent = o->cfg.variant == BIDI_NEO_AGENT;
  size_t gen_off = 0;
  for (int l = 0; l < L; l++) {
    coder* c = &o->cod[l];
    sparsecoder_activate(c, o->ld[l].iter_settle, o->ld[l].leak, gen_normals ? gen_normals + gen_off : NULL, gen_noise);
    gen_off += (size_t)o->ld[l].iter_settle * c->nh;
    if (l < L - 1) {
      float* nin = o->cod[l + 1].vis_input;
      if (agent) for (int i = 0; i < c->nh; i++) nin[i] = c->state[i] * o->pn[l].col.a_expl[i * 3 + 0]; /* _attention, neo/Agent.cpp:149 */
      else memcpy(nin, c->state, sizeof(float) * (size_t)c->nh);
    }
  }
  size_t nk = 0;
  for (int l = L - 1; l >= 0; l--) {
    const bidi_layer_desc* d = &o->ld[l];
    coder* c = &o->cod[l];
    pnodes* p = &o->pn[l];
    const pnodes* above = l < L - 1 ? &o->pn[l + 1] : NULL;
    col_params cp = layer_col_params(d);
    for (int pi = 0; pi < p->n; pi++) {
      float gate = 1.0f;
      if (agent) { /* neo/Agent.cpp:159-180 */
        float* in = p->col.inputs + p->col.in_off[pi];
        int ci = 0;
        for (int k = p->fb.off[pi]; k < p->fb.off[pi + 1]; k++) {
          in[ci++] = above->col.a_expl[p->fb.idx[k] * 3 + 2];
          in[ci++] = above->state[p->fb.idx[k]];
        }
        for (int k = p->pred.off[pi]; k < p->pred.off[pi + 1]; k++) in[ci++] = c->state[p->pred.idx[k]];
        column_sim_step(&p->col, pi, reward, &cp, u + nk, v + nk);
        nk += 3;
        gate = p->col.a_expl[pi * 3 + 1];
      }
      float err = 0.0f;
      if (learn) err = agent ? gate * (c->state[pi] - p->state_prev[pi]) : c->state[pi] - p->state_prev[pi];
      p->act[pi] = node_learn_and_activate(p, pi, learn, err, d->learn_fb, d->learn_pred, above, c->state, c->state_prev);
      p->state[pi] = clamp01(p->act[pi]);
    }
  }
  input_nodes(o, learn, reward, u, v, &nk);
  for (int l = 0; l <= L; l++) {
    pnodes* p = &o->pn[l];
    if (l < L) {
      const bidi_layer_desc* d = &o->ld[l];
      coder* c = &o->cod[l];
      if (learn) { /* neo/Agent.cpp:252-265 */
        float* rewards = fzeros(p->n);
        for (int pi = 0; pi < p->n; pi++) {
          float err = c->state[pi] - p->state_prev[pi];
          float error2 = err * err;
          rewards[pi] = sigmoidf(d->sensitivity * (error2 - p->baseline[pi]));
          p->baseline[pi] = (1.0f - d->baseline_decay) * p->baseline[pi] + d->baseline_decay * error2;
        }
        sparsecoder_learn(c, rewards, d->lambda, d->learn_ff, d->learn_rec, d->learn_lat, d->learn_threshold,
                          d->sparsity, d->weight_decay, d->max_weight_delta);
        free(rewards);
      }
      memcpy(c->state_prev, c->state, sizeof(float) * (size_t)c->nh);
    }
    memcpy(p->state_prev, p->state, sizeof(float) * (size_t)p->n);
    memcpy(p->act_prev, p->act, sizeof(float) * (size_t)p->n);
  }
}

static void step_neo(oracle* o, float reward, const float* u, const float* v, int learn) { step_neo_ex(o, reward, u, v, learn, NULL, 0.0f); }

/* N(0,1) draws of one simStepGenerate: layer 0 first, iteration outer, unit inner, a fresh normal_distribution per layer */
int orc_num_generate_noise(void* h) {
  oracle* o = (oracle*)h;
  if (o->cfg.variant != BIDI_NEO_HIERARCHY) return 0;
  int n = 0;
  for (int l = 0; l < o->L; l++) n += o->ld[l].iter_settle * o->cod[l].nh;
  return n;
}
static void draw_generate(const oracle* o, mt19937* g, float* out) {
  size_t k = 0;
  for (int l = 0; l < o->L; l++) {
    normal_state ns = { 0, 0.0f };
    const int n = o->ld[l].iter_settle * o->cod[l].nh;
    for (int j = 0; j < n; j++) out[k++] = normal_real(g, &ns, 0.0f, 1.0f);
  }
}
int orc_peek_generate(void* h, float* out) {
  oracle* o = (oracle*)h;
  if (o->cfg.variant != BIDI_NEO_HIERARCHY) return -1;
  mt19937 g = o->gen;
  draw_generate(o, &g, out);
  return 0;
}
/* simStepGenerate drawing from the oracle's own generator */
void orc_step_generate(void* h, float noise) {
  oracle* o = (oracle*)h;
  if (o->cfg.variant != BIDI_NEO_HIERARCHY) return;
  float* nz = fzeros((size_t)orc_num_generate_noise(h));
  draw_generate(o, &o->gen, nz);
  step_neo_ex(o, 0.0f, NULL, NULL, 0, nz, noise);
  free(nz);
}

/* ---- sdr::QPRSDR::simStep after _prsdr.simStep (sdr/QPRSDR.cpp:102-341) */
static float clamp11(float x) { return smin(1.0f, smax(-1.0f, x)); }
/* forward pass of the Q network from the given action values (:111-144, :219-253) */
static void q_forward(oracle* o, const float* actions) {
  qnet* q = &o->q;
  for (int l = 0; l < o->L; l++) {
    qlayer* ql = &q->ql[l];
    for (int qi = 0; qi < ql->n; qi++) {
      float sum = 0.0f;
      for (int k = ql->ff.off[qi]; k < ql->ff.off[qi + 1]; k++)
        sum += ql->ff.w[k] * (l == 0 ? actions[q->act_index[ql->ff.idx[k]]] : q->ql[l - 1].state[ql->ff.idx[k]]);
      ql->state[qi] = sigmoidf(sum) * o->pn[l].state[qi];
      ql->err[qi] = 0.0f;
    }
  }
}
/* backpropagation of dQ/dstate from the top (:154-190, :268-301); to_actions: also into the action nodes */
static void q_backward(oracle* o, int to_actions) {
  qnet* q = &o->q;
  const int L = o->L;
  for (int i = 0; i < q->n_top; i++) q->ql[L - 1].err[i] = q->qc_w[i];
  for (int l = L - 1; l >= 0; l--) {
    qlayer* ql = &q->ql[l];
    for (int qi = 0; qi < ql->n; qi++) {
      ql->err[qi] = ql->err[qi] * ql->state[qi] * (1.0f - ql->state[qi]);
      if (l > 0) for (int k = ql->ff.off[qi]; k < ql->ff.off[qi + 1]; k++) q->ql[l - 1].err[ql->ff.idx[k]] += ql->err[qi] * ql->ff.w[k];
      else if (to_actions) for (int k = ql->ff.off[qi]; k < ql->ff.off[qi + 1]; k++) q->a_err[q->act_index[ql->ff.idx[k]]] += ql->err[qi] * ql->ff.w[k];
    }
  }
}
static int q_num_noise(const oracle* o) { return o->cfg.variant == BIDI_QPRSDR ? o->q.half : 0; }
/* :202-213: one pertDist object for the whole loop, so its pair cache carries over between actions */
static void q_draw_noise(const oracle* o, mt19937* g, float* u, float* v) {
  normal_state ns = { 0, 0.0f };
  for (int i = 0; i < o->q.half; i++) {
    u[i] = uniform_real(g, 0.0f, 1.0f);
    v[i] = u[i] < o->cfg.q_expl_break ? uniform_real(g, 0.0f, 1.0f) : normal_real(g, &ns, 0.0f, o->cfg.q_expl_stddev);
  }
}
static void step_q(oracle* o, float reward, const float* u, const float* v, int learn) {
  const bidi_config* g = &o->cfg;
  qnet* q = &o->q;
  const int L = o->L, half = q->half;
  for (int i = 0; i < q->n_nodes; i++) q->a_pred[i] = q->a_derive[i] = clamp11(o->pn[L].state[q->a_input[i]]); /* :106-107 */
  for (int iter = 0; iter < g->q_derive_iterations; iter++) { /* :110-200 */
    q_forward(o, q->a_derive);
    for (int i = 0; i < q->n_nodes; i++) q->a_err[i] = 0.0f;
    q_backward(o, 1);
    for (int i = 0; i < half; i++) q->a_derive[i] = clamp11(q->a_derive[i] + (q->a_err[i] > 0.0f ? 1.0f : -1.0f) * g->q_derive_alpha);
    for (int i = 0; i < half; i++) q->a_derive[half + i] = 1.0f - q->a_derive[i];
  }
  for (int i = 0; i < half; i++) q->a_expl[i] = u[i] < g->q_expl_break ? v[i] : clamp11(q->a_derive[i] + v[i]); /* :205-213 */
  for (int i = 0; i < half; i++) q->a_expl[half + i] = 1.0f - q->a_expl[i];
  q_forward(o, q->a_expl);
  float qv = 0.0f;
  for (int i = 0; i < q->n_top; i++) qv += q->qc_w[i] * q->ql[L - 1].state[i];
  const float td = reward + g->q_gamma * qv - q->prev_value[0];
  const float q_alpha_td = g->q_alpha * td, action_alpha_td = g->q_action_alpha * td;
  q->prev_value[0] = qv;
  if (learn) { /* :266-331 */
    q_backward(o, 0);
    for (int l = 0; l < L; l++) {
      qlayer* ql = &q->ql[l];
      for (int qi = 0; qi < ql->n; qi++) {
        ql->bias_w[qi] += action_alpha_td * ql->bias_tr[qi];
        ql->bias_tr[qi] = g->q_gamma_lambda * ql->bias_tr[qi] + ql->err[qi];
        for (int k = ql->ff.off[qi]; k < ql->ff.off[qi + 1]; k++) {
          ql->ff.w[k] += action_alpha_td * ql->ff.tr[k];
          const float src = l == 0 ? q->a_expl[q->act_index[ql->ff.idx[k]]] : q->ql[l - 1].state[ql->ff.idx[k]];
          ql->ff.tr[k] = g->q_gamma_lambda * ql->ff.tr[k] + ql->err[qi] * src;
        }
      }
    }
  }
  for (int i = 0; i < q->n_top; i++) { /* :334-340, also when not learning */
    q->qc_w[i] += q_alpha_td * q->qc_tr[i];
    q->qc_tr[i] = g->q_gamma_lambda * q->qc_tr[i] + q->ql[L - 1].state[i];
  }
}

int orc_num_noise(void* h) {
  oracle* o = (oracle*)h;
  if (o->cfg.variant == BIDI_CSRL) return csrl_num_noise(o);
  return o->cfg.variant == BIDI_QPRSDR ? q_num_noise(o) : orc_num_columns(h) * 3;
}

void orc_step_noise(void* h, float reward, const float* u, const float* v, int learn) {
  oracle* o = (oracle*)h;
  switch (o->cfg.variant) {
    case BIDI_KWINNERS: step_kwinners(o, learn); break;
    case BIDI_ITERATIVE: step_iterative(o, learn); break;
    case BIDI_QPRSDR: step_iterative(o, learn); step_q(o, reward, u, v, learn); break;
    case BIDI_CSRL: step_csrl(o, reward, u, v, learn); break;
    default: step_neo(o, reward, u, v, learn); break;
  }
}
void orc_step(void* h, float reward, int learn) {
  oracle* o = (oracle*)h;
  if (o->cfg.variant == BIDI_QPRSDR) {
    float* u = fzeros((size_t)o->q.half);
    float* v = fzeros((size_t)o->q.half);
    step_iterative(o, learn);  /* the hierarchy draws nothing (sdr/IRSDR.cpp:126 never samples) */
    q_draw_noise(o, &o->gen, u, v);
    step_q(o, reward, u, v, learn);
    free(u); free(v);
    return;
  }
  if (o->cfg.variant == BIDI_CSRL) {
    size_t n = (size_t)csrl_num_noise(o);
    float* u = fzeros(n);
    float* v = fzeros(n);
    csrl_draw_noise(o, &o->gen, u, v);
    step_csrl(o, reward, u, v, learn);
    free(u); free(v);
    return;
  }
  if (o->cfg.variant == BIDI_NEO_AGENT) {
    size_t n = (size_t)orc_num_columns(h) * 3;
    float* u = fzeros(n);
    float* v = fzeros(n);
    draw_noise(o, &o->gen, u, v);
    step_neo(o, reward, u, v, learn);
    free(u); free(v);
  } else orc_step_noise(h, reward, NULL, NULL, learn);
}

/* ======================= accessors ============================================ */
static float* find_array(oracle* o, int which, int layer, long long* len) {
  const int L = o->L, V = o->cfg.variant;
  *len = 0;
  if (layer < 0 || layer > L) return NULL;
  if (which == BIDI_PREDICTION) {
    if (V != BIDI_KWINNERS || layer != 0) return NULL;
    *len = o->n_in; return o->prediction;
  }
  const int iter = V != BIDI_KWINNERS, traced = V == BIDI_NEO_HIERARCHY || V == BIDI_NEO_AGENT || V == BIDI_CSRL;
  if (which == BIDI_REWARD_OFFSETS) {
    if (V != BIDI_CSRL || layer != L - 1) return NULL;
    *len = o->cod[L - 1].nh; return o->reward_offsets;
  }
  if (which <= BIDI_REC_TRACE) {
    if (layer >= L) return NULL;
    coder* c = &o->cod[layer];
    switch (which) {
      case BIDI_VIS_INPUT: *len = c->nv; return c->vis_input;
      case BIDI_VIS_RECON: *len = c->nv; return c->vis_recon;
      case BIDI_HID_THRESHOLD: *len = c->nh; return c->thr;
      case BIDI_HID_STATE: *len = c->nh; return c->state;
      case BIDI_HID_STATE_PREV: *len = c->nh; return c->state_prev;
      case BIDI_HID_ACTIVATION: *len = c->nh; return c->act;
      case BIDI_HID_RECON: *len = c->nh; return c->recon;
      case BIDI_HID_SPIKE: if (!iter) return NULL; *len = c->nh; return c->spike;
      case BIDI_HID_SPIKE_PREV: if (!iter) return NULL; *len = c->nh; return c->spike_prev;
      case BIDI_FF_W: *len = c->ff.off[c->nh]; return c->ff.w;
      case BIDI_REC_W: *len = c->rec.off[c->nh]; return c->rec.w;
      case BIDI_LAT_W: if (!iter) return NULL; *len = c->lat.off[c->nh]; return c->lat.w;
      case BIDI_FF_TRACE: if (!traced) return NULL; *len = c->ff.off[c->nh]; return c->ff.tr;
      case BIDI_REC_TRACE: if (!traced) return NULL; *len = c->rec.off[c->nh]; return c->rec.tr;
    }
    return NULL;
  }
  if (layer == L && V == BIDI_KWINNERS) return NULL;
  pnodes* p = &o->pn[layer];
  switch (which) {
    case BIDI_P_STATE: *len = p->n; return p->state;
    case BIDI_P_STATE_PREV: *len = p->n; return p->state_prev;
    case BIDI_P_ACTIVATION: *len = p->n; return p->act;
    case BIDI_P_ACTIVATION_PREV: *len = p->n; return p->act_prev;
    case BIDI_P_BASELINE: if (layer == L) return NULL; *len = p->n; return p->baseline;
    case BIDI_FB_W: *len = p->fb.off[p->n]; return p->fb.w;
    case BIDI_PRED_W: if (layer == L) return NULL; *len = p->pred.off[p->n]; return p->pred.w;
    case BIDI_FB_TRACE: if (V != BIDI_CSRL) return NULL; *len = p->fb.off[p->n]; return p->fb.tr;
    case BIDI_PRED_TRACE: if (V != BIDI_CSRL || layer == L) return NULL; *len = p->pred.off[p->n]; return p->pred.tr;
    case BIDI_P_LOCAL_REWARD: if (V != BIDI_CSRL) return NULL; *len = p->n; return p->local_reward;
  }
  if (V == BIDI_QPRSDR) {
    qnet* q = &o->q;
    if (layer < L) {
      qlayer* ql = &q->ql[layer];
      switch (which) {
        case BIDI_Q_STATE: *len = ql->n; return ql->state;
        case BIDI_Q_ERROR: *len = ql->n; return ql->err;
        case BIDI_Q_BIAS_W: *len = ql->n; return ql->bias_w;
        case BIDI_Q_BIAS_TRACE: *len = ql->n; return ql->bias_tr;
        case BIDI_Q_FF_W: *len = ql->ff.off[ql->n]; return ql->ff.w;
        case BIDI_Q_FF_TRACE: *len = ql->ff.off[ql->n]; return ql->ff.tr;
      }
    }
    if (layer == 0)
      switch (which) {
        case BIDI_QCONN_W: *len = q->n_top; return q->qc_w;
        case BIDI_QCONN_TRACE: *len = q->n_top; return q->qc_tr;
        case BIDI_ACT_PREDICTED: *len = q->n_nodes; return q->a_pred;
        case BIDI_ACT_DERIVE: *len = q->n_nodes; return q->a_derive;
        case BIDI_ACT_EXPL: *len = q->n_nodes; return q->a_expl;
        case BIDI_ACT_ERROR: *len = q->n_nodes; return q->a_err;
        case BIDI_Q_PREV_VALUE: *len = 1; return q->prev_value;
      }
    return NULL;
  }
  if (V != BIDI_NEO_AGENT && V != BIDI_CSRL) return NULL;
  columns* c = &p->col;
  long long tot = c->in_off[c->n], nc = (long long)c->n * c->cells;
  if (V == BIDI_CSRL)
    switch (which) {
      case BIDI_COL_ACT_ERROR: *len = (long long)c->n * c->na; return c->a_err;
      case BIDI_COL_CELL_ACT_STATE: *len = nc; return c->cell_as;
      case BIDI_COL_CELL_ACT_ERROR: *len = nc; return c->cell_ae;
      case BIDI_COL_AVG_SURPRISE: *len = c->n; return c->avg_surprise;
    }
  switch (which) {
    case BIDI_COL_INPUT: *len = tot; return c->inputs;
    case BIDI_COL_RECON_ERR: *len = tot; return c->recon_err;
    case BIDI_COL_FF_W: *len = tot * c->cells; return c->ff_w;
    case BIDI_COL_LAT_W: *len = nc * c->cells; return c->lat_w;
    case BIDI_COL_THRESHOLD: *len = nc; return c->thr;
    case BIDI_COL_ACTIVATION: *len = nc; return c->act;
    case BIDI_COL_SPIKE: *len = nc; return c->spike;
    case BIDI_COL_SPIKE_PREV: *len = nc; return c->spike_prev;
    case BIDI_COL_STATE: *len = nc; return c->state;
    case BIDI_COL_Q_W: *len = nc; return c->q_w;
    case BIDI_COL_Q_TRACE: *len = nc; return c->q_tr;
    case BIDI_COL_ACT_STATE: *len = (long long)c->n * c->na; return c->a_state;
    case BIDI_COL_ACT_EXPL: *len = (long long)c->n * c->na; return c->a_expl;
    case BIDI_COL_ACT_W: *len = nc * c->na; return c->a_w;
    case BIDI_COL_ACT_TRACE: *len = nc * c->na; return c->a_tr;
    case BIDI_COL_PREV_VALUE: *len = c->n; return c->prev_value;
  }
  return NULL;
}
long long orc_array_len(void* h, int which, int layer) {
  long long n;
  find_array((oracle*)h, which, layer, &n);
  return n;
}
int orc_get_array(void* h, int which, int layer, float* dst) {
  long long n;
  float* p = find_array((oracle*)h, which, layer, &n);
  if (!p) return -1;
  memcpy(dst, p, sizeof(float) * (size_t)n);
  return 0;
}
int orc_set_array(void* h, int which, int layer, const float* src) {
  long long n;
  float* p = find_array((oracle*)h, which, layer, &n);
  if (!p) return -1;
  memcpy(p, src, sizeof(float) * (size_t)n);
  return 0;
}
int orc_conn_counts(void* h, int which, int layer, int* counts) {
  oracle* o = (oracle*)h;
  const int L = o->L;
  if (layer < 0 || layer > L) return -1;
  const conns* c = NULL;
  if (layer < L) {
    if (which == BIDI_FF_W) c = &o->cod[layer].ff;
    else if (which == BIDI_REC_W) c = &o->cod[layer].rec;
    else if (which == BIDI_LAT_W) c = &o->cod[layer].lat;
    else if (which == BIDI_PRED_W) c = &o->pn[layer].pred;
  }
  if (which == BIDI_FB_W && (layer < L || o->cfg.variant != BIDI_KWINNERS)) c = &o->pn[layer].fb;
  if (which == BIDI_Q_FF_W && o->cfg.variant == BIDI_QPRSDR && layer < L) c = &o->q.ql[layer].ff;
  if (c) { for (int i = 0; i < c->n_units; i++) counts[i] = c->off[i + 1] - c->off[i]; return 0; }
  if (which == BIDI_COL_INPUT && (o->cfg.variant == BIDI_NEO_AGENT || o->cfg.variant == BIDI_CSRL)) {
    const columns* q = &o->pn[layer].col;
    for (int i = 0; i < q->n; i++) counts[i] = q->in_off[i + 1] - q->in_off[i];
    return 0;
  }
  return -1;
}
void orc_set_input(void* h, int i, float v) { ((oracle*)h)->cod[0].vis_input[i] = v; }
void orc_set_inputs(void* h, const float* in, int n) { memcpy(((oracle*)h)->cod[0].vis_input, in, sizeof(float) * (size_t)n); }
float orc_get_prediction(void* h, int i) {
  oracle* o = (oracle*)h;
  return o->cfg.variant == BIDI_KWINNERS ? o->prediction[i] : o->pn[o->L].state[i];
}
void orc_get_predictions(void* h, float* out, int n) {
  for (int i = 0; i < n; i++) out[i] = orc_get_prediction(h, i);
}
void orc_run(void* h, const float* frames, int n_frames, int n_in, const float* rewards, int steps, int learn) {
  for (int t = 0; t < steps; t++) {
    orc_set_inputs(h, frames + (size_t)(t % n_frames) * n_in, n_in);
    orc_step(h, rewards ? rewards[t % n_frames] : 0.0f, learn);
  }
}

/* RunnerMain.cpp:235-251, headless; same loop as ref_run_runner in ref_harness.cpp. */
void orc_run_runner(void* h, const float* frames, int n_frames, int n_in, int steps, int n_fb, const int* src,
                    const int* dst, int learn) {
  oracle* o = (oracle*)h;
  for (int t = 0; t < steps; t++) {
    orc_set_inputs(h, frames + (size_t)(t % n_frames) * n_in, n_in);
    normal_state ns = { 0, 0.0f };
    float sum = 0.0f;
    for (int k = 0; k < n_fb; k++) {
      const float a = orc_get_prediction(h, src[k]) * 0.5f + 0.5f;
      sum += clamp01(a);
      orc_set_input(h, dst[k], clamp01(a + normal_real(&o->gen, &ns, 0.0f, 0.05f)));
    }
    orc_step(h, n_fb ? sum / (float)n_fb - 0.5f : 0.0f, learn);
  }
}

#include "csrl_oracle.inc"
