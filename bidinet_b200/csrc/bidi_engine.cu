// bidi_engine.cu -- host side of the engine and the extern "C" boundary
// (include/bidinet_b200.h).  Builds the geometry tables and the HBM layout
// (bidi_types.h), converts between the reference's serialised state and the slot
// planes, and records one simStep as a CUDA graph of the kernels in
// bidi_kernels.cuh.  There is no CPU compute path: without a CUDA device
// bidi_create fails with BIDI_ENODEV.
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "bidi_kernels.cuh"
#include "bidi_coder.cuh"
#include "bidi_coder_dense.cuh"
#include "bidi_tiles.cuh"
#include "bidi_grid.cuh"
#include "bidi_persist.cuh"
#include "bidi_qnet.cuh"
#include "bidi_env.cuh"
#include "bidi_csrl.cuh"

namespace {

thread_local std::string g_last_error;

struct WinHost {
  WinDev d;
  std::vector<int> cx, cy, gxlo, gxhi, gylo, gyhi;
};

struct ColumnsHost {
  std::vector<int> in_off, rec_off, stride, in_src;
  std::vector<int> in_loc;   // ColumnsDev::in_loc, two 16-bit entries per int
  ColPairTab tab;            // k_columns16's constant-bank copy of in_off / rec_off / stride
};
struct SdrrlHost {
  std::vector<int> in_off, in_src;
};

// Pointers to the sections of one column inside a host blob image.  Columns are
// stored in PAIRS (nodes 2p, 2p+1) with every element interleaved [..][2]
// (layout: bidi_types.h, ColumnsDev), so element k of a section is ptr[2*k].
struct ColRec {
  float *W, *Wreg, *latT, *thr, *q_w, *q_tr, *a_w, *a_tr, *sprev, *err, *scal;
  int S, Cq, nin;
  int es;        // element stride: 2 in the pair-interleaved layout, 1 in the planar one
  int cs;        // stride of the per-cell scalars thr, q_w, q_tr, sprev: es, or the pitch of the planar layout's cell rows
  bool planar;
  int rh, lane0; // planar: inputs per row kept in the lane-major register part; this column's first lane (0 or 16)
  // lateral weight of cell i from cell j: planar rows [i][BIDI_COL_ROW] = (lat[i][0..15], thr, q_w, q_tr, sprev);
  // interleaved: transposed, latT[j][i]
  float& lat(int i, int j) const { return planar ? latT[i * BIDI_COL_ROW + j] : latT[es * (j * Cq + i)]; }
  // weight of cell i, input j
  float& w(int i, int j) const {
    if (j < rh) return Wreg[((long long)(j >> 2) * 32 + lane0 + i) * 4 + (j & 3)];
    return W[(long long)es * ((long long)i * (S - rh) + (j - rh))];
  }
};
// One column's block of a pair record: weights (the part not kept in registers), transposed lateral weights, ten
// per-cell vectors, the reconstruction error, prevValue.  Planar blocks are padded to 16 (mod 32) floats so that the
// two columns' 16-lane accesses to the same section fall in different halves of the shared-memory banks.
int column_block_floats(const ColumnsDev& C, int S) {
  int block = C.cells * (S - (C.planar ? C.rh : 0)) + C.cells * C.cq + 10 * C.cq + S + 4;
  if (C.planar) block += (16 - block % 32 + 32) % 32;
  return block;
}
ColRec col_rec(const ColumnsDev& C, const ColumnsHost& H, float* blob, int node) {
  ColRec r;
  const int pair = node >> 1, half = node & 1;
  r.S = H.stride[pair]; r.Cq = C.cq; r.nin = H.in_off[node + 1] - H.in_off[node];
  r.rh = C.planar ? C.rh : 0; r.lane0 = half * 16;
  const int sh = r.S - r.rh;
  const int block = column_block_floats(C, r.S);  // floats of ONE column's shared-memory block
  r.es = C.planar ? 1 : 2;
  // planar: the register part, then column 2p's block, then column 2p+1's; interleaved: element k of a section is ptr[2*k]
  r.Wreg = blob + C.rec_base + H.rec_off[pair];
  r.W = r.Wreg + 32 * r.rh + (C.planar ? half * block : half);
  const int e = r.es;
  r.latT = r.W + (long long)e * C.cells * sh;
  r.planar = C.planar != 0;
  if (r.planar) {  // cell rows: 16 lateral weights + the cell's four scalars, one 128-bit access each (bidi_columns16.cuh)
    r.cs = BIDI_COL_ROW;
    r.thr = r.latT + 16; r.q_w = r.latT + 17; r.q_tr = r.latT + 18; r.sprev = r.latT + 19;
    r.a_w = r.latT + C.cells * BIDI_COL_ROW; r.a_tr = r.a_w + 3 * C.cq;
    r.err = r.a_tr + 3 * C.cq; r.scal = r.err + r.S;
    return r;
  }
  r.cs = e;
  r.thr = r.latT + e * C.cells * C.cq;
  r.q_w = r.thr + e * C.cq; r.q_tr = r.q_w + e * C.cq; r.a_w = r.q_tr + e * C.cq; r.a_tr = r.a_w + 3 * e * C.cq;
  r.sprev = r.a_tr + 3 * e * C.cq; r.err = r.sprev + e * C.cq; r.scal = r.err + e * r.S;
  return r;
}
// common row stride of a column pair, >= nIn of both.
// interleaved layout: even, with stride/2 odd so that lane i's 128-bit shared-memory loads of
// row i (2*stride floats apart) never conflict.  planar layout: stride = 4 (mod 8), so that
// the 128-bit loads of rows 0..7 (one per lane of a quarter warp) fall in distinct banks.
int column_pair_stride(int nin, bool planar) {
  if (planar) {
    int s = std::max(4, (nin + 3) / 4 * 4);
    if ((s & 7) != 4) s += 4;
    return s;
  }
  int s = std::max(2, (nin + 1) / 2 * 2);
  if (((s / 2) & 1) == 0) s += 2;
  return s;
}

// row pitch of a k-winners family laid out as rows (bidi_grid.cuh): >= its slots, 4 (mod 8) floats so that a warp's
// 128-bit shared-memory loads of 32 staged rows never conflict and every row starts 16-byte aligned
inline int kw_pitch(int nslots) { return bidi::dense_pitch(std::max(1, nslots)); }

struct LayerHost {
  WinHost ff, rec, lat, fb, pred;
  ColumnsHost col;
  SdrrlHost sd;
  std::vector<int> dense_mask;  // LayerDev::dense_mask
  // layers the dense fused coder can run: the region that holds the feed-forward + recurrent weights (and one for
  // their traces) in either layout -- slot planes (ff at region, rec at region + rec_plane) or rows (WinDev::row_pitch)
  bool dense_geom = false;
  long long w_region = -1, tr_region = -1, rec_plane = 0, region_len = 0;
  int row_nf = 0, row_pitch = 0;
};

} // namespace

struct bidi_engine {
  bidi_config cfg;
  std::vector<bidi_layer_desc> ld;
  int device = 0, A = 0, L = 0, n_in = 0, ncol = 0, max_units = 0;
  std::vector<LayerDev> layers;    // host copy, device pointers inside
  std::vector<LayerHost> hl;
  long long stride = 0;            // floats per agent blob
  std::vector<int> coder_cta;      // per layer: the fused one-CTA-per-agent coder kernel applies
  std::vector<int> tiled;          // per node layer: few threads, use the tile kernels (bidi_tiles.cuh)
  bool force_phase_kernels = false, no_dense_coder = false;
  // per layer: the 2-D tile kernels of bidi_grid.cuh apply (large k-winners layers); shared-memory floats they stage
  struct GridSizes { int on, ff, rec, lat, fb, pred, gather; int act_slab, act_box, act_cap, act_dense; int tm_bw, tm_bh; };  // act_*: k_g_kw_activate (row segments); tm_*: its input box through a TMA tensor map
  CUtensorMap act_tmap;            // layer 0's input array [agent][y][x] for k_g_kw_activate's tiled loads
  std::vector<GridSizes> gridk;
  std::vector<bidi::PersistShape> persist;  // per layer: the whole iterative coder as one cooperative launch (bidi_persist.cuh)
  unsigned* d_persist_ctr = nullptr;        // its grid barrier's arrival counters, one per layer
  int persist_mode = 1;            // option "persist_coder"
  int grid_mode = 1;               // 0 off, 1 layers at least 16 units wide that are not tiled, 2 every layer
  EngineDev P;                     // kernel argument
  // device memory
  float* d_blob = nullptr; float* d_in0 = nullptr; float* d_pred = nullptr;
  float* d_rewards = nullptr; float* d_noise_u = nullptr; float* d_noise_v = nullptr; float* d_col_actions = nullptr;
  float* d_frames = nullptr; float* d_frame_rewards = nullptr; int n_frames = 0;
  float* d_frame_noise_u = nullptr; float* d_frame_noise_v = nullptr; int n_noise_frames = 0;  // bidi_load_frame_noise
  int* d_tables = nullptr; LayerDev* d_layers = nullptr;
  long long device_bytes = 0;
  // pinned staging
  float* h_in = nullptr; float* h_pred = nullptr; float* h_rewards = nullptr; float* h_noise = nullptr; float* h_actions = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;                     // second branch of the step graph (coder learning, see record_step)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool overlap_learn = true;
  std::vector<int> dense_rows;                     // per layer: its feed-forward/recurrent weights are currently laid out as rows (apply_layout)
  bool state_touched = false;                      // some state beyond the zero/threshold initialisation may be in the blobs
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_in = nullptr;                     // the last bidi_set_inputs' host->device copy has read the host buffer
  // one graph per value of the learn flag
  cudaGraphExec_t graphs[3] = {nullptr, nullptr, nullptr};   // learn off, learn on, simStepGenerate
  int graph_kernels[3] = {0, 0, 0};
  // simStepGenerate
  int gen_per_agent = 0; float* d_gen = nullptr; float* d_gen_scale = nullptr; float* h_gen = nullptr;
  long long launches = 0;
  unsigned long long step_index = 0, noise_seed = 0x5DEECE66DULL;
  long long first_agent = 0;                       // global index of agent 0 (bidi_set_noise_stream)
  // batched headless environment (bidi_env_*, bidi_env.cuh)
  int env_kind = 0; float* d_env = nullptr; float* d_env_noise = nullptr; int env_steps = 0;
  // feedback taps (bidi_set_feedback)
  int fb_n = 0; int* d_fb_src = nullptr; int* d_fb_dst = nullptr; float fb_scale = 0, fb_bias = 0, fb_std = 0; int fb_reward = 0;
  // kernel profiling (bidi_profile_select): external events around the launches of one kernel
  std::string profile_name; std::vector<std::pair<cudaEvent_t, cudaEvent_t>> profile_events;
  double profile_ms = 0; long long profile_launches = 0;
  // host mirror of ONE agent's blob for set/get_array
  std::vector<float> mirror; int mirror_agent = -1; bool mirror_dirty = false;
  // the 16-cell column kernel gathers its inputs into shared memory only; the BIDI_COL_INPUT planes are regathered
  // on the host when somebody asks for them (stale since the last step, unless uploaded since)
  bool col_inputs_stale = false; std::vector<char> col_inputs_set;
  // spatial strip split of one oversized layer (bidi_strip.inl)
  int strip_part = 0, strip_n = 1;
  int strip_span[BIDI_SPAN_COUNT][2] = {};
  struct StripXfer { int point, peer; long long off, count; };
  std::vector<StripXfer> strip_send, strip_recv;   // per exchange point, peer: rows of one plane
  std::vector<bidi_engine*> strip_peers;           // same-process transport, indexed by part
  void* nccl_comm = nullptr;                       // one-process-per-GPU transport
  // peer-memory transport (bidi_strip_link_ipc): device tables for k_strip_push
  bool ipc_linked = false;
  std::vector<void*> ipc_opened;                   // peer mappings to close
  unsigned* d_flags = nullptr; unsigned* d_epoch = nullptr;
  float** d_peer_blob = nullptr; unsigned** d_peer_flags = nullptr;
  bidi::StripXferDev* d_sends = nullptr; int* d_notify = nullptr; int* d_await = nullptr;
  struct PushArgs { int send0, n_sends, notify0, n_notify, await0, n_await; } push[5] = {};
  // sdr::QPRSDR: the engine runs the ITERATIVE hierarchy (cfg.variant is normalised to it) plus the Q network
  bool qnet = false;
  bool csrl = false;                                // deep::CSRL: the ITERATIVE coders + one deep::SDRRL unit per node (bidi_csrl.cuh)
  std::vector<int> csrl_is_action; int* d_csrl_tables = nullptr;
  int q_half = 0, q_top = 0, n_noise = 0;           // n_noise: floats per agent in noise_u / noise_v
  std::vector<std::vector<char>> q_keep;            // per layer: [slot][unit] the reference keeps the connection
  std::vector<int> q_act_index, q_a_input;
  float* d_qmask = nullptr; int* d_qtables = nullptr;
  bool use_tma_box = true;                         // option "tma_box": k_g_kw_activate's input box through a TMA tensor map
  bool strip_use_graph = true;                     // record the split step (NCCL exchanges included) into a CUDA graph
  cudaEvent_t strip_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t strip_done = nullptr;
  std::string err;
};

namespace {

int fail(bidi_engine* h, int code, const std::string& msg) {
  g_last_error = msg;
  if (h) h->err = msg;
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(h, BIDI_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));              \
  } while (0)

// ------------------------------------------------------------------ geometry
// Window centres exactly as the reference computes them (sdr/RSDR.cpp:33-41):
// fp32 ratio, fp32 product, std::round (half away from zero), truncation to int.
void build_win(WinHost& w, int uw, int uh, int tw, int th, int r, bool identity, bool excl) {
  WinDev& d = w.d;
  memset(&d, 0, sizeof(d));
  d.uw = uw; d.uh = uh; d.tw = tw; d.th = th; d.r = r; d.excl = excl ? 1 : 0; d.U = uw * uh;
  d.w_off = -1; d.tr_off = -1;
  w.cx.assign(uw, 0); w.cy.assign(uh, 0);
  const float rx = static_cast<float>(tw) / static_cast<float>(uw);
  const float ry = static_cast<float>(th) / static_cast<float>(uh);
  for (int x = 0; x < uw; x++) w.cx[x] = identity ? x : static_cast<int>(std::round(x * rx));
  for (int y = 0; y < uh; y++) w.cy[y] = identity ? y : static_cast<int>(std::round(y * ry));
  w.gxlo.assign(tw, 0); w.gxhi.assign(tw, -1); w.gylo.assign(th, 0); w.gyhi.assign(th, -1);
  if (r < 0) { d.dxn = d.dyn = d.nslots = 0; return; }
  const int span = 2 * r + 1;
  d.absx = span >= tw; d.absy = span >= th;
  d.dxn = d.absx ? tw : span; d.dyn = d.absy ? th : span;
  d.nslots = d.dxn * d.dyn;
  auto ranges = [r](const std::vector<int>& c, int tn, std::vector<int>& lo, std::vector<int>& hi) {
    for (int t = 0; t < tn; t++) {
      int l = (int)c.size(), hgh = -1;
      for (int u = 0; u < (int)c.size(); u++)
        if (std::abs(c[u] - t) <= r) { l = std::min(l, u); hgh = std::max(hgh, u); }
      lo[t] = l; hi[t] = hgh;
    }
  };
  ranges(w.cx, tw, w.gxlo, w.gxhi);
  ranges(w.cy, th, w.gylo, w.gyhi);
  int mx = 0, my = 0;
  for (int t = 0; t < tw; t++) mx = std::max(mx, w.gxhi[t] - w.gxlo[t] + 1);
  for (int t = 0; t < th; t++) my = std::max(my, w.gyhi[t] - w.gylo[t] + 1);
  d.max_cand = mx * my;
}

// Host twin of bidi::for_each_conn: f(slot, target) in the reference's list order.
template <class F> void for_each_conn_host(const WinHost& w, int ux, int uy, F&& f) {
  const WinDev& d = w.d;
  if (d.r < 0) return;
  const int cx = w.cx[ux], cy = w.cy[uy];
  for (int sx = 0; sx < d.dxn; sx++) {
    const int tx = d.absx ? sx : cx - d.r + sx;
    if (tx < 0 || tx >= d.tw || std::abs(tx - cx) > d.r) continue;
    for (int sy = 0; sy < d.dyn; sy++) {
      const int ty = d.absy ? sy : cy - d.r + sy;
      if (ty < 0 || ty >= d.th || std::abs(ty - cy) > d.r) continue;
      if (d.excl && tx == cx && ty == cy) continue;
      f(sx * d.dyn + sy, tx + ty * d.tw);
    }
  }
}
long long conn_total(const WinHost& w) {
  long long n = 0;
  for (int uy = 0; uy < w.d.uh; uy++)
    for (int ux = 0; ux < w.d.uw; ux++) for_each_conn_host(w, ux, uy, [&](int, int) { n++; });
  return n;
}

struct Allocator {
  long long top = 0;
  long long take(long long n) { long long o = top; top += (n + 31) / 32 * 32; return o; }
};

bool is_iter(int v) { return v != BIDI_KWINNERS; }
bool is_traced(int v) { return v == BIDI_NEO_HIERARCHY || v == BIDI_NEO_AGENT; }

// ------------------------------------------------------------------ array table
// How a serialised array (include/bidinet_b200.h enum bidi_array) maps to storage.
struct ArrayRef {
  enum Kind { NONE, PLANE, LIST, DENSE_IN, DENSE_PRED, COLUMN, QLIST } kind = NONE;
  long long off = 0, len = 0;   // PLANE: blob offset and length
  const WinHost* win = nullptr; // LIST
  bool trace = false;
  int which = 0, layer = 0;     // COLUMN: a field of the column records
};

ArrayRef find_array(const bidi_engine* h, int which, int layer) {
  ArrayRef r;
  const int L = h->L, V = h->cfg.variant;
  if (layer < 0 || layer > L) return r;
  auto plane = [&](long long off, long long len) { r.kind = ArrayRef::PLANE; r.off = off; r.len = len; return r; };
  auto column = [&](long long len) { r.kind = ArrayRef::COLUMN; r.which = which; r.layer = layer; r.len = len; return r; };
  auto list = [&](const WinHost& w, bool tr) {
    if (w.d.r < 0 || (tr ? w.d.tr_off : w.d.w_off) < 0) return r;
    r.kind = ArrayRef::LIST; r.win = &w; r.trace = tr; r.len = conn_total(w); return r;
  };
  if (which == BIDI_PREDICTION) {
    if (V != BIDI_KWINNERS || layer != 0) return r;
    r.kind = ArrayRef::DENSE_PRED; r.len = h->n_in; return r;
  }
  const LayerDev& D = h->layers[layer];
  const LayerHost& H = h->hl[layer];
  if (which <= BIDI_REC_TRACE) {
    if (layer >= L) return r;
    switch (which) {
      case BIDI_VIS_INPUT:
        if (layer == 0) { r.kind = ArrayRef::DENSE_IN; r.len = h->n_in; return r; }
        return plane(D.vis_input, D.nv);
      case BIDI_VIS_RECON: return plane(D.vis_recon, D.nv);
      case BIDI_HID_THRESHOLD: return plane(D.thr, D.nh);
      case BIDI_HID_STATE: return plane(D.state, D.nh);
      case BIDI_HID_STATE_PREV: return plane(D.state_prev, D.nh);
      case BIDI_HID_ACTIVATION: return plane(D.act, D.nh);
      case BIDI_HID_RECON: return plane(D.recon, D.nh);
      case BIDI_HID_SPIKE: return is_iter(V) ? plane(D.spike, D.nh) : r;
      case BIDI_HID_SPIKE_PREV: return is_iter(V) ? plane(D.spike_prev, D.nh) : r;
      case BIDI_FF_W: return list(H.ff, false);
      case BIDI_REC_W: return list(H.rec, false);
      case BIDI_LAT_W: return is_iter(V) ? list(H.lat, false) : r;
      case BIDI_FF_TRACE: return (is_traced(V) || h->csrl) ? list(H.ff, true) : r;
      case BIDI_REC_TRACE: return (is_traced(V) || h->csrl) ? list(H.rec, true) : r;
    }
    return r;
  }
  if (layer == L && V == BIDI_KWINNERS) return r;
  switch (which) {
    case BIDI_P_STATE:
      if (layer == L) { r.kind = ArrayRef::DENSE_PRED; r.len = h->n_in; return r; }
      return plane(D.p_state, D.pn);
    case BIDI_P_STATE_PREV: return plane(D.p_state_prev, D.pn);
    case BIDI_P_ACTIVATION: return plane(D.p_act, D.pn);
    case BIDI_P_ACTIVATION_PREV: return plane(D.p_act_prev, D.pn);
    case BIDI_P_BASELINE: return layer == L ? r : plane(D.p_baseline, D.pn);
    case BIDI_FB_W: return list(H.fb, false);
    case BIDI_PRED_W: return layer == L ? r : list(H.pred, false);
  }
  if (h->csrl) {  // deep::CSRL: traces of the prediction weights, _localReward, and the SDRRL unit of every node (plain planes)
    const SdrrlDev& S = D.sd;
    const long long nc = (long long)S.n * S.cells, nan = (long long)S.n * S.na;
    switch (which) {
      case BIDI_FB_TRACE: return list(H.fb, true);
      case BIDI_PRED_TRACE: return layer == L ? r : list(H.pred, true);
      case BIDI_P_LOCAL_REWARD: return plane(D.p_mod, D.pn);
      case BIDI_REWARD_OFFSETS: return layer == L - 1 ? plane(h->P.csrl_reward_offsets, h->layers[L - 1].nh) : r;
      case BIDI_COL_INPUT: return plane(S.inputs, S.tot_in);
      case BIDI_COL_RECON_ERR: return plane(S.recon_err, S.tot_in);
      case BIDI_COL_FF_W: return plane(S.ff_w, (long long)S.tot_in * S.cells);
      case BIDI_COL_LAT_W: return plane(S.lat_w, nc * S.cells);
      case BIDI_COL_THRESHOLD: return plane(S.thr, nc);
      case BIDI_COL_ACTIVATION: return plane(S.act, nc);
      case BIDI_COL_SPIKE: return plane(S.spike, nc);
      case BIDI_COL_SPIKE_PREV: return plane(S.spike_prev, nc);
      case BIDI_COL_STATE: return plane(S.state, nc);
      case BIDI_COL_CELL_ACT_STATE: return plane(S.cell_as, nc);
      case BIDI_COL_CELL_ACT_ERROR: return plane(S.cell_ae, nc);
      case BIDI_COL_Q_W: return plane(S.q_w, nc);
      case BIDI_COL_Q_TRACE: return plane(S.q_tr, nc);
      case BIDI_COL_ACT_STATE: return plane(S.a_state, nan);
      case BIDI_COL_ACT_EXPL: return plane(S.a_expl, nan);
      case BIDI_COL_ACT_ERROR: return plane(S.a_err, nan);
      case BIDI_COL_ACT_W: return plane(S.a_w, nc * S.na);
      case BIDI_COL_ACT_TRACE: return plane(S.a_tr, nc * S.na);
      case BIDI_COL_PREV_VALUE: return plane(S.prev_value, S.n);
      case BIDI_COL_AVG_SURPRISE: return plane(S.avg_surprise, S.n);
    }
    return r;
  }
  if (h->qnet && which >= BIDI_Q_STATE) {
    if (layer < L) {
      switch (which) {
        case BIDI_Q_STATE: return plane(D.q_state, D.nh);
        case BIDI_Q_ERROR: return plane(D.q_err, D.nh);
        case BIDI_Q_BIAS_W: return plane(D.q_bias_w, D.nh);
        case BIDI_Q_BIAS_TRACE: return plane(D.q_bias_tr, D.nh);
        case BIDI_Q_FF_W: case BIDI_Q_FF_TRACE: {
          r.kind = ArrayRef::QLIST; r.win = &H.ff; r.trace = which == BIDI_Q_FF_TRACE; r.layer = layer;
          long long n = 0;
          for (char c : h->q_keep[layer]) n += c;
          r.len = n;
          return r;
        }
      }
    }
    if (layer != 0) return r;
    switch (which) {
      case BIDI_QCONN_W: return plane(h->P.q_qc_w, h->q_top);
      case BIDI_QCONN_TRACE: return plane(h->P.q_qc_tr, h->q_top);
      case BIDI_ACT_PREDICTED: return plane(h->P.q_act, 2 * h->q_half);
      case BIDI_ACT_DERIVE: return plane(h->P.q_act + 2 * h->q_half, 2 * h->q_half);
      case BIDI_ACT_EXPL: return plane(h->P.q_act + 4 * h->q_half, 2 * h->q_half);
      case BIDI_ACT_ERROR: return plane(h->P.q_act + 6 * h->q_half, 2 * h->q_half);
      case BIDI_Q_PREV_VALUE: return plane(h->P.q_prev, 1);
    }
    return r;
  }
  if (V != BIDI_NEO_AGENT) return r;
  const ColumnsDev& C = D.col;
  const long long tot = C.tot_in, nc = (long long)C.n * C.cells;
  switch (which) {
    case BIDI_COL_INPUT: return plane(C.inputs, tot);
    case BIDI_COL_ACTIVATION: return plane(C.act, nc);
    case BIDI_COL_SPIKE: return plane(C.spike, nc);
    case BIDI_COL_STATE: return plane(C.state, nc);
    case BIDI_COL_ACT_STATE: return plane(C.a_state, (long long)C.n * 3);
    case BIDI_COL_ACT_EXPL: return plane(C.a_expl, (long long)C.n * 3);
    case BIDI_COL_RECON_ERR: return column(tot);
    case BIDI_COL_FF_W: return column(tot * C.cells);
    case BIDI_COL_LAT_W: return column(nc * C.cells);
    case BIDI_COL_THRESHOLD: case BIDI_COL_SPIKE_PREV: case BIDI_COL_Q_W: case BIDI_COL_Q_TRACE: return column(nc);
    case BIDI_COL_ACT_W: case BIDI_COL_ACT_TRACE: return column(nc * 3);
    case BIDI_COL_PREV_VALUE: return column(C.n);
  }
  return r;
}

// ------------------------------------------------------------------ mirror
int mirror_flush(bidi_engine* h) {
  if (h->mirror_agent >= 0 && h->mirror_dirty) {
    CU(cudaMemcpyAsync(h->d_blob + (long long)h->mirror_agent * h->stride, h->mirror.data(),
                       sizeof(float) * h->stride, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->mirror_dirty = false;
  }
  return BIDI_OK;
}
int mirror_load(bidi_engine* h, int agent) {
  if (h->mirror_agent == agent) return BIDI_OK;
  int rc = mirror_flush(h);
  if (rc) return rc;
  h->mirror.resize((size_t)h->stride);
  CU(cudaMemcpyAsync(h->mirror.data(), h->d_blob + (long long)agent * h->stride, sizeof(float) * h->stride,
                     cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->mirror_agent = agent;
  h->mirror_dirty = false;
  if (h->col_inputs_stale && !h->col_inputs_set[agent])
    for (int l = 0; l <= h->L; l++) {  // neo/Agent.cpp:159-169: the inputs the columns of the last step consumed
      const ColumnsDev& C = h->layers[l].col;
      if (!C.planar || C.n == 0) continue;
      const std::vector<int>& src = h->hl[l].col.in_src;
      for (int j = 0; j < C.tot_in; j++) h->mirror[C.inputs + j] = h->mirror[src[j]];
    }
  return BIDI_OK;
}
// device state is about to change: the mirror must be written back and dropped
int mirror_release(bidi_engine* h) {
  int rc = mirror_flush(h);
  h->mirror_agent = -1;
  return rc;
}

// Serialised list <-> slot planes inside a host blob image.
template <class F> void walk_list(const WinHost& w, bool trace, float* blob, F&& f) {
  float* base = blob + (trace ? w.d.tr_off : w.d.w_off);
  const long long U = w.d.U, ss = w.d.row_pitch ? 1 : U, us = w.d.row_pitch ? w.d.row_pitch : 1;
  for (int uy = 0; uy < w.d.uh; uy++)
    for (int ux = 0; ux < w.d.uw; ux++) {
      const long long u = ux + (long long)uy * w.d.uw;
      for_each_conn_host(w, ux, uy, [&](int s, int) { f(base[(long long)s * ss + u * us]); });
    }
}

// The Q network's kept connections (sdr/QPRSDR.cpp:72-81) <-> its slot planes.
template <class F> void walk_qlist(const bidi_engine* h, int layer, bool trace, float* blob, F&& f) {
  const WinHost& w = h->hl[layer].ff;
  const LayerDev& D = h->layers[layer];
  float* base = blob + (trace ? D.q_tr_off : D.q_w_off);
  const std::vector<char>& keep = h->q_keep[layer];
  const long long U = w.d.U;
  for (int uy = 0; uy < w.d.uh; uy++)
    for (int ux = 0; ux < w.d.uw; ux++) {
      const long long u = ux + (long long)uy * w.d.uw;
      for_each_conn_host(w, ux, uy, [&](int s, int) { if (keep[(size_t)s * U + u]) f(base[(long long)s * U + u]); });
    }
}

// A record-resident column array in the reference's order ([node][cell][input] ...).
template <class F> void walk_column(const bidi_engine* h, int which, int layer, float* blob, F&& f) {
  const ColumnsDev& C = h->layers[layer].col;
  const ColumnsHost& H = h->hl[layer].col;
  for (int node = 0; node < C.n; node++) {
    const ColRec r = col_rec(C, H, blob, node);
    switch (which) {
      case BIDI_COL_RECON_ERR: for (int j = 0; j < r.nin; j++) f(r.err[r.es * j]); break;
      case BIDI_COL_FF_W: for (int i = 0; i < C.cells; i++) for (int j = 0; j < r.nin; j++) f(r.w(i, j)); break;
      case BIDI_COL_LAT_W: for (int i = 0; i < C.cells; i++) for (int j = 0; j < C.cells; j++) f(r.lat(i, j)); break;
      case BIDI_COL_THRESHOLD: for (int i = 0; i < C.cells; i++) f(r.thr[r.cs * i]); break;
      case BIDI_COL_SPIKE_PREV: for (int i = 0; i < C.cells; i++) f(r.sprev[r.cs * i]); break;
      case BIDI_COL_Q_W: for (int i = 0; i < C.cells; i++) f(r.q_w[r.cs * i]); break;
      case BIDI_COL_Q_TRACE: for (int i = 0; i < C.cells; i++) f(r.q_tr[r.cs * i]); break;
      case BIDI_COL_ACT_W: for (int a = 0; a < 3; a++) for (int i = 0; i < C.cells; i++) f(r.a_w[r.es * (a * r.Cq + i)]); break;
      case BIDI_COL_ACT_TRACE: for (int a = 0; a < 3; a++) for (int i = 0; i < C.cells; i++) f(r.a_tr[r.es * (a * r.Cq + i)]); break;
      case BIDI_COL_PREV_VALUE: f(r.scal[0]); break;
    }
  }
}

// ------------------------------------------------------------------ launches
inline unsigned grid_for(long long n, int block) { return (unsigned)((n + block - 1) / block); }
constexpr int kBlock = 128;

struct Recorder {
  bidi_engine* h;
  int kernels = 0;
  cudaStream_t st = nullptr;  // where launches go: the engine's stream, or the side branch while it is open
  cudaStream_t s() const { return st ? st : h->stream; }
  bool capturing = false;  // inside a stream capture: the step may fork into a second branch
  // bracket a launch with external event records when its kernel is the one being profiled
  bool profiled(const char* name) const { return !h->profile_name.empty() && h->profile_name == name; }
  void before(const char* name) {
    if (!profiled(name)) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    h->profile_events.push_back({a, b});
    cudaEventRecordWithFlags(a, s(), cudaEventRecordExternal);
  }
  void after(const char* name) {
    if (profiled(name)) cudaEventRecordWithFlags(h->profile_events.back().second, s(), cudaEventRecordExternal);
  }
  template <class K, class... Args> void launch(const char* name, K kernel, long long threads, Args... args) {
    before(name);
    const int block = kBlock;
    kernel<<<grid_for(threads, block), block, 0, s()>>>(h->P, args...);
    after(name);
    kernels++;
  }
};

// One column pair per warp: shared memory, not warps, limits how many pairs an SM holds
// (the pair record is ~18-22 KB), so giving every pair a whole warp doubles the warps
// that hide each other's latency at no shared-memory cost; lanes 16..31 idle only
// during the 16-cell sweep and share the reconstruction and learning loops.
int columns_pairs_per_warp(const ColumnsDev& C) {
#ifdef BIDI_COL_PPW2
  return C.cells <= 16 ? 2 : 1;
#else
  return 1;
#endif
}
template <class K, class... Args>
void launch_tile(Recorder& R, const char* name, K kernel, long long tiles, size_t smem, Args... args) {
  R.before(name);
  kernel<<<(unsigned)tiles, TILE_U * TILE_G, smem, R.s()>>>(R.h->P, args...);
  R.after(name);
  R.kernels++;
}
inline long long tiles_of(long long agents, int units) { return agents * ((units + TILE_U - 1) / TILE_U); }
inline size_t tile_bytes(int slots) { return sizeof(float) * TILE_U * (size_t)std::max(1, slots); }
inline int nslots_of(const WinDev& w) { return w.r < 0 ? 0 : w.nslots; }

size_t columns_smem_bytes(const ColumnsDev& C) {
  const int ppw = columns_pairs_per_warp(C);
  return sizeof(float) * BIDI_COL_WARPS * ppw * (size_t)bidi::column_slab_floats(C) + 8 * BIDI_COL_WARPS;
}
void record_columns(Recorder& R, int l) {
  bidi_engine* h = R.h;
  const ColumnsDev& C = h->layers[l].col;
  const int ppw = columns_pairs_per_warp(C);
  const int n_pairs = (C.n + 1) / 2;
  const long long warps = (long long)h->A * ((n_pairs + ppw - 1) / ppw);
  const size_t smem = columns_smem_bytes(C);
  R.before("columns");
  // x = pair (fastest: consecutive CTAs are one agent's pairs), (y, z) = agent
  const unsigned ay = (unsigned)std::min(h->A, 32768);
  const dim3 pa((unsigned)n_pairs, ay, (unsigned)((h->A + ay - 1) / ay));
  const ColPairTab& T = h->hl[l].col.tab;
  if (C.planar && C.rh == 64) bidi::k_columns16<64><<<pa, 32, smem, h->stream>>>(h->P, h->layers[l], l, T);
  else if (C.planar && C.rh == 32) bidi::k_columns16<32><<<pa, 32, smem, h->stream>>>(h->P, h->layers[l], l, T);
  else if (C.planar && C.max_in <= 28) bidi::k_columns16<0, true><<<pa, 32, smem, h->stream>>>(h->P, h->layers[l], l, T);
  else if (C.planar) bidi::k_columns16<0><<<pa, 32, smem, h->stream>>>(h->P, h->layers[l], l, T);
  else if (ppw == 2) bidi::k_columns<16><<<grid_for(warps, BIDI_COL_WARPS), BIDI_COL_WARPS * 32, smem, h->stream>>>(h->P, h->layers[l], l);
  else bidi::k_columns<32><<<grid_for(warps, BIDI_COL_WARPS), BIDI_COL_WARPS * 32, smem, h->stream>>>(h->P, h->layers[l], l);
  R.after("columns");
  R.kernels++;
}

template <class K, class... Args>
void launch_grid(Recorder& R, const char* name, K kernel, long long blocks, size_t smem_floats, Args... args) {
  if (blocks <= 0) return;
  R.before(name);
  kernel<<<(unsigned)blocks, GRID_TX * GRID_TY, sizeof(float) * smem_floats, R.s()>>>(R.h->P, args...);
  R.after(name);
  R.kernels++;
}
// Tensor map of the layer-0 input array [agent][in_height][in_width] with box (bw x bh x 1), zero fill outside.  The
// encoder is a driver entry point fetched at run time (no link-time libcuda dependency).  false: not available.
bool encode_input_tmap(bidi_engine* h, int bw, int bh) {
  typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Encode encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
      cudaGetLastError();
      return false;
    }
    encode = reinterpret_cast<Encode>(fn);
  }
  const cuuint64_t dims[3] = {(cuuint64_t)h->cfg.in_width, (cuuint64_t)h->cfg.in_height, (cuuint64_t)h->A};
  const cuuint64_t strides[2] = {(cuuint64_t)h->cfg.in_width * 4, (cuuint64_t)h->n_in * 4};
  const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return encode(&h->act_tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, h->d_in0, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
inline size_t act_smem_bytes(const bidi_engine::GridSizes& G) { return sizeof(float) * (size_t)(G.act_slab + G.act_box + 2 * G.act_cap) + 16; }
// k_g_kw_activate over the hidden rows [Y.row0, Y.row1): one warp per 32-unit row segment
void launch_kw_activate(bidi_engine* h, cudaStream_t st, const bidi_engine::GridSizes& G, const LayerDev& Y, int l) {
  const long long blocks = (long long)h->A * (Y.row1 - Y.row0) * ((Y.hw + 31) / 32);
  if (blocks <= 0) return;
  const size_t smem = act_smem_bytes(G);
  const int dyn = (G.act_dense && Y.ff.dxn == Y.ff.dyn && !Y.ff.absx && !Y.ff.absy) ? Y.ff.dyn : 0;
#define BIDI_ACT(D) bidi::k_g_kw_activate<D><<<(unsigned)blocks, 32, smem, st>>>(h->P, Y, l, G.act_dense, G.act_slab, G.act_box, G.act_cap, h->act_tmap, G.tm_bw, G.tm_bh)
  switch (dyn) {
    case 5: BIDI_ACT(5); break;
    case 7: BIDI_ACT(7); break;
    case 9: BIDI_ACT(9); break;
    case 13: BIDI_ACT(13); break;
    case 17: BIDI_ACT(17); break;
    default: BIDI_ACT(0); break;
  }
#undef BIDI_ACT
}
// ---- bidi_persist.cuh: which layers of a single ITERATIVE agent run their whole coder as one cooperative launch
template <int DYN> int persist_prepare(bidi_engine* h, size_t smem) {
  if (smem > 48 * 1024) CU(cudaFuncSetAttribute(bidi::k_ic_layer_persist<DYN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return BIDI_OK;
}
int select_persist(bidi_engine* h) {
  h->persist.assign(h->L, bidi::PersistShape{});
  if (h->cfg.variant != BIDI_ITERATIVE || h->A != 1 || !h->persist_mode || h->csrl || h->strip_n > 1) return BIDI_OK;
  int sms = 0, coop = 0;
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
  CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
  if (!coop || sms < 1) return BIDI_OK;
  size_t most = 0;
  for (int l = 0; l < h->L; l++) {
    const LayerDev& D = h->layers[l];
    const LayerHost& H = h->hl[l];
    if (D.lat.r < 0 || D.nh < 256 || h->coder_cta[l] == 2) continue;  // small or dense layers have their fused one-CTA kernels
    const int nslots = D.ff.nslots + (D.rec.r >= 0 ? D.rec.nslots : 0) + D.lat.nslots;
    bidi::PersistShape S{};
    // hidden tiles: as many as there are SMs at most -- but at least ~16 units each, the grid barrier costs with the number
    // of CTAs --, weights of a tile within ~150 KB, at most 128 units (two sums per unit on 256 threads)
    const int cap = std::min(sms, std::max(8, D.nh / 16));
    const int vcap = std::min(sms, std::max(cap, (D.nv + 127) / 128));
    int best = 0;
    for (int tx : {2, 4, 8, 16, 32})
      for (int ty : {1, 2, 4, 8, 16}) {
        const int ut = tx * ty, nx = (D.hw + tx - 1) / tx, ny = (D.hh + ty - 1) / ty;
        if (ut > 128 || (size_t)ut * nslots * 4 > 150 * 1024 || nx * ny > cap) continue;
        const int score = nx * ny * 64 - std::abs(tx - ty);  // most tiles, then the squarest
        if (score > best) { best = score; S.htx = tx; S.hty = ty; S.hnx = nx; S.hny = ny; }
      }
    if (!best) continue;
    best = 0;
    for (int tx : {2, 4, 8, 16, 32})
      for (int ty : {1, 2, 4, 8, 16}) {
        const int nx = (D.vw + tx - 1) / tx, ny = (D.vh + ty - 1) / ty;
        if (tx * ty > 128 || nx * ny > vcap) continue;
        const int score = nx * ny * 64 - std::abs(tx - ty);
        if (score > best) { best = score; S.vtx = tx; S.vty = ty; S.vnx = nx; S.vny = ny; }
      }
    if (!best) continue;
    S.ut = S.htx * S.hty;
    auto box = [](const WinHost& w, int tx, int ty) {
      if (w.d.r < 0) return 0;
      int bw = w.d.tw, bh = w.d.th;
      if (!w.d.absx) { bw = 0; for (int u = 0; u < w.d.uw; u += tx) bw = std::max(bw, w.cx[std::min(u + tx, w.d.uw) - 1] - w.cx[u] + 2 * w.d.r + 1); }
      if (!w.d.absy) { bh = 0; for (int u = 0; u < w.d.uh; u += ty) bh = std::max(bh, w.cy[std::min(u + ty, w.d.uh) - 1] - w.cy[u] + 2 * w.d.r + 1); }
      return bw * bh;
    };
    auto cand = [](const WinHost& w, int tx, int ty) {
      if (w.d.r < 0) return 0;
      int bw = 0, bh = 0;
      for (int t0 = 0; t0 < w.d.tw; t0 += tx) {
        int lo = w.d.uw, hi = -1;
        for (int t = t0; t < std::min(t0 + tx, w.d.tw); t++) { lo = std::min(lo, w.gxlo[t]); hi = std::max(hi, w.gxhi[t]); }
        bw = std::max(bw, hi - lo + 1);
      }
      for (int t0 = 0; t0 < w.d.th; t0 += ty) {
        int lo = w.d.uh, hi = -1;
        for (int t = t0; t < std::min(t0 + ty, w.d.th); t++) { lo = std::min(lo, w.gylo[t]); hi = std::max(hi, w.gyhi[t]); }
        bh = std::max(bh, hi - lo + 1);
      }
      return std::max(0, bw) * std::max(0, bh);
    };
    S.box_ff = (box(H.ff, S.htx, S.hty) + 3) & ~3; S.box_rec = (box(H.rec, S.htx, S.hty) + 3) & ~3; S.box_lat = (box(H.lat, S.htx, S.hty) + 3) & ~3;
    S.cap_ff = cand(H.ff, S.vtx, S.vty) + 4; S.cap_rec = cand(H.rec, S.htx, S.hty) + 4;
    S.blocks = std::max(S.hnx * S.hny, S.vnx * S.vny);
    const int dy = D.ff.dyn;
    S.dyn = (dy == D.lat.dyn && (D.rec.r < 0 || dy == D.rec.dyn) && (dy == 7 || dy == 13)) ? dy : 0;
    S.smem_floats = (long long)nslots * S.ut + S.box_ff + S.box_rec + S.box_lat + S.ut + 3LL * (S.cap_ff + S.cap_rec) + 2LL * (D.hw + D.hh) + 8;
    if (S.smem_floats * 4 > 200 * 1024) continue;
    S.on = 1;
    most = std::max(most, (size_t)S.smem_floats * 4);
    h->persist[l] = S;
  }
  if (most > 0) {
    int rc = persist_prepare<0>(h, most); if (rc) return rc;
    rc = persist_prepare<7>(h, most); if (rc) return rc;
    rc = persist_prepare<13>(h, most); if (rc) return rc;
    if (!h->d_persist_ctr) CU(cudaMalloc(&h->d_persist_ctr, sizeof(unsigned) * BIDI_MAX_LAYERS));
  }
  return BIDI_OK;
}
// one cooperative launch: the up pass of layer l
int launch_persist(bidi_engine* h, cudaStream_t st, int l) {
  const bidi::PersistShape& S = h->persist[l];
  unsigned* ctr = h->d_persist_ctr + l;
  CU(cudaMemsetAsync(ctr, 0, sizeof(unsigned), st));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)S.blocks); cfg.blockDim = dim3(BIDI_PERSIST_THREADS);
  cfg.dynamicSmemBytes = (size_t)S.smem_floats * 4; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  switch (S.dyn) {
    case 7: CU(cudaLaunchKernelEx(&cfg, bidi::k_ic_layer_persist<7>, h->P, h->layers[l], l, S, ctr)); break;
    case 13: CU(cudaLaunchKernelEx(&cfg, bidi::k_ic_layer_persist<13>, h->P, h->layers[l], l, S, ctr)); break;
    default: CU(cudaLaunchKernelEx(&cfg, bidi::k_ic_layer_persist<0>, h->P, h->layers[l], l, S, ctr)); break;
  }
  return BIDI_OK;
}
// which layers run the 2-D tile kernels, and how much shared memory they stage
int select_grid_kernels(bidi_engine* h) {
  h->gridk.assign(h->L, bidi_engine::GridSizes{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0});
  memset(&h->act_tmap, 0, sizeof(h->act_tmap));
  if (h->cfg.variant == BIDI_NEO_AGENT || h->grid_mode == 0) return BIDI_OK;
  size_t most = 0;
  for (int l = 0; l < h->L; l++) {
    const LayerHost& H = h->hl[l];
    bidi_engine::GridSizes g;
    g.ff = bidi::grid_box_floats(H.ff.d, H.ff.cx.data(), H.ff.cy.data());
    g.rec = bidi::grid_box_floats(H.rec.d, H.rec.cx.data(), H.rec.cy.data());
    g.lat = bidi::grid_box_floats(H.lat.d, H.lat.cx.data(), H.lat.cy.data());
    g.fb = bidi::grid_box_floats(H.fb.d, H.fb.cx.data(), H.fb.cy.data());
    g.pred = bidi::grid_box_floats(H.pred.d, H.pred.cx.data(), H.pred.cy.data());
    g.gather = std::max(bidi::grid_gather_floats(H.ff.d, H.ff.gxlo.data(), H.ff.gxhi.data(), H.ff.gylo.data(), H.ff.gyhi.data()),
                        bidi::grid_gather_floats(H.rec.d, H.rec.gxlo.data(), H.rec.gxhi.data(), H.rec.gylo.data(), H.rec.gyhi.data()));
    size_t need = sizeof(float) * (size_t)std::max({g.ff + g.rec + (h->cfg.variant == BIDI_KWINNERS ? 0 : g.lat), g.lat, g.fb + g.pred, 3 * g.gather});
    g.act_slab = g.act_box = g.act_cap = g.act_dense = 0;
    if (h->cfg.variant == BIDI_KWINNERS) {
      // k_g_kw_activate: layer 0 reads a dense input (rows staged in shared memory); a higher layer reads the binary state below
      g.act_dense = l == 0 ? 1 : 0;
      const int fbox = bidi::grid_rowbox_floats(H.ff.d, H.ff.cx.data()), rbox = bidi::grid_rowbox_floats(H.rec.d, H.rec.cx.data());
      g.act_slab = g.act_dense ? 32 * kw_pitch(H.ff.d.nslots) : 0;
      g.act_box = g.act_dense ? (fbox + 3) / 4 * 4 : 0;
      g.tm_bw = g.tm_bh = 0;
      if (g.act_dense && h->use_tma_box) {
        // the input box as ONE tiled TMA load: rows padded to 16 bytes, at most 256 cells a side; the input array's
        // rows must be 16-byte multiples
        int bw, bh;
        bidi::grid_rowbox_dims(H.ff.d, H.ff.cx.data(), bw, bh);
        const int bwp = (bw + 3 + 3) / 4 * 4;  // the box starts on a multiple of four cells: up to three cells of slack on the left
        if (bwp <= 256 && bh <= 256 && h->cfg.in_width % 4 == 0 && encode_input_tmap(h, bwp, bh)) {
          g.tm_bw = bwp; g.tm_bh = bh;
          g.act_box = (bwp * bh + 31) / 32 * 32;  // keeps the weight slab behind it 128-byte aligned
        }
      }
      g.act_cap = std::max(g.act_dense ? 0 : fbox, rbox) + 2;
      need = std::max({sizeof(float) * (size_t)g.lat, sizeof(float) * 3 * (size_t)g.gather, sizeof(float) * 4 * (size_t)(g.fb + g.pred)});
      if (sizeof(float) * (size_t)(g.act_slab + g.act_box + 2 * g.act_cap) + 16 > 200 * 1024) need = 1 << 30;  // (no such layer in practice)
    }
    // default: untiled (large-batch) layers at least 16 units wide; k-winners layers at least 24 wide even for one
    // agent (C4: 0.44 ms against 0.49 ms with the slot-group tile kernels); the iterative sweeps of ONE agent stay on
    // the tile kernels (C4: 7.4 ms against 9.1 ms: a 507-connection chain per thread is too long for a single agent)
    const bool wide_kw = h->cfg.variant == BIDI_KWINNERS && h->layers[l].hw >= 24;
    g.on = need <= 160 * 1024 && (h->grid_mode == 2 || ((!h->tiled[l] || wide_kw) && h->layers[l].hw >= 16)) ? 1 : 0;
    if (g.on) most = std::max(most, need);
    h->gridk[l] = g;
  }
  if (most > 48 * 1024) {
    CU(cudaFuncSetAttribute(bidi::k_g_kw_inhibit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)most));
    CU(cudaFuncSetAttribute(bidi::k_g_reconstruct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)most));
    CU(cudaFuncSetAttribute(bidi::k_g_kw_predict, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)most));
    CU(cudaFuncSetAttribute(bidi::k_g_ic_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)most));
  }
  size_t act_most = 0;
  for (int l = 0; l < h->L; l++) if (h->gridk[l].on) act_most = std::max(act_most, act_smem_bytes(h->gridk[l]));
  if (act_most > 48 * 1024) {
    CU(cudaFuncSetAttribute(bidi::k_g_kw_activate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)act_most));
    CU(cudaFuncSetAttribute(bidi::k_g_kw_activate<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)act_most));
    CU(cudaFuncSetAttribute(bidi::k_g_kw_activate<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)act_most));
    CU(cudaFuncSetAttribute(bidi::k_g_kw_activate<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)act_most));
    CU(cudaFuncSetAttribute(bidi::k_g_kw_activate<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)act_most));
    CU(cudaFuncSetAttribute(bidi::k_g_kw_activate<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)act_most));
  }
  return BIDI_OK;
}

// One simStep as a sequence of kernel launches on h->stream (captured into a graph).
void record_step(Recorder& R, bool learn) {
  bidi_engine* h = R.h;
  const int L = h->L, V = h->cfg.variant;
  const long long A = h->A;
  auto& ly = h->layers;
  if (V == BIDI_KWINNERS) {
    // The reconstructions of a layer (sdr/RSDR.cpp:165-194) are read by nothing before the learning rule: where the
    // step is captured into a graph they form a second branch, beside the rest of the up pass and the down pass.
    const bool fork_recon = R.capturing && h->side != nullptr && h->overlap_learn && h->profile_name.empty();
    bool forked = false;
    // sdr/PredictiveRSDR.cpp:102-112
    for (int l = 0; l < L; l++) {
      const LayerDev& Y = ly[l];
      const long long nxt = l < L - 1 ? ly[l + 1].vis_input : -1LL;
      if (h->gridk[l].on) {
        const auto& G = h->gridk[l];
        const long long th = A * bidi::grid_tiles(Y.hw, Y.hh), tv = A * bidi::grid_tiles(Y.vw, Y.vh);
        R.before("kw_activate");
        launch_kw_activate(h, R.s(), G, Y, l);
        R.after("kw_activate");
        R.kernels++;
        launch_grid(R, "kw_inhibit", bidi::k_g_kw_inhibit, th, G.lat, Y, l, Y.act, Y.state, nxt);
        if (fork_recon) {
          cudaEventRecord(h->ev_fork, h->stream);
          cudaStreamWaitEvent(h->side, h->ev_fork, 0);
          R.st = h->side;
          forked = true;
        }
        launch_grid(R, "reconstruct", bidi::k_g_reconstruct, tv + th, 3 * G.gather, Y, l, Y.state, 1, 0, G.gather, 0, 0.0f, 0);
        R.st = nullptr;
      } else if (h->tiled[l]) {
        launch_tile(R, "kw_activate", bidi::k_t_kw_activate, tiles_of(A, Y.nh), tile_bytes(nslots_of(Y.ff) + nslots_of(Y.rec)), Y, l);
        launch_tile(R, "kw_inhibit", bidi::k_t_kw_inhibit, tiles_of(A, Y.nh), 0, Y, l, Y.act, Y.state, nxt);
        launch_tile(R, "reconstruct", bidi::k_t_reconstruct, tiles_of(A, Y.nv) + tiles_of(A, Y.nh),
                    tile_bytes(std::max(Y.ff.max_cand, Y.rec.max_cand)), Y, l, Y.state, 0, 0.0f, 0, 0);
      } else {
        R.launch("kw_activate", bidi::k_kw_activate, A * Y.nh, Y, l);
        R.launch("kw_inhibit", bidi::k_kw_inhibit, A * Y.nh, Y, l, Y.act, Y.state, nxt);
        R.launch("reconstruct", bidi::k_reconstruct, A * (Y.nv + Y.nh), Y, l, Y.state, 0, 0.0f, 0);
      }
    }
    // :115-171
    for (int l = L - 1; l >= 0; l--) {
      const LayerDev& Y = ly[l];
      const LayerDev& Up = ly[l < L - 1 ? l + 1 : l];
      if (h->gridk[l].on) {
        const auto& G = h->gridk[l];
        const long long th = A * bidi::grid_tiles(Y.hw, Y.hh);
        launch_grid(R, "predict", bidi::k_g_kw_predict, th, 4 * (G.fb + G.pred), Y, l, (int)learn, Up.p_state, Up.p_state_prev, G.pred, G.fb);
        launch_grid(R, "kw_inhibit", bidi::k_g_kw_inhibit, th, G.lat, Y, l, Y.p_act, Y.p_state, -1LL);
      } else if (h->tiled[l]) {
        launch_tile(R, "predict", bidi::k_t_predict, tiles_of(A, Y.pn), tile_bytes(nslots_of(Y.fb) + nslots_of(Y.pred)), Y, l, (int)learn,
                    Up.p_state, Up.p_state_prev);
        launch_tile(R, "kw_inhibit", bidi::k_t_kw_inhibit, tiles_of(A, Y.nh), 0, Y, l, Y.p_act, Y.p_state, -1LL);
      } else {
        R.launch("predict", bidi::k_predict, A * Y.pn, Y, l, (int)learn, Up.p_state, Up.p_state_prev);
        R.launch("kw_inhibit", bidi::k_kw_inhibit, A * Y.nh, Y, l, Y.p_act, Y.p_state, -1LL);
      }
    }
    if (forked) {  // the reconstructions are in before anything reads them (learning) or overwrites their inputs (stepEnd)
      cudaEventRecord(h->ev_join, h->side);
      cudaStreamWaitEvent(h->stream, h->ev_join, 0);
    }
    // :173-185
    if (learn)
      for (int l = 0; l < L; l++) {
        if (h->gridk[l].on) launch_grid(R, "kw_learn", bidi::k_g_kw_learn, A * bidi::grid_tiles(ly[l].hw, ly[l].hh), 0, ly[l], l);
        else if (h->tiled[l]) launch_tile(R, "kw_learn", bidi::k_t_kw_learn, tiles_of(A, ly[l].nh), 0, ly[l], l);
        else R.launch("kw_learn", bidi::k_kw_learn, A * ly[l].nh, ly[l], l);
      }
    // :188-193 -- the input-space prediction gathers from the predicted states and the (just updated) feed-forward
    // weights; stepEnd's copies touch neither: in a captured graph the two run side by side
    if (fork_recon && h->gridk[0].on) {
      cudaEventRecord(h->ev_fork, h->stream);
      cudaStreamWaitEvent(h->side, h->ev_fork, 0);
      R.st = h->side;
      R.launch("step_end", bidi::k_step_end, A * h->max_units, h->max_units);
      R.st = nullptr;
      cudaEventRecord(h->ev_join, h->side);
      launch_grid(R, "kw_predict_input", bidi::k_g_reconstruct, A * bidi::grid_tiles(ly[0].vw, ly[0].vh), 3 * h->gridk[0].gather, ly[0], 0, ly[0].p_state, 0, 1, h->gridk[0].gather, 0, 0.0f, 0);
      cudaStreamWaitEvent(h->stream, h->ev_join, 0);
      return;
    }
    R.launch("step_end", bidi::k_step_end, A * h->max_units, h->max_units);
    if (h->gridk[0].on) launch_grid(R, "kw_predict_input", bidi::k_g_reconstruct, A * bidi::grid_tiles(ly[0].vw, ly[0].vh), 3 * h->gridk[0].gather, ly[0], 0, ly[0].p_state, 0, 1, h->gridk[0].gather, 0, 0.0f, 0);
    else if (h->tiled[0])
      launch_tile(R, "kw_predict_input", bidi::k_t_reconstruct, tiles_of(A, ly[0].nv), tile_bytes(ly[0].ff.max_cand), ly[0], 0, ly[0].p_state, 0,
                  0.0f, 0, 1);
    else R.launch("kw_predict_input", bidi::k_kw_predict_input, A * h->n_in, ly[0]);
    return;
  }
  // Coder learning (sdr/IRSDR.cpp:287-319, neo/SparseCoder.cpp:326-361) reads only what the layer's own up pass and
  // the PREVIOUS step left behind (errors, states, the previous prediction), and nothing before stepEnd reads what it
  // writes, so in the step graph it is a second branch: beside the down pass, and -- where every layer's up pass is
  // one fused launch -- already beside the up pass of the layers above.
  // layers on the dense fused coder kernel learn in that kernel's tail (bidi_coder_dense.cuh)
  auto learns_in_coder = [&](int l) { return h->dense_rows[l] != 0; };
  auto record_learn_layer = [&](int l) {
    if (learns_in_coder(l)) return;
    if (h->csrl) { R.launch("neo_learn", bidi::k_neo_learn, A * ly[l].nh, ly[l], l); return; }  // sdr/IRSDR.cpp:321-356 with the units' _reward actions
    if (h->tiled[l]) launch_tile(R, "ic_learn", bidi::k_t_ic_learn, tiles_of(A, ly[l].nh), 0, ly[l], l, (int)(V != BIDI_ITERATIVE));
    else if (V == BIDI_ITERATIVE) R.launch("ic_learn", bidi::k_ic_learn, A * ly[l].nh, ly[l], l);
    else R.launch("neo_learn", bidi::k_neo_learn, A * ly[l].nh, ly[l], l);
  };
  bool any_separate_learn = false;
  for (int l = 0; l < L; l++) any_separate_learn = any_separate_learn || !learns_in_coder(l);
  const bool fork_learn = learn && any_separate_learn && h->overlap_learn && h->side != nullptr && R.capturing && !h->csrl;
  const int plearn = h->csrl ? 0 : (int)learn;  // deep::CSRL's prediction weights learn through traces after the SDRRL pass
  for (int l = 0; l < L; l++) {
    const long long nh = A * ly[l].nh, nvh = A * (ly[l].nv + ly[l].nh);
    if (h->coder_cta[l] && !h->force_phase_kernels && (h->coder_cta[l] != 2 || h->dense_rows[l])) { // whole up pass of the layer in one launch
      const long long nxt = l < L - 1 ? ly[l + 1].vis_input : -1LL;
      R.before("coder");
      if (h->coder_cta[l] == 2) {  // windows span the grid: dense rows (bidi_coder_dense.cuh)
        const int threads = bidi::coder_dense_threads(ly[l]);
        const size_t smem = sizeof(float) * (size_t)bidi::coder_dense_shape(ly[l]).floats;
        const bool light = threads == 32;  // one warp per agent: the register-lean build, 28 CTAs per SM
        const int fused_learn = learn && learns_in_coder(l) ? 1 : 0;
        if (V == BIDI_ITERATIVE) {
          if (light) bidi::k_coder_dense<0, 1><<<(unsigned)A, threads, smem, h->stream>>>(h->P, ly[l], l, nxt, fused_learn);
          else bidi::k_coder_dense<0, 0><<<(unsigned)A, threads, smem, h->stream>>>(h->P, ly[l], l, nxt, fused_learn);
        } else {
          if (light) bidi::k_coder_dense<1, 1><<<(unsigned)A, threads, smem, h->stream>>>(h->P, ly[l], l, nxt, fused_learn);
          else bidi::k_coder_dense<1, 0><<<(unsigned)A, threads, smem, h->stream>>>(h->P, ly[l], l, nxt, fused_learn);
        }
      } else {
        const int threads = std::min(512, std::max(64, (ly[l].nv + ly[l].nh + 31) / 32 * 32));
        const size_t smem = sizeof(float) * (size_t)bidi::coder_cta_shape(ly[l]).floats;
        if (V == BIDI_ITERATIVE) bidi::k_coder_cta<0><<<(unsigned)A, threads, smem, h->stream>>>(h->P, ly[l], l, nxt);
        else bidi::k_coder_cta<1><<<(unsigned)A, threads, smem, h->stream>>>(h->P, ly[l], l, nxt);
      }
      R.after("coder");
      R.kernels++;
      continue;
    }
    const LayerDev& Y = ly[l];
    if (V == BIDI_ITERATIVE && h->persist[l].on && !h->force_phase_kernels) {  // bidi_persist.cuh
      R.before("coder");
      launch_persist(h, R.s(), l);
      R.after("coder");
      R.kernels++;
      if (l < L - 1) R.launch("ic_finish", bidi::k_ic_finish, nh, Y, l, 0, 0.0f, ly[l + 1].vis_input);
      continue;
    }
    const bool tl = h->tiled[l] != 0;
    const size_t sm_sweep = tile_bytes(nslots_of(Y.ff) + nslots_of(Y.rec) + nslots_of(Y.lat));
    const size_t sm_rec = tile_bytes(std::max(Y.ff.max_cand, Y.rec.max_cand));
    const long long t_sweep = tiles_of(A, Y.nh), t_rec = tiles_of(A, Y.nv) + tiles_of(A, Y.nh);
    const bool gk = h->gridk[l].on;  // large layer: 2-D tiles with the source boxes in shared memory
    const auto& G = h->gridk[l];
    const long long g_th = A * bidi::grid_tiles(Y.hw, Y.hh), g_tv = A * bidi::grid_tiles(Y.vw, Y.vh);
    int sweep_it = 0;
    auto sweep = [&](int first, int acc, float scale) {
      if (gk) launch_grid(R, "ic_sweep", bidi::k_g_ic_sweep, g_th, G.ff + G.rec + G.lat, Y, l, first, acc, scale, sweep_it, G.ff, G.rec);
      else if (tl) launch_tile(R, "ic_sweep", bidi::k_t_ic_sweep, t_sweep, sm_sweep, Y, l, first, acc, scale, sweep_it);
      else R.launch("ic_sweep", bidi::k_ic_sweep, nh, Y, l, first, acc, scale, sweep_it);
      sweep_it++;
    };
    auto recon = [&](long long drive, int scaled, float mult, int copy_spike) {
      if (gk) launch_grid(R, "reconstruct", bidi::k_g_reconstruct, g_tv + g_th, 3 * G.gather, Y, l, drive, 1, 0, G.gather, scaled, mult, copy_spike);
      else if (tl) launch_tile(R, "reconstruct", bidi::k_t_reconstruct, t_rec, sm_rec, Y, l, drive, scaled, mult, copy_spike, 0);
      else R.launch("reconstruct", bidi::k_reconstruct, nvh, Y, l, drive, scaled, mult, copy_spike);
    };
    if (V == BIDI_ITERATIVE) { // sdr/IRSDR.cpp:125-216
      const int settle = Y.iter_settle, measure = Y.iter_measure;
      for (int it = 0; it < settle; it++) { sweep((int)(it == 0), 0, 0.0f); recon(Y.spike, 0, 0.0f, 1); }
      const float inv = 1.0f / measure;
      for (int it = 0; it < measure; it++) { sweep((int)(settle == 0 && it == 0), 1, inv); recon(Y.spike, 0, 0.0f, 1); }
      recon(Y.state, 0, 0.0f, 0);
      if (l < L - 1) R.launch("ic_finish", bidi::k_ic_finish, nh, Y, l, 0, 0.0f, ly[l + 1].vis_input);
    } else { // neo/SparseCoder.cpp:116-176
      const int iter = Y.iter_settle;
      float counter = 0.0f;
      for (int it = 0; it < iter; it++) {
        sweep((int)(it == 0), 2, 0.0f);
        counter += 1.0f;
        recon(Y.state, 1, 1.0f / counter, 1);
      }
      R.launch("ic_finish", bidi::k_ic_finish, nh, Y, l, 1, 1.0f / counter, l < L - 1 ? ly[l + 1].vis_input : -1LL);
    }
  }
  auto record_learn = [&]() {
    for (int l = 0; l < L; l++) record_learn_layer(l);
  };
  if (fork_learn) {
    cudaEventRecord(h->ev_fork, h->stream);
    cudaStreamWaitEvent(h->side, h->ev_fork, 0);
    R.st = h->side;
    record_learn();
    R.st = nullptr;
    cudaEventRecord(h->ev_join, h->side);
  }
  for (int l = L - 1; l >= 0; l--) {
    const LayerDev& Y = ly[l];
    const LayerDev& Up = ly[l < L - 1 ? l + 1 : l];
    if (V == BIDI_NEO_AGENT) record_columns(R, l);
    if (h->tiled[l] && V != BIDI_NEO_AGENT)
      launch_tile(R, "predict", bidi::k_t_predict, tiles_of(A, Y.pn), tile_bytes(nslots_of(Y.fb) + nslots_of(Y.pred)), Y, l, plearn, Up.p_state,
                  Up.p_state_prev);
    else R.launch("predict", bidi::k_predict, A * Y.pn, Y, l, plearn, Up.p_state, Up.p_state_prev);
  }
  if (V == BIDI_NEO_AGENT) record_columns(R, L);
  if (h->tiled[L] && V != BIDI_NEO_AGENT)
    launch_tile(R, "predict_input", bidi::k_t_predict, tiles_of(A, ly[L].pn), tile_bytes(nslots_of(ly[L].fb)), ly[L], L, plearn, ly[0].p_state,
                ly[0].p_state_prev);
  else R.launch("predict_input", bidi::k_predict_input, A * h->n_in, ly[L], plearn, ly[0].p_state, ly[0].p_state_prev);
  if (h->csrl) {  // deep/CSRL.cpp:246-419
    if (h->P.csrl_n_actions > 0) R.launch("csrl_explore", bidi::k_csrl_explore, A * h->P.csrl_n_actions);
    for (int k = 0; k <= L; k++) {  // the SDRRL units: top layer first (reward flows down), the input nodes last
      const int l = k < L ? L - 1 - k : L;
      const size_t smem = sizeof(float) * BIDI_SDRRL_WARPS * (size_t)bidi::sdrrl_warp_floats(ly[l].sd.max_in);
      R.before("sdrrl");
      bidi::k_sdrrl<<<grid_for(A * ly[l].sd.n, BIDI_SDRRL_WARPS), 32 * BIDI_SDRRL_WARPS, smem, R.s()>>>(h->P, ly[l], l);
      R.after("sdrrl");
      R.kernels++;
    }
    if (learn) {
      for (int l = 0; l < L; l++) R.launch("csrl_pred_learn", bidi::k_csrl_pred_learn, A * ly[l].pn, ly[l], l, ly[l < L - 1 ? l + 1 : l].p_state_prev);
      R.launch("csrl_pred_learn", bidi::k_csrl_pred_learn, A * ly[L].pn, ly[L], L, ly[0].p_state_prev);
    }
  }
  if (fork_learn) cudaStreamWaitEvent(h->stream, h->ev_join, 0);
  else if (learn) record_learn();
  R.launch("step_end", bidi::k_step_end, A * h->max_units, h->max_units);
  if (h->qnet) {  // sdr/QPRSDR.cpp:102-341, after _prsdr.simStep
      R.before("qnet");
    bidi::k_qnet<<<(unsigned)A, 256, 0, h->stream>>>(h->P, (int)learn);
    R.after("qnet");
    R.kernels++;
  }
}

// ---- layout of the layers the dense fused coder runs (bidi_coder_dense.cuh): rows while that kernel is the layer's
// path, slot planes while any other kernel family is (the per-phase kernels, the general fused kernel -- which the
// tests select through bidi_set_option).  Switching converts every agent's region in place.
bool dense_wanted(const bidi_engine* h, int l) {
  // (a deep::CSRL's coder learns from rewards its SDRRL units produce in the down pass: not in the coder kernel's tail)
  return h->coder_cta[l] == 2 && !h->force_phase_kernels && !h->csrl;
}

void convert_region(float* R, const LayerHost& H, int nh, bool to_rows) {
  std::vector<float> tmp(R, R + H.region_len);
  std::fill(R, R + H.region_len, 0.0f);
  const int nsf = H.ff.d.nslots, nsr = H.rec.d.r >= 0 ? H.rec.d.nslots : 0, pitch = H.row_pitch, nf = H.row_nf;
  for (int u = 0; u < nh; u++) {
    for (int s = 0; s < nsf; s++) {
      const long long row = (long long)u * pitch + s, pl = (long long)s * nh + u;
      if (to_rows) R[row] = tmp[pl]; else R[pl] = tmp[row];
    }
    for (int k = 0; k < nsr; k++) {
      const long long row = (long long)u * pitch + nf + k, pl = H.rec_plane + (long long)k * nh + u;
      if (to_rows) R[row] = tmp[pl]; else R[pl] = tmp[row];
    }
  }
}

// one family of a k-winners layer between slot planes [slot][U] and rows [unit][pitch] (bidi_grid.cuh), in place
void convert_family(float* R, int nslots, long long U, int pitch, bool to_rows) {
  const long long len = (long long)pitch * U;
  std::vector<float> tmp(R, R + len);
  std::fill(R, R + len, 0.0f);
  for (long long u = 0; u < U; u++)
    for (int s = 0; s < nslots; s++) {
      const long long row = u * pitch + s, pl = (long long)s * U + u;
      if (to_rows) R[row] = tmp[pl]; else R[pl] = tmp[row];
    }
}

int apply_layout(bidi_engine* h) {
  bool changed = false;
  // k-winners layers: rows exactly while the 2-D grid kernels are the layer's path
  for (int l = 0; l < h->L && h->cfg.variant == BIDI_KWINNERS; l++) {
    LayerHost& H = h->hl[l];
    const int want = h->gridk[l].on ? 1 : 0;
    if (want == h->dense_rows[l]) continue;
    WinHost* fams[4] = {&H.ff, &H.rec, &H.fb, &H.pred};
    if (h->state_touched)
      for (int a = 0; a < h->A; a++) {
        int rc = mirror_load(h, a);
        if (rc) return rc;
        for (WinHost* w : fams)
          if (w->d.r >= 0 && w->d.w_off >= 0) convert_family(h->mirror.data() + w->d.w_off, w->d.nslots, w->d.U, kw_pitch(w->d.nslots), want != 0);
        h->mirror_dirty = true;
      }
    for (WinHost* w : fams) w->d.row_pitch = (want && w->d.r >= 0) ? kw_pitch(w->d.nslots) : 0;
    h->layers[l].ff = H.ff.d; h->layers[l].rec = H.rec.d; h->layers[l].fb = H.fb.d; h->layers[l].pred = H.pred.d;
    h->dense_rows[l] = want;
    changed = true;
  }
  for (int l = 0; l < h->L; l++) {
    LayerHost& H = h->hl[l];
    if (!H.dense_geom) continue;
    const int want = dense_wanted(h, l) ? 1 : 0;
    if (want == h->dense_rows[l]) continue;
    if (h->state_touched)
      for (int a = 0; a < h->A; a++) {
        int rc = mirror_load(h, a);
        if (rc) return rc;
        convert_region(h->mirror.data() + H.w_region, H, h->layers[l].nh, want != 0);
        if (H.tr_region >= 0) convert_region(h->mirror.data() + H.tr_region, H, h->layers[l].nh, want != 0);
        h->mirror_dirty = true;
      }
    const bool rec = H.rec.d.r >= 0;
    H.ff.d.row_pitch = want ? H.row_pitch : 0;
    if (rec) { H.rec.d.row_pitch = H.ff.d.row_pitch; H.rec.d.w_off = H.w_region + (want ? H.row_nf : H.rec_plane); }
    if (H.tr_region >= 0 && rec) H.rec.d.tr_off = H.tr_region + (want ? H.row_nf : H.rec_plane);
    h->layers[l].ff = H.ff.d; h->layers[l].rec = H.rec.d;
    h->dense_rows[l] = want;
    changed = true;
  }
  if (!changed) return BIDI_OK;
  int rc = mirror_release(h);
  if (rc) return rc;
  CU(cudaMemcpyAsync(h->d_layers, h->layers.data(), sizeof(LayerDev) * (h->L + 1), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (auto& g : h->graphs) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
  return BIDI_OK;
}

// mode 0 / 1: simStep with learn off / on; 2: simStepGenerate (no learning, noise-driven coders)
int ensure_graph(bidi_engine* h, int mode) {
  cudaGraphExec_t& exec = h->graphs[mode];
  if (exec) return BIDI_OK;
  cudaGraph_t graph = nullptr;
  const bool learn = mode == 1;
  CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  Recorder R{h};
  R.capturing = true;
  h->P.gen_on = mode == 2 ? 1 : 0;  // kernel arguments are captured by value
  record_step(R, learn);
  h->P.gen_on = 0;
  cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
  if (e != cudaSuccess) return fail(h, BIDI_ECUDA, std::string("graph capture: ") + cudaGetErrorString(e));
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, BIDI_ECUDA, std::string("kernel launch during capture: ") + cudaGetErrorString(e));
  CU(cudaGraphInstantiate(&exec, graph, 0));
  CU(cudaGraphDestroy(graph));
  h->graph_kernels[mode] = R.kernels;
  return BIDI_OK;
}

int strip_step_nccl(bidi_engine* h, bool learn);

struct Recorder;
void record_step(Recorder& R, bool learn);
int run_step(bidi_engine* h, bool learn, bool device_noise) {
  int rc = mirror_release(h);
  if (rc) return rc;
  h->state_touched = true;
  if (h->strip_n > 1) return strip_step_nccl(h, learn);
  if (h->qnet && device_noise) {
    bidi::k_noise_q<<<grid_for((long long)h->A * h->q_half, kBlock), kBlock, 0, h->stream>>>(h->P, h->d_noise_u, h->d_noise_v, h->noise_seed, h->step_index, h->first_agent);
    h->launches++;
  }
  if (h->csrl && device_noise) {
    bidi::k_csrl_noise<<<grid_for((long long)h->A * h->n_noise, kBlock), kBlock, 0, h->stream>>>(h->P, h->d_noise_u, h->d_noise_v, h->noise_seed, h->step_index, h->first_agent);
    h->launches++;
  }
  if (h->cfg.variant == BIDI_NEO_AGENT && device_noise) {
    // the step counter changes every call, so this launch stays outside the graph
    bidi::k_noise<<<grid_for((long long)h->A * h->ncol, kBlock), kBlock, 0, h->stream>>>(
        h->P, h->d_noise_u, h->d_noise_v, h->noise_seed, h->step_index, h->first_agent);
    h->launches++;
  }
  rc = ensure_graph(h, learn ? 1 : 0);
  if (rc) return rc;
  if (h->cfg.variant == BIDI_NEO_AGENT) { h->col_inputs_stale = true; std::fill(h->col_inputs_set.begin(), h->col_inputs_set.end(), 0); }
  CU(cudaGraphLaunch(h->graphs[learn], h->stream));
  h->launches += h->graph_kernels[learn];
  h->step_index++;
  return BIDI_OK;
}

void fill_columns_params(ColumnsDev& C, int cells, int iter, float sparsity, float leak, float gamma, float gl,
                         float aff, float alat, float athr, float aq, float aact, float std, float brk) {
  C.cells = cells; C.iter = iter; C.sparsity = sparsity; C.leak = leak; C.gamma = gamma; C.gamma_lambda = gl;
  C.a_ff = aff; C.a_lat = alat; C.a_thr = athr; C.a_q = aq; C.a_act = aact; C.expl_std = std; C.expl_break = brk;
}

void strip_release(bidi_engine* h);

} // namespace

// =================================================================== C ABI
extern "C" {

const char* bidi_last_error(const bidi_engine* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

void bidi_destroy(bidi_engine* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto& g : h->graphs) if (g) cudaGraphExecDestroy(g);
  cudaFree(h->d_blob); cudaFree(h->d_in0); cudaFree(h->d_pred); cudaFree(h->d_rewards); cudaFree(h->d_noise_u);
  cudaFree(h->d_noise_v); cudaFree(h->d_col_actions); cudaFree(h->d_frames); cudaFree(h->d_frame_rewards);
  cudaFree(h->d_frame_noise_u); cudaFree(h->d_frame_noise_v); cudaFree(h->d_env); cudaFree(h->d_env_noise);
  cudaFree(h->d_persist_ctr); cudaFree(h->d_tables); cudaFree(h->d_layers); cudaFree(h->d_fb_src); cudaFree(h->d_fb_dst);
  cudaFree(h->d_qmask); cudaFree(h->d_qtables); cudaFree(h->d_csrl_tables); cudaFree(h->d_gen); cudaFree(h->d_gen_scale); cudaFreeHost(h->h_gen);
  for (auto& e : h->profile_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  cudaFreeHost(h->h_in); cudaFreeHost(h->h_pred); cudaFreeHost(h->h_rewards); cudaFreeHost(h->h_noise); cudaFreeHost(h->h_actions);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_in) cudaEventDestroy(h->ev_in);
  strip_release(h);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int bidi_create(const bidi_config* cfg, const bidi_layer_desc* layers, int n_agents, int device, bidi_engine** out) {
  bidi_engine* h = nullptr;
  if (!cfg || !layers || !out) return fail(h, BIDI_EINVAL, "bidi_create: null argument");
  *out = nullptr;
  if (cfg->variant < BIDI_KWINNERS || cfg->variant > BIDI_CSRL) return fail(h, BIDI_EINVAL, "bidi_create: unknown variant");
  const bool qnet = cfg->variant == BIDI_QPRSDR, csrl = cfg->variant == BIDI_CSRL;
  if (csrl) {
    if (cfg->q_n_actions < 0 || cfg->q_n_actions > BIDI_MAX_ACTIONS) return fail(h, BIDI_EINVAL, "bidi_create: a CSRL has 0..BIDI_MAX_ACTIONS action inputs");
    for (int k = 0; k < cfg->q_n_actions; k++)
      if (cfg->q_action_idx[k] < 0 || cfg->q_action_idx[k] >= cfg->in_width * cfg->in_height || (k > 0 && cfg->q_action_idx[k] <= cfg->q_action_idx[k - 1]))
        return fail(h, BIDI_EINVAL, "bidi_create: action inputs must be inside the input and ascending");
    if (cfg->col_cells < 1 || cfg->col_cells > 32 || cfg->csrl_n_recurrent < 0 || 2 * (3 + cfg->csrl_n_recurrent) > BIDI_SDRRL_MAX_ACTIONS ||
        cfg->csrl_iter_measure < 1 || cfg->csrl_iter_settle < 0 || cfg->csrl_derive_iterations < 0)
      return fail(h, BIDI_EINVAL, "bidi_create: bad SDRRL parameters of the input nodes");
  }
  if (qnet) {
    if (cfg->q_n_actions < 1 || cfg->q_n_actions > BIDI_MAX_ACTIONS || cfg->q_derive_iterations < 0)
      return fail(h, BIDI_EINVAL, "bidi_create: a QPRSDR needs 1..BIDI_MAX_ACTIONS actions");
    for (int k = 0; k < cfg->q_n_actions; k++)
      if (cfg->q_action_idx[k] < 0 || cfg->q_action_idx[k] >= cfg->in_width * cfg->in_height || cfg->q_anti_idx[k] < 0 ||
          cfg->q_anti_idx[k] >= cfg->in_width * cfg->in_height)
        return fail(h, BIDI_EINVAL, "bidi_create: action index outside the input");
  }
  if (cfg->n_layers < 1 || cfg->n_layers > BIDI_MAX_LAYERS) return fail(h, BIDI_EINVAL, "bidi_create: n_layers out of range");
  if (n_agents < 1 || cfg->in_width < 1 || cfg->in_height < 1) return fail(h, BIDI_EINVAL, "bidi_create: bad sizes");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return fail(h, BIDI_ENODEV, "bidi_create: no usable CUDA device (this engine has no CPU path)");
  }
  const int V = (qnet || csrl) ? BIDI_ITERATIVE : cfg->variant, L = cfg->n_layers;  // a QPRSDR owns an IPredictiveRSDR; a CSRL runs IRSDR coders
  for (int l = 0; l < L; l++) {
    const bidi_layer_desc& d = layers[l];
    if (d.width < 1 || d.height < 1 || d.r_ff < 0 || d.r_lat < 0 || d.r_pred < 0 || d.r_rec < -1 || (l < L - 1 && d.r_fb < 0))
      return fail(h, BIDI_EINVAL, "bidi_create: bad layer descriptor");
    if (V == BIDI_ITERATIVE && d.iter_settle + d.iter_measure < 1) return fail(h, BIDI_EINVAL, "bidi_create: no coder iterations");
    if ((V == BIDI_NEO_HIERARCHY || V == BIDI_NEO_AGENT) && d.iter_settle < 1) return fail(h, BIDI_EINVAL, "bidi_create: no coder iterations");
    if (V == BIDI_NEO_AGENT && (d.col_cells < 1 || d.col_cells > 32 || d.col_iter < 1))
      return fail(h, BIDI_EINVAL, "bidi_create: columns need 1..32 cells and >= 1 iteration");
    if (csrl && (d.col_cells < 1 || d.col_cells > 32 || d.csrl_n_recurrent < 0 || 2 * (3 + d.csrl_n_recurrent) > BIDI_SDRRL_MAX_ACTIONS ||
                 d.iter_measure < 1 || d.csrl_derive_iterations < 0 || d.r_fb < 0))
      return fail(h, BIDI_EINVAL, "bidi_create: bad SDRRL parameters in a layer descriptor");
  }
  if (V != BIDI_KWINNERS && cfg->r_infb < 0) return fail(h, BIDI_EINVAL, "bidi_create: bad input feedback radius");
  if (V == BIDI_NEO_AGENT && (cfg->col_cells < 1 || cfg->col_cells > 32 || cfg->col_iter < 1))
    return fail(h, BIDI_EINVAL, "bidi_create: columns need 1..32 cells and >= 1 iteration");

  const bool h_no_col_stage = getenv("BIDI_NO_COL_STAGE") != nullptr;  // experiment switch: gather the column inputs straight from the blob
  const bool h_no_col_regs = getenv("BIDI_NO_COL_REGS") != nullptr;  // experiment switch: keep whole weight rows in shared memory
  h = new bidi_engine();
  h->cfg = *cfg; h->ld.assign(layers, layers + L);
  h->cfg.variant = V; h->qnet = qnet; h->csrl = csrl;
  const bool traced = is_traced(V) || csrl;
  h->device = device; h->A = n_agents; h->L = L; h->n_in = cfg->in_width * cfg->in_height;
  h->col_inputs_set.assign((size_t)n_agents, 0);
  h->layers.assign(L + 1, LayerDev());
  h->hl.resize(L + 1);
  for (auto& d : h->layers) memset(&d, 0, sizeof(LayerDev));

  // ---- geometry and blob layout
  Allocator al;
  int vw = cfg->in_width, vh = cfg->in_height;
  h->max_units = h->n_in;
  for (int l = 0; l <= L; l++) {
    LayerDev& D = h->layers[l];
    LayerHost& H = h->hl[l];
    if (l < L) {
      const bidi_layer_desc& d = layers[l];
      D.vw = vw; D.vh = vh; D.hw = d.width; D.hh = d.height; D.nv = vw * vh; D.nh = d.width * d.height;
      D.row0 = 0; D.row1 = D.hh; D.vrow0 = 0; D.vrow1 = D.vh;
      h->max_units = std::max(h->max_units, D.nh);
      D.vis_input = l == 0 ? -1 : al.take(D.nv);
      D.vis_recon = al.take(D.nv);
      D.thr = al.take(D.nh); D.state = al.take(D.nh); D.state_prev = al.take(D.nh); D.act = al.take(D.nh);
      D.recon = al.take(D.nh); D.spike = al.take(D.nh); D.spike_prev = al.take(D.nh);
      build_win(H.ff, D.hw, D.hh, vw, vh, d.r_ff, false, false);
      build_win(H.rec, D.hw, D.hh, D.hw, D.hh, d.r_rec, true, true);
      build_win(H.lat, D.hw, D.hh, D.hw, D.hh, d.r_lat, true, true);
      H.dense_geom = is_iter(V) && H.ff.d.absx && H.ff.d.absy && (d.r_rec < 0 || (H.rec.d.absx && H.rec.d.absy));
      if (H.dense_geom) {  // one region for either layout (slot planes now; rows once the dense coder is chosen, apply_layout)
        auto r32 = [](long long n) { return (n + 31) / 32 * 32; };
        H.row_nf = (D.nv + 3) & ~3;
        H.row_pitch = bidi::dense_pitch(H.row_nf + (d.r_rec >= 0 ? ((D.nh + 3) & ~3) : 0));
        H.rec_plane = r32((long long)H.ff.d.nslots * D.nh);
        const long long region = std::max(H.rec_plane + (d.r_rec >= 0 ? r32((long long)H.rec.d.nslots * D.nh) : 0), r32((long long)D.nh * H.row_pitch));
        H.region_len = region;
        H.w_region = al.take(region);
        H.ff.d.w_off = H.w_region;
        if (d.r_rec >= 0) H.rec.d.w_off = H.w_region + H.rec_plane;
        if (traced) {
          H.tr_region = al.take(region);
          H.ff.d.tr_off = H.tr_region;
          if (d.r_rec >= 0) H.rec.d.tr_off = H.tr_region + H.rec_plane;
        }
      } else if (V == BIDI_KWINNERS) {  // room for either layout: slot planes, or one row of kw_pitch(nslots) floats per unit (apply_layout)
        H.ff.d.w_off = al.take((long long)kw_pitch(H.ff.d.nslots) * D.nh);
        if (d.r_rec >= 0) H.rec.d.w_off = al.take((long long)kw_pitch(H.rec.d.nslots) * D.nh);
      } else {
        H.ff.d.w_off = al.take((long long)H.ff.d.nslots * D.nh);
        if (d.r_rec >= 0) H.rec.d.w_off = al.take((long long)H.rec.d.nslots * D.nh);
      }
      if (is_iter(V)) H.lat.d.w_off = al.take((long long)H.lat.d.nslots * D.nh);
      if (traced && !H.dense_geom) {
        H.ff.d.tr_off = al.take((long long)H.ff.d.nslots * D.nh);
        if (d.r_rec >= 0) H.rec.d.tr_off = al.take((long long)H.rec.d.nslots * D.nh);
      }
      D.pn = D.nh;
      D.p_state = al.take(D.pn);
      if (l < L - 1) build_win(H.fb, D.hw, D.hh, layers[l + 1].width, layers[l + 1].height, d.r_fb, false, false);
      else if (csrl) build_win(H.fb, D.hw, D.hh, D.hw, D.hh, d.r_fb, true, false);  // the top layer looks at the rewards around its own position, deep/CSRL.cpp:92-114
      else build_win(H.fb, D.hw, D.hh, 1, 1, -1, true, false);
      build_win(H.pred, D.hw, D.hh, D.hw, D.hh, d.r_pred, true, false);
      if (l < L - 1 || csrl) H.fb.d.w_off = al.take((long long)(V == BIDI_KWINNERS ? kw_pitch(H.fb.d.nslots) : H.fb.d.nslots) * D.pn);
      H.pred.d.w_off = al.take((long long)(V == BIDI_KWINNERS ? kw_pitch(H.pred.d.nslots) : H.pred.d.nslots) * D.pn);
      if (csrl) {
        H.fb.d.tr_off = al.take((long long)H.fb.d.nslots * D.pn);
        H.pred.d.tr_off = al.take((long long)H.pred.d.nslots * D.pn);
      }
      if (qnet) {
        D.q_state = al.take(D.nh); D.q_err = al.take(D.nh); D.q_bias_w = al.take(D.nh); D.q_bias_tr = al.take(D.nh);
        D.q_w_off = al.take((long long)H.ff.d.nslots * D.nh); D.q_tr_off = al.take((long long)H.ff.d.nslots * D.nh);
      }
      D.learn_ff = d.learn_ff; D.learn_rec = d.learn_rec; D.learn_lat = d.learn_lat; D.learn_threshold = d.learn_threshold;
      D.learn_fb = d.learn_fb; D.learn_pred = d.learn_pred; D.sparsity = d.sparsity;
      D.avg_surprise_decay = d.avg_surprise_decay; D.attention_factor = d.attention_factor;
      D.iter_settle = d.iter_settle; D.iter_measure = d.iter_measure; D.leak = d.leak; D.lambda = d.lambda;
      D.weight_decay = d.weight_decay; D.max_weight_delta = d.max_weight_delta;
      D.baseline_decay = d.baseline_decay; D.sensitivity = d.sensitivity;
      vw = d.width; vh = d.height;
    } else {
      if (V == BIDI_KWINNERS) continue;
      D.pn = h->n_in;
      D.p_state = -1;
      D.hw = cfg->in_width; D.hh = cfg->in_height; D.row0 = 0; D.row1 = D.hh;
      build_win(H.fb, cfg->in_width, cfg->in_height, layers[0].width, layers[0].height, cfg->r_infb, false, false);
      build_win(H.pred, cfg->in_width, cfg->in_height, 1, 1, -1, true, false);
      build_win(H.ff, 1, 1, 1, 1, -1, true, false); build_win(H.rec, 1, 1, 1, 1, -1, true, false); build_win(H.lat, 1, 1, 1, 1, -1, true, false);
      H.fb.d.w_off = al.take((long long)H.fb.d.nslots * D.pn);
      if (csrl) H.fb.d.tr_off = al.take((long long)H.fb.d.nslots * D.pn);
      D.learn_fb = cfg->learn_infb;
    }
    D.p_state_prev = al.take(D.pn); D.p_act = al.take(D.pn); D.p_act_prev = al.take(D.pn);
    D.p_baseline = al.take(D.pn); D.p_mod = al.take(D.pn);
    if (V == BIDI_NEO_AGENT) {
      ColumnsDev& C = D.col;
      C.n = D.pn;
      if (l < L) {
        const bidi_layer_desc& d = layers[l];
        fill_columns_params(C, d.col_cells, d.col_iter, d.col_sparsity, d.col_leak, d.col_gamma, d.col_gamma_lambda, d.col_alpha_ff,
                            d.col_alpha_lat, d.col_alpha_threshold, d.col_alpha_q, d.col_alpha_action, d.col_expl_stddev, d.col_expl_break);
      } else {
        fill_columns_params(C, cfg->col_cells, cfg->col_iter, cfg->col_sparsity, cfg->col_leak, cfg->col_gamma, cfg->col_gamma_lambda,
                            cfg->col_alpha_ff, cfg->col_alpha_lat, cfg->col_alpha_threshold, cfg->col_alpha_q, cfg->col_alpha_action,
                            cfg->col_expl_stddev, cfg->col_expl_break);
      }
      // nIn = |pred| + 2|fb|  (neo/Agent.cpp:90,137)
      H.col.in_off.assign(C.n + 1, 0);
      const int n_pairs = (C.n + 1) / 2;
      H.col.rec_off.assign(n_pairs + 1, 0);
      H.col.stride.assign(n_pairs, 0);
      C.cq = (C.cells + 3) / 4 * 4;
      C.planar = C.cells == 16 ? 1 : 0;  // the reference's default size runs the specialised kernel (bidi_columns16.cuh)
      C.max_in = C.max_s = C.max_rec = 0;
      for (int i = 0; i < C.n; i++) {
        const int ux = i % H.fb.d.uw, uy = i / H.fb.d.uw;
        int nfb = 0, npr = 0;
        for_each_conn_host(H.fb, ux, uy, [&](int, int) { nfb++; });
        for_each_conn_host(H.pred, ux, uy, [&](int, int) { npr++; });
        const int nin = npr + 2 * nfb;
        H.col.in_off[i + 1] = H.col.in_off[i] + nin;
        C.max_in = std::max(C.max_in, nin);
      }
      int s_min = 1 << 30;
      for (int p = 0; p < n_pairs; p++) {
        const int na = H.col.in_off[2 * p + 1] - H.col.in_off[2 * p];
        const int nb = 2 * p + 1 < C.n ? H.col.in_off[2 * p + 2] - H.col.in_off[2 * p + 1] : 0;
        H.col.stride[p] = column_pair_stride(std::max(na, nb), C.planar != 0);
        s_min = std::min(s_min, H.col.stride[p]);
      }
      // planar: how much of every weight row lives in registers (bidi_columns16.cuh); every row keeps >= 4 inputs in shared memory
      C.rh = !C.planar ? 0 : (s_min >= 68 ? 64 : (s_min >= 36 ? 32 : 0));
      if (h_no_col_regs) C.rh = 0;
      for (int p = 0; p < n_pairs; p++) {
        const int S = H.col.stride[p];
        const int part = 2 * column_block_floats(C, S);
        H.col.rec_off[p + 1] = H.col.rec_off[p] + 32 * C.rh + part;
        C.max_s = std::max(C.max_s, S); C.max_rec = std::max(C.max_rec, part);
      }
      if (C.planar) C.max_s += (16 - C.max_s % 32 + 32) % 32;  // pitch of the two gathered-input vectors: disjoint bank halves
      C.tot_in = H.col.in_off[C.n];
      const long long nc = (long long)C.n * C.cells;
      C.rec_base = al.take(H.col.rec_off[n_pairs]);
      C.inputs = al.take(C.tot_in);
      C.act = al.take(nc); C.spike = al.take(nc); C.state = al.take(nc);
      C.a_state = al.take(C.n * 3LL); C.a_expl = al.take(C.n * 3LL);
    }
  }
  long long csrl_offsets = 0;
  if (csrl) {  // one deep::SDRRL unit per node: plain planes in the reference's order (bidi_types.h, SdrrlDev)
    csrl_offsets = al.take(h->layers[L - 1].nh);
    for (int l = 0; l <= L; l++) {
      LayerDev& D = h->layers[l];
      LayerHost& H = h->hl[l];
      SdrrlDev& S = D.sd;
      S.n = D.pn;
      if (l < L) {
        const bidi_layer_desc& d = layers[l];
        S.cells = d.col_cells; S.n_rec = d.csrl_n_recurrent;
        S.sparsity = d.col_sparsity; S.gamma = d.col_gamma; S.leak = d.leak; S.a_ff = d.col_alpha_ff; S.a_lat = d.col_alpha_lat;
        S.a_thr = d.col_alpha_threshold; S.a_q = d.col_alpha_q; S.a_act = d.col_alpha_action; S.derive_alpha = d.csrl_derive_alpha;
        S.gamma_lambda = d.col_gamma_lambda; S.expl_std = d.col_expl_stddev; S.expl_break = d.col_expl_break;
        S.surprise_decay = d.avg_surprise_decay; S.settle = d.iter_settle; S.measure = d.iter_measure; S.derive_iter = d.csrl_derive_iterations;
      } else {
        S.cells = cfg->col_cells; S.n_rec = cfg->csrl_n_recurrent;
        S.sparsity = cfg->col_sparsity; S.gamma = cfg->col_gamma; S.leak = cfg->csrl_leak; S.a_ff = cfg->col_alpha_ff; S.a_lat = cfg->col_alpha_lat;
        S.a_thr = cfg->col_alpha_threshold; S.a_q = cfg->col_alpha_q; S.a_act = cfg->col_alpha_action; S.derive_alpha = cfg->csrl_derive_alpha;
        S.gamma_lambda = cfg->col_gamma_lambda; S.expl_std = cfg->col_expl_stddev; S.expl_break = cfg->col_expl_break;
        S.surprise_decay = cfg->csrl_avg_surprise_decay; S.settle = cfg->csrl_iter_settle; S.measure = cfg->csrl_iter_measure;
        S.derive_iter = cfg->csrl_derive_iterations;
      }
      S.na = 2 * (3 + S.n_rec);
      H.sd.in_off.assign(S.n + 1, 0);
      S.max_in = 0;
      for (int i = 0; i < S.n; i++) {  // numStates = |feedback| + |predictive| + recurrent inputs, deep/CSRL.cpp:136,184
        const int ux = i % H.fb.d.uw, uy = i / H.fb.d.uw;
        int nin = S.n_rec;
        for_each_conn_host(H.fb, ux, uy, [&](int, int) { nin++; });
        if (l < L) for_each_conn_host(H.pred, ux, uy, [&](int, int) { nin++; });
        H.sd.in_off[i + 1] = H.sd.in_off[i] + nin;
        S.max_in = std::max(S.max_in, nin);
      }
      S.tot_in = H.sd.in_off[S.n];
      const long long nc = (long long)S.n * S.cells, nan = (long long)S.n * S.na;
      S.inputs = al.take(S.tot_in); S.recon_err = al.take(S.tot_in); S.ff_w = al.take((long long)S.tot_in * S.cells);
      S.lat_w = al.take(nc * S.cells);
      S.thr = al.take(nc); S.act = al.take(nc); S.spike = al.take(nc); S.spike_prev = al.take(nc); S.state = al.take(nc);
      S.cell_as = al.take(nc); S.cell_ae = al.take(nc); S.q_w = al.take(nc); S.q_tr = al.take(nc);
      S.a_state = al.take(nan); S.a_expl = al.take(nan); S.a_err = al.take(nan);
      S.a_w = al.take(nc * S.na); S.a_tr = al.take(nc * S.na);
      S.prev_value = al.take(S.n); S.avg_surprise = al.take(S.n);
    }
    if (al.top >= (1LL << 31)) { delete h; return fail(nullptr, BIDI_EINVAL, "bidi_create: agent state too large for 32-bit gather offsets"); }
    for (int l = 0; l <= L; l++) {  // where every SDRRL input comes from, deep/CSRL.cpp:267-276, 295-304, 320-324
      const LayerDev& D = h->layers[l];
      LayerHost& H = h->hl[l];
      const SdrrlDev& S = D.sd;
      const bool input = l == L, top = l == L - 1;
      const LayerDev& Up = h->layers[input ? 0 : (top ? l : l + 1)];
      for (int i = 0; i < S.n; i++) {
        const int ux = i % H.fb.d.uw, uy = i / H.fb.d.uw;
        for_each_conn_host(H.fb, ux, uy, [&](int, int t) { H.sd.in_src.push_back(top ? -(1 + t) : (int)(Up.p_mod + t)); });
        if (!input) for_each_conn_host(H.pred, ux, uy, [&](int, int t) { H.sd.in_src.push_back((int)(D.state + t)); });
        for (int k = 0; k < S.n_rec; k++) H.sd.in_src.push_back((int)(S.a_expl + (long long)i * S.na + 3 + k));
      }
    }
    int base = cfg->q_n_actions;  // noise entries: the action inputs, then the units in visit order
    for (int l = L - 1; l >= 0; l--) { h->layers[l].sd.noise_base = base; base += h->layers[l].sd.n * h->layers[l].sd.na; }
    h->layers[L].sd.noise_base = base; base += h->layers[L].sd.n * h->layers[L].sd.na;
    h->n_noise = base;
    h->csrl_is_action.assign(h->n_in, 0);
    for (int k = 0; k < cfg->q_n_actions; k++) h->csrl_is_action[cfg->q_action_idx[k]] = 1;
  }
  long long q_qc_w = 0, q_qc_tr = 0, q_act = 0, q_prev = 0;
  if (qnet) {  // sdr/QPRSDR.cpp:13-31,86-96
    h->q_half = cfg->q_n_actions; h->q_top = layers[L - 1].width * layers[L - 1].height; h->n_noise = h->q_half;
    q_qc_w = al.take(h->q_top); q_qc_tr = al.take(h->q_top); q_act = al.take(8LL * h->q_half); q_prev = al.take(1);
    h->q_act_index.assign(h->n_in, -1);
    h->q_a_input.assign(2 * h->q_half, 0);
    for (int k = 0; k < h->q_half; k++) {
      h->q_act_index[cfg->q_action_idx[k]] = k;
      h->q_a_input[k] = cfg->q_action_idx[k]; h->q_a_input[h->q_half + k] = cfg->q_anti_idx[k];
    }
    // kept connections: layer 0 sees the action inputs, a higher layer the nodes below that kept any (sdr/QPRSDR.cpp:72-81)
    h->q_keep.resize(L);
    std::vector<char> below_any;
    for (int l = 0; l < L; l++) {
      const WinHost& w = h->hl[l].ff;
      const long long U = w.d.U;
      h->q_keep[l].assign((size_t)w.d.nslots * U, 0);
      std::vector<char> any((size_t)U, 0);
      for (int uy = 0; uy < w.d.uh; uy++)
        for (int ux = 0; ux < w.d.uw; ux++) {
          const long long u = ux + (long long)uy * w.d.uw;
          for_each_conn_host(w, ux, uy, [&](int s, int t) {
            const bool keep = l == 0 ? h->q_act_index[t] != -1 : below_any[t] != 0;
            if (keep) { h->q_keep[l][(size_t)s * U + u] = 1; any[u] = 1; }
          });
        }
      below_any.swap(any);
    }
  }
  h->stride = al.top;
  if (V == BIDI_NEO_AGENT) {
    if (h->stride >= (1LL << 31)) { delete h; return fail(nullptr, BIDI_EINVAL, "bidi_create: agent state too large for 32-bit gather offsets"); }
    // where every column input comes from: per feedback connection the upper column's
    // _signal action and the upper node's state, then the coder state per predictive
    // connection (neo/Agent.cpp:159-169; input nodes: :217-220)
    for (int l = 0; l <= L; l++) {
      LayerHost& H = h->hl[l];
      const LayerDev& D = h->layers[l];
      const bool input_layer = l == L;
      const LayerDev& Up = h->layers[input_layer ? 0 : (l == L - 1 ? l : l + 1)];
      H.col.in_src.clear();
      // the same inputs as positions in a staging of their source planes (ColumnsDev::stage_*): segment 0 = the upper
      // columns' _signal actions, 1 = the upper prediction states, 2 = this layer's coder states
      ColumnsDev& C = h->layers[l].col;
      const bool has_fb = input_layer || l < L - 1;
      const int n_up = has_fb ? H.fb.d.tw * H.fb.d.th : 0, n_own = input_layer ? 0 : H.pred.d.tw * H.pred.d.th;
      C.stage_len[0] = n_up; C.stage_step[0] = 3; C.stage_off[0] = Up.col.a_expl + 2;
      C.stage_len[1] = n_up; C.stage_step[1] = 1; C.stage_off[1] = Up.p_state;
      C.stage_len[2] = n_own; C.stage_step[2] = 1; C.stage_off[2] = D.state;
      std::vector<unsigned short> loc;
      for (int i = 0; i < D.col.n; i++) {
        const int ux = i % H.fb.d.uw, uy = i / H.fb.d.uw;
        if (has_fb)
          for_each_conn_host(H.fb, ux, uy, [&](int, int t) {
            H.col.in_src.push_back((int)(Up.col.a_expl + t * 3 + 2));
            H.col.in_src.push_back((int)(Up.p_state + t));
            loc.push_back((unsigned short)t); loc.push_back((unsigned short)(n_up + t));
          });
        if (!input_layer) for_each_conn_host(H.pred, ux, uy, [&](int, int t) {
          H.col.in_src.push_back((int)(D.state + t));
          loc.push_back((unsigned short)(2 * n_up + t));
        });
      }
      // staged only where the planes fit the kernel's scratch rows (bidi_columns16.cuh; 8*(rh+4) floats) or are tiny
      const int total = 2 * n_up + n_own;
      C.stage_n = (C.planar && !h_no_col_stage && total > 0 && total <= std::max(64, C.rh ? 8 * (C.rh + 4) : 0)) ? total : 0;
      H.col.in_loc.assign((loc.size() + 1) / 2 + 1, 0);
      memcpy(H.col.in_loc.data(), loc.data(), loc.size() * sizeof(unsigned short));
      ColPairTab& T = H.col.tab;
      memset(&T, 0, sizeof(T));
      const int n_pairs = (C.n + 1) / 2;
      if (C.planar && n_pairs <= BIDI_PAIRTAB_MAX && n_pairs > 0) {
        T.n = n_pairs;
        for (int p = 0; p <= n_pairs; p++) T.rec_off[p] = H.col.rec_off[p];
        for (int p = 0; p < n_pairs; p++) T.stride[p] = (short)H.col.stride[p];
        for (int i = 0; i <= 2 * n_pairs; i++) T.in_off[i] = H.col.in_off[std::min(i, C.n)];
      }
    }
  }
  if (V == BIDI_NEO_AGENT) { // visit order: layers top -> bottom, then the input nodes
    int base = 0;
    for (int l = L - 1; l >= 0; l--) { h->layers[l].col.noise_base = base; base += h->layers[l].col.n; }
    h->layers[L].col.noise_base = base; base += h->layers[L].col.n;
    h->ncol = base;
    h->n_noise = base * 3;
  }

  // ---- device memory
  if (cudaSetDevice(device) != cudaSuccess) { bidi_destroy(h); return fail(nullptr, BIDI_ENODEV, "bidi_create: cudaSetDevice failed"); }
  auto bail = [&](int code, const std::string& msg) { bidi_destroy(h); h = nullptr; return fail(nullptr, code, msg); };
#define CU_CREATE(call)                                                                            \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return bail(e_ == cudaErrorMemoryAllocation ? BIDI_ENOMEM : BIDI_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)
  CU_CREATE(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CU_CREATE(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
  CU_CREATE(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CU_CREATE(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  CU_CREATE(cudaEventCreate(&h->ev0));
  CU_CREATE(cudaEventCreate(&h->ev1));
  CU_CREATE(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
  const size_t A = (size_t)n_agents, nin = (size_t)h->n_in, nn = std::max<size_t>(1, (size_t)h->n_noise);
  CU_CREATE(cudaMalloc(&h->d_blob, sizeof(float) * A * (size_t)h->stride));
  CU_CREATE(cudaMemsetAsync(h->d_blob, 0, sizeof(float) * A * (size_t)h->stride, h->stream));
  CU_CREATE(cudaMalloc(&h->d_in0, sizeof(float) * A * nin));
  CU_CREATE(cudaMalloc(&h->d_pred, sizeof(float) * A * nin));
  CU_CREATE(cudaMalloc(&h->d_rewards, sizeof(float) * A));
  CU_CREATE(cudaMalloc(&h->d_noise_u, sizeof(float) * A * nn));
  CU_CREATE(cudaMalloc(&h->d_noise_v, sizeof(float) * A * nn));
  CU_CREATE(cudaMalloc(&h->d_col_actions, sizeof(float) * A * nn));
  if (V == BIDI_NEO_HIERARCHY) {  // simStepGenerate's N(0,1) draws: per layer iter x units
    int off = 0;
    for (int l = 0; l < L; l++) { h->layers[l].gen_off = off; off += layers[l].iter_settle * h->layers[l].nh; }
    h->gen_per_agent = off;
    CU_CREATE(cudaMalloc(&h->d_gen, sizeof(float) * A * (size_t)off));
    CU_CREATE(cudaMalloc(&h->d_gen_scale, sizeof(float)));
    CU_CREATE(cudaMemsetAsync(h->d_gen, 0, sizeof(float) * A * (size_t)off, h->stream));
    CU_CREATE(cudaMemsetAsync(h->d_gen_scale, 0, sizeof(float), h->stream));
    CU_CREATE(cudaMallocHost(&h->h_gen, sizeof(float) * (A * (size_t)off + 1)));
    h->device_bytes += (long long)sizeof(float) * (long long)(A * (size_t)off + 1);
  }
  CU_CREATE(cudaMemsetAsync(h->d_in0, 0, sizeof(float) * A * nin, h->stream));
  CU_CREATE(cudaMemsetAsync(h->d_pred, 0, sizeof(float) * A * nin, h->stream));
  CU_CREATE(cudaMemsetAsync(h->d_rewards, 0, sizeof(float) * A, h->stream));
  CU_CREATE(cudaMemsetAsync(h->d_noise_u, 0, sizeof(float) * A * nn, h->stream));
  CU_CREATE(cudaMemsetAsync(h->d_noise_v, 0, sizeof(float) * A * nn, h->stream));
  CU_CREATE(cudaMemsetAsync(h->d_col_actions, 0, sizeof(float) * A * nn, h->stream));
  h->device_bytes = (long long)sizeof(float) * (long long)(A * (size_t)h->stride + 2 * A * nin + A + 3 * A * nn);
  CU_CREATE(cudaMallocHost(&h->h_in, sizeof(float) * A * nin));
  CU_CREATE(cudaMallocHost(&h->h_pred, sizeof(float) * A * nin));
  CU_CREATE(cudaMallocHost(&h->h_rewards, sizeof(float) * A));
  CU_CREATE(cudaMallocHost(&h->h_noise, sizeof(float) * 2 * A * nn));
  CU_CREATE(cudaMallocHost(&h->h_actions, sizeof(float) * A * nn));

  // ---- integer tables: one device pool, pointers patched into the WinDev copies
  std::vector<int> pool;
  struct Patch { const int** dst; size_t off; };
  std::vector<Patch> patches;
  auto push = [&](const std::vector<int>& v, const int** dst) {
    patches.push_back({dst, pool.size()});
    pool.insert(pool.end(), v.begin(), v.end());
    if (v.empty()) pool.push_back(0);
  };
  for (int l = 0; l <= L; l++) {
    LayerHost& H = h->hl[l];
    for (WinHost* w : {&H.ff, &H.rec, &H.lat, &H.fb, &H.pred}) {
      push(w->cx, &w->d.cx); push(w->cy, &w->d.cy);
      push(w->gxlo, &w->d.gx_lo); push(w->gxhi, &w->d.gx_hi); push(w->gylo, &w->d.gy_lo); push(w->gyhi, &w->d.gy_hi);
    }
    if (l < L && H.dense_geom) {
      // the dense coder's learning mask: which slots of a unit's row [feed-forward | recurrent | pad] are its own connections
      const int nh = h->layers[l].nh, nf = H.row_nf, pq = H.row_pitch / 4;
      H.dense_mask.assign((size_t)pq * nh, 0);
      for (int u = 0; u < nh; u++) {
        const int ux = u % H.ff.d.uw, uy = u / H.ff.d.uw;
        for_each_conn_host(H.ff, ux, uy, [&](int s, int) { H.dense_mask[(size_t)u * pq + s / 4] |= 1 << (s % 4); });
        for_each_conn_host(H.rec, ux, uy, [&](int s, int) { H.dense_mask[(size_t)u * pq + (nf + s) / 4] |= 1 << ((nf + s) % 4); });
      }
      push(H.dense_mask, &h->layers[l].dense_mask);
    }
    if (csrl) {
      push(H.sd.in_off, &h->layers[l].sd.in_off);
      push(H.sd.in_src, &h->layers[l].sd.in_src);
    }
    if (V == BIDI_NEO_AGENT) {
      push(H.col.in_off, &h->layers[l].col.in_off);
      push(H.col.rec_off, &h->layers[l].col.rec_off);
      push(H.col.stride, &h->layers[l].col.stride);
      push(H.col.in_src, &h->layers[l].col.in_src);
      push(H.col.in_loc, reinterpret_cast<const int**>(&h->layers[l].col.in_loc));
    }
  }
  CU_CREATE(cudaMalloc(&h->d_tables, sizeof(int) * pool.size()));
  CU_CREATE(cudaMemcpyAsync(h->d_tables, pool.data(), sizeof(int) * pool.size(), cudaMemcpyHostToDevice, h->stream));
  for (auto& p : patches) *p.dst = h->d_tables + p.off;
  for (int l = 0; l <= L; l++) {
    LayerDev& D = h->layers[l];
    LayerHost& H = h->hl[l];
    D.ff = H.ff.d; D.rec = H.rec.d; D.lat = H.lat.d; D.fb = H.fb.d; D.pred = H.pred.d;
  }
  if (qnet) {  // one 0/1 mask per layer for all agents
    size_t tot = 0;
    for (int l = 0; l < L; l++) tot += h->q_keep[l].size();
    std::vector<float> mask;
    mask.reserve(tot);
    for (int l = 0; l < L; l++) for (char c : h->q_keep[l]) mask.push_back(c ? 1.0f : 0.0f);
    CU_CREATE(cudaMalloc(&h->d_qmask, sizeof(float) * std::max<size_t>(1, tot)));
    CU_CREATE(cudaMemcpyAsync(h->d_qmask, mask.data(), sizeof(float) * tot, cudaMemcpyHostToDevice, h->stream));
    CU_CREATE(cudaStreamSynchronize(h->stream));
    size_t off = 0;
    for (int l = 0; l < L; l++) { h->layers[l].q_mask = h->d_qmask + off; off += h->q_keep[l].size(); }
    h->device_bytes += (long long)(sizeof(float) * tot);
  }
  CU_CREATE(cudaMalloc(&h->d_layers, sizeof(LayerDev) * (L + 1)));
  CU_CREATE(cudaMemcpyAsync(h->d_layers, h->layers.data(), sizeof(LayerDev) * (L + 1), cudaMemcpyHostToDevice, h->stream));
  h->device_bytes += (long long)(sizeof(int) * pool.size() + sizeof(LayerDev) * (L + 1));

  EngineDev& P = h->P;
  P.variant = V; P.L = L; P.A = n_agents; P.n_in = h->n_in;
  P.blob = h->d_blob; P.stride = h->stride; P.in0 = h->d_in0; P.pred_out = h->d_pred;
  P.rewards = h->d_rewards; P.noise_u = h->d_noise_u; P.noise_v = h->d_noise_v; P.col_actions = h->d_col_actions;
  P.ncol = h->ncol; P.layers = h->d_layers;
  P.gen_on = 0; P.gen_per_agent = h->gen_per_agent; P.gen_normals = h->d_gen; P.gen_scale = h->d_gen_scale;
  P.q_on = qnet ? 1 : 0;
  P.csrl = csrl ? 1 : 0; P.n_noise = h->n_noise;
  if (csrl) {
    P.csrl_n_actions = cfg->q_n_actions; P.csrl_reward_offsets = csrl_offsets;
    P.csrl_expl_std = cfg->col_expl_stddev; P.csrl_expl_break = cfg->col_expl_break;
    std::vector<int> tab(cfg->q_action_idx, cfg->q_action_idx + cfg->q_n_actions);
    tab.insert(tab.end(), h->csrl_is_action.begin(), h->csrl_is_action.end());
    CU_CREATE(cudaMalloc(&h->d_csrl_tables, sizeof(int) * tab.size()));
    CU_CREATE(cudaMemcpyAsync(h->d_csrl_tables, tab.data(), sizeof(int) * tab.size(), cudaMemcpyHostToDevice, h->stream));
    CU_CREATE(cudaStreamSynchronize(h->stream));
    P.csrl_action_idx = h->d_csrl_tables; P.csrl_is_action = h->d_csrl_tables + cfg->q_n_actions;
    size_t sm = 0;
    for (int l = 0; l <= L; l++) sm = std::max(sm, sizeof(float) * BIDI_SDRRL_WARPS * (size_t)bidi::sdrrl_warp_floats(h->layers[l].sd.max_in));
    if (sm > 200 * 1024) return bail(BIDI_EINVAL, "bidi_create: an SDRRL unit has too many inputs for shared memory");
    CU_CREATE(cudaFuncSetAttribute(bidi::k_sdrrl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(sm, 48 * 1024)));
  }
  if (qnet) {
    P.q_half = h->q_half; P.q_top = h->q_top; P.q_iters = cfg->q_derive_iterations;
    P.q_qc_w = q_qc_w; P.q_qc_tr = q_qc_tr; P.q_act = q_act; P.q_prev = q_prev;
    P.q_alpha = cfg->q_alpha; P.q_action_alpha = cfg->q_action_alpha; P.q_derive_alpha = cfg->q_derive_alpha;
    P.q_expl_break = cfg->q_expl_break; P.q_gamma = cfg->q_gamma; P.q_gamma_lambda = cfg->q_gamma_lambda; P.q_expl_std = cfg->q_expl_stddev;
    std::vector<int> tab(h->q_act_index);
    tab.insert(tab.end(), h->q_a_input.begin(), h->q_a_input.end());
    CU_CREATE(cudaMalloc(&h->d_qtables, sizeof(int) * tab.size()));
    CU_CREATE(cudaMemcpyAsync(h->d_qtables, tab.data(), sizeof(int) * tab.size(), cudaMemcpyHostToDevice, h->stream));
    CU_CREATE(cudaStreamSynchronize(h->stream));
    P.q_act_index = h->d_qtables; P.q_a_input = h->d_qtables + h->n_in;
  }
  {  // constants the packed-pair arithmetic must not see at compile time (bidi_columns.cuh)
    const float one[2] = {1.0f, 1.0f}, negzero[2] = {-0.0f, -0.0f}, negone[2] = {-1.0f, -1.0f};
    memcpy(&P.k_one2, one, 8); memcpy(&P.k_negzero2, negzero, 8); memcpy(&P.k_negone2, negone, 8);
  }
  CU_CREATE(cudaStreamSynchronize(h->stream));

  // thresholds start at init_threshold (sdr/RSDR.cpp:43, neo/Column.cpp:27); everything else at 0
  {
    std::vector<float> img((size_t)h->stride, 0.0f);
    for (int l = 0; l <= L; l++) {
      const LayerDev& D = h->layers[l];
      if (l < L) for (int i = 0; i < D.nh; i++) img[D.thr + i] = cfg->init_threshold;
      if (V == BIDI_NEO_AGENT)
        for (int node = 0; node < D.col.n; node++) {
          const ColRec r = col_rec(D.col, h->hl[l].col, img.data(), node);
          for (int i = 0; i < D.col.cells; i++) r.thr[r.cs * i] = cfg->init_threshold;
        }
      if (csrl) for (long long k = 0; k < (long long)D.sd.n * D.sd.cells; k++) img[D.sd.thr + k] = cfg->init_threshold;
    }
    for (int a = 0; a < n_agents; a++)
      CU_CREATE(cudaMemcpyAsync(h->d_blob + (long long)a * h->stride, img.data(), sizeof(float) * h->stride, cudaMemcpyHostToDevice, h->stream));
    CU_CREATE(cudaStreamSynchronize(h->stream));
  }
  // layers whose planes fit in shared memory use the fused coder kernel
  h->coder_cta.assign(L, 0);
  if (V != BIDI_KWINNERS) {
    size_t smem = 0, smem_dense = 0;
    for (int l = 0; l < L; l++) {
      const size_t need = sizeof(float) * (size_t)bidi::coder_cta_shape(h->layers[l]).floats;
      const size_t need_dense = sizeof(float) * (size_t)bidi::coder_dense_shape(h->layers[l]).floats;
      if (bidi::coder_dense_applies(h->layers[l]) && need_dense <= 200 * 1024 && !h->no_dense_coder && !csrl) { h->coder_cta[l] = 2; smem_dense = std::max(smem_dense, need_dense); }
      else if (need <= 200 * 1024) { h->coder_cta[l] = 1; smem = std::max(smem, need); }
    }
    if (smem > 0) {
      CU_CREATE(cudaFuncSetAttribute(bidi::k_coder_cta<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CU_CREATE(cudaFuncSetAttribute(bidi::k_coder_cta<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (smem_dense > 0) {
      CU_CREATE(cudaFuncSetAttribute(bidi::k_coder_dense<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dense));
      CU_CREATE(cudaFuncSetAttribute(bidi::k_coder_dense<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dense));
      CU_CREATE(cudaFuncSetAttribute(bidi::k_coder_dense<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dense));
      CU_CREATE(cudaFuncSetAttribute(bidi::k_coder_dense<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dense));
    }
  }
  h->dense_rows.assign(L, 0);
  // node layers with few (agent, unit) threads use the tile kernels
  h->tiled.assign(L + 1, 0);
  {
    size_t sm = 0;
    for (int l = 0; l <= L; l++) {
      if (l == L && V == BIDI_KWINNERS) break;
      const LayerDev& Y = h->layers[l];
      const int units = l < L ? std::max(Y.nh, Y.nv) : Y.pn;
      size_t need = tile_bytes(nslots_of(Y.fb) + nslots_of(Y.pred));
      if (l < L) {
        need = std::max(need, tile_bytes(nslots_of(Y.ff) + nslots_of(Y.rec) + nslots_of(Y.lat)));
        need = std::max(need, tile_bytes(std::max(Y.ff.max_cand, Y.rec.max_cand)));
      }
      if ((long long)n_agents * units <= 32768 && need <= 200 * 1024) { h->tiled[l] = 1; sm = std::max(sm, need); }
    }
    if (sm > 0) {
      CU_CREATE(cudaFuncSetAttribute(bidi::k_t_kw_activate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      CU_CREATE(cudaFuncSetAttribute(bidi::k_t_reconstruct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      CU_CREATE(cudaFuncSetAttribute(bidi::k_t_ic_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      CU_CREATE(cudaFuncSetAttribute(bidi::k_t_predict, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    }
  }
  if (select_grid_kernels(h) != BIDI_OK) return bail(BIDI_ECUDA, g_last_error);
  if (select_persist(h) != BIDI_OK) return bail(BIDI_ECUDA, g_last_error);
  if (apply_layout(h) != BIDI_OK) return bail(BIDI_ECUDA, g_last_error);  // (after the kernel families are chosen: they decide the layouts)
  // the column kernel may need more than the default 48 KB of dynamic shared memory
  if (V == BIDI_NEO_AGENT) {
    size_t smem = 0;
    for (auto& D : h->layers) if (D.col.n > 0) smem = std::max(smem, columns_smem_bytes(D.col));
    if (smem > 227 * 1024) return bail(BIDI_EINVAL, "bidi_create: a column record does not fit in shared memory");
    CU_CREATE(cudaFuncSetAttribute(bidi::k_columns<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU_CREATE(cudaFuncSetAttribute(bidi::k_columns<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU_CREATE(cudaFuncSetAttribute(bidi::k_columns16<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU_CREATE(cudaFuncSetAttribute(bidi::k_columns16<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU_CREATE(cudaFuncSetAttribute(bidi::k_columns16<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU_CREATE(cudaFuncSetAttribute(bidi::k_columns16<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  *out = h;
  return BIDI_OK;
}

int bidi_num_agents(const bidi_engine* h) { return h ? h->A : 0; }
int bidi_num_inputs(const bidi_engine* h) { return h ? h->n_in : 0; }
int bidi_num_columns(const bidi_engine* h) { return h ? h->ncol : 0; }
int bidi_num_noise(const bidi_engine* h) { return h ? h->n_noise : 0; }
long long bidi_launch_count(const bidi_engine* h) { return h ? h->launches : 0; }
long long bidi_device_bytes(const bidi_engine* h) { return h ? h->device_bytes : 0; }

long long bidi_array_len(const bidi_engine* h, int which, int layer) {
  if (!h) return -1;
  return find_array(h, which, layer).len;
}

int bidi_conn_counts(const bidi_engine* h, int which, int layer, int* counts, int n_units) {
  bidi_engine* hm = const_cast<bidi_engine*>(h);
  if (!h || !counts || layer < 0 || layer > h->L) return fail(hm, BIDI_EINVAL, "bidi_conn_counts: bad argument");
  const LayerHost& H = h->hl[layer];
  const WinHost* w = nullptr;
  switch (which) {
    case BIDI_FF_W: w = &H.ff; break;
    case BIDI_REC_W: w = &H.rec; break;
    case BIDI_LAT_W: w = &H.lat; break;
    case BIDI_FB_W: w = &H.fb; break;
    case BIDI_PRED_W: w = &H.pred; break;
    case BIDI_Q_FF_W: {
      if (!h->qnet || layer >= h->L || n_units != h->layers[layer].nh) return fail(hm, BIDI_EINVAL, "bidi_conn_counts: bad Q network query");
      const WinHost& qw = H.ff;
      for (int i = 0; i < n_units; i++) {
        int n = 0;
        for_each_conn_host(qw, i % qw.d.uw, i / qw.d.uw, [&](int s, int) { n += h->q_keep[layer][(size_t)s * qw.d.U + i]; });
        counts[i] = n;
      }
      return BIDI_OK;
    }
    case BIDI_COL_INPUT:
      if (h->csrl && n_units == h->layers[layer].sd.n) {
        for (int i = 0; i < n_units; i++) counts[i] = H.sd.in_off[i + 1] - H.sd.in_off[i];
        return BIDI_OK;
      }
      if (h->cfg.variant != BIDI_NEO_AGENT || n_units != h->layers[layer].col.n) return fail(hm, BIDI_EINVAL, "bidi_conn_counts: bad column query");
      for (int i = 0; i < n_units; i++) counts[i] = H.col.in_off[i + 1] - H.col.in_off[i];
      return BIDI_OK;
    default: return fail(hm, BIDI_EINVAL, "bidi_conn_counts: not a list array");
  }
  if (n_units != w->d.U || w->cx.empty()) return fail(hm, BIDI_EINVAL, "bidi_conn_counts: unit count mismatch");
  for (int i = 0; i < n_units; i++) {
    int n = 0;
    for_each_conn_host(*w, i % w->d.uw, i / w->d.uw, [&](int, int) { n++; });
    counts[i] = n;
  }
  return BIDI_OK;
}

int bidi_sync(bidi_engine* h) {
  if (!h) return fail(h, BIDI_EINVAL, "null engine");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  return BIDI_OK;
}

int bidi_set_array(bidi_engine* h, int agent, int which, int layer, const float* src, long long len) {
  if (!h || !src || agent < 0 || agent >= h->A) return fail(h, BIDI_EINVAL, "bidi_set_array: bad argument");
  CU(cudaSetDevice(h->device));
  const ArrayRef r = find_array(h, which, layer);
  if (r.kind == ArrayRef::NONE || r.len != len) return fail(h, BIDI_EINVAL, "bidi_set_array: no such array or length mismatch");
  if (r.kind == ArrayRef::DENSE_IN || r.kind == ArrayRef::DENSE_PRED) {
    float* dst = (r.kind == ArrayRef::DENSE_IN ? h->d_in0 : h->d_pred) + (long long)agent * h->n_in;
    CU(cudaMemcpyAsync(dst, src, sizeof(float) * len, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return BIDI_OK;
  }
  if (h->ipc_linked) return fail(h, BIDI_EINVAL, "bidi_set_array: not after bidi_strip_link_ipc (the neighbours store into this memory): initialise the state before bidi_strip_ipc_export");
  int rc = mirror_load(h, agent);
  if (rc) return rc;
  if (which == BIDI_COL_INPUT) h->col_inputs_set[agent] = 1;
  h->state_touched = true;
  if (r.kind == ArrayRef::PLANE) memcpy(h->mirror.data() + r.off, src, sizeof(float) * len);
  else if (r.kind == ArrayRef::COLUMN) walk_column(h, r.which, r.layer, h->mirror.data(), [&](float& v) { v = *src++; });
  else if (r.kind == ArrayRef::QLIST) walk_qlist(h, r.layer, r.trace, h->mirror.data(), [&](float& v) { v = *src++; });
  else walk_list(*r.win, r.trace, h->mirror.data(), [&](float& v) { v = *src++; });
  h->mirror_dirty = true;
  return BIDI_OK;
}

int bidi_get_array(bidi_engine* h, int agent, int which, int layer, float* dst, long long len) {
  if (!h || !dst || agent < 0 || agent >= h->A) return fail(h, BIDI_EINVAL, "bidi_get_array: bad argument");
  CU(cudaSetDevice(h->device));
  const ArrayRef r = find_array(h, which, layer);
  if (r.kind == ArrayRef::NONE || r.len != len) return fail(h, BIDI_EINVAL, "bidi_get_array: no such array or length mismatch");
  if (r.kind == ArrayRef::DENSE_IN || r.kind == ArrayRef::DENSE_PRED) {
    const float* s = (r.kind == ArrayRef::DENSE_IN ? h->d_in0 : h->d_pred) + (long long)agent * h->n_in;
    CU(cudaMemcpyAsync(dst, s, sizeof(float) * len, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return BIDI_OK;
  }
  int rc = mirror_load(h, agent);
  if (rc) return rc;
  if (r.kind == ArrayRef::PLANE) memcpy(dst, h->mirror.data() + r.off, sizeof(float) * len);
  else if (r.kind == ArrayRef::COLUMN) walk_column(h, r.which, r.layer, h->mirror.data(), [&](float& v) { *dst++ = v; });
  else if (r.kind == ArrayRef::QLIST) walk_qlist(h, r.layer, r.trace, h->mirror.data(), [&](float& v) { *dst++ = v; });
  else walk_list(*r.win, r.trace, h->mirror.data(), [&](float& v) { *dst++ = v; });
  return BIDI_OK;
}

// The reference's createRandom draw order, replayed with the same
// standard-library engine and distribution the reference uses.
// `given`: draw from the caller's generator instead of std::mt19937(seed_base + agent) (one agent; bidi_init_random_generator)
static int init_random_impl(bidi_engine* h, int first, int count, unsigned seed_base, std::mt19937* given) {
  if (!h || first < 0 || count < 0 || first + count > h->A) return fail(h, BIDI_EINVAL, "bidi_init_random: bad agent range");
  if (h->ipc_linked) return fail(h, BIDI_EINVAL, "bidi_init_random: not after bidi_strip_link_ipc (the neighbours store into this memory): initialise the state before bidi_strip_ipc_export");
  CU(cudaSetDevice(h->device));
  int rc = mirror_release(h);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));  // a queued step may still be reading or writing the agents being re-seeded
  h->state_touched = true;
  const int L = h->L, V = h->cfg.variant;
  const bidi_config& g = h->cfg;
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const int nthreads = given ? 1 : (int)std::min<unsigned>(hw, (unsigned)std::max(1, count));
  std::vector<std::vector<float>> images(nthreads);
  std::vector<std::string> errors(nthreads);
  auto fill_agent = [&](std::vector<float>& img, std::mt19937& gen) {
    std::uniform_real_distribution<float> weightDist(g.init_min_w, g.init_max_w);
    std::uniform_real_distribution<float> inhibitionDist(g.init_min_inh, g.init_max_inh);
    std::fill(img.begin(), img.end(), 0.0f);
    auto draw_list = [&](const WinHost& w, long long u, int ux, int uy, std::uniform_real_distribution<float>& dist) {
      if (w.d.w_off < 0) return;
      float* base = img.data() + w.d.w_off;
      const long long ss = w.d.row_pitch ? 1 : w.d.U, us = w.d.row_pitch ? w.d.row_pitch : 1;
      for_each_conn_host(w, ux, uy, [&](int s, int) { base[(long long)s * ss + u * us] = dist(gen); });
    };
    auto draw_column = [&](const ColumnsDev& C, const ColumnsHost& CH, int node) { // neo/Column.cpp:22-43
      const ColRec r = col_rec(C, CH, img.data(), node);
      for (int i = 0; i < C.cells; i++) {
        r.thr[r.cs * i] = g.init_threshold;
        for (int j = 0; j < r.nin; j++) r.w(i, j) = weightDist(gen);
        for (int j = 0; j < C.cells; j++) r.lat(i, j) = inhibitionDist(gen);
        r.q_w[r.cs * i] = weightDist(gen);
      }
      for (int a = 0; a < 3; a++)
        for (int k = 0; k < C.cells; k++) r.a_w[r.es * (a * r.Cq + k)] = weightDist(gen);
    };
    auto draw_sdrrl = [&](const SdrrlDev& S, const SdrrlHost& SH, int node) {  // deep/SDRRL.cpp:23-43
      const int in0 = SH.in_off[node], nin = SH.in_off[node + 1] - in0;
      for (int i = 0; i < S.cells; i++) {
        const long long ck = (long long)node * S.cells + i;
        img[S.thr + ck] = g.init_threshold;
        for (int j = 0; j < nin; j++) img[S.ff_w + (long long)in0 * S.cells + (long long)i * nin + j] = weightDist(gen);
        for (int j = 0; j < S.cells; j++) img[S.lat_w + ck * S.cells + j] = inhibitionDist(gen);
        for (int j = 0; j < S.na; j++) img[S.a_w + ck * S.na + j] = weightDist(gen);
        img[S.q_w + ck] = weightDist(gen);
      }
    };
    if (h->csrl) {  // _lastLayerRewardOffsets, deep/CSRL.cpp:32-35
      std::uniform_real_distribution<float> dist01(0.0f, 1.0f);
      for (int i = 0; i < h->layers[L - 1].nh; i++) img[h->P.csrl_reward_offsets + i] = dist01(gen) * 2.0f - 1.0f;
    }
    for (int l = 0; l <= L; l++) {
      if (l == L && V == BIDI_KWINNERS) break;
      const LayerDev& D = h->layers[l];
      const LayerHost& H = h->hl[l];
      if (l < L) {
        for (int i = 0; i < D.nh; i++) {
          const int ux = i % D.hw, uy = i / D.hw;
          img[D.thr + i] = g.init_threshold;
          draw_list(H.ff, i, ux, uy, weightDist);
          draw_list(H.rec, i, ux, uy, weightDist);
          if (is_iter(V)) draw_list(H.lat, i, ux, uy, inhibitionDist);
        }
      }
      for (int i = 0; i < D.pn; i++) {
        const int ux = i % H.fb.d.uw, uy = i / H.fb.d.uw;
        (void)weightDist(gen); // the node's bias: drawn, never used (sdr/PredictiveRSDR.cpp:39)
        draw_list(H.fb, i, ux, uy, weightDist);
        if (l < L) draw_list(H.pred, i, ux, uy, weightDist);
        if (V == BIDI_NEO_AGENT) draw_column(D.col, H.col, i);
        if (h->csrl) draw_sdrrl(D.sd, H.sd, i);
      }
    }
    if (h->qnet) {  // sdr/QPRSDR.cpp:30-31, then per layer and node: bias, one draw per in-bounds slot -- kept or not (:50-83)
      for (int i = 0; i < h->q_top; i++) img[h->P.q_qc_w + i] = weightDist(gen) / static_cast<float>(h->q_top);
      for (int l = 0; l < L; l++) {
        const LayerDev& D = h->layers[l];
        const WinHost& w = h->hl[l].ff;
        const long long U = w.d.U;
        for (int i = 0; i < D.nh; i++) {
          img[D.q_bias_w + i] = weightDist(gen);
          for_each_conn_host(w, i % D.hw, i / D.hw, [&](int s, int) {
            const float v = weightDist(gen);
            if (h->q_keep[l][(size_t)s * U + i]) img[D.q_w_off + (long long)s * U + i] = v;
          });
        }
      }
    }
  };
  std::vector<std::thread> pool;
  std::vector<cudaError_t> cuerr(nthreads, cudaSuccess);
  for (int t = 0; t < nthreads; t++)
    pool.emplace_back([&, t]() {
      // every worker uploads on its own stream and waits for each copy to land before it reuses the image; the engine's
      // stream is idle (synchronised above) and nothing is queued on it before all workers have joined
      cudaSetDevice(h->device);
      cudaStream_t up = nullptr;
      cudaError_t e = cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking);
      if (e != cudaSuccess) { cuerr[t] = e; return; }
      images[t].resize((size_t)h->stride);
      for (int a = first + t; a < first + count; a += nthreads) {
        std::mt19937 own(seed_base + (unsigned)a);
        fill_agent(images[t], given ? *given : own);
        e = cudaMemcpyAsync(h->d_blob + (long long)a * h->stride, images[t].data(), sizeof(float) * h->stride, cudaMemcpyHostToDevice, up);
        if (e == cudaSuccess) e = cudaStreamSynchronize(up);
        if (e != cudaSuccess) { cuerr[t] = e; break; }
      }
      cudaStreamDestroy(up);
    });
  for (auto& t : pool) t.join();
  for (auto e : cuerr) if (e != cudaSuccess) return fail(h, BIDI_ECUDA, std::string("bidi_init_random upload: ") + cudaGetErrorString(e));
  return BIDI_OK;
}

int bidi_init_random(bidi_engine* h, int first, int count, unsigned seed_base) { return init_random_impl(h, first, count, seed_base, nullptr); }

int bidi_init_random_generator(bidi_engine* h, int agent, const char* state_in, char* state_out, int state_out_cap) {
  if (!h || !state_in) return fail(h, BIDI_EINVAL, "bidi_init_random_generator: null argument");
  std::mt19937 gen;
  std::istringstream is(state_in);
  is >> gen;
  if (is.fail()) return fail(h, BIDI_EINVAL, "bidi_init_random_generator: state_in is not the text form of a std::mt19937");
  const int rc = init_random_impl(h, agent, 1, 0u, &gen);
  if (rc) return rc;
  if (state_out) {
    std::ostringstream os;
    os << gen;
    const std::string t = os.str();
    if ((int)t.size() + 1 > state_out_cap) return fail(h, BIDI_EINVAL, "bidi_init_random_generator: state_out too small");
    memcpy(state_out, t.c_str(), t.size() + 1);
  }
  return BIDI_OK;
}

int bidi_set_inputs(bidi_engine* h, const float* in) {
  if (!h || !in) return fail(h, BIDI_EINVAL, "bidi_set_inputs: null argument");
  CU(cudaSetDevice(h->device));
  const size_t bytes = sizeof(float) * (size_t)h->A * h->n_in;
  CU(cudaStreamSynchronize(h->stream)); // the staging buffer may still be in flight
  if (h->strip_n > 1) {  // one part of a split layer reads only the input rows its windows reach
    const size_t w = (size_t)h->cfg.in_width;
    const size_t lo = std::min(h->strip_span[BIDI_SPAN_VIS_RECON][0], h->strip_span[BIDI_SPAN_OWN_VIS][0]) * w;
    const size_t hi = std::max(h->strip_span[BIDI_SPAN_VIS_RECON][1], h->strip_span[BIDI_SPAN_OWN_VIS][1]) * w;
    for (int a = 0; a < h->A; a++) {
      const size_t o = (size_t)a * h->n_in + lo;
      if (in != h->h_in) memcpy(h->h_in + o, in + o, sizeof(float) * (hi - lo));
      CU(cudaMemcpyAsync(h->d_in0 + o, h->h_in + o, sizeof(float) * (hi - lo), cudaMemcpyHostToDevice, h->stream));
    }
    CU(cudaEventRecord(h->ev_in, h->stream));
    return BIDI_OK;
  }
  if (in != h->h_in) memcpy(h->h_in, in, bytes);  // bidi_host_inputs: the caller wrote the pinned buffer itself
  CU(cudaMemcpyAsync(h->d_in0, h->h_in, bytes, cudaMemcpyHostToDevice, h->stream));
  CU(cudaEventRecord(h->ev_in, h->stream));
  return BIDI_OK;
}

int bidi_get_inputs(bidi_engine* h, float* out) {
  if (!h || !out) return fail(h, BIDI_EINVAL, "bidi_get_inputs: null argument");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(out, h->d_in0, sizeof(float) * (size_t)h->A * h->n_in, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BIDI_OK;
}

float* bidi_host_inputs(bidi_engine* h) { return h ? h->h_in : nullptr; }
const float* bidi_host_predictions(bidi_engine* h) { return h ? h->h_pred : nullptr; }

int bidi_wait_inputs(bidi_engine* h) {
  if (!h) return fail(h, BIDI_EINVAL, "null engine");
  CU(cudaSetDevice(h->device));
  CU(cudaEventSynchronize(h->ev_in));
  return BIDI_OK;
}

int bidi_set_input(bidi_engine* h, int agent, int index, float value) {
  if (!h || agent < 0 || agent >= h->A || index < 0 || index >= h->n_in) return fail(h, BIDI_EINVAL, "bidi_set_input: out of range");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(h->d_in0 + (long long)agent * h->n_in + index, &value, sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BIDI_OK;
}

int bidi_step(bidi_engine* h, const float* rewards, const float* noise_u, const float* noise_v, int learn) {
  if (!h) return fail(h, BIDI_EINVAL, "null engine");
  if ((noise_u == nullptr) != (noise_v == nullptr)) return fail(h, BIDI_EINVAL, "bidi_step: noise_u and noise_v go together");
  CU(cudaSetDevice(h->device));
  const bool agent = h->cfg.variant == BIDI_NEO_AGENT || h->qnet || h->csrl;  // variants with a reward and exploration draws
  if (agent) {
    CU(cudaStreamSynchronize(h->stream));
    if (rewards) memcpy(h->h_rewards, rewards, sizeof(float) * h->A);
    else memset(h->h_rewards, 0, sizeof(float) * h->A);
    CU(cudaMemcpyAsync(h->d_rewards, h->h_rewards, sizeof(float) * h->A, cudaMemcpyHostToDevice, h->stream));
    if (noise_u) {
      const size_t n = (size_t)h->A * h->n_noise;
      memcpy(h->h_noise, noise_u, sizeof(float) * n);
      memcpy(h->h_noise + n, noise_v, sizeof(float) * n);
      CU(cudaMemcpyAsync(h->d_noise_u, h->h_noise, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
      CU(cudaMemcpyAsync(h->d_noise_v, h->h_noise + n, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    }
  }
  return run_step(h, learn != 0, agent && noise_u == nullptr);
}

int bidi_num_generate_noise(const bidi_engine* h) { return h ? h->gen_per_agent : 0; }

int bidi_step_generate(bidi_engine* h, const float* normals, float noise) {
  if (!h || h->cfg.variant != BIDI_NEO_HIERARCHY || h->gen_per_agent == 0)
    return fail(h, BIDI_EINVAL, "bidi_step_generate: only neo::PredictiveHierarchy has simStepGenerate");
  CU(cudaSetDevice(h->device));
  int rc = mirror_release(h);
  if (rc) return rc;
  const size_t n = (size_t)h->A * h->gen_per_agent;
  h->state_touched = true;
  CU(cudaStreamSynchronize(h->stream));  // the staging buffer may still be in flight
  if (normals) {
    memcpy(h->h_gen, normals, sizeof(float) * n);
    CU(cudaMemcpyAsync(h->d_gen, h->h_gen, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
  } else {
    bidi::k_gen_normals<<<grid_for((long long)n, kBlock), kBlock, 0, h->stream>>>(h->d_gen, (long long)n, h->noise_seed ^ 0x6E6F697365ULL, h->step_index,
                                                                                 h->first_agent * h->gen_per_agent);
    h->launches++;
  }
  h->h_gen[n] = noise;
  CU(cudaMemcpyAsync(h->d_gen_scale, h->h_gen + n, sizeof(float), cudaMemcpyHostToDevice, h->stream));
  rc = ensure_graph(h, 2);
  if (rc) return rc;
  CU(cudaGraphLaunch(h->graphs[2], h->stream));
  h->launches += h->graph_kernels[2];
  h->step_index++;
  return BIDI_OK;
}

int bidi_get_predictions(bidi_engine* h, float* out) {
  if (!h || !out) return fail(h, BIDI_EINVAL, "bidi_get_predictions: null argument");
  CU(cudaSetDevice(h->device));
  const size_t bytes = sizeof(float) * (size_t)h->A * h->n_in;
  if (h->strip_n > 1) {  // one part of a split layer: only the rows it owns are meaningful, only they are copied
    const size_t w = (size_t)h->cfg.in_width;
    const size_t lo = h->strip_span[BIDI_SPAN_OWN_VIS][0] * w, hi = h->strip_span[BIDI_SPAN_OWN_VIS][1] * w;
    for (int a = 0; a < h->A; a++) {
      const size_t o = (size_t)a * h->n_in + lo;
      CU(cudaMemcpyAsync(h->h_pred + o, h->d_pred + o, sizeof(float) * (hi - lo), cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    for (int a = 0; a < h->A; a++) {
      const size_t o = (size_t)a * h->n_in + lo;
      if (out != h->h_pred) memcpy(out + o, h->h_pred + o, sizeof(float) * (hi - lo));
    }
    return BIDI_OK;
  }
  CU(cudaMemcpyAsync(h->h_pred, h->d_pred, bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (out != h->h_pred) memcpy(out, h->h_pred, bytes);  // bidi_host_predictions: the caller reads the pinned buffer itself
  return BIDI_OK;
}

int bidi_get_column_actions(bidi_engine* h, float* out) {
  if (!h || !out || h->ncol == 0) return fail(h, BIDI_EINVAL, "bidi_get_column_actions: no columns");
  CU(cudaSetDevice(h->device));
  const size_t bytes = sizeof(float) * (size_t)h->A * h->ncol * 3;
  CU(cudaMemcpyAsync(h->h_actions, h->d_col_actions, bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  memcpy(out, h->h_actions, bytes);
  return BIDI_OK;
}

int bidi_get_actions(bidi_engine* h, float* out) {
  if (!h || !out || !h->qnet) return fail(h, BIDI_EINVAL, "bidi_get_actions: not a QPRSDR engine");
  CU(cudaSetDevice(h->device));
  const size_t row = sizeof(float) * 2 * (size_t)h->q_half;
  CU(cudaMemcpy2DAsync(out, row, h->d_blob + h->P.q_act + 4 * h->q_half, sizeof(float) * (size_t)h->stride, row, (size_t)h->A,
                       cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BIDI_OK;
}

/* which of a layer's in-bounds feed-forward window slots the Q network keeps, list order (sdr/QPRSDR.cpp:72-81) */
int bidi_q_kept(const bidi_engine* h, int layer, unsigned char* kept, long long len) {
  bidi_engine* hm = const_cast<bidi_engine*>(h);
  if (!h || !kept || !h->qnet || layer < 0 || layer >= h->L) return fail(hm, BIDI_EINVAL, "bidi_q_kept: bad argument");
  const WinHost& w = h->hl[layer].ff;
  long long k = 0;
  for (int uy = 0; uy < w.d.uh; uy++)
    for (int ux = 0; ux < w.d.uw; ux++) {
      const long long u = ux + (long long)uy * w.d.uw;
      for_each_conn_host(w, ux, uy, [&](int s, int) { if (k < len) kept[k] = (unsigned char)h->q_keep[layer][(size_t)s * w.d.U + u]; k++; });
    }
  return k == len ? BIDI_OK : fail(hm, BIDI_EINVAL, "bidi_q_kept: length mismatch");
}

int bidi_load_frames(bidi_engine* h, const float* frames, const float* rewards, int n_frames) {
  if (!h || !frames || n_frames < 1) return fail(h, BIDI_EINVAL, "bidi_load_frames: bad argument");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  cudaFree(h->d_frames); cudaFree(h->d_frame_rewards);
  h->d_frames = nullptr; h->d_frame_rewards = nullptr; h->n_frames = 0;
  const size_t n = (size_t)n_frames * h->A * h->n_in;
  CU(cudaMalloc(&h->d_frames, sizeof(float) * n));
  CU(cudaMemcpyAsync(h->d_frames, frames, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
  if (rewards) {
    CU(cudaMalloc(&h->d_frame_rewards, sizeof(float) * (size_t)n_frames * h->A));
    CU(cudaMemcpyAsync(h->d_frame_rewards, rewards, sizeof(float) * (size_t)n_frames * h->A, cudaMemcpyHostToDevice, h->stream));
  }
  CU(cudaStreamSynchronize(h->stream));
  h->n_frames = n_frames;
  return BIDI_OK;
}

int bidi_step_frame(bidi_engine* h, int t, int learn) {
  if (!h || h->n_frames < 1 || t < 0) return fail(h, BIDI_EINVAL, "bidi_step_frame: no frames loaded");
  CU(cudaSetDevice(h->device));
  const int f = t % h->n_frames;
  const size_t n = (size_t)h->A * h->n_in;
  CU(cudaMemcpyAsync(h->d_in0, h->d_frames + (size_t)f * n, sizeof(float) * n, cudaMemcpyDeviceToDevice, h->stream));
  if (h->d_frame_rewards)
    CU(cudaMemcpyAsync(h->d_rewards, h->d_frame_rewards + (size_t)f * h->A, sizeof(float) * h->A, cudaMemcpyDeviceToDevice, h->stream));
  if (h->fb_n > 0) {
    bidi::k_feedback<<<grid_for(h->A, kBlock), kBlock, 0, h->stream>>>(h->P, h->fb_n, h->d_fb_src, h->d_fb_dst, h->fb_scale, h->fb_bias,
                                                                      h->fb_std, h->noise_seed ^ 0xFEEDULL, h->step_index,
                                                                      h->first_agent, h->fb_reward ? h->d_rewards : nullptr);
    h->launches++;
  }
  bool device_noise = h->cfg.variant == BIDI_NEO_AGENT || h->qnet || h->csrl;
  if (device_noise && h->n_noise_frames > 0) {  // the caller's draws (bidi_load_frame_noise) instead of the engine's generator
    const size_t nn = (size_t)h->A * h->n_noise, fo = (size_t)(t % h->n_noise_frames) * nn;
    CU(cudaMemcpyAsync(h->d_noise_u, h->d_frame_noise_u + fo, sizeof(float) * nn, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_noise_v, h->d_frame_noise_v + fo, sizeof(float) * nn, cudaMemcpyDeviceToDevice, h->stream));
    device_noise = false;
  }
  return run_step(h, learn != 0, device_noise);
}

int bidi_load_frame_noise(bidi_engine* h, const float* noise_u, const float* noise_v, int n_frames) {
  if (!h || n_frames < 0 || (n_frames > 0 && (!noise_u || !noise_v))) return fail(h, BIDI_EINVAL, "bidi_load_frame_noise: bad argument");
  if (n_frames > 0 && h->n_noise == 0) return fail(h, BIDI_EINVAL, "bidi_load_frame_noise: this variant draws no exploration noise");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  cudaFree(h->d_frame_noise_u); cudaFree(h->d_frame_noise_v);
  h->d_frame_noise_u = h->d_frame_noise_v = nullptr; h->n_noise_frames = 0;
  if (n_frames == 0) return BIDI_OK;
  const size_t n = (size_t)n_frames * h->A * h->n_noise;
  CU(cudaMalloc(&h->d_frame_noise_u, sizeof(float) * n));
  CU(cudaMalloc(&h->d_frame_noise_v, sizeof(float) * n));
  CU(cudaMemcpyAsync(h->d_frame_noise_u, noise_u, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->d_frame_noise_v, noise_v, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->n_noise_frames = n_frames;
  return BIDI_OK;
}

int bidi_set_noise_stream(bidi_engine* h, unsigned long long seed, long long first_global_agent) {
  if (!h || first_global_agent < 0) return fail(h, BIDI_EINVAL, "bidi_set_noise_stream: bad argument");
  h->noise_seed = seed; h->first_agent = first_global_agent;
  return BIDI_OK;
}

int bidi_get_noise(bidi_engine* h, float* noise_u, float* noise_v) {
  if (!h || h->n_noise == 0) return fail(h, BIDI_EINVAL, "bidi_get_noise: this variant draws no exploration noise");
  CU(cudaSetDevice(h->device));
  const size_t bytes = sizeof(float) * (size_t)h->A * h->n_noise;
  if (noise_u) CU(cudaMemcpyAsync(noise_u, h->d_noise_u, bytes, cudaMemcpyDeviceToHost, h->stream));
  if (noise_v) CU(cudaMemcpyAsync(noise_v, h->d_noise_v, bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BIDI_OK;
}

// ---- batched headless environment (bidi_env.cuh)
int bidi_env_attach(bidi_engine* h, int kind) {
  if (!h || kind != BIDI_ENV_RUNNER) return fail(h, BIDI_EINVAL, "bidi_env_attach: unknown environment");
  if (h->n_in < 20 + BIDI_ENV_JOINTS) return fail(h, BIDI_EINVAL, "bidi_env_attach: the runner needs at least 30 inputs (15 state, 4 clocks, 1 spare, 10 actions)");
  static_assert(BIDI_ENV_FLOATS == BIDI_ENV_STATE_FLOATS, "header and kernel disagree on the environment's state size");
  CU(cudaSetDevice(h->device));
  if (!h->d_env) {
    CU(cudaMalloc(&h->d_env, sizeof(float) * (size_t)h->A * BIDI_ENV_FLOATS));
    CU(cudaMalloc(&h->d_env_noise, sizeof(float) * (size_t)h->A * BIDI_ENV_JOINTS));
    h->device_bytes += (long long)sizeof(float) * h->A * (BIDI_ENV_FLOATS + BIDI_ENV_JOINTS);
  }
  h->env_kind = kind;
  return bidi_env_reset(h);
}

int bidi_env_reset(bidi_engine* h) {
  if (!h || !h->env_kind) return fail(h, BIDI_EINVAL, "bidi_env_reset: no environment attached");
  CU(cudaSetDevice(h->device));
  bidi::k_env_reset<<<grid_for(h->A, kBlock), kBlock, 0, h->stream>>>(bidi::env_default_params(), h->d_env, h->A);
  h->launches++;
  h->env_steps = 0;
  return BIDI_OK;
}

int bidi_env_step(bidi_engine* h, const float* noise_u, const float* noise_v, const float* env_noise, int learn) {
  if (!h || !h->env_kind) return fail(h, BIDI_EINVAL, "bidi_env_step: no environment attached");
  if ((noise_u == nullptr) != (noise_v == nullptr)) return fail(h, BIDI_EINVAL, "bidi_env_step: noise_u and noise_v go together");
  CU(cudaSetDevice(h->device));
  const bool draws = h->cfg.variant == BIDI_NEO_AGENT || h->qnet || h->csrl;
  if (noise_u || env_noise) CU(cudaStreamSynchronize(h->stream));  // the staging buffers may still be in flight
  if (noise_u && draws) {
    const size_t n = (size_t)h->A * h->n_noise;
    memcpy(h->h_noise, noise_u, sizeof(float) * n);
    memcpy(h->h_noise + n, noise_v, sizeof(float) * n);
    CU(cudaMemcpyAsync(h->d_noise_u, h->h_noise, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_noise_v, h->h_noise + n, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
  }
  if (env_noise) {
    CU(cudaMemcpyAsync(h->d_env_noise, env_noise, sizeof(float) * (size_t)h->A * BIDI_ENV_JOINTS, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));  // pageable source: the caller may reuse it on return
  }
  // the four clock inputs, RunnerMain.cpp:235-236, in the host's fp32 arithmetic like the reference's
  float ck[4];
  for (int a = 0; a < 4; a++) ck[a] = std::sin(h->env_steps / 60.0f * 2.0f * a * 2.0f * 3.141596f) * 0.5f + 0.5f;
  bidi::k_env_pre<<<grid_for(h->A, kBlock), kBlock, 0, h->stream>>>(h->P, h->d_env, make_float4(ck[0], ck[1], ck[2], ck[3]), BIDI_ENV_STATE, 20,
                                                                     BIDI_ENV_JOINTS, env_noise ? h->d_env_noise : nullptr, 0.05f,
                                                                     h->noise_seed ^ 0xE27ULL, h->step_index, h->first_agent, h->d_rewards);
  h->launches++;
  int rc = run_step(h, learn != 0, draws && noise_u == nullptr);
  if (rc) return rc;
  bidi::k_env_post<<<grid_for(h->A, kBlock), kBlock, 0, h->stream>>>(h->P, bidi::env_default_params(), h->d_env, 20);
  h->launches++;
  h->env_steps++;
  return BIDI_OK;
}

int bidi_env_get_state(bidi_engine* h, float* out) {
  if (!h || !out || !h->env_kind) return fail(h, BIDI_EINVAL, "bidi_env_get_state: no environment attached");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(out, h->d_env, sizeof(float) * (size_t)h->A * BIDI_ENV_FLOATS, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BIDI_OK;
}

int bidi_set_feedback(bidi_engine* h, int n, const int* src, const int* dst, float scale, float bias, float noise_std,
                      int reward_from_actions) {
  if (!h || n < 0 || (n > 0 && (!src || !dst))) return fail(h, BIDI_EINVAL, "bidi_set_feedback: bad argument");
  for (int k = 0; k < n; k++)
    if (src[k] < 0 || src[k] >= h->n_in || dst[k] < 0 || dst[k] >= h->n_in) return fail(h, BIDI_EINVAL, "bidi_set_feedback: index out of range");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  cudaFree(h->d_fb_src); cudaFree(h->d_fb_dst);
  h->d_fb_src = h->d_fb_dst = nullptr; h->fb_n = 0;
  if (n > 0) {
    CU(cudaMalloc(&h->d_fb_src, sizeof(int) * n));
    CU(cudaMalloc(&h->d_fb_dst, sizeof(int) * n));
    CU(cudaMemcpyAsync(h->d_fb_src, src, sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_fb_dst, dst, sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  h->fb_n = n; h->fb_scale = scale; h->fb_bias = bias; h->fb_std = noise_std; h->fb_reward = reward_from_actions;
  return BIDI_OK;
}

/* 1: always run the per-phase kernels, even where the fused coder kernel applies (tests compare both paths) */
int bidi_set_option(bidi_engine* h, const char* name, int value) {
  if (!h || !name) return fail(h, BIDI_EINVAL, "bidi_set_option: null argument");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  if (std::string(name) == "force_phase_kernels") h->force_phase_kernels = value != 0;
  else if (std::string(name) == "tile_kernels") {
    if (!value) std::fill(h->tiled.begin(), h->tiled.end(), 0);
    int rc = select_grid_kernels(h);
    if (rc) return rc;
  } else if (std::string(name) == "grid_kernels") {  // 0 off, 1 default (wide, untiled k-winners layers), 2 every k-winners layer
    h->grid_mode = value;
    int rc = select_grid_kernels(h);
    if (rc) return rc;
  }
  else if (std::string(name) == "strip_graph") h->strip_use_graph = value != 0;
  else if (std::string(name) == "tma_box") {
    h->use_tma_box = value != 0;
    int rc = select_grid_kernels(h);
    if (rc) return rc;
  }
  else if (std::string(name) == "persist_coder") {  // 0: the per-phase kernels for the large layers of a single iterative agent
    h->persist_mode = value != 0;
    int rc = select_persist(h);
    if (rc) return rc;
  }
  else if (std::string(name) == "overlap_learn") h->overlap_learn = value != 0;
  else if (std::string(name) == "dense_coder") {  // 0: use the general fused coder kernel where the dense one was chosen
    if (!value) {
      size_t smem = 0;
      for (int l = 0; l < h->L; l++) {
        const size_t need = sizeof(float) * (size_t)bidi::coder_cta_shape(h->layers[l]).floats;
        if (h->coder_cta[l] == 2) h->coder_cta[l] = need <= 200 * 1024 ? 1 : 0;
        if (h->coder_cta[l] == 1) smem = std::max(smem, need);
      }
      if (smem > 0) {
        CU(cudaFuncSetAttribute(bidi::k_coder_cta<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(bidi::k_coder_cta<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      }
    }
  }
  else return fail(h, BIDI_EINVAL, "bidi_set_option: unknown option");
  for (auto& g : h->graphs) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
  return apply_layout(h);  // the dense coder's layers are rows exactly while that kernel is their path
}

// ---- checkpoint / resume.  The reference has none for its hierarchies (only deep::FERL's text
// dump, deep/FERL.cpp:330-407).  File = header + configuration + the device arrays verbatim (the
// engine's own HBM layout), so saving and loading are plain copies.
namespace {
struct SnapHeader {
  char magic[8]; int version, config_bytes, desc_bytes, n_layers, n_agents, n_in; long long stride;
  unsigned long long step_index, noise_seed; long long first_agent;
  unsigned layout_rows;  // bit l: layer l's coder weights are stored as rows (bidi_engine::dense_rows)
};
constexpr int kSnapVersion = 3;
unsigned layout_mask(const bidi_engine* h) {
  unsigned m = 0;
  for (int l = 0; l < h->L; l++) if (h->dense_rows[l]) m |= 1u << l;
  return m;
}
const char kSnapMagic[8] = {'B', 'I', 'D', 'I', 'S', 'N', 'A', 'P'};
}

int bidi_save_snapshot(bidi_engine* h, const char* path) {
  if (!h || !path) return fail(h, BIDI_EINVAL, "bidi_save_snapshot: null argument");
  CU(cudaSetDevice(h->device));
  int rc = mirror_release(h);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));
  FILE* f = fopen(path, "wb");
  if (!f) return fail(h, BIDI_EINVAL, std::string("bidi_save_snapshot: cannot open ") + path);
  SnapHeader hd;
  memset(&hd, 0, sizeof(hd));
  memcpy(hd.magic, kSnapMagic, 8);
  hd.version = kSnapVersion; hd.config_bytes = (int)sizeof(bidi_config); hd.desc_bytes = (int)sizeof(bidi_layer_desc);
  hd.n_layers = h->L; hd.n_agents = h->A; hd.n_in = h->n_in; hd.stride = h->stride; hd.step_index = h->step_index;
  hd.noise_seed = h->noise_seed; hd.first_agent = h->first_agent; hd.layout_rows = layout_mask(h);
  bidi_config cfg = h->cfg;
  if (h->qnet) cfg.variant = BIDI_QPRSDR;
  if (h->csrl) cfg.variant = BIDI_CSRL;
  bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1 && fwrite(&cfg, sizeof(cfg), 1, f) == 1 &&
            fwrite(h->ld.data(), sizeof(bidi_layer_desc), (size_t)h->L, f) == (size_t)h->L;
  std::vector<float> buf;
  auto dump = [&](const float* dev, size_t n) {
    const size_t chunk = (size_t)16 << 20;
    for (size_t o = 0; o < n && ok; o += chunk) {
      const size_t m = std::min(chunk, n - o);
      buf.resize(m);
      if (cudaMemcpyAsync(buf.data(), dev + o, sizeof(float) * m, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
          cudaStreamSynchronize(h->stream) != cudaSuccess) { ok = false; break; }
      ok = fwrite(buf.data(), sizeof(float), m, f) == m;
    }
  };
  dump(h->d_in0, (size_t)h->A * h->n_in);
  dump(h->d_pred, (size_t)h->A * h->n_in);
  dump(h->d_blob, (size_t)h->A * (size_t)h->stride);
  ok = (fclose(f) == 0) && ok;
  cudaGetLastError();
  return ok ? BIDI_OK : fail(h, BIDI_EINVAL, std::string("bidi_save_snapshot: write to ") + path + " failed");
}

int bidi_load_snapshot(bidi_engine* h, const char* path) {
  if (!h || !path) return fail(h, BIDI_EINVAL, "bidi_load_snapshot: null argument");
  CU(cudaSetDevice(h->device));
  int rc = mirror_release(h);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));
  FILE* f = fopen(path, "rb");
  if (!f) return fail(h, BIDI_EINVAL, std::string("bidi_load_snapshot: cannot open ") + path);
  SnapHeader hd;
  bidi_config cfg;
  std::vector<bidi_layer_desc> ld((size_t)h->L);
  bidi_config mine = h->cfg;
  if (h->qnet) mine.variant = BIDI_QPRSDR;
  if (h->csrl) mine.variant = BIDI_CSRL;
  bool ok = fread(&hd, sizeof(hd), 1, f) == 1 && memcmp(hd.magic, kSnapMagic, 8) == 0 && hd.version == kSnapVersion &&
            hd.config_bytes == (int)sizeof(bidi_config) && hd.desc_bytes == (int)sizeof(bidi_layer_desc) && hd.n_layers == h->L &&
            hd.n_agents == h->A && hd.n_in == h->n_in && hd.stride == h->stride && hd.layout_rows == layout_mask(h) &&
            fread(&cfg, sizeof(cfg), 1, f) == 1 &&
            fread(ld.data(), sizeof(bidi_layer_desc), (size_t)h->L, f) == (size_t)h->L && memcmp(&cfg, &mine, sizeof(cfg)) == 0 &&
            memcmp(ld.data(), h->ld.data(), sizeof(bidi_layer_desc) * (size_t)h->L) == 0;
  if (!ok) { fclose(f); return fail(h, BIDI_EINVAL, std::string("bidi_load_snapshot: ") + path + " is not a snapshot of an engine with this configuration"); }
  std::vector<float> buf;
  auto fill = [&](float* dev, size_t n) {
    const size_t chunk = (size_t)16 << 20;
    for (size_t o = 0; o < n && ok; o += chunk) {
      const size_t m = std::min(chunk, n - o);
      buf.resize(m);
      ok = fread(buf.data(), sizeof(float), m, f) == m &&
           cudaMemcpyAsync(dev + o, buf.data(), sizeof(float) * m, cudaMemcpyHostToDevice, h->stream) == cudaSuccess &&
           cudaStreamSynchronize(h->stream) == cudaSuccess;  // on the engine's stream: ordered before the next step
    }
  };
  fill(h->d_in0, (size_t)h->A * h->n_in);
  fill(h->d_pred, (size_t)h->A * h->n_in);
  fill(h->d_blob, (size_t)h->A * (size_t)h->stride);
  fclose(f);
  cudaGetLastError();
  if (!ok) return fail(h, BIDI_EINVAL, std::string("bidi_load_snapshot: ") + path + " is truncated");
  h->step_index = hd.step_index; h->noise_seed = hd.noise_seed; h->first_agent = hd.first_agent;
  h->state_touched = true;
  if (h->cfg.variant == BIDI_NEO_AGENT) { h->col_inputs_stale = true; std::fill(h->col_inputs_set.begin(), h->col_inputs_set.end(), 0); }
  return BIDI_OK;
}

int bidi_profile_select(bidi_engine* h, const char* kernel_name) {
  if (!h) return fail(h, BIDI_EINVAL, "null engine");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  for (auto& g : h->graphs) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
  for (auto& e : h->profile_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  h->profile_events.clear();
  h->profile_name = kernel_name ? kernel_name : "";
  h->profile_ms = 0; h->profile_launches = 0;
  return BIDI_OK;
}

int bidi_profile_collect(bidi_engine* h, double* total_ms, long long* launches) {
  if (!h) return fail(h, BIDI_EINVAL, "null engine");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  for (auto& e : h->profile_events) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, e.first, e.second) == cudaSuccess) { h->profile_ms += ms; h->profile_launches++; }
    else cudaGetLastError(); // the pair belongs to the graph of the other learn flag: not recorded yet
  }
  if (total_ms) *total_ms = h->profile_ms;
  if (launches) *launches = h->profile_launches;
  return BIDI_OK;
}

int bidi_timer_start(bidi_engine* h) {
  if (!h) return fail(h, BIDI_EINVAL, "null engine");
  CU(cudaSetDevice(h->device));
  CU(cudaEventRecord(h->ev0, h->stream));
  return BIDI_OK;
}
int bidi_timer_stop_ms(bidi_engine* h, float* ms) {
  if (!h || !ms) return fail(h, BIDI_EINVAL, "null argument");
  CU(cudaSetDevice(h->device));
  CU(cudaEventRecord(h->ev1, h->stream));
  CU(cudaEventSynchronize(h->ev1));
  CU(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return BIDI_OK;
}

#ifdef BIDI_COL_TIMING
int bidi_debug_col_timing(unsigned long long* out16, int reset) {
  if (out16) cudaMemcpyFromSymbol(out16, bidi::g_col_timing, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(bidi::g_col_timing, z, sizeof(z)); }
  return 0;
}
#endif

/* sizeof the two POD structs, so a binding can check its mirror of the header */
int bidi_abi_sizes(int* config_bytes, int* layer_desc_bytes) {
  if (config_bytes) *config_bytes = (int)sizeof(bidi_config);
  if (layer_desc_bytes) *layer_desc_bytes = (int)sizeof(bidi_layer_desc);
  return BIDI_OK;
}

} // extern "C"

#include "bidi_strip.inl"
#include "bidi_sharded.inl"
