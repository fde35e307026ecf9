// bidi_strip.inl -- spatial split of ONE oversized layer over several GPUs (SURVEY 8e,
// BASELINE config 5).  Included at the end of bidi_engine.cu.
//
// A single-layer k-winners hierarchy (sdr::PredictiveRSDR, sdr/PredictiveRSDR.cpp:100-194)
// is cut into horizontal row strips, one engine ("part") per GPU.  The step becomes six
// compute phases with four halo exchanges of unit-state rows in between:
//
//   phase 0  activate            a = -theta + sum_ff + sum_rec          sdr/RSDR.cpp:123-133
//     X0     activations, r_lat rows either side
//   phase 1  k-winner rank       state                                  sdr/RSDR.cpp:135-145
//     X1     winners, rows the reconstruction / prediction / learning windows reach
//   phase 2  reconstruct         vis_recon (owned visible rows), recon  sdr/RSDR.cpp:165-194
//   phase 3  predict             delta rule + activation                sdr/PredictiveRSDR.cpp:122-160
//     X3     prediction activations (r_lat rows), attentions (halo units), and both reconstructions (made in
//            phase 2, first read by the learning rule in phase 5)
//   phase 4  k-winner rank       predicted state                        sdr/PredictiveRSDR.cpp:163-170
//     X4     predicted winners (rows the input prediction gathers from)
//   phase 5  learn, stepEnd, input prediction                           sdr/RSDR.cpp:241-266,196-212
//
// Weights never cross the link.  The reconstruction is a gather (bidi_kernels.cuh,
// gather_recon): an owned visible row sums over hidden units up to a window radius away,
// in ascending unit order -- the order the reference's scatter adds them, which a
// partial-sum exchange could not keep.  So a part also holds the weights of those few
// halo rows ("extended rows") and runs the winners-only learning rule on them
// redundantly, from exchanged inputs; identical inputs and arithmetic keep the copies
// bit-identical to the owner's.  Everything else is computed by the owner only.
//
// Transports: same process (events + cudaMemcpyPeerAsync, pull model) and NCCL
// send/recv (one process per GPU).  libnccl is opened at run time, so the library has
// no link-time NCCL dependency and single-GPU users never load it.
#include <dlfcn.h>

namespace {

struct StripGeom { int vh, hh, r_ff, r_rec, r_lat, r_pred; };

inline int clampi(int v, int lo, int hi) { return std::max(lo, std::min(hi, v)); }

// Row intervals of part `part` of `n` (include/bidinet_b200.h, enum bidi_strip_span).
// Pure geometry; window centres as the reference computes them (sdr/RSDR.cpp:33-41).
void strip_spans(const StripGeom& g, int part, int n, int spans[][2]) {
  auto set = [&](int k, int lo, int hi, int lim) { spans[k][0] = clampi(lo, 0, lim); spans[k][1] = clampi(hi, 0, lim); };
  const float ry = static_cast<float>(g.vh) / static_cast<float>(g.hh);
  auto cy = [&](int y) { return static_cast<int>(std::round(y * ry)); };
  const int o0 = (int)((long long)part * g.hh / n), o1 = (int)((long long)(part + 1) * g.hh / n);
  const int v0 = (int)((long long)part * g.vh / n), v1 = (int)((long long)(part + 1) * g.vh / n);
  set(BIDI_SPAN_OWN_HID, o0, o1, g.hh);
  set(BIDI_SPAN_OWN_VIS, v0, v1, g.vh);
  // hidden rows whose feed-forward window reaches an owned visible row
  int e0 = o0, e1 = o1;
  for (int y = 0; y < g.hh; y++) {
    const int c = cy(y);
    if (c + g.r_ff >= v0 && c - g.r_ff < v1) { e0 = std::min(e0, y); e1 = std::max(e1, y + 1); }
  }
  int f0 = g.hh, f1 = 0;
  for (int y = 0; y < g.hh; y++) {
    const int c = cy(y);
    if (c + g.r_ff >= v0 && c - g.r_ff < v1) { f0 = std::min(f0, y); f1 = std::max(f1, y + 1); }
  }
  if (f0 >= f1) { f0 = o0; f1 = o0; }
  set(BIDI_SPAN_PRED_STATE, f0, f1, g.hh);
  const int rr = std::max(0, g.r_rec);
  e0 = std::min(e0, o0 - rr); e1 = std::max(e1, o1 + rr);
  set(BIDI_SPAN_EXT_HID, e0, e1, g.hh);
  e0 = spans[BIDI_SPAN_EXT_HID][0]; e1 = spans[BIDI_SPAN_EXT_HID][1];
  set(BIDI_SPAN_STATE, std::min(e0 - rr, o0 - g.r_pred), std::max(e1 + rr, o1 + g.r_pred), g.hh);
  set(BIDI_SPAN_ACT, o0 - g.r_lat, o1 + g.r_lat, g.hh);
  set(BIDI_SPAN_VIS_RECON, cy(e0) - g.r_ff, cy(e1 - 1) + g.r_ff + 1, g.vh);
  set(BIDI_SPAN_HID_RECON, e0 - rr, e1 + rr, g.hh);
}

int strip_check(const bidi_config* cfg, const bidi_layer_desc* layers, int part, int n, StripGeom& g, std::string& why) {
  if (!cfg || !layers) { why = "null argument"; return BIDI_EINVAL; }
  if (cfg->variant != BIDI_KWINNERS || cfg->n_layers != 1) { why = "the spatial split handles a single-layer BIDI_KWINNERS hierarchy"; return BIDI_EINVAL; }
  g = StripGeom{cfg->in_height, layers[0].height, layers[0].r_ff, layers[0].r_rec, layers[0].r_lat, layers[0].r_pred};
  if (n < 1 || part < 0 || part >= n || n > g.hh || n > g.vh) { why = "bad part index or more parts than rows"; return BIDI_EINVAL; }
  return BIDI_OK;
}

// ---- NCCL, opened at run time
struct NcclUid { char internal[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
constexpr int kNcclFloat = 7;  // ncclFloat32 (nccl.h)

NcclApi* nccl_api(std::string& why) {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    // a process that already imported torch has its bundled libnccl.so.2 loaded: same soname, same handle
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      auto sym = [&](const char* n) { return dlsym(api.lib, n); };
      api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
      api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
      api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
      api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
      api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
      api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Send || !api.Recv || !api.GroupStart || !api.GroupEnd) {
        dlclose(api.lib);
        api.lib = nullptr;
      }
    }
  }
  if (!api.lib) { why = "libnccl.so.2 could not be opened"; return nullptr; }
  return &api;
}
#define NC(call)                                                                                       \
  do {                                                                                                 \
    int r_ = (call);                                                                                   \
    if (r_ != 0) return fail(h, BIDI_ECUDA, std::string(#call) + ": " + (nc->GetErrorString ? nc->GetErrorString(r_) : "NCCL error")); \
  } while (0)

void strip_release(bidi_engine* h) {
  if (h->nccl_comm) {
    std::string why;
    if (NcclApi* nc = nccl_api(why)) nc->CommDestroy(h->nccl_comm);
    h->nccl_comm = nullptr;
  }
  for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  h->ipc_opened.clear();
  h->ipc_linked = false;
  cudaFree(h->d_flags); cudaFree(h->d_epoch); cudaFree(h->d_peer_blob); cudaFree(h->d_peer_flags);
  cudaFree(h->d_sends); cudaFree(h->d_notify); cudaFree(h->d_await);
  h->d_flags = h->d_epoch = nullptr; h->d_peer_blob = nullptr; h->d_peer_flags = nullptr;
  h->d_sends = nullptr; h->d_notify = h->d_await = nullptr;
  for (auto& e : h->strip_ev) if (e) { cudaEventDestroy(e); e = nullptr; }
  if (h->strip_done) { cudaEventDestroy(h->strip_done); h->strip_done = nullptr; }
  for (bidi_engine* p : h->strip_peers)  // the other parts must not follow a dangling pointer
    if (p && p != h) p->strip_peers.clear();
  h->strip_peers.clear();
}

// ---- compute phases
template <class K, class... Args> void strip_launch(bidi_engine* h, K kernel, long long threads, Args... args) {
  if (threads <= 0) return;
  kernel<<<grid_for(threads, kBlock), kBlock, 0, h->stream>>>(h->P, args...);
  h->launches++;
}

void strip_phase(bidi_engine* h, int phase, bool learn) {
  LayerDev Y = h->layers[0];
  const long long A = h->A;
  auto span = [&](int k, int i) { return h->strip_span[k][i]; };
  auto rows = [&](int hid_span) {
    Y.row0 = span(hid_span, 0); Y.row1 = span(hid_span, 1);
    Y.vrow0 = span(BIDI_SPAN_OWN_VIS, 0); Y.vrow1 = span(BIDI_SPAN_OWN_VIS, 1);
  };
  rows(BIDI_SPAN_OWN_HID);
  const long long nh = A * (Y.row1 - Y.row0) * Y.hw, nv = A * (Y.vrow1 - Y.vrow0) * Y.vw;
  const auto& G = h->gridk[0];
  const long long th = A * bidi::grid_tiles(Y.hw, Y.row1 - Y.row0), tv = A * bidi::grid_tiles(Y.vw, Y.vrow1 - Y.vrow0);
  auto grid = [&](auto kernel, long long blocks, int smem_floats, auto... args) {
    if (blocks <= 0) return;
    kernel<<<(unsigned)blocks, GRID_TX * GRID_TY, sizeof(float) * (size_t)smem_floats, h->stream>>>(h->P, args...);
    h->launches++;
  };
  switch (phase) {
    case 0:
      if (G.on) { launch_kw_activate(h, h->stream, G, Y, 0); h->launches++; }
      else strip_launch(h, bidi::k_kw_activate, nh, Y, 0);
      break;
    case 1:
      if (G.on) grid(bidi::k_g_kw_inhibit, th, G.lat, Y, 0, Y.act, Y.state, -1LL);
      else strip_launch(h, bidi::k_kw_inhibit, nh, Y, 0, Y.act, Y.state, -1LL);
      break;
    case 2:
      if (G.on) grid(bidi::k_g_reconstruct, tv + th, 3 * G.gather, Y, 0, Y.state, 1, 0, G.gather, 0, 0.0f, 0);
      else strip_launch(h, bidi::k_reconstruct, nv + nh, Y, 0, Y.state, 0, 0.0f, 0);
      break;
    case 3:
      if (G.on) grid(bidi::k_g_kw_predict, th, 4 * (G.fb + G.pred), Y, 0, (int)learn, Y.p_state, Y.p_state_prev, G.pred, G.fb);
      else strip_launch(h, bidi::k_predict, nh, Y, 0, (int)learn, Y.p_state, Y.p_state_prev);
      break;
    case 4:
      if (G.on) grid(bidi::k_g_kw_inhibit, th, G.lat, Y, 0, Y.p_act, Y.p_state, -1LL);
      else strip_launch(h, bidi::k_kw_inhibit, nh, Y, 0, Y.p_act, Y.p_state, -1LL);
      break;
    case 5: {
      if (learn) {
        rows(BIDI_SPAN_EXT_HID);  // owned rows + the halo rows whose weights this part mirrors
        if (G.on) grid(bidi::k_g_kw_learn, A * bidi::grid_tiles(Y.hw, Y.row1 - Y.row0), 0, Y, 0);
        else strip_launch(h, bidi::k_kw_learn, A * (Y.row1 - Y.row0) * Y.hw, Y, 0);
        rows(BIDI_SPAN_OWN_HID);
      }
      const int s0 = span(BIDI_SPAN_STATE, 0), s1 = span(BIDI_SPAN_STATE, 1);
      strip_launch(h, bidi::k_strip_step_end, A * (s1 - s0) * Y.hw, Y, s0, s1, h->ipc_linked ? h->d_epoch : nullptr);
      if (G.on) grid(bidi::k_g_reconstruct, tv, 3 * G.gather, Y, 0, Y.p_state, 0, 1, G.gather, 0, 0.0f, 0);
      else strip_launch(h, bidi::k_kw_predict_input, nv, Y);
      break;
    }
  }
}
constexpr int kStripPhases = 6;
inline bool strip_point_used(int point, bool) { return point != 2; }  // X2's rows (the reconstructions) travel with X3

// ---- the exchange plan: which rows of which plane go to / come from which part
void strip_build_plan(bidi_engine* h) {
  const LayerDev& Y = h->layers[0];
  const bidi_layer_desc& d = h->ld[0];
  const StripGeom g{h->cfg.in_height, d.height, d.r_ff, d.r_rec, d.r_lat, d.r_pred};
  const int n = h->strip_n, me = h->strip_part;
  std::vector<std::array<std::array<int, 2>, BIDI_SPAN_COUNT>> sp(n);
  for (int p = 0; p < n; p++) {
    int s[BIDI_SPAN_COUNT][2];
    strip_spans(g, p, n, s);
    for (int k = 0; k < BIDI_SPAN_COUNT; k++) sp[p][k] = {s[k][0], s[k][1]};
  }
  for (int k = 0; k < BIDI_SPAN_COUNT; k++) { h->strip_span[k][0] = sp[me][k][0]; h->strip_span[k][1] = sp[me][k][1]; }
  struct Item { int point; long long off; int width, need, own; };
  std::vector<Item> items = {
      {0, Y.act, Y.hw, BIDI_SPAN_ACT, BIDI_SPAN_OWN_HID},
      {1, Y.state, Y.hw, BIDI_SPAN_STATE, BIDI_SPAN_OWN_HID},
      {3, Y.vis_recon, Y.vw, BIDI_SPAN_VIS_RECON, BIDI_SPAN_OWN_VIS},  // made in phase 2, first read in phase 5: rides with X3
      {3, Y.p_act, Y.hw, BIDI_SPAN_ACT, BIDI_SPAN_OWN_HID},
      {3, Y.p_mod, Y.hw, BIDI_SPAN_EXT_HID, BIDI_SPAN_OWN_HID},
      {4, Y.p_state, Y.hw, BIDI_SPAN_PRED_STATE, BIDI_SPAN_OWN_HID},
  };
  if (d.r_rec >= 0) items.insert(items.begin() + 3, Item{3, Y.recon, Y.hw, BIDI_SPAN_HID_RECON, BIDI_SPAN_OWN_HID});
  h->strip_send.clear(); h->strip_recv.clear();
  for (const Item& it : items)
    for (int q = 0; q < n; q++) {
      if (q == me) continue;
      // what I need that q owns
      int lo = std::max(sp[me][it.need][0], sp[q][it.own][0]), hi = std::min(sp[me][it.need][1], sp[q][it.own][1]);
      if (lo < hi) h->strip_recv.push_back({it.point, q, it.off + (long long)lo * it.width, (long long)(hi - lo) * it.width});
      // what q needs that I own
      lo = std::max(sp[q][it.need][0], sp[me][it.own][0]); hi = std::min(sp[q][it.need][1], sp[me][it.own][1]);
      if (lo < hi) h->strip_send.push_back({it.point, q, it.off + (long long)lo * it.width, (long long)(hi - lo) * it.width});
    }
}

// same-process transport: pull the rows from the owner's blob once its phase has run
int strip_exchange_local(bidi_engine* h, int point) {
  for (const auto& x : h->strip_recv) {
    if (x.point != point) continue;
    bidi_engine* q = h->strip_peers[x.peer];
    CU(cudaStreamWaitEvent(h->stream, q->strip_ev[point], 0));
    for (int a = 0; a < h->A; a++)
      CU(cudaMemcpyPeerAsync(h->d_blob + (long long)a * h->stride + x.off, h->device, q->d_blob + (long long)a * q->stride + x.off,
                             q->device, sizeof(float) * x.count, h->stream));
  }
  return BIDI_OK;
}

int strip_exchange_nccl(bidi_engine* h, int point) {
  std::string why;
  NcclApi* nc = nccl_api(why);
  if (!nc) return fail(h, BIDI_EINVAL, why);
  bool any = false;
  for (const auto& x : h->strip_send) any |= x.point == point;
  for (const auto& x : h->strip_recv) any |= x.point == point;
  if (!any) return BIDI_OK;
  NC(nc->GroupStart());
  for (const auto& x : h->strip_send)
    if (x.point == point)
      for (int a = 0; a < h->A; a++) NC(nc->Send(h->d_blob + (long long)a * h->stride + x.off, (size_t)x.count, kNcclFloat, x.peer, h->nccl_comm, h->stream));
  for (const auto& x : h->strip_recv)
    if (x.point == point)
      for (int a = 0; a < h->A; a++) NC(nc->Recv(h->d_blob + (long long)a * h->stride + x.off, (size_t)x.count, kNcclFloat, x.peer, h->nccl_comm, h->stream));
  NC(nc->GroupEnd());
  return BIDI_OK;
}

void strip_exchange_push(bidi_engine* h, int point) {
  const auto& a = h->push[point];
  if (a.n_sends == 0 && a.n_await == 0) return;
  bidi::k_strip_push<<<1, 1024, 0, h->stream>>>(h->P, h->d_sends + a.send0, a.n_sends, h->d_peer_blob, h->d_peer_flags, h->d_flags, h->d_epoch, point,
                                                h->strip_part, h->strip_n, h->d_notify + a.notify0, a.n_notify, h->d_await + a.await0, a.n_await);
  h->launches++;
}

int strip_record_nccl(bidi_engine* h, bool learn) {
  // inside a stream capture the reconstructions (phase 2, first needed by exchange X3) become a second branch of the
  // graph, beside the prediction phase
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(h->stream, &cap);
  const bool fork = cap == cudaStreamCaptureStatusActive && h->side != nullptr && h->overlap_learn;
  for (int phase = 0; phase < kStripPhases; phase++) {
    if (fork && phase == 2) {
      CU(cudaEventRecord(h->ev_fork, h->stream));
      CU(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
      cudaStream_t main_stream = h->stream;
      h->stream = h->side;
      strip_phase(h, 2, learn);
      h->stream = main_stream;
      CU(cudaEventRecord(h->ev_join, h->side));
      continue;
    }
    strip_phase(h, phase, learn);
    if (fork && phase == 3) CU(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    if (phase < kStripPhases - 1 && strip_point_used(phase, learn)) {
      if (h->ipc_linked) strip_exchange_push(h, phase);
      else {
        int rc = strip_exchange_nccl(h, phase);
        if (rc) return rc;
      }
    }
  }
  return BIDI_OK;
}

// The whole step -- six compute phases and five grouped send/recv exchanges -- is recorded
// once into a CUDA graph per value of `learn` (NCCL operations can be captured), so a step
// costs one graph launch instead of ~15 separate submissions.
int strip_step_nccl(bidi_engine* h, bool learn) {
  if (!h->nccl_comm && !h->ipc_linked)
    return fail(h, BIDI_EINVAL, "this engine is one part of a split layer: link it (bidi_strip_link_ipc / bidi_strip_link_nccl) or step all parts with bidi_strip_step_local");
  if (h->strip_use_graph) {
    cudaGraphExec_t& exec = h->graphs[learn];
    if (!exec) {
      const long long before = h->launches;
      cudaGraph_t graph = nullptr;
      CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
      int rc = strip_record_nccl(h, learn);
      cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
      h->graph_kernels[learn] = (int)(h->launches - before);
      h->launches = before;
      if (rc == BIDI_OK && e == cudaSuccess && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
        cudaGraphDestroy(graph);
      } else {  // capture not possible with this NCCL: submit the launches one by one from now on
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        exec = nullptr;
        h->strip_use_graph = false;
      }
    }
    if (exec) {
      CU(cudaGraphLaunch(exec, h->stream));
      h->launches += h->graph_kernels[learn];
      h->step_index++;
      return BIDI_OK;
    }
  }
  int rc = strip_record_nccl(h, learn);
  if (rc) return rc;
  CU(cudaGetLastError());
  h->step_index++;
  return BIDI_OK;
}

int strip_step_local_all(bidi_engine** parts, int n, bool learn, int frame) {
  bidi_engine* h = parts && n > 0 ? parts[0] : nullptr;
  if (!h || (int)h->strip_peers.size() != n || h->strip_n != n) return fail(h, BIDI_EINVAL, "bidi_strip_step_local: the parts are not linked");
  for (int p = 0; p < n; p++)
    if (!parts[p] || parts[p] != h->strip_peers[p]) return fail(h, BIDI_EINVAL, "bidi_strip_step_local: pass the parts in the order they were linked");
  for (int p = 0; p < n; p++) {
    bidi_engine* e = parts[p];
    CU(cudaSetDevice(e->device));
    int rc = mirror_release(e);
    if (rc) return rc;
    if (frame >= 0) {
      if (e->n_frames < 1) return fail(h, BIDI_EINVAL, "bidi_strip_step_frame_local: no frames loaded");
      const size_t nn = (size_t)e->A * e->n_in;
      CU(cudaMemcpyAsync(e->d_in0, e->d_frames + (size_t)(frame % e->n_frames) * nn, sizeof(float) * nn, cudaMemcpyDeviceToDevice, e->stream));
    }
    // nobody overwrites a plane a neighbour may still be pulling from the previous step
    for (int q = 0; q < n; q++)
      if (q != p) CU(cudaStreamWaitEvent(e->stream, parts[q]->strip_done, 0));
  }
  for (int phase = 0; phase < kStripPhases; phase++) {
    const bool xch = phase < kStripPhases - 1 && strip_point_used(phase, learn);
    for (int p = 0; p < n; p++) {
      bidi_engine* e = parts[p];
      CU(cudaSetDevice(e->device));
      strip_phase(e, phase, learn);
      if (xch) CU(cudaEventRecord(e->strip_ev[phase], e->stream));
    }
    if (xch)
      for (int p = 0; p < n; p++) {
        CU(cudaSetDevice(parts[p]->device));
        int rc = strip_exchange_local(parts[p], phase);
        if (rc) return fail(h, rc, parts[p]->err);
      }
  }
  for (int p = 0; p < n; p++) {
    bidi_engine* e = parts[p];
    CU(cudaSetDevice(e->device));
    CU(cudaGetLastError());
    CU(cudaEventRecord(e->strip_done, e->stream));
    e->step_index++;
  }
  return BIDI_OK;
}

} // namespace

extern "C" {

int bidi_strip_plan(const bidi_config* cfg, const bidi_layer_desc* layers, int part, int n_parts, int* spans) {
  StripGeom g;
  std::string why;
  int rc = strip_check(cfg, layers, part, n_parts, g, why);
  if (rc || !spans) return fail(nullptr, BIDI_EINVAL, "bidi_strip_plan: " + (spans ? why : std::string("null argument")));
  strip_spans(g, part, n_parts, reinterpret_cast<int(*)[2]>(spans));
  return BIDI_OK;
}

int bidi_strip_configure(bidi_engine* h, int part, int n_parts) {
  if (!h) return fail(h, BIDI_EINVAL, "null engine");
  StripGeom g;
  std::string why;
  int rc = strip_check(&h->cfg, h->ld.data(), part, n_parts, g, why);
  if (rc) return fail(h, rc, "bidi_strip_configure: " + why);
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  strip_release(h);
  for (auto& gr : h->graphs) if (gr) { cudaGraphExecDestroy(gr); gr = nullptr; }
  h->strip_part = part; h->strip_n = n_parts;
  if (n_parts == 1) return BIDI_OK;
  std::fill(h->tiled.begin(), h->tiled.end(), 0);
  {
    int rc2 = select_grid_kernels(h);
    if (rc2) return rc2;
    rc2 = apply_layout(h);  // the chosen kernel family decides the layer's layout
    if (rc2) return rc2;
  }
  strip_build_plan(h);
  for (auto& e : h->strip_ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&h->strip_done, cudaEventDisableTiming));
  CU(cudaEventRecord(h->strip_done, h->stream));
  return BIDI_OK;
}

int bidi_strip_link_local(bidi_engine** parts, int n_parts) {
  bidi_engine* h = parts && n_parts > 0 ? parts[0] : nullptr;
  if (!h || n_parts < 2) return fail(h, BIDI_EINVAL, "bidi_strip_link_local: need at least two parts");
  for (int p = 0; p < n_parts; p++) {
    bidi_engine* e = parts[p];
    if (!e || e->strip_n != n_parts || e->strip_part != p || e->stride != h->stride || e->A != h->A)
      return fail(h, BIDI_EINVAL, "bidi_strip_link_local: part p must be configured as (p, n_parts) with the same shape");
  }
  for (int p = 0; p < n_parts; p++) {
    bidi_engine* e = parts[p];
    e->strip_peers.assign(parts, parts + n_parts);
    CU(cudaSetDevice(e->device));
    for (int q = 0; q < n_parts; q++)
      if (parts[q]->device != e->device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, e->device, parts[q]->device));
        if (can) {
          cudaError_t er = cudaDeviceEnablePeerAccess(parts[q]->device, 0);
          if (er != cudaSuccess && er != cudaErrorPeerAccessAlreadyEnabled) return fail(h, BIDI_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(er));
          cudaGetLastError();
        }  // without peer access cudaMemcpyPeerAsync stages through the host: slower, still correct
      }
  }
  return BIDI_OK;
}

int bidi_strip_step_local(bidi_engine** parts, int n_parts, int learn) { return strip_step_local_all(parts, n_parts, learn != 0, -1); }
int bidi_strip_step_frame_local(bidi_engine** parts, int n_parts, int t, int learn) {
  if (t < 0) return fail(parts && n_parts > 0 ? parts[0] : nullptr, BIDI_EINVAL, "bidi_strip_step_frame_local: negative frame index");
  return strip_step_local_all(parts, n_parts, learn != 0, t);
}

int bidi_strip_nccl_unique_id(void* id128) {
  bidi_engine* h = nullptr;
  std::string why;
  NcclApi* nc = nccl_api(why);
  if (!nc || !id128) return fail(h, BIDI_EINVAL, "bidi_strip_nccl_unique_id: " + (nc ? std::string("null argument") : why));
  NcclUid id;
  NC(nc->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return BIDI_OK;
}

int bidi_strip_link_nccl(bidi_engine* h, const void* id128) {
  if (!h || !id128 || h->strip_n < 2) return fail(h, BIDI_EINVAL, "bidi_strip_link_nccl: configure the part first (bidi_strip_configure, n_parts >= 2)");
  std::string why;
  NcclApi* nc = nccl_api(why);
  if (!nc) return fail(h, BIDI_EINVAL, "bidi_strip_link_nccl: " + why);
  CU(cudaSetDevice(h->device));
  NcclUid id;
  memcpy(&id, id128, sizeof(id));
  NC(nc->CommInitRank(&h->nccl_comm, h->strip_n, id, h->strip_part));
  return BIDI_OK;
}

// Peer-memory transport.  Export: the CUDA IPC handles of this part's blob and flag words.
int bidi_strip_ipc_export(bidi_engine* h, void* handle128) {
  if (!h || !handle128 || h->strip_n < 2) return fail(h, BIDI_EINVAL, "bidi_strip_ipc_export: configure the part first (bidi_strip_configure, n_parts >= 2)");
  static_assert(2 * sizeof(cudaIpcMemHandle_t) <= 128, "two IPC handles fit the 128-byte record");
  CU(cudaSetDevice(h->device));
  {  // everything uploaded so far is in place before any neighbour can learn this memory's address
    int rc = mirror_release(h);
    if (rc) return rc;
    CU(cudaStreamSynchronize(h->stream));
  }
  if (!h->d_flags) {
    CU(cudaMalloc(&h->d_flags, sizeof(unsigned) * 5 * h->strip_n));
    CU(cudaMemset(h->d_flags, 0, sizeof(unsigned) * 5 * h->strip_n));
    CU(cudaMalloc(&h->d_epoch, sizeof(unsigned)));
    CU(cudaMemset(h->d_epoch, 0, sizeof(unsigned)));
  }
  cudaIpcMemHandle_t hb, hf;
  CU(cudaIpcGetMemHandle(&hb, h->d_blob));
  CU(cudaIpcGetMemHandle(&hf, h->d_flags));
  memset(handle128, 0, 128);
  memcpy(handle128, &hb, sizeof(hb));
  memcpy(static_cast<char*>(handle128) + 64, &hf, sizeof(hf));
  return BIDI_OK;
}

// Link: handles[n_parts][128] as exported by every part (this part's own record is ignored).
int bidi_strip_link_ipc(bidi_engine* h, const void* handles) {
  if (!h || !handles || h->strip_n < 2 || !h->d_flags) return fail(h, BIDI_EINVAL, "bidi_strip_link_ipc: export this part's handles first (bidi_strip_ipc_export)");
  CU(cudaSetDevice(h->device));
  const int n = h->strip_n, me = h->strip_part;
  std::vector<float*> blobs(n, nullptr);
  std::vector<unsigned*> flags(n, nullptr);
  blobs[me] = h->d_blob; flags[me] = h->d_flags;
  std::vector<char> need(n, 0);
  for (const auto& x : h->strip_send) need[x.peer] = 1;
  for (const auto& x : h->strip_recv) need[x.peer] = 1;
  for (int q = 0; q < n; q++) {
    if (q == me || !need[q]) continue;
    cudaIpcMemHandle_t hb, hf;
    memcpy(&hb, static_cast<const char*>(handles) + 128 * q, sizeof(hb));
    memcpy(&hf, static_cast<const char*>(handles) + 128 * q + 64, sizeof(hf));
    void *pb = nullptr, *pf = nullptr;
    CU(cudaIpcOpenMemHandle(&pb, hb, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_opened.push_back(pb);
    CU(cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_opened.push_back(pf);
    blobs[q] = static_cast<float*>(pb); flags[q] = static_cast<unsigned*>(pf);
  }
  // per exchange point: the rows to push, whom to notify, whom to wait for
  std::vector<bidi::StripXferDev> sends;
  std::vector<int> notify, await;
  for (int point = 0; point < 5; point++) {
    auto& a = h->push[point];
    a.send0 = (int)sends.size(); a.notify0 = (int)notify.size(); a.await0 = (int)await.size();
    std::vector<char> to(n, 0), from(n, 0);
    for (const auto& x : h->strip_send) if (x.point == point) { sends.push_back({x.peer, 0, x.off, x.count}); to[x.peer] = 1; }
    for (const auto& x : h->strip_recv) if (x.point == point) from[x.peer] = 1;
    for (int q = 0; q < n; q++) { if (to[q]) notify.push_back(q); if (from[q]) await.push_back(q); }
    a.n_sends = (int)sends.size() - a.send0; a.n_notify = (int)notify.size() - a.notify0; a.n_await = (int)await.size() - a.await0;
    if (a.n_notify > 1024 || a.n_await > 1024) return fail(h, BIDI_EINVAL, "bidi_strip_link_ipc: too many neighbours");
  }
  auto upload = [&](auto** dst, const auto& v) -> cudaError_t {
    using T = typename std::remove_reference<decltype(v)>::type::value_type;
    cudaError_t e = cudaMalloc(dst, sizeof(T) * std::max<size_t>(1, v.size()));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, h->stream);  // ordered before the next step
    return e != cudaSuccess ? e : cudaStreamSynchronize(h->stream);
  };
  CU(upload(&h->d_peer_blob, blobs));
  CU(upload(&h->d_peer_flags, flags));
  CU(upload(&h->d_sends, sends));
  CU(upload(&h->d_notify, notify));
  CU(upload(&h->d_await, await));
  for (auto& g : h->graphs) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
  h->ipc_linked = true;
  return BIDI_OK;
}

long long bidi_strip_halo_floats(const bidi_engine* h) {
  if (!h) return -1;
  long long n = 0;
  for (const auto& x : h->strip_send) n += x.count * h->A;
  return n;
}

} // extern "C"
