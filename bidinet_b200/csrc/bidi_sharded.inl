// bidi_sharded.inl -- many independent hierarchies of one shape spread over several devices of ONE process
// (include/bidinet_b200.h, bidi_sharded_*).  In the reference a second agent is simply a second object
// (RunnerMain.cpp:100-106); agents never interact (no cross-agent term anywhere in neo/Agent.cpp), so the
// population is cut into contiguous blocks, one bidi_engine per device, and every call fans out: each shard
// enqueues on its own device's stream and the devices run concurrently.  There is no data-path collective.
struct bidi_sharded {
  std::vector<bidi_engine*> parts;
  std::vector<int> first, count;   // global agent range of every shard
  int A = 0, n_in = 0, n_noise = 0, n_col = 0;
  std::vector<float> slice;        // staging for calls whose per-shard slices are not contiguous in the caller's array
  std::string err;
};

namespace {
int sfail(bidi_sharded* s, int code, const std::string& msg) {
  g_last_error = msg;
  if (s) s->err = msg;
  return code;
}
// run f(shard index, engine) on every shard; the first failure wins
template <class F> int for_shards(bidi_sharded* s, const char* what, F&& f) {
  if (!s) return sfail(s, BIDI_EINVAL, std::string(what) + ": null handle");
  for (size_t k = 0; k < s->parts.size(); k++) {
    const int rc = f((int)k, s->parts[k]);
    if (rc != BIDI_OK) return sfail(s, rc, std::string(what) + ", shard " + std::to_string(k) + ": " + bidi_last_error(s->parts[k]));
  }
  return BIDI_OK;
}
}  // namespace

extern "C" {

const char* bidi_sharded_last_error(const bidi_sharded* s) { return s ? s->err.c_str() : g_last_error.c_str(); }

void bidi_sharded_destroy(bidi_sharded* s) {
  if (!s) return;
  for (bidi_engine* e : s->parts) bidi_destroy(e);
  delete s;
}

int bidi_create_sharded(const bidi_config* cfg, const bidi_layer_desc* layers, int n_agents, const int* devices, int n_devices,
                        bidi_sharded** out) {
  if (!cfg || !layers || !devices || !out || n_devices < 1 || n_agents < n_devices)
    return sfail(nullptr, BIDI_EINVAL, "bidi_create_sharded: bad argument (at least one agent per shard)");
  *out = nullptr;
  bidi_sharded* s = new bidi_sharded();
  s->A = n_agents;
  for (int k = 0; k < n_devices; k++) {
    const int lo = (int)((long long)n_agents * k / n_devices), hi = (int)((long long)n_agents * (k + 1) / n_devices);
    bidi_engine* e = nullptr;
    const int rc = bidi_create(cfg, layers, hi - lo, devices[k], &e);
    if (rc != BIDI_OK) {
      const std::string msg = std::string("bidi_create_sharded, shard ") + std::to_string(k) + ": " + bidi_last_error(nullptr);
      bidi_sharded_destroy(s);
      return sfail(nullptr, rc, msg);
    }
    bidi_set_noise_stream(e, e->noise_seed, lo);  // every global agent explores on its own stream of the engines' generator
    s->parts.push_back(e); s->first.push_back(lo); s->count.push_back(hi - lo);
  }
  s->n_in = s->parts[0]->n_in; s->n_noise = s->parts[0]->n_noise; s->n_col = s->parts[0]->ncol;
  *out = s;
  return BIDI_OK;
}

int bidi_sharded_num_shards(const bidi_sharded* s) { return s ? (int)s->parts.size() : 0; }
int bidi_sharded_num_agents(const bidi_sharded* s) { return s ? s->A : 0; }

bidi_engine* bidi_sharded_shard(bidi_sharded* s, int k, int* first_agent, int* n_agents) {
  if (!s || k < 0 || k >= (int)s->parts.size()) return nullptr;
  if (first_agent) *first_agent = s->first[k];
  if (n_agents) *n_agents = s->count[k];
  return s->parts[k];
}

int bidi_sharded_init_random(bidi_sharded* s, unsigned seed_base) {
  return for_shards(s, "bidi_sharded_init_random", [&](int k, bidi_engine* e) { return bidi_init_random(e, 0, s->count[k], seed_base + (unsigned)s->first[k]); });
}

int bidi_sharded_set_inputs(bidi_sharded* s, const float* in) {
  if (!s || !in) return sfail(s, BIDI_EINVAL, "bidi_sharded_set_inputs: null argument");
  return for_shards(s, "bidi_sharded_set_inputs", [&](int k, bidi_engine* e) { return bidi_set_inputs(e, in + (size_t)s->first[k] * s->n_in); });
}

int bidi_sharded_step(bidi_sharded* s, const float* rewards, const float* noise_u, const float* noise_v, int learn) {
  return for_shards(s, "bidi_sharded_step", [&](int k, bidi_engine* e) {
    const size_t a0 = (size_t)s->first[k];
    return bidi_step(e, rewards ? rewards + a0 : nullptr, noise_u ? noise_u + a0 * s->n_noise : nullptr,
                     noise_v ? noise_v + a0 * s->n_noise : nullptr, learn);
  });
}

int bidi_sharded_get_predictions(bidi_sharded* s, float* out) {
  if (!s || !out) return sfail(s, BIDI_EINVAL, "bidi_sharded_get_predictions: null argument");
  return for_shards(s, "bidi_sharded_get_predictions", [&](int k, bidi_engine* e) { return bidi_get_predictions(e, out + (size_t)s->first[k] * s->n_in); });
}

int bidi_sharded_get_column_actions(bidi_sharded* s, float* out) {
  if (!s || !out) return sfail(s, BIDI_EINVAL, "bidi_sharded_get_column_actions: null argument");
  return for_shards(s, "bidi_sharded_get_column_actions", [&](int k, bidi_engine* e) { return bidi_get_column_actions(e, out + (size_t)s->first[k] * s->n_col * 3); });
}

int bidi_sharded_load_frames(bidi_sharded* s, const float* frames, const float* rewards, int n_frames) {
  if (!s || !frames || n_frames < 1) return sfail(s, BIDI_EINVAL, "bidi_sharded_load_frames: bad argument");
  return for_shards(s, "bidi_sharded_load_frames", [&](int k, bidi_engine* e) {
    // the caller's arrays are [frame][global agent][...]: gather this shard's agents of every frame
    const size_t c = (size_t)s->count[k], a0 = (size_t)s->first[k], row = (size_t)s->n_in;
    s->slice.resize((size_t)n_frames * c * row + (rewards ? (size_t)n_frames * c : 0));
    float* fr = s->slice.data();
    float* rw = rewards ? fr + (size_t)n_frames * c * row : nullptr;
    for (int t = 0; t < n_frames; t++) {
      memcpy(fr + (size_t)t * c * row, frames + ((size_t)t * s->A + a0) * row, sizeof(float) * c * row);
      if (rw) memcpy(rw + (size_t)t * c, rewards + (size_t)t * s->A + a0, sizeof(float) * c);
    }
    return bidi_load_frames(e, fr, rw, n_frames);
  });
}

int bidi_sharded_step_frame(bidi_sharded* s, int t, int learn) {
  return for_shards(s, "bidi_sharded_step_frame", [&](int, bidi_engine* e) { return bidi_step_frame(e, t, learn); });
}

int bidi_sharded_set_feedback(bidi_sharded* s, int n, const int* src, const int* dst, float scale, float bias, float noise_std,
                              int reward_from_actions) {
  return for_shards(s, "bidi_sharded_set_feedback", [&](int, bidi_engine* e) { return bidi_set_feedback(e, n, src, dst, scale, bias, noise_std, reward_from_actions); });
}

int bidi_sharded_sync(bidi_sharded* s) {
  return for_shards(s, "bidi_sharded_sync", [&](int, bidi_engine* e) { return bidi_sync(e); });
}

int bidi_sharded_timer_start(bidi_sharded* s) {
  return for_shards(s, "bidi_sharded_timer_start", [&](int, bidi_engine* e) { return bidi_timer_start(e); });
}

/* the slowest shard's time between timer_start and now, on its own device's stream (CUDA events) */
int bidi_sharded_timer_stop_ms(bidi_sharded* s, float* ms) {
  if (!s || !ms) return sfail(s, BIDI_EINVAL, "bidi_sharded_timer_stop_ms: null argument");
  *ms = 0.0f;
  // record every shard's stop event first, then wait: the shards keep running concurrently
  int rc = for_shards(s, "bidi_sharded_timer_stop_ms", [&](int, bidi_engine* e) {
    cudaSetDevice(e->device);
    return cudaEventRecord(e->ev1, e->stream) == cudaSuccess ? BIDI_OK : BIDI_ECUDA;
  });
  if (rc) return rc;
  return for_shards(s, "bidi_sharded_timer_stop_ms", [&](int, bidi_engine* e) {
    cudaSetDevice(e->device);
    float t = 0.0f;
    if (cudaEventSynchronize(e->ev1) != cudaSuccess || cudaEventElapsedTime(&t, e->ev0, e->ev1) != cudaSuccess) return (int)BIDI_ECUDA;
    *ms = std::max(*ms, t);
    return (int)BIDI_OK;
  });
}

}  // extern "C"
