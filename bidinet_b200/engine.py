"""ctypes binding of the C ABI (include/bidinet_b200.h).

``Engine`` is the Python-side mirror of what the reference's experiment mains do
with a hierarchy object: create it from a layer-descriptor list, then per step
``set_inputs`` / ``step`` / ``get_predictions`` (RunnerMain.cpp:139-154,239-249,
Experiment_Prediction.cpp:124-148).  All compute happens in the CUDA library; if it
is missing or no GPU is usable this module raises -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from .build import LIB_PATH
from .config import ARRAY, ARRAY_NAMES, Config, LayerDesc

_FP = C.POINTER(C.c_float)
_IP = C.POINTER(C.c_int)

# every symbol include/bidinet_b200.h declares: (restype, argtypes)
ABI = {
    "bidi_last_error": (C.c_char_p, [C.c_void_p]),
    "bidi_create": (C.c_int, [C.POINTER(Config), C.POINTER(LayerDesc), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "bidi_destroy": (None, [C.c_void_p]),
    "bidi_init_random": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint]),
    "bidi_init_random_generator": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int]),
    "bidi_num_agents": (C.c_int, [C.c_void_p]),
    "bidi_num_inputs": (C.c_int, [C.c_void_p]),
    "bidi_num_columns": (C.c_int, [C.c_void_p]),
    "bidi_num_noise": (C.c_int, [C.c_void_p]),
    "bidi_num_generate_noise": (C.c_int, [C.c_void_p]),
    "bidi_step_generate": (C.c_int, [C.c_void_p, _FP, C.c_float]),
    "bidi_get_actions": (C.c_int, [C.c_void_p, _FP]),
    "bidi_q_kept": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_ubyte), C.c_longlong]),
    "bidi_array_len": (C.c_longlong, [C.c_void_p, C.c_int, C.c_int]),
    "bidi_conn_counts": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _IP, C.c_int]),
    "bidi_set_array": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _FP, C.c_longlong]),
    "bidi_get_array": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _FP, C.c_longlong]),
    "bidi_set_inputs": (C.c_int, [C.c_void_p, _FP]),
    "bidi_get_inputs": (C.c_int, [C.c_void_p, _FP]),
    "bidi_host_inputs": (_FP, [C.c_void_p]),
    "bidi_host_predictions": (_FP, [C.c_void_p]),
    "bidi_wait_inputs": (C.c_int, [C.c_void_p]),
    "bidi_set_input": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float]),
    "bidi_step": (C.c_int, [C.c_void_p, _FP, _FP, _FP, C.c_int]),
    "bidi_get_predictions": (C.c_int, [C.c_void_p, _FP]),
    "bidi_get_column_actions": (C.c_int, [C.c_void_p, _FP]),
    "bidi_sync": (C.c_int, [C.c_void_p]),
    "bidi_load_frames": (C.c_int, [C.c_void_p, _FP, _FP, C.c_int]),
    "bidi_step_frame": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "bidi_load_frame_noise": (C.c_int, [C.c_void_p, _FP, _FP, C.c_int]),
    "bidi_set_noise_stream": (C.c_int, [C.c_void_p, C.c_ulonglong, C.c_longlong]),
    "bidi_get_noise": (C.c_int, [C.c_void_p, _FP, _FP]),
    "bidi_set_feedback": (C.c_int, [C.c_void_p, C.c_int, _IP, _IP, C.c_float, C.c_float, C.c_float, C.c_int]),
    "bidi_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "bidi_env_attach": (C.c_int, [C.c_void_p, C.c_int]),
    "bidi_env_reset": (C.c_int, [C.c_void_p]),
    "bidi_env_step": (C.c_int, [C.c_void_p, _FP, _FP, _FP, C.c_int]),
    "bidi_env_get_state": (C.c_int, [C.c_void_p, _FP]),
    "bidi_save_snapshot": (C.c_int, [C.c_void_p, C.c_char_p]),
    "bidi_load_snapshot": (C.c_int, [C.c_void_p, C.c_char_p]),
    "bidi_profile_select": (C.c_int, [C.c_void_p, C.c_char_p]),
    "bidi_profile_collect": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "bidi_timer_start": (C.c_int, [C.c_void_p]),
    "bidi_timer_stop_ms": (C.c_int, [C.c_void_p, _FP]),
    "bidi_launch_count": (C.c_longlong, [C.c_void_p]),
    "bidi_device_bytes": (C.c_longlong, [C.c_void_p]),
    "bidi_abi_sizes": (C.c_int, [_IP, _IP]),
    "bidi_strip_plan": (C.c_int, [C.POINTER(Config), C.POINTER(LayerDesc), C.c_int, C.c_int, _IP]),
    "bidi_strip_configure": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "bidi_strip_link_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "bidi_strip_step_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "bidi_strip_step_frame_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]),
    "bidi_strip_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bidi_strip_link_ipc": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bidi_strip_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "bidi_strip_link_nccl": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bidi_strip_halo_floats": (C.c_longlong, [C.c_void_p]),
    "bidi_create_sharded": (C.c_int, [C.POINTER(Config), C.POINTER(LayerDesc), C.c_int, _IP, C.c_int, C.POINTER(C.c_void_p)]),
    "bidi_sharded_destroy": (None, [C.c_void_p]),
    "bidi_sharded_last_error": (C.c_char_p, [C.c_void_p]),
    "bidi_sharded_num_shards": (C.c_int, [C.c_void_p]),
    "bidi_sharded_num_agents": (C.c_int, [C.c_void_p]),
    "bidi_sharded_shard": (C.c_void_p, [C.c_void_p, C.c_int, _IP, _IP]),
    "bidi_sharded_init_random": (C.c_int, [C.c_void_p, C.c_uint]),
    "bidi_sharded_set_inputs": (C.c_int, [C.c_void_p, _FP]),
    "bidi_sharded_step": (C.c_int, [C.c_void_p, _FP, _FP, _FP, C.c_int]),
    "bidi_sharded_get_predictions": (C.c_int, [C.c_void_p, _FP]),
    "bidi_sharded_get_column_actions": (C.c_int, [C.c_void_p, _FP]),
    "bidi_sharded_load_frames": (C.c_int, [C.c_void_p, _FP, _FP, C.c_int]),
    "bidi_sharded_step_frame": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "bidi_sharded_set_feedback": (C.c_int, [C.c_void_p, C.c_int, _IP, _IP, C.c_float, C.c_float, C.c_float, C.c_int]),
    "bidi_sharded_sync": (C.c_int, [C.c_void_p]),
    "bidi_sharded_timer_start": (C.c_int, [C.c_void_p]),
    "bidi_sharded_timer_stop_ms": (C.c_int, [C.c_void_p, _FP]),
}
SPAN_NAMES = ["OWN_HID", "OWN_VIS", "EXT_HID", "STATE", "ACT", "VIS_RECON", "HID_RECON", "PRED_STATE"]

_lib = None


class BidiError(RuntimeError):
    pass


def load_library():
    """dlopen the in-tree CUDA library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BidiError("%s is missing: run __graft_entry__.build() (no CPU fallback exists)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in ABI.items():
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
        cb, lb = C.c_int(), C.c_int()
        lib.bidi_abi_sizes(C.byref(cb), C.byref(lb))
        if cb.value != C.sizeof(Config) or lb.value != C.sizeof(LayerDesc):
            raise BidiError("config.py structs do not match include/bidinet_b200.h")
        _lib = lib
    return _lib


def _fptr(a):
    return a.ctypes.data_as(_FP)


class Engine:
    """n_agents independent hierarchies of one shape on one CUDA device."""

    def __init__(self, cfg, layers, n_agents=1, device=0, seed=None):
        self.lib = load_library()
        self.cfg, self.layers = cfg, layers
        self.n_layers = cfg.n_layers
        self.n_agents = n_agents
        self.n_in = cfg.in_width * cfg.in_height
        h = C.c_void_p()
        rc = self.lib.bidi_create(C.byref(cfg), layers, n_agents, device, C.byref(h))
        if rc != 0:
            raise BidiError("bidi_create failed (%d): %s" % (rc, self.lib.bidi_last_error(None).decode()))
        self.h = h
        self.n_columns = self.lib.bidi_num_columns(self.h)
        self.n_noise = self.lib.bidi_num_noise(self.h)
        if seed is not None:
            self.init_random(seed)

    def _check(self, rc, what):
        if rc != 0:
            raise BidiError("%s failed (%d): %s" % (what, rc, self.lib.bidi_last_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.bidi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- createRandom
    def init_random(self, seed_base, first=0, count=None):
        """agent a gets the weights std::mt19937(seed_base + a) yields in the reference's draw order."""
        count = self.n_agents - first if count is None else count
        self._check(self.lib.bidi_init_random(self.h, first, count, seed_base), "bidi_init_random")

    def init_random_generator(self, state_text, agent=0):
        """createRandom's draws for ONE agent from the caller's generator: `state_text` is the text form of a std::mt19937
        (operator<<: 624 words and the position); returns the generator's text form after the draws."""
        out = C.create_string_buffer(16384)
        self._check(self.lib.bidi_init_random_generator(self.h, agent, state_text.encode(), out, len(out)), "bidi_init_random_generator")
        return out.value.decode()

    # ---- serialised state
    def array_len(self, which, layer):
        return int(self.lib.bidi_array_len(self.h, ARRAY[which] if isinstance(which, str) else which, layer))

    def get_array(self, which, layer, agent=0):
        w = ARRAY[which] if isinstance(which, str) else which
        n = self.array_len(w, layer)
        out = np.zeros(n, dtype=np.float32)
        if n:
            self._check(self.lib.bidi_get_array(self.h, agent, w, layer, _fptr(out), n), "bidi_get_array")
        return out

    def set_array(self, which, layer, values, agent=0):
        w = ARRAY[which] if isinstance(which, str) else which
        a = np.ascontiguousarray(values, dtype=np.float32)
        self._check(self.lib.bidi_set_array(self.h, agent, w, layer, _fptr(a), a.size), "bidi_set_array")

    def conn_counts(self, which, layer, n_units):
        out = np.zeros(n_units, dtype=np.int32)
        self._check(self.lib.bidi_conn_counts(self.h, ARRAY[which], layer, out.ctypes.data_as(_IP), n_units),
                    "bidi_conn_counts")
        return out

    def state(self, agent=0):
        out = {}
        for name in ARRAY_NAMES:
            for layer in range(self.n_layers + 1):
                if self.array_len(name, layer) > 0:
                    out[(name, layer)] = self.get_array(name, layer, agent)
        return out

    def load_state(self, state, agent=0):
        for (name, layer), a in state.items():
            self.set_array(name, layer, a, agent)

    # ---- per step
    def set_inputs(self, x):
        a = np.ascontiguousarray(x, dtype=np.float32)
        assert a.size == self.n_agents * self.n_in, (a.size, self.n_agents, self.n_in)
        self._check(self.lib.bidi_set_inputs(self.h, _fptr(a)), "bidi_set_inputs")

    def set_input(self, agent, index, value):
        self._check(self.lib.bidi_set_input(self.h, agent, index, value), "bidi_set_input")

    def step(self, rewards=None, noise=None, learn=True):
        r = None if rewards is None else np.ascontiguousarray(rewards, dtype=np.float32).reshape(self.n_agents)
        u = v = None
        if noise is not None:
            u = np.ascontiguousarray(noise[0], dtype=np.float32)
            v = np.ascontiguousarray(noise[1], dtype=np.float32)
            assert u.size == v.size == self.n_agents * self.n_noise
        self._check(self.lib.bidi_step(self.h, None if r is None else _fptr(r), None if u is None else _fptr(u),
                                       None if v is None else _fptr(v), int(bool(learn))), "bidi_step")

    def step_generate(self, noise, normals=None):
        """neo::PredictiveHierarchy::simStepGenerate; normals: [agent][bidi_num_generate_noise] N(0,1) draws or None."""
        z = None
        if normals is not None:
            z = np.ascontiguousarray(normals, dtype=np.float32)
            assert z.size == self.n_agents * self.lib.bidi_num_generate_noise(self.h)
        self._check(self.lib.bidi_step_generate(self.h, None if z is None else _fptr(z), float(noise)), "bidi_step_generate")

    def get_predictions(self, out=None):
        """out: optional [agent][index] float32 array to fill (host_predictions() saves the staging copy)."""
        if out is None:
            out = np.empty((self.n_agents, self.n_in), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == self.n_agents * self.n_in
        self._check(self.lib.bidi_get_predictions(self.h, _fptr(out)), "bidi_get_predictions")
        return out

    def get_inputs(self):
        """[agent][index]: the layer-0 inputs as the last step consumed them."""
        out = np.empty((self.n_agents, self.n_in), dtype=np.float32)
        self._check(self.lib.bidi_get_inputs(self.h, _fptr(out)), "bidi_get_inputs")
        return out

    def host_inputs(self):
        """The engine's page-locked input buffer as an [agent][index] array: fill it, pass it to set_inputs (no
        staging copy), and do not rewrite it before wait_inputs() returns."""
        p = self.lib.bidi_host_inputs(self.h)
        return np.ctypeslib.as_array(p, shape=(self.n_agents, self.n_in))

    def host_predictions(self):
        """The engine's page-locked prediction buffer: pass it to get_predictions(out=...) (no staging copy)."""
        p = self.lib.bidi_host_predictions(self.h)
        return np.ctypeslib.as_array(p, shape=(self.n_agents, self.n_in))

    def wait_inputs(self):
        self._check(self.lib.bidi_wait_inputs(self.h), "bidi_wait_inputs")

    def get_actions(self):
        """sdr::QPRSDR::getActionRel of every action node: [agent][2*n_actions] (actions, then anti-actions)."""
        out = np.zeros((self.n_agents, 2 * self.cfg.q_n_actions), dtype=np.float32)
        self._check(self.lib.bidi_get_actions(self.h, _fptr(out)), "bidi_get_actions")
        return out

    def get_column_actions(self):
        out = np.zeros((self.n_agents, self.n_columns, 3), dtype=np.float32)
        self._check(self.lib.bidi_get_column_actions(self.h, _fptr(out)), "bidi_get_column_actions")
        return out

    def sync(self):
        self._check(self.lib.bidi_sync(self.h), "bidi_sync")

    # ---- device-resident stepping and timing (bench.py)
    def load_frames(self, frames, rewards=None):
        f = np.ascontiguousarray(frames, dtype=np.float32).reshape(-1, self.n_agents, self.n_in)
        r = None if rewards is None else np.ascontiguousarray(rewards, dtype=np.float32).reshape(f.shape[0], self.n_agents)
        self._check(self.lib.bidi_load_frames(self.h, _fptr(f), None if r is None else _fptr(r), f.shape[0]),
                    "bidi_load_frames")

    def step_frame(self, t, learn=True):
        self._check(self.lib.bidi_step_frame(self.h, int(t), int(bool(learn))), "bidi_step_frame")

    def load_frame_noise(self, noise_u, noise_v):
        """[T][agent][n_noise] exploration draws for step_frame (frame t % T); None unloads (engine generator again)."""
        if noise_u is None:
            self._check(self.lib.bidi_load_frame_noise(self.h, None, None, 0), "bidi_load_frame_noise")
            return
        u = np.ascontiguousarray(noise_u, dtype=np.float32).reshape(-1, self.n_agents, self.n_noise)
        v = np.ascontiguousarray(noise_v, dtype=np.float32).reshape(-1, self.n_agents, self.n_noise)
        assert u.shape == v.shape
        self._check(self.lib.bidi_load_frame_noise(self.h, _fptr(u), _fptr(v), u.shape[0]), "bidi_load_frame_noise")

    def set_noise_stream(self, seed, first_global_agent=0):
        self._check(self.lib.bidi_set_noise_stream(self.h, int(seed), int(first_global_agent)), "bidi_set_noise_stream")

    def get_noise(self):
        """(u, v), [agent][n_noise] each: the exploration draws the last step consumed."""
        u = np.zeros((self.n_agents, self.n_noise), dtype=np.float32)
        v = np.zeros((self.n_agents, self.n_noise), dtype=np.float32)
        self._check(self.lib.bidi_get_noise(self.h, _fptr(u), _fptr(v)), "bidi_get_noise")
        return u, v

    def set_feedback(self, src, dst, scale=0.5, bias=0.5, noise_std=0.05, reward_from_actions=True):
        s = np.ascontiguousarray(src, dtype=np.int32)
        d = np.ascontiguousarray(dst, dtype=np.int32)
        assert s.size == d.size
        self._check(self.lib.bidi_set_feedback(self.h, s.size, s.ctypes.data_as(_IP), d.ctypes.data_as(_IP), scale, bias,
                                               noise_std, int(bool(reward_from_actions))), "bidi_set_feedback")

    # ---- batched headless environment on the device (bidi_env_*)
    def env_attach(self, kind=1):
        self._check(self.lib.bidi_env_attach(self.h, kind), "bidi_env_attach")

    def env_reset(self):
        self._check(self.lib.bidi_env_reset(self.h), "bidi_env_reset")

    def env_step(self, noise=None, env_noise=None, learn=True):
        """One pass of the RunnerMain loop for every agent: reward and inputs from the environment, simStep, motors, physics."""
        u = v = z = None
        if noise is not None:
            u = np.ascontiguousarray(noise[0], dtype=np.float32)
            v = np.ascontiguousarray(noise[1], dtype=np.float32)
            assert u.size == v.size == self.n_agents * self.n_noise
        if env_noise is not None:
            z = np.ascontiguousarray(env_noise, dtype=np.float32)
            assert z.size == self.n_agents * 10
        self._check(self.lib.bidi_env_step(self.h, None if u is None else _fptr(u), None if v is None else _fptr(v),
                                           None if z is None else _fptr(z), int(bool(learn))), "bidi_env_step")

    def env_state(self):
        out = np.zeros((self.n_agents, 24), dtype=np.float32)
        self._check(self.lib.bidi_env_get_state(self.h, _fptr(out)), "bidi_env_get_state")
        return out

    def save_snapshot(self, path):
        self._check(self.lib.bidi_save_snapshot(self.h, os.fsencode(path)), "bidi_save_snapshot")

    def load_snapshot(self, path):
        self._check(self.lib.bidi_load_snapshot(self.h, os.fsencode(path)), "bidi_load_snapshot")

    def set_option(self, name, value):
        self._check(self.lib.bidi_set_option(self.h, name.encode(), int(value)), "bidi_set_option")

    def profile_select(self, kernel_name):
        self._check(self.lib.bidi_profile_select(self.h, kernel_name.encode() if kernel_name else None),
                    "bidi_profile_select")

    def profile_collect(self):
        ms, n = C.c_double(), C.c_longlong()
        self._check(self.lib.bidi_profile_collect(self.h, C.byref(ms), C.byref(n)), "bidi_profile_collect")
        return ms.value, n.value

    def timer_start(self):
        self._check(self.lib.bidi_timer_start(self.h), "bidi_timer_start")

    def timer_stop_ms(self):
        ms = C.c_float()
        self._check(self.lib.bidi_timer_stop_ms(self.h, C.byref(ms)), "bidi_timer_stop_ms")
        return ms.value

    # ---- spatial split of one oversized layer (include/bidinet_b200.h, bidi_strip_*)
    def strip_configure(self, part, n_parts):
        self._check(self.lib.bidi_strip_configure(self.h, part, n_parts), "bidi_strip_configure")
        self.strip_part, self.strip_n = part, n_parts
        self.spans = strip_plan(self.cfg, self.layers, part, n_parts)

    def strip_link_nccl(self, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self.lib.bidi_strip_link_nccl(self.h, buf), "bidi_strip_link_nccl")

    def strip_ipc_export(self):
        """128 bytes (CUDA IPC handles of this part's state and flag words) for strip_link_ipc."""
        buf = C.create_string_buffer(128)
        self._check(self.lib.bidi_strip_ipc_export(self.h, buf), "bidi_strip_ipc_export")
        return buf.raw

    def strip_link_ipc(self, handles):
        """handles: the 128-byte records of ALL parts, in part order."""
        raw = b"".join(bytes(x) for x in handles)
        assert len(raw) == 128 * self.strip_n
        buf = C.create_string_buffer(raw, len(raw))
        self._check(self.lib.bidi_strip_link_ipc(self.h, buf), "bidi_strip_link_ipc")

    @property
    def strip_halo_floats(self):
        return int(self.lib.bidi_strip_halo_floats(self.h))

    @property
    def launch_count(self):
        return int(self.lib.bidi_launch_count(self.h))

    @property
    def device_bytes(self):
        return int(self.lib.bidi_device_bytes(self.h))


def strip_plan(cfg, layers, part, n_parts):
    """{span name: (lo, hi)} row intervals of one part of a split layer (pure geometry, no GPU)."""
    lib = load_library()
    out = (C.c_int * (2 * len(SPAN_NAMES)))()
    rc = lib.bidi_strip_plan(C.byref(cfg), layers, part, n_parts, out)
    if rc != 0:
        raise BidiError("bidi_strip_plan failed (%d): %s" % (rc, lib.bidi_last_error(None).decode()))
    return {n: (out[2 * i], out[2 * i + 1]) for i, n in enumerate(SPAN_NAMES)}


def nccl_unique_id():
    """128 bytes for bidi_strip_link_nccl; call on one rank and ship them to the others."""
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.bidi_strip_nccl_unique_id(buf)
    if rc != 0:
        raise BidiError("bidi_strip_nccl_unique_id failed (%d): %s" % (rc, lib.bidi_last_error(None).decode()))
    return buf.raw


class _ShardView(Engine):
    """An Engine view of one shard of a ShardedEngine (owned by the sharded handle: never destroyed here)."""

    def __init__(self, lib, h, cfg, layers, n_agents):
        self.lib, self.h, self.cfg, self.layers = lib, C.c_void_p(h), cfg, layers
        self.n_layers, self.n_agents, self.n_in = cfg.n_layers, n_agents, cfg.in_width * cfg.in_height
        self.n_columns = lib.bidi_num_columns(self.h)
        self.n_noise = lib.bidi_num_noise(self.h)

    def close(self):
        self.h = None


class ShardedEngine:
    """n_agents independent hierarchies over several devices of this process (bidi_create_sharded): global agent g
    lives on shard k with first[k] <= g < first[k] + count[k].  Same calls as Engine, arrays indexed by global agent."""

    def __init__(self, cfg, layers, n_agents, devices, seed=None):
        self.lib = load_library()
        self.cfg, self.layers = cfg, layers
        self.n_layers, self.n_agents, self.n_in = cfg.n_layers, n_agents, cfg.in_width * cfg.in_height
        dev = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.lib.bidi_create_sharded(C.byref(cfg), layers, n_agents, dev, len(devices), C.byref(h))
        if rc != 0:
            raise BidiError("bidi_create_sharded failed (%d): %s" % (rc, self.lib.bidi_sharded_last_error(None).decode()))
        self.h = h
        self.shards, self.first, self.count = [], [], []
        for k in range(self.lib.bidi_sharded_num_shards(self.h)):
            f, c = C.c_int(), C.c_int()
            e = self.lib.bidi_sharded_shard(self.h, k, C.byref(f), C.byref(c))
            self.shards.append(_ShardView(self.lib, e, cfg, layers, c.value))
            self.first.append(f.value)
            self.count.append(c.value)
        self.n_columns, self.n_noise = self.shards[0].n_columns, self.shards[0].n_noise
        if seed is not None:
            self.init_random(seed)

    def _check(self, rc, what):
        if rc != 0:
            raise BidiError("%s failed (%d): %s" % (what, rc, self.lib.bidi_sharded_last_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.bidi_sharded_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def locate(self, agent):
        """(shard view, local agent index) of a global agent."""
        for e, f, c in zip(self.shards, self.first, self.count):
            if f <= agent < f + c:
                return e, agent - f
        raise IndexError(agent)

    def state(self, agent=0):
        e, a = self.locate(agent)
        return e.state(agent=a)

    def init_random(self, seed_base):
        self._check(self.lib.bidi_sharded_init_random(self.h, seed_base), "bidi_sharded_init_random")

    def set_inputs(self, x):
        a = np.ascontiguousarray(x, dtype=np.float32)
        assert a.size == self.n_agents * self.n_in
        self._check(self.lib.bidi_sharded_set_inputs(self.h, _fptr(a)), "bidi_sharded_set_inputs")

    def step(self, rewards=None, noise=None, learn=True):
        r = None if rewards is None else np.ascontiguousarray(rewards, dtype=np.float32).reshape(self.n_agents)
        u = v = None
        if noise is not None:
            u = np.ascontiguousarray(noise[0], dtype=np.float32)
            v = np.ascontiguousarray(noise[1], dtype=np.float32)
            assert u.size == v.size == self.n_agents * self.n_noise
        self._check(self.lib.bidi_sharded_step(self.h, None if r is None else _fptr(r), None if u is None else _fptr(u),
                                               None if v is None else _fptr(v), int(bool(learn))), "bidi_sharded_step")

    def get_predictions(self):
        out = np.empty((self.n_agents, self.n_in), dtype=np.float32)
        self._check(self.lib.bidi_sharded_get_predictions(self.h, _fptr(out)), "bidi_sharded_get_predictions")
        return out

    def get_column_actions(self):
        out = np.zeros((self.n_agents, self.n_columns, 3), dtype=np.float32)
        self._check(self.lib.bidi_sharded_get_column_actions(self.h, _fptr(out)), "bidi_sharded_get_column_actions")
        return out

    def load_frames(self, frames, rewards=None):
        f = np.ascontiguousarray(frames, dtype=np.float32).reshape(-1, self.n_agents, self.n_in)
        r = None if rewards is None else np.ascontiguousarray(rewards, dtype=np.float32).reshape(f.shape[0], self.n_agents)
        self._check(self.lib.bidi_sharded_load_frames(self.h, _fptr(f), None if r is None else _fptr(r), f.shape[0]),
                    "bidi_sharded_load_frames")

    def step_frame(self, t, learn=True):
        self._check(self.lib.bidi_sharded_step_frame(self.h, int(t), int(bool(learn))), "bidi_sharded_step_frame")

    def set_feedback(self, src, dst, scale=0.5, bias=0.5, noise_std=0.05, reward_from_actions=True):
        s = np.ascontiguousarray(src, dtype=np.int32)
        d = np.ascontiguousarray(dst, dtype=np.int32)
        self._check(self.lib.bidi_sharded_set_feedback(self.h, s.size, s.ctypes.data_as(_IP), d.ctypes.data_as(_IP), scale, bias,
                                                       noise_std, int(bool(reward_from_actions))), "bidi_sharded_set_feedback")

    def sync(self):
        self._check(self.lib.bidi_sharded_sync(self.h), "bidi_sharded_sync")

    def timer_start(self):
        self._check(self.lib.bidi_sharded_timer_start(self.h), "bidi_sharded_timer_start")

    def timer_stop_ms(self):
        ms = C.c_float()
        self._check(self.lib.bidi_sharded_timer_stop_ms(self.h, C.byref(ms)), "bidi_sharded_timer_stop_ms")
        return ms.value

    @property
    def launch_count(self):
        return sum(e.launch_count for e in self.shards)

    @property
    def device_bytes(self):
        return sum(e.device_bytes for e in self.shards)


class StripGroup:
    """One oversized single-layer k-winners hierarchy split into row strips over several
    engines of THIS process (devices[p] hosts part p; the same device may repeat).  The
    parts exchange halo rows by peer copies; results equal the unsplit engine's bit for
    bit.  State arrays come back assembled from the rows each part owns."""

    def __init__(self, cfg, layers, devices, n_agents=1, seed=None):
        self.cfg, self.layers = cfg, layers
        n = len(devices)
        self.parts = [Engine(cfg, layers, n_agents=n_agents, device=d) for d in devices]
        for p, e in enumerate(self.parts):
            e.strip_configure(p, n)
        self.lib = self.parts[0].lib
        self._handles = (C.c_void_p * n)(*[e.h for e in self.parts])
        rc = self.lib.bidi_strip_link_local(self._handles, n)
        if rc != 0:
            raise BidiError("bidi_strip_link_local failed (%d): %s" % (rc, self.lib.bidi_last_error(self.parts[0].h).decode()))
        self.n_in, self.n_agents = self.parts[0].n_in, n_agents
        if seed is not None:
            for e in self.parts:
                e.init_random(seed)

    def close(self):
        for e in self.parts:
            e.close()

    def load_state(self, state, agent=0):
        for e in self.parts:
            e.load_state(state, agent)

    def set_inputs(self, x):
        for e in self.parts:
            e.set_inputs(x)

    def load_frames(self, frames):
        for e in self.parts:
            e.load_frames(frames)

    def step(self, learn=True):
        rc = self.lib.bidi_strip_step_local(self._handles, len(self.parts), int(bool(learn)))
        if rc != 0:
            raise BidiError("bidi_strip_step_local failed (%d): %s" % (rc, self.lib.bidi_last_error(self.parts[0].h).decode()))

    def step_frame(self, t, learn=True):
        rc = self.lib.bidi_strip_step_frame_local(self._handles, len(self.parts), int(t), int(bool(learn)))
        if rc != 0:
            raise BidiError("bidi_strip_step_frame_local failed (%d): %s" % (rc, self.lib.bidi_last_error(self.parts[0].h).decode()))

    def sync(self):
        for e in self.parts:
            e.sync()

    def get_predictions(self):
        return assemble_rows([e.get_predictions() for e in self.parts], [e.spans["OWN_VIS"] for e in self.parts],
                             self.cfg.in_width)

    def state(self, agent=0):
        return assemble_state(self.parts, agent)


def assemble_rows(arrays, spans, width):
    """Whole-grid plane [..., H*W] from per-part copies: rows [lo,hi) come from the part that owns them."""
    out = np.array(arrays[0], copy=True)
    for a, (lo, hi) in zip(arrays, spans):
        out[..., lo * width:hi * width] = a[..., lo * width:hi * width]
    return out


_VISIBLE_PLANES = ("VIS_RECON", "PREDICTION", "VIS_INPUT")  # a part uploads only the input rows its windows reach
_REPLICATED = ()


def assemble_state(parts, agent=0):
    """The serialised state of a split layer: every array taken from the part that owns the rows
    (unit planes) or the units (connection lists, which are concatenated unit after unit)."""
    cfg, ld = parts[0].cfg, parts[0].layers
    vw, hw = cfg.in_width, ld[0].width
    states = [e.state(agent) for e in parts]
    out = {}
    for key in states[0]:
        name, layer = key
        if name in _REPLICATED:
            out[key] = states[0][key]
        elif name in _VISIBLE_PLANES:
            out[key] = assemble_rows([s[key] for s in states], [e.spans["OWN_VIS"] for e in parts], vw)
        elif name in ("FF_W", "REC_W", "PRED_W", "LAT_W", "FB_W"):
            counts = parts[0].conn_counts(name, layer, ld[0].width * ld[0].height)
            offs = np.concatenate([[0], np.cumsum(counts)])
            res = np.array(states[0][key], copy=True)
            for s, e in zip(states, parts):
                lo, hi = e.spans["OWN_HID"]
                res[offs[lo * hw]:offs[hi * hw]] = s[key][offs[lo * hw]:offs[hi * hw]]
            out[key] = res
        else:
            out[key] = assemble_rows([s[key] for s in states], [e.spans["OWN_HID"] for e in parts], hw)
    return out
