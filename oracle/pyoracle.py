"""TEST INFRASTRUCTURE: ctypes front-ends of the two CPU checkers.

``RefHierarchy``    -> oracle/_ref/libbidiref.so (the unmodified reference, compiled
                       from /root/reference by oracle/Makefile)
``OracleHierarchy`` -> oracle/liboracle.so (the C restatement, bidi_oracle.c)

Both expose the same methods so tests can swap them.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from bidinet_b200.config import ARRAY, ARRAY_NAMES, Config, LayerDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libbidiref.so")
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_FP = C.POINTER(C.c_float)
_IP = C.POINTER(C.c_int)


def build(verbose=False):
    """Compile liboracle.so and, where /root/reference exists, _ref/libbidiref.so."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def have_ref():
    return os.path.exists(REF_SO)


def _bind(lib, p):
    sig = {
        "create": (C.c_void_p, [C.POINTER(Config), C.POINTER(LayerDesc), C.c_uint]),
        "destroy": (None, [C.c_void_p]),
        "array_len": (C.c_longlong, [C.c_void_p, C.c_int, C.c_int]),
        "get_array": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _FP]),
        "set_array": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _FP]),
        "conn_counts": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _IP]),
        "set_inputs": (None, [C.c_void_p, _FP, C.c_int]),
        "get_predictions": (None, [C.c_void_p, _FP, C.c_int]),
        "num_columns": (C.c_int, [C.c_void_p]),
        "num_noise": (C.c_int, [C.c_void_p]),
        "num_generate_noise": (C.c_int, [C.c_void_p]),
        "peek_generate": (C.c_int, [C.c_void_p, _FP]),
        "step_generate": (None, [C.c_void_p, C.c_float]),
        "peek_noise": (C.c_int, [C.c_void_p, _FP, _FP]),
        "step": (None, [C.c_void_p, C.c_float, C.c_int]),
        "run": (None, [C.c_void_p, _FP, C.c_int, C.c_int, _FP, C.c_int, C.c_int]),
        "run_runner": (None, [C.c_void_p, _FP, C.c_int, C.c_int, C.c_int, C.c_int, _IP, _IP, C.c_int]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, p + name)
        f.restype, f.argtypes = res, args
    if p == "orc_":
        lib.orc_step_noise.restype = None
        lib.orc_step_noise.argtypes = [C.c_void_p, C.c_float, _FP, _FP, C.c_int]
    return lib


_libs = {}


def _lib(prefix):
    if prefix not in _libs:
        path = REF_SO if prefix == "ref_" else ORACLE_SO
        if not os.path.exists(path):
            build()
        _libs[prefix] = _bind(C.CDLL(path), prefix)
    return _libs[prefix]


def _fptr(a):
    return a.ctypes.data_as(_FP)


class _Hierarchy:
    prefix = None

    def __init__(self, cfg, layers, seed=1234):
        self.lib = _lib(self.prefix)
        self.cfg, self.layers = cfg, layers
        self.n_layers = cfg.n_layers
        self.n_in = cfg.in_width * cfg.in_height
        self._f = lambda n: getattr(self.lib, self.prefix + n)
        self.h = self._f("create")(C.byref(cfg), layers, seed)
        if not self.h:
            raise RuntimeError("create failed")

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def array_len(self, which, layer):
        return int(self._f("array_len")(self.h, ARRAY[which] if isinstance(which, str) else which, layer))

    def get_array(self, which, layer):
        w = ARRAY[which] if isinstance(which, str) else which
        n = self.array_len(w, layer)
        out = np.zeros(n, dtype=np.float32)
        if n and self._f("get_array")(self.h, w, layer, _fptr(out)) != 0:
            raise RuntimeError("get_array(%s,%d) failed" % (which, layer))
        return out

    def set_array(self, which, layer, values):
        w = ARRAY[which] if isinstance(which, str) else which
        a = np.ascontiguousarray(values, dtype=np.float32)
        assert a.size == self.array_len(w, layer), (which, layer, a.size)
        if a.size and self._f("set_array")(self.h, w, layer, _fptr(a)) != 0:
            raise RuntimeError("set_array(%s,%d) failed" % (which, layer))

    def conn_counts(self, which, layer, n_units):
        out = np.zeros(n_units, dtype=np.int32)
        rc = self._f("conn_counts")(self.h, ARRAY[which], layer, out.ctypes.data_as(_IP))
        return out if rc == 0 else None

    def state(self):
        """{(array name, layer): ndarray} of every non-empty state array."""
        out = {}
        for name in ARRAY_NAMES:
            for layer in range(self.n_layers + 1):
                if self.array_len(name, layer):
                    out[(name, layer)] = self.get_array(name, layer)
        return out

    def load_state(self, state):
        for (name, layer), a in state.items():
            self.set_array(name, layer, a)

    def set_inputs(self, x):
        a = np.ascontiguousarray(x, dtype=np.float32).ravel()
        assert a.size == self.n_in
        self._f("set_inputs")(self.h, _fptr(a), a.size)

    def get_predictions(self):
        out = np.zeros(self.n_in, dtype=np.float32)
        self._f("get_predictions")(self.h, _fptr(out), out.size)
        return out

    def num_columns(self):
        return int(self._f("num_columns")(self.h))

    def num_noise(self):
        """floats per step in each of the exploration-noise arrays (neo::Agent columns, sdr::QPRSDR actions)"""
        return int(self._f("num_noise")(self.h))

    def peek_noise(self):
        n = self.num_noise()
        u = np.zeros(n, dtype=np.float32)
        v = np.zeros(n, dtype=np.float32)
        if n:
            assert self._f("peek_noise")(self.h, _fptr(u), _fptr(v)) == 0
        return u, v

    def step(self, reward=0.0, learn=True):
        self._f("step")(self.h, float(reward), int(bool(learn)))

    def peek_generate(self):
        """the N(0,1) values the next simStepGenerate will draw (neo::PredictiveHierarchy only)"""
        out = np.zeros(int(self._f("num_generate_noise")(self.h)), dtype=np.float32)
        assert out.size and self._f("peek_generate")(self.h, _fptr(out)) == 0
        return out

    def step_generate(self, noise):
        self._f("step_generate")(self.h, float(noise))

    def run(self, frames, steps, rewards=None, learn=True):
        fr = np.ascontiguousarray(frames, dtype=np.float32).reshape(-1, self.n_in)
        rw = None if rewards is None else np.ascontiguousarray(rewards, dtype=np.float32)
        assert rw is None or rw.size == fr.shape[0]
        self._f("run")(self.h, _fptr(fr), fr.shape[0], self.n_in, None if rw is None else _fptr(rw),
                       int(steps), int(bool(learn)))


    def run_runner(self, frames, steps, src, dst, learn=True):
        """The headless RunnerMain loop (closed action feedback) for `steps` steps."""
        fr = np.ascontiguousarray(frames, dtype=np.float32).reshape(-1, self.n_in)
        s = np.ascontiguousarray(src, dtype=np.int32)
        d = np.ascontiguousarray(dst, dtype=np.int32)
        self._f("run_runner")(self.h, _fptr(fr), fr.shape[0], self.n_in, int(steps), s.size,
                              s.ctypes.data_as(_IP), d.ctypes.data_as(_IP), int(bool(learn)))


class RefHierarchy(_Hierarchy):
    prefix = "ref_"


class OracleHierarchy(_Hierarchy):
    prefix = "orc_"

    def step_noise(self, reward, u, v, learn=True):
        u = np.ascontiguousarray(u, dtype=np.float32)
        v = np.ascontiguousarray(v, dtype=np.float32)
        self.lib.orc_step_noise(self.h, float(reward), _fptr(u) if u.size else None,
                                _fptr(v) if v.size else None, int(bool(learn)))


class RunnerEnv:
    """CPU twin of the device environment BIDI_ENV_RUNNER (oracle/runner_env_oracle.c), one agent's world."""

    def __init__(self):
        lib = _lib("orc_")
        if not hasattr(lib, "_renv_bound"):
            lib.renv_create.restype = C.c_void_p
            lib.renv_create.argtypes = []
            lib.renv_destroy.restype = None
            lib.renv_destroy.argtypes = [C.c_void_p]
            lib.renv_reward.restype = C.c_float
            lib.renv_reward.argtypes = [C.c_void_p]
            for n in ("renv_state", "renv_clocks", "renv_motor_update", "renv_dump"):
                getattr(lib, n).restype = None
                getattr(lib, n).argtypes = [C.c_void_p, _FP]
            lib._renv_bound = True
        self.lib = lib
        self.h = lib.renv_create()

    def __del__(self):
        try:
            self.lib.renv_destroy(self.h)
        except Exception:
            pass

    def reward(self):
        return float(self.lib.renv_reward(self.h))

    def _get(self, name, n):
        out = np.zeros(n, dtype=np.float32)
        getattr(self.lib, name)(self.h, _fptr(out))
        return out

    def state(self):
        return self._get("renv_state", 15)

    def clocks(self):
        return self._get("renv_clocks", 4)

    def dump(self):
        return self._get("renv_dump", 24)

    def motor_update(self, action):
        a = np.ascontiguousarray(action, dtype=np.float32)
        assert a.size == 10
        self.lib.renv_motor_update(self.h, _fptr(a))


def fnv1a64(h, arr):
    """FNV-1a-64 over the raw little-endian bytes of a float32 array."""
    for b in np.ascontiguousarray(arr, dtype="<f4").tobytes():
        h ^= b
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def c1_frame(t, w=16, h=16):
    """SURVEY 8d C1 input: x(x,y,t) = ((x+p)%8<2 && (y+2p)%8<3), p = t%10."""
    p = t % 10
    x = np.arange(w)[None, :]
    y = np.arange(h)[:, None]
    return (((x + p) % 8 < 2) & ((y + 2 * p) % 8 < 3)).astype(np.float32).ravel()
