"""bidinet_b200: B200-native engine for BIDInet's per-timestep hierarchy update.

The product is the CUDA library behind include/bidinet_b200.h; this package holds
its sources (csrc/), the build recipe and a thin ctypes binding used by the tests
and bench.py.
"""
from . import config  # noqa: F401
from .engine import (ABI, BidiError, Engine, ShardedEngine, StripGroup, assemble_rows, assemble_state, load_library,  # noqa: F401
                     nccl_unique_id, strip_plan)
