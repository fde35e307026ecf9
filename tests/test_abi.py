"""The C-ABI library loads and exports every symbol include/bidinet_b200.h declares;
no compute is called (CPU box).  Also: the product never routes through oracle/."""
import ctypes
import os
import re

import pytest

import bidinet_b200
from bidinet_b200.build import LIB_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bidinet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bidi_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(bidinet_b200.ABI) == names  # the ctypes binding covers the whole header


def test_struct_mirrors_match_header():
    bidinet_b200.load_library()  # raises if sizeof(Config/LayerDesc) differ from the C structs


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from bidinet_b200 import config as cf
    cfg, ld = cf.config_c1(cf.KWINNERS)
    with pytest.raises(bidinet_b200.BidiError, match="no usable CUDA device"):
        bidinet_b200.Engine(cfg, ld)


def test_bad_config_is_rejected():
    lib = bidinet_b200.load_library()
    from bidinet_b200 import config as cf
    cfg, ld = cf.config_c1(cf.KWINNERS)
    cfg.n_layers = 0
    h = ctypes.c_void_p()
    assert lib.bidi_create(ctypes.byref(cfg), ld, 1, 0, ctypes.byref(h)) == -1
    assert b"n_layers" in lib.bidi_last_error(None)


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "bidinet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("# oracle", ""), f
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "oracle/" not in open(p).read()
    # developer tools outside tests/ may not run the oracle either (checker scripts live in tests/checks/)
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            src = open(os.path.join(ROOT, "tools", f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f


def test_facade_headers_build_and_link_against_the_abi():
    """The drop-in C++ facade (include/bidinet/{sdr,neo,deep}/*.h) compiles and links against the C-ABI library on a box
    without a GPU: one demo source drives every facade class (tests/facade/facade_demo.cpp; it is RUN by the gpu tests)."""
    import subprocess
    d = os.path.join(ROOT, "tests", "facade")
    subprocess.check_call(["make", "-C", d, "facade_demo", "batch_demo"])
    assert os.path.exists(os.path.join(d, "facade_demo"))


def test_facade_golden_is_what_the_reference_prints():
    """tests/golden/facade_demo.txt is the output of the same demo source built against the unmodified reference
    (only where /root/reference exists: this container; the GPU box compares the facade build with the committed text)."""
    import subprocess
    if not os.path.isdir("/root/reference/BIDInet/source"):
        pytest.skip("/root/reference absent")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "facade"), "ref"])
    out = subprocess.check_output([os.path.join(ROOT, "oracle", "_ref", "facade_demo_ref")]).decode()
    assert out == open(os.path.join(ROOT, "tests", "golden", "facade_demo.txt")).read()
