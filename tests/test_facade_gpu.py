"""The C++ facade (include/bidinet/{sdr,neo}/*.h over the C ABI) as a drop-in: the SAME
source file that drives the reference's classes like its experiment mains do
(tests/facade/facade_demo.cpp) is built against the facade and must print, line for line,
what the build against the unmodified reference printed here
(tests/golden/facade_demo.txt, produced by oracle/_ref/facade_demo_ref).  That covers
createRandom's draw order from the caller's std::mt19937, the per-step prediction vectors
of all four classes, the exploration draws of neo::Agent::simStep and the generator's
final state."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "tests", "facade", "facade_demo")
GOLDEN = os.path.join(ROOT, "tests", "golden", "facade_demo.txt")


@pytest.mark.gpu
def test_facade_demo_prints_what_the_reference_prints():
    subprocess.check_call(["make", "-C", os.path.dirname(DEMO), "facade_demo"])  # incremental: never a stale binary
    out = subprocess.run([DEMO], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    got, want = out.stdout.splitlines(), open(GOLDEN).read().splitlines()
    assert len(got) == len(want)
    for k, (g, w) in enumerate(zip(got, want)):
        assert g == w, "line %d: facade %r != reference %r" % (k, g, w)


@pytest.mark.gpu
def test_agent_batch_equals_separate_agents():
    """neo::AgentBatch over two shards (bidi_create_sharded) against separate neo::Agent objects, bit for bit."""
    d = os.path.dirname(DEMO)
    subprocess.check_call(["make", "-C", d, "batch_demo"])
    out = subprocess.run([os.path.join(d, "batch_demo")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("mismatches 0")


@pytest.mark.gpu
@pytest.mark.gpu2
def test_agent_batch_on_two_devices():
    from conftest import require_two_gpus
    require_two_gpus()
    d = os.path.dirname(DEMO)
    subprocess.check_call(["make", "-C", d, "batch_demo"])
    out = subprocess.run([os.path.join(d, "batch_demo"), "0", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "shards 2" in out.stdout and out.stdout.strip().endswith("mismatches 0")


def test_reference_build_still_matches_golden():
    """Where the reference tree exists, its own build of the demo reproduces the golden file."""
    ref = os.path.join(ROOT, "oracle", "_ref", "facade_demo_ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/facade_demo_ref not built (needs /root/reference)")
    out = subprocess.run([ref], capture_output=True, text=True, timeout=600)
    assert out.stdout == open(GOLDEN).read()
