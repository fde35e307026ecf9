"""bench.py's reference arm, run here on the CPU with a tiny sample: exactly ONE line on stdout, valid JSON,
carrying the keys the driver reads (metric / unit / config shared with the GPU arm, impl, cpu_baseline, e2e)."""
import json
import os
import subprocess
import sys

import pytest

from oracle.pyoracle import have_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libbidiref.so not built (needs /root/reference)")
def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--ref-inner-steps", "3"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "agent_steps_per_s" and j["unit"] == "agent-steps/s"
    assert j["higher_is_better"] is True and j["steps"] == 2 and j["warmup"] == 1 and j["n_gpus"] == 1
    assert j["value"] > 0 and j["vs_baseline"] is None and j["gpu_launches"] == 0
    assert j["config"]["workload"].startswith("C3")
    assert j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_stdout_guard_keeps_library_banners_off_stdout():
    """bench._emit_one_line: what native code writes to file descriptor 1 while the bench runs goes to stderr."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench\n"
            "def fake_main():\n"
            "    os.write(1, b'NCCL version banner\\n')\n"
            "    print('{\"ok\": 1}')\n"
            "bench.main = fake_main; bench._emit_one_line()\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout == '{"ok": 1}\n'
    assert "NCCL version banner" in out.stderr
