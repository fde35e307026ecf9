"""Builds bidinet_b200/libbidinet_b200.so (CUDA kernels + C ABI) in-tree for sm_100a."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbidinet_b200.so")


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (see csrc/Makefile)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc")]
    if force:
        cmd.append("-B")
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libbidinet_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout + out.stderr)
    build_runner_env()
    return LIB_PATH


RUNNER_ENV_PATH = os.path.join(os.path.dirname(_HERE), "tools", "librunner_env.so")


def build_runner_env():
    """gcc: the synthetic environment of bench.py's end-to-end loop (tools/runner_env.c), host code only."""
    src = os.path.join(os.path.dirname(_HERE), "tools", "runner_env.c")
    if os.path.exists(RUNNER_ENV_PATH) and os.path.getmtime(RUNNER_ENV_PATH) >= os.path.getmtime(src):
        return RUNNER_ENV_PATH
    tmp = "%s.%d.tmp" % (RUNNER_ENV_PATH, os.getpid())  # several ranks may build at once: write aside, rename atomically
    out = subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", tmp, src], capture_output=True, text=True)
    if out.returncode != 0:
        if os.path.exists(RUNNER_ENV_PATH):
            return RUNNER_ENV_PATH
        raise RuntimeError("building librunner_env.so failed:\n" + out.stdout + out.stderr)
    os.replace(tmp, RUNNER_ENV_PATH)
    return RUNNER_ENV_PATH
