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
    return LIB_PATH
