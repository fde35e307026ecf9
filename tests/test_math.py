"""bidi_math.h's expf restatement (host build) against the running C library."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROG = r'''
#include "bidi_math.h"
#include <stdio.h>
int main(void) {
  unsigned long bad = 0, n = 0;
  for (uint32_t b = 0; b < 0xffffff00u; b += 197) {
    float x; memcpy(&x, &b, 4);
    if (!(x == x) || x > 90.f || x < -110.f) continue;
    float a = expf(x), c = bidi_expf(x);
    uint32_t ua, uc; memcpy(&ua, &a, 4); memcpy(&uc, &c, 4);
    n++; if (ua != uc) bad++;
  }
  float s0 = bidi_sigmoid(0.0f), s1 = 1.0f / (1.0f + expf(-1.25f));
  printf("%lu %lu %d\n", n, bad, s0 == 0.5f && bidi_sigmoid(1.25f) == s1);
  return 0;
}
'''


def test_expf_matches_libm_bit_for_bit():
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(PROG)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "bidinet_b200", "csrc"),
                               "-o", exe, src, "-lm"])
        n, bad, ok = map(int, subprocess.check_output([exe]).split())
    assert n > 10_000_000 and bad == 0 and ok == 1
