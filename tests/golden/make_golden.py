#!/usr/bin/env python3
"""Copies the reference's own golden result files that pin the hot path into
tests/golden/ (they are DATA of the reference's test-suite, 16 lines each) so
that the oracle can be checked on machines without /root/reference.

  mus/examples/fluid/benchmark/gaussianPulse/reference/
     gaussianPulse_pressAlongLength_p00000_t10.001E+00.res   (level 4, np=2, t=10.001)

Run in the build container only:  python tests/golden/make_golden.py
"""
import os
import shutil

REF = "/root/reference/mus/examples/fluid/benchmark/gaussianPulse/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
for f in ("gaussianPulse_pressAlongLength_p00000_t10.001E+00.res",):
    shutil.copy(os.path.join(REF, f), os.path.join(HERE, f))
    print("copied", f)
