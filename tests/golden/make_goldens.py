#!/usr/bin/env python
"""Regenerates the fixtures under tests/golden/ that come from the reference
checkout (run in the build container, where /root/reference is mounted; the GPU
box only ever sees the committed copies).

* ``examples/*.qasm``      -> circuit inputs (configs 1-2 of BASELINE.json)
* ``examples/ghz_3.json`` + ``ghz_3_plan.json`` -> ``ghz_3_contracted.json``:
  the reference's own stored input network / plan / contracted tensor, the one
  place where the reference pins tensor *values and layout* of a contraction.
"""
import os
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

FILES = ["ghz_3.qasm", "qft_2.qasm", "qft_3.qasm", "qft_5.qasm", "qft_10.qasm",
         "ghz_3.json", "ghz_3_plan.json", "ghz_3_contracted.json"]

for name in FILES:
    shutil.copyfile(os.path.join(REF, "examples", name), os.path.join(HERE, name))
    print("copied", name)
