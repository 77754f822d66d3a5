#!/usr/bin/env python
"""Writes tests/golden/bench_specs.json: the twb_spec (raw bytes, hex) of every bench workload, so that bench.py's reference
arm (the CPU oracle) can evaluate the same problem WITHOUT loading libtowr_b200.so.  Generated with the product's host
code (towr_b200.make_formulation); tests/test_structure.py checks the fixture against a fresh build of the recipes."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import towr_b200 as tb  # noqa: E402

out = {}
for name in ("hopper", "anymal_trot_block", "biped_walk_stairs", "hyq_gallop_gap", "anymal_trot_mixed", "go1_trot_flat"):
    spec = tb.make_formulation(name).to_spec()
    out[name] = bytes(spec).hex()
path = os.path.join(ROOT, "tests", "golden", "bench_specs.json")
with open(path, "w") as fh:
    json.dump(out, fh, indent=0)
print("wrote", path, {k: len(v) // 2 for k, v in out.items()})
