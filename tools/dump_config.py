#!/usr/bin/env python
"""Dump a synthetic config as raw little-endian files for baseline/julia/bench_ref.jl (the true reference, for anyone
with Julia): python tools/dump_config.py C2 8 /tmp/c2_8"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from lowrankmodels_b200 import synth

name, scale, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
cfg = {"C2": synth.config2, "C3": synth.config3}[name](scale=scale)
os.makedirs(out, exist_ok=True)
cfg["rows"].astype("<i8").tofile(os.path.join(out, "rows.i64"))
cfg["cols"].astype("<i8").tofile(os.path.join(out, "cols.i64"))
cfg["vals"].astype("<f8").tofile(os.path.join(out, "vals.f64"))
np.asfortranarray(cfg["X0"]).ravel(order="F").astype("<f8").tofile(os.path.join(out, "X0.f64"))
np.asfortranarray(cfg["Y0"]).ravel(order="F").astype("<f8").tofile(os.path.join(out, "Y0.f64"))
open(os.path.join(out, "meta.txt"), "w").write(f"{cfg['m']} {cfg['n']} {cfg['k']} {len(cfg['vals'])}\n")
print("wrote", out)
