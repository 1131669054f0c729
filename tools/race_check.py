"""Determinism check: repeated device-path runs must agree bitwise with each other and with the
single-stream (profiling) schedule."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from starfish_b200 import synth
from starfish_b200.engine import LikelihoodEngine

N, B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, int(sys.argv[2]) if len(sys.argv) > 2 else 64
ws = int(sys.argv[3]) if len(sys.argv) > 3 else 0
d = synth.stage_inputs_direct(N, B)
eng = LikelihoodEngine(N, 6, 2, B, workspace_walkers=ws)
eng.set_data(d["wave"], d["sigma"], d["data_flux"])
print("slots", eng.workspace_walkers)
eng.profile(True)
ref, info = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"], loc=d["loc"])
ref = ref.cpu().numpy(); eng.profile_read(); eng.profile(False)
bad = 0
for it in range(6):
    out, info = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"], loc=d["loc"])
    out = out.cpu().numpy()
    diff = np.flatnonzero(out != ref)
    print(it, "mismatches", diff.size, (np.abs(out - ref) / np.abs(ref)).max(), diff[:10])
    bad += diff.size
print("RACE" if bad else "deterministic")
