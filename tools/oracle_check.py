"""Debug helper: CUDA path vs CPU oracle per walker (dev tool, not shipped logic)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import starfish_oracle as O
from starfish_b200 import synth
from starfish_b200.engine import LikelihoodEngine

N, B = int(sys.argv[1]), int(sys.argv[2])
ws = int(sys.argv[3]) if len(sys.argv) > 3 else 0
d = synth.stage_inputs_direct(N, B)
eng = LikelihoodEngine(N, 6, 2, B, workspace_walkers=ws)
eng.set_data(d["wave"], d["sigma"], d["data_flux"])
print("slots", eng.workspace_walkers)
ref = np.empty(B)
for b in range(B):
    cov = O.assemble_covariance(d["wave"], d["sigma"], None, None, d["glob"][b], d["loc"][b])
    cov += d["X"][b].T @ d["A"][b] @ d["X"][b]
    ref[b] = O.log_likelihood(cov, d["model_flux"][b], d["data_flux"])[0]
for it in range(3):
    out, info = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"], loc=d["loc"])
    out = out.cpu().numpy(); info = info.cpu().numpy()
    err = np.abs(out - ref) / np.abs(ref)
    bad = np.flatnonzero(~(err < 1e-10))
    print(it, "bad", bad.size, "info!=0:", np.flatnonzero(info != 0)[:10], "first bad:", bad[:16], "max err", np.nanmax(err))
