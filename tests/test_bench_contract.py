"""CPU: the reference arm of bench.py prints ONE JSON line with the keys the driver reads (the b200 arm needs a GPU;
its line is checked on the box and kept under profiles/)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--n-pix", "512", "--walkers", "4", "--cpu-sample", "2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["scaling"] in ("weak", "strong")
    assert d["value"] > 0 and d["ms_per_step"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


def test_committed_b200_line_has_the_contract_keys():
    """A default `python bench.py` line measured on a B200 at the end of round 2 (profiles/r4i_bench.json)."""
    d = json.load(open(os.path.join(ROOT, "profiles", "r4i_bench.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["parity"]["ok"] and d["parity"]["max_rel_vs_dense_oracle"] <= 1e-10 and d["parity"]["walkers"] >= 8
    assert set(d["configs"]) == {"configs[1]", "configs[4]"} and d["e2e"]["steps"] >= 10
    assert d["frozen_shared"]["max_rel_diff_vs_per_walker_factorisation"] <= 1e-10
    assert d["roofline"]["traffic"] and 0 < d["roofline"]["int8_mma_issued_frac"] <= 1
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1


def test_no_collective_inside_rank0_only_blocks():
    """bench.py under torchrun: a step function that ends in the lnL all-gather must never be called from a block only
    rank 0 executes (rank 0 would wait for the others for ever — seen once with the frozen-groups leg at 2 GPUs)."""
    src = open(os.path.join(ROOT, "bench.py")).read().splitlines()
    collective = ("step_frozen(", "step_device(", "step_model(", "gather(", "allgather_lnl(", "gather_lnl(")
    i = 0
    found = []
    while i < len(src):
        line = src[i]
        stripped = line.lstrip()
        if stripped.startswith("if rank == 0"):
            indent = len(line) - len(stripped)
            j = i + 1
            while j < len(src) and (not src[j].strip() or len(src[j]) - len(src[j].lstrip()) > indent):
                code = src[j].split("#")[0]
                if any(c in code for c in collective) and "def " not in code:
                    found.append((j + 1, src[j].strip()))
                j += 1
            i = j
        else:
            i += 1
    assert not found, found
