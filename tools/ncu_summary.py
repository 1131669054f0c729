"""Summarise ncu outputs brought back in gpurun_out/ (run here, no GPU needed).
   python tools/ncu_summary.py launches <launches.csv>
   python tools/ncu_summary.py raw <file.ncu-rep>"""
import collections, csv, subprocess, sys

def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("sfb::<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total device time {tot/1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:58]:58s} n={v[0]:5d} total={v[1]/1e6:10.3f} ms share={v[1]/tot*100:5.1f}% avg={v[1]/v[0]/1e3:9.1f} us")

WANT = ["launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]

def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    for r in body:
        print("==", r[ki].split("(")[0].replace("sfb::<unnamed>::", ""))
        for w in WANT:
            if w in hdr:
                print(f"   {w:75s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
        st = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.25:
                    st.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        print("   stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)))

if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
