"""Turn gpurun_out/r01_*.ncu-rep + r01_launches.csv into the small text/JSON summaries that are
committed under profiles/ (the .ncu-rep files themselves stay in gpurun_out/, scratch)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "sm__cycles_elapsed.avg.per_second",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2]


def main():
    os.makedirs(DST, exist_ok=True)
    traffic = {}
    lines = [f"# ncu --set full --clock-control none, one launch each, bench.py --steps 2 --warmup 3 ({TAG})", ""]
    for f in sorted(os.listdir(SRC)):
        if not (f.startswith(TAG + "_") and f.endswith(".ncu-rep")):
            continue
        hdr, units, vals = raw(os.path.join(SRC, f))
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else f
        lines.append(f"## {f[len(TAG) + 1:-8]}  ({name[:90]})")
        rec = {}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"  {k:88s} {vals[i]:>16s} {units[i]}")
                rec[k] = (vals[i], units[i])
        lines.append("")

        def to_bytes(key):
            v, u = rec.get(key, ("0", "byte"))
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            return float(v.replace(",", "")) * mult
        traffic[f[len(TAG) + 1:-8]] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    with open(os.path.join(DST, f"{TAG}_ncu_summary.txt"), "w") as fh:
        fh.write("\n".join(lines))
    # launch list -> per-kernel totals and shares
    lc = os.path.join(SRC, f"{TAG}_launches.csv")
    if os.path.exists(lc):
        with open(lc) as fh:
            body = [l for l in fh if not l.startswith("==")]
        agg = collections.defaultdict(list)
        for row in csv.DictReader(body):
            try:
                agg[row["Kernel Name"]].append(float(row["Metric Value"]))
            except (KeyError, ValueError):
                pass
        tot = sum(sum(v) for v in agg.values())
        with open(os.path.join(DST, f"{TAG}_launches_summary.txt"), "w") as fh:
            fh.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 2 --warmup 3 ({TAG}); "
                     "cold-cache serialised launches: compare SHARES\n")
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
                fh.write(f"{sum(v) / tot * 100:6.2f}%  n={len(v):4d}  total={sum(v) / 1e6:9.3f} ms  avg={sum(v) / len(v) / 1e3:9.1f} us  {k[:110]}\n")
    with open(os.path.join(DST, "roofline_traffic.json"), "w") as fh:
        json.dump({"msda_bwd": traffic.get("msda_backward"), "msda_fwd": traffic.get("msda_forward"),
                   "all_dram_bytes_per_launch": traffic, "source": f"profiles/{TAG}_ncu_summary.txt"}, fh, indent=1)
    print(open(os.path.join(DST, f"{TAG}_launches_summary.txt")).read()[:1800])


if __name__ == "__main__":
    main()
