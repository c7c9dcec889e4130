"""Turns gpurun_out/*.ncu-rep / launch-list csv into the small text summaries committed under profiles/.

    python scripts/summarize_ncu.py rep  gpurun_out/x.ncu-rep  profiles/out.txt
    python scripts/summarize_ncu.py list gpurun_out/launches.csv profiles/out.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum ",
        "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor", "sm__inst_executed_pipe_tensor",
        "sm__inst_executed_pipe_fmaheavy", "sm__inst_executed_pipe_fmalite", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum ", "l1tex__t_bytes.sum ", "smsp__cycles_active.avg ", "sm__cycles_elapsed.max"]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none, summary of {path}\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"\n== {name}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}\n")
            for i, h in enumerate(hdr):
                if any(h.startswith(k.strip()) if k.endswith(" ") else k in h for k in KEEP) and r[i] not in ("", "0"):
                    if ".max" in h or ".min" in h or (".sum." in h and "per_second" not in h):
                        continue
                    f.write(f"{h:95s} {units[i]:12s} {r[i]}\n")


def lst(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        name = r[kn].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none launch list of {path}\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':72s} {'launches':>8s} {'total_ms':>10s} {'share':>8s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:72]:72s} {v[0]:8d} {v[1] / 1e6:10.3f} {v[1] / tot * 100:7.2f}%\n")


if __name__ == "__main__":
    {"rep": rep, "list": lst}[sys.argv[1]](sys.argv[2], sys.argv[3])
