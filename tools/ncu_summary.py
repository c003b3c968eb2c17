"""Summarise ncu outputs into text for profiles/ (run here, no GPU needed).
   python tools/ncu_summary.py launches gpurun_out/launches_r01.csv
   python tools/ncu_summary.py kernel gpurun_out/prof_conv.ncu-rep"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed.sum.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__inst_executed_pipe_uniform.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, data = rows[0], rows[1:]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = {}, 0.0
    for r in data:
        name = r[ik].split("(")[0].replace("rvsr::", "")
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] == "ns" else (v * 1e3 if r[iu] == "ms" else v)
        a = agg.setdefault(name, [0.0, 0])
        a[0] += v; a[1] += 1; tot += v
    print("total %.1f us over %d launches (cold-cache, serialised: compare shares)" % (tot, len(data)))
    for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("%-60s %9.1f us  n=%3d  %5.1f%%" % (k, v, n, 100 * v / tot))


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, vals = rows[0], rows[-1]
    print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for i, h in enumerate(hdr):
        if h in KEYS:
            print("  %-70s %s %s" % (h, vals[i], rows[1][i]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = next(r for r in rows if "Source" in r and "Address" in r)
    data = rows[rows.index(hdr) + 1:]
    isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    tot_ex = sum(int(r[iex] or 0) for r in data); tot_s = sum(int(r[ismp] or 0) for r in data)
    print("  warp instructions executed: %d   stall samples: %d" % (tot_ex, tot_s))
    op, ops = Counter(), Counter()
    for r in data:
        m = r[isrc].split()
        if not m:
            continue
        name = (m[1] if m[0].startswith("@") else m[0]).split(".")[0]
        op[name] += int(r[iex] or 0); ops[name] += int(r[ismp] or 0)
    for k, v in op.most_common(14):
        print("    %-10s exec %5.1f%%   stall samples %5.1f%%" % (k, 100 * v / max(tot_ex, 1), 100 * ops[k] / max(tot_s, 1)))
    st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = Counter()
    for r in data:
        for i in st:
            try:
                agg[hdr[i]] += int(r[i] or 0)
            except ValueError:
                pass
    print("  stall reasons:", ", ".join("%s %.0f%%" % (k, 100 * v / max(tot_s, 1)) for k, v in agg.most_common(8)))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
