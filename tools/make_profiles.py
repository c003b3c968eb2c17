"""Turn the ncu artefacts in gpurun_out/ into the tracked summaries under profiles/.
    python tools/make_profiles.py r01
Reads: gpurun_out/launches_<tag>.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_*),
       gpurun_out/launch_labels.json (ordered labels of one forward, same launch order),
       gpurun_out/prof_conv.ncu-rep, gpurun_out/prof_dcn.ncu-rep (ncu --set full).
Writes: profiles/<tag>_launches.txt, profiles/<tag>_ncu_conv3x3.txt, profiles/<tag>_ncu_dcn.txt,
        profiles/ncu_traffic.json (average DRAM bytes per launch per kernel class, used by bench.py)."""
import contextlib
import csv
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402


def main(tag):
    go, pr = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    rows = [r for r in csv.reader(open(os.path.join(go, "launches_%s.csv" % tag))) if len(r) > 5]
    hdr, data = rows[0], rows[1:]
    iid, ik, im, iv, iu = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    launches = {}
    for r in data:
        L = launches.setdefault(int(r[iid]), dict(kernel=r[ik].split("(")[0].replace("rvsr::", "")))
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        if r[im] == "gpu__time_duration.sum":
            L["us"] = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        else:
            scale = dict(byte=1, Kbyte=1e3, Mbyte=1e6, Gbyte=1e9).get(u, 1)
            L[r[im]] = v * scale
    # torch dtype copies and the one-time weight repacks precede the forward; keep the engine's own launches
    order = [launches[k] for k in sorted(launches)]
    order = [L for L in order if not L["kernel"].startswith("void at::") and "pack_weight" not in L["kernel"]
             and "pad_cin" not in L["kernel"] and "us" in L]
    labels = json.load(open(os.path.join(go, "launch_labels.json")))
    assert len(order) >= len(labels), (len(labels), len(order))
    order = order[:len(labels)]  # the first forward
    agg, tot = {}, 0.0
    for L, lab in zip(order, labels):
        key = ":".join(lab["label"].split(":")[:2])
        a = agg.setdefault(key, dict(us=0.0, n=0, dram=0.0, flops=0.0, bytes=0.0, kernel=L["kernel"]))
        a["us"] += L["us"]; a["n"] += 1; a["flops"] += lab["flops"]; a["bytes"] += lab["bytes"]
        a["dram"] += L.get("dram__bytes_read.sum", 0.0) + L.get("dram__bytes_write.sum", 0.0)
        tot += L["us"]
    out = io.StringIO()
    out.write("ncu launch list of ONE cfg2 forward at batch 4 (python tools/prof_forward.py 1 4), per kernel class.\n"
              "ncu times are cold-cache and serialised: compare SHARES with bench.py's `kernels`, not absolutes.\n"
              "dram = dram__bytes_read.sum + dram__bytes_write.sum; alg = algorithmic bytes (DESIGN.md).\n\n")
    out.write("%-28s %-28s %4s %10s %7s %12s %12s\n" % ("class", "kernel", "n", "us", "share", "dram MB/launch", "alg MB/launch"))
    traffic = {}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        out.write("%-28s %-28s %4d %10.1f %6.1f%% %12.1f %12.1f\n" % (k, a["kernel"][:28], a["n"], a["us"], 100 * a["us"] / tot,
                                                                     a["dram"] / a["n"] / 1e6, a["bytes"] / a["n"] / 1e6))
        traffic[k] = a["dram"] / a["n"]
    out.write("\ntotal %.1f us over %d launches\n" % (tot, len(order)))
    open(os.path.join(pr, "%s_launches.txt" % tag), "w").write(out.getvalue())
    json.dump(traffic, open(os.path.join(pr, "ncu_traffic.json"), "w"), indent=1)
    print(out.getvalue())
    for rep, name in (("prof_conv.ncu-rep", "ncu_conv3x3"), ("prof_dcn.ncu-rep", "ncu_dcn"), ("prof_tapn.ncu-rep", "ncu_conv_last_tapn"),
                      ("prof_fused.ncu-rep", "ncu_dcn_pack_fused")):
        path = os.path.join(go, rep)
        if os.path.exists(path):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                ncu_summary.kernel(path)
            open(os.path.join(pr, "%s_%s.txt" % (tag, name)), "w").write(
                "ncu --set full --clock-control none, one launch inside a batch-4 cfg2 forward\n" + buf.getvalue())
            print(buf.getvalue())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
