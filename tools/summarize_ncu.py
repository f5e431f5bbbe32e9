"""Summarise ncu outputs brought back from the GPU box into small text files under profiles/.
  python tools/summarize_ncu.py launches gpurun_out/launches_r01.csv > profiles/ncu_launches_r01.txt
  python tools/summarize_ncu.py details gpurun_out/prof.ncu-rep > profiles/ncu_details_r01.txt"""
import collections
import csv
import re
import subprocess
import sys

mode, path = sys.argv[1], sys.argv[2]
if mode == "launches":
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            agg[re.sub(r"\(.*", "", r[ik])[:100]].append(float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0))
        except ValueError:
            pass
    tot = sum(sum(v) for k, v in agg.items() if "k_fma_peak" not in k)
    print("# gpu__time_duration.sum per launch (ncu --clock-control none; cold-cache, serialised: compare SHARES, not absolutes)")
    print("# share excludes the one-off k_fma_peak roofline probe")
    print("%-102s %6s %11s %9s %7s" % ("kernel", "count", "total_us", "mean_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        share = "" if "k_fma_peak" in k else "%5.1f%%" % (100 * sum(v) / tot)
        print("%-102s %6d %11.1f %9.2f %7s" % (k, len(v), sum(v), sum(v) / len(v), share))
else:
    out = subprocess.run(["ncu", "-i", path, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    keep = ("GPU Speed Of Light Throughput", "Launch Statistics", "Occupancy", "Compute Workload Analysis", "Memory Workload Analysis",
            "Warp State Statistics", "Scheduler Statistics")
    last = None
    for r in rows[1:]:
        if r[idx["Section Name"]] in keep and r[idx["Metric Name"]]:
            if r[idx["ID"]] != last:
                last = r[idx["ID"]]
                print("\n== launch %s: %s" % (last, r[idx["Kernel Name"]][:110]))
            print("%-30s %-46s %-16s %s" % (r[idx["Section Name"]][:30], r[idx["Metric Name"]][:46], r[idx["Metric Unit"]], r[idx["Metric Value"]]))
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h2 = rr[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
            "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
            "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__inst_executed.sum"]
    print("\n== raw counters")
    for w in want:
        if w in h2:
            i = h2.index(w)
            print("%-66s %-10s %s" % (w, rr[1][i], [r[i][:60] for r in rr[2:]]))
