#!/usr/bin/env python
"""Turn one gpurun round's ncu outputs into the tracked summaries under profiles/.

  python tools/ncu_summary.py <tag>

Reads gpurun_out/launches_<tag>.csv (ncu --metrics gpu__time_duration.sum launch
list), gpurun_out/prof_<tag>.ncu-rep (ncu --set full capture), bench_<tag>.json and
table_<tag>.json; writes profiles/<tag>_launches_by_kernel.csv, profiles/<tag>_launches.csv
(the raw launch list), profiles/<tag>_ncu_full.csv (selected counters per captured launch)
and copies the bench line / kernel table.
"""
import collections
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

WANT = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum",
]


def main():
    tag = sys.argv[1]
    os.makedirs(P, exist_ok=True)
    lp = os.path.join(G, f"launches_{tag}.csv")
    if os.path.exists(lp):
        lines = [l for l in open(lp) if l.startswith('"')]
        with open(os.path.join(P, f"{tag}_launches.csv"), "w") as f:
            f.writelines(lines)
        agg = collections.OrderedDict()
        for row in csv.DictReader(lines):
            v = float(row["Metric Value"].replace(",", ""))
            v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
            a = agg.setdefault(row["Kernel Name"], [0, 0.0, row["Grid Size"], row["Block Size"]])
            a[0] += 1
            a[1] += v
        tot = sum(a[1] for a in agg.values())
        with open(os.path.join(P, f"{tag}_launches_by_kernel.csv"), "w") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "launches", "total_us", "avg_us", "share_pct", "grid(last)", "block(last)"])
            for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                w.writerow([k, a[0], f"{a[1] / 1e3:.1f}", f"{a[1] / a[0] / 1e3:.2f}",
                            f"{100 * a[1] / tot:.2f}", a[2], a[3]])
    rp = os.path.join(G, f"prof_{tag}.ncu-rep")
    if os.path.exists(rp):
        raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, data = rows[0], rows[1], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        cols = [c for c in WANT if c in idx]
        with open(os.path.join(P, f"{tag}_ncu_full.csv"), "w") as f:
            w = csv.writer(f)
            w.writerow([f"{c} [{units[idx[c]]}]" if units[idx[c]] else c for c in cols])
            for d in data:
                w.writerow([d[idx[c]] for c in cols])
    # dram traffic per launch of the step's kernels (CIFAR net) -> profiles/traffic.json,
    # read by bench.py for roofline.traffic
    full = os.path.join(P, f"{tag}_ncu_full.csv")
    if os.path.exists(full):
        import json
        import re
        rows = list(csv.DictReader(open(full)))
        FW, BW = "evaluate+activation+pool (tcgen05)", "input_delta+activation+pool backward (tcgen05)"
        label_of = [
            (r"conv_wgrad_mma_rgb_kernel", "convolution_layer_1:weight_update"),
            (r"conv_wgrad_rows_kernel", "convolution_layer_1:activation+pool backward+weight_update"),
            (r"conv_rows_kernel", "convolution_layer_1:" + FW),
            (r"conv_wgrad_umma_kernel<unnamed>::GCfg<16, 20, 16,", "convolution_layer_4:weight_update"),
            (r"conv_wgrad_umma_kernel<unnamed>::GCfg<20, 20, 8,", "convolution_layer_7:weight_update"),
            (r"UCfg<3, 16, 32, 32, 5, 0, 0>, 2>", "convolution_layer_1:" + FW),
            (r"TCfg<16, 20, 16, 4, 0>, 2>", "convolution_layer_4:" + FW),
            (r"TCfg<20, 20, 8, 4, 0>, 2>", "convolution_layer_7:" + FW),
            (r"UCfg<16, 3, 32, 32, 5, 1, 1>, 2>", "convolution_layer_1:" + BW),
            (r"TCfg<20, 16, 16, 4, 1>, 2>", "convolution_layer_4:" + BW),
            (r"TCfg<20, 20, 8, 4, 1>, 2>", "convolution_layer_7:" + BW),
        ]
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        traffic = {}
        for r in rows:
            name = r["Kernel Name"]
            total = 0.0
            for col, v in r.items():
                m = re.match(r"dram__bytes_(read|write)\.sum \[(\w+)\]", col)
                if m and v:
                    total += float(v.replace(",", "")) * unit.get(m.group(2), 1.0)
            for pat, label in label_of:
                if pat in name.replace("(anonymous namespace)", "<unnamed>"):
                    traffic[label] = total
        if traffic:
            traffic["_source"] = f"ncu --set full capture {tag} (dram__bytes_read.sum + dram__bytes_write.sum per launch)"
            json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
            json.dump(traffic, open(os.path.join(P, f"{tag}_traffic.json"), "w"), indent=1)
    # the tensor-core GEMM configurations (wide MLP, YOLOv1): full captures + bench lines
    for kind in ("dense", "yolo"):
        rp = os.path.join(G, f"prof_{kind}_{tag}.ncu-rep")
        if not os.path.exists(rp):
            continue
        raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, data = rows[0], rows[1], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        cols = [c for c in WANT + ["lts__t_sectors_srcunit_tex_op_read.sum",
                                   "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
                                   "sm__cycles_elapsed.max"] if c in idx]
        with open(os.path.join(P, f"{tag}_{kind}_ncu_full.csv"), "w") as f:
            w = csv.writer(f)
            w.writerow([f"{c} [{units[idx[c]]}]" if units[idx[c]] else c for c in cols])
            for d in data:
                w.writerow([d[idx[c]] for c in cols])
    for name in (f"mlp_{tag}.json", f"yolo_{tag}.json"):
        src = os.path.join(G, name)
        if os.path.exists(src):
            lines = [l for l in open(src) if l.startswith("{")]
            if lines:
                open(os.path.join(P, f"{tag}_{name.split('_')[0]}.json"), "w").write(lines[-1])
    for name in (f"bench_{tag}.json", f"table_{tag}.json"):
        src = os.path.join(G, name)
        if os.path.exists(src):
            shutil.copy(src, os.path.join(P, f"{tag}_{name.split('_')[0]}.json"))


if __name__ == "__main__":
    main()
