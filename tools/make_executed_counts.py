"""profiles/executed_counts.json from ncu metric CSVs: what the dominant kernels EXECUTE per unit of work, tied to the
hash of the sources they were compiled from (bench.py prints `executed_counts_stale` when the hash no longer matches).

One gpurun call produces the inputs (see profiles/README or DESIGN.md section 4):
    M=smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,\
smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,\
dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
    ncu --metrics $M --clock-control none -k regex:pv_state_bits_sorted -s 2 -c 1 --csv --log-file gpurun_out/state_counts.csv python tools/prof_state.py
    ncu --metrics $M --clock-control none -k regex:pv_edge -s 8 -c 4 --csv --log-file gpurun_out/edge_counts.csv python tools/prof_edge.py
(one pv_check_edges call of the config-3 workload = four launches: pv_edge_cert_kernel, pv_edge_kernel<list, scene only>,
pv_edge_cert2_kernel, pv_edge_kernel<list, self-collision only>; their counts are summed and kept per kernel;
tools/debug/validate_gpu.sh runs both commands)
then here:
    python tools/make_executed_counts.py gpurun_out/state_counts.csv gpurun_out/edge_counts.csv [--tag r2a]
(prof_state.py checks 1 048 576 configurations, prof_edge.py 1 048 576 edges x 64 states: the bench workloads.)
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (source_hash)


def metrics(path):
    vals = {}
    for row in csv.reader(open(path)):
        if len(row) > 14 and row[12].startswith(("smsp__", "gpu__", "dram__")):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0}.get(row[13], 1.0)
            vals[row[12]] = float(row[14].replace(",", "")) * scale
    return vals


def metrics_per_launch(path):
    """[(kernel name, {metric: value})] per profiled launch, in launch order"""
    launches = {}
    for row in csv.reader(open(path)):
        if len(row) > 14 and row[0].isdigit() and row[12].startswith(("smsp__", "gpu__", "dram__")):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0}.get(row[13], 1.0)
            name = row[4].split("(")[0].replace("void ", "")
            launches.setdefault(int(row[0]), (name, {}))[1][row[12]] = float(row[14].replace(",", "")) * scale
    return [launches[k] for k in sorted(launches)]


def fp(v, n):
    ffma = v["smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"]
    fadd = v["smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"]
    fmul = v["smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"]
    return {"ffma": ffma / n, "fadd": fadd / n, "fmul": fmul / n, "fp32_flops": (2 * ffma + fadd + fmul) / n}


args = [a for a in sys.argv[1:] if not a.startswith("--")]
tag = sys.argv[sys.argv.index("--tag") + 1] if "--tag" in sys.argv else ""
if "--tag" in sys.argv:
    args = [a for a in args if a != tag]
n = 1 << 20
out = {"source_hash": bench.source_hash(), "tag": tag,
       "note": "FMNMX / FSETP / abs / compare work is not counted as FLOPs; FFMA counts 2.  Counts are per launch of the "
               "bench workload under ncu (same kernel, same inputs as bench.py)"}
v = metrics(args[0])
f = fp(v, n)
out["state"] = {
    "kernel": "pv_state_bits_sorted_kernel<SoA, no carry>", "workload": f"{n} random Panda configs vs goal1_scattered",
    "ffma_per_check": f["ffma"], "fadd_per_check": f["fadd"], "fmul_per_check": f["fmul"],
    "fp32_flops_per_check": f["fp32_flops"],
    "thread_inst_per_check": v["smsp__thread_inst_executed.sum"] / n,
    "warp_inst_per_32_checks": v["smsp__inst_executed.sum"] / (n / 32),
    "ncu_duration_us": v.get("gpu__time_duration.sum", 0) / 1e3,
    "dram_bytes_per_launch": v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0),
    "algorithmic_bytes_per_launch": n * (32 + 0.125), "source": os.path.basename(args[0]),
}
if len(args) > 1:
    per = metrics_per_launch(args[1])
    v = {}
    for _, m in per:
        for k, x in m.items():
            v[k] = v.get(k, 0.0) + x
    f = fp(v, n)
    out["edges"] = {
        "kernel": " + ".join(name for name, _ in per), "workload": f"{n} edges x 64 states vs goal4_task1_pentagon (gaussian 0.3 pairs)",
        "per_kernel": [{"kernel": name, "warp_inst": m["smsp__inst_executed.sum"], "ncu_duration_us": m.get("gpu__time_duration.sum", 0) / 1e3}
                       for name, m in per],
        "warp_inst_per_edge": v["smsp__inst_executed.sum"] / n, "thread_inst_per_edge": v["smsp__thread_inst_executed.sum"] / n,
        "fp32_flops_per_edge": f["fp32_flops"], "ncu_duration_us": v.get("gpu__time_duration.sum", 0) / 1e3,
        "dram_bytes_per_launch": v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0),
        "algorithmic_bytes_per_launch": n * 64.125, "source": os.path.basename(args[1]),
    }
dst = os.path.join(ROOT, "profiles", "executed_counts.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
