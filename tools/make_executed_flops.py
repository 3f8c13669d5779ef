"""profiles/executed_flops.json from an ncu metrics CSV (thread-level FP32 op counts of pv_state_bits_kernel).

    ncu --metrics smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,\
smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,gpu__time_duration.sum \
        --clock-control none -k regex:pv_state_bits -s 2 -c 1 --csv --log-file gpurun_out/fpops.csv python tools/prof_state.py
    python tools/make_executed_flops.py gpurun_out/fpops.csv 1048576
"""
import csv, json, os, sys
path, n = sys.argv[1], int(sys.argv[2])
vals = {}
for row in csv.reader(open(path)):
    if len(row) > 14 and row[12].startswith(("smsp__", "gpu__")):
        vals[row[12]] = float(row[14].replace(",", ""))
ffma = vals["smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"]
fadd = vals["smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"]
fmul = vals["smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"]
out = {
    "kernel": "pv_state_bits_sorted_kernel<SoA>: per-lane culling + scene-level cull, each block visits its share in order of (wrist distance class, elbow bin, wrist-flex bin), 512-thread lockstep blocks",
    "workload": f"{n} random Panda configs vs goal1_scattered (bench.py workload)",
    "ffma_per_check": ffma / n, "fadd_per_check": fadd / n, "fmul_per_check": fmul / n,
    "fp32_flops_per_check": (2 * ffma + fadd + fmul) / n,
    "thread_inst_per_check": vals.get("smsp__thread_inst_executed.sum", 0) / n,
    "warp_inst_per_32_checks": vals.get("smsp__inst_executed.sum", 0) / (n / 32),
    "ncu_duration_us": vals.get("gpu__time_duration.sum", 0) / 1e3,
    "source": os.path.basename(path),
    "note": "FMNMX/FSETP/abs/compare work is not counted as FLOPs; FFMA counts 2",
}
if len(sys.argv) > 3:
    # raw page of a `--set full` capture of the same launch: DRAM traffic
    rows = list(csv.reader(open(sys.argv[3])))
    hdr, unit, val = rows[0], rows[1], rows[2]
    def grab(name):
        i = hdr.index(name)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit[i]]
        return float(val[i]) * scale
    out["dram_bytes_per_launch"] = grab("dram__bytes_read.sum") + grab("dram__bytes_write.sum")
    out["algorithmic_bytes_per_launch"] = n * (32 + 0.125)
    out["dram_source"] = os.path.basename(sys.argv[3])
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "executed_flops.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
