"""Runs bench.py (rotating batches, the contract workload) on every alternative build csrc/libpv_*.so. Developer tool.
usage: python tools/variant_bench_main.py [bench.py arguments]"""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:] or ["--steps", "1000", "--warmup", "20", "--no-plan", "--no-cpu-baseline"]
for lib in sorted(glob.glob(os.path.join(ROOT, "rbe550_final_project_b200", "csrc", "libpv_*.so"))):
    code = ("import sys, runpy; sys.path.insert(0, %r); from rbe550_final_project_b200 import _cabi; _cabi.LIB_PATH = %r; "
            "sys.argv = ['bench.py'] + %r; runpy.run_path(%r, run_name='__main__')" % (ROOT, lib, args, os.path.join(ROOT, "bench.py")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(os.path.basename(lib), "FAILED", out.stderr[-300:])
        continue
    d = json.loads(line[-1])
    print(f"{os.path.basename(lib):24s} {d['value'] / 1e9:7.3f} G checks/s  {d['ms_per_step'] * 1e3:7.2f} us/step  e2e {d['e2e']['value'] / 1e9:.3f} G/s")
