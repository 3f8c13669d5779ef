"""Build csrc/libpv_<name>.so with extra nvcc -D flags, for same-box A/B runs (tools/variant_bench*.py).
usage: python tools/build_variant.py <name> [-DFLAG[=v] ...]"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rbe550_final_project_b200 import _cabi, panda_model

name, flags = sys.argv[1], sys.argv[2:]
panda_model.write_header()
nvcc = _cabi._nvcc()
objs, procs = [], []
for src in _cabi.SOURCES:
    obj = os.path.join("/tmp", f"pvvar_{name}_{src.replace('.cu', '.o')}")
    procs.append(subprocess.Popen([nvcc] + _cabi.NVCC_FLAGS + flags + ["-c", src, "-o", obj], cwd=_cabi.CSRC))
    objs.append(obj)
assert all(p.wait() == 0 for p in procs)
out = os.path.join(_cabi.CSRC, f"libpv_{name}.so")
subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + ["-lcudart"], check=True)
print(out)
