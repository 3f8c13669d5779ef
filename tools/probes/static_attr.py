"""Static SASS size of one kernel per section of pv_check_config (nvdisasm line info of an object file).  Developer probe.
usage: python tools/probes/static_attr.py <obj-or-cubin> <kernel-symbol-substring>"""
import re, subprocess, sys, collections, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
obj, sym = sys.argv[1], sys.argv[2]
txt = subprocess.run(["nvdisasm", "--print-line-info-inline", obj], capture_output=True, text=True).stdout
if not txt:
    import tempfile
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
lines = txt.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and sym in l][0]
src = open(os.path.join(ROOT, "rbe550_final_project_b200/csrc/pv_device.cuh")).read().split("\n")
def find(t): return [i + 1 for i, l in enumerate(src) if t in l][0]
marks = [("limits", find("Joint limits are part of the model")), ("place-call", find("PvPlaced P;\n") if False else find("    PvPlaced P;")),
         ("plane", find("robot vs ground plane")), ("carry", find("carried box: placed by the hand, checked")),
         ("self:ss-blocks", find("#define PV_SS(a, b, rr2, rr)")),
         ("self:sbh-blocks", find("sphere-vs-gripper pairs: the three")), ("scene-level", find("// ---- robot vs scene boxes ----")),
         ("end", find("#undef PV_EARLY_EXIT\n") if False else find("#undef PV_LOCKSTEP"))]
sec_lo, sec_hi = find("__device__ __forceinline__ void pv_scene_section("), find("// The scene section OUT OF LINE")
fk_lo, fk_hi = find("template <bool FAST = false, class F>"), find("// ---- the state check")
pl_lo, pl_hi = find("// FK -> sphere centres + gripper boxes"), find("#define PV_EARLY_EXIT_RET")
def cat(fr):
    for f, l in fr:
        if f == "pv_device.cuh":
            if sec_lo <= l < sec_hi: return "scene-section"
    for f, l in fr:
        if f == "pv_device.cuh":
            if fk_lo <= l < fk_hi: return "fk"
            if pl_lo <= l < pl_hi: return "place"
    for f, l in fr:
        if f == "pv_device.cuh" and marks[0][1] <= l < marks[-1][1]:
            sec = None
            for name, ln in marks:
                if l >= ln: sec = name
            return sec
    for f, l in fr:
        if f == "panda_model_gen.h": return "gen.h(?)"
    return "kernel-shell/other"
n = collections.Counter(); frames = []; pending = []
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2)))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+", l):
        if pending: frames = pending; pending = []
        n[cat(frames)] += 1
T = sum(n.values())
for k, v in sorted(n.items(), key=lambda kv: -kv[1]): print(f"{k:24s} {v:5d} instr {v * 16 / 1024:6.1f} KB")
print("total", T, T * 16 / 1024, "KB")
if len(sys.argv) > 3:  # run-length layout: category sequence in address order
    seq = []; frames = []; pending = []
    for l in lines[start + 1:]:
        if l.startswith(".text.") or l.startswith(".section"): break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            pending.append((m.group(1).split("/")[-1], int(m.group(2)))); continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+", l):
            if pending: frames = pending; pending = []
            c = cat(frames)
            if seq and seq[-1][0] == c: seq[-1][1] += 1
            else: seq.append([c, 1])
    pos = 0
    for c, k in seq:
        if k >= 12: print(f"  @{pos * 16 / 1024:5.1f} KB  {c:24s} {k:5d}")
        pos += k
