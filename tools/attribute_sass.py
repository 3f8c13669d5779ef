"""Where do the executed instructions of a kernel go?  Joins the per-SASS-instruction "Instructions Executed" column of an
ncu report (--page source --print-source sass) with the source lines nvdisasm attributes to the same instructions of the
in-tree library (compiled with -lineinfo) and sums them (a) per innermost device function and (b) per `// ----` section
of the file the outermost non-kernel frame lies in.

    python tools/attribute_sass.py gpurun_out/prof_state.ncu-rep [kernel-symbol-substring] [cubin-name]

The report must come from the same build as csrc/libpanda_validity.so (opcodes are compared).  Developer tool."""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "rbe550_final_project_b200/csrc")
rep = sys.argv[1]
sym = sys.argv[2] if len(sys.argv) > 2 else "pv_state_bits_sorted_kernelILi0ELb0ELb1E"
cubin = sys.argv[3] if len(sys.argv) > 3 else "pv_kernels"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(CSRC, "libpanda_validity.so")], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if cubin in f][0]
sass_txt = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
page_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = sass_txt.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and sym in l][0]
ins, frames, pending = [], [], []
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        if pending:
            frames, pending = pending, []
        ins.append((m.group(2).strip(), list(frames)))
rows = list(csv.reader(io.StringIO(page_csv)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr, body = rows[hdr_i], rows[hdr_i + 1:]
ci, si, ti = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Thread Instructions Executed")
opc = lambda s: re.sub(r"@!?U?P\d+\s+", "", s.strip()).split()[0]
print(len(ins), "instructions in cubin,", len(body), "in the ncu page, opcode mismatches",
      sum(1 for a, b in zip(ins, body) if opc(a[0]) != opc(b[si])))
src = {}
def text(f):
    if f not in src:
        p = os.path.join(CSRC, f)
        src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    return src[f]
def func_of(f, ln):
    t = text(f)
    for i in range(min(ln, len(t)) - 1, -1, -1):
        m = re.match(r"\s*(?:static\s+)?(?:__device__|__global__|template).*?\b(pv_\w+|v_\w+)\s*\(", t[i])
        if m and ("__device__" in t[i] or "__global__" in t[i] or (i + 1 < len(t) and "__device__" in t[i + 1])):
            return m.group(1)
        m = re.match(r"\s*(pv_\w+)\(const __grid_constant__", t[i])
        if m:
            return m.group(1)
    return f
def section_of(f, ln):
    t = text(f)
    for i in range(min(ln, len(t)) - 1, -1, -1):
        m = re.match(r"\s*// ----+ ?(.*?)[- ]*$", t[i])
        if m and m.group(1):
            return f"{f}: {m.group(1)[:60]}"
    return f
by_fn, by_sec, th_fn, th_sec, st_fn = (collections.Counter() for _ in range(5))
for (a, fr), b in zip(ins, body):
    n, th = int(b[ci]), int(b[ti])
    inner = fr[0] if fr else ("?", 0)
    outer = next((x for x in reversed(fr) if x[0] != cubin + ".cu"), fr[-1] if fr else ("?", 0))
    k1, k2 = func_of(*inner), section_of(*outer)
    by_fn[k1] += n; th_fn[k1] += th; st_fn[k1] += 1
    by_sec[k2] += n; th_sec[k2] += th
T = sum(by_fn.values())
print("--- per innermost function")
for k, v in by_fn.most_common(25):
    print(f"{k:34s} static {st_fn[k]:5d}  warp-inst {v:11d} {100 * v / T:5.1f}%  lanes/inst {th_fn[k] / max(v, 1):5.1f}")
print("--- per section (outermost non-kernel frame)")
for k, v in by_sec.most_common(25):
    print(f"{k:80s} {v:11d} {100 * v / T:5.1f}%  lanes/inst {th_sec[k] / max(v, 1):5.1f}")
print("total", T)
