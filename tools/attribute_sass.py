"""Where do the executed instructions of the state kernel go?  Joins the per-SASS-instruction "Instructions Executed"
column of an ncu report (--page source --print-source sass) with the source lines nvdisasm attributes to the same
instructions of the in-tree library (compiled with -lineinfo), and sums them per section of pv_check_config.

    python tools/attribute_sass.py gpurun_out/prof_state.ncu-rep [kernel-symbol-substring]

The report must come from the same build as csrc/libpanda_validity.so (instruction counts are compared).  Developer tool."""
import csv, os, re, subprocess, sys, collections, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
sym = sys.argv[2] if len(sys.argv) > 2 else "pv_state_bits_sorted_kernelILi0ELb0E"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "rbe550_final_project_b200/csrc/libpanda_validity.so")], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
sass_txt = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, "pv_kernels.sm_100a.cubin")], capture_output=True, text=True).stdout
page_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
# 1. parse nvdisasm section
lines=sass_txt.split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('.text.') and sym in l][0]
ins=[]  # (opcode text, frames)
frames=[]; pending=[]
for l in lines[start+1:]:
    if l.startswith('.text.') or l.startswith('.section'): break
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?',l)
    if m:
        pending.append((m.group(1).split('/')[-1],int(m.group(2))))
        continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m:
        if pending: frames=pending; pending=[]
        ins.append((m.group(2).strip(),list(frames)))
print(len(ins),'instructions in cubin')
# 2. ncu sass page
import io
rows=list(csv.reader(io.StringIO(page_csv)))
hdr_i=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hdr_i]; body=rows[hdr_i+1:]
ci=hdr.index('Instructions Executed'); si=hdr.index('Source'); ti=hdr.index('Thread Instructions Executed')
print(len(body),'instructions in ncu page')
def opc(s): 
    s=re.sub(r'@!?U?P\d+\s+','',s.strip()); return s.split()[0]
mism=sum(1 for a,b in zip(ins,body) if opc(a[0])!=opc(b[si]))
print('opcode mismatches',mism)
json_out={}
# 3. attribute
src=open(os.path.join(ROOT,'rbe550_final_project_b200/csrc/pv_device.cuh')).read().split('\n')
def find(txt): return [i+1 for i,l in enumerate(src) if txt in l][0]
marks=[('junk/limits',find('bool junk = false')),('fk+place',find('float3 s[PV_N_SPHERES];')),('table',find('robot vs ground plane')),
       ('carry-place',find('carried box: placed by the hand')),('self:grip_r',find('radius of a ball centred on the hand box')),('self:ss',find('if (S.flags & PV_FLAG_SELF)')),
       ('self:sbh',find('sphere-vs-gripper pairs: the three')),('env:obb-load',find('const int nb = S.n_obb;')),('env:group-macros',find('one uniform yaw / general decision per GROUP')),
       ('env:groups',find('PV_LINK_GROUPS(PV_ENV_GROUP)')),('env:gripper',find('one bounding ball around the whole gripper')),('end',find('#undef PV_EARLY_EXIT'))]
fk_lo,fk_hi=find('template <bool FAST = false, class F>'),find('// ---- the state check')
prim_lo,prim_hi=find('// ---- primitive tests'),find('// ---- forward kinematics')
def cat(fr):
    # fr: innermost first
    # find the frame that's inside pv_check_config in pv_device.cuh
    names=[]
    for f,l in fr:
        if f=='pv_device.cuh':
            sec=None
            for name,ln in marks:
                if l>=ln: sec=name
            if l>=marks[0][1] and l<marks[-1][1]: return sec
    for f,l in fr:
        if f=='pv_kernels.cu': return 'kernel-shell'
    return 'other'
tot=collections.Counter(); tth=collections.Counter(); n_static=collections.Counter()
for (a,fr),b in zip(ins,body):
    c=cat(fr)
    # sub-categorise env:groups into cull vs test by innermost function
    if c=='env:groups':
        inner=fr[0]
        if inner[0]=='pv_device.cuh' and prim_lo<=inner[1]<prim_hi: c='env:groups:sphere-tests'
        else: c='env:groups:culls'
    tot[c]+=int(b[ci]); tth[c]+=int(b[ti]); n_static[c]+=1
T=sum(tot.values())
for k,v in sorted(tot.items(), key=lambda kv:-kv[1]):
    print(f"{k:28s} static {n_static[k]:5d}  warp-inst {v:10d} {100*v/T:5.1f}%  lanes/inst {tth[k]/max(v,1):5.1f}")
print('total',T)
if os.environ.get("PV_ATTR_LINES"):
    per = collections.Counter()
    for (a, fr), b in zip(ins, body):
        if cat(fr) == os.environ["PV_ATTR_LINES"]:
            per[tuple(fr[:3])] += int(b[ci])
    for k, v in per.most_common(12):
        print(v, k)
