import os, sys, time, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for n in sorted(glob.glob("/sys/devices/system/node/node*")):
    print(n, open(n + "/cpulist").read().strip())
bus = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), "pci_bus_id") else None
import subprocess
q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.split()
print("gpus", q)
def local_cpus(busid):
    b = busid.lower()
    if len(b.split(":")[0]) == 8: b = b[4:]
    p = f"/sys/bus/pci/devices/{b}/local_cpulist"
    try:
        return open(p).read().strip(), open(f"/sys/bus/pci/devices/{b}/numa_node").read().strip()
    except Exception as e:
        return repr(e), None
print("gpu0 local", local_cpus(q[0]))
def parse(s):
    out = set()
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-"); out |= set(range(int(a), int(b) + 1))
        elif part: out.add(int(part))
    return out
def bw(tag):
    n = 1 << 20
    h = torch.empty((n, 9), dtype=torch.float32).pin_memory()
    h.fill_(1.0)
    d = torch.empty((n, 9), dtype=torch.float32, device="cuda")
    for sz in (n, n // 4):
        for _ in range(3): d[:sz].copy_(h[:sz], non_blocking=True)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(20): d[:sz].copy_(h[:sz], non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / 20
        print(f"{tag}: H2D pinned {sz*36/1e6:.1f} MB: {sz*36/dt/1e9:.1f} GB/s")
bw("default affinity")
allc = os.sched_getaffinity(0)
lc, node = local_cpus(q[0])
try:
    loc = parse(lc) & allc
    if loc:
        os.sched_setaffinity(0, loc); bw(f"local cpus ({len(loc)})")
        rem = allc - loc
        if rem:
            os.sched_setaffinity(0, rem); bw(f"remote cpus ({len(rem)})")
except Exception as e:
    print("affinity test failed", e)
