"""Developer check: the sweep mask must not depend on how the index range is cut into launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rbe550_final_project_b200 import scenes as sc
from rbe550_final_project_b200.validity import PandaValidity
pv = PandaValidity(0); pv.set_scene(sc.goal1_scattered())
n = 104_857_600; seed = 20251212
a, ca = pv.sweep(0, n, seed)
a2, ca2 = pv.sweep(0, n, seed)
print("repeat identical:", bool(torch.equal(a, a2)), int(ca.item()), int(ca2.item()))
for parts in (2, 4, 8):
    per = n // parts
    pieces = [pv.sweep(k * per, per, seed) for k in range(parts)]
    b = torch.cat([p[0] for p in pieces]); cb = sum(int(p[1].item()) for p in pieces)
    diff = (a != b).nonzero().flatten()
    print(parts, "pieces identical:", bool(torch.equal(a, b)), cb, "differing words:", diff[:10].tolist())
    if len(diff):
        w = int(diff[0].item()); x = int(a[w].item()) ^ int(b[w].item())
        print("   word", w, "xor", hex(x & 0xffffffff))
