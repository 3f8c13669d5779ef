"""Export (q -> link poses, contact pairs, verdict) golden vectors FROM THE REAL REFERENCE STACK (Genesis), for the
day it is importable -- SURVEY.md 8c: "parity vs Genesis: unpinned" until this has been run.

NOT RUNNABLE IN THE BUILD CONTAINER (no `genesis`, no Panda MJCF/meshes, no network) and therefore untested here;
it only uses the public Genesis calls the reference itself uses (scenes.py:41-92 to build the scene,
planning.py:210-211 `set_qpos` + `detect_collision`, planning.py:221-230 for the forgiveness rule).

    python tools/export_genesis_goldens.py [scene] [n]   ->  tests/golden/genesis_<scene>.npz

The seeded inputs are the same ones the in-repo oracle is tested on (numpy default_rng(20251212), joints uniform in
the Panda limits, fingers open), so `tests/test_golden_cpu.py::test_oracle_vs_genesis_goldens` -- which is skipped
while the file is absent -- then measures how far the primitive-based robot model is from the Genesis meshes:
FK should agree to 1e-5; verdicts are expected to differ only within a few millimetres of contact.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rbe550_final_project_b200 import panda_model as pm  # noqa: E402
from rbe550_final_project_b200 import scenes as sc  # noqa: E402

LINKS = ["link0", "link1", "link2", "link3", "link4", "link5", "link6", "link7", "hand", "left_finger", "right_finger"]
GRIPPER = ("hand", "left_finger", "right_finger")


def main():
    try:
        import genesis as gs
    except ImportError:
        print("genesis is not importable here: nothing exported (parity vs Genesis stays unpinned)")
        return 0
    scene_name = sys.argv[1] if len(sys.argv) > 1 else "goal1_scattered"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    snap = sc.FIXTURES[scene_name]()

    gs.init(backend=gs.cpu, logging_level="warning")
    scene = gs.Scene(show_viewer=False)
    scene.add_entity(gs.morphs.Plane())
    blocks = []
    for k in range(snap.n_obb):
        o = np.asarray(snap.obb[k], dtype=np.float64)
        R = o[6:15].reshape(3, 3)
        yaw = float(np.degrees(np.arctan2(R[1, 0], R[0, 0])))
        blocks.append(scene.add_entity(gs.morphs.Box(size=tuple(2.0 * o[3:6]), pos=tuple(o[0:3]), euler=(0.0, 0.0, yaw))))
    robot = scene.add_entity(gs.morphs.MJCF(file="xml/franka_emika_panda/panda.xml"))
    scene.build()
    base = np.asarray(robot.get_pos(), dtype=float)
    robot.set_pos(base + np.array([0.0, 0.0, float(snap.base[2])]))  # scenes.py:29-34

    rng = np.random.default_rng(20251212)
    q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9))
    q[:, 7:] = 0.04
    q[0], q[1], q[2] = 0.0, pm.Q_SAFE_HOME, pm.Q_SCENE_INIT
    links = [robot.get_link(nm) for nm in LINKS]
    geoms = scene.rigid_solver.geoms
    pos = np.zeros((n, 11, 3))
    quat = np.zeros((n, 11, 4))
    valid = np.zeros(n, dtype=bool)
    valid_attached0 = np.zeros(n, dtype=bool)
    n_pairs = np.zeros(n, dtype=np.int32)
    first_pair = np.full((n, 2), -1, dtype=np.int32)
    for i in range(n):
        robot.set_qpos(q[i])                       # planning.py:210
        pairs = np.asarray(robot.detect_collision())   # planning.py:211
        for l, lk in enumerate(links):
            pos[i, l] = np.asarray(lk.get_pos(), dtype=float)
            quat[i, l] = np.asarray(lk.get_quat(), dtype=float)
        pairs = pairs.reshape(-1, 2) if pairs.size else np.zeros((0, 2), dtype=np.int64)
        n_pairs[i] = len(pairs)
        valid[i] = len(pairs) == 0                 # planning.py:212-215
        if len(pairs):
            first_pair[i] = pairs[0]
        # forgiveness rule of planning.py:221-230 with block 0 attached
        ok = True
        for a, b in pairs:
            la, lb = geoms[int(a)].link.name, geoms[int(b)].link.name
            if (la in GRIPPER and int(b) == blocks[0].idx) or (lb in GRIPPER and int(a) == blocks[0].idx):
                continue
            ok = False
            break
        valid_attached0[i] = ok
    out = os.path.join(ROOT, "tests", "golden", f"genesis_{scene_name}.npz")
    np.savez_compressed(out, q=q, link_pos=pos, link_quat_wxyz=quat, valid=valid, valid_attached0=valid_attached0,
                        n_pairs=n_pairs, first_pair=first_pair, scene=scene_name,
                        genesis_version=str(getattr(gs, "__version__", "unknown")))
    print("wrote", out, "valid fraction", valid.mean())
    return 0


if __name__ == "__main__":
    sys.exit(main())
