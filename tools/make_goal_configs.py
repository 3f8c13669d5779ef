"""Generate tests/golden/goal_configs.json: joint configurations for the hand poses the reference's motion
primitives plan to (motion_primitives.py:263-280, 379-385, 659-674), solved numerically with the CPU
oracle's FK (scipy least squares, seeded at safe_home).  Genesis' IK is not available here, so these stand in
for `robot.inverse_kinematics(link=hand, pos, quat)`; the planner only needs *a* valid goal configuration.

Run:  python tools/make_goal_configs.py
"""
import json
import os
import sys

import numpy as np
from scipy.optimize import least_squares

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import panda_oracle as po  # noqa: E402
from rbe550_final_project_b200 import panda_model as pm  # noqa: E402
from rbe550_final_project_b200 import scenes as sc  # noqa: E402


def quat_to_mat(q):
    return sc.quat_wxyz_to_mat(q)


def ik(pos, quat=(0.0, 1.0, 0.0, 0.0), seed=pm.Q_SAFE_HOME, fingers=0.04):
    Rt = quat_to_mat(quat)

    def resid(x):
        q = np.concatenate([x, [fingers, fingers]])
        R, p = po.fk(q[None])
        e_p = p[0, 8] - pos
        e_r = (R[0, 8] - Rt).ravel()  # full matrix difference: no spurious minimum at a half-turn
        return np.concatenate([e_p, 0.2 * e_r])

    best = None
    rng = np.random.default_rng(0)
    for trial in range(20):
        x0 = seed[:7] if trial == 0 else np.clip(seed[:7] + rng.normal(0, 0.4, 7), pm.Q_LOWER[:7], pm.Q_UPPER[:7])
        r = least_squares(resid, x0, bounds=(pm.Q_LOWER[:7] + 1e-3, pm.Q_UPPER[:7] - 1e-3), xtol=1e-12, ftol=1e-12)
        if best is None or r.cost < best.cost:
            best = r
        if best.cost < 1e-16:
            break
    assert best.cost < 1e-12, f"IK did not converge for {pos}: cost {best.cost}"
    return np.concatenate([best.x, [fingers, fingers]])


def main():
    model = pm.model_arrays()
    out = {}
    cases = {
        # goal1: PICK-UP(r): approach = block top + 0.18 -> centre + 0.20; grasp = centre + 0.12
        "goal1_scattered": [("approach_r", (0.65, 0.0, 0.02 + 0.02 + 0.18)), ("grasp_r", (0.65, 0.0, 0.02 + 0.12)),
                            ("approach_c", (0.45, 0.4, 0.22)), ("place_050_000", (0.50, 0.0, 0.02 + 0.12 + 0.15)),
                            # carry-mode cases (block r in the hand, grasped at grasp_r): pushed 15 mm into the table,
                            # resting exactly on block g, hovering 3 cm above it, and sunk half-way into block g
                            ("carry_low_r", (0.65, 0.0, 0.02 + 0.12 - 0.015)),
                            ("carry_on_g", (0.65, 0.2, 0.02 + 0.04 + 0.12)),
                            ("carry_above_g", (0.65, 0.2, 0.02 + 0.04 + 0.12 + 0.03)),
                            ("carry_into_g", (0.65, 0.2, 0.02 + 0.02 + 0.12))],
        # goal3 tower: approach above the 8-high tower top (z = 0.30 centre) and the loose blocks
        "goal3_tower": [("approach_top", (0.45, 0.0, 0.30 + 0.02 + 0.18)), ("approach_r2", (0.65, -0.4, 0.22)),
                        ("approach_o2", (0.65, 0.4, 0.22))],
        "goal4_task1_pentagon": [("approach_b6", (0.55384, 0.13477, 0.06 + 0.02 + 0.18)),
                                 ("approach_centre", (0.50, 0.10, 0.30))],
    }
    for scene_name, lst in cases.items():
        scene = sc.FIXTURES[scene_name]().as_oracle_scene()
        out[scene_name] = {}
        for name, pos in lst:
            q = ik(np.array(pos))
            m = float(po.state_margin(q[None], scene, model)[0])
            R, p = po.fk(q[None])
            print(f"{scene_name:22s} {name:16s} margin {m:+.4f}  hand {np.round(p[0, 8], 4)}  q {np.round(q, 4)}")
            out[scene_name][name] = {"q": [float(v) for v in q], "hand_pos": list(map(float, pos)), "oracle_margin": m}
    out["safe_home"] = [float(v) for v in pm.Q_SAFE_HOME]
    path = os.path.join(ROOT, "tests", "golden", "goal_configs.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
