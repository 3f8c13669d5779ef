"""The certified never-collide pruning (tools/certify_never_collide.py) must not change any verdict:
dense random sampling against the UNPRUNED pair list, plus a re-run of the certificate on a few pairs."""
import json
import os
import sys

import numpy as np

from rbe550_final_project_b200 import panda_model as pm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_certificate_is_current():
    assert pm.NEVER_COLLIDE_CERT is not None, "data/never_collide.json missing or stale: rerun tools/certify_never_collide.py"
    cert = pm.NEVER_COLLIDE_CERT
    assert cert["model_fingerprint"] == pm.model_fingerprint()
    assert len(cert["never_ss"]) + pm.N_SS_PAIRS == len(pm.SS_PAIRS_UNPRUNED)
    assert len(cert["never_sb"]) + pm.N_SB_PAIRS == len(pm.SB_PAIRS_UNPRUNED)
    assert cert["margin"] >= 1e-3


def test_pruned_pairs_never_come_close_in_2m_samples():
    from oracle.c_oracle import COracle
    full = pm.model_arrays()
    full["ss_pairs"], full["sb_pairs"] = pm.SS_PAIRS_UNPRUNED.copy(), pm.SB_PAIRS_UNPRUNED.copy()
    dropped = pm.model_arrays()
    keep_ss = {tuple(p) for p in pm.SS_PAIRS.tolist()}
    keep_sb = {tuple(p) for p in pm.SB_PAIRS.tolist()}
    dropped["ss_pairs"] = np.array([p for p in pm.SS_PAIRS_UNPRUNED.tolist() if tuple(p) not in keep_ss], dtype=np.int32)
    dropped["sb_pairs"] = np.array([p for p in pm.SB_PAIRS_UNPRUNED.tolist() if tuple(p) not in keep_sb], dtype=np.int32)
    pruned = pm.model_arrays()
    ora_full, ora_drop, ora_pruned = COracle(full, "f64"), COracle(dropped, "f64"), COracle(pruned, "f64")
    empty = {"obb": np.zeros((0, 16)), "table_z": -10.0}
    rng = np.random.default_rng(2025)
    n = 2_000_000
    q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9))
    # bias a third of the samples towards the joint-limit corners where links fold onto each other
    k = n // 3
    corner = rng.random((k, 9)) < 0.5
    q[:k] = np.where(rng.random((k, 9)) < 0.6, np.where(corner, pm.Q_LOWER, pm.Q_UPPER), q[:k])
    m_drop = ora_drop.state_margin(q, empty)
    assert m_drop.min() >= pm.NEVER_COLLIDE_CERT["margin"] - 1e-9, f"a pruned pair came within {m_drop.min()} m"
    m_full = ora_full.state_margin(q, empty)
    m_pruned = ora_pruned.state_margin(q, empty)
    assert ((m_full >= 0) == (m_pruned >= 0)).all()
    close = m_full < pm.NEVER_COLLIDE_CERT["margin"]
    assert np.array_equal(m_full[close], m_pruned[close])


def test_recertify_a_few_pairs():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import certify_never_collide as cert
    never = pm.NEVER_COLLIDE_CERT["never_ss"]
    for a, b in never[:: max(1, len(never) // 6)][:6]:
        kind, _, _, ok, fmin, used = cert.certify(("ss", a, b))
        assert ok and fmin >= cert.MARGIN
    # and a pair that does collide is refuted
    a, b = pm.SS_PAIRS[0]
    assert cert.certify(("ss", int(a), int(b)))[3] in (False,)


def test_generated_cull_constants_are_conservative():
    """The culling constants in csrc/panda_model_gen.h: every moving sphere appears exactly once in the ground-plane
    min-tree; the cull ball of each self-collision block contains every sphere that takes part in the block."""
    import re
    from rbe550_final_project_b200 import panda_model as pm
    hdr = pm.header_text()
    lowest = re.search(r"#define PV_TABLE_LOWEST\(s\) (.*)", hdr).group(1)
    idx = sorted(int(i) for i in re.findall(r"s\[(\d+)\]\.z", lowest))
    assert idx == [i for i in range(pm.N_SPHERES) if int(pm.SPHERE_LINK[i]) != 0]
    # each radius group subtracts the radius of its own spheres
    for grp, r in re.findall(r"\(((?:fminf\(|s\[\d+\]\.z|, |\))+) - ([0-9.e+-]+)f\)", lowest):
        for i in re.findall(r"s\[(\d+)\]", grp):
            assert abs(float(pm.SPHERE_RADIUS[int(i)]) - float(r)) < 1e-7
    block = re.search(r"#define PV_SS_LINKPAIRS\(LP\) \\\n((?:.*\\\n)+)", hdr).group(1)
    n_blocks = 0
    for la, lb, ca, cb, c2 in re.findall(r"LP\((\d+), (\d+), (\d+), (\d+), ([0-9.e+-]+)f\)", block):
        la, lb, ca, cb, c2 = int(la), int(lb), int(ca), int(cb), float(c2)
        pairs = [(int(p), int(q)) for p, q in pm.SS_PAIRS if (int(pm.SPHERE_LINK[p]), int(pm.SPHERE_LINK[q])) == (la, lb)]
        assert pairs and int(pm.SPHERE_LINK[ca]) == la and int(pm.SPHERE_LINK[cb]) == lb
        ra = max(np.linalg.norm(pm.SPHERE_CENTER[p] - pm.SPHERE_CENTER[ca]) + pm.SPHERE_RADIUS[p] for p, _ in pairs)
        rb = max(np.linalg.norm(pm.SPHERE_CENTER[q] - pm.SPHERE_CENTER[cb]) + pm.SPHERE_RADIUS[q] for _, q in pairs)
        # |s[ca] - s[cb]| >= ra + rb  =>  no pair of the block can touch
        assert np.sqrt(c2) >= ra + rb + 0.5 * pm.CULL_SLACK
        n_blocks += 1
    assert n_blocks == len({(int(pm.SPHERE_LINK[p]), int(pm.SPHERE_LINK[q])) for p, q in pm.SS_PAIRS})
    block = re.search(r"#define PV_SBH_LINKS\(LB\) \\\n((?:.*\\\n)+)", hdr).group(1)
    for la, ca, c0, c1, c2, rla in re.findall(r"LB\((\d+), (\d+), ([0-9.e+-]+)f, ([0-9.e+-]+)f, ([0-9.e+-]+)f, ([0-9.e+-]+)f\)", block):
        la, ca, rla = int(la), int(ca), float(rla)
        sph = sorted({int(p) for p, _ in pm.SB_PAIRS if int(pm.SPHERE_LINK[p]) == la})
        need = max(np.linalg.norm(pm.SPHERE_CENTER[p] - pm.SPHERE_CENTER[ca]) + pm.SPHERE_RADIUS[p] for p in sph)
        assert rla >= need + 0.5 * pm.CULL_SLACK
        for k, c in enumerate((float(c0), float(c1), float(c2))):
            has = any(int(pm.SPHERE_LINK[p]) == la and int(kk) == k for p, kk in pm.SB_PAIRS)
            assert (c > 0) == has
            if has:
                assert np.sqrt(c) >= need + pm.BOX_BOUND_RADIUS[k] + 0.5 * pm.CULL_SLACK


def test_motion_reach_bounds_hold():
    """The displacement bound behind the motion certificates (panda_model.motion_reach_bounds; pv_edge_kernel): when the
    joints move by dq, no tested point of the robot -- sphere centres, gripper box centres and corners -- moves farther
    than sum_j |dq_j| R_j + |dq_8| + |dq_9|, in the world and relative to every link proximal to it."""
    from oracle import panda_oracle as po
    rng = np.random.default_rng(5)
    n = 40_000
    Rj = pm.motion_reach_bounds()
    qa = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9))
    step = rng.choice([1e-3, 0.02, 0.3], size=(n, 1)) * rng.standard_normal((n, 9))
    step[n // 2:, :] *= (rng.random((n - n // 2, 9)) < 0.3)  # sparse moves: single joints, where the bound is tightest
    step[:, 7:] *= 0.05
    qb = np.clip(qa + step, pm.Q_LOWER, pm.Q_UPPER)
    bound = (np.abs(qb - qa)[:, :7] * Rj[None]).sum(1) + np.abs(qb - qa)[:, 7:].sum(1)

    def points(q):
        R, p = po.fk(q)
        sl = pm.SPHERE_LINK
        cen = p[:, sl] + np.einsum("nsij,sj->nsi", R[:, sl], pm.SPHERE_CENTER)
        sg = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 0, 1)], float)  # corners + centres
        loc = pm.BOX_CENTER[:, None, :] + sg[None] * pm.BOX_HALF[:, None, :]
        box = p[:, pm.BOX_LINK, None, :] + np.einsum("nkij,kcj->nkci", R[:, pm.BOX_LINK], loc)
        link = np.concatenate([sl, np.repeat(pm.BOX_LINK, sg.shape[0])])
        return np.concatenate([cen, box.reshape(n, -1, 3)], 1), link, R, p

    Pa, link, Ra, pa = points(qa)
    Pb, _, Rb, pb = points(qb)
    world = np.linalg.norm(Pb - Pa, axis=2).max(1)
    assert (world <= bound * (1 + 1e-9) + 1e-12).all()
    # the chord of a straight joint-space motion is what the validator needs: sampled states of the SAME motion
    # relative to a proximal link frame a: x_a = R_a^T (x - p_a)
    for a in (1, 2, 3, 4):
        la = np.einsum("nji,npj->npi", Ra[:, a], Pa - pa[:, a, None, :])
        lb = np.einsum("nji,npj->npi", Rb[:, a], Pb - pb[:, a, None, :])
        distal = link > a
        rel = np.linalg.norm(lb - la, axis=2)[:, distal].max(1)
        assert (rel <= bound * (1 + 1e-9) + 1e-12).all(), a
    # and the bound is not vacuous: some single-joint moves come within 25 % of it
    assert (world / np.maximum(bound, 1e-12)).max() > 0.75
