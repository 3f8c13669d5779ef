"""CPU tests (-m "not gpu"): the oracle against the analytic known answers of SURVEY.md App. A and the
acceptance facts of App. F, numpy oracle vs C oracle, host logic, and the C-ABI library's exports."""
import ctypes
import json
import os

import numpy as np
import pytest

from conftest import random_configs
from oracle import panda_oracle as po
from rbe550_final_project_b200 import panda_model as pm
from rbe550_final_project_b200 import scenes as sc
from rbe550_final_project_b200.pathutil import interpolate

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_fk_known_answers():
    """SURVEY.md App. A table (base offset 0): public Franka kinematics."""
    R, p = po.fk(np.zeros((1, 9)), base=(0, 0, 0))
    assert np.allclose(p[0, 7], [0.088, 0, 1.033], atol=1e-9)
    assert np.allclose(p[0, 8], [0.088, 0, 0.926], atol=1e-9)
    assert np.allclose(p[0, 9], [0.088, 0, 0.8676], atol=1e-9)
    assert np.allclose(R[0, 8][:, 2], [0, 0, -1], atol=1e-9)
    R, p = po.fk(pm.Q_SAFE_HOME[None], base=(0, 0, 0))
    assert np.allclose(p[0, 3], [-0.22336, 0, 0.55653], atol=1e-5)
    assert np.allclose(R[0, 3][:, 2], [-0.707, 0, 0.707], atol=1e-3)
    assert np.allclose(p[0, 5], [0.21902, 0, 0.69727], atol=1e-5)
    assert np.allclose(p[0, 8], [0.30702, 0, 0.59027], atol=1e-5)
    assert np.allclose(R[0, 8][:, 2], [0, 0, -1], atol=1e-3)
    assert np.allclose(p[0, 9], [0.30704, -0.04, 0.53187], atol=1e-5)
    assert np.allclose(p[0, 10], [0.30700, 0.04, 0.53187], atol=1e-5)
    R, p = po.fk(pm.Q_SCENE_INIT[None], base=(0, 0, 0))
    assert np.allclose(p[0, 8], [0.15327, -0.08921, 0.97041], atol=1e-5)
    assert np.allclose(R[0, 8][:, 2], [0.479, 0, -0.878], atol=1e-3)


def test_fk_matches_the_published_dh_table():
    """An independent pin of the oracle's kinematics: the modified (Craig) Denavit-Hartenberg table Franka publishes for
    the Panda (a, d, alpha per joint, flange at d = 0.107), composed as Rx(alpha) Tx(a) Rz(theta) Tz(d), must give the
    frames the oracle derives from the MJCF body chain -- every link frame, over random configurations."""
    dh = [  # a, d, alpha
        (0.0, 0.333, 0.0), (0.0, 0.0, -np.pi / 2), (0.0, 0.316, np.pi / 2), (0.0825, 0.0, np.pi / 2),
        (-0.0825, 0.384, -np.pi / 2), (0.0, 0.0, np.pi / 2), (0.088, 0.0, np.pi / 2),
    ]

    def rx(a):
        c, s = np.cos(a), np.sin(a)
        return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1.0]])

    def rz(a):
        c, s = np.cos(a), np.sin(a)
        return np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])

    def tr(x, y, z):
        t = np.eye(4)
        t[:3, 3] = (x, y, z)
        return t

    q = random_configs(300, 5, fingers="random").astype(np.float64)
    q[0] = 0.0
    q[1] = pm.Q_SAFE_HOME
    base = (0.1, -0.2, 0.01)
    R, p = po.fk(q, base=base)
    for n in range(len(q)):
        T = tr(*base)
        for j, (a, d, al) in enumerate(dh):
            T = T @ rx(al) @ tr(a, 0, 0) @ rz(q[n, j]) @ tr(0, 0, d)
            assert np.allclose(T[:3, :3], R[n, j + 1], atol=1e-12) and np.allclose(T[:3, 3], p[n, j + 1], atol=1e-12), (n, j)
        flange = T @ tr(0, 0, 0.107)
        assert np.allclose(flange[:3, 3], p[n, 8], atol=1e-12)  # the hand body sits on the flange ...
        assert np.allclose(flange[:3, :3] @ rz(-np.pi / 4)[:3, :3], R[n, 8], atol=1e-7)  # ... turned by -45 deg about z


def test_derived_ompl_constants():
    assert abs(pm.SPACE_EXTENT - 13.03716) < 1e-5
    assert abs(pm.VALIDITY_RESOLUTION - 0.130372) < 1e-6
    assert abs(pm.RRTC_RANGE - 2.607432) < 1e-6


def test_c_oracle_matches_numpy_oracle(model, c64, c32):
    q = random_configs(4000, 3, fingers="random").astype(np.float64)
    R, p = po.fk(q)
    R2, p2 = c64.fk(q)
    assert np.abs(R - R2).max() < 1e-12 and np.abs(p - p2).max() < 1e-12
    for name in ("goal1_scattered", "goal4_task1_pentagon", "goal3_tower"):
        s = sc.FIXTURES[name]().as_oracle_scene()
        for att in (-1, 1):
            a = po.state_margin(q, s, model, attached=att)
            b = c64.state_margin(q, s, attached=att)
            c = c32.state_margin(q, s, attached=att)
            assert np.abs(a - b).max() < 1e-12
            assert np.abs(a - c).max() < 5e-6
    s = sc.goal4_task1_pentagon().as_oracle_scene()
    qb = np.clip(q + 0.25, pm.Q_LOWER, pm.Q_UPPER)
    for steps in (0, 16):
        a = po.edge_margin(q[:300], qb[:300], s, model, n_steps=steps)
        b = c64.edge_margin(q[:300], qb[:300], s, n_steps=steps)
        assert np.abs(a - b).max() < 1e-12
    # early exit keeps the verdict
    full = c64.edge_margin(q, qb, s, n_steps=0)
    ee, cnt = c64.edge_margin(q, qb, s, n_steps=0, early_exit=True, return_count=True)
    assert ((full >= 0) == (ee >= 0)).all() and cnt > 0


def test_acceptance_constraints(model):
    """poses the reference plans from / to are valid in every scene."""
    goals = json.load(open(os.path.join(GOLD, "goal_configs.json")))
    for name, f in sc.FIXTURES.items():
        s = f().as_oracle_scene()
        q = np.stack([pm.Q_SAFE_HOME, pm.Q_SAFE_HOME_039, pm.Q_SCENE_INIT])
        assert (po.state_margin(q, s, model) > 0).all(), name
    for scene_name, cases in goals.items():
        if scene_name == "safe_home":
            continue
        s = sc.FIXTURES[scene_name]().as_oracle_scene()
        for nm, rec in cases.items():
            m = po.state_margin(np.array(rec["q"])[None], s, model)[0]
            assert m > 0 and abs(m - rec["oracle_margin"]) < 1e-9, (scene_name, nm)
            R, p = po.fk(np.array(rec["q"])[None])
            assert np.allclose(p[0, 8], rec["hand_pos"], atol=1e-6)
            assert np.allclose(R[0, 8][:, 2], [0, 0, -1], atol=1e-5)  # quat [0,1,0,0]: hand z down


def test_carry_mode_semantics(model, c64, c32):
    """SURVEY.md 8f-3 / App. E-3 (not reference behaviour): block r rides on the hand; resting contacts pass
    (1 mm allowance), pushing the block into the table or into block g does not; numpy == C oracle."""
    scene = sc.goal1_scattered()
    s = scene.as_oracle_scene()
    goals = json.load(open(os.path.join(GOLD, "goal_configs.json")))["goal1_scattered"]
    qg = np.array(goals["grasp_r"]["q"])
    R, p = po.fk(qg[None])
    Rh, ph = R[0, 8], p[0, 8]
    rec = np.asarray(s["obb"], dtype=np.float64).reshape(-1, 16)[0]
    s["carried"] = dict(index=0, R=Rh.T @ rec[6:15].reshape(3, 3), t=Rh.T @ (rec[0:3] - ph), shrink=1e-3)
    expect = {"grasp_r": 0.001, "place_050_000": 0.103, "carry_on_g": 0.001, "carry_low_r": -0.014, "carry_into_g": -0.019}
    for name, m in expect.items():
        got = po.state_margin(np.array(goals[name]["q"])[None], s, model)[0]
        assert abs(got - m) < 1e-6, (name, got)
    q = random_configs(4000, 77, fingers="random").astype(np.float64)
    a = po.state_margin(q, s, model)
    assert np.abs(a - c64.state_margin(q, s)).max() < 1e-12
    assert np.abs(a - c32.state_margin(q, s)).max() < 1e-5
    base = dict(s)
    del base["carried"]
    assert ((a >= 0) != (po.state_margin(q, base, model, attached=0) >= 0)).mean() > 0.002
    qb = np.clip(q + np.random.default_rng(3).normal(0, 0.3, q.shape), pm.Q_LOWER, pm.Q_UPPER)
    e = po.edge_margin(q[:300], qb[:300], s, model, n_steps=0)
    assert np.abs(e - c64.edge_margin(q[:300], qb[:300], s, n_steps=0)).max() < 1e-12


def test_attached_and_self_semantics(model):
    s = sc.goal1_scattered().as_oracle_scene()
    goals = json.load(open(os.path.join(GOLD, "goal_configs.json")))
    q = np.array(goals["goal1_scattered"]["grasp_r"]["q"])
    q[7:] = 0.015  # fingers closed into block r
    assert po.state_margin(q[None], s, model, attached=-1)[0] < 0
    assert po.state_margin(q[None], s, model, attached=0)[0] > 0
    assert po.state_margin(q[None], s, model, attached=1)[0] < 0
    # monotone under obstacle inflation
    qq = random_configs(3000, 9).astype(np.float64)
    m0 = po.state_margin(qq, s, model)
    s2 = {"obb": s["obb"].copy(), "table_z": s["table_z"]}
    s2["obb"][:, 3:6] *= 1.5
    assert (po.state_margin(qq, s2, model) <= m0 + 1e-12).all()
    # self-collision off can only raise the margin
    assert (po.state_margin(qq, s, model, self_collision=False) >= m0 - 1e-12).all()
    # joint limits (planning.py:139-150)
    bad = pm.Q_SAFE_HOME.copy()
    bad[7] = 0.0405
    assert not po.in_bounds(bad[None], model)[0] and po.in_bounds(pm.Q_SAFE_HOME[None], model)[0]


def test_pair_filter(model):
    """Static pair filter (App. C): no same-link, no parent-child pairs; counts frozen with the header."""
    for a, b in model["ss_pairs"]:
        la, lb = int(pm.SPHERE_LINK[a]), int(pm.SPHERE_LINK[b])
        assert la != lb and pm.PARENT[la] != lb and pm.PARENT[lb] != la
    for a, k in model["sb_pairs"]:
        la, lk = int(pm.SPHERE_LINK[a]), int(pm.BOX_LINK[k])
        assert la != lk and pm.PARENT[lk] != la
    txt = open(pm.HEADER_PATH).read()
    assert txt == pm.header_text(), "csrc/panda_model_gen.h is stale: run python -m rbe550_final_project_b200.panda_model"
    # the culling balls really contain their spheres
    for l, c, r in pm.link_groups():
        for i in range(pm.N_SPHERES):
            if int(pm.SPHERE_LINK[i]) == l:
                assert np.linalg.norm(pm.SPHERE_CENTER[i] - pm.SPHERE_CENTER[c]) + pm.SPHERE_RADIUS[i] <= r + 1e-12


def test_interpolate_matches_oracle_restatement():
    rng = np.random.default_rng(4)
    for n_pts, count in [(2, 150), (3, 150), (7, 100), (5, 5), (12, 8), (4, 11)]:
        pts = rng.uniform(-1, 1, size=(n_pts, 9))
        a = interpolate(pts, count)
        b = po.interpolate_path(pts, count)
        assert a.shape == b.shape and np.allclose(a, b, atol=1e-15)
        assert np.array_equal(a[0], pts[0]) and np.array_equal(a[-1], pts[-1])
        if count >= n_pts:
            assert len(a) == count
            # every original vertex survives, in order
            idx = 0
            for v in pts:
                while not np.array_equal(a[idx], v):
                    idx += 1
        else:
            assert len(a) == n_pts


def test_bit_packing_roundtrip():
    rng = np.random.default_rng(1)
    for n in (1, 31, 32, 33, 1000):
        v = rng.random(n) < 0.5
        w = po.pack_bits(v)
        assert w.dtype == np.uint32 and len(w) == (n + 31) // 32
        assert np.array_equal(po.unpack_bits(w, n), v)
        from rbe550_final_project_b200.validity import unpack_bits
        assert np.array_equal(unpack_bits(w, n), v)


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    r = po.philox4x32(np.array([[0, 0, 0, 0]], dtype=np.uint32), (0, 0))
    assert [hex(int(x)) for x in r[0]] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    r = po.philox4x32(np.array([[0xffffffff] * 4], dtype=np.uint32), (0xffffffff, 0xffffffff))
    assert [hex(int(x)) for x in r[0]] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    r = po.philox4x32(np.array([[0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], dtype=np.uint32),
                      (0xa4093822, 0x299f31d0))
    assert [hex(int(x)) for x in r[0]] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_sweep_stream_is_shard_invariant(model):
    a = po.sweep_configs(0, 4096, 7, model)
    b = np.concatenate([po.sweep_configs(0, 1024, 7, model), po.sweep_configs(1024, 3072, 7, model)])
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert (a[:, :7] >= pm.Q_LOWER[:7].astype(np.float32)).all() and (a[:, :7] <= pm.Q_UPPER[:7].astype(np.float32)).all()


def test_scene_fixtures_and_snapshot():
    pent = sc.goal4_task1_pentagon()
    assert pent.n_obb == 10
    # App. B.3 numbers
    assert np.allclose(pent.obb[0, :3], [0.56450, 0.10000, 0.02], atol=1e-5)
    assert np.allclose(pent.obb[1, :3], [0.52304, 0.15706, 0.02], atol=1e-5)
    assert np.allclose(pent.obb[5, :3], [0.55384, 0.13477, 0.06], atol=1e-5)
    assert np.allclose(pent.obb[9, :3], [0.55384, 0.06423, 0.06], atol=1e-5)
    R1 = pent.obb[1, 6:15].reshape(3, 3)
    assert np.allclose(np.degrees(np.arctan2(R1[1, 0], R1[0, 0])), 72.0, atol=1e-4)
    tower = sc.goal3_tower()
    assert tower.n_obb == 10 and np.allclose(tower.obb[7, :3], [0.45, 0, 0.30], atol=1e-6)
    # snapshot of a (stub) live scene reproduces the fixture, plane and robot skipped
    from rbe550_final_project_b200.sim_stub import create_scene
    scene, franka, blocks = create_scene("goal4_task1_pentagon")
    snap = sc.snapshot_from_sim(scene, franka)
    assert snap.n_obb == 10 and np.allclose(snap.obb[:, :15], pent.obb[:, :15], atol=1e-6)
    assert snap.entity_idx == list(range(1, 11)) and snap.index_of_entity(blocks["b3"].idx) == 2
    assert np.allclose(snap.base, pm.BASE_LIFT)


def test_robot_adapter_forwards():
    from rbe550_final_project_b200.robot_adapter import RobotAdapter
    from rbe550_final_project_b200.sim_stub import StubPanda
    raw = StubPanda(7)
    ad = RobotAdapter(raw, scene=None)
    assert ad.raw is raw and ad.n_qs == 9 and ad._solver.n_envs == 0
    ad.set_qpos(pm.Q_SAFE_HOME)
    assert np.allclose(ad.get_qpos(), pm.Q_SAFE_HOME) and raw.set_qpos_calls == 1
    with pytest.raises(AttributeError):
        ad.no_such_attribute


def test_cabi_library_exports_every_declared_symbol():
    from rbe550_final_project_b200 import _cabi
    lib_path = _cabi.build()
    lib = ctypes.CDLL(lib_path)
    hdr = open(os.path.join(os.path.dirname(GOLD), "..", "include", "panda_validity.h")).read()
    import re
    declared = sorted(set(re.findall(r"\b(pv_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found in the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/panda_validity.h but not exported"
    assert sorted(_cabi.EXPORTS) == declared
    # model constants agree between the header the kernels were compiled with and the Python table
    ns, nb, nss, nsb = (ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int())
    lib.pv_model_info(ctypes.byref(ns), ctypes.byref(nb), ctypes.byref(nss), ctypes.byref(nsb))
    assert (ns.value, nb.value, nss.value, nsb.value) == (pm.N_SPHERES, pm.N_BOXES, pm.N_SS_PAIRS, pm.N_SB_PAIRS)
    lo, hi = (ctypes.c_float * 9)(), (ctypes.c_float * 9)()
    lib.pv_joint_limits(lo, hi)
    assert np.allclose(list(lo), pm.Q_LOWER.astype(np.float32)) and np.allclose(list(hi), pm.Q_UPPER.astype(np.float32))


def test_no_gpu_means_loud_failure():
    """There is no CPU fallback: without a device pv_create fails and the wrapper raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rbe550_final_project_b200.validity import PandaValidity, PandaValidityError
    with pytest.raises(PandaValidityError):
        PandaValidity(0)
