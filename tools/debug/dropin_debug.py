import sys, os, logging
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, dataclasses, importlib, tempfile
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from rbe550_final_project_b200.sim_stub import create_scene
from rbe550_final_project_b200.validity import PandaValidity
logging.basicConfig(level=logging.INFO)
shim = tempfile.mkdtemp()
open(os.path.join(shim, "planning.py"), "w").write("from rbe550_final_project_b200.planning import *\nfrom rbe550_final_project_b200.planning import PlannerInterface\n")
open(os.path.join(shim, "robot_adapter.py"), "w").write("from rbe550_final_project_b200.robot_adapter import *\n")
sys.path[:0] = [shim, os.path.join(ROOT, "oracle", "_ref", "reference_caller")]
real = dataclasses.dataclass
def dc(cls=None, **kw):
    def wrap(c):
        for name, val in list(vars(c).items()):
            if isinstance(val, np.ndarray):
                setattr(c, name, dataclasses.field(default_factory=lambda v=val: v))
        return real(c, **kw)
    return wrap if cls is None else wrap(cls)
dataclasses.dataclass = dc
import motion_primitives as mp
dataclasses.dataclass = real
pv = PandaValidity(0)
scene, franka, blocks = create_scene("goal1_scattered")
franka.raw.attach_validity(pv)
franka.set_qpos(pm.Q_SAFE_HOME)
ex = mp.MotionPrimitiveExecutor(scene, franka, blocks)
ex.planner.validity = pv
print("pick", ex.pick_up("r"), ex.planner.last_stats)
print("held", scene.held is not None, "q", franka.get_qpos(), "block r", blocks["r"].get_pos(), "hand", franka.get_link("hand").get_pos())
hand = franka.get_link("hand")
q_app = ex._ik_for_pose(np.array([0.5, -0.2, 0.29]), ex.grasp_quat)
print("ik approach", q_app)
if q_app is not None:
    path = ex.planner.plan_path(qpos_goal=q_app, num_waypoints=150, attached_object=blocks["r"], timeout=10.0)
    print("plan", len(path), ex.planner.last_stats)
    import torch
    snap = ex.planner._snapshot
    pv.set_scene(snap); pv.set_attached(snap.index_of_entity(blocks["r"].idx))
    for name, q in (("start", franka.get_qpos()), ("goal", q_app)):
        qq = torch.as_tensor(np.asarray(q, np.float32)[None], device="cuda")
        m, cu = pv.state_margins(qq, want_culprit=True)
        from rbe550_final_project_b200.validity import decode_culprit
        print(name, float(m[0]), decode_culprit(int(cu[0])), pv.contacts(qq)[0])
