"""Golden table of the reference's robot constants (build container only; needs /root/reference):

    python tools/make_robot_constants_golden.py      # -> tests/golden/robot_constants.json

What `tests/test_g1_constants.py` / `tests/test_go1_constants.py` of the reference assert about the COMPILED MjModel - actuator
gains / limits per joint pattern, keyframe base pose and joint positions, collision-geom condim / priority / friction / solimp,
counts, control / force limiting - comes from `asset_zoo/robots/unitree_{g1,go1}/{g1,go1}_constants.py`.  The module imports
under tools/refstub.py (no MuJoCo needed for the numbers), so the values are dumped here and `tests/test_robot_constants.py`
holds this repo's own model compiler (mjlab_b200/mjcf.py + robots.py, which never saw MuJoCo's compiler) to them.
"""
import dataclasses
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import refstub  # noqa: E402

refstub.install()
from mjlab.asset_zoo.robots.unitree_g1 import g1_constants as G  # noqa: E402
from mjlab.asset_zoo.robots.unitree_go1 import go1_constants as O  # noqa: E402


def actuators(art):
  return [dict(joint_names_expr=list(a.joint_names_expr), stiffness=a.stiffness, damping=a.damping, effort_limit=a.effort_limit,
               armature=a.armature, frictionloss=a.frictionloss) for a in art.actuators]


def collision(c):
  d = dataclasses.asdict(c)
  return {k: (list(v) if isinstance(v, tuple) else v) for k, v in d.items()}


def keyframe(k):
  return dict(pos=list(k.pos), rot=list(k.rot), joint_pos=dict(k.joint_pos), joint_vel=dict(k.joint_vel))


out = {
  "g1": dict(actuators=actuators(G.G1_ARTICULATION), soft_joint_pos_limit_factor=G.G1_ARTICULATION.soft_joint_pos_limit_factor,
             keyframe=keyframe(G.KNEES_BENT_KEYFRAME), collision=collision(G.FULL_COLLISION), action_scale=dict(G.G1_ACTION_SCALE),
             num_actuators=29, num_joints=29, foot_geoms=14),
  "go1": dict(actuators=actuators(O.GO1_ARTICULATION), soft_joint_pos_limit_factor=O.GO1_ARTICULATION.soft_joint_pos_limit_factor,
              keyframe=keyframe(O.INIT_STATE), collision=collision(O.FULL_COLLISION), action_scale=dict(O.GO1_ACTION_SCALE),
              num_actuators=12, num_joints=12, foot_geoms=4),
}
path = ROOT / "tests" / "golden" / "robot_constants.json"
path.write_text(json.dumps(out, indent=1))
print("->", path, {k: (len(v["actuators"]), len(v["action_scale"])) for k, v in out.items()})
