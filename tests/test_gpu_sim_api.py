"""The reference's own boundary tests, mirrored against the B200 engine through the same public API:

  tests/test_sim.py:44-135                   option plumbing, full reset, selective reset
  tests/test_sim_data.py:31-83               bridge shares memory / keeps addresses / is read-only / caches
  tests/test_domain_randomization.py:99-301  randomize_field ranges + diversity, graph re-captured after expansion, no
                                             accumulation, partial axes, single env without expansion
  sim/sim_data.py:23-31                      stride-0 model views index like per-world ones (before and after expansion)

Same robot XML as the reference's tests (three boxes: its two feet form a box-box pair), compiled by mjlab_b200.mjcf instead
of MuJoCo; `Simulation` is also built from a `mujoco.MjModel` stand-in (tests/fake_mujoco.py) and must behave identically."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

from tests.test_boxbox_cpu import BOX_SCENE, REF_TEST_XML  # noqa: E402

FRICTION_RANGE, MASS_SCALE_RANGE, DAMPING_RANGE, NUM_ENVS = (0.3, 1.2), (0.8, 1.2), (0.1, 0.5), 4


def _model(xml=REF_TEST_XML):
  from mjlab_b200 import mjcf
  return mjcf.Spec.from_string(xml).compile()


def _sim(num_envs, cfg=None, model=None):
  from mjlab_b200.sim import Simulation, SimulationCfg
  return Simulation(num_envs=num_envs, cfg=cfg or SimulationCfg(), model=model or _model(), device=DEV)


# ---- tests/test_sim.py ---------------------------------------------------------------------------------------------
def test_simulation_config_is_piped():
  from mjlab_b200 import mjcf
  from mjlab_b200.sim import MujocoCfg, SimulationCfg
  cfg = SimulationCfg(contact_sensor_maxmatch=128, ls_parallel=False,
                      mujoco=MujocoCfg(timestep=0.02, integrator="euler", solver="cg", iterations=7, ls_iterations=14,
                                       ccd_iterations=20, gravity=(0, 0, 7.5)))
  sim = _sim(1, cfg)
  assert sim.mj_model.opt.timestep == cfg.mujoco.timestep
  assert sim.mj_model.opt.integrator == mjcf.INT_EULER
  assert sim.mj_model.opt.solver == mjcf.SOL_CG
  assert sim.mj_model.opt.iterations == cfg.mujoco.iterations
  assert sim.mj_model.opt.ls_iterations == cfg.mujoco.ls_iterations
  assert sim.mj_model.opt.ccd_iterations == cfg.mujoco.ccd_iterations
  assert tuple(sim.mj_model.opt.gravity) == cfg.mujoco.gravity
  np.testing.assert_almost_equal(sim.model.opt.timestep[0].cpu().numpy(), cfg.mujoco.timestep)
  np.testing.assert_almost_equal(sim.model.opt.gravity[0].cpu().numpy(), cfg.mujoco.gravity)
  assert sim.model.opt.integrator == mjcf.INT_EULER
  assert sim.model.opt.solver == mjcf.SOL_CG
  assert sim.model.opt.iterations == cfg.mujoco.iterations
  assert sim.wp_model.opt.contact_sensor_maxmatch == cfg.contact_sensor_maxmatch
  assert sim.wp_model.opt.ls_parallel == cfg.ls_parallel
  sim.wp_model.opt.ls_parallel = True                       # assignable, as the reference does at construction (sim.py:140-141)
  assert sim.model.opt.ls_parallel is True
  # gravity (0, 0, +7.5) really reaches the kernels: the free body accelerates upwards
  sim.step()
  assert float(sim.data.qvel[0, 2]) == pytest.approx(7.5 * 0.02, rel=1e-5)
  # PGS is outside the supported subset: refused at setup, not ignored; the elliptic cone is piped like every other option
  with pytest.raises(ValueError, match="PGS"):
    _sim(1, SimulationCfg(mujoco=MujocoCfg(solver="pgs")))
  ell = _sim(1, SimulationCfg(mujoco=MujocoCfg(cone="elliptic", impratio=3.0)))
  assert ell.mj_model.opt.cone == mjcf.CONE_ELLIPTIC and ell.model.opt.cone == mjcf.CONE_ELLIPTIC
  assert ell.mj_model.opt.impratio == 3.0


def test_cg_solver_on_the_gpu_matches_oracle():
  import copy
  import mjlab_b200
  from mjlab_b200.sim import MujocoCfg, SimulationCfg
  from tests import util
  m = copy.deepcopy(mjlab_b200.load_model("go1_flat"))
  nw = 8
  qpos, qvel, ctrl, xfrc = util.random_states(m, nw, seed=5)
  sim = _sim(nw, SimulationCfg(nconmax=35, njmax=300, mujoco=MujocoCfg(timestep=0.005, solver="cg", iterations=100, ls_iterations=50)), m)
  f = lambda a: torch.as_tensor(a, dtype=torch.float32, device=DEV)  # noqa: E731
  sim.data.qpos[:], sim.data.qvel[:], sim.data.ctrl[:] = f(qpos), f(qvel), f(ctrl)
  sim.forward()
  ds = util.oracle_rollout(sim.mj_model, qpos, qvel, ctrl, xfrc, 1, 35, 300, integrate_last=False)
  assert sim.mj_model.opt_solver == 1
  for w, d in enumerate(ds):
    assert int(sim.data.nefc[w]) == d.nefc
    q = sim.data.qacc[w].cpu().numpy()
    assert np.abs(q - d.qacc).max() < 2e-3 * max(1.0, np.abs(d.qacc).max()), w
  assert int(sim.data.solver_niter.max()) > 3


def test_sim_reset_restores_initial_state():
  sim = _sim(2)
  qpos0, qvel0 = sim.data.qpos.clone(), sim.data.qvel.clone()
  for _ in range(10):
    sim.step()
  assert not torch.allclose(sim.data.qpos, qpos0)
  assert not torch.allclose(sim.data.qvel, qvel0)
  sim.reset()
  torch.testing.assert_close(sim.data.qpos[:], qpos0)
  torch.testing.assert_close(sim.data.qvel[:], qvel0)
  assert (sim.data.qacc_warmstart == 0).all()


def test_sim_reset_selective():
  sim = _sim(4)
  qpos0 = sim.data.qpos.clone()
  for _ in range(10):
    sim.step()
  after = sim.data.qpos.clone()
  sim.reset(torch.tensor([1, 3], device=DEV))
  torch.testing.assert_close(sim.data.qpos[1], qpos0[1])
  torch.testing.assert_close(sim.data.qpos[3], qpos0[3])
  torch.testing.assert_close(sim.data.qpos[0], after[0])
  torch.testing.assert_close(sim.data.qpos[2], after[2])


def test_simulation_accepts_an_mjmodel():
  """Simulation(num_envs, cfg, mujoco.MjModel, device) (sim.py:122-149): an object with MjModel's attribute surface is converted
  in-process and steps bit-identically to the same scene passed as a ModelDesc."""
  import mjlab_b200
  from mjlab_b200.sim import MujocoCfg, SimulationCfg
  from tests.fake_mujoco import FakeMjModel
  from tests import util
  m = mjlab_b200.load_model("go1_flat")
  cfg = lambda: SimulationCfg(nconmax=35, njmax=300, mujoco=MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20))  # noqa: E731
  a, b = _sim(16, cfg(), m), _sim(16, cfg(), FakeMjModel(m))
  qpos, qvel, ctrl, _ = util.random_states(m, 16, seed=4)
  for s in (a, b):
    s.data.qpos[:] = torch.as_tensor(qpos, dtype=torch.float32, device=DEV)
    s.data.qvel[:] = torch.as_tensor(qvel, dtype=torch.float32, device=DEV)
    s.data.ctrl[:] = torch.as_tensor(ctrl, dtype=torch.float32, device=DEV)
    for _ in range(5):
      s.step()
  assert int(a.data.ncon.sum()) > 0
  for f in ("qpos", "qvel", "sensordata", "ncon", "efc_force"):
    assert torch.equal(getattr(a.data, f), getattr(b.data, f)), f
  assert b.mj_model.body(1).name == m.body_names[1] and b.mj_model.opt.timestep == 0.005


# ---- tests/test_sim_data.py ------------------------------------------------------------------------------------------
def test_bridge_on_engine_buffers():
  from mjlab_b200.sim import WarpBridge
  sim = _sim(2)
  assert isinstance(sim.data, WarpBridge) and isinstance(sim.model, WarpBridge)
  q = sim.data.qpos
  assert sim.data.qpos is q                                   # cached: the same tensor object on every access
  ptr = q.data_ptr()
  sim.data.qpos[:] = torch.zeros_like(q)
  assert sim.data.qpos.data_ptr() == ptr and torch.all(sim.data.qpos == 0)
  sim.data.qpos[0, 2] = 0.75                                  # writes land in the buffer the kernels read:
  sim.forward()
  assert float(sim.data.xpos[0, 1, 2]) == pytest.approx(0.75)  # ... forward kinematics sees them
  assert sim.data.qpos.data_ptr() == ptr
  assert torch.allclose(sim.data.qpos * 2, q * 2)
  with pytest.raises(AttributeError, match="Cannot set attribute 'qpos' on WarpBridge"):
    sim.data.qpos = torch.zeros_like(q)
  with pytest.raises(AttributeError, match="Use in-place operations instead"):
    sim.model.body_mass = 1.0
  assert sim.data.nworld == 2


def test_stride0_model_views_index_like_per_world_ones():
  sim = _sim(NUM_ENVS)
  f = sim.model.geom_friction
  assert f.shape == (NUM_ENVS, 3, 3) and f.stride(0) == 0     # shared row, leading world axis (sim_data.py:23-31)
  eg, ig = torch.meshgrid(torch.arange(NUM_ENVS, device=DEV), torch.tensor([0, 2], device=DEV), indexing="ij")
  before = f[eg, ig].clone()
  assert before.shape == (NUM_ENVS, 2, 3) and torch.all(before[..., 0] == 0.5)
  sim.expand_model_fields(("geom_friction",))
  g = sim.model.geom_friction
  assert g.stride(0) == 9 and g.data_ptr() != f.data_ptr()    # a real per-world copy at a new address
  assert torch.equal(g[eg, ig], before)
  g[1, 2, 0] = 0.9
  assert float(sim.model.geom_friction[1, 2, 0]) == pytest.approx(0.9) and float(sim.model.geom_friction[0, 2, 0]) == 0.5


# ---- tests/test_domain_randomization.py ------------------------------------------------------------------------------
class _Env:
  def __init__(self, sim):
    from mjlab_b200.entity import Entity
    self.sim, self.num_envs, self.device = sim, sim.num_envs, DEV
    self.scene = {"robot": Entity("robot", sim, prefix="")}


def _dr_env(num_envs=NUM_ENVS, expand=("geom_friction", "body_mass", "dof_damping")):
  sim = _sim(num_envs)
  if expand:
    sim.expand_model_fields(expand)
  return _Env(sim)


@pytest.mark.parametrize("field,ranges,operation,asset_names,axes,seed", [
  ("geom_friction", FRICTION_RANGE, "abs", {"geom_names": [".*"]}, [0], 123),
  ("body_mass", MASS_SCALE_RANGE, "scale", {"body_names": [".*"]}, None, 456),
  ("dof_damping", DAMPING_RANGE, "abs", {"joint_names": [".*"]}, None, 789),
])
def test_randomize_field(field, ranges, operation, asset_names, axes, seed):
  from mjlab_b200.events import randomize_field
  from mjlab_b200.managers import SceneEntityCfg
  torch.manual_seed(seed)
  env = _dr_env()
  robot = env.scene["robot"]
  if field == "geom_friction":
    model_field = env.sim.model.geom_friction[:, robot.indexing.geom_ids[0], 0]
  elif field == "body_mass":
    model_field = env.sim.model.body_mass[:, robot.indexing.body_ids[0]]
  else:
    idx = robot.indexing.joint_v_adr
    env.sim.model.dof_damping[:, idx] = 0.0
    model_field = env.sim.model.dof_damping[:, idx[0]]
  initial = model_field.clone()
  randomize_field(env, env_ids=None, field=field, ranges=ranges, operation=operation,
                  asset_cfg=SceneEntityCfg("robot", **asset_names), axes=axes)
  new = model_field
  assert not torch.all(new == initial)
  if operation == "abs":
    assert torch.all((new >= ranges[0]) & (new <= ranges[1]))
  else:
    assert torch.all((new >= ranges[0] * initial) & (new <= ranges[1] * initial))
  assert len(torch.unique(new)) >= 2
  # and the kernels read the randomised values: heavier / lighter bodies change the mass matrix of the worlds differently
  if field == "body_mass":
    env.sim.forward()
    assert len(torch.unique(env.sim.data.qM[:, 0, 0])) >= 2


def test_expand_model_fields_recreates_cuda_graph():
  sim = _sim(NUM_ENVS)
  assert sim.use_cuda_graph
  g0 = sim.step_graph
  sim.expand_model_fields(("geom_friction",))
  assert sim.step_graph is not g0, "CUDA graph was not recreated after expand_model_fields"
  # the recaptured graph reads the new arrays: zero friction on the feet of env 1 only changes env 1
  with pytest.raises(ValueError, match="Fields not found"):
    sim.expand_model_fields(("no_such_field",))


def test_randomize_field_scale_uses_defaults():
  from mjlab_b200.events import randomize_field
  from mjlab_b200.managers import SceneEntityCfg
  env = _dr_env(2)
  body_idx = env.scene["robot"].indexing.body_ids[0]
  default_mass = env.sim.get_default_field("body_mass")[body_idx].item()
  for _ in range(3):
    randomize_field(env, env_ids=None, field="body_mass", ranges=(2.0, 2.0), operation="scale", asset_cfg=SceneEntityCfg("robot", body_ids=[0]))
  assert abs(env.sim.model.body_mass[0, body_idx].item() - default_mass * 2.0) < 1e-5


def test_randomize_field_scale_partial_axes():
  from mjlab_b200.events import randomize_field
  from mjlab_b200.managers import SceneEntityCfg
  env = _dr_env(2)
  geom_idx = env.scene["robot"].indexing.geom_ids[0]
  default = env.sim.get_default_field("geom_friction")[geom_idx].clone()
  randomize_field(env, env_ids=None, field="geom_friction", ranges=(2.0, 2.0), operation="scale",
                  asset_cfg=SceneEntityCfg("robot", geom_ids=[0]), axes=[0])
  final = env.sim.model.geom_friction[0, geom_idx]
  assert abs(final[0] - default[0] * 2.0) < 1e-5 and abs(final[1] - default[1]) < 1e-5 and abs(final[2] - default[2]) < 1e-5


def test_randomize_field_single_env_without_expand():
  from mjlab_b200.events import randomize_field
  from mjlab_b200.managers import SceneEntityCfg
  env = _dr_env(1, expand=None)
  body_idx = env.scene["robot"].indexing.body_ids[0]
  original = env.sim.model.body_mass[0, body_idx].item()
  for _ in range(2):
    randomize_field(env, env_ids=None, field="body_mass", ranges=(2.0, 2.0), operation="scale", asset_cfg=SceneEntityCfg("robot", body_ids=[0]))
    assert abs(env.sim.model.body_mass[0, body_idx].item() - original * 2.0) < 1e-5


# ---- box-box on the GPU ------------------------------------------------------------------------------------------------
def test_gpu_box_pile_matches_oracle():
  from oracle import mjstep as O
  m = _model(BOX_SCENE)
  nw, nstep = 8, 12
  rng = np.random.default_rng(5)
  from mjlab_b200.sim import SimulationCfg
  sim = _sim(nw, SimulationCfg(nconmax=64, njmax=300), m)
  sim.use_cuda_graph = False
  ds = []
  for w in range(nw):
    d = O.Data(m)
    d.qpos[:] = m.qpos0
    for b in range(4):
      d.qpos[7 * b: 7 * b + 3] += rng.normal(size=3) * 0.004 * w
      q = d.qpos[7 * b + 3: 7 * b + 7] + rng.normal(size=4) * 0.01 * w
      d.qpos[7 * b + 3: 7 * b + 7] = q / np.linalg.norm(q)
    d.qvel[:] = rng.normal(size=m.nv) * 0.1
    ds.append(d)
  sim.data.qpos[:] = torch.as_tensor(np.stack([d.qpos for d in ds]), dtype=torch.float32, device=DEV)
  sim.data.qvel[:] = torch.as_tensor(np.stack([d.qvel for d in ds]), dtype=torch.float32, device=DEV)
  nbb = same = 0
  for s in range(nstep):
    sim.step()
    ncon = sim.data.ncon.cpu().numpy()
    geom = sim.data.contact_geom.cpu().numpy()
    for w, d in enumerate(ds):
      O.step(m, d, 64, 300)
      if ncon[w] == len(d.contacts) and geom[w, :ncon[w]].tolist() == [list(c.geom) for c in d.contacts]:
        same += 1
      nbb += sum(1 for c in d.contacts if m.geom_type[c.geom[0]] == 6 and m.geom_type[c.geom[1]] == 6)
  # float32 against float64: a contact entering one step apart is allowed in a few (world, step) cells, trajectories must agree
  assert same >= 0.9 * nw * nstep and nbb >= 100
  qpos = sim.data.qpos.cpu().numpy()
  for w, d in enumerate(ds):
    assert np.abs(qpos[w] - d.qpos).max() < 2e-3, w


def test_go1_rough_env_steps():
  """Mjlab-Velocity-Rough-Unitree-Go1 (refused in round 1): the env runs on the box terrain, trunk contacts included."""
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  env = ManagerBasedRlEnv(VelocityEnvCfg(robot="go1", terrain="rough", num_envs=64, seed=3), DEV)
  obs, _ = env.reset()
  g = torch.Generator(device=DEV).manual_seed(0)
  for _ in range(60):
    obs, rew, term, trunc, _ = env.step(torch.rand(64, env.vcfg.nj, device=DEV, generator=g) * 2 - 1)
  assert torch.isfinite(obs["policy"]).all() and torch.isfinite(rew).all()
  assert int(env.sim.data.ncon.sum()) > 0 and int(env.sim.data.ncon_dropped.sum()) == 0


def test_env_scene_sensor_views():
  """`env.scene[...]` hands out the reference's sensor objects (contact_sensor.py:231-281, builtin_sensor.py:327-340) over the
  buffers the fused kernels write: feet air-time state, contact found / force slices, IMU slices."""
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  env = ManagerBasedRlEnv(VelocityEnvCfg(robot="go1", num_envs=32, seed=2), DEV)
  env.reset()
  feet = env.scene["feet_ground_contact"]
  assert env.scene["feet_ground_contact"] is feet
  saw_first = False
  for _ in range(40):
    env.step(torch.zeros(32, env.vcfg.nj, device=DEV))
    d = feet.data
    assert d.found.shape == (32, 4) and d.force.shape == (32, 4, 3)
    assert torch.equal(d.current_air_time, env._buf["air_state"][:, 0]) and torch.equal(d.current_contact_time, env._buf["air_state"][:, 2])
    fc = feet.compute_first_contact(env.step_dt)
    assert torch.equal(fc, (d.current_contact_time > 0) & (d.current_contact_time < env.step_dt + 1e-8))
    saw_first |= bool(fc.any())
    in_contact = d.found > 0
    assert torch.all(d.current_contact_time[in_contact] > 0) and torch.all(d.current_air_time[in_contact] == 0)
  assert saw_first and bool((feet.data.found > 0).any())
  imu = env.scene["robot/imu_ang_vel"]
  i = env.sim.mj_model.sensor_names.index("robot/imu_ang_vel")
  a = int(env.sim.mj_model.sensor_adr[i])
  assert torch.equal(imu.data, env.sim.data.sensordata[:, a: a + 3])
  assert env.scene["nonfoot_ground_touch"].data.found.shape[0] == 32
  with pytest.raises(KeyError):
    env.scene["no_such_thing"]
  assert env.scene.env_origins.shape == (32, 3)


def test_reset_consumes_its_own_draws():
  """reset() and the first step() shared draw counter 0.  Repeated reset() calls must give different states, and the
  noise of the first step's observations must differ from reset()'s."""
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  env = ManagerBasedRlEnv(VelocityEnvCfg(robot="go1", num_envs=16, seed=11), DEV)
  env.reset()
  q1 = env.sim.data.qpos.clone()
  c1 = int(env._buf["step_counter"][0])
  env.reset()
  q2 = env.sim.data.qpos.clone()
  assert int(env._buf["step_counter"][0]) == c1 + 1
  assert not torch.equal(q1[:, :2], q2[:, :2])                     # a fresh x / y / yaw draw on every reset()
  env.step(torch.zeros(16, env.vcfg.nj, device=DEV))
  assert int(env._buf["step_counter"][0]) == c1 + 2


def test_nan_guard_covers_the_fused_env_step(tmp_path):
  """NanGuard was only wired into Simulation.step(); the env's fused step bypasses it."""
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg, _default_sim_cfg
  from mjlab_b200.nan_guard import NanGuardCfg
  sc = _default_sim_cfg()
  sc.nan_guard = NanGuardCfg(enabled=True, buffer_size=4, output_dir=str(tmp_path), max_envs_to_dump=2)
  env = ManagerBasedRlEnv(VelocityEnvCfg(robot="go1", num_envs=8, seed=1, sim=sc), DEV)
  env.reset()
  z = torch.zeros(8, env.vcfg.nj, device=DEV)
  for _ in range(3):
    env.step(z)
  assert not list(tmp_path.glob("nan_dump_*.npz"))
  env.sim.data.qvel[5, 7] = float("nan")
  env.step(z)
  dump = np.load(tmp_path / "nan_dump_latest.npz", allow_pickle=True)
  meta = dump["_metadata"].item()
  assert 5 in meta["nan_env_ids"] and meta["buffer_size"] == 4


def _task_cfg(task, play=False):
  import json
  from pathlib import Path
  return json.loads((Path(__file__).resolve().parent / "golden" / "task_cfg" / f"{task}{'-play' if play else ''}.json").read_text())


def test_env_from_the_reference_manager_cfg_equals_the_builtin_task():
  """`VelocityEnvCfg.from_manager_cfg(load_env_cfg(task))`: the env built from the reference's own configuration object (its JSON
  dump, tests/golden/task_cfg) steps bit-identically to the built-in task of the same name."""
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  d = _task_cfg("Mjlab-Velocity-Flat-Unitree-Go1")
  a = ManagerBasedRlEnv(VelocityEnvCfg.from_task("Mjlab-Velocity-Flat-Unitree-Go1", num_envs=64, seed=5), DEV)
  torch.manual_seed(123)
  b = ManagerBasedRlEnv(VelocityEnvCfg.from_manager_cfg(d, "go1", "flat", num_envs=64, seed=5), DEV)
  assert b.max_episode_length == a.max_episode_length == 1000 and b.step_dt == a.step_dt
  outs = []
  for env in (a, b):
    torch.manual_seed(9)       # the startup foot-friction event draws from torch's generator
    env2 = env
    g = torch.Generator(device=DEV).manual_seed(1)
    env2.sim.model.geom_friction[:] = a.sim.model.geom_friction      # same startup DR on both
    env2.reset()
    env2.episode_length_buf = torch.randint(0, 1000, (64,), device=DEV, generator=g)
    for _ in range(12):
      obs, rew, term, trunc, _ = env2.step(torch.rand(64, env2.vcfg.nj, device=DEV, generator=g) * 2 - 1)
    outs.append((obs["policy"].clone(), obs["critic"].clone(), rew.clone(), term.clone(), env2.sim.data.qpos.clone()))
  for x, y in zip(*outs):
    assert torch.equal(x, y)


def test_env_consumes_edited_manager_cfg_with_history_and_delay():
  """A user-edited reference cfg: a reward re-weighted and one removed, no push event, 3-step history on the critic group and a
  delayed, clipped policy term -> consumed (weights, flags) and post-processed (observation shapes, delay, history back-fill)."""
  from mjlab_b200 import envcfg
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  d = _task_cfg("Mjlab-Velocity-Flat-Unitree-Go1")
  d["rewards"]["foot_slip"]["weight"] = -0.7
  d["rewards"].pop("soft_landing")
  d["events"].pop("push_robot")
  d["observations"]["critic"]["history_length"] = 3
  d["observations"]["policy"]["enable_corruption"] = False
  d["observations"]["policy"]["terms"]["base_ang_vel"].update(delay_min_lag=2, delay_max_lag=2)
  d["observations"]["policy"]["terms"]["joint_vel"].update(clip=[-1.0, 1.0], scale=0.5)
  # two more startup domain-randomisation events (any field of randomize_field, events.py:292-354): trunk mass scaled, hip damping set
  ev = lambda field, ranges, op, **names: dict(func="randomize_field", mode="startup", interval_range_s=None, params=dict(  # noqa: E731
    field=field, ranges=ranges, operation=op, asset_cfg=dict(name="robot", **names)))
  d["events"]["trunk_mass"] = ev("body_mass", [0.8, 1.2], "scale", body_names=["trunk"])
  d["events"]["hip_damping"] = ev("dof_damping", [0.5, 1.5], "abs", joint_names=[".*_hip_joint"])
  env = ManagerBasedRlEnv(VelocityEnvCfg.from_manager_cfg(d, "go1", "flat", num_envs=32, seed=2), DEV)
  mj = env.sim.mj_model
  trunk = [i for i, n in enumerate(mj.body_names) if n.endswith("trunk")][0]
  ratio = env.sim.model.body_mass[:, trunk] / float(mj.body_mass[trunk])
  assert ratio.shape == (32,) and float(ratio.min()) >= 0.8 and float(ratio.max()) <= 1.2 and float(ratio.std()) > 0.02
  hips = [int(mj.jnt_dofadr[i]) for i, n in enumerate(mj.jnt_names) if n.endswith("_hip_joint")]
  dmp = env.sim.model.dof_damping[:, hips]
  assert dmp.shape == (32, 4) and float(dmp.min()) >= 0.5 and float(dmp.max()) <= 1.5 and float(dmp.std()) > 0.05
  fr = env.sim.model.geom_friction[:, [i for i, n in enumerate(mj.geom_names) if n.endswith("_foot_collision")], 0]
  assert float(fr.min()) >= 0.3 and float(fr.max()) <= 1.2 and float(fr.std()) > 0.05        # the cfg's own foot_friction event
  k = envcfg.REWARD_NAMES.index
  assert env.vcfg.reward_weights[k("foot_slip")] == -0.7 and env.vcfg.reward_weights[k("soft_landing")] == 0 and not env.vcfg.push_enabled
  obs, _ = env.reset()
  nj = env.vcfg.nj
  assert obs["policy"].shape == (32, 48) and obs["critic"].shape == (32, 3 * 72)
  raw_hist = []
  z = torch.zeros(32, nj, device=DEV)
  for t in range(6):
    obs, rew, term, trunc, _ = env.step(z)
    raw = env._raw_obs["policy"].clone()
    raw_hist.append(raw)
    alive = ~(term | trunc)
    # clipped + scaled term
    jv = obs["policy"][:, 9 + nj: 9 + 2 * nj]
    torch.testing.assert_close(jv, raw[:, 9 + nj: 9 + 2 * nj].clip(-1, 1) * 0.5)
    # delayed term: the value of two control steps ago (envs that have not been reset since)
    if t >= 2:
      keep = alive & alive_prev & alive_prev2
      torch.testing.assert_close(obs["policy"][keep, 3:6], raw_hist[t - 2][keep, 3:6])
    alive_prev2 = alive_prev if t >= 1 else alive
    alive_prev = alive
    # critic history, term-major, oldest first: the newest slot of the first term is this step's raw critic term
    torch.testing.assert_close(obs["critic"][:, 6:9], env._raw_obs["critic"][:, 0:3])
  with pytest.raises(envcfg.UnsupportedTermError):
    d["rewards"]["x"] = dict(func="my_power_cost", weight=-1.0, params={})
    ManagerBasedRlEnv(VelocityEnvCfg.from_manager_cfg(d, "go1", "flat", num_envs=8), DEV)


def test_env_consumes_custom_observation_groups():
  """Observation groups of a consumed cfg are gathers from the kernels' term bank: the usual sim-to-real edit (no base linear
  velocity for the policy), a reordered group and an extra group; noisy terms come from the corrupted copy, the others clean."""
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  d = _task_cfg("Mjlab-Velocity-Flat-Unitree-Go1")
  pt = d["observations"]["policy"]["terms"]
  pt.pop("base_lin_vel")
  d["observations"]["policy"]["terms"] = {"command": pt["command"], **{k: v for k, v in pt.items() if k != "command"}}
  d["observations"]["feet"] = dict(d["observations"]["critic"], terms={k: d["observations"]["critic"]["terms"][k] for k in ("foot_contact", "foot_height")})
  d["observations"]["feet"]["terms"]["root_ang_vel"] = dict(d["observations"]["critic"]["terms"]["foot_height"], func="base_ang_vel", params={})
  env = ManagerBasedRlEnv(VelocityEnvCfg.from_manager_cfg(d, "go1", "flat", num_envs=48, seed=4), DEV)
  nj, nf = env.vcfg.nj, env.vcfg.nfeet
  om = env.observation_manager
  assert om.active_terms["policy"] == ["command", "base_ang_vel", "projected_gravity", "joint_pos", "joint_vel", "actions"]
  assert om.group_obs_dim == {"policy": (9 + 3 * nj,), "critic": (12 + 3 * nj + 6 * nf,), "feet": (2 * nf + 3,)}
  obs, _ = env.reset()
  g = torch.Generator(device=DEV).manual_seed(0)
  for _ in range(5):
    obs, rew, term, trunc, _ = env.step(torch.rand(48, nj, device=DEV, generator=g) * 2 - 1)
    raw_p, raw_c = env._raw_obs["policy"], env._raw_obs["critic"]
    assert obs["policy"].shape == (48, 9 + 3 * nj) and obs["feet"].shape == (48, 2 * nf + 3) and obs["critic"].shape == raw_c.shape
    torch.testing.assert_close(obs["policy"][:, 0:3], raw_c[:, 9 + 3 * nj: 12 + 3 * nj])           # command: no noise, clean copy
    torch.testing.assert_close(obs["policy"][:, 3:6], raw_p[:, 3:6])                               # noisy imu angular velocity
    torch.testing.assert_close(obs["policy"][:, 9: 9 + nj], raw_p[:, 9: 9 + nj])                   # noisy joint positions
    assert float((raw_p[:, 3:6] - raw_c[:, 3:6]).abs().max()) > 1e-3                               # (the two copies do differ)
    torch.testing.assert_close(obs["feet"][:, : 2 * nf], torch.cat([raw_c[:, 12 + 3 * nj + 2 * nf: 12 + 3 * nj + 3 * nf], raw_c[:, 12 + 3 * nj: 12 + 3 * nj + nf]], 1))
    # base_ang_vel (envs/mdp/observations.py:31-35) = EntityData.root_link_ang_vel_b, the same quantity through the entity view
    torch.testing.assert_close(obs["feet"][:, 2 * nf:], env.robot.data.root_link_ang_vel_b, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(obs["critic"], raw_c)
