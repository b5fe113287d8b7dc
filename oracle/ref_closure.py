"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

Runs the reference's OWN env step -- ``ManagerBasedDiffRLEnv.step`` / ``_reset_idx`` / ``detach``
(L/envs/manager_based_diff_rl_env.py:160-267,362-416), ``DiffActionManager.process_action``
(L/managers/action_manager.py:31-52), ``LossManager.compute`` (L/managers/loss_manager.py:84-104), ``DiffActions``
(QD/mdp/diff_action.py), ``RacingCommand`` (QD/mdp/commands.py:166-401) and the free MDP term functions of
QD/mdp/{rewards,observation,termination,losses,events,curriculums}.py -- UNMODIFIED, executed where they lie under
/root/reference, so that tests/test_oracle_vs_reference_env.py can pin oracle/racing_oracle.py (rows a1, a5-a10 of
SURVEY.md §8a and the step order of §3.3) against the reference's code rather than against a reading of it.  The same harness
runs the reach-target tasks (``make_reference_reach_env``: ``UniformWorldPoseCommand`` QD/mdp/commands.py:33-134 and the terms of
QD/reach_target_lv_env.py / reach_target_ctbr_env.py) for tests/test_reach_oracle_vs_reference_env.py, records the golden vectors of
tests/golden/make_{ref_env,reach_env,c1}_golden.py, and carries the reference's own trainers (tests/test_bptt_trainer_vs_reference.py,
tests/test_ppo_trainer_on_kernels_vs_reference.py).

What is NOT the reference here, and therefore restated (marked [isaac] below): Isaac Lab's manager base classes
(``ActionManager``, ``CommandManager`` / ``CommandTerm``, ``RewardManager``, ``TerminationManager``, ``CurriculumManager``,
``ObservationManager``, ``EventManager``, ``TerrainImporter.update_env_origins``, ``mdp.time_out``, ``sample_uniform``) --
third-party code (omni-isaac-lab >= 0.27.15, pyproject.toml:32) absent from /root/reference; they follow Isaac Lab's
published behaviour as SURVEY.md §3.3 / Appendix B record it.  The simulator is the PhysX-free closure of SURVEY.md
Appendix A.1: ``sim.step`` makes the nominal model's next state the simulated truth.  Term parameters are the values of
QD/racing_ctbr_env.py (cited at each use; tests/test_oracle_vs_reference_env.py greps the file for them).

``is_differentiable_physics`` is always True on the reference side: Appendix A.1 re-anchors the nominal model on the simulated
truth every step, which is what the reference's step does when the flag is set (manager_based_diff_rl_env.py:206-212).  With the
flag False the reference never re-anchors -- its nominal model free-runs beside PhysX and nothing reads it -- a mode without a
PhysX-free meaning; the oracle (and the kernels) therefore align in both modes and only export the losses when asked.

Random numbers: the reference draws from torch's global generator.  ``replay_*`` below repeat, after the same
``torch.manual_seed``, exactly the calls the reference made (same order, shapes and dtypes) and place the values in the
oracle's explicit ``rnd`` rows, so both sides see identical noise.
"""
from __future__ import annotations

import os
import sys
import types

import torch

from generalizableracing_b200 import layout as L_
from . import isaac_math as M
from . import ref_modules as RM

_QD = RM._QD
_L = RM._L


# ----------------------------------------------------------------------------------------------- [isaac] stubs
class SceneEntityCfg:
    def __init__(self, name, body_names=None, **kw):
        self.name = name
        self.body_names = body_names
        self.body_ids = slice(None)


class ManagerTermBase:
    pass


class ManagerTermBaseCfg:
    params: dict = {}


class ManagerBase:
    def __init__(self, cfg, env):
        self.cfg = cfg
        self._env = env
        self._prepare_terms()

    num_envs = property(lambda self: self._env.num_envs)
    device = property(lambda self: self._env.device)

    def _resolve_common_term_cfg(self, term_name, term_cfg, min_argc=1):
        pass


class ActionTerm:
    def __init__(self, cfg, env):
        self.cfg = cfg
        self._env = env
        self._asset = env.scene[cfg.asset_name]

    num_envs = property(lambda self: self._env.num_envs)
    device = property(lambda self: self._env.device)


class ActionManager(ManagerBase):
    def __init__(self, cfg, env):
        self._terms = {}
        super().__init__(cfg, env)
        self._action = torch.zeros(self.num_envs, self.total_action_dim)
        self._prev_action = torch.zeros_like(self._action)

    def _prepare_terms(self):
        for name, term_cfg in self.cfg.items():
            self._terms[name] = term_cfg.class_type(term_cfg, self._env)

    total_action_dim = property(lambda self: sum(t.action_dim for t in self._terms.values()))
    action = property(lambda self: self._action)
    prev_action = property(lambda self: self._prev_action)

    def get_term(self, name):
        return self._terms[name]

    def reset(self, env_ids=None):
        self._prev_action[env_ids] = 0.0
        self._action[env_ids] = 0.0
        for term in self._terms.values():
            term.reset(env_ids=env_ids)
        return {}

    def apply_action(self):
        for term in self._terms.values():
            term.apply_actions()


class CommandTerm:
    def __init__(self, cfg, env):
        self.cfg = cfg
        self._env = env
        self.metrics = dict()
        self.time_left = torch.zeros(self.num_envs)
        self.command_counter = torch.zeros(self.num_envs, dtype=torch.long)

    num_envs = property(lambda self: self._env.num_envs)
    device = property(lambda self: self._env.device)

    def reset(self, env_ids=None):
        extras = {}
        for name, value in self.metrics.items():
            extras[name] = torch.mean(value[env_ids]).item()
            value[env_ids] = 0.0
        self.command_counter[env_ids] = 0
        self._resample(env_ids)
        return extras

    def compute(self, dt):
        self._update_metrics()
        self.time_left -= dt
        ids = (self.time_left <= 0.0).nonzero().flatten()
        if len(ids) > 0:
            self._resample(ids)
        self._update_command()

    def _resample(self, env_ids):
        if len(env_ids) != 0:
            self.time_left[env_ids] = self.time_left[env_ids].uniform_(*self.cfg.resampling_time_range)
            self._resample_command(env_ids)
            self.command_counter[env_ids] += 1


class CommandManager:
    def __init__(self, terms, env):
        self._terms = terms
        self._env = env
        self.last_achieved = None

    def get_term(self, name):
        return self._terms[name]

    def get_command(self, name):
        return self._terms[name].command

    def compute(self, dt):
        for term in self._terms.values():
            # recorded for the random-stream replay only (which envs the term is about to draw noise / new targets for)
            if hasattr(term, "gate_pose_gt_w"):
                term_pos = term.gate_pose_gt_w[:, :3] - term.robot.data.root_state_w[:, :3]
                self.last_achieved = torch.norm(term_pos, dim=-1) < term.cfg.update_threshold
            self.last_timer_ids = ((term.time_left - dt) <= 0.0).nonzero().flatten()
            term.compute(dt)

    def reset(self, env_ids=None):
        extras = {}
        for name, term in self._terms.items():
            for metric, value in term.reset(env_ids=env_ids).items():
                extras[f"Metrics/{name}/{metric}"] = value
        return extras


class _TermCfg:
    def __init__(self, func, params=None, weight=None, time_out=False):
        self.func, self.params, self.weight, self.time_out = func, dict(params or {}), weight, time_out


class RewardManager:
    def __init__(self, terms, env):
        self._env = env
        self._term_names = list(terms.keys())
        self._term_cfgs = list(terms.values())
        n = env.num_envs
        self._episode_sums = {k: torch.zeros(n) for k in self._term_names}
        self._reward_buf = torch.zeros(n)
        self._step_reward = torch.zeros(n, len(self._term_names))

    def compute(self, dt):
        self._reward_buf[:] = 0.0
        for name, cfg in zip(self._term_names, self._term_cfgs):
            if cfg.weight == 0.0:
                continue
            value = cfg.func(self._env, **cfg.params) * cfg.weight * dt
            self._reward_buf += value
            self._episode_sums[name] += value
            self._step_reward[:, self._term_names.index(name)] = value / dt
        return self._reward_buf

    def reset(self, env_ids=None):
        extras = {}
        for key in self._episode_sums:
            extras["Episode_Reward/" + key] = torch.mean(self._episode_sums[key][env_ids]) / self._env.max_episode_length_s
            self._episode_sums[key][env_ids] = 0.0
        return extras


class TerminationManager:
    def __init__(self, terms, env):
        self._env = env
        self._terms = terms
        self._truncated = torch.zeros(env.num_envs, dtype=torch.bool)
        self._terminated = torch.zeros_like(self._truncated)

    dones = property(lambda self: self._truncated | self._terminated)
    time_outs = property(lambda self: self._truncated)
    terminated = property(lambda self: self._terminated)

    def compute(self):
        self._truncated[:] = False
        self._terminated[:] = False
        for cfg in self._terms.values():
            value = cfg.func(self._env, **cfg.params)
            if cfg.time_out:
                self._truncated |= value
            else:
                self._terminated |= value
        return self._truncated | self._terminated

    def reset(self, env_ids=None):
        return {}


class CurriculumManager:
    def __init__(self, terms, env):
        self._env = env
        self._terms = terms
        self._state = {k: None for k in terms}

    def compute(self, env_ids=None):
        for name, cfg in self._terms.items():
            self._state[name] = cfg.func(self._env, env_ids, **cfg.params)

    def reset(self, env_ids=None):
        return {"Curriculum/" + k: v for k, v in self._state.items() if v is not None}


class ObservationManager:
    def __init__(self, groups, env):
        self._env = env
        self._groups = groups

    def compute(self):
        return {g: torch.cat([c.func(self._env, **c.params).clone() for c in terms.values()], dim=-1) for g, terms in self._groups.items()}

    def reset(self, env_ids=None):
        return {}


class EventManager:
    def __init__(self, terms, env):
        self._env = env
        self._terms = terms                 # mode -> {name: _TermCfg}

    available_modes = property(lambda self: list(self._terms.keys()))

    def apply(self, mode, env_ids=None, dt=None, global_env_step_count=None):
        for cfg in self._terms.get(mode, {}).values():
            cfg.func(self._env, env_ids, **cfg.params)

    def reset(self, env_ids=None):
        return {}


class _Null:
    """Recorder manager / visualisers / anything whose calls have no effect in the closure."""
    active_terms = ()

    def __getattr__(self, name):
        return lambda *a, **k: {}


class Recorder(_Null):
    """[isaac] RecorderManager with no terms; ``pre_reset_hook(env_ids)`` lets a test act between the step's reward
    computation and ``_reset_idx`` (manager_based_diff_rl_env.py:232-236)."""
    pre_reset_hook = None

    def record_pre_reset(self, env_ids):
        if self.pre_reset_hook is not None:
            self.pre_reset_hook(env_ids)


def time_out(env):
    return env.episode_length_buf >= env.max_episode_length


def sample_uniform(lower, upper, size, device):
    if isinstance(size, int):
        size = (size,)
    return torch.rand(*size, device=device) * (upper - lower) + lower


def quat_unique(q):
    return torch.where(q[..., 0:1] < 0, -q, q)


def compute_pose_error(t01, q01, t02, q02, rot_error_type="axis_angle"):
    # [isaac] position part only: the racing call site (QD/mdp/commands.py:247-257) uses nothing else, the reach one (:98-108)
    # only logs the norm of the rotation part ("orientation_error", not an input of any term)
    return t02 - t01, torch.zeros_like(t01)


class ManagerBasedRLEnvBase:
    num_envs = property(lambda self: self.scene.num_envs)
    device = property(lambda self: self.scene.device)
    physics_dt = property(lambda self: self.cfg.sim.dt)
    step_dt = property(lambda self: self.cfg.sim.dt * self.cfg.decimation)


class _GymEnv:
    pass


# ----------------------------------------------------------------------------------------------- closure "simulator"
class RobotData:
    def __init__(self, n, default_pos):
        self.root_pos_w = torch.zeros(n, 3)
        self.root_quat_w = torch.zeros(n, 4)
        self.root_quat_w[:, 0] = 1.0
        self.root_lin_vel_w = torch.zeros(n, 3)
        self.root_ang_vel_w = torch.zeros(n, 3)
        self.body_ang_acc_w = torch.zeros(n, 1, 3)
        self.body_lin_acc_w = torch.zeros(n, 1, 3)
        self.default_root_state = torch.zeros(n, 13)
        self.default_root_state[:, :3] = torch.tensor(default_pos)
        self.default_root_state[:, 3] = 1.0

    root_state_w = property(lambda s: torch.hstack([s.root_pos_w, s.root_quat_w, s.root_lin_vel_w, s.root_ang_vel_w]))
    root_lin_vel_b = property(lambda s: M.quat_rotate_inverse(s.root_quat_w, s.root_lin_vel_w))
    root_com_lin_vel_b = root_lin_vel_b
    root_ang_vel_b = property(lambda s: M.quat_rotate_inverse(s.root_quat_w, s.root_ang_vel_w))
    root_vel_w = property(lambda s: torch.cat([s.root_lin_vel_w, s.root_ang_vel_w], dim=-1))


class Robot:
    is_initialized = True
    device = "cpu"

    def __init__(self, n, mass, default_pos):
        self.data = RobotData(n, default_pos)
        masses = torch.full((n, 1), mass)
        self.root_physx_view = types.SimpleNamespace(get_masses=lambda: masses)

    def find_bodies(self, names, preserve_order=False):
        return ([0], ["body"]) if names == "body" else ([1, 2, 3, 4], ["m1_prop", "m2_prop", "m3_prop", "m4_prop"])

    def set_external_force_and_torque(self, *a, **k):
        pass

    def write_root_link_pose_to_sim(self, pose, env_ids):
        self.data.root_pos_w[env_ids] = pose[:, :3]
        self.data.root_quat_w[env_ids] = pose[:, 3:7]

    def write_root_com_velocity_to_sim(self, vel, env_ids):
        self.data.root_lin_vel_w[env_ids] = vel[:, :3]
        self.data.root_ang_vel_w[env_ids] = vel[:, 3:]
        self.data.body_ang_acc_w[env_ids] = 0.0            # closure A.1: accelerations restart at zero
        self.data.body_lin_acc_w[env_ids] = 0.0


class Terrain:
    """[isaac] TerrainImporter in curriculum mode; the level re-draw of ``update_env_origins`` takes its uniform from
    ``pending_level_u`` (rows aligned with env ids) so that it can be shared with the oracle's RND_LEVEL column."""

    def __init__(self, table, n, startup_rnd, max_init_level, num_gate):
        self.extras = {"gate_pose": torch.tensor(table.gate_pose), "next_gate_id": torch.tensor(table.next_gate_id, dtype=torch.long)}
        self.terrain_origins = torch.tensor(table.terrain_origins)
        self.max_terrain_level = table.num_levels
        mi = min(max_init_level, table.num_levels - 1)
        self.terrain_levels = torch.floor(startup_rnd[:, L_.SRND_LEVEL].double() * (mi + 1)).long().clamp(max=mi)
        self.terrain_types = torch.div(torch.arange(n), (n / table.num_types), rounding_mode="floor").to(torch.long)
        self.env_origins = self.terrain_origins[self.terrain_levels, self.terrain_types].clone()
        self.cfg = types.SimpleNamespace(terrain_generator=types.SimpleNamespace(sub_terrains={"circular": types.SimpleNamespace(num_gate=num_gate)}))
        self.pending_level_u = torch.zeros(n)

    def update_env_origins(self, env_ids, move_up, move_down):
        lv = self.terrain_levels[env_ids] + 1 * move_up - 1 * move_down
        rand_lv = torch.floor(self.pending_level_u[env_ids].double() * self.max_terrain_level).long().clamp(max=self.max_terrain_level - 1)
        self.terrain_levels[env_ids] = torch.where(lv >= self.max_terrain_level, rand_lv, torch.clip(lv, 0))
        self.env_origins[env_ids] = self.terrain_origins[self.terrain_levels[env_ids], self.terrain_types[env_ids]]


class Scene:
    device = "cpu"

    def __init__(self, n, robot, terrain):
        self.num_envs = n
        self._robot = robot
        self.terrain = terrain

    env_origins = property(lambda self: self.terrain.env_origins)

    def __getitem__(self, name):
        assert name == "robot", name
        return self._robot

    def write_data_to_sim(self):
        pass

    def update(self, dt):
        pass

    def reset(self, env_ids=None):
        pass


class Sim:
    """Closure A.1: the simulated truth after a step is the nominal model's next state (idempotent over the
    ``decimation`` calls of one env step)."""

    def __init__(self, env):
        self._env = env

    def has_gui(self):
        return False

    def has_rtx_sensors(self):
        return False

    def forward(self):
        pass

    def render(self):
        pass

    def step(self, render=False):
        env = self._env
        term = env.action_manager.get_term("force_torque")
        d = term.drone_dynamics
        w = term.ang_vel_b.detach()                                   # pre-step body rates (QD/mdp/diff_action.py:136)
        torque = term.processed_actions[:, 1:4]
        alpha = (d.inertia_inv @ torque.unsqueeze(-1)).squeeze(-1) - (
            d.inertia_inv @ torch.linalg.cross(w, (d.inertia @ w.unsqueeze(-1)).squeeze(-1)).unsqueeze(-1)).squeeze(-1)
        nom = term.nominal_next_state.detach()
        data = env.scene["robot"].data
        data.root_pos_w = nom[:, :3] + env.scene.env_origins
        data.root_quat_w = nom[:, 3:7].clone()
        data.root_lin_vel_w = nom[:, 7:10].clone()
        data.root_ang_vel_w = nom[:, 10:13].clone()
        data.body_ang_acc_w = M.quat_rotate(data.root_quat_w, alpha).unsqueeze(1)
        data.body_lin_acc_w = term.a.detach().clone().unsqueeze(1)


# ----------------------------------------------------------------------------------------------- loading
_cache = {}


def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        if "." in name:
            parent, leaf = name.rsplit(".", 1)
            setattr(_mod(parent), leaf, m)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def load():
    """Namespace with the reference's env class, managers and MDP term modules (unmodified, executed in place)."""
    if "ns" in _cache:
        return _cache["ns"]
    ref = RM.load()                                        # dynamics / controllers (also installs the math stub)
    math_mod = sys.modules["omni.isaac.lab.utils.math"]
    for fn in ("quat_mul", "quat_rotate", "quat_rotate_inverse", "matrix_from_quat", "quat_inv", "quat_from_euler_xyz",
               "euler_xyz_from_quat", "wrap_to_pi"):
        setattr(math_mod, fn, getattr(M, fn))
    math_mod.sample_uniform = sample_uniform
    math_mod.quat_unique = quat_unique
    math_mod.compute_pose_error = compute_pose_error
    math_mod.yaw_quat = math_mod.subtract_frame_transforms = None   # imported by name, never called on the racing path
    cls = lambda n: type(n, (), {})
    _mod("omni.isaac.lab.utils", configclass=lambda c: c, math=math_mod)
    _mod("omni.isaac.lab.managers", SceneEntityCfg=SceneEntityCfg, ManagerBase=ManagerBase, ManagerTermBase=ManagerTermBase,
         CommandTerm=CommandTerm, ActionTerm=ActionTerm, ActionTermCfg=cls("ActionTermCfg"), CommandManager=CommandManager,
         CurriculumManager=CurriculumManager, RewardManager=RewardManager, TerminationManager=TerminationManager)
    _mod("omni.isaac.lab.managers.manager_base", ManagerBase=ManagerBase)
    _mod("omni.isaac.lab.managers.manager_term_cfg", ActionTermCfg=cls("ActionTermCfg"), ManagerTermBaseCfg=ManagerTermBaseCfg)
    _mod("omni.isaac.lab.managers.action_manager", ActionTerm=ActionTerm, ActionManager=ActionManager)
    _mod("omni.isaac.lab.managers.command_manager", CommandTerm=CommandTerm)
    _mod("omni.isaac.lab.assets", Articulation=cls("Articulation"), RigidObject=cls("RigidObject"), AssetBase=cls("AssetBase"))
    _mod("omni.isaac.lab.sensors", **{n: cls(n) for n in ("FrameTransformerData", "TiledCamera", "Camera", "RayCasterCamera",
                                                          "ContactSensor", "RayCasterCameraCfg")})
    _mod("omni.isaac.lab.markers", VisualizationMarkers=_Null)
    _mod("omni.isaac.lab.terrains", TerrainImporter=cls("TerrainImporter"))
    _mod("omni.isaac.lab.ui.widgets", ManagerLiveVisualizer=_Null)
    _mod("omni.isaac.lab.envs", ManagerBasedRLEnv=ManagerBasedRLEnvBase, ManagerBasedEnv=ManagerBasedRLEnvBase)
    _mod("omni.isaac.lab.envs.common", VecEnvStepReturn=tuple)
    _mod("omni.isaac.lab.envs.manager_based_rl_env", ManagerBasedRLEnv=ManagerBasedRLEnvBase)
    _mod("omni.isaac.lab.envs.mdp", time_out=time_out)
    _mod("omni.isaac.lab.envs.mdp.events", _randomize_prop_by_op=None)
    _mod("omni.isaac.version", get_version=lambda: ("closure",))
    _mod("omni.kit.app")
    if "prettytable" not in sys.modules:
        _mod("prettytable", PrettyTable=_Null)
    if "gymnasium" not in sys.modules:
        _mod("gymnasium", Env=_GymEnv)
    ctrl_mod = sys.modules["_gr_ref_controllers.controller_diff"]
    _mod("diff.lab.controllers", ThrustController=sys.modules["_gr_ref_controllers.thrust_controller_diff"].ThrustController,
         PSController=ctrl_mod.PSController, LVController=ctrl_mod.LVController, CTBRController=ctrl_mod.CTBRController)
    _mod("diff.lab.utils", get_uav_collision_num_ray=None, LATTICE_TENSOR=None)     # Warp mesh query: out of scope
    _mod("diff.lab.terrains", TerrainImporterCfg=cls("TerrainImporterCfg"))
    # L/managers: loss_term_cfg.py, loss_manager.py, action_manager.py under a synthetic package on the real directory
    mgr = _mod("_gr_ref_managers")
    mgr.__path__ = [os.path.join(RM.REF_ROOT, _L, "managers")]
    import importlib
    loss_cfg = importlib.import_module("_gr_ref_managers.loss_term_cfg")
    loss_mgr = importlib.import_module("_gr_ref_managers.loss_manager")
    act_mgr = importlib.import_module("_gr_ref_managers.action_manager")
    _mod("diff.lab.managers", LossManager=loss_mgr.LossManager, DiffActionManager=act_mgr.DiffActionManager, LossTermCfg=loss_cfg.LossTermCfg)
    # L/envs/manager_based_diff_rl_env.py (its cfg sibling needs Isaac's cfg classes: stubbed)
    envs = _mod("_gr_ref_envs")
    _mod("_gr_ref_envs.manager_based_diff_rl_env_cfg", ManagerBasedDiffRLEnvCfg=cls("ManagerBasedDiffRLEnvCfg"))
    env_mod = RM._load("_gr_ref_envs.manager_based_diff_rl_env", os.path.join(RM.REF_ROOT, _L, "envs/manager_based_diff_rl_env.py"))
    _mod("diff.lab.envs", ManagerBasedDiffRLEnv=env_mod.ManagerBasedDiffRLEnv)
    # QD/mdp term modules under a synthetic package on the real directory (mdp/__init__.py is not executed)
    mdp = _mod("_gr_ref_mdp")
    mdp.__path__ = [os.path.join(RM.REF_ROOT, _QD, "mdp")]
    mods = {n: importlib.import_module("_gr_ref_mdp." + n)
            for n in ("diff_action", "rewards", "commands", "observation", "termination", "losses", "events", "curriculums")}
    ns = types.SimpleNamespace(Env=env_mod.ManagerBasedDiffRLEnv, DiffActionManager=act_mgr.DiffActionManager, LossManager=loss_mgr.LossManager,
                               LossTermCfg=loss_cfg.LossTermCfg, CTBRController=ref.CTBRController, **mods)
    _cache["ns"] = ns
    return ns


# ----------------------------------------------------------------------------------------------- the closure env
def make_reference_env(cfg, table, num_envs, startup_rnd, seed):
    """The reference env of DiffLab-Quadcopter-CTBR-Racing for ``cfg.stage`` over the closure simulator.  Returns
    (env, startup_rnd') where startup_rnd' carries the draws the reference made while constructing (thr_est_error,
    startup DR) in the oracle's SRND_* columns."""
    ns = load()
    n = num_envs
    NS = types.SimpleNamespace
    env = object.__new__(ns.Env)
    env.cfg = NS(sim=NS(dt=cfg.sim_dt, gravity=(0.0, 0.0, -cfg.gravity), render_interval=cfg.decimation), decimation=cfg.decimation,
                 episode_length_s=cfg.episode_length_s, is_differentiable_physics=True, rerender_on_reset=False)
    robot = Robot(n, cfg.mass, cfg.default_root_pos)
    env.scene = Scene(n, robot, Terrain(table, n, startup_rnd, cfg.max_init_terrain_level, table.num_gates))
    env.sim = Sim(env)
    env.extras = {}
    env._sim_step_counter = 0
    env.common_step_counter = 0
    env.episode_length_buf = torch.zeros(n, dtype=torch.long)
    env.recorder_manager = Recorder()
    srnd = startup_rnd.clone()
    # -- managers, in the order of ManagerBasedDiffRLEnv.load_managers (L/envs/manager_based_diff_rl_env.py:108-135)
    pr, yr = cfg.cmd_noise_pos, cfg.cmd_noise_yaw
    cmd_cfg = NS(asset_name="robot", resampling_time_range=(20.0, 20.0), debug_vis=False, consecutive_commands=True, make_quat_unique=False,
                 noise_ranges=NS(pos_x=(-pr, pr), pos_y=(-pr, pr), pos_z=(-pr, pr), roll=(-0.0, 0.0), pitch=(-0.0, 0.0), yaw=(-yr, yr)),
                 add_noise=cfg.add_cmd_noise, update_threshold=cfg.update_threshold)                     # QD/racing_ctbr_env.py:98-121
    env.command_manager = CommandManager({}, env)
    env.command_manager._terms["next_gate_pose"] = ns.commands.RacingCommand(cmd_cfg, env)
    ctrl_cfg = RM.ctbr_cfg(cfg)
    ctrl_cfg.class_type = ns.CTBRController
    act_cfg = NS(class_type=ns.diff_action.DiffActions, asset_name="robot", rotor_names="m.*_prop", command_type="CTBRController",
                 controller_cfg=ctrl_cfg, gravity=9.81, random_drag=cfg.random_drag, action_lag=cfg.action_lag, sim2real_test=False,
                 max_thrust_weight_ratio=cfg.max_thrust_weight_ratio)                                     # QD/racing_ctbr_env.py:124-136
    torch.manual_seed(seed)
    env.action_manager = ns.DiffActionManager({"force_torque": act_cfg}, env)          # draws: randn(n) (diff_action.py:86)
    torch.manual_seed(seed)
    srnd[:, L_.SRND_THR_ERR] = torch.randn(n)
    T = _TermCfg
    R, O, Te, Lo, Ev, Cu = ns.rewards, ns.observation, ns.termination, ns.losses, ns.events, ns.curriculums
    name = {"command_name": "next_gate_pose"}
    obs = {                                                                                              # QD/racing_ctbr_env.py:139-174 (depth image: out of scope)
        "policy": {"base_lin_vel": T(O.modified_base_lin_vel, {"add_noise": True}), "base_orientation": T(O.base_orientation_r, {"add_noise": True}),
                   "target_cmd": T(O.modified_generated_commands, name), "last_action": T(O.modified_last_action, {"action_name": "force_torque"})},
        "critic": {"base_lin_vel": T(O.modified_base_lin_vel, {"add_noise": False}), "base_orientation": T(O.base_orientation_r, {"add_noise": False}),
                   "target_cmd": T(O.modified_generated_commands_gt, name), "last_action": T(O.modified_last_action, {"action_name": "force_torque"})},
        "auxiliary": {"cross_obs": T(O.cross_obs, {"reward_name": "success_cross"})},
    }
    env.observation_manager = ObservationManager(obs, env)
    terms = {"time_out": T(time_out, time_out=True)}                                                     # QD/racing_ctbr_env.py:245-260 (contact: out of scope)
    if cfg.term_out_of_bound:
        terms["outofbound"] = T(Te.out_of_bound, {"bounds": (cfg.oob_lo, cfg.oob_hi)})
    if cfg.term_bad_pose:
        terms["bad_pose"] = T(Te.bad_pose, {"asset_cfg": SceneEntityCfg("robot")})
    env.termination_manager = TerminationManager(terms, env)
    rew = {"progress_rewards": T(R.progress_reward_mine, name, cfg.w_progress),                        # QD/racing_ctbr_env.py:280-328 (collision: out of scope)
           "command_bodyrate_penalty": T(R.command_body_rate_penalty, {"action_name": "force_torque"}, cfg.w_bodyrate),
           "action_rate": T(R.command_rate_penalty, {"action_name": "force_torque"}, cfg.w_action_rate),
           "perception_reward": T(R.perception_reward, name, cfg.w_perception),
           "success_cross": T(R.success_cross, {"command_name": "next_gate_pose", "threshold": cfg.update_threshold}, cfg.w_success)}
    if cfg.w_bad_pose != 0.0:
        rew["bad_pose_penalty"] = T(R.penalize_bad_pose, {}, cfg.w_bad_pose)
    env.reward_manager = RewardManager(rew, env)

    def loss_term(func, weight, params=None):
        t = ns.LossTermCfg()
        t.func, t.weight, t.params, t.use_diff_states, t.use_action = func, weight, dict(params or {}), True, False
        return t
    env.loss_manager = ns.LossManager({"move_towards_goal": loss_term(Lo.racing_target_diff, cfg.w_loss_target, name),     # QD/racing_ctbr_env.py:338-353
                                       "falling": loss_term(Lo.racing_vel_diff, cfg.w_loss_vel),
                                       "falling_speed": loss_term(Lo.racing_falling_diff, cfg.w_loss_fall)}, env)
    cur = {"terrain_levels": T(Cu.racing_terrain_levels, {"cmd_name": "next_gate_pose", "move_on_threshold": cfg.level_up_gates,
                                                          "move_down_threshold": cfg.level_down_gates})}  # QD/racing_ctbr_env.py:262-278
    if cfg.noise_curriculum:
        cur["command_noise_level"] = T(Cu.racing_cmd_noise_levels, {"cmd_name": "next_gate_pose", "enhance_threshold": cfg.noise_up_gates,
                                                                    "decay_threshold": cfg.noise_down_gates, "enhance_percent": cfg.noise_up,
                                                                    "decay_percent": cfg.noise_down})
    env.curriculum_manager = CurriculumManager(cur, env)
    rp, rr, ry, rv = cfg.reset_pos, cfg.reset_roll_pitch, cfg.reset_yaw, cfg.reset_vel
    ev = {"reset": {"reset_base": T(Ev.reset_root_state_racing, {                                         # QD/racing_ctbr_env.py:177-197
        "pose_range": {"x": (-rp, rp), "y": (-rp, rp), "z": (-rp, rp), "roll": (-rr, rr), "pitch": (-rr, rr), "yaw": (-ry, ry)},
        "velocity_range": {k: (-rv, rv) for k in ("x", "y", "z", "roll", "pitch", "yaw")}})}}
    env.event_manager = EventManager(ev, env)
    # -- startup event (QD/racing_ctbr_env.py:211-219; the PhysX mass / inertia randomisation is out of scope)
    torch.manual_seed(seed + 1)
    Ev.randomize_rate_controller_gain_and_thrust_delay(env, None, "force_torque", pid_scale_factor=tuple(cfg.pid_scale),
                                                       thrust_delay_scale_factor=tuple(cfg.delay_scale))
    torch.manual_seed(seed + 1)
    srnd[:, L_.SRND_KP:L_.SRND_KP + 3] = torch.rand(n, 3)
    torch.rand(n, 3)                                                   # rate_gain_i (all-zero gains, QD/racing_ctbr_env.py:129)
    srnd[:, L_.SRND_KD:L_.SRND_KD + 3] = torch.rand(n, 3)
    srnd[:, L_.SRND_THRUST_DELAY:L_.SRND_THRUST_DELAY + 1] = torch.rand(n, 1)
    srnd[:, L_.SRND_TORQUE_DELAY:L_.SRND_TORQUE_DELAY + 3] = torch.rand(n, 3)
    return env, srnd


# ----------------------------------------------------------------------------------------------- reach-target tasks
class PlaneTerrain:
    def __init__(self, n):
        self.env_origins = torch.zeros(n, 3)            # reach_oracle R.3


def reset_root_state_uniform(env, env_ids, pose_range, velocity_range, asset_cfg=SceneEntityCfg("robot")):
    """[isaac] omni.isaac.lab.envs.mdp.reset_root_state_uniform."""
    asset = env.scene[asset_cfg.name]
    root_states = asset.data.default_root_state[env_ids].clone()
    ranges = torch.tensor([pose_range.get(k, (0.0, 0.0)) for k in ["x", "y", "z", "roll", "pitch", "yaw"]])
    rs = sample_uniform(ranges[:, 0], ranges[:, 1], (len(env_ids), 6), device="cpu")
    positions = root_states[:, 0:3] + env.scene.env_origins[env_ids] + rs[:, 0:3]
    orientations = M.quat_mul(root_states[:, 3:7], M.quat_from_euler_xyz(rs[:, 3], rs[:, 4], rs[:, 5]))
    ranges = torch.tensor([velocity_range.get(k, (0.0, 0.0)) for k in ["x", "y", "z", "roll", "pitch", "yaw"]])
    rs = sample_uniform(ranges[:, 0], ranges[:, 1], (len(env_ids), 6), device="cpu")
    velocities = root_states[:, 7:13] + rs
    asset.write_root_link_pose_to_sim(torch.cat([positions, orientations], dim=-1), env_ids=env_ids)
    asset.write_root_com_velocity_to_sim(velocities, env_ids=env_ids)


def _repaired_diff_actions(DiffActions):
    """DiffActions with the LV / PS branch of ``_get_scale_factor`` repaired: the reference builds ``action_offset`` (and ``action_scale``)
    as ``torch.tensor([...])[None].repeat(num_envs, 1)`` on an already 2-d literal for the offset (QD/mdp/diff_action.py:272-280), which
    raises for every num_envs; the evident intent -- one row per env -- is restated here.  Everything else is the reference's class."""
    class RepairedDiffActions(DiffActions):
        def _get_scale_factor(self, normal_range=(-1, 1), method="medium"):
            if self.command_type not in ("LVController", "PSController"):
                return super()._get_scale_factor(normal_range, method)
            self.motor_omega, self.thrustmap = self.controller_cfg.motor_omega, self.controller_cfg.thrustmap
            self.max_thrust_weight_ratio = self.cfg.max_thrust_weight_ratio
            bound = self.cfg.lin_vel_bound if self.command_type == "LVController" else self.cfg.pos_bound
            self.action_scale = torch.tensor([3.1415926, bound[1], bound[1], bound[1]], device=self.device)[None].repeat(self.num_envs, 1)
            self.action_offset = torch.tensor([[0.0, 0.0, 0.0, 0.0]], device=self.device).repeat(self.num_envs, 1)
    return RepairedDiffActions


def make_reference_reach_env(cfg, num_envs, seed, repair_lv_ps=False):
    """The reference env of DiffLab-Quadcopter-{LV,CTBR}-ReachTarget (QD/reach_target_lv_env.py, reach_target_ctbr_env.py) over
    the closure simulator, with the substitutions R.1-R.3 of oracle/reach_oracle.py.  [isaac] terms: base_lin_vel, base_ang_vel,
    last_action, action_rate_l2, body_lin_acc_l2, is_terminated, time_out, reset_root_state_uniform."""
    ns = load()
    n = num_envs
    NS = types.SimpleNamespace
    env = object.__new__(ns.Env)
    env.cfg = NS(sim=NS(dt=cfg.sim_dt, gravity=(0.0, 0.0, -cfg.gravity), render_interval=cfg.decimation), decimation=cfg.decimation,
                 episode_length_s=cfg.episode_length_s, is_differentiable_physics=True, rerender_on_reset=False)
    robot = Robot(n, cfg.mass, cfg.default_root_pos)
    env.scene = Scene(n, robot, PlaneTerrain(n))
    env.sim = Sim(env)
    env.extras = {}
    env._sim_step_counter = 0
    env.common_step_counter = 0
    env.episode_length_buf = torch.zeros(n, dtype=torch.long)
    env.recorder_manager = Recorder()
    lo, hi = cfg.cmd_lo, cfg.cmd_hi
    cmd_cfg = NS(asset_name="robot", resampling_time_range=(cfg.resampling_time, cfg.resampling_time), debug_vis=False, make_quat_unique=False,
                 ranges=NS(pos_x=(lo[0], hi[0]), pos_y=(lo[1], hi[1]), pos_z=(lo[2], hi[2]), roll=(0.0, 0.0), pitch=(0.0, 0.0), yaw=(0.0, 0.0)))
    env.command_manager = CommandManager({}, env)
    env.command_manager._terms["desired_pos_b"] = ns.commands.UniformWorldPoseCommand(cmd_cfg, env)
    if cfg.controller == "CTBRController":
        ctrl_cfg = RM.ctbr_cfg(cfg)
        ctrl_cfg.class_type = ns.CTBRController
    else:
        ctrl_cfg = RM.outer_loop_cfg(cfg)
        ref = RM.load()
        ctrl_cfg.class_type = ref.LVController if cfg.controller == "LVController" else ref.PSController
    b = cfg.lin_vel_bound
    act_cls = _repaired_diff_actions(ns.diff_action.DiffActions) if repair_lv_ps else ns.diff_action.DiffActions
    act_cfg = NS(class_type=act_cls, asset_name="robot", rotor_names="m.*_prop", command_type=cfg.controller,
                 controller_cfg=ctrl_cfg, gravity=9.81, random_drag=cfg.random_drag, action_lag=cfg.action_lag, sim2real_test=cfg.sim2real_test,
                 max_thrust_weight_ratio=cfg.max_thrust_weight_ratio, lin_vel_bound=(-b, b), pos_bound=(-cfg.pos_bound, cfg.pos_bound))
    torch.manual_seed(seed)
    env.action_manager = ns.DiffActionManager({"force_torque": act_cfg}, env)
    T = _TermCfg
    R, O, Te, Lo = ns.rewards, ns.observation, ns.termination, ns.losses
    data = robot.data
    last = T(O.modified_last_action, {"action_name": "force_torque"}) if cfg.last_action_modified else T(lambda e, action_name: e.action_manager.action, {"action_name": "force_torque"})
    env.observation_manager = ObservationManager({"policy": {
        "base_lin_vel": T(lambda e: data.root_lin_vel_b), "base_ang_vel": T(lambda e: data.root_ang_vel_b), "last_action": last,
        "base_orientation": T(O.base_orientation_q), "desired_pos_b": T(O.desired_position_b, {"command_name": "desired_pos_b"})}}, env)
    terms = {"time_out": T(time_out, time_out=True)}
    if cfg.term_out_of_bound:
        terms["outofbound"] = T(Te.out_of_bound, {"bounds": (cfg.oob_lo, cfg.oob_hi)})               # R.2
    env.termination_manager = TerminationManager(terms, env)
    w = cfg.w_reward
    body = SceneEntityCfg("robot", body_names="body")
    rew = {"move_towards": T(R.target_reward, {"command_name": "desired_pos_b"}, w[0]), "orientation_reward": T(R.orientation_reward, {}, w[1]),
           "move_in_dir": T(R.move_in_dir, {"threshold": cfg.move_in_dir_threshold}, w[2]),
           "action_rate": T(lambda e: torch.sum(torch.square(e.action_manager.action - e.action_manager.prev_action), dim=1), {}, w[3]),
           "reach_target": T(R.reach_target, {"threshold": cfg.reach_threshold}, w[4]), "smooth_ang_vel": T(R.ang_vel_reward, {}, w[5]),
           "smooth_lin_acc": T(lambda e: torch.sum(torch.norm(data.body_lin_acc_w[:, body.body_ids, :], dim=-1), dim=1), {}, w[6]),
           "smooth_ang_acc": T(R.body_ang_acc_l2, {"robot_cfg": body}, w[7]),
           "early_termination": T(lambda e: e.termination_manager.terminated.float(), {}, w[8]),
           "hover_state": T(R.hover_state, {"threshold": cfg.hover_threshold, "ratio": cfg.hover_ratio}, w[9])}
    env.reward_manager = RewardManager(rew, env)

    def loss_term(func, weight, params=None):
        t = ns.LossTermCfg()
        t.func, t.weight, t.params, t.use_diff_states, t.use_action = func, float(weight), dict(params or {}), True, False
        return t
    wl = cfg.w_loss
    env.loss_manager = ns.LossManager({"move_towards_goal": loss_term(Lo.target_diff, wl[0], {"command_name": "desired_pos_b"}),
                                       "orientation_tracking": loss_term(Lo.orientation_diff, wl[1]),
                                       "move_in_dir": loss_term(Lo.move_in_dir_diff, wl[2], {"command_name": "desired_pos_b", "threshold": cfg.loss_dir_threshold}),
                                       "smooth_vel": loss_term(Lo.smooth_vel_diff, wl[3], {"ratio": cfg.loss_smooth_ratio})}, env)
    env.curriculum_manager = CurriculumManager({}, env)
    pose = dict(zip(("x", "y", "z", "roll", "pitch", "yaw"), zip(cfg.reset_lo, cfg.reset_hi)))
    env.event_manager = EventManager({"reset": {"reset_base": T(reset_root_state_uniform, {"pose_range": pose, "velocity_range": {}})}}, env)
    return env


def replay_reach_reset_draws(rnd, ids, random_drag):
    """The global-generator calls of one reference ``_reset_idx(ids)`` of a reach-target env, into the oracle's REACH_RND_* columns."""
    n = len(ids)
    if n == 0:
        return
    rnd[ids, L_.REACH_RND_RESET_POSE:L_.REACH_RND_RESET_POSE + 6] = torch.rand(n, 6)     # [isaac] reset_root_state_uniform: pose
    torch.rand(n, 6)                                                                     # velocity ranges (all zero)
    if random_drag:                                                                      # droneDynamics.py:52-57
        rnd[ids, L_.REACH_RND_Z_DRAG] = torch.rand(n)
        rnd[ids, L_.REACH_RND_DRAG2:L_.REACH_RND_DRAG2 + 3] = torch.rand(n, 3)
        rnd[ids, L_.REACH_RND_DRAG1:L_.REACH_RND_DRAG1 + 3] = torch.rand(n, 3)
    rnd[ids, L_.REACH_RND_THR_ERR] = torch.randn(n)                                      # diff_action.py:233
    replay_reach_command_draws(rnd, ids, L_.REACH_RND_CMD)


def replay_reach_command_draws(rnd, ids, col):
    """[isaac] CommandTerm._resample (time_left) then UniformWorldPoseCommand._resample_command (QD/mdp/commands.py:113-126)."""
    n = len(ids)
    if n == 0:
        return
    torch.empty(n).uniform_(10.0, 10.0)
    for k in range(3):
        rnd[ids, col + k] = torch.empty(n).uniform_()
    for k in range(3):
        torch.empty(n).uniform_()                                                        # roll / pitch / yaw ranges (0, 0)


def replay_reset_draws(rnd, ids, add_noise):
    """The global-generator calls of one reference ``_reset_idx(ids)``, in order, into the oracle's columns."""
    n = len(ids)
    if n == 0:
        return
    rnd[ids, L_.RND_RESET_POSE:L_.RND_RESET_POSE + 6] = torch.rand(n, 6)        # events.py:153 (sample_uniform)
    rnd[ids, L_.RND_RESET_VEL:L_.RND_RESET_VEL + 6] = torch.rand(n, 6)          # events.py:171
    rnd[ids, L_.RND_Z_DRAG] = torch.rand(n)                                     # droneDynamics.py:53
    rnd[ids, L_.RND_DRAG2:L_.RND_DRAG2 + 3] = torch.rand(n, 3)                  # :54
    rnd[ids, L_.RND_DRAG1:L_.RND_DRAG1 + 3] = torch.rand(n, 3)                  # :56
    rnd[ids, L_.RND_THR_ERR] = torch.randn(n)                                   # diff_action.py:233
    torch.empty(n).uniform_(20.0, 20.0)                                         # [isaac] CommandTerm._resample time_left
    if not add_noise:                                                           # commands.py:286
        return
    for cols in (L_.RND_RESET_GATE, L_.RND_RESET_NEXT):                         # commands.py:287-306
        for c in cols:
            rnd[ids, c] = torch.empty(n).uniform_()


def replay_pass_draws(rnd, achieved, add_noise):
    k = int(achieved.sum())
    if k == 0 or not add_noise:
        return
    for cols in (L_.RND_PASS_GATE, L_.RND_PASS_NEXT):                           # commands.py:329-350
        for c in cols:
            rnd[achieved, c] = torch.empty(k).uniform_()


def replay_obs_draws(rnd):
    n = rnd.shape[0]
    rnd[:, L_.RND_OBS_VEL:L_.RND_OBS_VEL + 3] = torch.randn(n, 3)               # observation.py:52 (randn_like)
    rnd[:, L_.RND_OBS_EUL:L_.RND_OBS_EUL + 3] = torch.randn(n, 3)               # observation.py:27
