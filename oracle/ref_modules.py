"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

Loads the UNMODIFIED reference modules that import without Isaac Sim, by file
path from /root/reference (present only in the build container, never on the GPU
box).  Used by tests/test_oracle_vs_reference.py and tests/golden/make_golden.py
to pin the restatement in oracle/racing_oracle.py and oracle/rollout_oracle.py.

Recipe: stub ``omni.isaac.lab.utils.math`` with the three functions
the files use (oracle/isaac_math.py) and ``rsl_rl.utils.split_and_pad_trajectories``;
load ``controller_diff.py`` + ``thrust_controller_diff.py`` under a synthetic parent
package so the relative import resolves without running controllers/__init__.py.
Nothing is copied: the sources are executed where they lie.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("GRACING_REFERENCE_ROOT", "/root/reference")
_QD = "extensions/diff.lab_tasks/diff/lab_tasks/tasks/quadcopter_diff"
_L = "extensions/diff.lab/diff/lab"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, _QD, "mdp/dynamics/droneDynamics.py"))


def _install_stubs():
    from . import isaac_math

    for name in ("omni", "omni.isaac", "omni.isaac.lab", "omni.isaac.lab.utils"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    if "omni.isaac.lab.utils.math" not in sys.modules:
        m = types.ModuleType("omni.isaac.lab.utils.math")
        for fn in ("quat_mul", "quat_rotate", "quat_rotate_inverse", "matrix_from_quat"):
            setattr(m, fn, getattr(isaac_math, fn))
        sys.modules["omni.isaac.lab.utils.math"] = m
        sys.modules["omni.isaac.lab.utils"].math = m
    if "rsl_rl" not in sys.modules:
        r = types.ModuleType("rsl_rl")
        r.__path__ = []
        u = types.ModuleType("rsl_rl.utils")
        u.split_and_pad_trajectories = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("stub"))
        r.utils = u
        sys.modules["rsl_rl"] = r
        sys.modules["rsl_rl.utils"] = u


def _load(modname: str, path: str):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def load():
    """Returns a namespace with DroneDynamics, CTBRController, RolloutStorage of the reference."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise FileNotFoundError(f"reference tree not found under {REF_ROOT}")
    _install_stubs()
    dyn = _load("_gr_ref_droneDynamics", os.path.join(REF_ROOT, _QD, "mdp/dynamics/droneDynamics.py"))
    pkg = types.ModuleType("_gr_ref_controllers")
    pkg.__path__ = [os.path.join(REF_ROOT, _L, "controllers")]
    sys.modules["_gr_ref_controllers"] = pkg
    _load("_gr_ref_controllers.thrust_controller_diff", os.path.join(REF_ROOT, _L, "controllers/thrust_controller_diff.py"))
    ctrl = _load("_gr_ref_controllers.controller_diff", os.path.join(REF_ROOT, _L, "controllers/controller_diff.py"))
    sto = _load("_gr_ref_rollout_storage", os.path.join(REF_ROOT, "standalone/rsl_rl/ext/storage/rollout_storage.py"))
    ns = types.SimpleNamespace(DroneDynamics=dyn.DroneDynamics, CTBRController=ctrl.CTBRController, LVController=ctrl.LVController, PSController=ctrl.PSController,
                               RolloutStorage=sto.RolloutStorage)
    _cache["ns"] = ns
    return ns


def ctbr_cfg(cfg):
    """Plain object carrying the CTBRControllerCfg fields (L/controllers/controller_diff_cfg.py:21-54) with the
    racing overrides of QD/racing_ctbr_env.py:127-134."""
    return types.SimpleNamespace(
        arm_length=0.09, kappa=0.016, motor_tau=0.0001, motor_omega=(150, 3000),
        thrustmap=[1.3298253500372892e-06, 0.0038360810526746033, -1.7689986848125325],
        rate_gain_p=list(cfg.rate_gain_p), rate_gain_i=[0.0, 0.0, 0.0], rate_gain_d=list(cfg.rate_gain_d),
        body_rate_bound=[-cfg.body_rate_bound, cfg.body_rate_bound],
        thrust_ctrl_delay=cfg.thrust_ctrl_delay, torque_ctrl_delay=tuple(cfg.torque_ctrl_delay),
        use_motor_model=False,
    )


def outer_loop_cfg(cfg):
    """Plain object carrying the LVControllerCfg / PSControllerCfg fields (L/controllers/controller_diff_cfg.py:56-79)."""
    return types.SimpleNamespace(
        arm_length=0.09, kappa=0.016, motor_tau=0.0001, motor_omega=(150, 3000), g=cfg.gravity,
        thrustmap=[1.3298253500372892e-06, 0.0038360810526746033, -1.7689986848125325],
        max_feedback_accel=cfg.max_feedback_accel, body_rate_bound=[-cfg.body_rate_bound, cfg.body_rate_bound],
        speed_gain=list(cfg.speed_gain), pose_gain=list(cfg.pose_gain), rate_gain=list(cfg.rate_gain), pos_gain=list(cfg.pos_gain),
        thrust_ctrl_delay=cfg.thrust_ctrl_delay, torque_ctrl_delay=tuple(cfg.torque_ctrl_delay), use_motor_model=False,
    )


def load_track_families():
    """The reference's track-family functions (L/terrains/trimesh/racing_terrains.py) AND its mesh helpers
    (L/terrains/trimesh/utils.py: make_gate / make_wall / make_orbit / make_ground_high_obs / make_ground_little_obj), executed
    unmodified where they lie.  Only the third-party ``trimesh`` package (absent here) is stubbed: its constructors return an inert
    object, so no geometry is built, but every ``random`` / ``np.random`` draw of the reference -- including the obstacle branch
    (``add_obs=True``) and the shape lotteries inside the mesh helpers -- happens exactly as in the reference, which is what fixes
    the gate poses / origins / next_gate_id of every FOLLOWING tile."""
    if "tracks" in _cache:
        return _cache["tracks"]
    if not available():
        raise FileNotFoundError(f"reference tree not found under {REF_ROOT}")
    _install_stubs()
    for name in ("omni.isaac.lab.terrains", "omni.isaac.lab.terrains.trimesh", "omni.isaac.lab.terrains.trimesh.utils"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["omni.isaac.lab.terrains.trimesh.utils"].make_border = lambda *a, **k: []

    class _InertMesh:
        def difference(self, other):
            return self

        def apply_transform(self, m):
            return self

        def apply_translation(self, t):
            return self

    mk = lambda *a, **k: _InertMesh()
    had_trimesh = sys.modules.get("trimesh")
    tm = types.ModuleType("trimesh")
    tm.Trimesh = _InertMesh
    tm.creation = types.SimpleNamespace(box=mk, cylinder=mk, icosphere=mk, cone=mk, capsule=mk)
    tm.transformations = types.SimpleNamespace(translation_matrix=lambda *a, **k: None, euler_matrix=lambda *a, **k: None)
    sys.modules["trimesh"] = tm
    try:
        pkg = types.ModuleType("_gr_ref_trimesh_terrains")
        pkg.__path__ = [os.path.join(REF_ROOT, _L, "terrains/trimesh")]
        sys.modules["_gr_ref_trimesh_terrains"] = pkg
        _load("_gr_ref_trimesh_terrains.utils", os.path.join(REF_ROOT, _L, "terrains/trimesh/utils.py"))
        mod = _load("_gr_ref_trimesh_terrains.racing_terrains", os.path.join(REF_ROOT, _L, "terrains/trimesh/racing_terrains.py"))
    finally:
        if had_trimesh is None:
            sys.modules.pop("trimesh", None)
        else:
            sys.modules["trimesh"] = had_trimesh
    ns = types.SimpleNamespace(square=mod.SquareRacingTrackTerrain, zigzag=mod.ZigzagRacingTerrain, ellipse=mod.EllipseRacingTerrain,
                               figure_eight=mod.FigureEightTrackTerrain)
    _cache["tracks"] = ns
    return ns
