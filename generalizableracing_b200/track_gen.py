"""Racing-track generators -> gate tables.

Restates the *gate / spawn-origin* part of the reference's track families and the curriculum tiling that turns them into
the ``gate_pose[type, level, gate, 7]`` / ``next_gate_id`` / ``terrain_origins`` tables the racing command indexes
(paths relative to /root/reference, L = extensions/diff.lab/diff/lab, QD = extensions/diff.lab_tasks/diff/lab_tasks/
tasks/quadcopter_diff):

* ``SquareRacingTrackTerrain``  L/terrains/trimesh/racing_terrains.py:167-337
* ``FigureEightTrackTerrain``   L/terrains/trimesh/racing_terrains.py:340-415
* ``ZigzagRacingTerrain``       L/terrains/trimesh/racing_terrains.py:423-622
* ``EllipseRacingTerrain``      L/terrains/trimesh/racing_terrains.py:625-833
* pose conversion + origin recentring + cache files  L/terrains/terrain_generator.py:57-99
* curriculum tiling: Isaac Lab ``TerrainGenerator._generate_curriculum_terrains`` / ``_add_sub_terrain`` (third party,
  ``omni-isaac-lab``; restated from the published semantics) and ``TerrainImporter`` reshape L/terrains/terrain_importer.py:47-55
* the task's generator configs  QD/terrains/racing_terrains.py:114-211

The reference draws from the two GLOBAL streams ``random`` and ``np.random`` (seeded by the launcher's ``set_seed``); here they
are explicit objects (``random.Random`` / ``np.random.RandomState`` = the same MT19937 streams), consumed in exactly the
reference's order, with the same numpy expressions and dtypes, so that for a given stream state a family returns bit-identical
``gate_pose`` / ``origin`` / ``next_gate_id`` (tests/test_track_gen.py runs the unmodified reference functions, with the mesh
library stubbed out, against these).  Obstacle / wall / ground MESHES are out of scope (trimesh + Warp + USD), but with
``add_obs=True`` (the setting of all three families in RacingComplexTerrainCfg, QD/terrains/racing_terrains.py:163,184,204) the
reference consumes draws for them *after* a tile's gates and origin are fixed, which moves the stream position of every FOLLOWING
tile.  ``_ObstacleReplay`` therefore replays exactly those draws -- the scatter points with their rejection loops, wall sizes, the
shape lotteries of L/terrains/trimesh/utils.py:56-134 -- without building any geometry, so the whole 20 x 10 table is the reference's.
"""
from __future__ import annotations

import os
import random as _random
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .tracks import GateTable, gate_euler_to_quat_wxyz


class Streams:
    """The two global random streams of the reference as explicit objects."""

    def __init__(self, seed: int = 42):
        self.py = _random.Random(seed)
        self.np = np.random.RandomState(seed)


# ---------------------------------------------------------------------------------------------------------------
# family configs (only the fields the gate / origin part reads; L/terrains/trimesh/racing_terrains_cfg.py)
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class _FamilyCfg:
    proportion: float = 1.0
    size: Tuple[float, float] = (40.0, 40.0)
    num_gate: int = 8
    gate_size: Sequence[float] = (0.8, 1.2)
    gate_thickness: Sequence[float] = (0.03, 0.06)
    pos_noise_scale: Sequence[float] = (0.2, 1.0)
    rot_noise_scale: Sequence[float] = (0.0, 30.0)
    only_yaw: bool = True
    # obstacle branch: only the stream position of the following tiles depends on these (L/terrains/trimesh/racing_terrains_cfg.py)
    add_obs: bool = False
    add_ground_obs: bool = False
    num_ground_obs: Sequence[float] = (1, 4)
    num_wall_seg: Sequence[float] = (1, 5)
    wall_size: Sequence[float] = (0.4, 1.0)
    wall_thickness: Sequence[float] = (0.04, 0.08)
    num_orbit_seg: Sequence[float] = (1, 5)
    adj_dir_shift_prop: Sequence[float] = (0.2, 0.5)
    radius_dir_shift_prop: Sequence[float] = (0.2, 0.5)
    no_obs_range: float = 0.8


@dataclass
class SquareTrackCfg(_FamilyCfg):
    radius: Sequence[float] = (5.0, 8.0)
    add_border: bool = False


@dataclass
class ZigzagTrackCfg(_FamilyCfg):
    track_length: float = 35.0
    pos_noise_scale: Sequence[float] = (1.0, 4.0)
    pos_z_noise_scale: Sequence[float] = (0.1, 1.0)
    adj_dir_shift_prop: Sequence[float] = (1, 2)
    radius_dir_shift_prop: Sequence[float] = (10, 10)


@dataclass
class EllipseTrackCfg(_FamilyCfg):
    gate_distance: float = 5.0
    short_axis_prop: Sequence[float] = (1.414, 0.8)
    long_axis_prop: Sequence[float] = (3.1414, 4.8)
    adj_dir_shift_prop: Sequence[float] = (1, 2)
    radius_dir_shift_prop: Sequence[float] = (10, 10)


@dataclass
class FigureEightTrackCfg(_FamilyCfg):
    num_gate: int = 6
    gate_thickness: Sequence[float] = (0.08, 0.12)
    pos_noise_scale: Sequence[float] = (0.0, 0.0)
    rot_noise_scale: Sequence[float] = (0.0, 0.0)
    size: Tuple[float, float] = (18.0, 18.0)


def _lerp(rng_pair, d):
    return d * (rng_pair[1] - rng_pair[0]) + rng_pair[0]


def _gate_shape_draws(s: Streams, n: int, edge_hi: float):
    """Gate frame dimensions (mesh-only quantities): drawn to keep the stream aligned, values unused."""
    s.np.uniform(-0.05, 0.05, n)
    s.np.uniform(-0.05, 0.05, n)
    s.np.uniform(-1, 1, n)
    s.np.uniform(0.15, edge_hi, n)


def _pose6(gate_pts, gate_euler):
    pose = np.zeros((gate_pts.shape[0], 6), dtype=np.float32)
    pose[:, 0:3] = gate_pts
    pose[:, 3:6] = gate_euler
    return pose


class _ObstacleReplay:
    """Draw-only replay of the obstacle branch of one track segment.  Geometry is never built; what is reproduced is the ORDER and
    COUNT of the draws on the two streams (and the few values that steer them: the collinearity rejection loop, the zigzag
    family's keep-out test around the gates), so that the next tile starts from the reference's stream position."""

    def __init__(self, s: Streams, cfg: _FamilyCfg, difficulty: float, lateral_scale: float, probe_range: float):
        self.s, self.cfg = s, cfg
        self.adj = _lerp(cfg.adj_dir_shift_prop, difficulty)
        self.lat = _lerp(cfg.radius_dir_shift_prop, difficulty)
        self.lateral_scale = lateral_scale      # what the unit lateral offset is multiplied with (ring radius / gate distance / half gate size)
        self.probe_range = probe_range          # range of the random vector crossed with the segment direction (10 or 1)

    def segment(self, a, b):
        self.mid, self.vec = (a + b) / 2, b - a

    def scatter(self, along=None):
        """a point next to the segment: along-track shift, then a lateral shift along (segment x random vector)"""
        s = self.s.np
        along = self.adj if along is None else along
        off1 = self.vec / 2 * s.uniform(-along, along)
        while True:
            cross = np.cross(self.vec, s.uniform(-self.probe_range, self.probe_range, 3))
            if not np.allclose(cross, np.zeros(3)):
                break
        off2 = cross / np.linalg.norm(cross) * s.uniform(-self.lat, self.lat) * self.lateral_scale
        return self.mid + off1 + off2

    def raise_point(self, pt):
        pt[2] = self.s.py.uniform(0.5, 3.0)
        return pt

    def wall(self):
        s, c = self.s.np, self.cfg
        s.uniform(-180, 180, 3)
        s.uniform(c.wall_size[0], c.wall_size[1], 2)
        s.uniform(c.wall_thickness[0], c.wall_thickness[1])

    def orbit(self):
        """utils.py:56-83 (`make_orbit`): euler first (argument), then the shape lottery"""
        s = self.s
        s.np.uniform(-180, 180, 3)
        prob = s.py.random()
        if prob < 0.2:
            s.np.uniform(0.1, 0.5, 3)
        elif 0.4 <= prob < 0.6:
            s.np.uniform(0.1, 0.3)
        else:                                   # cylinder, capsule (the reference's cone branch is unreachable: 0.8 <= p < 0.8)
            s.np.uniform(0.1, 0.3)
            s.np.uniform(0.2, 0.6)

    def ground_pillar(self):
        """utils.py:85-104 (`make_ground_high_obs`)"""
        s = self.s
        s.np.uniform(-180, 180, 3)
        s.py.uniform(0.0, 2.0)
        if s.py.random() < 0.5:
            s.np.uniform(0.05, 1.0, 2)
        else:
            s.np.uniform(0.025, 0.5)

    def ground_clutter(self):
        """utils.py:106-134 (`make_ground_little_obj`)"""
        s = self.s
        s.np.uniform(-180, 180, 3)
        prob = s.py.random()
        if prob < 0.33:
            s.np.uniform(0.1, 1.5, 3)
            s.py.uniform(-0.2, 0.5)
        elif prob < 0.66:
            s.py.uniform(0.025, 0.5)
            s.py.uniform(0.1, 1.0)
            s.py.uniform(-0.2, 0.5)
        else:
            radius = s.py.uniform(0.05, 0.5)
            s.py.uniform(-radius, radius)
            s.py.uniform(-0.2, 0.5)


def _closed_loop_obstacles(rp: _ObstacleReplay, pts, start_seg: int, n_wall: int, n_orbit: int, n_ground: int, clutter_hi: int):
    """Ring / ellipse families (racing_terrains.py:266-321, 751-812): every segment but the spawn segment gets a fixed number of
    walls, orbits, ground pillars and 1..clutter_hi small ground objects."""
    n = pts.shape[0]
    for i in range(n):
        if i == start_seg:
            continue
        rp.segment(pts[i], pts[(i + 1) % n])
        for _ in range(n_wall):
            rp.raise_point(rp.scatter())
            rp.wall()
        for _ in range(n_orbit):
            rp.raise_point(rp.scatter())
            rp.orbit()
        if rp.cfg.add_ground_obs:
            for _ in range(n_ground):
                rp.scatter()
                rp.ground_pillar()
            for _ in range(rp.s.py.randint(1, clutter_hi)):
                rp.scatter()
                rp.ground_clutter()


def _open_line_obstacles(rp: _ObstacleReplay, pts, n_wall: int, n_orbit: int, n_ground: int):
    """Zigzag family (racing_terrains.py:511-603): between consecutive gates, candidates closer than `no_obs_range` to either gate
    are redrawn (walls, orbits: 3-D distance after the height draw; ground objects: planar distance); clutter candidates inside
    the range are dropped, not redrawn."""
    keep_out = rp.cfg.no_obs_range
    for i in range(pts.shape[0] - 1):
        a, b = pts[i], pts[i + 1]
        rp.segment(a, b)
        for count, finish in ((n_wall, rp.wall), (n_orbit, rp.orbit)):
            placed = 0
            while placed < count:
                pt = rp.raise_point(rp.scatter())
                if np.linalg.norm(pt - a) < keep_out or np.linalg.norm(pt - b) < keep_out:
                    continue
                finish()
                placed += 1
        if rp.cfg.add_ground_obs:
            placed = 0
            while placed < n_ground:
                pt = rp.scatter()
                if np.linalg.norm(pt[:2] - a[:2]) < keep_out or np.linalg.norm(pt[:2] - b[:2]) < keep_out:
                    continue
                rp.ground_pillar()
                placed += 1
            for _ in range(rp.s.py.randint(1, 4)):
                pt = rp.scatter(along=0.5)
                if np.linalg.norm(pt[:2] - a[:2]) < keep_out or np.linalg.norm(pt[:2] - b[:2]) < keep_out:
                    continue
                rp.ground_clutter()


# ---------------------------------------------------------------------------------------------------------------
# families: (difficulty, cfg, streams) -> (gate_pose[G,6] float32 = xyz | euler deg, origin[3], next_gate_id)
# ---------------------------------------------------------------------------------------------------------------
def square_track(difficulty: float, cfg: SquareTrackCfg, s: Streams):
    """Ring of gates of random radius (racing_terrains.py:179-262, 323-337)."""
    radius = s.py.uniform(cfg.radius[0], cfg.radius[1])
    n = cfg.num_gate
    pos_scale = _lerp(cfg.pos_noise_scale, difficulty)
    rot_scale = _lerp(cfg.rot_noise_scale, difficulty)
    theta = np.linspace(0, 2 * np.pi, n, endpoint=False)
    pts = np.zeros((n, 3), dtype=np.float32)
    pts[:, 0] = np.cos(theta) * radius
    pts[:, 1] = np.sin(theta) * radius
    pts[:, 2] = 1.0
    pts[:, 0] += cfg.size[0] / 2
    pts[:, 1] += cfg.size[1] / 2
    pts[:, 2] += 0
    eul = np.zeros((n, 3), dtype=np.float32)
    eul[:, 0] = 90.0
    eul[:, 1] = theta / np.pi * 180.0
    pos_noise = s.np.uniform(-1, 1, (n, 3)) * pos_scale
    rot_noise = s.np.uniform(-1, 1, (n, 3)) * rot_scale
    if cfg.only_yaw:
        rot_noise[:, 0] = 0.0
        rot_noise[:, 2] = 0.0
    pts += pos_noise
    pts[:, 2] = pts[:, 2].clip(0.8, 2.0)
    eul += rot_noise
    _gate_shape_draws(s, n, 0.25)
    reverse = 1
    if s.py.random() < 0.5:
        pts, eul, reverse = pts[::-1, :], eul[::-1, :], -1
    start_seg = s.py.randint(0, n - 1)
    nxt = (start_seg + 1) % n
    heading = eul[nxt][1] / 180 * np.pi + np.pi / 2
    origin = pts[nxt] - reverse * s.py.uniform(2, 4) * np.array([np.cos(heading), np.sin(heading), 0])
    origin[2] = s.py.uniform(0.7, 1.5)
    if cfg.add_obs:
        _closed_loop_obstacles(_ObstacleReplay(s, cfg, difficulty, lateral_scale=radius, probe_range=10), pts, start_seg,
                               int(_lerp(cfg.num_wall_seg, difficulty) * radius / cfg.radius[1]), int(_lerp(cfg.num_orbit_seg, difficulty) * radius / cfg.radius[1]),
                               int(_lerp(cfg.num_ground_obs, difficulty) * radius / cfg.radius[1]), clutter_hi=4)
    return _pose6(pts, eul), origin, nxt


def figure_eight_tile(difficulty: float, cfg: FigureEightTrackCfg, s: Streams):
    """Six gates on a figure eight (racing_terrains.py:340-415)."""
    pos_scale = _lerp(cfg.pos_noise_scale, difficulty)
    rot_scale = _lerp(cfg.rot_noise_scale, difficulty)
    pts = np.array([[3.0, 3.0, 1.0], [5.0, 0.0, 1.0], [3.0, -3.0, 1.0], [-3.0, 3.0, 1.0], [-5.0, 0.0, 1.0], [-3.0, -3.0, 1.0]], dtype=np.float32)
    eul = np.array([[90.0, 90.0, 0.0], [90.0, 0.0, 0.0], [90.0, 90.0, 0.0], [90.0, 90.0, 0.0], [90.0, 0.0, 0.0], [90.0, 90.0, 0.0]], dtype=np.float32)
    pos_noise = s.np.uniform(-1, 1, (6, 3)) * pos_scale
    rot_noise = s.np.uniform(-1, 1, (6, 3)) * rot_scale
    if cfg.only_yaw:
        rot_noise[:, 0] = 0.0
        rot_noise[:, 2] = 0.0
    pts += pos_noise
    pts[:, 2] = pts[:, 2].clip(1.0, 2.0)
    eul += rot_noise
    _gate_shape_draws(s, 6, 0.22)
    if s.py.random() < 0.5:
        pts, eul = pts[::-1, :], eul[::-1, :]
    origin = s.np.uniform(-1, 1, 3) * 0.5 + np.array([0.0, 0.0, 1.5])
    origin[2] = s.py.uniform(0.7, 1.5)
    return _pose6(pts, eul), origin, 0


def zigzag_track(difficulty: float, cfg: ZigzagTrackCfg, s: Streams):
    """Gates along a randomly oriented line, lateral / vertical noise growing along it (racing_terrains.py:433-509, 613-622)."""
    length, n = cfg.track_length, cfg.num_gate
    pos_scale = _lerp(cfg.pos_noise_scale, difficulty)
    z_scale = _lerp(cfg.pos_z_noise_scale, difficulty)
    rot_scale = _lerp(cfg.rot_noise_scale, difficulty)
    eul = np.zeros((n, 3), dtype=np.float32)
    theta = s.np.uniform(0, 2 * np.pi)
    direction = np.array([np.cos(theta), np.sin(theta), 0])
    start_point = -0.5 * length * direction
    end_point = 0.5 * length * direction
    t_values = np.linspace(0, 1, n)
    points = start_point + np.outer(t_values, end_point - start_point)
    for i in range(1, n - 1):
        f = t_values[i]
        lateral = np.array([-direction[1], direction[0], 0])
        lateral = lateral / np.linalg.norm(lateral)
        points[i] += 2.0 * (s.np.rand() - 0.5) * pos_scale * f * lateral
        s.np.rand()                                   # along-track noise: drawn by the reference, never applied
        points[i] += 2.0 * (s.np.rand() - 0.5) * z_scale * f * np.array([0, 0, 1])
    eul[:, 0] = 90.0
    eul[:, 1] = theta / np.pi * 180.0 + 90
    pts = points
    pts[:, 0] += cfg.size[0] / 2
    pts[:, 1] += cfg.size[1] / 2
    pts[:, 2] += 1.0
    pts[:, 2] = pts[:, 2].clip(0.8, 2.0)
    rot_noise = s.np.uniform(-1, 1, (n, 3)) * rot_scale
    if cfg.only_yaw:
        rot_noise[:, 0] = 0.0
        rot_noise[:, 2] = 0.0
    eul += rot_noise
    _gate_shape_draws(s, n, 0.25)
    first_dir = pts[1, :] - pts[0, :]
    first_dir = first_dir / np.linalg.norm(first_dir)
    origin = pts[0].copy() - first_dir * s.py.uniform(2, 3)
    origin[2] = s.py.uniform(0.7, 1.5)
    if cfg.add_obs:
        _open_line_obstacles(_ObstacleReplay(s, cfg, difficulty, lateral_scale=cfg.gate_size[1] / 2, probe_range=1), pts,
                             int(_lerp(cfg.num_wall_seg, difficulty)), int(_lerp(cfg.num_orbit_seg, difficulty)), int(_lerp(cfg.num_ground_obs, difficulty)))
    return _pose6(pts, eul), origin, 0


def ellipse_track(difficulty: float, cfg: EllipseTrackCfg, s: Streams):
    """Eight gates on a stadium: two on the long axis, three on each long side (racing_terrains.py:640-748, 826-833)."""
    n = cfg.num_gate
    if n != 8:
        raise ValueError("the ellipse family places exactly 8 gates")
    a_ellipse = (cfg.long_axis_prop[0] + difficulty * (cfg.long_axis_prop[1] - cfg.long_axis_prop[0])) * cfg.gate_distance
    b_ellipse = (cfg.short_axis_prop[0] + difficulty * (cfg.short_axis_prop[1] - cfg.short_axis_prop[0])) * cfg.gate_distance
    pos_scale = _lerp(cfg.pos_noise_scale, difficulty)
    rot_scale = _lerp(cfg.rot_noise_scale, difficulty)
    eul = np.zeros((n, 3), dtype=np.float32)
    eul[:, 0] = 90.0
    theta = s.np.uniform(0, 2 * np.pi)
    theta_deg = theta / np.pi * 180.0
    long_dir = np.array([np.cos(theta), np.sin(theta), 0])
    short_dir = np.array([-np.sin(theta), np.cos(theta), 0])
    pts = np.zeros((n, 3), dtype=np.float32)
    pts[0] = -0.5 * a_ellipse * long_dir
    pts[4] = 0.5 * a_ellipse * long_dir
    eul[0, 1] = theta_deg
    eul[4, 1] = 180 + theta_deg
    pts[2] = 0.5 * b_ellipse * short_dir
    pts[6] = -0.5 * b_ellipse * short_dir
    eul[2, 1] = theta_deg + 90
    eul[6, 1] = theta_deg + 270
    pts[1] = pts[2] - cfg.gate_distance * long_dir
    pts[3] = pts[2] + cfg.gate_distance * long_dir
    pts[5] = pts[6] + cfg.gate_distance * long_dir
    pts[7] = pts[6] - cfg.gate_distance * long_dir
    eul[1, 1] = theta_deg + 90
    eul[3, 1] = theta_deg + 90
    eul[5, 1] = theta_deg + 270
    eul[7, 1] = theta_deg + 270
    pts[:, 0] += cfg.size[0] / 2
    pts[:, 1] += cfg.size[1] / 2
    pts[:, 2] += 1.0
    pos_noise = s.np.uniform(-1, 1, (n, 3)) * pos_scale
    rot_noise = s.np.uniform(-1, 1, (n, 3)) * rot_scale
    if cfg.only_yaw:
        rot_noise[:, 0] = 0.0
        rot_noise[:, 2] = 0.0
    pts += pos_noise
    pts[:, 2] = pts[:, 2].clip(0.8, 2.0)
    eul += rot_noise
    _gate_shape_draws(s, n, 0.22)
    if s.py.random() < 0.5:
        pts, eul = pts[::-1, :], eul[::-1, :]
    start_seg = s.py.randint(0, n - 1)
    nxt = (start_seg + 1) % n
    seg = pts[nxt] - pts[start_seg]
    seg = seg / np.linalg.norm(seg)
    origin = pts[start_seg % n] + seg * s.py.uniform(2, 3)
    origin[2] = s.py.uniform(0.7, 1.5)
    if cfg.add_obs:
        _closed_loop_obstacles(_ObstacleReplay(s, cfg, difficulty, lateral_scale=cfg.gate_distance, probe_range=10), pts, start_seg,
                               int(_lerp(cfg.num_wall_seg, difficulty)), int(_lerp(cfg.num_orbit_seg, difficulty)), int(_lerp(cfg.num_ground_obs, difficulty)),
                               clutter_hi=2)
    return _pose6(pts, eul), origin, nxt


FAMILIES = {SquareTrackCfg: square_track, ZigzagTrackCfg: zigzag_track, EllipseTrackCfg: ellipse_track, FigureEightTrackCfg: figure_eight_tile}


# ---------------------------------------------------------------------------------------------------------------
# tile -> table entry (L/terrains/terrain_generator.py:57-77)
# ---------------------------------------------------------------------------------------------------------------
def tile_entry(gate_pose6: np.ndarray, origin: np.ndarray, size: Sequence[float]):
    """(gate_pose7 relative to the spawn origin with wxyz quaternions, spawn origin in tile-centred coordinates)."""
    pose7 = np.zeros((gate_pose6.shape[0], 7))
    pose7[:, :3] = gate_pose6[:, :3] - origin
    pose7[:, 3:] = gate_euler_to_quat_wxyz(gate_pose6[:, 3:6])
    centred = np.array(origin, copy=True)
    centred += np.array([-size[0] * 0.5, -size[1] * 0.5, 0.0])
    return pose7, centred


# ---------------------------------------------------------------------------------------------------------------
# generator config + curriculum tiling (QD/terrains/racing_terrains.py; Isaac Lab TerrainGenerator, restated)
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class TrackGeneratorCfg:
    sub_terrains: Dict[str, _FamilyCfg]
    size: Tuple[float, float] = (40.0, 40.0)
    num_rows: int = 10            # curriculum levels
    num_cols: int = 20            # terrain types
    seed: int = 42                # TerrainGeneratorCfg.seed: the generator's own np_rng (difficulty jitter)
    difficulty_range: Tuple[float, float] = (0.0, 1.0)
    curriculum: bool = True


def racing_complex_cfg(add_obs: bool = True) -> TrackGeneratorCfg:
    """RacingComplexTerrainCfg (QD/terrains/racing_terrains.py:137-211).  ``add_obs=True`` is the reference's setting: no obstacle
    geometry exists here, but its draws are replayed so the gate table is the one the reference builds from seed 42;
    ``add_obs=False`` gives the obstacle-free variant of the same families (a different table from the second tile on)."""
    size = (40.0, 40.0)
    obs = dict(add_obs=add_obs, add_ground_obs=True, wall_size=(0.4, 1.0), wall_thickness=(0.04, 0.08), adj_dir_shift_prop=(0.6, 0.6))
    return TrackGeneratorCfg(size=size, num_rows=10, num_cols=20, seed=42, sub_terrains={
        "zigzag": ZigzagTrackCfg(proportion=0.3, size=size, track_length=35.0, num_gate=8, pos_noise_scale=(1.0, 4.0), pos_z_noise_scale=(0.1, 1.0),
                                 num_wall_seg=(2, 6), num_orbit_seg=(2, 6), num_ground_obs=(1, 4), radius_dir_shift_prop=(6, 6), no_obs_range=1.5, **obs),
        "circular": SquareTrackCfg(proportion=0.3, size=size, radius=(5.0, 8.0), num_gate=8,
                                   num_wall_seg=(1, 4), num_orbit_seg=(1, 4), num_ground_obs=(1, 4), radius_dir_shift_prop=(0.5, 0.5), **obs),
        "ellipse": EllipseTrackCfg(proportion=0.4, size=size, gate_distance=5.0, num_gate=8,
                                   num_wall_seg=(1, 4), num_orbit_seg=(1, 4), num_ground_obs=(1, 2), radius_dir_shift_prop=(0.5, 0.5), **obs)})


def racing_test_cfg() -> TrackGeneratorCfg:
    """RacingTestTerrainCfg (QD/terrains/racing_terrains.py:114-134): one figure-eight tile, zero noise."""
    size = (18.0, 18.0)
    return TrackGeneratorCfg(size=size, num_rows=1, num_cols=1, seed=42, curriculum=False,
                             sub_terrains={"circular": FigureEightTrackCfg(proportion=1.0, size=size)})


def generate_track_table(cfg: TrackGeneratorCfg, global_seed: int = 42, name: str = "generated") -> GateTable:
    """Isaac Lab's curriculum layout: columns (= terrain types) are assigned to families by cumulative proportion, rows (= levels)
    get difficulty ``(row + U[0,1)) / num_rows`` from the generator's own ``default_rng(cfg.seed)``, tiles are generated
    column-major and laid out on a grid centred on the world origin.  ``global_seed`` seeds the two global streams."""
    s = Streams(global_seed)
    np_rng = np.random.default_rng(cfg.seed)
    fams = list(cfg.sub_terrains.values())
    prop = np.array([f.proportion for f in fams], dtype=np.float64)
    prop /= np.sum(prop)
    cum = np.cumsum(prop)
    rows, cols = cfg.num_rows, cfg.num_cols
    if cfg.curriculum:
        order = [(r, c, fams[int(np.min(np.where(c / cols + 0.001 < cum)[0]))], (r + np_rng.uniform()) / rows) for c in range(cols) for r in range(rows)]
    else:      # _generate_random_terrains: row-major, family and difficulty drawn per tile
        order = []
        for index in range(rows * cols):
            r, c = np.unravel_index(index, (rows, cols))
            fam = fams[int(np_rng.choice(len(fams), p=prop))]
            order.append((int(r), int(c), fam, np_rng.uniform(*cfg.difficulty_range)))
    gates = fams[0].num_gate
    pose = np.zeros((cols, rows, gates, 7), dtype=np.float32)
    nxt = np.zeros((cols, rows), dtype=np.int32)
    origins = np.zeros((rows, cols, 3), dtype=np.float32)
    lo, hi = cfg.difficulty_range
    for r, c, fam, d in order:
        if fam.num_gate != gates:
            raise ValueError("every family of one table must place the same number of gates (QD/terrains/racing_terrains.py:136)")
        difficulty = lo + (hi - lo) * d if cfg.curriculum else d
        pose6, origin, next_id = FAMILIES[type(fam)](float(difficulty), fam, s)
        pose7, centred = tile_entry(pose6, origin, cfg.size)
        pose[c, r] = pose7
        nxt[c, r] = next_id
        # _add_sub_terrain: tile (r, c) sits at ((r + .5) sx, (c + .5) sy); the whole grid is then centred on the origin
        origins[r, c] = centred + np.array([(r + 0.5) * cfg.size[0] - cfg.size[0] * rows * 0.5, (c + 0.5) * cfg.size[1] - cfg.size[1] * cols * 0.5, 0.0])
    return GateTable(pose, nxt, origins, name=name)


# ---------------------------------------------------------------------------------------------------------------
# on-disk tile cache (L/terrains/terrain_generator.py:36-39, 88-99): <dir>/gate_info.yaml + <dir>/origin.csv  (mesh.obj: out of scope)
# ---------------------------------------------------------------------------------------------------------------
def save_tile_cache(tile_dir: str, pose7: np.ndarray, next_gate_id: int, origin_centred: np.ndarray) -> None:
    import yaml
    os.makedirs(tile_dir, exist_ok=True)
    np.savetxt(os.path.join(tile_dir, "origin.csv"), origin_centred, delimiter=",", header="x,y,z")
    with open(os.path.join(tile_dir, "gate_info.yaml"), "w") as f:
        yaml.dump({"gate_pose": np.asarray(pose7).tolist(), "next_gate_id": int(next_gate_id)}, f, default_flow_style=None, sort_keys=False)


def load_tile_cache(tile_dir: str):
    import yaml
    origin = np.loadtxt(os.path.join(tile_dir, "origin.csv"), delimiter=",")
    with open(os.path.join(tile_dir, "gate_info.yaml")) as f:
        info = yaml.full_load(f)
    return np.array(info["gate_pose"]), int(info["next_gate_id"]), origin


def table_from_cache(tile_dirs: List[List[str]], size: Sequence[float]) -> GateTable:
    """tile_dirs[col][row] -> GateTable, as TerrainImporter assembles it from cached tiles (terrain_importer.py:47-50)."""
    cols, rows = len(tile_dirs), len(tile_dirs[0])
    first = load_tile_cache(tile_dirs[0][0])[0]
    pose = np.zeros((cols, rows, first.shape[0], 7), dtype=np.float32)
    nxt = np.zeros((cols, rows), dtype=np.int32)
    origins = np.zeros((rows, cols, 3), dtype=np.float32)
    for c in range(cols):
        for r in range(rows):
            p7, nid, o = load_tile_cache(tile_dirs[c][r])
            pose[c, r], nxt[c, r] = p7, nid
            origins[r, c] = o + np.array([(r + 0.5) * size[0] - size[0] * rows * 0.5, (c + 0.5) * size[1] - size[1] * cols * 0.5, 0.0])
    return GateTable(pose, nxt, origins, name="cache")
