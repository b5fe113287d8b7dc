"""Gate-pose tables ("tracks") consumed by the racing command.

Data contract restated from the reference (relative to /root/reference,
L = extensions/diff.lab/diff/lab):

* ``gate_pose[type, level, gate, 7] = (xyz - tile_origin, quat wxyz)``
  (L/terrains/terrain_generator.py:64-77, reshaped ``(num_cols, num_rows, G, 7)``
  by L/terrains/terrain_importer.py:47-50 and indexed
  ``[terrain_types, terrain_levels, gate_id]`` at
  extensions/diff.lab_tasks/.../quadcopter_diff/mdp/commands.py:272).
* ``next_gate_id[type, level]`` = id of the first gate to fly
  (L/terrains/trimesh/racing_terrains.py:410).
* ``terrain_origins[level, type, 3]`` = world position of the spawn origin of
  each tile; ``env_origins[n] = terrain_origins[level[n], type[n]]``.

Only the *centre-line* (gate positions/orientations) of the reference's track
families is produced here; obstacle / wall / ground meshes need trimesh + USD
and are out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def _quat_mul_xyzw(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def _axis_quat_xyzw(axis: int, deg: np.ndarray) -> np.ndarray:
    half = np.deg2rad(np.asarray(deg, dtype=np.float64)) * 0.5
    q = np.zeros(half.shape + (4,), dtype=np.float64)
    q[..., axis] = np.sin(half)
    q[..., 3] = np.cos(half)
    return q


def gate_euler_to_quat_wxyz(gate_euler_deg: np.ndarray) -> np.ndarray:
    """Gate euler triple (degrees) -> quaternion (w, x, y, z).

    Restates L/terrains/terrain_generator.py:69-73:
    ``R.from_euler('YXZ', [e0, -e1, e2]) * R.from_euler('XYZ', [-90, -90, 0])``
    (both intrinsic sequences), scipy xyzw re-ordered to wxyz.
    """
    e = np.asarray(gate_euler_deg, dtype=np.float64).reshape(-1, 3)
    # intrinsic 'YXZ' = Ry(e0) * Rx(-e1) * Rz(e2)
    ori = _quat_mul_xyzw(_quat_mul_xyzw(_axis_quat_xyzw(1, e[:, 0]), _axis_quat_xyzw(0, -e[:, 1])),
                         _axis_quat_xyzw(2, e[:, 2]))
    # intrinsic 'XYZ' = Rx(-90) * Ry(-90) * Rz(0)
    off = _quat_mul_xyzw(_axis_quat_xyzw(0, np.array([-90.0])), _axis_quat_xyzw(1, np.array([-90.0])))
    q = _quat_mul_xyzw(ori, np.broadcast_to(off, ori.shape))
    # scipy returns the canonical sign (w >= 0)
    q = np.where(q[:, 3:4] < 0, -q, q)
    return np.concatenate([q[:, 3:4], q[:, :3]], axis=1)


@dataclass
class GateTable:
    """A set of racing tracks laid out on a (types x levels) tile grid."""

    gate_pose: np.ndarray        # [types, levels, G, 7] float32
    next_gate_id: np.ndarray     # [types, levels] int32
    terrain_origins: np.ndarray  # [levels, types, 3] float32
    name: str = "custom"

    def __post_init__(self):
        self.gate_pose = np.ascontiguousarray(self.gate_pose, dtype=np.float32)
        self.next_gate_id = np.ascontiguousarray(self.next_gate_id, dtype=np.int32)
        self.terrain_origins = np.ascontiguousarray(self.terrain_origins, dtype=np.float32)
        t, l, g, c = self.gate_pose.shape
        if c != 7:
            raise ValueError("gate_pose must be [types, levels, G, 7]")
        if self.next_gate_id.shape != (t, l):
            raise ValueError("next_gate_id must be [types, levels]")
        if self.terrain_origins.shape != (l, t, 3):
            raise ValueError("terrain_origins must be [levels, types, 3]")
        if not (1 <= g <= 32):
            raise ValueError("1..32 gates per track are supported")

    @property
    def num_types(self) -> int:
        return self.gate_pose.shape[0]

    @property
    def num_levels(self) -> int:
        return self.gate_pose.shape[1]

    @property
    def num_gates(self) -> int:
        return self.gate_pose.shape[2]


def _tile_world_origin(local_origin, row, col, num_rows, num_cols, size):
    """World spawn origin of tile (row=level, col=type) as Isaac Lab's TerrainGenerator lays tiles out:
    local origin, centred on the tile, tile placed at ((row+.5)sx, (col+.5)sy), grid centred on 0."""
    o = np.array(local_origin, dtype=np.float64).copy()
    o[0] += -size[0] * 0.5 + (row + 0.5) * size[0] - size[0] * num_rows * 0.5
    o[1] += -size[1] * 0.5 + (col + 0.5) * size[1] - size[1] * num_cols * 0.5
    return o


def figure_eight_track(origin=(0.0, 0.0, 1.0), world_origin=(0.0, 0.0, 0.0)) -> GateTable:
    """The fixed figure-8 test track: 6 gates, zero noise, no sequence reversal.

    Gate points/eulers: L/terrains/trimesh/racing_terrains.py:350-366 with the
    RacingTestTerrainCfg of QD/terrains/racing_terrains.py:114-134 (noise 0).
    The reference draws the tile origin at random (:406-407) and reverses the
    order with p=0.5 (:386-390); the pinned C1 case uses ``origin=(0,0,1)``
    and the forward order.
    """
    pts = np.array([[3.0, 3.0, 1.0], [5.0, 0.0, 1.0], [3.0, -3.0, 1.0],
                    [-3.0, 3.0, 1.0], [-5.0, 0.0, 1.0], [-3.0, -3.0, 1.0]], dtype=np.float32)
    eul = np.array([[90.0, 90.0, 0.0], [90.0, 0.0, 0.0], [90.0, 90.0, 0.0],
                    [90.0, 90.0, 0.0], [90.0, 0.0, 0.0], [90.0, 90.0, 0.0]], dtype=np.float32)
    pose = np.zeros((1, 1, 6, 7), dtype=np.float32)
    pose[0, 0, :, :3] = pts - np.asarray(origin, dtype=np.float32)
    pose[0, 0, :, 3:] = gate_euler_to_quat_wxyz(eul)
    return GateTable(pose, np.zeros((1, 1), dtype=np.int32),
                     np.asarray(world_origin, dtype=np.float32).reshape(1, 1, 3), name="figure8")


def _ring_track(rng, difficulty, size, num_gate, radius_range=(5.0, 8.0), pos_noise=(0.2, 1.0), rot_noise=(0.0, 30.0)):
    # centre-line of SquareRacingTrackTerrain (L/terrains/trimesh/racing_terrains.py:197-219)
    radius = radius_range[1] - (radius_range[1] - radius_range[0]) * difficulty
    pn = difficulty * (pos_noise[1] - pos_noise[0]) + pos_noise[0]
    rn = difficulty * (rot_noise[1] - rot_noise[0]) + rot_noise[0]
    theta = np.linspace(0, 2 * np.pi, num_gate, endpoint=False)
    pts = np.zeros((num_gate, 3), dtype=np.float32)
    pts[:, 0] = np.cos(theta) * radius + size[0] / 2
    pts[:, 1] = np.sin(theta) * radius + size[1] / 2
    pts[:, 2] = 1.0
    eul = np.zeros((num_gate, 3), dtype=np.float32)
    eul[:, 0] = 90.0
    eul[:, 1] = theta / np.pi * 180.0
    pts += (rng.uniform(-1, 1, (num_gate, 3)) * pn).astype(np.float32)
    pts[:, 2] = pts[:, 2].clip(0.8, 2.0)
    eul[:, 1] += (rng.uniform(-1, 1, num_gate) * rn).astype(np.float32)
    return pts, eul


def _ellipse_track(rng, difficulty, size, num_gate, gate_distance=5.0, pos_noise=(0.2, 1.0), rot_noise=(0.0, 30.0)):
    # centre-line of EllipseRacingTerrain (L/terrains/trimesh/racing_terrains.py:657-712); 8 gates
    if num_gate != 8:
        return _ring_track(rng, difficulty, size, num_gate)
    pn = difficulty * (pos_noise[1] - pos_noise[0]) + pos_noise[0]
    rn = difficulty * (rot_noise[1] - rot_noise[0]) + rot_noise[0]
    a_ell, b_ell = 4.0 * gate_distance, 2.0 * gate_distance
    theta = rng.uniform(0, 2 * np.pi)
    th_deg = theta / np.pi * 180.0
    ld = np.array([np.cos(theta), np.sin(theta), 0.0])
    sd = np.array([-np.sin(theta), np.cos(theta), 0.0])
    pts = np.zeros((8, 3), dtype=np.float64)
    eul = np.zeros((8, 3), dtype=np.float32)
    eul[:, 0] = 90.0
    pts[0], pts[4] = -0.5 * a_ell * ld, 0.5 * a_ell * ld
    pts[2], pts[6] = 0.5 * b_ell * sd, -0.5 * b_ell * sd
    pts[1], pts[3] = pts[2] - gate_distance * ld, pts[2] + gate_distance * ld
    pts[5], pts[7] = pts[6] + gate_distance * ld, pts[6] - gate_distance * ld
    eul[:, 1] = th_deg + np.array([0, 90, 90, 90, 180, 270, 270, 270])
    pts[:, 0] += size[0] / 2
    pts[:, 1] += size[1] / 2
    pts[:, 2] += 1.0
    pts += rng.uniform(-1, 1, (8, 3)) * pn
    pts[:, 2] = pts[:, 2].clip(0.8, 2.0)
    eul[:, 1] += (rng.uniform(-1, 1, 8) * rn).astype(np.float32)
    return pts.astype(np.float32), eul


def _zigzag_track(rng, difficulty, size, num_gate, track_length=35.0, pos_noise=(1.0, 4.0), z_noise=(0.1, 1.0),
                  rot_noise=(0.0, 30.0)):
    # centre-line of ZigzagRacingTerrain (L/terrains/trimesh/racing_terrains.py:445-487)
    pn = difficulty * (pos_noise[1] - pos_noise[0]) + pos_noise[0]
    zn = difficulty * (z_noise[1] - z_noise[0]) + z_noise[0]
    rn = difficulty * (rot_noise[1] - rot_noise[0]) + rot_noise[0]
    theta = rng.uniform(0, 2 * np.pi)
    d = np.array([np.cos(theta), np.sin(theta), 0.0])
    lat = np.array([-d[1], d[0], 0.0])
    t = np.linspace(0, 1, num_gate)
    pts = -0.5 * track_length * d + np.outer(t, track_length * d)
    for i in range(1, num_gate - 1):
        pts[i] += 2.0 * (rng.random() - 0.5) * pn * t[i] * lat
        pts[i] += 2.0 * (rng.random() - 0.5) * zn * t[i] * np.array([0.0, 0.0, 1.0])
    eul = np.zeros((num_gate, 3), dtype=np.float32)
    eul[:, 0] = 90.0
    eul[:, 1] = theta / np.pi * 180.0 + 90
    pts[:, 0] += size[0] / 2
    pts[:, 1] += size[1] / 2
    pts[:, 2] += 1.0
    pts[:, 2] = pts[:, 2].clip(0.8, 2.0)
    eul[:, 1] += (rng.uniform(-1, 1, num_gate) * rn).astype(np.float32)
    return pts.astype(np.float32), eul


def synthetic_track_table(num_types: int = 20, num_levels: int = 10, num_gates: int = 8, seed: int = 42,
                          size=(40.0, 40.0), proportions=(0.3, 0.3, 0.4)) -> GateTable:
    """Curriculum table shaped like RacingComplexTerrainCfg (QD/terrains/racing_terrains.py:137-211):
    ``num_types`` columns split zigzag/ring/ellipse by ``proportions`` (Isaac Lab assigns columns by the
    cumulative proportion), ``num_levels`` rows of increasing difficulty, ``num_gates`` gates each.
    Generated column-major (for col: for row) like the curriculum generator, numpy seed 42 (:138)."""
    rng = np.random.default_rng(seed)
    families = (_zigzag_track, _ring_track, _ellipse_track)
    cum = np.cumsum(np.asarray(proportions, dtype=np.float64) / np.sum(proportions))
    pose = np.zeros((num_types, num_levels, num_gates, 7), dtype=np.float32)
    nxt = np.zeros((num_types, num_levels), dtype=np.int32)
    origins = np.zeros((num_levels, num_types, 3), dtype=np.float32)
    for col in range(num_types):
        fam = families[int(np.min(np.where(col / num_types + 0.001 < cum)[0]))]
        for row in range(num_levels):
            difficulty = (row + rng.uniform()) / num_levels
            pts, eul = fam(rng, difficulty, size, num_gates)
            # spawn origin: a point ~2.5 m before the first gate on the line last->first gate, random height
            # (the reference families draw it the same way: racing_terrains.py:146-158)
            back = pts[0] - pts[-1]
            back[2] = 0.0
            back = back / max(np.linalg.norm(back), 1e-6)
            local_origin = pts[0] - 2.5 * back
            local_origin[2] = rng.uniform(0.7, 1.5)
            pose[col, row, :, :3] = pts - local_origin.astype(np.float32)
            pose[col, row, :, 3:] = gate_euler_to_quat_wxyz(eul)
            nxt[col, row] = 0
            origins[row, col] = _tile_world_origin(local_origin, row, col, num_levels, num_types, size)
    return GateTable(pose, nxt, origins, name=f"synthetic{num_types}x{num_levels}x{num_gates}")
