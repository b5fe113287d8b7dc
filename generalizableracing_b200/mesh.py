"""Terrain meshes and the UAV collision count (last item).

Mirrors, relative to /root/reference (L = extensions/diff.lab/diff/lab, QD = extensions/diff.lab_tasks/.../quadcopter_diff):

* ``get_uav_collision_num_ray`` -- L/utils/mesh_tools.py:237-295 (launcher of the Warp kernel ``check_uav_collision_ray_kernel``,
  :128-233), same argument order and meaning; ``LATTICE_TENSOR`` -- L/utils/__init__.py:19-37;
* ``collision_penalty_custom`` -- QD/mdp/rewards.py:226-242 (the STAGE-0 reward term, weight -50: QD/racing_ctbr_env.py:299-303);
* ``TerrainMesh`` -- what ``TerrainImporter.warp_meshes["terrain"]`` is to the reference (a ``wp.Mesh``: points + face indices with an
  acceleration structure, L/terrains/terrain_importer.py + omni.isaac.lab ``convert_to_warp_mesh``).

The ray casts run in ``libgracing.so`` (csrc/mesh_collision.cu: BVH built on the host by ``gr_mesh_build_bvh``, traversed on the
device).  There is no CPU fallback.  The reference's terrain MESHES (trimesh boxes / cylinders / spheres / capsules / boolean gate
frames, USD import) stay out of scope; :func:`track_table_mesh` builds a plain box model of a gate table -- ground
slab plus one four-bar frame per gate -- so that the collision term can be exercised on the tracks this package generates.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib as B

# L/utils/__init__.py:19-37: centre, the 8 corners and the 8 half-way points of the box collider
LATTICE_TENSOR = torch.tensor(
    [[0.0, 0.0, 0.0]] +
    [[sx, sy, sz] for sz in (1.0, -1.0) for sx in (1.0, -1.0) for sy in (1.0, -1.0)] +
    [[sx, sy, sz] for sz in (0.5, -0.5) for sx in (0.5, -0.5) for sy in (0.5, -0.5)], dtype=torch.float32)


class TerrainMesh:
    """Triangle mesh + BVH resident on a CUDA device (the role of ``wp.Mesh``)."""

    def __init__(self, points, indices, device="cuda:0", _lib=None):
        self.device = torch.device(device)
        if _lib is None:
            if self.device.type != "cuda":
                raise RuntimeError("TerrainMesh lives on a CUDA device (sm_100a); there is no CPU fallback")
            _lib = B.load()
        self._lib = _lib
        pts = np.ascontiguousarray(np.asarray(points, dtype=np.float32).reshape(-1, 3))
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.int32).reshape(-1, 3))
        if pts.shape[0] == 0 or idx.shape[0] == 0:
            raise ValueError("empty mesh")
        if idx.min() < 0 or idx.max() >= pts.shape[0]:
            raise ValueError("face index out of range")
        self.points, self.indices = pts, idx
        F = idx.shape[0]
        lib = B.load()                      # the builder is host code of the same library (runs without a GPU)
        max_nodes = int(lib.gr_mesh_bvh_max_nodes(F))
        nodes = np.zeros((max_nodes, 8), dtype=np.float32)
        tris = np.zeros((F, 12), dtype=np.float32)
        face_ids = np.zeros(F, dtype=np.int32)
        n = C.c_int32(0)
        B.check(lib.gr_mesh_build_bvh(pts.ctypes.data, idx.ctypes.data, pts.shape[0], F, nodes.ctypes.data, max_nodes, tris.ctypes.data,
                                      face_ids.ctypes.data, C.byref(n)), "gr_mesh_build_bvh")
        self.num_nodes, self.num_faces = int(n.value), F
        self.nodes_host, self.tris_host, self.face_ids = nodes[: self.num_nodes].copy(), tris, face_ids
        self._nodes = torch.from_numpy(self.nodes_host).to(self.device).contiguous()
        self._tris = torch.from_numpy(tris).to(self.device).contiguous()
        self._mesh = B.GrMesh(self._nodes.data_ptr(), self._tris.data_ptr(), self.num_nodes, F)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else None

    def query_rays(self, origins: torch.Tensor, dirs: torch.Tensor, max_t: float = 1.0e6):
        """``wp.mesh_query_ray`` over arrays of rays: returns (t [R], sign [R]); t = max_t and sign = 0 where nothing was hit,
        sign = +1 for a front-face hit, -1 for a back-face hit."""
        o = origins.to(self.device, torch.float32).reshape(-1, 3).contiguous()
        d = dirs.to(self.device, torch.float32).reshape(-1, 3).contiguous()
        if o.shape != d.shape:
            raise ValueError("origins and dirs must have the same shape")
        t = torch.empty(o.shape[0], device=self.device)
        s = torch.empty(o.shape[0], device=self.device)
        B.check(self._lib.gr_mesh_query_rays(C.byref(self._mesh), o.data_ptr(), d.data_ptr(), o.shape[0], float(max_t), t.data_ptr(), s.data_ptr(),
                                             self._stream()), "gr_mesh_query_rays")
        return t, s


def get_uav_collision_num_ray(mesh: TerrainMesh, uav_position: torch.Tensor, uav_orientation: torch.Tensor, arm_length: float, height: float,
                              max_dist: float = 1e6, lattices: Optional[torch.Tensor] = None) -> torch.Tensor:
    """L/utils/mesh_tools.py:237-295.  ``uav_position`` [N,3], ``uav_orientation`` [N,4] (w,x,y,z), ``lattices`` [M,3] or None (centre
    point only) -> int32 [N]: the number of lattice points of each UAV's box collider that lie inside the terrain mesh.  Unlike the
    reference launcher there is no ``wp.synchronize()``: the result is ordered on the current stream like any torch op."""
    dev = mesh.device
    pos = uav_position.to(dev, torch.float32).reshape(-1, 3).contiguous()
    quat = uav_orientation.to(dev, torch.float32).reshape(-1, 4).contiguous()
    n = pos.shape[0]
    if quat.shape[0] != n:
        raise ValueError("uav_position and uav_orientation disagree on the number of UAVs")
    out = torch.empty(n, dtype=torch.int32, device=dev)
    lat_ptr, n_lat = None, 0
    if lattices is not None:
        lat = lattices.to(dev, torch.float32).reshape(-1, 3).contiguous()
        lat_ptr, n_lat = lat.data_ptr(), lat.shape[0]
    B.check(mesh._lib.gr_uav_collision_ray(C.byref(mesh._mesh), pos.data_ptr(), quat.data_ptr(), n, lat_ptr, n_lat, float(max_dist), float(arm_length),
                                           float(height), out.data_ptr(), mesh._stream()), "gr_uav_collision_ray")
    return out.to(uav_position.device)


def collision_penalty_custom(mesh: TerrainMesh, root_pos_w: torch.Tensor, root_quat_w: torch.Tensor) -> torch.Tensor:
    """QD/mdp/rewards.py:226-242 with the scene look-ups replaced by their values: 0.09 m arms, 0.05 m height, rays of 1e3 m, the
    17-point lattice; 1.0 where more than two lattice points are inside the terrain."""
    num = get_uav_collision_num_ray(mesh, root_pos_w, root_quat_w, 0.09, 0.05, 1e3, LATTICE_TENSOR)
    return (num > 2.0).float()


# ---------------------------------------------------------------------------------------------------------------------------
# box meshes (outward-facing, counter-clockwise faces)
# ---------------------------------------------------------------------------------------------------------------------------
_BOX_CORNERS = np.array([[sx, sy, sz] for sx in (-0.5, 0.5) for sy in (-0.5, 0.5) for sz in (-0.5, 0.5)], dtype=np.float64)
_BOX_FACES = np.array([[0, 1, 3], [0, 3, 2],      # -x
                       [4, 6, 7], [4, 7, 5],      # +x
                       [0, 4, 5], [0, 5, 1],      # -y
                       [2, 3, 7], [2, 7, 6],      # +y
                       [0, 2, 6], [0, 6, 4],      # -z
                       [1, 5, 7], [1, 7, 3]], dtype=np.int32)      # +z


def _quat_to_matrix_wxyz(q) -> np.ndarray:
    w, x, y, z = [float(v) for v in q]
    n = w * w + x * x + y * y + z * z
    s = 2.0 / n if n > 0 else 0.0
    return np.array([[1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
                     [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
                     [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)]])


def box_mesh(extents, position=(0.0, 0.0, 0.0), rotation: Optional[np.ndarray] = None):
    """Closed box: (points [8,3], faces [12,3]); ``rotation`` = 3x3 matrix applied before the translation."""
    pts = _BOX_CORNERS * np.asarray(extents, dtype=np.float64)
    if rotation is not None:
        pts = pts @ np.asarray(rotation, dtype=np.float64).T
    return (pts + np.asarray(position, dtype=np.float64)).astype(np.float32), _BOX_FACES.copy()


def merge_meshes(parts):
    """[(points, faces), ...] -> one (points, faces) with re-based indices."""
    pts, faces, base = [], [], 0
    for p, f in parts:
        pts.append(np.asarray(p, dtype=np.float32))
        faces.append(np.asarray(f, dtype=np.int32) + base)
        base += len(p)
    return np.concatenate(pts, axis=0), np.concatenate(faces, axis=0)


def gate_frame_mesh(inner_w: float, inner_h: float, edge: float, thickness: float, position, quat_wxyz):
    """A gate as four bars around an ``inner_w`` x ``inner_h`` opening (the solid the reference obtains as the boolean difference of two
    boxes of equal depth, L/terrains/trimesh/utils.py:10-33), in the gate frame of the pose table: x to the left, y up, z through the
    opening."""
    R = _quat_to_matrix_wxyz(quat_wxyz)
    pos = np.asarray(position, dtype=np.float64)
    outer_h = inner_h + 2 * edge
    bars = [((edge, outer_h, thickness), (-(inner_w + edge) / 2, 0.0, 0.0)), ((edge, outer_h, thickness), ((inner_w + edge) / 2, 0.0, 0.0)),
            ((inner_w, edge, thickness), (0.0, (inner_h + edge) / 2, 0.0)), ((inner_w, edge, thickness), (0.0, -(inner_h + edge) / 2, 0.0))]
    return merge_meshes([box_mesh(ext, pos + R @ np.asarray(off), R) for ext, off in bars])


def track_table_mesh(table, tile_size=None, gate_inner=(1.5, 1.5), gate_edge: float = 0.15, gate_thickness: float = 0.1,
                     ground_thickness: float = 1.0, ground_margin: float = 10.0):
    """Box model of every tile of a gate table: a ground slab below z = 0 (the reference's ``terrain_height = 1.0`` slab,
    racing_terrains.py:322-327) and a four-bar frame per gate at the table's pose.  ``tile_size`` = (x, y) extents of the slab, centred
    on the tile's spawn origin; None: the bounding rectangle of the tile's gates grown by ``ground_margin``.  World frame = the frame
    of ``table.terrain_origins`` (what ``root_pos_w`` of an env lives in)."""
    parts = []
    types, levels, gates, _ = table.gate_pose.shape
    for ty in range(types):
        for lv in range(levels):
            origin = table.terrain_origins[lv, ty].astype(np.float64)
            if tile_size is None:
                xy = table.gate_pose[ty, lv, :, :2].astype(np.float64) + origin[:2]
                lo, hi = xy.min(0) - ground_margin, xy.max(0) + ground_margin
                parts.append(box_mesh((hi[0] - lo[0], hi[1] - lo[1], ground_thickness), ((lo[0] + hi[0]) / 2, (lo[1] + hi[1]) / 2, -ground_thickness / 2)))
            else:
                parts.append(box_mesh((tile_size[0], tile_size[1], ground_thickness), (origin[0], origin[1], -ground_thickness / 2)))
            for g in range(gates):
                pose = table.gate_pose[ty, lv, g]
                parts.append(gate_frame_mesh(gate_inner[0], gate_inner[1], gate_edge, gate_thickness, pose[:3].astype(np.float64) + origin, pose[3:7]))
    return merge_meshes(parts)
