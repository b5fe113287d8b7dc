"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

CPU restatement (numpy, brute force over all faces) of the reference's UAV collision count:
extensions/diff.lab/diff/lab/utils/mesh_tools.py:128-233 (Warp kernel ``check_uav_collision_ray_kernel``) and :237-295 (its launcher
``get_uav_collision_num_ray``), and of the reward term built on it, quadcopter_diff/mdp/rewards.py:226-242.

PARITY UNPINNED against the reference: the kernel needs NVIDIA Warp (``wp.mesh_query_ray`` against a ``wp.Mesh`` BVH) and the
terrain meshes need trimesh -- both third party, both absent from this image and from /root/reference -- and the reference's own test of
this path (standalone/diff_rl/test/test_collider.py:131-141) only times the call inside Isaac Sim, it asserts nothing.  What is restated:

* the kernel's control flow, line by line (lattice offsets 0.707 * arm / 0.5 * height rotated by the attitude quaternion, the order of the
  six axis rays, "count the point and go to the next one on the first back-face hit", the overwrite semantics of the lattice-free branch);
* ``wp.mesh_query_ray(mesh, start, dir, max_t)`` from its published contract (Warp documentation of ``mesh_query_ray``): the CLOSEST
  intersection with 0 <= t <= max_t over all faces, faces two-sided, ``sign > 0`` if the ray hit the front of the face (the side the
  counter-clockwise normal points to), ``< 0`` otherwise.  Intersection test: Moeller-Trumbore, fp64 here.

What pins this oracle instead: analytic cases in tests/test_mesh_collision.py (points inside / outside closed boxes, rays along known
axes with known hit distances), independent of any mesh library.
"""
from __future__ import annotations

import numpy as np

# mesh_tools.py:150-155: front, back, left, right, up, down
AXIS_DIRS = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64)


def mesh_query_ray(points: np.ndarray, faces: np.ndarray, origin: np.ndarray, direction: np.ndarray, max_t: float):
    """Closest hit of ONE ray against every face.  Returns (hit, t, sign, margin); margin = distance of the decision from its nearest
    tie / edge (smallest of: barycentric slack of the winning face, gap in t to the runner-up of opposite facing) -- tests use it to tell
    a genuine mismatch from a coin toss at an edge."""
    p = np.asarray(points, dtype=np.float64)
    f = np.asarray(faces)
    o = np.asarray(origin, dtype=np.float64)
    d = np.asarray(direction, dtype=np.float64)
    v0, e1, e2 = p[f[:, 0]], p[f[:, 1]] - p[f[:, 0]], p[f[:, 2]] - p[f[:, 0]]
    pv = np.cross(d, e2)
    det = np.einsum("ij,ij->i", e1, pv)
    ok = det != 0.0
    inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
    tv = o - v0
    u = np.einsum("ij,ij->i", tv, pv) * inv
    q = np.cross(tv, e1)
    v = np.einsum("j,ij->i", d, q) * inv
    t = np.einsum("ij,ij->i", e2, q) * inv
    inside = ok & (u >= 0.0) & (u <= 1.0) & (v >= 0.0) & (u + v <= 1.0) & (t >= 0.0) & (t <= max_t)
    if not inside.any():
        # how close did any face come to being hit (in barycentric units)?
        near = ok & (t >= 0.0) & (t <= max_t)
        slack = np.minimum(np.minimum(u, v), 1.0 - u - v)
        margin = float(-slack[near].max()) if near.any() else np.inf
        return False, float(max_t), 0.0, margin
    tt = np.where(inside, t, np.inf)
    k = int(np.argmin(tt))
    sign = 1.0 if det[k] > 0.0 else -1.0
    slack_k = float(min(u[k], v[k], 1.0 - u[k] - v[k]))
    other = inside & ((det > 0.0) != (det[k] > 0.0))
    gap = float((tt[other] - tt[k]).min()) if other.any() else np.inf
    return True, float(tt[k]), sign, min(slack_k, gap)


def quat_rotate_xyzw(q_xyzw: np.ndarray, v: np.ndarray) -> np.ndarray:
    """wp.quat_rotate: v (2w^2 - 1) + 2w (u x v) + 2u (u . v)."""
    u, w = q_xyzw[:3], q_xyzw[3]
    return v * (2.0 * w * w - 1.0) + 2.0 * w * np.cross(u, v) + 2.0 * u * np.dot(u, v)


def uav_collision_num_ray(points, faces, uav_position, uav_orientation_wxyz, arm_length, height, max_dist=1e6, lattices=None):
    """get_uav_collision_num_ray (mesh_tools.py:237-295).  Returns (num_collisions int32 [N], margin [N]): margin = the smallest decision
    margin over the rays that decided the env's count."""
    pos = np.asarray(uav_position, dtype=np.float64).reshape(-1, 3)
    quat = np.asarray(uav_orientation_wxyz, dtype=np.float64).reshape(-1, 4)
    quat_xyzw = np.concatenate([quat[:, 1:], quat[:, 0:1]], axis=1)            # :261
    n = pos.shape[0]
    out = np.zeros(n, dtype=np.int32)
    margins = np.full(n, np.inf)
    for i in range(n):
        if lattices is None:                                                    # :157-187: every hit overwrites
            for d in AXIS_DIRS:
                hit, _, sign, m = mesh_query_ray(points, faces, pos[i], d, max_dist)
                margins[i] = min(margins[i], m)
                if hit:
                    out[i] = 1 if sign <= 0.0 else 0
            continue
        lat = np.asarray(lattices, dtype=np.float64).reshape(-1, 3)
        for k in range(lat.shape[0]):                                           # :188-233
            vec = np.array([lat[k, 0] * 0.707 * arm_length, lat[k, 1] * 0.707 * arm_length, lat[k, 2] * 0.5 * height])
            pt = pos[i] + quat_rotate_xyzw(quat_xyzw[i], vec)
            for d in AXIS_DIRS:
                hit, _, sign, m = mesh_query_ray(points, faces, pt, d, max_dist)
                margins[i] = min(margins[i], m)
                if hit and sign <= 0.0:
                    out[i] += 1
                    break
    return out, margins


def collision_penalty_custom(points, faces, root_pos_w, root_quat_w, lattice):
    """rewards.py:226-242."""
    num, margins = uav_collision_num_ray(points, faces, root_pos_w, root_quat_w, 0.09, 0.05, 1e3, lattice)
    return (num > 2.0).astype(np.float32), margins
