"""UAV-vs-terrain-mesh collision count (csrc/mesh_collision.cu, generalizableracing_b200/mesh.py) against the oracle's brute-force
restatement of the reference's Warp kernel (oracle/mesh_oracle.py; L/utils/mesh_tools.py:128-295, QD/mdp/rewards.py:226-242).

The oracle is pinned by analytic cases first (Warp and trimesh are absent: there is no reference output to record); then the kernels --
BVH built by gr_mesh_build_bvh, traversed by the device code -- must reproduce its integer counts exactly on seeded inputs; a disagreement
is only tolerated where the oracle itself reports a decision margin below 1e-5 (a ray through an edge / two faces at the same distance)."""
import os
import re

import numpy as np
import pytest
import torch

from tests.conftest import backend_params
from generalizableracing_b200 import mesh as M
from generalizableracing_b200.tracks import synthetic_track_table
from oracle import mesh_oracle as O

REF = os.environ.get("GRACING_REFERENCE_ROOT", "/root/reference")


def _scene(seed=0, boxes=24):
    """ground slab + rotated boxes scattered over it"""
    rng = np.random.default_rng(seed)
    parts = [M.box_mesh((20.0, 20.0, 1.0), (0.0, 0.0, -0.5))]
    for _ in range(boxes):
        a = rng.uniform(-np.pi, np.pi, 3)
        cx, sx, cy, sy, cz, sz = np.cos(a[0]), np.sin(a[0]), np.cos(a[1]), np.sin(a[1]), np.cos(a[2]), np.sin(a[2])
        R = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        parts.append(M.box_mesh(rng.uniform(0.3, 2.0, 3), (rng.uniform(-8, 8), rng.uniform(-8, 8), rng.uniform(0.3, 3.0)), R))
    return M.merge_meshes(parts)


def _rand_quat(rng, n):
    q = rng.normal(size=(n, 4))
    return (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)


# ----------------------------------------------------------------------------------------------- the oracle, pinned analytically
def test_oracle_ray_against_a_unit_box_matches_the_analytic_answer():
    pts, faces = M.box_mesh((2.0, 2.0, 2.0))                      # [-1, 1]^3
    # from outside along +x: front face at distance 4
    hit, t, sign, _ = O.mesh_query_ray(pts, faces, [-5.0, 0.2, 0.3], [1.0, 0.0, 0.0], 1e3)
    assert hit and abs(t - 4.0) < 1e-12 and sign > 0
    # from inside: every axis ray sees a back face at distance 1 -/+ offset
    for d, expect in zip(O.AXIS_DIRS, [0.9, 1.1, 0.8, 1.2, 0.7, 1.3]):
        hit, t, sign, _ = O.mesh_query_ray(pts, faces, [0.1, 0.2, 0.3], d, 1e3)
        assert hit and abs(t - expect) < 1e-12 and sign < 0
    # pointing away from the box, and a hit beyond max_t: nothing
    assert not O.mesh_query_ray(pts, faces, [-5.0, 0.0, 0.0], [-1.0, 0.0, 0.0], 1e3)[0]
    assert not O.mesh_query_ray(pts, faces, [-5.0, 0.0, 0.0], [1.0, 0.0, 0.0], 3.5)[0]


def test_oracle_counts_lattice_points_inside_a_box():
    pts, faces = M.box_mesh((2.0, 2.0, 2.0))
    lat = M.LATTICE_TENSOR.numpy()
    ident = np.array([[1.0, 0.0, 0.0, 0.0]])
    # collider fully inside / fully outside / centre on a face (x = 1): the 1 + 4 + 4 points with x <= 0 offsets... only x < 1 count
    assert O.uav_collision_num_ray(pts, faces, [[0.0, 0.0, 0.0]], ident, 0.09, 0.05, 1e3, lat)[0][0] == 17
    assert O.uav_collision_num_ray(pts, faces, [[3.0, 0.0, 0.0]], ident, 0.09, 0.05, 1e3, lat)[0][0] == 0
    n = O.uav_collision_num_ray(pts, faces, [[1.0 + 1e-4, 0.013, 0.021]], ident, 0.09, 0.05, 1e3, lat)[0][0]
    assert n == 8                                                 # the 4 + 4 points with a negative x offset
    # lattice-free branch: the last ray that hits decides (down: back face of the bottom from inside -> 1; outside above: front -> 0)
    assert O.uav_collision_num_ray(pts, faces, [[0.0, 0.0, 0.0]], ident, 0.09, 0.05, 1e3, None)[0][0] == 1
    assert O.uav_collision_num_ray(pts, faces, [[0.0, 0.0, 5.0]], ident, 0.09, 0.05, 1e3, None)[0][0] == 0
    pen, _ = O.collision_penalty_custom(pts, faces, [[0.0, 0.0, 0.0], [3.0, 0.0, 0.0]], np.repeat(ident, 2, 0), lat)
    assert pen.tolist() == [1.0, 0.0]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_lattice_is_the_reference_lattice():
    src = open(os.path.join(REF, "extensions/diff.lab/diff/lab/utils/__init__.py")).read()
    body = src[src.index("LATTICE_TENSOR"):]
    rows = re.findall(r"\[\s*(-?[\d.]+),\s*(-?[\d.]+),\s*(-?[\d.]+)\]", body)
    ref = np.array(rows, dtype=np.float32)
    assert ref.shape == (17, 3) and np.array_equal(ref, M.LATTICE_TENSOR.numpy())


def test_box_meshes_are_closed_and_outward():
    pts, faces = _scene(3, boxes=5)
    p = pts.astype(np.float64)
    n = np.cross(p[faces[:, 1]] - p[faces[:, 0]], p[faces[:, 2]] - p[faces[:, 0]])
    for b in range(6):                                            # per box: normals point away from the centroid, signed volume > 0
        f = faces[12 * b: 12 * b + 12]
        c = p[np.unique(f)].mean(0)
        assert (np.einsum("ij,ij->i", n[12 * b: 12 * b + 12], p[f[:, 0]] - c) > 0).all()
        assert np.einsum("ij,ij->i", p[f[:, 0]], np.cross(p[f[:, 1]], p[f[:, 2]])).sum() > 0


# ----------------------------------------------------------------------------------------------- the BVH (host code of libgracing.so)
def test_bvh_covers_every_face_once_and_bounds_it():
    pts, faces = _scene(1)
    m = M.TerrainMesh(pts, faces, device="cpu", _lib=object())
    assert sorted(m.face_ids.tolist()) == list(range(len(faces)))
    nodes = m.nodes_host
    lo, hi = nodes[:, :3], nodes[:, 4:7]
    first, count = nodes[:, 3].copy().view(np.int32), nodes[:, 7].copy().view(np.int32)
    seen = np.zeros(len(faces), dtype=np.int32)
    reached = np.zeros(m.num_nodes, dtype=bool)
    stack = [0]
    while stack:
        k = stack.pop()
        reached[k] = True
        if count[k] > 0:
            assert count[k] <= 4
            for f in range(first[k], first[k] + count[k]):
                seen[f] += 1
                tri = m.tris_host[f].reshape(3, 4)[:, :3]
                v = np.stack([tri[0], tri[0] + tri[1], tri[0] + tri[2]])
                assert (v >= lo[k] - 1e-6).all() and (v <= hi[k] + 1e-6).all()
                assert np.allclose(v, pts[faces[m.face_ids[f]]], atol=1e-6)
        else:
            for c in (first[k], first[k] + 1):
                assert (lo[c] >= lo[k] - 1e-6).all() and (hi[c] <= hi[k] + 1e-6).all()
                stack.append(int(c))
    assert (seen == 1).all() and reached.all()


# ----------------------------------------------------------------------------------------------- kernels vs oracle
def _assert_counts(got, want, margins, what):
    bad = np.nonzero(got != want)[0]
    assert all(margins[b] < 1e-5 for b in bad), (what, bad[:8], got[bad[:8]], want[bad[:8]], margins[bad[:8]])
    assert len(bad) <= max(1, len(got) // 500), (what, len(bad))


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_collision_counts_match_the_oracle(backend):
    device, lib = backend
    pts, faces = _scene(0)
    mesh = M.TerrainMesh(pts, faces, device=device, _lib=lib)
    rng = np.random.default_rng(5)
    n = 240
    pos = np.stack([rng.uniform(-9, 9, n), rng.uniform(-9, 9, n), rng.uniform(-0.3, 3.5, n)], 1).astype(np.float32)
    quat = _rand_quat(rng, n)
    lat = M.LATTICE_TENSOR
    got = M.get_uav_collision_num_ray(mesh, torch.from_numpy(pos), torch.from_numpy(quat), 0.09, 0.05, 1e3, lat).cpu().numpy()
    want, margins = O.uav_collision_num_ray(pts, faces, pos, quat, 0.09, 0.05, 1e3, lat.numpy())
    assert got.dtype == np.int32 and (want > 0).sum() > 20 and (want == 0).sum() > 20 and ((want > 0) & (want < 17)).sum() > 5
    _assert_counts(got, want, margins, "lattice")
    # a larger collider (more partial overlaps), a short ray (max_dist below most hit distances), and the centre-only branch
    got = M.get_uav_collision_num_ray(mesh, torch.from_numpy(pos), torch.from_numpy(quat), 0.6, 0.4, 1e3, lat).cpu().numpy()
    want, margins = O.uav_collision_num_ray(pts, faces, pos, quat, 0.6, 0.4, 1e3, lat.numpy())
    _assert_counts(got, want, margins, "large collider")
    got = M.get_uav_collision_num_ray(mesh, torch.from_numpy(pos), torch.from_numpy(quat), 0.09, 0.05, 0.4, lat).cpu().numpy()
    want, margins = O.uav_collision_num_ray(pts, faces, pos, quat, 0.09, 0.05, 0.4, lat.numpy())
    _assert_counts(got, want, margins, "short rays")
    got = M.get_uav_collision_num_ray(mesh, torch.from_numpy(pos), torch.from_numpy(quat), 0.09, 0.05, 1e3, None).cpu().numpy()
    want, margins = O.uav_collision_num_ray(pts, faces, pos, quat, 0.09, 0.05, 1e3, None)
    _assert_counts(got, want, margins, "centre only")
    pen = M.collision_penalty_custom(mesh, torch.from_numpy(pos), torch.from_numpy(quat)).cpu().numpy()
    want_pen, margins = O.collision_penalty_custom(pts, faces, pos, quat, lat.numpy())
    assert (pen != want_pen).sum() <= 1


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_ray_queries_match_the_oracle(backend):
    device, lib = backend
    pts, faces = _scene(2)
    mesh = M.TerrainMesh(pts, faces, device=device, _lib=lib)
    rng = np.random.default_rng(9)
    n = 600
    org = np.stack([rng.uniform(-9, 9, n), rng.uniform(-9, 9, n), rng.uniform(0.05, 4.0, n)], 1).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d[: n // 3] = O.AXIS_DIRS[rng.integers(0, 6, n // 3)]                                  # a third of the rays axis-aligned
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    t, s = mesh.query_rays(torch.from_numpy(org), torch.from_numpy(d), 50.0)
    t, s = t.cpu().numpy(), s.cpu().numpy()
    hits = 0
    for i in range(n):
        hit, tt, sign, margin = O.mesh_query_ray(pts, faces, org[i], d[i], 50.0)
        if margin < 1e-5:
            continue
        assert s[i] == sign, (i, s[i], sign, margin)
        assert abs(t[i] - tt) <= 1e-4 * max(1.0, tt), (i, t[i], tt)
        hits += int(hit)
    assert hits > n // 4


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_track_table_mesh_gates_collide_and_openings_do_not(backend):
    device, lib = backend
    table = synthetic_track_table()
    pts, faces = M.track_table_mesh(table, gate_inner=(1.5, 1.5), gate_edge=0.15, gate_thickness=0.1)
    mesh = M.TerrainMesh(pts, faces, device=device, _lib=lib)
    ty, lv = 1, 2
    origin = table.terrain_origins[lv, ty].astype(np.float64)
    pose = table.gate_pose[ty, lv]
    R = [M._quat_to_matrix_wxyz(q) for q in pose[:, 3:7]]
    centre = pose[:, :3].astype(np.float64) + origin
    in_bar = np.stack([centre[g] + R[g] @ np.array([0.825, 0.0, 0.0]) for g in range(len(R))])          # middle of the left bar
    ident = np.tile(np.array([[1.0, 0.0, 0.0, 0.0]], dtype=np.float32), (len(R), 1))
    P = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    # mass point in the opening: no collision; in a bar: all rays see back faces; below the ground surface: inside the slab
    assert M.get_uav_collision_num_ray(mesh, P(centre), P(ident), 0.09, 0.05, 1e3, None).sum() == 0
    assert (M.get_uav_collision_num_ray(mesh, P(in_bar), P(ident), 0.01, 0.01, 1e3, M.LATTICE_TENSOR).cpu().numpy() == 17).all()
    under = centre.copy(); under[:, 2] = -0.3
    assert (M.collision_penalty_custom(mesh, P(under), P(ident)).cpu().numpy() == 1.0).all()
    assert (M.collision_penalty_custom(mesh, P(centre), P(ident)).cpu().numpy() == 0.0).all()
    want, _ = O.uav_collision_num_ray(pts, faces, in_bar, ident, 0.09, 0.05, 1e3, M.LATTICE_TENSOR.numpy())
    got = M.get_uav_collision_num_ray(mesh, P(in_bar), P(ident), 0.09, 0.05, 1e3, M.LATTICE_TENSOR).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_collision_at_c4_size_properties(cuda_lib):
    """65,536 UAVs on the bench's gate table: counts are in [0, 17], translation of mesh + UAVs together leaves them unchanged,
    and the count of a UAV does not depend on which other UAVs are in the batch (thread mapping / atomics)."""
    from generalizableracing_b200.track_gen import generate_track_table, racing_complex_cfg
    table = generate_track_table(racing_complex_cfg())
    pts, faces = M.track_table_mesh(table)
    mesh = M.TerrainMesh(pts, faces)
    g = torch.Generator(device="cuda").manual_seed(0)
    N = 65536
    types = torch.randint(0, table.gate_pose.shape[0], (N,), device="cuda", generator=g)
    levels = torch.randint(0, table.gate_pose.shape[1], (N,), device="cuda", generator=g)
    gates = torch.randint(0, table.gate_pose.shape[2], (N,), device="cuda", generator=g)
    gp = torch.from_numpy(table.gate_pose).cuda()[types, levels, gates, :3] + torch.from_numpy(table.terrain_origins).cuda()[levels, types]
    pos = gp + torch.randn(N, 3, device="cuda", generator=g) * torch.tensor([1.0, 1.0, 0.8], device="cuda")
    quat = torch.nn.functional.normalize(torch.randn(N, 4, device="cuda", generator=g), dim=-1)
    num = M.get_uav_collision_num_ray(mesh, pos, quat, 0.09, 0.05, 1e3, M.LATTICE_TENSOR)
    assert int(num.min()) >= 0 and int(num.max()) <= 17 and int((num > 0).sum()) > 100 and int((num == 0).sum()) > N // 2
    sub = torch.arange(0, N, 7, device="cuda")
    assert torch.equal(M.get_uav_collision_num_ray(mesh, pos[sub], quat[sub], 0.09, 0.05, 1e3, M.LATTICE_TENSOR), num[sub])
    shift = np.array([1000.0, -2000.0, 0.0], dtype=np.float32)
    mesh2 = M.TerrainMesh(pts + shift, faces)
    num2 = M.get_uav_collision_num_ray(mesh2, pos + torch.from_numpy(shift).cuda(), quat, 0.09, 0.05, 1e3, M.LATTICE_TENSOR)
    assert int((num2 != num).sum()) <= N // 500            # fp32 at |x| ~ 2000 m (1.2e-4 m spacing): lattice points that close to a face of a 0.1 m bar may flip (measured: 42)
    # against the oracle on a sample
    k = torch.arange(0, N, 2003, device="cuda")
    want, margins = O.uav_collision_num_ray(pts, faces, pos[k].cpu().numpy(), quat[k].cpu().numpy(), 0.09, 0.05, 1e3, M.LATTICE_TENSOR.numpy())
    _assert_counts(num[k].cpu().numpy(), want, margins, "c4 sample")


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_env_step_adds_the_collision_term_on_the_pre_reset_pose(backend):
    """RacingVecEnv.set_terrain_mesh: reward += -50 * dt * collision_penalty_custom(pose after the step's physics, before any reset)
    (QD/racing_ctbr_env.py:299-303, QD/mdp/rewards.py:226-242); everything else of the step is untouched."""
    from generalizableracing_b200.config import RacingCfg
    from generalizableracing_b200.env import RacingVecEnv
    from generalizableracing_b200.tracks import GateTable
    device, lib = backend
    full = synthetic_track_table()
    table = GateTable(full.gate_pose[:2, :2].copy(), full.next_gate_id[:2, :2].copy(), full.terrain_origins[:2, :2].copy(), "2x2 tiles")
    cfg, N = RacingCfg.for_stage(0), 48
    pts, faces = M.track_table_mesh(table)                     # 4 tiles: the brute-force oracle stays fast
    mesh = M.TerrainMesh(pts, faces, device=device, _lib=lib)
    a, b = (RacingVecEnv(cfg, table, N, device=device, seed=11, _lib=lib) for _ in range(2))
    b.set_terrain_mesh(mesh, -50.0)
    a.reset(); b.reset()
    g = torch.Generator().manual_seed(3)
    hits = resets = 0
    for t in range(48):
        act = (torch.randn(N, 4, generator=g) * (0.5 if t % 3 else 2.5)).to(device)
        oa, ra, da, _ = a.step(act)
        ob, rb, db, ex = b.step(act)
        pen = ex["collision_penalty"]
        assert torch.equal(oa, ob) and torch.equal(da, db)
        assert torch.equal(rb, ra + pen * (-50.0 * cfg.step_dt))
        pos, quat = b._pre_pose
        if t % 4 == 3:
            want, margins = O.collision_penalty_custom(pts, faces, pos.cpu().numpy(), quat.cpu().numpy(), M.LATTICE_TENSOR.numpy())
            assert (pen.cpu().numpy() != want).sum() <= 1
        sv = b.state_dict_view()
        keep = ~db.bool()                                       # envs that did not reset: the exported pose IS the stored state
        assert torch.equal(sv["root_pos_w"][keep], pos[keep]) and torch.equal(sv["root_quat_w"][keep], quat[keep])
        hits += int(pen.sum()); resets += int(db.sum())
    assert hits > 0 and resets > 0
    assert torch.equal(a.planes, b.planes)

