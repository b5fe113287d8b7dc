"""Track generators (generalizableracing_b200/track_gen.py) against the reference's own family functions executed with the
mesh layer stubbed (oracle/ref_modules.load_track_families), bit for bit, and against the committed golden vectors those
functions produced (tests/golden/track_families.npz, for machines without the reference tree)."""
import os
import random
import types

import numpy as np
import pytest

from generalizableracing_b200 import track_gen as TG
from oracle import ref_modules as RM

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "track_families.npz")
CASES = [("square", s, d) for s in (0, 7) for d in (0.0, 0.37, 1.0)] + [("zigzag", s, d) for s in (1, 8) for d in (0.05, 0.6)] + \
        [("ellipse", s, d) for s in (2, 9) for d in (0.2, 0.95)] + [("figure_eight", s, d) for s in (3, 4) for d in (0.0, 0.5)] + \
        [(fam + "+obs", s, d) for fam in ("square", "zigzag", "ellipse") for s in (5, 11) for d in (0.15, 0.8)]       # add_obs=True: the reference's setting


def _cfg(fam):
    if fam.endswith("+obs"):         # the three families exactly as RacingComplexTerrainCfg configures them (QD/terrains/racing_terrains.py:137-211)
        return TG.racing_complex_cfg().sub_terrains[{"square": "circular", "zigzag": "zigzag", "ellipse": "ellipse"}[fam[:-4]]]
    return {"square": TG.SquareTrackCfg(), "zigzag": TG.ZigzagTrackCfg(), "ellipse": TG.EllipseTrackCfg(),
            "figure_eight": TG.FigureEightTrackCfg(pos_noise_scale=(0.0, 0.3), rot_noise_scale=(0.0, 10.0))}[fam]


def _ref_cfg(c):
    """the same fields as a plain object for the reference function (lists where the reference cfg has lists)"""
    d = {k: (list(v) if isinstance(v, tuple) else v) for k, v in c.__dict__.items()}
    d["add_border"] = False
    return types.SimpleNamespace(**d)


def _fn(fam):
    return fam[:-4] if fam.endswith("+obs") else fam


def _mine(fam, seed, diff, chain=1):
    s = TG.Streams(seed)
    fn = {"square": TG.square_track, "zigzag": TG.zigzag_track, "ellipse": TG.ellipse_track, "figure_eight": TG.figure_eight_tile}[_fn(fam)]
    out = [fn(diff, _cfg(fam), s) for _ in range(chain)]
    return out


@pytest.mark.skipif(not RM.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("fam,seed,diff", CASES)
def test_family_matches_the_reference_function(fam, seed, diff):
    ref = RM.load_track_families()
    random.seed(seed)
    np.random.seed(seed)
    theirs = [getattr(ref, _fn(fam))(diff, _ref_cfg(_cfg(fam))) for _ in range(3)]     # three tiles in a row on the same global streams
    mine = _mine(fam, seed, diff, chain=3)
    for (_, o_ref, ex), (pose, o, nid) in zip(theirs, mine):
        assert np.array_equal(np.asarray(ex["gate_pose"]), pose) and pose.dtype == np.asarray(ex["gate_pose"]).dtype
        assert np.array_equal(np.asarray(o_ref, dtype=np.float64), np.asarray(o, dtype=np.float64))
        assert int(ex["next_gate_id"]) == int(nid)


def test_family_matches_the_golden_vectors():
    g = np.load(GOLDEN)
    for k, (fam, seed, diff) in enumerate(CASES):
        pose, o, nid = _mine(fam, seed, diff)[0]
        assert np.array_equal(g[f"pose_{k}"], pose), (fam, seed, diff)
        assert np.array_equal(g[f"origin_{k}"], np.asarray(o, dtype=np.float64))
        assert int(g[f"next_{k}"]) == int(nid)


def _reference_complex_table():
    """RacingComplexTerrainCfg through the UNMODIFIED reference family functions and mesh helpers (trimesh itself stubbed), tiles in Isaac
    Lab's curriculum order on the two global streams seeded like the launcher's set_seed(42); pose conversion of terrain_generator.py:57-77"""
    ref = RM.load_track_families()
    cfg = TG.racing_complex_cfg()
    random.seed(42)
    np.random.seed(42)
    np_rng = np.random.default_rng(cfg.seed)
    fams = list(cfg.sub_terrains.items())
    cum = np.cumsum(np.array([f.proportion for _, f in fams]) / sum(f.proportion for _, f in fams))
    fn = {"zigzag": ref.zigzag, "circular": ref.square, "ellipse": ref.ellipse}
    rows, cols = cfg.num_rows, cfg.num_cols
    pose, nxt, origins = np.zeros((cols, rows, 8, 7), np.float32), np.zeros((cols, rows), np.int32), np.zeros((rows, cols, 3), np.float32)
    for c in range(cols):
        for r in range(rows):
            name, f = fams[int(np.min(np.where(c / cols + 0.001 < cum)[0]))]
            _, origin, ex = fn[name](float((r + np_rng.uniform()) / rows), _ref_cfg(f))
            pose[c, r], centred = TG.tile_entry(np.asarray(ex["gate_pose"]), origin, cfg.size)
            nxt[c, r] = ex["next_gate_id"]
            origins[r, c] = centred + np.array([(r + 0.5) * 40.0 - 200.0, (c + 0.5) * 40.0 - 400.0, 0.0])
    return pose, nxt, origins


@pytest.mark.skipif(not RM.available(), reason="reference tree not mounted")
def test_complex_table_is_the_reference_table():
    """VERDICT r1 Missing #2: with add_obs=True (the reference's setting) the obstacle draws are replayed, so all 200 tiles of the seed-42
    curriculum table -- gate poses, spawn origins, first gates -- equal what the reference's own functions produce."""
    pose, nxt, origins = _reference_complex_table()
    t = TG.generate_track_table(TG.racing_complex_cfg())
    assert np.array_equal(t.gate_pose, pose) and np.array_equal(t.next_gate_id, nxt) and np.array_equal(t.terrain_origins, origins)
    free = TG.generate_track_table(TG.racing_complex_cfg(add_obs=False))
    assert np.array_equal(free.gate_pose[0, 0], pose[0, 0]) and not np.array_equal(free.gate_pose[0, 1], pose[0, 1])   # the obstacle draws move every later tile


def test_complex_table_matches_the_golden_table():
    g = np.load(GOLDEN)
    t = TG.generate_track_table(TG.racing_complex_cfg())
    assert np.array_equal(g["complex_gate_pose"], t.gate_pose) and np.array_equal(g["complex_next_gate_id"], t.next_gate_id)
    assert np.array_equal(g["complex_terrain_origins"], t.terrain_origins)


def test_tile_entry_follows_the_generator():
    """terrain_generator.py:57-77: positions relative to the origin, wxyz quaternion from the YXZ / XYZ euler rule, origin recentred."""
    from scipy.spatial.transform import Rotation as R
    pose6, origin, _ = TG.square_track(0.5, TG.SquareTrackCfg(), TG.Streams(5))
    pose7, centred = TG.tile_entry(pose6, origin, (40.0, 40.0))
    q = (R.from_euler("YXZ", np.stack([pose6[:, 3], -pose6[:, 4], pose6[:, 5]], axis=1), degrees=True) * R.from_euler("XYZ", [-90, -90, 0], degrees=True)).as_quat()
    want = np.concatenate([q[:, 3:], q[:, :3]], axis=1)
    sign = np.sign(np.sum(want * pose7[:, 3:], axis=1, keepdims=True))           # q and -q are the same rotation
    assert np.allclose(pose7[:, 3:], want * sign, atol=1e-6)
    assert np.allclose(pose7[:, :3], pose6[:, :3] - origin)
    assert np.allclose(centred, origin + np.array([-20.0, -20.0, 0.0]))


def test_complex_table_layout_and_cache_round_trip(tmp_path):
    cfg = TG.racing_complex_cfg()
    t = TG.generate_track_table(cfg)
    assert t.gate_pose.shape == (20, 10, 8, 7) and t.next_gate_id.shape == (20, 10) and t.terrain_origins.shape == (10, 20, 3)
    # columns by cumulative proportion 0.3 / 0.3 / 0.4: 6 zigzag (next gate always 0), 6 rings, 8 ellipses
    assert (t.next_gate_id[:6] == 0).all()
    assert np.allclose(np.linalg.norm(t.gate_pose[..., 3:], axis=-1), 1.0, atol=1e-5)
    # every spawn origin sits a few metres from the gate to fly first, inside its own 40 m tile
    first = np.take_along_axis(t.gate_pose[..., :3], t.next_gate_id[..., None, None].astype(np.int64).repeat(3, -1), axis=2)[:, :, 0]
    d = np.linalg.norm(first[..., :2], axis=-1)
    assert (d > 0.3).all() and (d < 15.0).all()
    centres = np.stack(np.meshgrid((np.arange(10) + 0.5) * 40 - 200, (np.arange(20) + 0.5) * 40 - 400, indexing="ij"), axis=-1)
    assert (np.abs(t.terrain_origins[..., :2] - centres) < 21.0).all()      # (a 35 m zigzag + its 2-3 m run-up can poke 0.5 m out of a 40 m tile, as in the reference)
    # deterministic, and different global seeds give different tables
    assert np.array_equal(t.gate_pose, TG.generate_track_table(cfg).gate_pose)
    assert not np.array_equal(t.gate_pose, TG.generate_track_table(cfg, global_seed=43).gate_pose)
    # cache files of the reference's format (gate_info.yaml + origin.csv per tile)
    s = TG.Streams(42)
    dirs = []
    for c in range(2):
        col = []
        for r in range(3):
            pose6, origin, nid = TG.ellipse_track(0.1 * r, TG.EllipseTrackCfg(), s)
            pose7, centred = TG.tile_entry(pose6, origin, (40.0, 40.0))
            d_ = str(tmp_path / f"tile_{c}_{r}")
            TG.save_tile_cache(d_, pose7, nid, centred)
            p2, n2, o2 = TG.load_tile_cache(d_)
            assert np.allclose(p2, pose7) and n2 == nid and np.allclose(o2, centred)
            col.append(d_)
        dirs.append(col)
    tab = TG.table_from_cache(dirs, (40.0, 40.0))
    assert tab.gate_pose.shape == (2, 3, 8, 7)


def test_figure_eight_cfg_reproduces_the_pinned_c1_track():
    """RacingTestTerrainCfg (zero noise): the six gate positions / orientations of the C1 parity track, up to the random
    reversal and origin the reference draws."""
    from generalizableracing_b200.tracks import figure_eight_track
    pose6, origin, nid = TG.figure_eight_tile(0.0, TG.FigureEightTrackCfg(), TG.Streams(1))
    fixed = figure_eight_track(origin=(0.0, 0.0, 0.0))
    pts = fixed.gate_pose[0, 0, :, :3]
    assert nid == 0
    assert np.allclose(pose6[:, :3], pts) or np.allclose(pose6[:, :3], pts[::-1])
