"""Generates tests/golden/track_families.npz by running the UNMODIFIED reference family functions (mesh layer stubbed, see
oracle/ref_modules.load_track_families) on the cases of tests/test_track_gen.py.  Run in the build container:
    python tests/golden/make_track_golden.py"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_modules as RM  # noqa: E402
from tests.test_track_gen import CASES, _cfg, _fn, _ref_cfg, _reference_complex_table  # noqa: E402

ref = RM.load_track_families()
out = {}
for k, (fam, seed, diff) in enumerate(CASES):
    random.seed(seed)
    np.random.seed(seed)
    _, origin, ex = getattr(ref, _fn(fam))(diff, _ref_cfg(_cfg(fam)))
    out[f"pose_{k}"] = np.asarray(ex["gate_pose"])
    out[f"origin_{k}"] = np.asarray(origin, dtype=np.float64)
    out[f"next_{k}"] = np.int64(ex["next_gate_id"])
# the whole RacingComplexTerrainCfg table (seed 42, add_obs=True) from the reference's own functions
out["complex_gate_pose"], out["complex_next_gate_id"], out["complex_terrain_origins"] = _reference_complex_table()
np.savez(os.path.join(ROOT, "tests", "golden", "track_families.npz"), **out)
print("wrote", len(CASES), "cases")
