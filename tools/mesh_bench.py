#!/usr/bin/env python
"""Timing of the terrain-mesh collision count at BASELINE C4 size (bench.py's `mesh_collision` extra alone)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from generalizableracing_b200.config import RacingCfg  # noqa: E402
from generalizableracing_b200.track_gen import generate_track_table, racing_complex_cfg  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    table = generate_track_table(racing_complex_cfg())
    print(json.dumps(bench.bench_mesh_collision(dev, RacingCfg.for_stage(1), table), indent=1))
