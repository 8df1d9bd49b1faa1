#!/usr/bin/env python
"""One C4 step (count pass, voxelize, build) through the default library: the process ncu wraps.
  ncu --set full --clock-control none --import-source on -k regex:'k_reduce|k_emit|k_radix|k_onesweep' -c 12 \
      -o gpurun_out/full python tools/profile_step.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsevoxeloctree_b200 import api, scenes  # noqa: E402

w = sys.argv[1] if len(sys.argv) > 1 else "C4"
cfg = scenes.CONFIGS[w]
mode = api.CENTER if cfg["mode"] == "center" else api.CONSERVATIVE_EXACT
scene, vox, b = api.build_svo(cfg["gen"](), cfg["level"], mode)
print(w, vox.GetVoxelFragmentCount(), b.GetLeafCount(), b.LastMs())
