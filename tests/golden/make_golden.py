#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (oracle/svo_oracle.c).

The reference ships no fixtures for this path and cannot be executed in this image (no Vulkan ICD), so these
vectors are ORACLE outputs on fixed seeded inputs -- a regression pin for the oracle itself (CPU tests) and a
run-anywhere expected value for the CUDA path (GPU tests) -- not reference outputs.  Re-run after any
intentional change of the pinned arithmetic:   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from sparsevoxeloctree_b200 import scenes  # noqa: E402
from tests.parity import morton_np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "heightfield21_L6_center": (lambda: scenes.heightfield(21), 6, oracle.CENTER),
    "heightfield21_L6_conservative": (lambda: scenes.heightfield(21), 6, oracle.CONSERVATIVE_EXACT),
    "soup80_L7_conservative": (lambda: scenes.random_soup(80, 5, 0.01, 1.2), 7, oracle.CONSERVATIVE_EXACT),
    "soup200_L5_center": (lambda: scenes.random_soup(200, 6, 0.005, 0.6), 5, oracle.CENTER),
    "texsoup120_L7_conservative": (lambda: scenes.textured_soup(120, 12, size_hi=0.35), 7, oracle.CONSERVATIVE_EXACT),
    "texsoup60_L7_center": (lambda: scenes.textured_soup(60, 13, size_hi=0.35), 7, oracle.CENTER),
}


def make(name):
    gen, level, mode = CASES[name]
    mesh = gen()
    texset = oracle.TexSet(mesh.textures) if mesh.textures else None
    fr = oracle.voxelize(mesh.positions, mesh.indices, mesh.draws, level, mode,
                         texcoords=mesh.texcoords if texset is not None else None, texset=texset)
    keys = (morton_np(fr["x"], fr["y"], fr["z"], level) << np.uint64(24)) | fr["rgb"].astype(np.uint64)
    words, rng = oracle.build_octree(fr, level)
    d, m, w = oracle.canonicalise(words, level)
    # colours depend on the fragment order per voxel: store the flags+count byte only (order independent) and
    # the exact word where every contributing fragment has the same colour
    out = dict(level=level, mode=mode, n_fragments=len(fr), sorted_keys=np.sort(keys), range_bytes=rng,
               depth=d, morton=m, word_hi=(w >> np.uint32(24)).astype(np.uint8))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return out


if __name__ == "__main__":
    for n in CASES:
        o = make(n)
        print(n, o["n_fragments"], o["range_bytes"], len(o["depth"]))
