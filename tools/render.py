#!/usr/bin/env python
"""Render a built octree with the CUDA port of the reference's primary-ray traversal (Octree_RayMarchLeaf,
shader/octree.glsl:179-340; views of octree_tracer.frag:43-47) -- verification without Vulkan.
  python tools/render.py --workload C2 --size 512 --out gpurun_out/c2.png
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsevoxeloctree_b200 import api, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--view", default="shaded", choices=["shaded", "diffuse", "normal", "position", "iteration"])
    ap.add_argument("--out", default="gpurun_out/render.png")
    args = ap.parse_args()
    cfg = scenes.CONFIGS[args.workload]
    mesh = cfg["gen"]()
    mode = api.CENTER if cfg["mode"] == "center" else api.CONSERVATIVE_EXACT
    level = min(cfg["level"], 13)
    _, _, builder = api.build_svo(mesh, level, mode)
    # a camera inside the [1,2]^3 cube in the middle of the scene
    o, d = api.camera_rays([1.5, 1.3, 1.5], [0.8, 0.1, 0.6], [0.48, 0.0, -0.64], [0.0, 0.8, 0.0], args.size, args.size)
    t = time.time()
    hits = api.raymarch_leaf(builder.GetOctree(), o, d)
    dt = time.time() - t
    hit = hits["hit"] != 0
    if args.view == "shaded":
        light = np.array([0.35, 0.8, 0.5]) / np.linalg.norm([0.35, 0.8, 0.5])
        img = np.where(hit[:, None], np.power(hits["colour"], 1 / 2.2) * (0.6 + 0.4 * (hits["normal"] @ light))[:, None], 0.1)
    elif args.view == "diffuse":
        img = np.where(hit[:, None], np.power(hits["colour"], 1 / 2.2), 0.1)
    elif args.view == "normal":
        img = np.where(hit[:, None], hits["normal"] * 0.5 + 0.5, 0.5)
    elif args.view == "position":
        img = np.where(hit[:, None], hits["pos"] - 1.0, 0.0)
    else:
        x = np.clip(hits["iter"][:, None] / 128.0, 0, 1)
        img = np.sin(x * 3.0 - np.array([1.0, 2.0, 3.0])) * 0.5 + 0.5
    img = (np.clip(img, 0, 1) * 255).astype(np.uint8).reshape(args.size, args.size, 3)
    from PIL import Image
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    Image.fromarray(img).save(args.out)
    print(f"{args.workload} L={level}: {hit.sum()} / {len(hit)} rays hit, mean {hits['iter'].mean():.1f} iterations, "
          f"{dt * 1e3:.1f} ms incl. copies -> {args.out}")


if __name__ == "__main__":
    main()
