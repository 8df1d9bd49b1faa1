#!/usr/bin/env python
"""Kernel tuning harness (run on the GPU box): builds variants of libsvo_b200.so with different compile-time
tile shapes and prints the per-phase cudaEvent times of the C4 (or other) workload for each.

  python tools/tune.py --workload C4 --variants "512,8,2;256,16,3;256,16,4;512,8,3;384,12,2"
"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
from sparsevoxeloctree_b200 import api, scenes  # noqa: E402


def build_variant(tag, defs):
    out = f"/tmp/libsvo_{tag}.so"
    cmd = ["nvcc"] + graft.NVCC_FLAGS + [f"-D{d}" for d in defs] + ["-o", out, os.path.join(graft.CSRC, "svo_b200.cu")]
    subprocess.run(cmd, check=True, capture_output=True)
    return out


def run(lib, mesh, level, mode, reps=6):
    scene = api.Scene.Create(mesh, lib=lib)
    vox = api.Voxelizer.Create(scene, level, mode)
    b = api.OctreeBuilder.Create(vox)
    best = None
    for _ in range(reps):
        vox.CmdVoxelize()
        b.CmdBuild()
        ms, npass = b.LastMs()
        tot = sum(ms.values())
        if best is None or tot < best[0]:
            best = (tot, ms, npass)
    import zlib
    info = (vox.GetVoxelFragmentCount(), b.GetLeafCount(), zlib.crc32(b.octree_to_host().tobytes()))
    b.Destroy(), vox.Destroy(), scene.Destroy()
    return best, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--variants", default="256,22,3")
    ap.add_argument("--extra", default="", help="extra -D definitions, comma separated")
    args = ap.parse_args()
    cfg = scenes.CONFIGS[args.workload]
    mesh = cfg["gen"]()
    mode = api.CENTER if cfg["mode"] == "center" else api.CONSERVATIVE_EXACT
    for v in args.variants.split(";"):
        blk, items, minb = v.split(",")
        defs = [f"SVO_OS_BLOCK={blk}", f"SVO_OS_ITEMS={items}", f"SVO_OS_MINB={minb}"] + [d for d in args.extra.split(",") if d]
        try:
            path = build_variant(f"{blk}_{items}_{minb}_" + "_".join(args.extra.replace("=", "").split(",")), defs)
        except subprocess.CalledProcessError:
            print(f"variant {v}: does not compile", flush=True)
            continue
        lib = api.Library(path)
        (tot, ms, npass), (F, U, crc) = run(lib, mesh, cfg["level"], mode)
        per = ms["sort_passes"] / max(npass, 1)
        gbs = 16.0 * F / (per * 1e-3) / 1e9 if per > 0 else 0
        print(f"variant {v:12s} total {tot:7.3f} ms | " + " ".join(f"{k}={x:.3f}" for k, x in ms.items()) +
              f" | passes={npass} per-pass {per:.3f} ms = {gbs:.0f} GB/s  (F={F} U={U} crc={crc:08x})", flush=True)


if __name__ == "__main__":
    main()
