#!/usr/bin/env python
"""Kernel tuning harness (run on the GPU box): builds variants of libsvo_b200.so with different compile-time
switches and prints the per-phase cudaEvent times (and the time of every sort kernel) of a workload for each.

  python tools/tune.py --workload C4 --variants "SVO_OS_ITEMS=22;SVO_OS_ITEMS=18,SVO_OS_TMA=0,SVO_OS_MINB=3"

A variant is a comma-separated list of -D definitions; "base" = no definition; "git:<rev>" builds the sources of a
commit (the baseline to compare with).  All variants are compiled in parallel first.
"""
import argparse
import concurrent.futures as cf
import os
import subprocess
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
from sparsevoxeloctree_b200 import api, scenes  # noqa: E402


def build_variant(i, spec):
    out = f"/tmp/libsvo_var{i}.so"
    src = os.path.join(graft.CSRC, "svo_b200.cu")
    defs = []
    if spec.startswith("git:"):
        rev = spec[4:]
        d = os.path.join(ROOT, ".baseline_src", rev)  # exported beforehand (the GPU box has no .git):
        if not os.path.isdir(d):                      #   git archive <rev> sparsevoxeloctree_b200/csrc include | tar -x -C .baseline_src/<rev>
            os.makedirs(d)
            subprocess.run(f"git -C {ROOT} archive {rev} sparsevoxeloctree_b200/csrc include | tar -x -C {d}", shell=True, check=True)
        src = os.path.join(d, "sparsevoxeloctree_b200", "csrc", "svo_b200.cu")
    elif spec != "base":
        defs = [f"-D{x}" for x in spec.split(",") if x]
    cmd = ["nvcc"] + graft.NVCC_FLAGS + defs + ["-o", out, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return out if r.returncode == 0 else None, r.stderr[-400:]


def run(lib, mesh, level, mode, reps=6):
    scene = api.Scene.Create(mesh, lib=lib)
    vox = api.Voxelizer.Create(scene, level, mode)
    b = api.OctreeBuilder.Create(vox)
    prof = hasattr(lib.dll, "svo_debug_profile_passes")
    best = None
    for r in range(reps):
        if prof:
            lib.dll.svo_debug_profile_passes(1 if r == reps - 1 else 0)
        vox.CmdVoxelize()
        b.CmdBuild()
        ms, npass = b.LastMs()
        tot = sum(ms.values())
        if best is None or tot < best[0]:
            best = (tot, ms, npass)
    steps = b.SortStepMs() if prof else []
    if hasattr(lib.dll, "svo_debug_onesweep_clocks"):  # -DSVO_OS_CLOCKS=1 builds: cycles per phase and tile
        import ctypes as C
        clk = (C.c_ulonglong * 20)()
        lib.dll.svo_debug_onesweep_clocks(clk, 1)
        vox.CmdVoxelize()
        b.CmdBuild()
        lib.dll.svo_stream_synchronize(0, None)
        lib.dll.svo_debug_onesweep_clocks(clk, 1)
        tiles = -(-vox.GetVoxelFragmentCount() // 5632) * max(b.LastMs()[1], 1)
        names = ["load", "rank(w0)", "rank(all)", "digit scan", "reorder", "lookback(t0)", "lookback(all)", "scatter"]
        print("   cycles per tile: " + "  ".join(f"{n}={c / tiles:.0f}" for n, c in zip(names, clk)), flush=True)
        print(f"   thread 0 walks per tile: steps={clk[8] / tiles:.1f} states={clk[9] / tiles:.1f} empty polls={clk[10] / tiles:.1f} "
              f"walks={clk[11] / tiles:.2f}", flush=True)
    if prof:
        lib.dll.svo_debug_profile_passes(0)
    if hasattr(lib.dll, "svo_builder_build_path") and lib.dll.svo_builder_build_path(b._h) == 1:
        best = (best[0], {k2: v for k2, v in zip(("raster", "small_sort_reduce", "pairs", "bricks", "levels", "emit"), best[1].values())}, best[2])
    info = (vox.GetVoxelFragmentCount(), b.GetLeafCount(), zlib.crc32(b.octree_to_host().tobytes()))
    b.Destroy(), vox.Destroy(), scene.Destroy()
    return best, steps, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--variants", default="base")
    ap.add_argument("--reps", type=int, default=6)
    args = ap.parse_args()
    specs = [v for v in args.variants.split(";") if v]
    with cf.ThreadPoolExecutor(max_workers=min(len(specs), os.cpu_count() or 4)) as ex:
        built = list(ex.map(lambda t: build_variant(*t), enumerate(specs)))
    for wl in args.workload.split(","):
        cfg = scenes.CONFIGS[wl]
        mesh = cfg["gen"]()
        mode = api.CENTER if cfg["mode"] == "center" else api.CONSERVATIVE_EXACT
        for spec, (path, err) in zip(specs, built):
            if path is None:
                print(f"{wl} variant {spec}: does not compile: {err}", flush=True)
                continue
            try:
                if spec.startswith("git:"):  # an older ABI: bind only what this script calls
                    lib = api.Library.__new__(api.Library)
                    import ctypes as C
                    lib.path, lib.dll = path, C.CDLL(path)
                    for name, res, a in api.SYMBOLS:
                        if hasattr(lib.dll, name):
                            fn = getattr(lib.dll, name)
                            fn.restype, fn.argtypes = res, a
                else:
                    lib = api.Library(path)
                (tot, ms, npass), steps, (F, U, crc) = run(lib, mesh, cfg["level"], mode, args.reps)
            except Exception as e:  # noqa: BLE001
                print(f"{wl} variant {spec}: FAILED {e}", flush=True)
                continue
            per = ms.get("sort_passes", 0.0) / max(npass, 1)
            gbs = 16.0 * F / (per * 1e-3) / 1e9 if per > 0 else 0
            print(f"{wl} {spec:60s} total {tot:7.3f} ms | " + " ".join(f"{k}={x:.3f}" for k, x in ms.items()) +
                  f" | passes={npass} per-pass {per:.3f} ms = {gbs:.0f} GB/s | sort kernels: " +
                  " ".join(f"{x:.3f}" for x in steps) + f" | F={F} U={U} crc={crc:08x}", flush=True)


if __name__ == "__main__":
    main()
