#!/usr/bin/env python
"""bench.py -- SVO build throughput on B200(s): `python bench.py --gpus N --steps K --warmup W`.

A step = one pass of the hot path over one synthetic scene: Voxelizer::CmdVoxelize + OctreeBuilder::CmdBuild,
the span the reference times with its 4 GPU timestamps (src/LoaderThread.cpp:57-97, "SVO build time" of the
README).  Workload = BASELINE.json configs[3]: the Living-Room-scale synthetic scene at level 12 (2^12 is
the resolution BASELINE.json's metric is quoted on; it fits one GPU).  metric = leaf voxels / s (and
ms_per_step = build ms).  N > 1: the same scene octant-sharded over N ranks with the NVLink subtree gather
(strong scaling; one process per GPU under torchrun).

`--impl reference` times the reference algorithm's CPU port (oracle/, all host threads) on a bounded sample
of the same workload; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC, UNIT = "svo_build_leaf_voxels_per_s", "leaf voxels/s"
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C4", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--level", type=int, default=0, help="override the workload's level")
    ap.add_argument("--mode", default="", choices=["", "center", "conservative"])
    ap.add_argument("--no-ipc", action="store_true", help="multi-GPU: NCCL send/recv instead of P2P stores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stitch-check", action="store_true", help="multi-GPU: skip the stitched-vs-single-GPU tree comparison")
    return ap.parse_args()


def workload(args):
    from sparsevoxeloctree_b200 import scenes
    cfg = scenes.CONFIGS[args.workload]
    level = args.level or cfg["level"]
    mode_name = args.mode or cfg["mode"]
    mesh = cfg["gen"]()
    return mesh, level, mode_name


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])), mx.append(float(parts[1])), pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v == "Active":
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def sample_box(level):
    """The bounded sample of the in-line cpu_baseline: one top-level octant of the grid at the full level."""
    half = (1 << level) // 2
    return ((0, 0, 0), (half, half, half))


def cpu_reference_run(mesh, level, mode_name, steps, warmup, budget_s=25.0, whole_scene=False, keep=None):
    """The reference algorithm's CPU port (oracle/, OpenMP over all host threads): voxelize + the literal level loop.
    whole_scene: the workload itself (the --impl reference arm); otherwise a bounded sample of it, the fragments of ONE
    top-level octant at the full level (so the tree is as deep as the real one).  Stops early once budget_s is spent
    (at least one timed run).  keep: dict that receives the first run's fragments and node words (parity check)."""
    from oracle import oracle
    cores = os.cpu_count() or 1
    mode = oracle.CENTER if mode_name == "center" else oracle.CONSERVATIVE_EXACT
    box = None if whole_scene or level > 12 else sample_box(level)
    times, leaves, frags_n = [], 0, 0
    t_start = time.perf_counter()
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        frags = oracle.voxelize(mesh.positions, mesh.indices, mesh.draws, level, mode, shard=box, nthreads=cores)
        words, rng = oracle.build_octree(frags, level, nthreads=cores)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if it == 0:
            d, _, _ = oracle.canonicalise(words, level)
            leaves, frags_n = int((d == level).sum()), len(frags)
            if keep is not None:
                keep["words"], keep["n_frags"] = words, len(frags)
        del frags, words
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    what = "the whole scene" if box is None else "octant (0,0,0) of the scene"
    sample = (f"{what} {mesh.name} at level {level} ({frags_n} fragments, {leaves} leaves), "
              f"oracle port (literal level loop), OpenMP {cores} threads, {len(times)} timed runs")
    return times, leaves, cores, sample


def parity_check_sample(lib, api, mesh, level, mode, mode_name, device, stream, keep):
    """The oracle tree of the cpu_baseline sample against a CUDA build of the same voxel window: fragment count,
    topology and occupancy bit exact; then the oracle's level loop is run once more on the CUDA fragment order (the
    reference's colour average depends on the order its atomics resolve in) and every leaf word must agree."""
    from oracle import oracle
    from tests.parity import keys_to_oracle_frags
    lo, hi = sample_box(level)
    scene = api.Scene.Create(mesh, device=device, stream=stream, lib=lib)
    vox = api.Voxelizer.CreateWindowed(scene, level, mode, lo, hi, stream=stream)
    b = api.OctreeBuilder.Create(vox, stream=stream)
    vox.CmdVoxelize(stream)
    frags = vox.fragments_to_host(stream)
    b.CmdBuild(stream)
    words = b.octree_to_host(stream)
    b.Destroy(), vox.Destroy(), scene.Destroy()
    out = {"window": [list(lo), list(hi)], "fragments": int(len(frags))}
    if len(frags) != keep["n_frags"]:
        out["result"] = f"FAILED: {len(frags)} fragments, oracle {keep['n_frags']}"
        return out
    d1, m1, w1 = oracle.canonicalise(words, level)
    d2, m2, w2 = oracle.canonicalise(keep["words"], level)
    keep.clear()
    if len(d1) != len(d2) or not ((d1 == d2).all() and (m1 == m2).all()):
        out["result"] = "FAILED: topology / occupancy differs from the oracle's tree"
        return out
    if not ((w1 >> 24) == (w2 >> 24)).all():
        out["result"] = "FAILED: leaf flags / fragment counts differ"
        return out
    # colours: a voxel with one fragment holds that colour; voxels with several are folded in list order (the reference's
    # running average depends on the order its atomics resolve in), so the oracle's level loop is run single-threaded on
    # exactly those fragments, in the CUDA list's order
    leaf = d1 == level
    lm, lw = m1[leaf], w1[leaf]
    mort = frags >> np.uint64(24)
    uniq, inv, cnt = np.unique(mort, return_inverse=True, return_counts=True)
    if len(uniq) != len(lm) or not (uniq == lm).all():
        out["result"] = "FAILED: the leaves are not the voxels of the fragment list"
        return out
    expected = np.zeros(len(uniq), np.uint32)
    single = cnt[inv] == 1
    expected[inv[single]] = np.uint32(0xC1000000) | (frags[single] & np.uint64(0xFFFFFF)).astype(np.uint32)
    multi = frags[~single]
    if len(multi):
        ow, _ = oracle.build_octree(keys_to_oracle_frags(multi, level), level, nthreads=1)
        d3, m3, w3 = oracle.canonicalise(ow, level)
        expected[cnt > 1] = w3[d3 == level]
    ok = bool((expected == lw).all())
    out.update(nodes=int(len(d1)), leaves=int(len(lm)), multi_fragment_leaves=int((cnt > 1).sum()),
               result="ok" if ok else "FAILED: leaf colours differ from the oracle's fold of the same fragment order")
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mesh, level, mode_name = workload(args)
    # the workload itself (same config as our arm): ~30 s per step for C4 on 16 cores, so one warm-up and as many of the K
    # steps as fit a 4-minute budget (the line's "steps" says how many were timed)
    warm = min(args.warmup, 1)
    times, leaves, cores, sample = cpu_reference_run(mesh, level, mode_name, max(1, args.steps), warm, budget_s=240.0, whole_scene=True)
    sec = float(np.mean(times))
    value = leaves / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{args.workload} {mesh.name} level {level} {mode_name}", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from sparsevoxeloctree_b200 import api, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = api.get_library()  # raises without the CUDA library: no fallback
    mesh, level, mode_name = workload(args)
    mode = api.CENTER if mode_name == "center" else api.CONSERVATIVE_EXACT
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # -------- device-resident arm: inputs already in HBM, handles created (count pass done) ----------------
    phases_acc = {k: 0.0 for k in api.PHASES}
    sort_passes = 0
    brick_acc, brick_counts, build_path = {"raster": 0.0, "scans": 0.0, "keys": 0.0, "emit": 0.0}, None, 0
    # one device holds levels <= 13 in a 64-bit fragment; level 14 runs as 8 cube-local octant builds ("virtual shards")
    single = world == 1 and level <= 13
    if single:
        scene = api.Scene.Create(mesh, device=local_rank, stream=stream, lib=lib)
        vox = api.Voxelizer.Create(scene, level, mode, stream=stream)
        builder = api.OctreeBuilder.Create(vox, stream=stream)

        def step():
            vox.CmdVoxelize(stream)
            builder.CmdBuild(stream)
    else:
        sh = sharded.ShardedSVO(torch if world > 1 else None, dist if world > 1 else None, mesh, level, mode, local_rank,
                                lib=lib, use_ipc=not args.no_ipc)

        def step():
            sh.step(stream)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = int(lib.dll.svo_launch_count())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
        bld = None
        if single:  # per-phase cudaEvent times recorded by the library on the same stream
            bld = builder
        elif rank == 0 and getattr(sh, "slab", False) and sh.builders and sh.builders[0].GetLeafCount():
            bld = sh.builders[0]  # rank 0's own slab (the roofline line below describes this rank)
        if bld is not None:
            ms, sort_passes = bld.LastMs()
            for k in api.PHASES:
                phases_acc[k] += ms[k]
            build_path = bld.BuildPath()
            if build_path == 1:
                brick_counts, bms = bld.BrickStats()
                for k in brick_acc:
                    brick_acc[k] += bms[k]
    e1.record(stream)
    barrier()
    launches = int(lib.dll.svo_launch_count()) - launches0
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps

    if single:
        leaves, frags = builder.GetLeafCount(), vox.GetVoxelFragmentCount()
        octree_bytes = builder.GetOctreeRange()
    else:
        c = torch.tensor([sh.leaf_count_local(), sh.fragment_count_local()], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(c)
        leaves, frags = int(c[0].item()), int(c[1].item())
        octree_bytes = sh.total_words * 4
    value = leaves / (ms_per_step * 1e-3)

    # -------- end-to-end arm: host (pinned) mesh -> handles -> build -> result read-back, every step --------
    pos_pin = torch.from_numpy(np.ascontiguousarray(mesh.positions)).pin_memory()
    idx_pin = torch.from_numpy(mesh.indices.astype(np.int32)).pin_memory()
    h2d = pos_pin.numel() * 4 + idx_pin.numel() * 4 + mesh.draws.nbytes
    d2h = 8 + 32 + 8 * (level + 1)

    def e2e_step():
        pm = mesh.__class__(pos_pin.numpy(), idx_pin.numpy().view(np.uint32), mesh.draws, mesh.name)
        if single:
            s = api.Scene.Create(pm, device=local_rank, stream=stream, lib=lib)
            v = api.Voxelizer.Create(s, level, mode, stream=stream)
            b = api.OctreeBuilder.Create(v, stream=stream)
            v.CmdVoxelize(stream)
            b.CmdBuild(stream)
            rng = b.GetOctreeRange()
            root = lib.to_host(b.GetOctree(), np.uint32, 8, local_rank, int(stream.cuda_stream))  # D2H + sync
            counts = b.GetLevelCounts()
            b.Destroy(), v.Destroy(), s.Destroy()
            return rng, root, counts
        s2 = sharded.ShardedSVO(torch if world > 1 else None, dist if world > 1 else None, pm, level, mode, local_rank, lib=lib,
                                use_ipc=not args.no_ipc)
        rng = s2.step(stream)
        root = lib.to_host(s2.final, np.uint32, 8, local_rank) if rank == 0 else None
        s2.destroy()
        return rng, root, None

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())

    # -------- N > 1: the stitched tree on rank 0 against a single-GPU build of the same scene (canonical compare) --------
    stitch = None
    if not single and world > 1 and not args.no_stitch_check:
        sh.step(stream)
        if rank == 0 and level <= 13:
            from oracle import oracle  # the canonicaliser (checker only)
            t0 = time.perf_counter()
            stitched = sh.octree_to_host()
            sc1 = api.Scene.Create(mesh, device=local_rank, stream=stream, lib=lib)
            v1 = api.Voxelizer.Create(sc1, level, mode, stream=stream)
            b1 = api.OctreeBuilder.Create(v1, stream=stream)
            v1.CmdVoxelize(stream)
            b1.CmdBuild(stream)
            whole = b1.octree_to_host(stream)
            b1.Destroy(), v1.Destroy(), sc1.Destroy()
            d1, m1, w1 = oracle.canonicalise(stitched, level)
            d2, m2, w2 = oracle.canonicalise(whole, level)
            same = len(d1) == len(d2) and (d1 == d2).all() and (m1 == m2).all() and (w1 == w2).all()
            stitch = {"result": "ok" if same else "FAILED: the stitched tree differs from the single-GPU tree",
                      "nodes": int(len(d1)), "words_stitched": int(len(stitched)), "words_single": int(len(whole)),
                      "seconds": round(time.perf_counter() - t0, 1)}
            del stitched, whole, d1, m1, w1, d2, m2, w2
        barrier()

    if rank == 0:
        peak, peak_src = hbm_peak()
        roofline, kernels = None, None
        frags_rank0 = frags if single else sh.vox[0].GetVoxelFragmentCount()
        traffic_db = {}
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                with open(tp) as f:
                    traffic_db = json.load(f)
            except Exception:
                traffic_db = {}
        if build_path == 1 and brick_acc["raster"] > 0:
            # Brick path: the large triangles never become fragments.  Its two big kernels (cudaEvents around each):
            #  k_brick_emit   HBM bound.  Algorithmic bytes: every 32-byte block of the two deepest windows written once
            #                 (leaf blocks: N1, pointer blocks: N2); in: per brick its record (16) and two ranks (16), and the
            #                 leaf blocks of the bricks that were rasterized (flat bricks -- one triangle, one depth voxel:
            #                 64 leaves in 16 identical blocks -- are generated from their record).
            #  k_brick_raster (+ k_brick_flat) instruction bound.  8-byte pairs and per-brick table entries (first pair 4,
            #                 code 8) in, record 16 out, 32-byte leaf blocks out for the rasterized bricks.
            bld0 = builder if single else sh.builders[0]
            lc = bld0.GetLevelCounts()
            kl = bld0.GetLevel()
            n1, n2 = lc[kl - 1], lc[kl - 2]
            flat = brick_counts["bricks"] - brick_counts["raster_bricks"]
            raster_blocks = n1 - 16 * flat
            emit_ms, raster_ms = brick_acc["emit"] / args.steps, brick_acc["raster"] / args.steps
            emit_bytes = 32.0 * (n1 + n2) + 32.0 * brick_counts["bricks"] + 32.0 * raster_blocks
            raster_bytes = 8.0 * brick_counts["pairs"] + 28.0 * brick_counts["bricks"] + 32.0 * raster_blocks
            t_emit, t_raster = traffic_db.get("k_brick_emit", {}), traffic_db.get("k_brick_raster", {})
            k_emit = {"bound": "hbm", "kernel": "k_brick_emit", "achieved": emit_bytes / (emit_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                      "frac": emit_bytes / (emit_ms * 1e-3) / 1e9 / peak, "traffic": t_emit.get("dram_bytes_per_launch") if single else None,
                      "peak_source": peak_src, "algorithmic_bytes_per_launch": emit_bytes, "launch_ms": emit_ms, "launches_per_step": 1,
                      "rank": 0}
            k_raster = {"bound": "hbm", "kernel": "k_brick_flat + k_brick_raster", "achieved": raster_bytes / (raster_ms * 1e-3) / 1e9,
                        "peak": peak, "unit": "GB/s", "frac": raster_bytes / (raster_ms * 1e-3) / 1e9 / peak,
                        "traffic": t_raster.get("dram_bytes_per_launch") if single else None, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": raster_bytes, "launch_ms": raster_ms, "launches_per_step": 1, "rank": 0,
                        "limiter": "instruction issue / per-warp latency, not HBM: 64 exact pixel tests (64-bit edge functions, fp64 depth) "
                                   "per (brick, triangle) pair that is not flat",
                        "issue_slots_busy_pct": t_raster.get("issue_slots_busy_pct")}
            for k in (k_emit, k_raster):
                k.update(pairs=brick_counts["pairs"], bricks=brick_counts["bricks"], flat_bricks=flat)
            roofline, other = (k_emit, k_raster) if emit_ms >= raster_ms else (k_raster, k_emit)  # the dominant kernel first
            kernels = [other]
        elif sort_passes and phases_acc["sort_passes"] > 0:
            per_launch_ms = phases_acc["sort_passes"] / args.steps / sort_passes
            # one read + one write of every 8-byte fragment per onesweep pass (N > 1: the fragments of rank 0's first part)
            alg_bytes = 16.0 * frags_rank0
            achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": "k_onesweep_pass", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic_db.get("k_onesweep_pass", {}).get("dram_bytes_per_launch") if single else None,
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": per_launch_ms,
                        "launches_per_step": sort_passes, "rank": 0}
        cpu_baseline, parity = None, None
        if single and not args.no_cpu_baseline:
            keep = {}
            times, cl, cores, sample = cpu_reference_run(mesh, level, mode_name, 3, 0, budget_s=25.0, keep=keep)
            cpu_baseline = {"value": cl / float(np.mean(times)), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
            if level <= 12:
                parity = parity_check_sample(lib, api, mesh, level, mode, mode_name, local_rank, stream, keep)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{args.workload} {mesh.name} level {level} {mode_name}", "level": level,
                       "triangles": mesh.n_triangles, "fragments": frags, "leaf_voxels": leaves,
                       "octree_bytes": octree_bytes, "raster_mode": mode_name,
                       "l2": "no flush needed: every step streams the fragment list (8 B x fragments, > 126 MB L2)",
                       "parallelism": "single GPU" if single else ("single GPU, 8 octant builds stitched (virtual shards)" if world == 1 else
                                      f"octant-sharded x{world}, NVLink subtree gather")
                                      + (f" ({'P2P stores via CUDA IPC' if not args.no_ipc else 'NCCL send/recv'})" if world > 1 else "")
                                      + ("; slabs on the brick path cross in compact form (upper windows, rasterized bricks' leaf blocks, "
                                         "32 B per brick) and GPU 0 expands the rest" if world > 1 and getattr(sh, "slab", False)
                                         and getattr(sh, "compact", False) and build_path == 1 else "")},
            "build_ms": ms_per_step,
            "build_path": "bricks" if build_path == 1 else "fragment sort",
            "phases_ms": ({(api.BRICK_PHASES[i] if build_path == 1 else k): phases_acc[k] / args.steps for i, k in enumerate(api.PHASES)}
                          if (sort_passes or build_path == 1) else None),
            "brick_ms": {k: v / args.steps for k, v in brick_acc.items()} if build_path == 1 else None,
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline, "parity_check": parity, "stitch_check": stitch,
            "e2e": {"value": leaves / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "note": "host wall clock around pinned-host mesh -> Scene/Voxelizer(count pass)/OctreeBuilder create -> "
                            "voxelize -> build -> read-back of range, root block and level counts -> destroy"},
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if not single:
        sh.destroy()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
