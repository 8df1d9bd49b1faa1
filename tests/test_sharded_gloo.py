"""Multi-rank host logic of the octant-sharded build on CPU: world_size 2 and 4 over gloo.
Each rank builds its octants' subtrees with the oracle (standing in for the CUDA builder), then runs the
product's exchange plan (sizes all_gather -> offsets -> rebase -> gather -> root block) and rank 0 checks the
stitched tree against the whole-grid oracle tree."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle
from sparsevoxeloctree_b200 import scenes, sharded


def _worker(rank, world, port, level, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = scenes.random_soup(300, 21, 0.02, 1.0)
        mode = oracle.CONSERVATIVE_EXACT
        half = 1 << (level - 1)
        octs = sharded.octants_of_rank(rank, world)
        subtrees, sizes = {}, []
        for o in octs:
            cx, cy, cz = sharded.octant_cube(o)
            lo = (cx * half, cy * half, cz * half)
            hi = tuple(v + half for v in lo)
            fr = oracle.voxelize(mesh.positions, mesh.indices, mesh.draws, level, mode, shard=(lo, hi))
            fr["x"] -= lo[0]
            fr["y"] -= lo[1]
            fr["z"] -= lo[2]
            if len(fr):
                w, _ = oracle.build_octree(fr, level - 1)
            else:
                w = np.zeros(0, np.uint32)
            subtrees[o] = w
            sizes.append(len(w))
        words = sharded.exchange_sizes(torch, dist, sizes, world, torch.device("cpu"))
        bases, total = sharded.plan_offsets(words)
        if rank == 0:
            final = np.zeros(total, np.uint32)
            for o in range(8):
                if not words[o]:
                    continue
                if o % world == 0:
                    final[bases[o]:bases[o] + words[o]] = sharded.rebase_words_numpy(subtrees[o], bases[o])
                else:
                    buf = torch.zeros(words[o], dtype=torch.int32)
                    dist.recv(buf, o % world)
                    final[bases[o]:bases[o] + words[o]] = buf.numpy().view(np.uint32)
            final[:8] = sharded.root_block(bases, words)
            fr = oracle.voxelize(mesh.positions, mesh.indices, mesh.draws, level, mode)
            ow, orng = oracle.build_octree(fr, level)
            d1, m1, w1 = oracle.canonicalise(final, level)
            d2, m2, w2 = oracle.canonicalise(ow, level)
            ok = len(d1) == len(d2) and (d1 == d2).all() and (m1 == m2).all() and ((w1 >> 24) == (w2 >> 24)).all()
            # single-material voxels: colours exact too
            q.put(("ok" if ok and total * 4 == orng else "mismatch", total * 4, orng))
        else:
            for o in octs:
                if words[o]:
                    w = sharded.rebase_words_numpy(subtrees[o], bases[o])
                    dist.send(torch.from_numpy(w.view(np.int32).copy()), 0)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_stitch_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + world + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, world, port, 6, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == "ok", res


def test_plan_and_root_block():
    assert sharded.octants_of_rank(1, 2) == [1, 3, 5, 7]   # x split
    assert sharded.octants_of_rank(2, 4) == [2, 6]         # x,y split
    assert sharded.octants_of_rank(5, 8) == [5]
    bases, total = sharded.plan_offsets([16, 0, 24, 0, 0, 0, 0, 8])
    assert bases == [8, 0, 24, 0, 0, 0, 0, 48] and total == 56
    rb = sharded.root_block(bases, [16, 0, 24, 0, 0, 0, 0, 8])
    assert rb.tolist() == [0x80000008, 0, 0x80000018, 0, 0, 0, 0, 0x80000030]
    with pytest.raises(OverflowError):
        sharded.plan_offsets([1 << 29] * 2 + [0] * 6)
    w = np.array([0x80000008, 0xC1000005, 0, 0x80000010], np.uint32)
    assert sharded.rebase_words_numpy(w, 64).tolist() == [0x80000048, 0xC1000005, 0, 0x80000050]


def test_sub_windows_and_header_merge():
    # the slab of a rank cut at depth-2 cell borders: 2 parts along y, 4 along y and z
    assert sharded.sub_windows(1, 2, 5, 2) == [([16, 0, 0], [32, 16, 32]), ([16, 16, 0], [32, 32, 32])]
    boxes = sharded.sub_windows(5, 8, 6, 4)   # octant (1, 0, 1) of a 64^3 grid
    assert boxes == [([32, 0, 32], [64, 16, 48]), ([32, 0, 48], [64, 16, 64]), ([32, 16, 32], [64, 32, 48]), ([32, 16, 48], [64, 32, 64])]
    for world in (2, 4, 8):
        for n_sub in (1, 2, 4):
            vol = sum(int(np.prod(np.array(hi) - np.array(lo))) for r in range(world) for lo, hi in sharded.sub_windows(r, world, 7, n_sub))
            assert vol == 128 ** 3  # the parts tile the grid
    with pytest.raises(ValueError):
        sharded.sub_windows(0, 8, 6, 3)       # 1, 2 or 4 parts
    # two parts of octant 3 (different depth-2 cells) and one of octant 5, built separately
    a = np.zeros((2, 8), np.uint32); a[0, 3] = 0x80000008; a[1, 0] = 0x80000100; a[1, 2] = 0x80000108
    b = np.zeros((2, 8), np.uint32); b[0, 3] = 0x80000008; b[1, 5] = 0x80000200
    c = np.zeros((2, 8), np.uint32); c[0, 5] = 0x80000008; c[1, 7] = 0x80000300
    h0, h1 = sharded.merge_top_blocks([a, b]), sharded.merge_top_blocks([c])
    hdr = sharded.merge_headers(np.stack([h0, h1]))
    assert hdr[3] == 0x80000000 | 32 and hdr[5] == 0x80000000 | 48 and not hdr[[0, 1, 2, 4, 6, 7]].any()
    assert hdr[32:40].tolist() == [0x80000100, 0, 0x80000108, 0, 0, 0x80000200, 0, 0]
    assert hdr[48:56].tolist() == [0, 0, 0, 0, 0, 0, 0, 0x80000300]
    with pytest.raises(ValueError):
        sharded.merge_top_blocks([a, a])      # the same child slot from two parts
