"""Shared parity checker: the CUDA path (through the C ABI / api.py) against the CPU oracle."""
from __future__ import annotations

import numpy as np

from oracle import oracle
from sparsevoxeloctree_b200 import api


def morton_np(x, y, z, level):
    m = np.zeros(len(x), dtype=np.uint64)
    x, y, z = (np.asarray(v).astype(np.uint64) for v in (x, y, z))
    for b in range(level):
        m |= ((x >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
        m |= ((y >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + 1)
        m |= ((z >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + 2)
    return m


def demorton_np(m, level):
    out = []
    for s in range(3):
        c = np.zeros(len(m), dtype=np.uint32)
        for b in range(level):
            c |= ((m >> np.uint64(3 * b + s)) & np.uint64(1)).astype(np.uint32) << np.uint32(b)
        out.append(c)
    return out


def oracle_fragment_keys(mesh, level, mode, shard_box=None, origin=(0, 0, 0), key_level=None):
    """The oracle's fragments as 64-bit keys (morton << 24 | rgb), in oracle emission order."""
    texset = oracle.TexSet(mesh.textures) if getattr(mesh, "textures", None) else None
    fr = oracle.voxelize(mesh.positions, mesh.indices, mesh.draws, level, mode, shard=shard_box,
                         texcoords=mesh.texcoords if texset is not None else None, texset=texset)
    kl = level if key_level is None else key_level
    m = morton_np(fr["x"] - np.uint32(origin[0]), fr["y"] - np.uint32(origin[1]), fr["z"] - np.uint32(origin[2]), kl)
    return (m << np.uint64(24)) | fr["rgb"].astype(np.uint64)


def keys_to_oracle_frags(keys, level):
    x, y, z = demorton_np(keys >> np.uint64(24), level)
    return oracle.frags_from_xyzc(x, y, z, (keys & np.uint64(0xFFFFFF)).astype(np.uint32))


def assert_same_tree(words_a, words_b, level):
    da, ma, wa = oracle.canonicalise(words_a, level)
    db, mb, wb = oracle.canonicalise(words_b, level)
    assert len(da) == len(db), f"node count {len(da)} != {len(db)}"
    assert (da == db).all() and (ma == mb).all(), "topology / occupancy differs"
    assert (wa == wb).all(), "leaf words differ"
    return da, ma, wa


def check_layout_invariants(words, level, level_counts):
    """The reference layout: root block at 0, level windows top-down, pointers inside the next window."""
    words = np.asarray(words)
    assert len(words) % 8 == 0
    internal = (words & np.uint32(0xC0000000)) == np.uint32(0x80000000)
    leaf = (words & np.uint32(0xC0000000)) == np.uint32(0xC0000000)
    assert ((words == 0) | internal | leaf).all()
    ptr = words[internal] & np.uint32(0x3FFFFFFF)
    assert (ptr % 8 == 0).all() and (ptr > 0).all() and (ptr.astype(np.int64) + 8 <= len(words)).all()
    # every block except the root is pointed to exactly once
    assert len(np.unique(ptr)) == len(ptr) == len(words) // 8 - 1
    blocks = 1 + sum(level_counts[1:level])
    assert len(words) == 8 * blocks
    assert int(leaf.sum()) == level_counts[level]


def check_against_oracle(lib, mesh, level, mode, shard=None, device=0, stream=None, window=None):
    """Full-path parity: (1) fragment multiset == oracle voxelizer's; (2) octree built by CUDA ==
    oracle level loop fed with the same fragment order, after Morton canonicalisation (bit exact, colours
    included); (3) range and layout invariants.
    shard = (shard_level, cube index): one cube of the grid in cube-local coordinates (level - shard_level deep);
    window = (lo, hi): a half-open voxel box of the whole grid in global coordinates (full level deep) -- how the
    BASELINE configs that are too large for the CPU oracle are checked piecewise."""
    scene = api.Scene.Create(mesh, device=device, stream=stream, lib=lib)
    if window is not None:
        vox = api.Voxelizer.CreateWindowed(scene, level, mode, window[0], window[1], stream=stream)
    else:
        vox = api.Voxelizer.Create(scene, level, mode, shard=shard, stream=stream)
    builder = api.OctreeBuilder.Create(vox, stream=stream)
    vox.CmdVoxelize(stream)
    frags = vox.fragments_to_host(stream)
    key_level = builder.GetLevel()
    box, origin = None, (0, 0, 0)
    if window is not None:
        box = (tuple(int(v) for v in window[0]), tuple(int(v) for v in window[1]))
    if shard is not None:
        sl, ci = shard
        side = 1 << (level - sl)
        origin = tuple(c * side for c in ci)
        box = (origin, tuple(o + side for o in origin))
    okeys = oracle_fragment_keys(mesh, level, mode, box, origin, key_level)
    assert len(frags) == len(okeys) == vox.GetVoxelFragmentCount(), (len(frags), len(okeys))
    assert (np.sort(frags) == np.sort(okeys)).all(), "fragment multiset differs from the oracle"
    builder.CmdBuild(stream)
    words = builder.octree_to_host(stream)
    rng = builder.GetOctreeRange()
    ow, orng = oracle.build_octree(keys_to_oracle_frags(frags, key_level), key_level)
    assert rng == orng == len(words) * 4, (rng, orng)
    d, m, w = assert_same_tree(words, ow, key_level)
    counts = builder.GetLevelCounts()
    check_layout_invariants(words, key_level, counts)
    assert counts[key_level] == builder.GetLeafCount() == int((d == key_level).sum())
    info = dict(fragments=len(frags), leaves=counts[key_level], range=rng, counts=counts,
                path=int(lib.dll.svo_builder_build_path(builder._h)))
    builder.Destroy(), vox.Destroy(), scene.Destroy()
    return info


def depth2_parts_check(lib, world, n_sub, mesh, level, mode=api.CONSERVATIVE_EXACT, compact=False):
    """Builds the scene as world * n_sub separately built parts (sharded.sub_windows) assembled in one arena the way the
    pipelined multi-GPU path does, and compares with the whole-grid build (canonical, colours included).
    compact: parts that took the brick path go through the compact gather (EmitCompactTo into a staging area, then
    expand_compact once every part has been sent), as the slab mode does between GPUs; returns how many parts did."""
    from sparsevoxeloctree_b200 import sharded
    scene = api.Scene.Create(mesh, lib=lib)
    parts, bodies, tops = [], [], []
    for r in range(world):
        for lo, hi in sharded.sub_windows(r, world, level, n_sub):
            v = api.Voxelizer.CreateWindowed(scene, level, mode, lo, hi)
            b = api.OctreeBuilder.Create(v)
            v.CmdVoxelize()
            b.Prepare()
            counts = b.GetLevelCounts() if b.GetLeafCount() else None
            parts.append((v, b))
            bodies.append(b.GetOctreeRange() // 4 - 8 * (1 + counts[1]) if counts else 0)
    total = sharded.HEADER_WORDS + sum(bodies)
    arena = lib.malloc(total * 4)
    stage_bytes = [(b.CompactBytes() + 255) // 256 * 256 if compact else 0 for _, b in parts]
    stage = lib.malloc(max(sum(stage_bytes), 256))
    pending = []
    for k, (v, b) in enumerate(parts):
        base = sharded.HEADER_WORDS + sum(bodies[:k])
        if b.GetLeafCount():
            if stage_bytes[k]:
                tables = stage + sum(stage_bytes[:k])
                if k % 2:  # the two-call form the slab mode uses: tables first, the other stores later
                    plan = b.PushTables(base, 2, tables)
                    assert b.EmitCompactTo(arena + base * 4, base, 2, None) is None
                else:
                    plan = b.EmitCompactTo(arena + base * 4, base, 2, tables)
                pending.append((tables, plan, arena + base * 4))
            else:
                b.EmitTo(arena + base * 4, base, 2)
            tops.append(b.TopWords())
    for tables, plan, dst in pending:
        api.expand_compact(lib, 0, tables, plan, dst)
    header = sharded.merge_top_blocks(tops)
    lib.check(lib.dll.svo_memcpy_h2d(0, arena, header.ctypes.data, header.nbytes, 0))
    lib.check(lib.dll.svo_stream_synchronize(0, 0))
    stitched = lib.to_host(arena, np.uint32, total)
    _, vox, builder = api.build_svo(mesh, level, mode, lib=lib)
    assert_same_tree(stitched, builder.octree_to_host(), level)
    assert sum(v.GetVoxelFragmentCount() for v, _ in parts) == vox.GetVoxelFragmentCount()
    lib.free(arena)
    lib.free(stage)
    for v, b in parts:
        b.Destroy(), v.Destroy()
    return len(pending)
