"""Kernel-LOGIC tests on the CPU: the product's .cu sources compiled against tests/cpu_emu/cuda_emu.h
(one OS thread per CUDA thread) and driven through the same C ABI / api.py as on the GPU.
This is test infrastructure only -- the product never loads this library.  Sizes are tiny: the emulator
is slow; the real parity tests are tests/test_gpu_parity.py (-m gpu)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu_emu"))
import build_emu  # noqa: E402

from sparsevoxeloctree_b200 import api, scenes  # noqa: E402
from tests.parity import check_against_oracle, depth2_parts_check as _depth2_parts_check  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    L = api.Library(build_emu.build())
    assert b"emulation" in L.dll.svo_version()
    L.dll.svo_emu_set_lookback_aggregate_only(0)
    return L


@pytest.mark.parametrize("n", [0, 1, 33, 4096, 4097, 9000])
def test_onesweep_logic(emu, n):
    rng = np.random.default_rng(n)
    k = rng.integers(0, 1 << 62, n, dtype=np.uint64)
    out = emu.sort_u64(k, 24, 24 + 21)
    key = (k >> np.uint64(24)) & np.uint64((1 << 21) - 1)
    assert (out == k[np.argsort(key, kind="stable")]).all()


def test_onesweep_wide_lookback_words(emu):
    rng = np.random.default_rng(3)
    k = rng.integers(0, 1 << 62, 12000, dtype=np.uint64)
    emu.dll.svo_debug_force_wide_sort_state(1)
    try:
        out = emu.sort_u64(k, 24, 24 + 18)
    finally:
        emu.dll.svo_debug_force_wide_sort_state(0)
    key = (k >> np.uint64(24)) & np.uint64((1 << 18) - 1)
    assert (out == k[np.argsort(key, kind="stable")]).all()


def test_onesweep_long_lookback_chain(emu):
    # tiles publish aggregates only: every tile walks the full chain (the path sequential emulation never takes)
    emu.dll.svo_emu_set_lookback_aggregate_only(1)
    try:
        rng = np.random.default_rng(3)
        k = rng.integers(0, 1 << 40, 40000, dtype=np.uint64) << np.uint64(24)
        out = emu.sort_u64(k, 24, 24 + 16)
        key = (k >> np.uint64(24)) & np.uint64(0xFFFF)
        assert (out == k[np.argsort(key, kind="stable")]).all()
    finally:
        emu.dll.svo_emu_set_lookback_aggregate_only(0)


@pytest.mark.parametrize("mode", [api.CENTER, api.CONSERVATIVE_EXACT, api.CONSERVATIVE_DILATE])
def test_full_path_small_soup(emu, mode):
    check_against_oracle(emu, scenes.random_soup(250, 7), 6, mode)


def test_full_path_large_triangles_and_chained_scans(emu):
    # big triangles -> row-span path; aggregate-only look-back -> 32-wide window walk in every chained scan
    emu.dll.svo_emu_set_lookback_aggregate_only(1)
    try:
        info = check_against_oracle(emu, scenes.random_soup(120, 8, 0.2, 1.5), 7, api.CONSERVATIVE_EXACT)
        assert info["fragments"] > 30000
    finally:
        emu.dll.svo_emu_set_lookback_aggregate_only(0)


def test_full_path_shard(emu):
    check_against_oracle(emu, scenes.random_soup(150, 9, 0.05, 1.0), 6, api.CONSERVATIVE_EXACT, shard=(1, (1, 0, 1)))


def test_full_path_heightfield(emu):
    check_against_oracle(emu, scenes.heightfield(21), 6, api.CENTER)


@pytest.mark.parametrize("mode", [api.CENTER, api.CONSERVATIVE_EXACT])
def test_full_path_textured(emu, mode):
    """voxelizer.frag:27-36: trilinear sRGB sampling, alpha-test discard, packUnorm4x8 -- kernels vs oracle."""
    m = scenes.textured_soup(60, 11, size_hi=0.3)
    info = check_against_oracle(emu, m, 6, mode)
    assert info["fragments"] > 500


def test_raymarcher_and_builder_against_reference_tracer_golden(emu):
    """Kernel logic of the builder + the ray marcher on the CPU emulator vs the executed octree_tracer.frag."""
    from tests.test_spirv_golden import TRACER, cuda_tree_traced_like_reference
    cuda_tree_traced_like_reference(emu, TRACER[0])


@pytest.mark.parametrize("which", ["modeA", "modeB"])
def test_whole_path_against_reference_shaders_end_to_end(emu, which):
    from tests.test_spirv_golden import cuda_whole_path_like_reference
    cuda_whole_path_like_reference(emu, which)


def test_textured_path_against_reference_shaders_end_to_end(emu):
    from tests.test_spirv_golden import cuda_textured_path_like_reference
    cuda_textured_path_like_reference(emu)


@pytest.mark.parametrize("case", ["soup", "dilate", "textured", "shard", "stack", "room"])
def test_brick_path_logic(emu, case):
    """brick.cuh on the emulator: pair generation, pair sort, k_brick_raster (shared-memory grids, folds in triangle
    order, leaf blocks), the rank scans and k_brick_emit -- against the oracle, with the path forced."""
    emu.dll.svo_debug_set_build_path(1)
    try:
        if case == "soup":
            info = check_against_oracle(emu, scenes.random_soup(60, 6, 0.01, 1.5), 7, api.CONSERVATIVE_EXACT)
        elif case == "dilate":
            info = check_against_oracle(emu, scenes.random_soup(50, 7, 0.01, 1.2), 6, api.CONSERVATIVE_DILATE)
        elif case == "textured":
            info = check_against_oracle(emu, scenes.textured_soup(40, 11, size_hi=1.0), 6, api.CENTER)
        elif case == "shard":
            info = check_against_oracle(emu, scenes.random_soup(50, 3, 0.01, 1.5), 6, api.CONSERVATIVE_EXACT, shard=(1, (1, 0, 1)))
        elif case == "room":  # walls parallel to the grid: flat pairs, bricks written from their record alone
            info = check_against_oracle(emu, scenes.living_room_like(n_boxes=2, n_small=30, level=6), 6, api.CONSERVATIVE_EXACT)
            assert info["fragments"] > 15000
        else:  # several large triangles through the same voxels
            rng = np.random.default_rng(5)
            pos = (rng.uniform(-0.05, 0.05, (8, 1, 3)) + rng.uniform(-0.7, 0.7, (8, 3, 3))).reshape(-1, 3).astype(np.float32)
            pos[:, 1] *= 0.05
            idx = np.arange(len(pos), dtype=np.uint32)
            draws = np.array([(0, 12, 0xFFFFFFFF, 0x00FF2010), (12, 12, 0xFFFFFFFF, 0x0010C0FF)], scenes.DRAW_DTYPE)
            info = check_against_oracle(emu, scenes.Mesh(pos, idx, draws, "stack"), 6, api.CONSERVATIVE_EXACT)
            assert info["fragments"] > info["leaves"]
        assert info["path"] == 1
    finally:
        emu.dll.svo_debug_set_build_path(-1)


def test_empty_scene(emu):
    m = scenes.Mesh(np.zeros((0, 3), np.float32), np.zeros(0, np.uint32), np.zeros(0, scenes.DRAW_DTYPE), "empty")
    info = check_against_oracle(emu, m, 4, api.CENTER)
    assert info["range"] == 32 and info["fragments"] == 0


# ---- regression tests for the round-1 advisor findings ------------------------------------------------------
def _voxelize_only_matches_oracle(lib, mesh, level, mode):
    from tests.parity import oracle_fragment_keys
    scene = api.Scene.Create(mesh, lib=lib)
    vox = api.Voxelizer.Create(scene, level, mode)
    vox.CmdVoxelize()
    frags = vox.fragments_to_host()
    okeys = oracle_fragment_keys(mesh, level, mode)
    assert len(frags) == len(okeys) == vox.GetVoxelFragmentCount()
    assert (np.sort(frags) == np.sort(okeys)).all()
    vox.Destroy(), scene.Destroy()
    return len(frags)


def test_large_triangles_level8_two_dimensional_row_grid(emu):
    """Large triangles at level >= 8 launch k_large_rows / k_rows_compact with gridDim.y > 1 (rows of a triangle
    shared out in chunks among several warps): the blockIdx.y chunking against the oracle."""
    rng = np.random.default_rng(21)
    big = rng.uniform(-0.9, 0.9, (3, 3, 3))
    pos = big.reshape(-1, 3).astype(np.float32)
    idx = np.arange(len(pos), dtype=np.uint32)
    draws = np.array([(0, len(idx), 0xFFFFFFFF, 0x00306090)], scenes.DRAW_DTYPE)
    n = _voxelize_only_matches_oracle(emu, scenes.Mesh(pos, idx, draws, "big8"), 8, api.CONSERVATIVE_EXACT)
    assert n > 20_000


def test_more_than_2pow22_large_triangle_rows(emu):
    """The large-triangle row table used to be numbered by a scan of (count << 40 | sum) words whose look-back keeps
    62 bits, so the count wrapped at 2^22 rows.  ~100k thin large-class triangles of 49 rows each (4.9M rows)."""
    level, res = 12, 4096
    nx, ny, planes = 600, 83, 2
    gx, gy, gz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(planes), indexing="ij")
    x0 = (gx.ravel() * 6.5 + 20.25)
    y0 = (gy.ravel() * 49.0 + 10.25)
    z = (gz.ravel() * 700.0 + 1000.3)
    def ndc(p):
        return p / res * 2.0 - 1.0
    v0 = np.stack([ndc(x0), ndc(y0), ndc(z)], 1)
    v1 = np.stack([ndc(x0 + 5.4), ndc(y0 + 48.4), ndc(z)], 1)
    v2 = np.stack([ndc(x0 + 1.3), ndc(y0 + 0.2), ndc(z + 0.4)], 1)
    pos = np.stack([v0, v1, v2], 1).reshape(-1, 3).astype(np.float32)
    idx = np.arange(len(pos), dtype=np.uint32)
    draws = np.array([(0, len(idx), 0xFFFFFFFF, 0x00112233)], scenes.DRAW_DTYPE)
    mesh = scenes.Mesh(pos, idx, draws, "slivers")
    assert len(idx) // 3 * 49 > (1 << 22)
    n = _voxelize_only_matches_oracle(emu, mesh, level, api.CENTER)
    assert n > 1_000_000


def test_build_consumes_the_fragment_list(emu):
    """svo_builder_build sorts the fragment list in place: a second build or a fragment export without a new
    CmdVoxelize must fail with SVO_ERR_NOT_READY instead of building a wrong tree."""
    mesh = scenes.random_soup(250, 7)
    scene = api.Scene.Create(mesh, lib=emu)
    vox = api.Voxelizer.Create(scene, 6, api.CONSERVATIVE_EXACT)
    b = api.OctreeBuilder.Create(vox)
    vox.CmdVoxelize()
    b.CmdBuild()
    first = b.octree_to_host().copy()
    with pytest.raises(api.SvoError) as e:
        b.CmdBuild()
    assert e.value.code == -5  # SVO_ERR_NOT_READY
    with pytest.raises(api.SvoError):
        vox.reference_fragments_to_host()
    vox.CmdVoxelize()
    b.CmdBuild()
    assert (b.octree_to_host() == first).all()
    b.Destroy(), vox.Destroy(), scene.Destroy()


def test_index_out_of_range_is_rejected(emu):
    mesh = scenes.random_soup(20, 3)
    bad = mesh.indices.copy()
    bad[7] = len(mesh.positions)
    with pytest.raises(api.SvoError) as e:
        api.Scene.Create(mesh.positions, bad, mesh.draws, lib=emu)
    assert e.value.code == -1  # SVO_ERR_INVALID_ARGUMENT


@pytest.mark.parametrize("n_dev", [2, 8])
def test_build_sharded_c_abi_logic(emu, n_dev):
    """svo_build_sharded's host logic (slab windows, offsets, emit into one buffer, merged root) on the emulator."""
    from tests.parity import assert_same_tree
    mesh = scenes.random_soup(200, 17, 0.02, 0.9)
    sh = api.ShardedBuild.Create(mesh, 6, api.CONSERVATIVE_EXACT, devices=[0] * n_dev, lib=emu)
    _, vox, builder = api.build_svo(mesh, 6, api.CONSERVATIVE_EXACT, lib=emu)
    assert sh.GetOctreeRange() == builder.GetOctreeRange() and sh.GetLeafCount() == builder.GetLeafCount()
    assert_same_tree(sh.octree_to_host(), builder.octree_to_host(), 6)
    sh.Destroy()


def test_build_sharded_c_abi_compact_gather(emu):
    """svo_build_sharded with the slabs on the brick path: the slabs of devices 1.. cross in compact form."""
    from tests.parity import assert_same_tree
    mesh = scenes.living_room_like(n_boxes=2, n_small=30, level=6)
    emu.dll.svo_debug_set_build_path(1)
    try:
        sh = api.ShardedBuild.Create(mesh, 6, api.CONSERVATIVE_EXACT, devices=[0] * 4, lib=emu)
        _, vox, builder = api.build_svo(mesh, 6, api.CONSERVATIVE_EXACT, lib=emu)
        assert builder.BuildPath() == 1
    finally:
        emu.dll.svo_debug_set_build_path(-1)
    assert sh.GetOctreeRange() == builder.GetOctreeRange() and sh.GetLeafCount() == builder.GetLeafCount()
    assert_same_tree(sh.octree_to_host(), builder.octree_to_host(), 6)
    sh.Destroy()


@pytest.mark.parametrize("world,n_sub", [(2, 2), (8, 2), (4, 4)])
def test_depth2_parts_assemble_the_whole_tree(emu, world, n_sub):
    """Pipelined slab mode on one (emulated) device: every rank's slab cut into parts at depth-2 cell borders, each part
    built on its own and emitted with skip_root = 2 behind a 72-word header; the merged header + bodies = the whole tree."""
    _depth2_parts_check(emu, world, n_sub, scenes.random_soup(250, 23, 0.02, 0.9), 6)


@pytest.mark.parametrize("world", [2, 8])
def test_compact_gather_assembles_the_whole_tree(emu, world):
    """The compact gather of the slab mode (svo_builder_emit_compact_to + svo_expand_compact) on one emulated device:
    every slab built on the brick path, upper windows and rasterized bricks' leaf blocks sent to their places, 32 bytes
    per brick to a staging area, flat bricks and pointer blocks generated from the staged tables afterwards."""
    emu.dll.svo_debug_set_build_path(1)
    try:
        n = _depth2_parts_check(emu, world, 1, scenes.living_room_like(n_boxes=2, n_small=30, level=6), 6, compact=True)
        assert n >= world // 2
    finally:
        emu.dll.svo_debug_set_build_path(-1)


@pytest.mark.parametrize("level", [1, 2, 3, 4, 5])
def test_low_levels_through_the_tail_kernel(emu, level):
    """k_parent_tail walks every level above depth 4 in one single-block launch; at these levels it is the whole upper
    part of the build (level 1: the leaves are the root's children, no tail at all)."""
    check_against_oracle(emu, scenes.random_soup(40, 100 + level, 0.05, 1.2), level, api.CONSERVATIVE_EXACT)


@pytest.mark.parametrize("level", [5, 6])
def test_brick_path_at_its_lowest_levels(emu, level):
    """Level 5 is the lowest level with large triangles (a candidate rectangle of more than 256 pixels needs a grid of more
    than 16 x 16): k_brick_ranks writes depth 2 (the bricks), the tail kernel everything above."""
    emu.dll.svo_debug_set_build_path(1)
    try:
        info = check_against_oracle(emu, scenes.random_soup(30, 200 + level, 0.3, 1.5), level, api.CONSERVATIVE_EXACT)
        assert info["path"] == 1
    finally:
        emu.dll.svo_debug_set_build_path(-1)
