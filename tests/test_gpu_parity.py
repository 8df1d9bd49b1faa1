"""Parity tests proper: the sm_100a CUDA path, called through the C ABI, against the CPU oracle.
Bit-exact bar (integer / index work): fragment multisets, octree range, canonicalised node words."""
import numpy as np
import pytest

from sparsevoxeloctree_b200 import api, scenes
from tests.parity import check_against_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    L = api.get_library()  # raises if libsvo_b200.so is missing: no fallback
    assert L.dll.svo_device_count() >= 1, "no CUDA device"
    assert b"sm_100a" in L.dll.svo_version()
    return L


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 4095, 4096, 4097, 100_000, 3_000_001])
@pytest.mark.parametrize("bits", [(24, 48), (24, 60), (0, 64), (24, 29)])
def test_onesweep_sort_matches_stable_numpy(lib, n, bits):
    rng = np.random.default_rng(n + bits[1])
    k = rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
    out = lib.sort_u64(k, *bits)
    width = bits[1] - bits[0]
    key = (k >> np.uint64(bits[0])) & np.uint64((1 << width) - 1 if width < 64 else 0xFFFFFFFFFFFFFFFF)
    assert (out == k[np.argsort(key, kind="stable")]).all()


def test_onesweep_sort_clustered_keys(lib):
    # spatially coherent keys (few distinct upper digits, long runs): the warp-uniform histogram path and
    # long equal-digit runs in the ranking
    rng = np.random.default_rng(5)
    n = 1_000_003
    k = (np.repeat(rng.integers(0, 1 << 36, n // 1000 + 1, dtype=np.uint64), 1000)[:n] << np.uint64(24)) | \
        rng.integers(0, 1 << 24, n, dtype=np.uint64)
    out = lib.sort_u64(k, 24, 60)
    assert (out == k[np.argsort(k >> np.uint64(24), kind="stable")]).all()


def test_onesweep_sort_wide_lookback_words(lib):
    """The 64-bit look-back words the sort switches to at >= 2^30 keys, forced on at a testable size."""
    rng = np.random.default_rng(6)
    k = rng.integers(0, 1 << 63, 2_000_003, dtype=np.uint64)
    lib.dll.svo_debug_force_wide_sort_state(1)
    try:
        out = lib.sort_u64(k, 24, 60)
    finally:
        lib.dll.svo_debug_force_wide_sort_state(0)
    key = (k >> np.uint64(24)) & np.uint64((1 << 36) - 1)
    assert (out == k[np.argsort(key, kind="stable")]).all()


@pytest.mark.parametrize("mode", [api.CENTER, api.CONSERVATIVE_EXACT, api.CONSERVATIVE_DILATE])
def test_config1_heightfield_level8(lib, mode):
    # BASELINE.json configs[0]: 10k-triangle heightfield, level 8
    info = check_against_oracle(lib, scenes.heightfield(), 8, mode)
    assert info["fragments"] > 50_000


@pytest.mark.parametrize("level", [1, 2, 3, 5, 9, 10])
@pytest.mark.parametrize("mode", [api.CENTER, api.CONSERVATIVE_EXACT, api.CONSERVATIVE_DILATE])
def test_random_soup_levels(lib, level, mode):
    # mixed triangle sizes: exercises both work classes (single-thread walk and row spans)
    check_against_oracle(lib, scenes.random_soup(400, 100 + level, 0.002, 1.2), level, mode)


def test_large_triangles_level9(lib):
    # huge triangles only: the row-span / output-parallel path at a high level
    m = scenes.living_room_like(n_boxes=3, n_small=50)
    info = check_against_oracle(lib, m, 9, api.CONSERVATIVE_EXACT)
    assert info["fragments"] > 1_000_000


@pytest.mark.parametrize("level,seed", [(11, 41), (12, 42)])
def test_big_triangles_above_level10(lib, level, seed):
    """Levels above 10 take the two-table Morton path (coordinate bits 10.. are spread separately) and rows that cross
    multiples of 1024 pixels; a few wall-sized triangles in general position plus some small ones, against the oracle."""
    rng = np.random.default_rng(seed)
    big = rng.uniform(-0.95, 0.95, (3 if level == 11 else 2, 3, 3))
    small_c = rng.uniform(-0.9, 0.9, (40, 1, 3))
    small = small_c + rng.uniform(-0.01, 0.01, (40, 3, 3))
    pos = np.concatenate([big, small]).reshape(-1, 3).astype(np.float32)
    idx = np.arange(len(pos), dtype=np.uint32)
    nb = 3 * len(big)
    draws = np.array([(0, nb, 0xFFFFFFFF, 0x00204060), (nb, len(idx) - nb, 0xFFFFFFFF, 0x00A0B0C0)], scenes.DRAW_DTYPE)
    info = check_against_oracle(lib, scenes.Mesh(pos, idx, draws, f"big{level}"), level, api.CONSERVATIVE_EXACT)
    assert info["fragments"] > 500_000


@pytest.mark.parametrize("case", range(16))
def test_random_configurations(lib, case):
    """Seeded fuzz over scene size, triangle sizes, level, raster mode, sharding and texturing."""
    rng = np.random.default_rng(1000 + case)
    level = int(rng.integers(3, 11))
    mode = [api.CENTER, api.CONSERVATIVE_EXACT, api.CONSERVATIVE_DILATE][int(rng.integers(0, 3))]
    n = int(rng.integers(20, 600))
    lo = float(np.exp(rng.uniform(np.log(0.002), np.log(0.05))))
    hi = float(lo * np.exp(rng.uniform(np.log(2), np.log(60))))
    if rng.random() < 0.4:
        mesh = scenes.textured_soup(n, 2000 + case, size_lo=lo, size_hi=min(hi, 0.6), big_quads=bool(rng.integers(0, 2)))
    else:
        mesh = scenes.random_soup(n, 2000 + case, lo, min(hi, 1.5), n_mat=int(rng.integers(1, 5)))
    shard = None
    if level >= 4 and rng.random() < 0.35:
        sl = int(rng.integers(1, 3))
        shard = (sl, tuple(int(v) for v in rng.integers(0, 1 << sl, 3)))
    check_against_oracle(lib, mesh, level, mode, shard=shard)


@pytest.mark.parametrize("cube", [(0, 0, 0), (1, 0, 1), (1, 1, 1)])
def test_octant_shard(lib, cube):
    # one top-level octant in cube-local coordinates
    check_against_oracle(lib, scenes.random_soup(300, 11, 0.01, 1.0), 7, api.CONSERVATIVE_EXACT, shard=(1, cube))


def test_mode_b_matches_mode_a_up_to_ties(lib):
    # the reference's two conservative variants describe the same pixel set mathematically:
    # software dilation + centre sampling (Mode B) vs exact square/triangle overlap (Mode A) differ only in
    # fp rounding and exact-touch ties -- a small fraction of voxels, and never a large hole
    m = scenes.heightfield()
    leaves = {}
    for mode in (api.CONSERVATIVE_EXACT, api.CONSERVATIVE_DILATE):
        scene, vox, b = api.build_svo(m, 8, mode, lib=lib)
        leaves[mode] = np.unique(vox.fragments_to_host() >> np.uint64(24)) if False else None
        vox.CmdVoxelize()
        leaves[mode] = np.unique(vox.fragments_to_host() >> np.uint64(24))
    a, b_ = leaves[api.CONSERVATIVE_EXACT], leaves[api.CONSERVATIVE_DILATE]
    only_a, only_b = np.setdiff1d(a, b_), np.setdiff1d(b_, a)
    assert len(only_b) <= 0.002 * len(a)          # dilation never finds voxels the exact test misses (up to rounding)
    assert len(only_a) <= 0.02 * len(a)           # exact-touch ties the centre sample excludes


def test_empty_and_degenerate_scenes(lib):
    # no triangles at all
    m = scenes.Mesh(np.zeros((0, 3), np.float32), np.zeros(0, np.uint32), np.zeros(0, scenes.DRAW_DTYPE), "empty")
    info = check_against_oracle(lib, m, 5, api.CENTER)
    assert info["fragments"] == 0 and info["range"] == 32
    # zero-area triangles: nothing in centre mode, segments / points in conservative mode
    pos = np.array([[-0.5, 0.1, 0.2], [0.0, 0.1, 0.2], [0.5, 0.1, 0.2], [0.3, 0.3, 0.3]], np.float32)
    idx = np.array([0, 1, 2, 3, 3, 3], np.uint32)
    draws = np.array([(0, 6, 0xFFFFFFFF, 0x00FF00)], scenes.DRAW_DTYPE)
    m = scenes.Mesh(pos, idx, draws, "degenerate")
    assert check_against_oracle(lib, m, 6, api.CENTER)["fragments"] == 0
    assert check_against_oracle(lib, m, 6, api.CONSERVATIVE_EXACT)["fragments"] > 0


def test_colour_average_many_fragments_per_voxel(lib):
    # 200 coincident small triangles of alternating materials in one voxel: count saturates at 63 and the
    # running average is order dependent (octree_tag_node.comp:48-57)
    base = np.array([[0.101, 0.101, 0.101], [0.104, 0.101, 0.101], [0.101, 0.104, 0.101]], np.float32)
    pos = np.tile(base, (200, 1))
    idx = np.arange(600, dtype=np.uint32)
    draws = np.array([(0, 300, 0xFFFFFFFF, 0x0000FF), (300, 300, 0xFFFFFFFF, 0xFF0000)], scenes.DRAW_DTYPE)
    info = check_against_oracle(lib, scenes.Mesh(pos, idx, draws, "pile"), 4, api.CONSERVATIVE_EXACT)
    assert info["leaves"] >= 1


def test_reference_fragment_packing(lib):
    # GetVoxelFragmentList in the reference's uvec2 packing (voxelizer.frag:40-42)
    from oracle import oracle
    m = scenes.heightfield(31)
    scene = api.Scene.Create(m, lib=lib)
    vox = api.Voxelizer.Create(scene, 7, api.CENTER)
    vox.CmdVoxelize()
    packed = vox.reference_fragments_to_host()
    ofr = oracle.voxelize(m.positions, m.indices, m.draws, 7, oracle.CENTER)
    exp = np.array([oracle.pack_fragment(int(f["x"]), int(f["y"]), int(f["z"]), int(f["rgb"])) for f in ofr], dtype=np.uint32)
    a = packed[np.lexsort((packed[:, 1], packed[:, 0]))]
    b = exp[np.lexsort((exp[:, 1], exp[:, 0]))]
    assert (a == b).all()


def test_stride20_reference_vertex_layout(lib):
    # the reference's interleaved Vertex {vec3 pos; vec2 uv} (Scene.cpp:16-19), stride 20
    m = scenes.heightfield(31)
    v5 = np.zeros((len(m.positions), 5), np.float32)
    v5[:, :3] = m.positions
    v5[:, 3:] = 0.25
    a = check_against_oracle(lib, scenes.Mesh(v5, m.indices, m.draws, "stride20"), 7, api.CENTER)
    b = check_against_oracle(lib, m, 7, api.CENTER)
    assert a == b


def test_error_behaviour(lib):
    m = scenes.heightfield(11)
    scene = api.Scene.Create(m, lib=lib)
    with pytest.raises(api.SvoError):
        api.Voxelizer.Create(scene, 0)  # level below kOctreeLevelMin (Config.hpp:18)
    with pytest.raises(api.SvoError):
        api.Voxelizer.Create(scene, 15)
    vox = api.Voxelizer.Create(scene, 5)
    b = api.OctreeBuilder.Create(vox)
    with pytest.raises(api.SvoError):
        b.CmdBuild()  # fragments not emitted yet
    assert b.GetOctreeRange() == 0 and b.GetOctree() == 0
    draws = m.draws.copy()
    draws["texture_id"][0] = 3
    with pytest.raises(api.SvoError):
        api.Scene.Create(m.positions, m.indices, draws, lib=lib)  # a textured draw without its texture / texcoords


# ---- textured materials (voxelizer.frag:27-36) ----------------------------------------------------------
def test_mip_chain_matches_oracle(lib):
    """Scene::load_textures' linear-blit mip chain (Scene.cpp:290-295), incl. non-power-of-two sizes."""
    from oracle import oracle
    rng = np.random.default_rng(3)
    tex = scenes.procedural_textures(5) + [rng.integers(0, 256, (h, w, 4), dtype=np.uint8) for h, w in ((1, 7), (37, 3), (128, 96))]
    m = scenes.textured_soup(8, 3, big_quads=False)
    m.textures = tex
    scene = api.Scene.Create(m, lib=lib)
    ts = oracle.TexSet(tex)
    for t in range(len(tex)):
        n = scene.texture_level_count(t)
        assert n == ts.level_count(t)
        for lv in range(n):
            assert (scene.texture_level_to_host(t, lv) == ts.level(t, lv)).all(), (t, lv)
    scene.Destroy()


@pytest.mark.parametrize("level", [6, 9, 11])
@pytest.mark.parametrize("mode", [api.CENTER, api.CONSERVATIVE_EXACT, api.CONSERVATIVE_DILATE])
def test_textured_soup(lib, level, mode):
    """trilinear sRGB sampling at every LOD regime, alpha-test discards (count pass included), packUnorm4x8, on both
    raster paths -- fragment multiset (colours included) and tree bit-exact against the oracle."""
    info = check_against_oracle(lib, scenes.textured_soup(400, 20 + level, size_hi=0.3), level, mode)
    assert info["fragments"] > 1000


def test_wall_sized_alpha_tested_quad(lib):
    """A quad across the whole grid with an alpha-tested texture: the large path counts and emits its rows by sampling
    (k_large_rows / k_emit_alpha_rows); about half of its pixels are discarded."""
    import time
    tex = scenes.procedural_textures(4)
    pos = np.array([[-0.93, -0.9, 0.1], [0.95, -0.92, 0.3], [0.94, 0.91, 0.35], [-0.9, 0.93, 0.05]], np.float32)
    uv = np.array([[0, 0], [7.3, 0.2], [7.1, 6.4], [-0.2, 6.1]], np.float32)
    idx = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    draws = np.array([(0, 6, 1, 0x00808080)], scenes.DRAW_DTYPE)  # texture 1: the one with alpha holes
    mesh = scenes.Mesh(pos, idx, draws, "alpha_wall", texcoords=uv, textures=tex)
    t = time.time()
    info = check_against_oracle(lib, mesh, 10, api.CONSERVATIVE_EXACT)
    assert 200_000 < info["fragments"] < 900_000  # ~0.9 M covered pixels, a good part of them discarded
    assert time.time() - t < 60


def test_textured_octant_shard(lib):
    check_against_oracle(lib, scenes.textured_soup(300, 31, size_hi=0.4), 8, api.CONSERVATIVE_EXACT, shard=(1, (0, 1, 1)))


def test_textured_reference_vertex_layout(lib):
    # interleaved Vertex {vec3 pos; vec2 uv}, stride 20: texcoords taken from positions + 12 (Scene.cpp:16-19)
    m = scenes.textured_soup(100, 32)
    v5 = np.concatenate([m.positions, m.texcoords], axis=1).astype(np.float32)
    a = check_against_oracle(lib, scenes.Mesh(v5, m.indices, m.draws, "tex20", texcoords=m.texcoords, textures=m.textures), 7,
                             api.CENTER)
    scene = api.Scene.Create(v5, m.indices, m.draws, lib=lib, textures=m.textures)  # no separate texcoords array
    vox = api.Voxelizer.Create(scene, 7, api.CENTER)
    assert vox.GetVoxelFragmentCount() == a["fragments"]


def test_virtual_shards_stitch_equals_whole_grid(lib):
    # all 8 octants built as cube-local subtrees on one GPU, rebased + stitched by k_rebase_copy:
    # canonically identical to the whole-grid build (bit exact incl. colours)
    from oracle import oracle
    from sparsevoxeloctree_b200 import sharded
    from tests.parity import assert_same_tree
    mesh = scenes.random_soup(500, 31, 0.01, 1.2)
    level, mode = 8, api.CONSERVATIVE_EXACT
    sh = sharded.ShardedSVO(None, None, mesh, level, mode, 0, lib=lib)
    nbytes = sh.step()
    stitched = sh.octree_to_host()
    assert nbytes == len(stitched) * 4
    scene, vox, builder = api.build_svo(mesh, level, mode, lib=lib)
    whole = builder.octree_to_host()
    assert_same_tree(stitched, whole, level)
    assert sh.leaf_count_local() == builder.GetLeafCount()
    assert sh.fragment_count_local() == vox.GetVoxelFragmentCount()
    sh.destroy()


def test_two_rank_nccl_stitch(lib):
    # real multi-process path (one process per GPU, NCCL + CUDA IPC) when the box has >= 2 GPUs
    import subprocess, sys, os
    if lib.dll.svo_device_count() < 2:
        pytest.skip("needs 2 GPUs (covered on CPU by tests/test_sharded_gloo.py and on 1 GPU by the virtual-shard test)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # (--room with the brick path forced: the slabs cross in compact form, svo_builder_emit_compact_to + svo_expand_compact)
    for extra, env in (([], {}), (["--no-ipc"], {}), (["--room"], {"SVO_BUILD_PATH": "1"})):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
               "--master-port", "29611", os.path.join(root, "tests", "multi_gpu_check.py")] + extra
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, **env})
        assert r.returncode == 0 and "STITCH_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_level14_virtual_shards_against_oracle(lib):
    # 16384^3: 42 Morton bits + 24 colour bits exceed a 64-bit fragment, so the grid is built as 8 cube-local
    # level-13 octants stitched under one root; checked against the oracle's whole-grid level-14 tree
    from oracle import oracle
    from sparsevoxeloctree_b200 import sharded
    from tests.parity import assert_same_tree
    mesh = scenes.random_soup(300, 77, 0.0005, 0.01)
    level, mode = 14, api.CONSERVATIVE_EXACT
    sh = sharded.ShardedSVO(None, None, mesh, level, mode, 0, lib=lib)
    sh.step()
    stitched = sh.octree_to_host()
    fr = oracle.voxelize(mesh.positions, mesh.indices, mesh.draws, level, oracle.CONSERVATIVE_EXACT)
    assert len(fr) == sh.fragment_count_local()
    ow, orng = oracle.build_octree(fr, level, cap_words=8 * (1 + len(fr) * (level - 1)))
    d1, m1, w1 = oracle.canonicalise(stitched, level)
    d2, m2, w2 = oracle.canonicalise(ow, level)
    assert (d1 == d2).all() and (m1 == m2).all()          # occupancy and topology bit-exact
    assert ((w1 >> 24) == (w2 >> 24)).all()               # flags and fragment counts exact
    sh.destroy()
    with pytest.raises(api.SvoError):                      # a single 64-bit fragment cannot hold level 14
        api.Voxelizer.Create(api.Scene.Create(mesh, lib=lib), 14)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_split_emit_to_assembles_the_whole_tree(lib, world):
    # the multi-GPU slab path on ONE device: every "rank" builds the window of its octants in global coordinates
    # (svo_voxelizer_create_windowed), prepares, and emits its node words with final pointers straight into a shared
    # arena (svo_builder_emit_to, skip_root); the merged root + bodies must equal the whole-grid tree canonically
    from sparsevoxeloctree_b200 import sharded
    from tests.parity import assert_same_tree
    mesh = scenes.random_soup(600, 51, 0.01, 1.2)
    level, mode = 8, api.CONSERVATIVE_EXACT
    scene = api.Scene.Create(mesh, lib=lib)
    parts, bodies = [], []
    for r in range(world):
        lo, hi = sharded.slab_window(r, world, level)
        v = api.Voxelizer.CreateWindowed(scene, level, mode, lo, hi)
        b = api.OctreeBuilder.Create(v)
        v.CmdVoxelize()
        b.Prepare()
        parts.append((v, b))
        bodies.append(b.GetOctreeRange() // 4 - 8 if b.GetLeafCount() else 0)
    total = 8 + sum(bodies)
    arena = lib.malloc(total * 4)
    root = np.zeros(8, np.uint64)
    for r, (v, b) in enumerate(parts):
        base = 8 + sum(bodies[:r])
        if bodies[r]:
            b.EmitTo(arena + base * 4, base, True)
            root += b.RootWords().astype(np.uint64)
    rb = root.astype(np.uint32)
    lib.check(lib.dll.svo_memcpy_h2d(0, arena, rb.ctypes.data, 32, 0))
    stitched = lib.to_host(arena, np.uint32, total)
    _, vox, builder = api.build_svo(mesh, level, mode, lib=lib)
    assert_same_tree(stitched, builder.octree_to_host(), level)
    assert total * 4 == builder.GetOctreeRange()   # same node count: the split adds no blocks
    assert sum(v.GetVoxelFragmentCount() for v, _ in parts) == vox.GetVoxelFragmentCount()
    lib.free(arena)


def test_export_fd_roundtrip(lib):
    """Row f1 (src/Octree.cpp:22-35 takes the builder's buffer): the node words in exportable memory, handed over as a
    file descriptor the way VK_KHR_external_memory_fd takes it, re-imported and compared with the builder's own buffer.
    Both a prepared builder (the emit kernel writes the exported memory directly) and a built one (device copy)."""
    import ctypes as C
    import os
    from tests.parity import assert_same_tree
    mesh = scenes.random_soup(500, 91, 0.01, 1.0)
    level = 8
    scene, vox, built = api.build_svo(mesh, level, api.CONSERVATIVE_EXACT, lib=lib)
    ref_words = built.octree_to_host()
    for prepared_only in (True, False):
        v = api.Voxelizer.Create(scene, level, api.CONSERVATIVE_EXACT)
        b = api.OctreeBuilder.Create(v)
        v.CmdVoxelize()
        if prepared_only:
            b.Prepare()
        else:
            b.CmdBuild()
        fd, size, d_ptr = b.ExportFd()
        assert fd >= 0 and size >= b.GetOctreeRange() and size % 4096 == 0 and d_ptr
        assert os.fstat(fd).st_mode  # a live descriptor
        own = lib.to_host(d_ptr, np.uint32, b.GetOctreeRange() // 4)
        assert (own == ref_words).all()
        handle, mapped = C.c_void_p(), C.c_void_p()
        path = lib.dll.svo_external_memory_import_fd(0, fd, size, C.byref(handle), C.byref(mapped))
        assert path in (1, 2), lib.dll.svo_last_error()
        imported = lib.to_host(mapped.value, np.uint32, size // 4)
        assert (imported[: len(ref_words)] == ref_words).all() and not imported[len(ref_words):].any()
        assert_same_tree(imported[: len(ref_words)], ref_words, level)
        lib.check(lib.dll.svo_external_memory_release(0, handle))
        print(f"export_fd: {size} bytes, imported through path {path} (1 = cudaImportExternalMemory, 2 = cuMemImportFromShareableHandle)")
        b.Destroy(), v.Destroy()


@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
def test_build_sharded_c_abi_on_one_device(lib, n_dev):
    """svo_build_sharded (the multi-GPU build behind the C ABI) with every 'device' = GPU 0: the slab split, the emit into
    one stitched buffer and the merged root must give the single-build tree, canonically, colours included."""
    from tests.parity import assert_same_tree
    mesh = scenes.random_soup(600, 52, 0.01, 1.2)
    level, mode = 8, api.CONSERVATIVE_EXACT
    sh = api.ShardedBuild.Create(mesh, level, mode, devices=[0] * n_dev, lib=lib)
    _, vox, builder = api.build_svo(mesh, level, mode, lib=lib)
    assert sh.GetOctreeRange() == builder.GetOctreeRange()
    assert sh.GetLeafCount() == builder.GetLeafCount() and sh.GetVoxelFragmentCount() == vox.GetVoxelFragmentCount()
    assert_same_tree(sh.octree_to_host(), builder.octree_to_host(), level)
    sh.Rebuild()  # same buffers again
    assert_same_tree(sh.octree_to_host(), builder.octree_to_host(), level)
    sh.Destroy()


def test_build_sharded_c_abi_compact_gather_gpu(lib):
    """svo_build_sharded on a scene of wall-sized triangles, brick path forced: the slabs of 'devices' 1.. cross in compact
    form (svo_builder_emit_compact_to) and are completed on devices[0] (svo_expand_compact)."""
    from tests.parity import assert_same_tree
    mesh = scenes.living_room_like(n_boxes=8, n_small=2000, level=9)
    level, mode = 9, api.CONSERVATIVE_EXACT
    lib.dll.svo_debug_set_build_path(1)
    try:
        sh = api.ShardedBuild.Create(mesh, level, mode, devices=[0] * 8, lib=lib)
        _, vox, builder = api.build_svo(mesh, level, mode, lib=lib)
        assert builder.BuildPath() == 1
        assert sh.GetOctreeRange() == builder.GetOctreeRange() and sh.GetLeafCount() == builder.GetLeafCount()
        assert_same_tree(sh.octree_to_host(), builder.octree_to_host(), level)
        sh.Rebuild()
        assert_same_tree(sh.octree_to_host(), builder.octree_to_host(), level)
        sh.Destroy()
    finally:
        lib.dll.svo_debug_set_build_path(-1)


def test_build_sharded_c_abi_level14(lib):
    """Level 14 through svo_build_sharded: 8 cube-local level-13 octant builds, rebased into one buffer."""
    from oracle import oracle
    mesh = scenes.random_soup(300, 77, 0.0005, 0.01)
    sh = api.ShardedBuild.Create(mesh, 14, api.CONSERVATIVE_EXACT, devices=[0, 0], lib=lib)
    fr = oracle.voxelize(mesh.positions, mesh.indices, mesh.draws, 14, oracle.CONSERVATIVE_EXACT)
    assert len(fr) == sh.GetVoxelFragmentCount()
    ow, _ = oracle.build_octree(fr, 14, cap_words=8 * (1 + len(fr) * 13))
    d1, m1, w1 = oracle.canonicalise(sh.octree_to_host(), 14)
    d2, m2, w2 = oracle.canonicalise(ow, 14)
    assert (d1 == d2).all() and (m1 == m2).all() and ((w1 >> 24) == (w2 >> 24)).all()
    sh.Destroy()


@pytest.mark.parametrize("world,n_sub", [(2, 2), (4, 2), (8, 2), (8, 4)])
def test_depth2_parts_assemble_the_whole_tree_gpu(lib, world, n_sub):
    """The pipelined multi-GPU layout on ONE device: each rank's slab as n_sub parts cut at depth-2 cell borders, emitted
    with skip_root = 2 behind the 72-word header; merged header + bodies must equal the whole-grid tree."""
    from tests.parity import depth2_parts_check
    depth2_parts_check(lib, world, n_sub, scenes.random_soup(700, 53, 0.01, 1.2), 9)


@pytest.mark.parametrize("world,n_sub", [(2, 1), (8, 1), (4, 2)])
def test_compact_gather_assembles_the_whole_tree_gpu(lib, world, n_sub):
    """The compact gather (svo_builder_emit_compact_to + svo_expand_compact) on ONE device, at a size where most bricks
    are flat: slabs built on the brick path send their upper windows, the rasterized bricks' leaf blocks and 32 bytes per
    brick; the flat bricks' blocks and all pointer blocks are generated from the staged tables.  Must equal the
    whole-grid tree word for word after canonicalisation."""
    from tests.parity import depth2_parts_check
    lib.dll.svo_debug_set_build_path(1)
    try:
        n = depth2_parts_check(lib, world, n_sub, scenes.living_room_like(n_boxes=8, n_small=2000, level=9), 9, compact=True)
        assert n >= world * n_sub // 2
    finally:
        lib.dll.svo_debug_set_build_path(-1)
