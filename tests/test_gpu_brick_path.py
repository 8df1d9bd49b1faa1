"""The brick path (csrc/brick.cuh: large triangles binned to 8^3-voxel bricks, no fragment list for them) against
the CPU oracle and against the fragment-sort path of the same library, through the C ABI.  svo_debug_set_build_path
forces one path or the other; the automatic choice is exercised by the full-size tests (C4 takes the brick path)."""
import numpy as np
import pytest

from sparsevoxeloctree_b200 import api, scenes
from tests.parity import check_against_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    L = api.get_library()
    if L.dll.svo_device_count() < 1:
        pytest.skip("no CUDA device")
    return L


@pytest.fixture()
def bricks(lib):
    lib.dll.svo_debug_set_build_path(1)
    yield lib
    lib.dll.svo_debug_set_build_path(-1)


@pytest.mark.parametrize("level", [5, 6, 7, 9, 10])
@pytest.mark.parametrize("mode", [api.CENTER, api.CONSERVATIVE_EXACT, api.CONSERVATIVE_DILATE])
def test_mixed_soup_on_the_brick_path(bricks, level, mode):
    # mixed triangle sizes: small ones through the fragment sort, large ones through the bricks, shared voxels folded
    # in list order (small class first) -- fragment multiset, topology and every leaf word against the oracle
    info = check_against_oracle(bricks, scenes.random_soup(400, 300 + level, 0.002, 1.2), level, mode)
    assert info["path"] == 1


@pytest.mark.parametrize("level,seed", [(11, 41), (12, 42)])
def test_wall_sized_triangles_on_the_brick_path(bricks, level, seed):
    rng = np.random.default_rng(seed)
    big = rng.uniform(-0.95, 0.95, (3 if level == 11 else 2, 3, 3))
    small = rng.uniform(-0.9, 0.9, (40, 1, 3)) + rng.uniform(-0.01, 0.01, (40, 3, 3))
    pos = np.concatenate([big, small]).reshape(-1, 3).astype(np.float32)
    idx = np.arange(len(pos), dtype=np.uint32)
    nb = 3 * len(big)
    draws = np.array([(0, nb, 0xFFFFFFFF, 0x00204060), (nb, len(idx) - nb, 0xFFFFFFFF, 0x00A0B0C0)], scenes.DRAW_DTYPE)
    info = check_against_oracle(bricks, scenes.Mesh(pos, idx, draws, f"big{level}"), level, api.CONSERVATIVE_EXACT)
    assert info["path"] == 1 and info["fragments"] > 500_000


def test_many_large_triangles_per_voxel(bricks):
    """Twenty large triangles through the same region: bricks with many pairs, voxels folded over many triangles."""
    rng = np.random.default_rng(5)
    c = rng.uniform(-0.05, 0.05, (20, 1, 3))
    pos = (c + rng.uniform(-0.7, 0.7, (20, 3, 3))).reshape(-1, 3).astype(np.float32)
    pos[:, 1] *= 0.01  # nearly coplanar: many voxels receive several fragments (4.8e4 fragments, 3.0e4 leaves)
    idx = np.arange(len(pos), dtype=np.uint32)
    draws = np.array([(0, 30, 0xFFFFFFFF, 0x00FF2010), (30, 30, 0xFFFFFFFF, 0x0010C0FF)], scenes.DRAW_DTYPE)
    info = check_against_oracle(bricks, scenes.Mesh(pos, idx, draws, "stack"), 8, api.CONSERVATIVE_EXACT)
    assert info["path"] == 1 and info["fragments"] > 1.5 * info["leaves"], info


@pytest.mark.parametrize("mode", [api.CENTER, api.CONSERVATIVE_EXACT])
def test_textured_and_alpha_tested_large_triangles(bricks, mode):
    m = scenes.textured_soup(200, 77, size_lo=0.01, size_hi=0.6, big_quads=True)
    info = check_against_oracle(bricks, m, 8, mode)
    assert info["path"] == 1


@pytest.mark.parametrize("cube", [(0, 0, 0), (1, 0, 1)])
def test_octant_shard_on_the_brick_path(bricks, cube):
    info = check_against_oracle(bricks, scenes.random_soup(300, 11, 0.01, 1.0), 7, api.CONSERVATIVE_EXACT, shard=(1, cube))
    assert info["path"] == 1


def test_unaligned_window_on_the_brick_path(bricks):
    # a voxel window whose corners are not multiples of the brick size: the bricks stay aligned to the grid
    info = check_against_oracle(bricks, scenes.random_soup(300, 12, 0.01, 1.0), 8, api.CONSERVATIVE_EXACT,
                                window=((13, 5, 27), (201, 250, 190)))
    assert info["path"] == 1


def test_both_paths_write_the_same_node_buffer(lib):
    """Not only canonically equal: the two paths produce the same words at the same places."""
    mesh = scenes.living_room_like(n_boxes=12, n_small=20000, level=10)
    out = {}
    for path in (0, 1):
        lib.dll.svo_debug_set_build_path(path)
        try:
            scene, vox, b = api.build_svo(mesh, 10, api.CONSERVATIVE_EXACT, lib=lib)
            assert b.BuildPath() == path
            out[path] = (b.octree_to_host(), b.GetLevelCounts(), vox.GetVoxelFragmentCount())
            b.Destroy(), vox.Destroy(), scene.Destroy()
        finally:
            lib.dll.svo_debug_set_build_path(-1)
    assert out[0][1] == out[1][1] and out[0][2] == out[1][2]
    assert (out[0][0] == out[1][0]).all()


def test_fragment_list_is_completed_on_request(bricks):
    """On the brick path CmdVoxelize emits only the small triangles' fragments; asking for the list completes it."""
    from tests.parity import oracle_fragment_keys
    mesh = scenes.living_room_like(n_boxes=4, n_small=500, level=9)
    scene = api.Scene.Create(mesh, lib=bricks)
    vox = api.Voxelizer.Create(scene, 9, api.CONSERVATIVE_EXACT)
    vox.CmdVoxelize()
    frags = vox.fragments_to_host()
    ref = vox.reference_fragments_to_host()
    okeys = oracle_fragment_keys(mesh, 9, api.CONSERVATIVE_EXACT)
    assert len(frags) == len(okeys) == len(ref) and (np.sort(frags) == np.sort(okeys)).all()
    vox.Destroy(), scene.Destroy()
