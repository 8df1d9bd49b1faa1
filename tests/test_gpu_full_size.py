"""Full-size checks at BASELINE.json's sizes, through size-independent properties (no oracle run needed):
sortedness + multiset checksums of the radix sort, leaf set == unique fragment voxels, layout invariants,
level-count consistency, range formula, run-to-run determinism."""
import numpy as np
import pytest

from sparsevoxeloctree_b200 import api, scenes
from tests.parity import check_against_oracle, check_layout_invariants

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return api.get_library()


def test_sort_100m_keys_properties(lib):
    n = 100_000_000
    rng = np.random.default_rng(1)
    k = rng.integers(0, 1 << 62, n, dtype=np.uint64)
    out = lib.sort_u64(k, 24, 60)
    key = out >> np.uint64(24) & np.uint64((1 << 36) - 1)
    assert (key[1:] >= key[:-1]).all()                                   # sorted on the key bits
    assert np.bitwise_xor.reduce(out) == np.bitwise_xor.reduce(k)        # same multiset (checksums)
    assert int(out.sum(dtype=np.uint64)) == int(k.sum(dtype=np.uint64))
    # stability: among equal keys the low (unsorted) bits keep their input order -> re-sorting is the identity
    again = lib.sort_u64(out, 24, 60)
    assert (again == out).all()


def _build(lib, name):
    cfg = scenes.CONFIGS[name]
    mesh = cfg["gen"]()
    mode = api.CENTER if cfg["mode"] == "center" else api.CONSERVATIVE_EXACT
    scene = api.Scene.Create(mesh, lib=lib)
    vox = api.Voxelizer.Create(scene, cfg["level"], mode)
    b = api.OctreeBuilder.Create(vox)
    vox.CmdVoxelize()
    frags = vox.fragments_to_host()
    b.CmdBuild()
    return cfg["level"], scene, vox, b, frags


def test_config2_sponza_scale_level10(lib):
    from oracle import oracle  # canonicaliser only (checker)
    level, scene, vox, b, frags = _build(lib, "C2")
    uniq = np.unique(frags >> np.uint64(24))
    counts = b.GetLevelCounts()
    assert counts[level] == len(uniq) == b.GetLeafCount()
    words = b.octree_to_host()
    assert b.GetOctreeRange() == 32 * (1 + sum(counts[1:level])) == len(words) * 4
    check_layout_invariants(words, level, counts)
    d, m, w = oracle.canonicalise(words, level)
    assert (np.sort(m[d == level]) == uniq).all()                        # occupancy bit-exact
    for dd in range(1, level):                                           # every ancestor set bit-exact
        assert (np.sort(m[d == dd]) == np.unique(uniq >> np.uint64(3 * (level - dd)))).all()
    # determinism: a second voxelize + build gives the identical buffer
    vox.CmdVoxelize()
    b.CmdBuild()
    assert (b.octree_to_host() == words).all()


def test_config4_living_room_scale_level12(lib):
    level, scene, vox, b, frags = _build(lib, "C4")
    assert len(frags) > 100_000_000
    uniq = np.unique(frags >> np.uint64(24))
    del frags
    counts = b.GetLevelCounts()
    assert counts[level] == len(uniq)
    for dd in range(level - 1, 0, -1):  # node count per depth == number of distinct Morton prefixes
        uniq = np.unique(uniq >> np.uint64(3))
        assert counts[dd] == len(uniq)
    words = b.octree_to_host()
    assert b.GetOctreeRange() == 32 * (1 + sum(counts[1:level])) == len(words) * 4
    check_layout_invariants(words, level, counts)


# ---- the BASELINE configs against the CPU oracle, whole (C2) or piecewise (C3, C4, C5) ------------------------------
# The oracle restates the reference's level loop (F*L pointer chases): the large configs are compared on voxel windows
# it finishes in seconds -- full depth, full leaf words (flags, fragment count and colour), fragment multisets included.
def _vertex_voxel(mesh, level, frac):
    """Voxel of the vertex at position `frac` of the vertex list (a point that is certainly on the surface)."""
    p = mesh.positions[int(frac * (len(mesh.positions) - 1))].astype(np.float64)
    return np.clip(((p + 1.0) * 0.5 * (1 << level)).astype(np.int64), 0, (1 << level) - 1)


def test_config2_whole_scene_against_oracle(lib):
    """BASELINE configs[1] (Sponza-scale, level 10, centre sampling): the whole scene, bit exact incl. colours."""
    cfg = scenes.CONFIGS["C2"]
    info = check_against_oracle(lib, cfg["gen"](), cfg["level"], api.CENTER)
    assert info["fragments"] > 4_000_000


def test_config3_window_against_oracle(lib):
    """BASELINE configs[2] (10 M small triangles, level 11, conservative): a 256 x 2048 x 256 column of the grid.  Several
    fragments per voxel: the leaf fold of the 'listed' branch of k_reduce_fused (runs dealt out one per thread)."""
    cfg = scenes.CONFIGS["C3"]
    mesh = cfg["gen"]()
    info = check_against_oracle(lib, mesh, cfg["level"], api.CONSERVATIVE_EXACT, window=((896, 0, 896), (1152, 2048, 1152)))
    assert info["fragments"] > 300_000
    assert info["fragments"] > 2 * info["leaves"]  # many fragments per voxel: the tile-wide run lists are exercised


@pytest.mark.parametrize("corner", [(0, 0, 0), (3584, 0, 3584)])
def test_config4_window_against_oracle(lib, corner):
    """BASELINE configs[3] (Living-Room-scale, level 12, conservative): a 512^3 corner of the room -- floor and two
    wall triangles of ~8.4 M pixels each through the large-triangle path, across 1024-pixel multiples of the Morton
    tables, plus furniture and clutter."""
    cfg = scenes.CONFIGS["C4"]
    lo = np.array(corner)
    info = check_against_oracle(lib, cfg["gen"](), cfg["level"], api.CONSERVATIVE_EXACT, window=(lo, lo + 512))
    assert info["fragments"] > 600_000


@pytest.mark.parametrize("frac", [0.31, 0.77])
def test_config5_level14_cube_against_oracle(lib, frac):
    """BASELINE configs[4] (dense surface, level 14): one 512^3 cube of the 16384^3 grid (shard level 5), built
    cube-local at full depth -- every leaf word, not only the flags."""
    cfg = scenes.CONFIGS["C5"]
    mesh = cfg["gen"]()
    cube = tuple(int(v) for v in _vertex_voxel(mesh, 14, frac) // 512)
    info = check_against_oracle(lib, mesh, 14, api.CENTER, shard=(5, cube))
    assert info["fragments"] > 100_000
