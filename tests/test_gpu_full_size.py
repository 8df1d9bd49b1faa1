"""Full-size checks at BASELINE.json's sizes, through size-independent properties (no oracle run needed):
sortedness + multiset checksums of the radix sort, leaf set == unique fragment voxels, layout invariants,
level-count consistency, range formula, run-to-run determinism."""
import numpy as np
import pytest

from sparsevoxeloctree_b200 import api, scenes
from tests.parity import check_layout_invariants

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
