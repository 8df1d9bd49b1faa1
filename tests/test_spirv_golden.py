"""The reference's own shader binaries, executed (tests/golden/make_spirv_golden.py + oracle/spirv_interp.py),
versus the CPU oracle (CPU tests) and versus the CUDA builder (GPU tests).

This is what pins the oracle to the reference itself rather than to a reading of its sources: everything except
the fixed-function rasterizer is defined by these SPIR-V modules."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle
from sparsevoxeloctree_b200 import api
from tests.parity import morton_np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = sorted(glob.glob(os.path.join(HERE, "golden", "spirv_build_*.npz")))
VOX = os.path.join(HERE, "golden", "spirv_voxelizer_L6_conservative.npz")


def unpack(packed):
    x = packed[:, 0] & 0xFFF
    y = (packed[:, 0] >> 12) & 0xFFF
    z = (packed[:, 0] >> 24) | ((packed[:, 1] >> 28) << 8)
    return x, y, z, packed[:, 1] & 0xFFFFFF


@pytest.mark.parametrize("path", BUILD, ids=[os.path.basename(p) for p in BUILD])
def test_oracle_level_loop_equals_reference_compute_shaders(path):
    g = np.load(path)
    level = int(g["level"])
    x, y, z, c = unpack(g["packed"])
    words, rng = oracle.build_octree(oracle.frags_from_xyzc(x, y, z, c), level)
    # same invocation order, same allocation order -> the node BUFFER is identical word for word
    assert rng == int(g["range_bytes"])
    assert (words == g["words"]).all()


def test_oracle_geometry_stage_equals_reference_geom_shader():
    g = np.load(VOX)
    level = int(g["level"])
    res = 1 << level
    for t, go in zip(g["triangles"], g["geom_out"]):
        a, xy = oracle.debug_tri_setup(t[0], t[1], t[2], level)
        assert a.tolist() == go[:7].tolist()          # gAxis, gAABB, gDepthRange  (voxelizer.geom:30-42)
        ndc = go[7:].astype(np.uint32).view(np.float32).reshape(3, 3)
        # the emitted gl_Position (Project(), voxelizer.geom:15-19) through the viewport transform + snapping
        sx = np.rint((((ndc[:, 0] + np.float32(1)) * np.float32(0.5)) * np.float32(res)) * np.float32(256)).astype(np.int64)
        sy = np.rint((((ndc[:, 1] + np.float32(1)) * np.float32(0.5)) * np.float32(res)) * np.float32(256)).astype(np.int64)
        got = sorted(zip(sx.tolist(), sy.tolist()))
        assert got == sorted(map(tuple, xy.tolist()))  # the oracle may have swapped two vertices (winding)


def test_oracle_fragment_stage_equals_reference_frag_shader():
    g = np.load(VOX)
    level, mode = int(g["level"]), int(g["mode"])
    res = 1 << level
    fi, fo = g["frag_in"], g["frag_out"]
    draws = np.array([(0, 3, 0xFFFFFFFF, int(g["albedo"]))], oracle.DRAW_DTYPE)
    n_pix = n_emit = n_zdiff = 0
    for ti, t in enumerate(g["triangles"]):
        sel = fi["tri"] == ti
        outs = fo[sel]
        ofr = oracle.voxelize(t, np.array([0, 1, 2], np.uint32), draws, level, mode)
        kept = outs[outs[:, 0] == 1]
        # the discards (gAABB test, voxelizer.frag:21-22) agree: same number of surviving fragments, same order
        assert len(kept) == len(ofr), (ti, len(kept), len(ofr))
        x, y, z, c = unpack(kept[:, 1:3])
        n_pix += int(sel.sum())
        n_emit += len(kept)
        assert (c == ofr["rgb"]).all()
        axis = int(g["geom_out"][ti][0])
        zax = {0: "x", 1: "y", 2: "z"}[axis]   # the depth axis carries fp32-vs-fp64 rounding of gl_FragCoord.z
        for name, got in (("x", x), ("y", y), ("z", z)):
            d = got.astype(np.int64) - ofr[name].astype(np.int64)
            if name == zax:
                n_zdiff += int((d != 0).sum())
                assert (np.abs(d) <= 1).all()
            else:
                assert (d == 0).all()
    # gl_FragCoord.z reaches the shader as fp32; the oracle floors the fp64 plane value.  They may differ by one
    # voxel only where z*res is within an fp32 ulp of an integer: a handful of pixels
    assert n_emit > 10000 and n_zdiff <= 1e-3 * n_emit, (n_emit, n_zdiff)


def test_oracle_textured_shading_equals_reference_frag_shader():
    """voxelizer.frag:27-36 on injected texture() results: discard iff alpha < 0.5, colour = packUnorm4x8(x) & 0xffffff,
    counter bumped only for surviving fragments."""
    g = np.load(os.path.join(HERE, "golden", "spirv_frag_textured.npz"))
    vals = g["values"].view(np.float32)
    for v, (alive, rgb, counter) in zip(vals, g["out"]):
        o = oracle.debug_shade(v)
        assert (o is not None) == bool(alive) == bool(counter), v
        if alive:
            assert o == int(rgb), (v, hex(o), hex(int(rgb)))


TRACER = sorted(glob.glob(os.path.join(HERE, "golden", "spirv_tracer_*.npz")))
PIPELINES = {"modeA": os.path.join(HERE, "golden", "spirv_pipeline_soup260_L6.npz"),        # VK_EXT_conservative_rasterization
             "modeB": os.path.join(HERE, "golden", "spirv_pipeline_soup260_L6_modeB.npz")}  # voxelizer_conservative.geom fallback


def check_hits_against_tracer_golden(g, hit, pos, colour, normal, iters):
    """The executed octree_tracer.frag wrote oColor for the views normal*0.5+0.5, pos-1, pow(colour, 1/2.2) (a miss
    shows normal 0 and pos 1, octree_tracer.frag:37-41) and the iteration count."""
    r = g["rays"]
    ghit = (r["normal"] != 0.5).any(axis=1)
    assert (hit == ghit).all(), "hit / miss differs from the reference tracer"
    assert (iters == r["iter"]).all(), "iteration counts differ"
    h = ghit
    assert ((normal[h] * np.float32(0.5) + np.float32(0.5)) == r["normal"][h]).all()
    assert ((pos[h] - np.float32(1.0)) == r["pos"][h]).all(), "hit positions differ (bit-exact expected)"
    assert np.allclose(np.power(colour[h].astype(np.float64), 1.0 / 2.2), r["colour"][h], atol=1e-6)  # pow: driver precision
    return int(h.sum())


@pytest.mark.parametrize("path", TRACER, ids=[os.path.basename(p) for p in TRACER])
def test_oracle_raymarch_equals_reference_tracer_shader(path):
    """Octree_RayMarchLeaf as compiled into octree_tracer.frag, executed on a node buffer built by the reference's own
    builder shaders, versus the oracle's restatement: hits, iteration counts, normals and positions bit for bit."""
    g = np.load(path)
    words = np.load(os.path.join(HERE, "golden", "spirv_build_" + str(g["build_case"]) + ".npz"))["words"]
    r, cams = g["rays"], g["cameras"]
    out = [oracle.raymarch_leaf(words, cams[c][:3], d) for c, d in zip(r["cam"], r["d"])]
    n = check_hits_against_tracer_golden(g, np.array([o[0] for o in out]), np.array([o[1] for o in out]),
                                         np.array([o[2] for o in out]), np.array([o[3] for o in out]),
                                         np.array([o[4] for o in out]))
    assert n > 100


def cuda_tree_traced_like_reference(lib, path):
    """The drop-in claim end to end: the tree the CUDA builder makes from the fragments the reference's shaders consumed,
    traversed by the CUDA port of the reference's ray marcher, shows what the reference's tracer shows on the
    reference-built tree -- although the two node buffers order their blocks differently."""
    g = np.load(path)
    b = np.load(os.path.join(HERE, "golden", "spirv_build_" + str(g["build_case"]) + ".npz"))
    level = int(b["level"])
    x, y, z, c = unpack(b["packed"])
    keys = (morton_np(x, y, z, level) << np.uint64(24)) | c.astype(np.uint64)
    vox = api.Voxelizer.CreateFromFragments(keys, level, lib=lib)
    builder = api.OctreeBuilder.Create(vox)
    vox.CmdVoxelize()
    builder.CmdBuild()
    r, cams = g["rays"], g["cameras"]
    hits = api.raymarch_leaf(builder.GetOctree(), cams[r["cam"]][:, :3], r["d"], lib=lib)
    n = check_hits_against_tracer_golden(g, hits["hit"] != 0, hits["pos"], hits["colour"], hits["normal"], hits["iter"])
    assert n > 100


def pipeline_mesh():
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_spirv_golden  # only its scene parameters: nothing of /root/reference is touched
    return make_spirv_golden.pipeline_mesh(), make_spirv_golden.PIPELINE_SCENE


@pytest.mark.parametrize("which", sorted(PIPELINES))
def test_oracle_whole_path_equals_reference_shaders_end_to_end(which):
    """One scene through every programmable stage of the reference (geometry, fragment, the four builder shaders, the
    tracer), all executed from its binaries, versus the oracle: the SAME fragment list in the same order, the same node
    buffer word for word, the same traced image."""
    g = np.load(PIPELINES[which])
    mesh, c = pipeline_mesh()
    level = int(g["level"])
    fr = oracle.voxelize(mesh.positions, mesh.indices, mesh.draws, level, int(g["mode"]))
    packed = np.array([oracle.pack_fragment(int(f["x"]), int(f["y"]), int(f["z"]), int(f["rgb"])) for f in fr], np.uint32)
    assert packed.shape == g["packed"].shape and (packed == g["packed"]).all()
    words, rng = oracle.build_octree(fr, level)
    assert rng == int(g["range_bytes"]) and (words == g["words"]).all()
    r, cams = g["rays"], g["cameras"]
    out = [oracle.raymarch_leaf(words, cams[c_][:3], d) for c_, d in zip(r["cam"], r["d"])]
    n = check_hits_against_tracer_golden(g, np.array([o[0] for o in out]), np.array([o[1] for o in out]),
                                         np.array([o[2] for o in out]), np.array([o[3] for o in out]),
                                         np.array([o[4] for o in out]))
    assert n > 40


def cuda_whole_path_like_reference(lib, which="modeA"):
    """Scene -> Voxelizer -> OctreeBuilder -> ray marcher, all CUDA, versus the executed reference shaders end to end.
    Every triangle of the scene is small, so the CUDA emission order is the reference's draw order and even the colours
    of voxels shared by several materials come out identical."""
    from tests.parity import assert_same_tree
    g = np.load(PIPELINES[which])
    mesh, c = pipeline_mesh()
    level = int(g["level"])
    scene = api.Scene.Create(mesh, lib=lib)
    vox = api.Voxelizer.Create(scene, level, int(g["mode"]))
    builder = api.OctreeBuilder.Create(vox)
    vox.CmdVoxelize()
    x, y, z, col = unpack(g["packed"])
    gkeys = (morton_np(x, y, z, level) << np.uint64(24)) | col.astype(np.uint64)
    assert (vox.fragments_to_host() == gkeys).all(), "fragment list differs from voxelizer.geom/.frag (order included)"
    assert (vox.reference_fragments_to_host() == g["packed"]).all()  # and in the reference's own uvec2 packing
    builder.CmdBuild()
    assert builder.GetOctreeRange() == int(g["range_bytes"])
    assert_same_tree(builder.octree_to_host(), g["words"], level)
    r, cams = g["rays"], g["cameras"]
    hits = api.raymarch_leaf(builder.GetOctree(), cams[r["cam"]][:, :3], r["d"], lib=lib)
    assert check_hits_against_tracer_golden(g, hits["hit"] != 0, hits["pos"], hits["colour"], hits["normal"], hits["iter"]) > 40


TEXTURED_PIPELINE = os.path.join(HERE, "golden", "spirv_pipeline_texsoup160_L6.npz")


def textured_pipeline_mesh():
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_spirv_golden
    return make_spirv_golden.textured_pipeline_mesh()


def test_oracle_textured_path_equals_reference_shaders_end_to_end():
    """Textured and alpha-tested materials through voxelizer.geom / voxelizer.frag / the builder shaders (the texture
    unit's value injected from the pinned sampler): same fragment list in the same order, same node buffer."""
    g = np.load(TEXTURED_PIPELINE)
    assert int(g["discarded"]) > 20 and int(g["sampled"]) > 500
    mesh = textured_pipeline_mesh()
    level = int(g["level"])
    fr = oracle.voxelize(mesh.positions, mesh.indices, mesh.draws, level, int(g["mode"]), texcoords=mesh.texcoords,
                         texset=oracle.TexSet(mesh.textures))
    packed = np.array([oracle.pack_fragment(int(f["x"]), int(f["y"]), int(f["z"]), int(f["rgb"])) for f in fr], np.uint32)
    assert packed.shape == g["packed"].shape and (packed == g["packed"]).all()
    words, rng = oracle.build_octree(fr, level)
    assert rng == int(g["range_bytes"]) and (words == g["words"]).all()


def cuda_textured_path_like_reference(lib):
    from tests.parity import assert_same_tree
    g = np.load(TEXTURED_PIPELINE)
    mesh = textured_pipeline_mesh()
    level = int(g["level"])
    scene = api.Scene.Create(mesh, lib=lib)
    vox = api.Voxelizer.Create(scene, level, int(g["mode"]))
    builder = api.OctreeBuilder.Create(vox)
    vox.CmdVoxelize()
    assert (vox.reference_fragments_to_host() == g["packed"]).all(), "textured fragment list differs (order included)"
    builder.CmdBuild()
    assert builder.GetOctreeRange() == int(g["range_bytes"])
    assert_same_tree(builder.octree_to_host(), g["words"], level)


@pytest.mark.gpu
def test_cuda_textured_path_equals_reference_shaders_end_to_end():
    cuda_textured_path_like_reference(api.get_library())


@pytest.mark.gpu
@pytest.mark.parametrize("which", sorted(PIPELINES))
def test_cuda_whole_path_equals_reference_shaders_end_to_end(which):
    cuda_whole_path_like_reference(api.get_library(), which)


@pytest.mark.gpu
@pytest.mark.parametrize("path", TRACER, ids=[os.path.basename(p) for p in TRACER])
def test_cuda_tree_and_raymarcher_equal_reference_tracer(path):
    cuda_tree_traced_like_reference(api.get_library(), path)


@pytest.mark.gpu
@pytest.mark.parametrize("path", BUILD, ids=[os.path.basename(p) for p in BUILD])
def test_cuda_builder_equals_reference_compute_shaders(path):
    """OctreeBuilder on the very fragment list the reference's shaders consumed: canonically identical tree,
    colours included (the stable sort keeps the emission order that the running average depends on)."""
    from tests.parity import assert_same_tree
    g = np.load(path)
    level = int(g["level"])
    x, y, z, c = unpack(g["packed"])
    keys = (morton_np(x, y, z, level) << np.uint64(24)) | c.astype(np.uint64)
    vox = api.Voxelizer.CreateFromFragments(keys, level)
    b = api.OctreeBuilder.Create(vox)
    vox.CmdVoxelize()
    b.CmdBuild()
    assert b.GetOctreeRange() == int(g["range_bytes"])
    assert_same_tree(b.octree_to_host(), g["words"], level)


def test_oracle_mode_b_dilation_equals_reference_conservative_geom_shader():
    """Reference Mode B (used when VK_EXT_conservative_rasterization is absent, Voxelizer.cpp:92-99): the dilated
    vertices voxelizer_conservative.geom emits, bit for bit (incl. where its compiled SPIR-V fuses multiply-adds)."""
    g = np.load(os.path.join(HERE, "golden", "spirv_conservative_geom_L6.npz"))
    level = int(g["level"])
    n_checked = 0
    for t, em, flat in zip(g["triangles"], g["emitted"], g["flat"]):
        a, _ = oracle.debug_tri_setup(t[0], t[1], t[2], level)
        assert a.tolist() == flat.tolist()                       # gAxis / gAABB / gDepthRange: those of the ORIGINAL triangle
        mine = oracle.debug_dilate(t[0], t[1], t[2], level).view(np.uint32)
        ref = em.astype(np.uint32)
        if np.isnan(ref.view(np.float32)).any():                 # degenerate input: both sides produce NaN
            assert np.isnan(mine.view(np.float32)).any()
            continue
        assert (mine == ref).all()
        n_checked += 1
    assert n_checked > 250
