#!/usr/bin/env python
"""Runs the REFERENCE'S OWN shader binaries (shader/include/spirv/*.u32 under /root/reference) through the small
SPIR-V interpreter in oracle/spirv_interp.py and records their outputs as golden vectors.

  part A (OctreeBuilder): octree_init_node / octree_tag_node / octree_alloc_node / octree_modify_arg, driven by
          a restatement of the dispatch sequence of OctreeBuilder::CmdBuild (src/OctreeBuilder.cpp:142-210),
          on fragment lists in the reference's uvec2 packing  ->  the node buffer, word for word.
  part B (Voxelizer): voxelizer.geom per triangle (gAxis, gAABB, gDepthRange, projected vertices) and
          voxelizer.frag per covered pixel (voxel position + packed fragment).  The pixel list and the depth at
          the pixel centre come from the pinned rasterizer arithmetic (the fixed-function stage is the one thing
          the shaders do not define); gl_FragCoord.z is that depth rounded to fp32, as a shader input would be.

The interpreter needs /root/reference, which does not exist on the GPU box: the vectors are committed
(tests/golden/spirv_*.npz) and tests/test_spirv_golden.py checks the oracle against them anywhere.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle, spirv_interp as si  # noqa: E402
from sparsevoxeloctree_b200 import scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SPV = "/root/reference/shader/include/spirv/"


def gx64(x):  # group_x_64, src/OctreeBuilder.cpp:6
    return (x >> 6) + (1 if x & 63 else 0)


def spirv_build(frag_packed, level, cap_words):
    """OctreeBuilder::CmdBuild with the reference's compute shaders (subgroup size 1, invocations in order)."""
    F, res = len(frag_packed), 1 << level
    init = si.Module.from_u32_file(SPV + "octree_init_node.comp.u32")
    tag = si.Module.from_u32_file(SPV + "octree_tag_node.comp.u32", spec={0: res, 1: F})  # OctreeBuilder.cpp:102-104
    alloc = si.Module.from_u32_file(SPV + "octree_alloc_node.comp.u32")
    mod = si.Module.from_u32_file(SPV + "octree_modify_arg.comp.u32")
    counter = np.zeros(1, np.uint32)                       # Counter::Reset(0), OctreeBuilder.cpp:14-15
    octree = np.full(cap_words, 0xCDCDCDCD, np.uint32)     # uninitialised device memory
    frags = np.ascontiguousarray(frag_packed, dtype=np.uint32).reshape(-1)
    info = np.array([0, 8], np.uint32)                     # build_info staging, OctreeBuilder.cpp:27-31
    indirect = np.array([1, 1, 1], np.uint32)              # indirect staging, :36-40
    bufs = {(0, 0): counter, (0, 1): octree, (0, 2): frags, (0, 3): info, (0, 4): indirect}
    for i in range(1, level + 1):                          # OctreeBuilder.cpp:167
        for g in range(int(indirect[0]) * 64):
            init.run({("builtin", 28): [g, 0, 0]}, bufs)
        for g in range(gx64(F) * 64):
            tag.run({("builtin", 28): [g, 0, 0]}, bufs)
        if i != level:
            for g in range(int(indirect[0]) * 64):
                alloc.run({("builtin", 28): [g, 0, 0]}, bufs)
            mod.run({}, bufs)
    rng = (int(counter[0]) + 1) * 8 * 4                    # GetOctreeRange, OctreeBuilder.cpp:212-214
    return octree[: rng // 4].copy(), rng


BUILD_CASES = {
    "soup200_L5_center": (lambda: scenes.random_soup(200, 6, 0.005, 0.6), 5, oracle.CENTER),
    "soup60_L7_conservative": (lambda: scenes.random_soup(60, 9, 0.01, 0.5), 7, oracle.CONSERVATIVE_EXACT),
    "heightfield11_L5_conservative": (lambda: scenes.heightfield(11), 5, oracle.CONSERVATIVE_EXACT),
}


def make_build_case(name):
    gen, level, mode = BUILD_CASES[name]
    mesh = gen()
    fr = oracle.voxelize(mesh.positions, mesh.indices, mesh.draws, level, mode)
    packed = np.array([oracle.pack_fragment(int(f["x"]), int(f["y"]), int(f["z"]), int(f["rgb"])) for f in fr], np.uint32)
    cap = 8 * (1 + sum(min(8 ** d, len(fr)) for d in range(1, level)))
    words, rng = spirv_build(packed, level, cap)
    return dict(level=level, packed=packed, words=words, range_bytes=rng)


def voxelizer_triangles():
    """A few hundred triangles: random soups of several size classes, a heightfield patch, axis ties, slivers."""
    tris = []
    for seed, lo, hi in ((11, 0.02, 0.3), (12, 0.2, 1.5), (13, 0.005, 0.05)):
        m = scenes.random_soup(70, seed, lo, hi)
        tris += [m.positions[m.indices[3 * k:3 * k + 3]] for k in range(m.n_triangles)]
    m = scenes.heightfield(7)
    tris += [m.positions[m.indices[3 * k:3 * k + 3]] for k in range(m.n_triangles)]
    tris.append(np.array([[0.0, 0.0, 0.0], [0.5, -0.5, 0.0], [0.0, 0.0, 0.5]], np.float32))      # |nx| == |ny| tie
    tris.append(np.array([[-0.9, 0.3, 0.1], [0.9, 0.3001, 0.1], [0.0, 0.3, 0.1003]], np.float32))  # sliver
    tris.append(np.array([[0.25, -1.0, -1.0], [0.25, 1.0, -1.0], [0.25, -1.0, 1.0]], np.float32))  # big, axis 0
    return [np.ascontiguousarray(t, dtype=np.float32) for t in tris]


def make_voxelizer_case(level=6, mode=oracle.CONSERVATIVE_EXACT, albedo=0x00A1B2C3):
    res = 1 << level
    geom = si.Module.from_u32_file(SPV + "voxelizer.geom.u32", spec={0: res})
    frag = si.Module.from_u32_file(SPV + "voxelizer.frag.u32", spec={0: res, 1: 1})
    g_pos = next(v for v, (t, sc) in geom.vars.items() if sc == 3 and geom.types[geom.types[t][2]][0] == "struct")
    g_axis, g_aabb, g_zr = (geom.var_by_location(k, 3) for k in (1, 2, 3))
    f_axis, f_aabb, f_zr, f_uv = (frag.var_by_location(k, 1) for k in (1, 2, 3, 0))
    tris = voxelizer_triangles()
    geo_out, frag_in, frag_out = [], [], []
    for t in tris:
        emitted = []
        null_vtx = geom._null(geom.types[geom.types[geom.vars[next(v for v, (tt, sc) in geom.vars.items() if sc == 1 and 30 not in geom.decor.get(v, {}))][0]][2]][1])
        gl_in = []
        for k in range(3):
            v = [x if not isinstance(x, list) else list(x) for x in null_vtx]
            v[0] = [np.float32(t[k][0]), np.float32(t[k][1]), np.float32(t[k][2]), np.float32(1.0)]  # voxelizer.vert:9
            gl_in.append(v)
        geom.run({"gl_in": gl_in, ("loc", 0): [[np.float32(0), np.float32(0)]] * 3}, {}, None, on_emit=emitted.append)
        assert len(emitted) == 3
        e = emitted[0]
        axis, aabb, zr = int(e[g_axis]), [int(x) for x in e[g_aabb]], [int(x) for x in e[g_zr]]
        ndc = np.array([[float(c) for c in em[g_pos][0][:3]] for em in emitted], np.float32)
        geo_out.append([axis] + aabb + zr + [int(v) for v in ndc.view(np.uint32).reshape(-1)])
        # fixed-function stage: the pinned rasterizer arithmetic (coverage + depth at the pixel centre)
        px, py, z = oracle.debug_raster_pixels(t[0], t[1], t[2], level, mode)
        for x, y, zz in zip(px, py, z):
            counter = np.zeros(1, np.uint32)
            flist = np.zeros(2, np.uint32)
            fc = [np.float32(x + 0.5), np.float32(y + 0.5), np.float32(zz), np.float32(1.0)]
            try:
                frag.run({("builtin", 15): fc, f_axis: axis, f_aabb: aabb, f_zr: zr, f_uv: [np.float32(0), np.float32(0)]},
                         {(0, 0): counter, (0, 1): flist}, [0, 0xFFFFFFFF, albedo])  # push: uCountOnly, uTextureId, uAlbedo
                out = (1, int(flist[0]), int(flist[1]))
            except si.Discard:
                out = (0, 0, 0)
            frag_in.append((len(geo_out) - 1, int(x), int(y), float(zz)))
            frag_out.append(out)
    fi = np.array(frag_in, dtype=[("tri", "<i4"), ("px", "<i4"), ("py", "<i4"), ("z", "<f8")])
    return dict(level=level, mode=mode, albedo=albedo, triangles=np.stack(tris), geom_out=np.array(geo_out, np.uint32),
                frag_in=fi, frag_out=np.array(frag_out, np.uint32))


def make_dilate_case(level=6):
    """voxelizer_conservative.geom (reference Mode B, used without VK_EXT_conservative_rasterization): the three
    emitted vertices (dilated ndc x, y and the barycentric-extrapolated depth) per triangle, as fp32 bit patterns."""
    res = 1 << level
    geom = si.Module.from_u32_file(SPV + "voxelizer_conservative.geom.u32", spec={0: res})
    g_pos = next(v for v, (t, sc) in geom.vars.items() if sc == 3 and geom.types[geom.types[t][2]][0] == "struct")
    in_var = next(v for v, (tt, sc) in geom.vars.items() if sc == 1 and 30 not in geom.decor.get(v, {}))
    null_vtx = geom._null(geom.types[geom.types[geom.vars[in_var][0]][2]][1])
    g_axis, g_aabb, g_zr = (geom.var_by_location(k, 3) for k in (1, 2, 3))
    tris = voxelizer_triangles()
    emitted_all, flat_all = [], []
    for t in tris:
        emitted, gl_in = [], []
        for k in range(3):
            v = [x if not isinstance(x, list) else list(x) for x in null_vtx]
            v[0] = [np.float32(t[k][0]), np.float32(t[k][1]), np.float32(t[k][2]), np.float32(1.0)]
            gl_in.append(v)
        with np.errstate(all="ignore"):
            geom.run({"gl_in": gl_in, ("loc", 0): [[np.float32(0), np.float32(0)]] * 3}, {}, None, on_emit=emitted.append)
        e = emitted[0]
        flat_all.append([int(e[g_axis])] + [int(x) for x in e[g_aabb]] + [int(x) for x in e[g_zr]])
        emitted_all.append(np.array([[np.float32(c) for c in em[g_pos][0][:3]] for em in emitted], np.float32))
    return dict(level=level, triangles=np.stack(tris), emitted=np.stack(emitted_all).view(np.uint32), flat=np.array(flat_all, np.uint32))


def make_textured_frag_case(level=6, n=600, seed=9):
    """voxelizer.frag with uTextureId != 0xffffffff: what the shader does with the value texture() returns (alpha-test
    discard, packUnorm4x8, & 0xffffff) -- texture() itself is the driver's sampler and is injected here."""
    res = 1 << level
    frag = si.Module.from_u32_file(SPV + "voxelizer.frag.u32", spec={0: res, 1: 4})  # kTextureNum = 4
    f_axis, f_aabb, f_zr, f_uv = (frag.var_by_location(k, 1) for k in (1, 2, 3, 0))
    rng = np.random.default_rng(seed)
    vals = rng.uniform(-0.1, 1.1, (n, 4)).astype(np.float32)
    vals[: n // 4, 3] = rng.choice(np.array([0.5, np.nextafter(np.float32(0.5), np.float32(0)), 0.4999, 0.5001, 0.0, 1.0], np.float32), n // 4)
    k = np.arange(n // 4, n // 2)  # exact .5/255 ties of packUnorm4x8 and their neighbours
    vals[k, 0] = ((k % 255) + 0.5).astype(np.float32) / np.float32(255.0)
    vals[k, 1] = np.nextafter(vals[k, 0], np.float32(2))
    vals[k, 2] = np.nextafter(vals[k, 0], np.float32(-1))
    out, seen = [], []
    for i, v in enumerate(vals):
        counter, flist = np.zeros(1, np.uint32), np.zeros(2, np.uint32)
        tex_id, uv = int(i % 4), [np.float32(0.25 + i), np.float32(-1.5)]

        def sampler(t, coord, v=v):
            seen.append((t, float(coord[0]), float(coord[1])))
            return list(v)

        try:
            frag.run({("builtin", 15): [np.float32(3.5), np.float32(4.5), np.float32(0.3), np.float32(1.0)], f_axis: 2,
                      f_aabb: [0, 0, res - 1, res - 1], f_zr: [0, res - 1], f_uv: uv}, {(0, 0): counter, (0, 1): flist},
                     [0, tex_id, 0x00112233], sampler=sampler)
            out.append((1, int(flist[1]) & 0xFFFFFF, int(counter[0])))
        except si.Discard:
            out.append((0, 0, int(counter[0])))
        assert seen[-1] == (tex_id, float(uv[0]), float(uv[1]))  # texture(uTextures[uTextureId], gTexcoord)
    return dict(values=vals.view(np.uint32), out=np.array(out, np.uint32))


TRACER_CAMERAS = {  # position, look, side, up (Camera.cpp builds such a basis; the octree occupies [1,2]^3, octree.glsl:53)
    "outside": ([1.35, 1.6, -0.4], [0.05, -0.1, 1.0], [0.45, 0.0, 0.0], [0.0, 0.45, 0.0]),
    "inside": ([1.5, 1.52, 1.47], [0.7, 0.2, -0.68], [0.5, 0.0, 0.5], [-0.1, 0.6, 0.1]),
}


def make_tracer_case(build_case, size=28):
    """octree_tracer.frag (the reference's primary-ray consumer of the node buffer: Octree_RayMarchLeaf,
    octree.glsl:179-340) executed on a node buffer the reference's own builder shaders produced.  Per pixel: the
    ray direction the shader derived (captured at its normalize()), and oColor for the view types 'normal' (1),
    'position' (2), 'diffuse' (0), plus the iteration count behind the 'iteration' view (3)."""
    g = np.load(os.path.join(HERE, "spirv_build_" + build_case + ".npz"))
    words = g["words"].astype(np.uint32)
    frag = si.Module.from_u32_file(SPV + "octree_tracer.frag.u32")
    out_var = frag.var_by_location(0, 3)
    rows = []
    for cname, (pos, look, side, up) in TRACER_CAMERAS.items():
        cam = np.zeros(16, np.float32)
        cam[0:3], cam[4:7], cam[8:11], cam[12:15] = pos, look, side, up
        for py in range(size):
            for px in range(size):
                rec = {}

                def hook(inst, args, res, rec=rec):
                    if inst == 69: rec["d"] = [np.float32(c) for c in res]
                    if inst == 43 and not isinstance(args[0], list): rec["iter"] = int(round(float(args[0]) * 128.0))

                outs = []
                for vt in (1, 2, 0, 3):
                    with np.errstate(all="ignore"):
                        r = frag.run({("builtin", 15): [np.float32(px + 0.5), np.float32(py + 0.5), np.float32(0.5), np.float32(1)]},
                                     {(0, 0): words, (1, 0): cam.view(np.uint32)},
                                     [size, size, vt, 0, 0, 1, [np.float32(0.1), np.float32(0.2), np.float32(0.3)], np.float32(0)],
                                     on_ext=hook)
                    outs.append([np.float32(c) for c in r[out_var][:3]])
                rows.append((list(TRACER_CAMERAS).index(cname), px, py, rec["d"], outs[0], outs[1], outs[2], rec["iter"]))
    dt = np.dtype([("cam", "<i4"), ("px", "<i4"), ("py", "<i4"), ("d", "<f4", 3), ("normal", "<f4", 3), ("pos", "<f4", 3),
                   ("colour", "<f4", 3), ("iter", "<i4")])
    arr = np.array([(c, x, y, d, n, p, col, it) for c, x, y, d, n, p, col, it in rows], dtype=dt)
    cams = np.array([np.concatenate([np.asarray(v, np.float32) for v in TRACER_CAMERAS[k]]) for k in TRACER_CAMERAS], np.float32)
    return dict(build_case=build_case, size=size, cameras=cams, rays=arr)


PIPELINE_SCENE = dict(n_tri=260, seed=21, size_lo=0.02, size_hi=0.24, n_mat=3, level=6, mode=oracle.CONSERVATIVE_EXACT)


def pipeline_mesh():
    c = PIPELINE_SCENE
    return scenes.random_soup(c["n_tri"], c["seed"], c["size_lo"], c["size_hi"], c["n_mat"])


def make_pipeline_case(mode=None, tag="soup260_L6"):
    """The whole reference path on one small multi-material scene, every programmable stage executed from the
    reference's binaries: voxelizer.geom per triangle -> [pinned rasterizer: covered pixels + depth] -> voxelizer.frag
    appending to ONE fragment list through its atomic counter (Scene::CmdDraw order: draw by draw, Scene.cpp:450-463)
    -> the four builder shaders (CmdBuild) -> octree_tracer.frag on the resulting node buffer."""
    c = PIPELINE_SCENE
    level, res = c["level"], 1 << c["level"]
    mode = c["mode"] if mode is None else mode
    mesh = pipeline_mesh()
    # Mode B (no VK_EXT_conservative_rasterization, Voxelizer.cpp:92-99): the dilating geometry shader, then plain
    # centre sampling of the dilated triangle with depth clipping (depthClampEnable = 0)
    dilate = mode == oracle.CONSERVATIVE_DILATE
    geom = si.Module.from_u32_file(SPV + ("voxelizer_conservative.geom.u32" if dilate else "voxelizer.geom.u32"), spec={0: res})
    frag = si.Module.from_u32_file(SPV + "voxelizer.frag.u32", spec={0: res, 1: 1})
    g_axis, g_aabb, g_zr = (geom.var_by_location(k, 3) for k in (1, 2, 3))
    f_axis, f_aabb, f_zr, f_uv = (frag.var_by_location(k, 1) for k in (1, 2, 3, 0))
    in_var = next(v for v, (tt, sc) in geom.vars.items() if sc == 1 and 30 not in geom.decor.get(v, {}))
    null_vtx = geom._null(geom.types[geom.types[geom.vars[in_var][0]][2]][1])
    cap = 200000
    counter, flist = np.zeros(1, np.uint32), np.zeros(2 * cap, np.uint32)
    for d in mesh.draws:
        for k in range(int(d["index_count"]) // 3):
            ix = mesh.indices[int(d["first_index"]) + 3 * k: int(d["first_index"]) + 3 * k + 3]
            t = mesh.positions[ix]
            emitted, gl_in = [], []
            for q in range(3):
                v = [x if not isinstance(x, list) else list(x) for x in null_vtx]
                v[0] = [np.float32(t[q][0]), np.float32(t[q][1]), np.float32(t[q][2]), np.float32(1.0)]
                gl_in.append(v)
            with np.errstate(all="ignore"):
                geom.run({"gl_in": gl_in, ("loc", 0): [[np.float32(0), np.float32(0)]] * 3}, {}, None, on_emit=emitted.append)
            e = emitted[0]
            axis, aabb, zr = int(e[g_axis]), [int(x) for x in e[g_aabb]], [int(x) for x in e[g_zr]]
            px, py, z = oracle.debug_raster_pixels(t[0], t[1], t[2], level, mode)
            for x, y, zz in zip(px, py, z):
                if dilate and not (0.0 <= zz <= 1.0):
                    continue  # depth clip happens before the fragment shader
                try:
                    frag.run({("builtin", 15): [np.float32(x + 0.5), np.float32(y + 0.5), np.float32(zz), np.float32(1.0)],
                              f_axis: axis, f_aabb: aabb, f_zr: zr, f_uv: [np.float32(0), np.float32(0)]},
                             {(0, 0): counter, (0, 1): flist}, [0, 0xFFFFFFFF, int(d["albedo_rgba8"])])
                except si.Discard:
                    pass
    F = int(counter[0])
    packed = flist[: 2 * F].reshape(F, 2).copy()
    words, rng = spirv_build(packed, level, 8 * (1 + sum(min(8 ** q, F) for q in range(1, level))))
    np.savez_compressed(os.path.join(HERE, "spirv_build_pipeline.npz"), level=level, packed=packed, words=words, range_bytes=rng)
    rays = make_tracer_case("pipeline", size=24 if not dilate else 12)
    return dict(level=level, mode=mode, packed=packed, words=words, range_bytes=rng, cameras=rays["cameras"], rays=rays["rays"])


TEXTURED_SCENE = dict(n_tri=160, seed=23, size_lo=0.02, size_hi=0.2, level=6, mode=oracle.CONSERVATIVE_EXACT)


def textured_pipeline_mesh():
    c = TEXTURED_SCENE
    return scenes.textured_soup(c["n_tri"], c["seed"], size_lo=c["size_lo"], size_hi=c["size_hi"], big_quads=False)


def make_textured_pipeline_case():
    """The reference's geometry / fragment / builder shaders on a scene with textured and alpha-tested materials.  The two
    fixed-function stages in between are the pinned ones: the rasterizer supplies the covered pixels, the depth and the
    interpolated gTexcoord; the texture unit supplies the value of texture() -- the shaders do the rest (texture id from
    the push constant, alpha-test discard, packUnorm4x8, counter, packing, tree)."""
    c = TEXTURED_SCENE
    level, mode, res = c["level"], c["mode"], 1 << c["level"]
    mesh = textured_pipeline_mesh()
    ts = oracle.TexSet(mesh.textures)
    geom = si.Module.from_u32_file(SPV + "voxelizer.geom.u32", spec={0: res})
    frag = si.Module.from_u32_file(SPV + "voxelizer.frag.u32", spec={0: res, 1: len(mesh.textures)})
    g_axis, g_aabb, g_zr, g_uv = (geom.var_by_location(k, 3) for k in (1, 2, 3, 0))
    f_axis, f_aabb, f_zr, f_uv = (frag.var_by_location(k, 1) for k in (1, 2, 3, 0))
    in_var = next(v for v, (tt, sc) in geom.vars.items() if sc == 1 and 30 not in geom.decor.get(v, {}))
    null_vtx = geom._null(geom.types[geom.types[geom.vars[in_var][0]][2]][1])
    cap = 200000
    counter, flist = np.zeros(1, np.uint32), np.zeros(2 * cap, np.uint32)
    n_sampled = n_discarded = 0
    for d in mesh.draws:
        tex_id = int(d["texture_id"])
        for k in range(int(d["index_count"]) // 3):
            ix = mesh.indices[int(d["first_index"]) + 3 * k: int(d["first_index"]) + 3 * k + 3]
            t, tuv = mesh.positions[ix], mesh.texcoords[ix]
            emitted, gl_in = [], []
            for q in range(3):
                v = [x if not isinstance(x, list) else list(x) for x in null_vtx]
                v[0] = [np.float32(t[q][0]), np.float32(t[q][1]), np.float32(t[q][2]), np.float32(1.0)]
                gl_in.append(v)
            geom.run({"gl_in": gl_in, ("loc", 0): [[np.float32(tuv[q][0]), np.float32(tuv[q][1])] for q in range(3)]}, {}, None,
                     on_emit=emitted.append)
            for q in range(3):  # voxelizer.geom passes the texture coordinates through (voxelizer.geom:46-52)
                assert [float(x) for x in emitted[q][g_uv]] == [float(tuv[q][0]), float(tuv[q][1])]
            e = emitted[0]
            axis, aabb, zr = int(e[g_axis]), [int(x) for x in e[g_aabb]], [int(x) for x in e[g_zr]]
            px, py, z = oracle.debug_raster_pixels(t[0], t[1], t[2], level, mode)
            for x, y, zz in zip(px, py, z):
                uv, rgba = (np.zeros(2), np.zeros(4, np.float32))
                if tex_id != 0xFFFFFFFF:
                    uv, rgba = ts.fetch(tex_id, t, tuv, level, int(x), int(y))

                def sampler(tid, coord, rgba=rgba, uv=uv, tex_id=tex_id):
                    assert tid == tex_id and abs(float(coord[0]) - uv[0]) < 1e-3 and abs(float(coord[1]) - uv[1]) < 1e-3
                    return [np.float32(c_) for c_ in rgba]

                before = int(counter[0])
                try:
                    frag.run({("builtin", 15): [np.float32(x + 0.5), np.float32(y + 0.5), np.float32(zz), np.float32(1.0)],
                              f_axis: axis, f_aabb: aabb, f_zr: zr, f_uv: [np.float32(uv[0]), np.float32(uv[1])]},
                             {(0, 0): counter, (0, 1): flist}, [0, tex_id, int(d["albedo_rgba8"])], sampler=sampler)
                except si.Discard:
                    n_discarded += tex_id != 0xFFFFFFFF and int(counter[0]) == before
                n_sampled += tex_id != 0xFFFFFFFF
    F = int(counter[0])
    packed = flist[: 2 * F].reshape(F, 2).copy()
    words, rng = spirv_build(packed, level, 8 * (1 + sum(min(8 ** q, F) for q in range(1, level))))
    return dict(level=level, mode=mode, packed=packed, words=words, range_bytes=rng, sampled=n_sampled, discarded=n_discarded)


if __name__ == "__main__":
    import time
    if "--textured-pipeline-only" in sys.argv or len(sys.argv) == 1:
        t = time.time()
        o = make_textured_pipeline_case()
        np.savez_compressed(os.path.join(HERE, "spirv_pipeline_texsoup160_L6.npz"), **o)
        print("textured pipeline", len(o["packed"]), "fragments,", o["sampled"], "texture() calls,", o["discarded"], "alpha discards",
              f"{time.time() - t:.0f}s", flush=True)
        if "--textured-pipeline-only" in sys.argv:
            sys.exit(0)
    if "--pipeline-only" in sys.argv or len(sys.argv) == 1:
        t = time.time()
        for mode, name in ((None, "spirv_pipeline_soup260_L6.npz"), (oracle.CONSERVATIVE_DILATE, "spirv_pipeline_soup260_L6_modeB.npz")):
            o = make_pipeline_case(mode)
            os.remove(os.path.join(HERE, "spirv_build_pipeline.npz"))  # (scratch for make_tracer_case)
            np.savez_compressed(os.path.join(HERE, name), **o)
            print("pipeline", name, len(o["packed"]), "fragments ->", o["range_bytes"], "bytes,", len(o["rays"]), "rays",
                  f"{time.time() - t:.0f}s", flush=True)
        if "--pipeline-only" in sys.argv:
            sys.exit(0)
    if "--tracer-only" in sys.argv or "--all" in sys.argv or len(sys.argv) == 1:
        for bc in ("soup60_L7_conservative", "heightfield11_L5_conservative"):
            t = time.time()
            o = make_tracer_case(bc)
            np.savez_compressed(os.path.join(HERE, "spirv_tracer_" + bc + ".npz"), **o)
            hit = (o["rays"]["normal"] != 0.5).any(axis=1)
            print("tracer", bc, len(o["rays"]), "rays,", int(hit.sum()), "hits", f"{time.time() - t:.0f}s", flush=True)
        if "--tracer-only" in sys.argv:
            sys.exit(0)
    o = make_textured_frag_case()
    np.savez_compressed(os.path.join(HERE, "spirv_frag_textured.npz"), **o)
    print("textured frag", len(o["out"]), "samples,", int((o["out"][:, 0] == 0).sum()), "discarded", flush=True)
    if "--textured-only" in sys.argv:
        sys.exit(0)
    for n in BUILD_CASES:
        t = time.time()
        o = make_build_case(n)
        np.savez_compressed(os.path.join(HERE, "spirv_build_" + n + ".npz"), **o)
        print("build", n, len(o["packed"]), "fragments ->", o["range_bytes"], "bytes", f"{time.time() - t:.0f}s", flush=True)
    o = make_dilate_case()
    np.savez_compressed(os.path.join(HERE, "spirv_conservative_geom_L6.npz"), **o)
    print("conservative geom", len(o["triangles"]), "triangles", flush=True)
    t = time.time()
    o = make_voxelizer_case()
    np.savez_compressed(os.path.join(HERE, "spirv_voxelizer_L6_conservative.npz"), **o)
    print("voxelizer", len(o["triangles"]), "triangles", len(o["frag_in"]), "pixels", f"{time.time() - t:.0f}s")
