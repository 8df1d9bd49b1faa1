"""ctypes binding of the CPU oracle (oracle/svo_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; the product package (sparsevoxeloctree_b200) never does.
Parity status (pinned against the executed reference SPIR-V, rasterizer stage unpinned): see oracle/svo_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsvo_oracle.so")

CENTER, CONSERVATIVE_EXACT, CONSERVATIVE_DILATE = 0, 1, 2

DRAW_DTYPE = np.dtype([("first_index", "<u4"), ("index_count", "<u4"), ("texture_id", "<u4"), ("albedo_rgba8", "<u4")])
FRAG_DTYPE = np.dtype([("x", "<u4"), ("y", "<u4"), ("z", "<u4"), ("rgb", "<u4")])


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("svo_oracle.c", "svo_oracle.h", "Makefile"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < src_m:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_pack_fragment.argtypes = [C.c_uint32] * 4 + [C.POINTER(C.c_uint32)]
        L.orc_pack_fragment.restype = None
        L.orc_unpack_fragment.argtypes = [C.POINTER(C.c_uint32)] + [C.POINTER(C.c_uint32)] * 4
        L.orc_unpack_fragment.restype = None
        L.orc_octree_entry_num.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_octree_entry_num.restype = C.c_uint32
        L.orc_voxelize.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.orc_voxelize.restype = C.c_int64
        L.orc_texset_create.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_texset_create.restype = C.c_void_p
        L.orc_texset_destroy.argtypes = [C.c_void_p]
        L.orc_texset_destroy.restype = None
        L.orc_texset_level.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                       C.POINTER(C.c_void_p)]
        L.orc_texset_level.restype = C.c_int
        L.orc_voxelize_textured.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                            C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                            C.c_int]
        L.orc_voxelize_textured.restype = C.c_int64
        L.orc_debug_sample.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 6 + [C.c_uint32, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_debug_sample.restype = C.c_uint32
        L.orc_raymarch_leaf.argtypes = [C.c_void_p] * 6 + [C.POINTER(C.c_uint32)]
        L.orc_raymarch_leaf.restype = C.c_int
        L.orc_debug_texture_fetch.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 6 + [C.c_uint32, C.c_int32, C.c_int32, C.c_void_p,
                                                                                         C.c_void_p]
        L.orc_debug_texture_fetch.restype = None
        L.orc_debug_shade.argtypes = [C.c_void_p]
        L.orc_debug_shade.restype = C.c_uint32
        L.orc_build.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p, C.c_uint64, C.c_int]
        L.orc_build.restype = C.c_int64
        L.orc_canonicalise.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_canonicalise.restype = C.c_int64
        L.orc_debug_tri_setup.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_debug_tri_setup.restype = None
        L.orc_debug_raster_pixels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_int64]
        L.orc_debug_raster_pixels.restype = C.c_int64
        L.orc_debug_dilate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_debug_dilate.restype = None
        L.orc_morton.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_morton.restype = C.c_uint64
        _lib = L
    return _lib


def pack_fragment(x, y, z, colour):
    out = (C.c_uint32 * 2)()
    lib().orc_pack_fragment(x, y, z, colour, out)
    return int(out[0]), int(out[1])


def unpack_fragment(lo, hi):
    a = (C.c_uint32 * 2)(lo, hi)
    x, y, z, c = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    lib().orc_unpack_fragment(a, C.byref(x), C.byref(y), C.byref(z), C.byref(c))
    return x.value, y.value, z.value, c.value


def octree_entry_num(fragment_count, level):
    return int(lib().orc_octree_entry_num(fragment_count, level))


def morton(x, y, z, level):
    return int(lib().orc_morton(x, y, z, level))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class _orc_texture(C.Structure):
    _fields_ = [("rgba8", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class TexSet:
    """Textures with the mip chains Scene::load_textures builds.  images: list of uint8 [H,W,4] arrays (sRGB RGBA)."""

    def __init__(self, images):
        self.images = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        arr = (_orc_texture * max(1, len(self.images)))()
        for i, im in enumerate(self.images):
            assert im.ndim == 3 and im.shape[2] == 4
            arr[i] = _orc_texture(im.ctypes.data, im.shape[1], im.shape[0])
        self._h = lib().orc_texset_create(C.cast(arr, C.c_void_p), len(self.images))
        if not self._h:
            raise ValueError("orc_texset_create failed")

    def level(self, tex, level):
        w, h, d = C.c_uint32(), C.c_uint32(), C.c_void_p()
        n = lib().orc_texset_level(self._h, tex, level, C.byref(w), C.byref(h), C.byref(d))
        if n < 0:
            raise IndexError((tex, level))
        buf = (C.c_uint8 * (w.value * h.value * 4)).from_address(d.value)
        return np.frombuffer(buf, dtype=np.uint8).reshape(h.value, w.value, 4).copy()

    def level_count(self, tex):
        w, h, d = C.c_uint32(), C.c_uint32(), C.c_void_p()
        return lib().orc_texset_level(self._h, tex, 0, C.byref(w), C.byref(h), C.byref(d))

    def sample(self, tex, tri_pos, tri_uv, level, px, py):
        """(rgb or None when discarded, (hi, lo, delta*256)) of pixel (px,py) of one textured triangle."""
        p = [np.ascontiguousarray(v, dtype=np.float32) for v in tri_pos]
        t = [np.ascontiguousarray(v, dtype=np.float32) for v in tri_uv]
        lod = np.zeros(3, np.uint32)
        r = lib().orc_debug_sample(self._h, tex, _ptr(p[0]), _ptr(p[1]), _ptr(p[2]), _ptr(t[0]), _ptr(t[1]), _ptr(t[2]), level,
                                   px, py, _ptr(lod))
        return (None if r == 0 else int(r & 0xffffff)), tuple(int(x) for x in lod)

    def fetch(self, tex, tri_pos, tri_uv, level, px, py):
        """(uv float64[2], rgba float32[4]) that reach voxelizer.frag for pixel (px,py) of one textured triangle."""
        p = [np.ascontiguousarray(v, dtype=np.float32) for v in tri_pos]
        t = [np.ascontiguousarray(v, dtype=np.float32) for v in tri_uv]
        uv, rgba = np.zeros(2, np.float64), np.zeros(4, np.float32)
        lib().orc_debug_texture_fetch(self._h, tex, _ptr(p[0]), _ptr(p[1]), _ptr(p[2]), _ptr(t[0]), _ptr(t[1]), _ptr(t[2]), level,
                                      px, py, _ptr(uv), _ptr(rgba))
        return uv, rgba

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_texset_destroy(self._h)
            self._h = None


def voxelize(positions, indices, draws, level, mode=CENTER, shard=None, nthreads=1, count_only=False, texcoords=None,
             texset=None):
    """positions: float32 [V,3] (or [V,5] pos+uv like the reference Vertex); indices uint32; draws DRAW_DTYPE.
    texcoords (float32 [V,2]) + texset (TexSet) enable textured draws.
    Returns a FRAG_DTYPE array (emission order when nthreads == 1)."""
    if texset is not None or texcoords is not None:
        return _voxelize_textured(positions, texcoords, indices, draws, texset, level, mode, shard, nthreads, count_only)
    positions = np.ascontiguousarray(positions, dtype=np.float32)
    indices = np.ascontiguousarray(indices, dtype=np.uint32)
    draws = np.ascontiguousarray(draws, dtype=DRAW_DTYPE)
    stride = positions.shape[1] * 4
    lo = hi = None
    if shard is not None:
        lo = np.ascontiguousarray(shard[0], dtype=np.uint32)
        hi = np.ascontiguousarray(shard[1], dtype=np.uint32)
    L = lib()
    n = L.orc_voxelize(_ptr(positions), stride, _ptr(indices), _ptr(draws), len(draws), level, mode, _ptr(lo), _ptr(hi),
                       None, 0, nthreads)
    if n < 0:
        raise ValueError(f"orc_voxelize failed ({n})")
    if count_only:
        return int(n)
    out = np.zeros(n, dtype=FRAG_DTYPE)
    n2 = L.orc_voxelize(_ptr(positions), stride, _ptr(indices), _ptr(draws), len(draws), level, mode, _ptr(lo), _ptr(hi),
                        _ptr(out), n, nthreads)
    assert n2 == n
    return out


def _voxelize_textured(positions, texcoords, indices, draws, texset, level, mode, shard, nthreads, count_only):
    positions = np.ascontiguousarray(positions, dtype=np.float32)
    texcoords = np.ascontiguousarray(texcoords, dtype=np.float32)
    indices = np.ascontiguousarray(indices, dtype=np.uint32)
    draws = np.ascontiguousarray(draws, dtype=DRAW_DTYPE)
    lo = hi = None
    if shard is not None:
        lo = np.ascontiguousarray(shard[0], dtype=np.uint32)
        hi = np.ascontiguousarray(shard[1], dtype=np.uint32)
    L = lib()

    def call(out, cap):
        return L.orc_voxelize_textured(_ptr(positions), positions.shape[1] * 4, _ptr(texcoords), texcoords.shape[1] * 4,
                                       _ptr(indices), _ptr(draws), len(draws), texset._h if texset is not None else None, level,
                                       mode, _ptr(lo), _ptr(hi), out, cap, nthreads)

    n = call(None, 0)
    if n < 0:
        raise ValueError(f"orc_voxelize_textured failed ({n})")
    if count_only:
        return int(n)
    out = np.zeros(n, dtype=FRAG_DTYPE)
    assert call(_ptr(out), n) == n
    return out


def build_octree(frags, level, cap_words=None, nthreads=1):
    """The reference level loop.  Returns (words[:range/4] uint32, range_bytes)."""
    frags = np.ascontiguousarray(frags, dtype=FRAG_DTYPE)
    if cap_words is None:
        # exact upper bound for the literal loop: one 8-word block per non-leaf node plus the root block
        n = max(1, len(frags))
        cap_words = 8 * (1 + sum(min(8 ** d, n) for d in range(1, level)))
        cap_words = min(cap_words, 1 << 30)
    words = np.zeros(cap_words, dtype=np.uint32)
    r = lib().orc_build(_ptr(frags), len(frags), level, _ptr(words), cap_words, nthreads)
    if r < 0:
        raise MemoryError("octree buffer too small (the reference would write out of bounds)")
    return words[: r // 4].copy(), int(r)


def canonicalise(words, level):
    """Morton-DFS canonical form: (depth uint8[], morton uint64[], word uint32[])."""
    words = np.ascontiguousarray(words, dtype=np.uint32)
    L = lib()
    n = L.orc_canonicalise(_ptr(words), len(words), level, None, None, None, 0)
    if n < 0:
        raise ValueError(f"malformed octree ({n})")
    depth = np.zeros(n, dtype=np.uint8)
    mort = np.zeros(n, dtype=np.uint64)
    word = np.zeros(n, dtype=np.uint32)
    n2 = L.orc_canonicalise(_ptr(words), len(words), level, _ptr(depth), _ptr(mort), _ptr(word), n)
    assert n2 == n
    return depth, mort, word


def debug_tri_setup(p0, p1, p2, level):
    """(axis, gAABB[4], gDepthRange[2]), snapped xy[3][2] of one triangle (geometry-stage outputs)."""
    p = [np.ascontiguousarray(v, dtype=np.float32) for v in (p0, p1, p2)]
    a = np.zeros(7, np.uint32)
    xy = np.zeros(6, np.int32)
    lib().orc_debug_tri_setup(_ptr(p[0]), _ptr(p[1]), _ptr(p[2]), level, _ptr(a), _ptr(xy))
    return a, xy.reshape(3, 2)


def debug_dilate(p0, p1, p2, level):
    """The three vertices voxelizer_conservative.geom emits (ndc x, ndc y, depth), float32 [3,3]."""
    p = [np.ascontiguousarray(v, dtype=np.float32) for v in (p0, p1, p2)]
    out = np.zeros(9, np.float32)
    lib().orc_debug_dilate(_ptr(p[0]), _ptr(p[1]), _ptr(p[2]), level, _ptr(out))
    return out.reshape(3, 3)


def debug_raster_pixels(p0, p1, p2, level, mode):
    """Covered pixels (px, py) and the pinned fp64 depth of one triangle, row-major."""
    p = [np.ascontiguousarray(v, dtype=np.float32) for v in (p0, p1, p2)]
    L = lib()
    n = L.orc_debug_raster_pixels(_ptr(p[0]), _ptr(p[1]), _ptr(p[2]), level, mode, None, None, None, 0)
    px, py, z = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64)
    L.orc_debug_raster_pixels(_ptr(p[0]), _ptr(p[1]), _ptr(p[2]), level, mode, _ptr(px), _ptr(py), _ptr(z), n)
    return px, py, z


def raymarch_leaf(words, o, d):
    """Octree_RayMarchLeaf (octree.glsl:179-340) -> (hit, pos[3], colour[3], normal[3], iterations)."""
    words = np.ascontiguousarray(words, dtype=np.uint32)
    o = np.ascontiguousarray(o, dtype=np.float32)
    d = np.ascontiguousarray(d, dtype=np.float32)
    pos, col, nrm = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
    it = C.c_uint32()
    hit = lib().orc_raymarch_leaf(_ptr(words), _ptr(o), _ptr(d), _ptr(pos), _ptr(col), _ptr(nrm), C.byref(it))
    return bool(hit), pos, col, nrm, int(it.value)


def debug_shade(rgba):
    """voxelizer.frag's use of a texture sample: packed rgb (int) or None when the alpha test discards it."""
    v = np.ascontiguousarray(rgba, dtype=np.float32)
    r = lib().orc_debug_shade(_ptr(v))
    return None if r == 0 else int(r & 0xFFFFFF)


def frags_from_xyzc(x, y, z, rgb):
    out = np.zeros(len(x), dtype=FRAG_DTYPE)
    out["x"], out["y"], out["z"], out["rgb"] = x, y, z, rgb
    return out
