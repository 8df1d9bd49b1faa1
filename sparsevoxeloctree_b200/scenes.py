"""Synthetic input meshes for the five BASELINE.json configs.

The reference ships no assets, so every workload is generated: untextured materials
(texture_id = 0xffffffff, per-draw albedo packed like Scene.cpp:164), positions in [-1,1]^3
like Scene.cpp:90-99 leaves them, one draw per material sorted by size (Scene.cpp:124-132).
RNG = counter-based splitmix64 so every scene is a pure function of its seed.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

DRAW_DTYPE = np.dtype([("first_index", "<u4"), ("index_count", "<u4"), ("texture_id", "<u4"), ("albedo_rgba8", "<u4")])

# the four albedos named in SURVEY.md section 8d; packUnorm4x8(vec4(albedo, 0)): R in bits 0-7
ALBEDOS = [(200, 60, 50), (60, 180, 75), (70, 90, 200), (220, 220, 210)]


def pack_albedo(rgb) -> int:
    r, g, b = rgb
    return int(r) | (int(g) << 8) | (int(b) << 16)


@dataclass
class Mesh:
    """Host-side mesh hand-off: what Scene keeps private (Scene.hpp:26,35-40)."""
    positions: np.ndarray  # float32 [V,3]
    indices: np.ndarray    # uint32 [3T]
    draws: np.ndarray      # DRAW_DTYPE [D]
    name: str = ""
    texcoords: np.ndarray | None = None  # float32 [V,2] (only for textured draws)
    textures: list | None = None         # uint8 [H,W,4] sRGB RGBA images, indexed by draws["texture_id"]

    @property
    def n_triangles(self) -> int:
        return len(self.indices) // 3


class SplitMix64:
    """Counter-based splitmix64: value(i) = mix(seed + (i+1)*golden)."""

    def __init__(self, seed: int):
        self.seed = np.uint64(seed)
        self.ctr = 0

    def u64(self, n: int) -> np.ndarray:
        i = np.arange(self.ctr + 1, self.ctr + n + 1, dtype=np.uint64)
        self.ctr += n
        with np.errstate(over="ignore"):
            z = self.seed + i * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        return z

    def uniform(self, n: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
        u = (self.u64(n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        return lo + (hi - lo) * u


def normalize_like_scene(pos: np.ndarray) -> np.ndarray:
    """Scene.cpp:90-99 in fp32: centre on the AABB, scale the largest half-extent to 1."""
    pos = pos.astype(np.float32)
    pmin, pmax = pos.min(axis=0), pos.max(axis=0)
    extent3 = pmax - pmin
    extent = np.float32(max(extent3)) * np.float32(0.5)
    inv_extent = np.float32(1.0) / extent
    center = (pmax + pmin) * np.float32(0.5)
    return ((pos - center) * inv_extent).astype(np.float32)


def _assemble(parts, name) -> Mesh:
    """parts: list of (positions[V,3], triangles[T,3] local indices, material id or per-triangle ids[T]).
    Vertices are concatenated and indices rebased; triangles are grouped (stably) into one draw per
    material, largest draw first (Scene.cpp:124-132)."""
    pos_all, tri_all, mat_all, base = [], [], [], 0
    for pos, tri, mat in parts:
        pos = np.asarray(pos, dtype=np.float32).reshape(-1, 3)
        tri = np.asarray(tri, dtype=np.int64).reshape(-1, 3)
        pos_all.append(pos)
        tri_all.append(tri + base)
        mat_all.append(np.broadcast_to(np.asarray(mat, dtype=np.int64), (len(tri),)))
        base += len(pos)
    tri = np.concatenate(tri_all)
    mat = np.concatenate(mat_all)
    mats, counts = np.unique(mat, return_counts=True)
    order = sorted(range(len(mats)), key=lambda k: (-counts[k], mats[k]))
    idx, draws, first = [], [], 0
    for k in order:
        t = tri[mat == mats[k]].reshape(-1)
        idx.append(t)
        draws.append((first, len(t), 0xFFFFFFFF, pack_albedo(ALBEDOS[int(mats[k]) % len(ALBEDOS)])))
        first += len(t)
    return Mesh(np.concatenate(pos_all).astype(np.float32), np.concatenate(idx).astype(np.uint32),
                np.array(draws, dtype=DRAW_DTYPE), name)


def _grid_tris(nu: int, nv: int, wrap_u: bool = False) -> np.ndarray:
    """Two triangles per cell of an nu x nv vertex grid (row-major, v fastest)."""
    cu = nu if wrap_u else nu - 1
    i, j = np.meshgrid(np.arange(cu), np.arange(nv - 1), indexing="ij")
    i, j = i.reshape(-1), j.reshape(-1)
    i1 = (i + 1) % nu
    a, b, c, d = i * nv + j, i1 * nv + j, i1 * nv + j + 1, i * nv + j + 1
    return np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)])


def _box(lo, hi, inward=False):
    """12 triangles of an axis-aligned box."""
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    c = np.array([[lo[0] if not (k & 1) else hi[0], lo[1] if not (k & 2) else hi[1], lo[2] if not (k & 4) else hi[2]]
                  for k in range(8)])
    q = [(0, 2, 3, 1), (4, 5, 7, 6), (0, 1, 5, 4), (2, 6, 7, 3), (0, 4, 6, 2), (1, 3, 7, 5)]
    t = []
    for a, b, cc, d in q:
        t += [(a, b, cc), (a, cc, d)]
    t = np.array(t)
    if inward:
        t = t[:, ::-1]
    return c, t


def _small_triangles(rng: SplitMix64, n: int, centers: np.ndarray, edge_lo: float, edge_hi: float):
    """n small triangles with log-uniform edge length around the given centres (soup, 3 verts each)."""
    e = np.exp(rng.uniform(n, np.log(edge_lo), np.log(edge_hi)))
    d1 = rng.uniform(3 * n, -1.0, 1.0).reshape(n, 3)
    d2 = rng.uniform(3 * n, -1.0, 1.0).reshape(n, 3)
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True) + 1e-12
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True) + 1e-12
    p0 = centers
    p1 = centers + d1 * e[:, None]
    p2 = centers + d2 * e[:, None]
    pos = np.stack([p0, p1, p2], 1).reshape(-1, 3)
    return np.clip(pos, -0.999, 0.999), np.arange(3 * n).reshape(n, 3)


# ---------------------------------------------------------------------------------------------
# C1: 10k-triangle heightfield, level 8
# ---------------------------------------------------------------------------------------------
def heightfield(n: int = 71, seed: int = 1) -> Mesh:
    rng = SplitMix64(seed)
    g = np.linspace(-1.0, 1.0, n)
    cell = 2.0 / (n - 1)
    x, z = np.meshgrid(g, g, indexing="ij")
    jx = rng.uniform(n * n, -0.2, 0.2).reshape(n, n) * cell
    jz = rng.uniform(n * n, -0.2, 0.2).reshape(n, n) * cell
    jx[[0, -1], :] = 0.0
    jx[:, [0, -1]] = 0.0
    jz[[0, -1], :] = 0.0
    jz[:, [0, -1]] = 0.0
    x, z = x + jx, z + jz
    y = 0.35 * np.sin(3.1 * x) * np.cos(2.3 * z) + 0.1 * np.sin(9.0 * x + 1.0)
    pos = normalize_like_scene(np.stack([x, y, z], -1).reshape(-1, 3))
    tri = _grid_tris(n, n)
    # four materials by quadrant of the triangle's first vertex
    v0 = pos[tri[:, 0]]
    mat = (v0[:, 0] > 0).astype(int) + 2 * (v0[:, 2] > 0).astype(int)
    return _assemble([(pos, tri, mat)], f"heightfield{n}")


# ---------------------------------------------------------------------------------------------
# C2: Sponza-scale (~260k triangles, mixed sizes), level 10
# ---------------------------------------------------------------------------------------------
def sponza_like(seed: int = 2, n_columns: int = 48, col_u: int = 64, col_v: int = 32, n_detail: int = 60000,
                level: int = 10) -> Mesh:
    rng = SplitMix64(seed)
    parts = []
    hall_lo, hall_hi = (-1.0, -0.45, -0.6), (1.0, 0.45, 0.6)
    c, t = _box(hall_lo, hall_hi, inward=True)  # floor + 4 walls + ceiling = 12 large triangles
    parts.append((c, t, 3))
    # tessellated columns in two rows
    ang = np.arange(col_u) * (2 * np.pi / col_u)
    hv = np.linspace(0.0, 1.0, col_v + 1)
    ctris = _grid_tris(col_u, col_v + 1, wrap_u=True)
    for k in range(n_columns):
        row, col = k % 2, k // 2
        cx = -0.92 + 1.84 * (col + 0.5) / (n_columns // 2)
        cz = -0.38 if row == 0 else 0.38
        r0 = 0.022 + 0.008 * rng.uniform(1)[0]
        a, h = np.meshgrid(ang, hv, indexing="ij")
        r = r0 * (1.0 + 0.15 * np.sin(8.0 * a) * np.sin(np.pi * h))  # fluting
        px = cx + r * np.cos(a)
        pz = cz + r * np.sin(a)
        py = -0.45 + 0.9 * h
        parts.append((np.stack([px, py, pz], -1).reshape(-1, 3), ctris, k % 3))
    # small detail triangles (edge 0.5-4 voxels) scattered on the hall surfaces
    vox = 2.0 / (1 << level)
    face = (rng.u64(n_detail) % np.uint64(6)).astype(int)
    u = rng.uniform(n_detail, -0.98, 0.98)
    v = rng.uniform(n_detail, -0.98, 0.98)
    cen = np.zeros((n_detail, 3))
    lo, hi = np.array(hall_lo), np.array(hall_hi)
    for f in range(6):
        ax, side = f // 2, f % 2
        o = [a for a in range(3) if a != ax]
        m = face == f
        cen[m, ax] = (hi[ax] if side else lo[ax]) * 0.995
        cen[m, o[0]] = u[m] * hi[o[0]]
        cen[m, o[1]] = v[m] * hi[o[1]]
    p, t = _small_triangles(rng, n_detail, cen, 0.5 * vox, 4.0 * vox)
    parts.append((p, t, np.arange(n_detail) % 3))
    return _assemble(parts, "sponza_like")


# ---------------------------------------------------------------------------------------------
# C3: San-Miguel-scale (~10M small triangles), level 11
# ---------------------------------------------------------------------------------------------
def san_miguel_like(seed: int = 3, n: int = 2236, layers: int = 2) -> Mesh:
    rng = SplitMix64(seed)
    parts = []
    g = np.linspace(-1.0, 1.0, n)
    x, z = np.meshgrid(g, g, indexing="ij")
    tri = _grid_tris(n, n)
    for k in range(layers):
        ph = rng.uniform(4, 0.0, 6.28)
        y = (-0.45 + 0.9 * k
             + 0.22 * np.sin(2.7 * x + ph[0]) * np.cos(3.3 * z + ph[1])
             + 0.05 * np.sin(17.0 * x + ph[2]) * np.sin(13.0 * z + ph[3])
             + 0.004 * rng.uniform(n * n, -1.0, 1.0).reshape(n, n))
        pos = np.stack([x, y, z], -1).reshape(-1, 3)
        half = len(tri) // 2
        parts.append((pos, tri, np.where(np.arange(len(tri)) < half, 2 * k, 2 * k + 1)))
    mesh = _assemble(parts, "san_miguel_like")
    mesh.positions = np.clip(mesh.positions, -1.0, 1.0)
    return mesh


# ---------------------------------------------------------------------------------------------
# C4: Living-Room-scale (huge wall/floor triangles + furniture + small clutter), level 12
# ---------------------------------------------------------------------------------------------
def living_room_like(seed: int = 4, n_boxes: int = 40, n_small: int = 200000, level: int = 12) -> Mesh:
    rng = SplitMix64(seed)
    parts = []
    c, t = _box((-0.99, -0.99, -0.99), (0.99, 0.99, 0.99), inward=True)  # 12 triangles, ~8.4e6 pixels each at L=12
    parts.append((c, t, 3))
    for k in range(n_boxes):  # furniture: 40 boxes standing on the floor = 480 large triangles
        sz = rng.uniform(3, 0.08, 0.35)
        cx, cz = rng.uniform(2, -0.85, 0.85)
        lo = np.array([cx - sz[0] / 2, -0.99 + 0.004, cz - sz[2] / 2])
        hi = np.array([cx + sz[0] / 2, -0.99 + 0.004 + sz[1], cz + sz[2] / 2])
        lo, hi = np.clip(lo, -0.985, 0.985), np.clip(hi, -0.985, 0.985)
        c, t = _box(lo, hi)
        parts.append((c, t, k % 3))
    vox = 2.0 / (1 << level)
    cen = rng.uniform(3 * n_small, -0.95, 0.95).reshape(n_small, 3)
    p, t = _small_triangles(rng, n_small, cen, 1.0 * vox, 6.0 * vox)
    parts.append((p, t, np.arange(n_small) % 3))
    return _assemble(parts, "living_room_like")


# ---------------------------------------------------------------------------------------------
# C5: dense surface for level 14 (octant-sharded)
# ---------------------------------------------------------------------------------------------
def dense_surface(seed: int = 5, n: int = 2048, layers: int = 1) -> Mesh:
    """Stacked wavy sheets: a dense surface whose leaf count scales with layers * res^2.  One sheet at level
    14 gives ~2.8e8 leaves and ~7.5e8 node words, which still fits the reference's 30-bit child pointers
    (octree.glsl:110) when the octant subtrees are gathered into one buffer."""
    rng = SplitMix64(seed)
    g = np.linspace(-0.999, 0.999, n)
    x, z = np.meshgrid(g, g, indexing="ij")
    tri = _grid_tris(n, n)
    parts = []
    for k in range(layers):
        ph = rng.uniform(2, 0.0, 6.28)
        y0 = -0.7 + 1.4 * (k + 0.5) / layers
        y = y0 + 0.12 * np.sin(5.1 * x + ph[0]) * np.cos(4.3 * z + ph[1])
        parts.append((np.stack([x, y, z], -1).reshape(-1, 3), tri, k % 4))
    return _assemble(parts, "dense_surface")


# ---------------------------------------------------------------------------------------------
# small random meshes for property tests
# ---------------------------------------------------------------------------------------------
def random_soup(n_tri: int, seed: int, size_lo: float = 0.005, size_hi: float = 0.8, n_mat: int = 3) -> Mesh:
    rng = SplitMix64(seed)
    cen = rng.uniform(3 * n_tri, -0.9, 0.9).reshape(n_tri, 3)
    p, t = _small_triangles(rng, n_tri, cen, size_lo, size_hi)
    return _assemble([(p, t, np.arange(n_tri) % n_mat)], f"soup{n_tri}_{seed}")


def procedural_textures(seed: int = 7):
    """Three small sRGB RGBA images: 64x64 opaque noise, 32(w)x16(h) with alpha holes (alpha-tested like the foliage
    textures voxelizer.frag:29-30 exists for), 20x12 opaque (not a power of two: uneven blits in the mip chain)."""
    rng = SplitMix64(seed)

    def noise(h, w):
        return (rng.u64(h * w * 4) >> np.uint64(56)).astype(np.uint8).reshape(h, w, 4)

    a = noise(64, 64)
    a[..., 3] = 255
    b = noise(16, 32)
    yy, xx = np.mgrid[0:16, 0:32]
    b[..., 3] = np.where(((xx // 4) + (yy // 4)) % 2 == 0, 255, np.where(xx % 3 == 0, 120, 10)).astype(np.uint8)
    c = noise(12, 20)
    c[..., 3] = 200
    return [a, b, c]


def textured_soup(n_tri: int = 300, seed: int = 11, size_lo: float = 0.01, size_hi: float = 0.5, uv_scale_hi: float = 6.0,
                  big_quads: bool = True) -> Mesh:
    """Triangle soup over 4 materials: untextured, opaque texture, alpha-tested texture, non-power-of-two texture;
    random texture coordinates whose scale spans magnification to several mip levels, plus (big_quads) two wall-sized
    quads so the large-triangle path samples textures too."""
    rng = SplitMix64(seed)
    cen = rng.uniform(3 * n_tri, -0.9, 0.9).reshape(n_tri, 3)
    pos, tri = _small_triangles(rng, n_tri, cen, size_lo, size_hi)
    scale = np.exp(rng.uniform(n_tri, np.log(0.05), np.log(uv_scale_hi)))
    uv = (rng.uniform(6 * n_tri, -1.0, 1.0).reshape(n_tri, 3, 2) * scale[:, None, None] + rng.uniform(2 * n_tri, -2, 2).reshape(n_tri, 1, 2))
    uv = uv.reshape(-1, 2)
    mat = np.arange(n_tri) % 4
    parts_pos, parts_tri, parts_mat, parts_uv = [pos], [tri], [mat], [uv]
    if big_quads:
        for k, (m, z) in enumerate(((1, -0.31), (2, 0.47), (3, 0.13))):
            q = np.array([[-0.8, -0.7, z], [0.75, -0.72, z + 0.2], [0.8, 0.77, z + 0.25], [-0.78, 0.7, z + 0.05]])
            if k == 1:
                q = q[:, [2, 0, 1]]
            quv = np.array([[0.0, 0.0], [3.0 + k, 0.2], [3.3 + k, 2.5], [-0.1, 2.2]])
            parts_pos.append(q), parts_tri.append(np.array([[0, 1, 2], [0, 2, 3]]) + sum(len(p) for p in parts_pos[:-1]))
            parts_mat.append(np.array([m, m])), parts_uv.append(quv)
    P = np.concatenate(parts_pos).astype(np.float32)
    T = np.concatenate(parts_tri)
    M = np.concatenate(parts_mat)
    UV = np.concatenate(parts_uv).astype(np.float32)
    idx, draws, first = [], [], 0
    for m in (1, 0, 2, 3):  # any fixed draw order (the reference sorts by size, Scene.cpp:124-132)
        t = T[M == m].reshape(-1)
        idx.append(t)
        draws.append((first, len(t), 0xFFFFFFFF if m == 0 else m - 1, pack_albedo(ALBEDOS[m])))
        first += len(t)
    return Mesh(P, np.concatenate(idx).astype(np.uint32), np.array(draws, dtype=DRAW_DTYPE), f"texsoup{n_tri}_{seed}",
                texcoords=UV, textures=procedural_textures(seed + 1))


CONFIGS = {
    "C1": dict(gen=heightfield, level=8, mode="center"),
    "C2": dict(gen=sponza_like, level=10, mode="center"),
    "C3": dict(gen=san_miguel_like, level=11, mode="conservative"),
    "C4": dict(gen=living_room_like, level=12, mode="conservative"),
    "C5": dict(gen=dense_surface, level=14, mode="center"),  # 3*14 Morton bits + 24 colour bits > 64: octant shards
}
