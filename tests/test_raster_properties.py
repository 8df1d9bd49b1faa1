"""Properties every conformant Vulkan rasterizer has, checked on the oracle's pinned rasterizer arithmetic.

The fixed-function stage is the one part of the path the reference does not define (the driver does), so it cannot
be pinned to reference outputs; what CAN be checked is that the pinned rules obey the
specification's own invariants, with predicates written independently of the oracle's formulation (exact integer
geometry on the snapped 1/256-pixel vertices):
  * centre sampling + top-left rule: two triangles sharing an edge never both cover a pixel, and together cover
    every pixel centre strictly inside their union (watertight, no double hits) -- Vulkan spec, basic polygon
    rasterization;
  * conservative overestimation with extra size 0 (Voxelizer.cpp:120-127): exactly the pixel squares that intersect
    the triangle as closed sets (VK_EXT_conservative_rasterization), and a superset of the centre-sampled set;
  * coverage does not depend on vertex order / winding (CULL_NONE, Voxelizer.cpp:117-118) nor on translation by
    whole pixels.
The CUDA path is bit-identical to the oracle (GPU parity tests), so these properties carry over."""
import itertools

import numpy as np
import pytest

from oracle import oracle

L = 6
RES = 1 << L


def ndc(w):  # window coordinate (pixels) -> the NDC value whose viewport transform gives it back exactly
    return np.float32(w / RES * 2.0 - 1.0)


def tri(pts, z=0.25):
    """z-facing triangle (axis 2: screen x,y = world x,y) from window-space points."""
    return [np.array([ndc(x), ndc(y), np.float32(z)], np.float32) for x, y in pts]


def covered(pts, mode):
    p = tri(pts)
    px, py, _ = oracle.debug_raster_pixels(p[0], p[1], p[2], L, mode)
    return set(zip(px.tolist(), py.tolist()))


def snapped(pts):
    p = tri(pts)
    a, xy = oracle.debug_tri_setup(p[0], p[1], p[2], L)
    assert a[0] == 2
    return [(int(x), int(y)) for x, y in xy]  # 1/256 pixel units


def orient(a, b, c):
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])


def point_in_triangle_closed(p, t):
    d = [orient(t[i], t[(i + 1) % 3], p) for i in range(3)]
    return all(x >= 0 for x in d) or all(x <= 0 for x in d)


def on_segment(a, b, p):
    return orient(a, b, p) == 0 and min(a[0], b[0]) <= p[0] <= max(a[0], b[0]) and min(a[1], b[1]) <= p[1] <= max(a[1], b[1])


def segments_intersect(a, b, c, d):
    o1, o2, o3, o4 = orient(a, b, c), orient(a, b, d), orient(c, d, a), orient(c, d, b)
    if ((o1 > 0) != (o2 > 0)) and ((o3 > 0) != (o4 > 0)) and 0 not in (o1, o2, o3, o4):
        return True
    return on_segment(a, b, c) or on_segment(a, b, d) or on_segment(c, d, a) or on_segment(c, d, b)


def square_touches_triangle(px, py, t):
    """Closed pixel square [px,px+1] x [py,py+1] (in 1/256 units) vs closed triangle t: independent of the SAT form."""
    c = [(256 * px, 256 * py), (256 * px + 256, 256 * py), (256 * px + 256, 256 * py + 256), (256 * px, 256 * py + 256)]
    if any(c[0][0] <= v[0] <= c[2][0] and c[0][1] <= v[1] <= c[2][1] for v in t):
        return True
    if orient(*t) != 0 and any(point_in_triangle_closed(q, t) for q in c):
        return True
    return any(segments_intersect(t[i], t[(i + 1) % 3], c[k], c[(k + 1) % 4]) for i in range(3) for k in range(4))


def lattice_points(rng, n, step=0.5, lo=2.0, hi=RES - 2.0):
    return [tuple(float(v) for v in np.round(rng.uniform(lo, hi, 2) / step) * step) for _ in range(n)]


def test_shared_edge_is_covered_exactly_once():
    rng = np.random.default_rng(11)
    checked = 0
    for _ in range(300):
        # convex quad: four points around a centre, sorted by angle; half-pixel lattice so that edges and the diagonal
        # pass exactly through pixel centres all the time
        c = rng.uniform(12, RES - 12, 2)
        ang = np.sort(rng.uniform(0, 2 * np.pi, 4))
        rad = rng.uniform(3, 10, 4)
        q = [tuple(float(v) for v in np.round((c + r * np.array([np.cos(a), np.sin(a)])) * 2) / 2) for a, r in zip(ang, rad)]
        s = snapped(q[:3]) + [snapped([q[0], q[2], q[3]])[2]]
        turns = [orient(s[i], s[(i + 1) % 4], s[(i + 2) % 4]) for i in range(4)]
        if not (all(t > 0 for t in turns) or all(t < 0 for t in turns)):
            continue  # not strictly convex after snapping
        a, b = covered([q[0], q[1], q[2]], oracle.CENTER), covered([q[0], q[2], q[3]], oracle.CENTER)
        assert not (a & b), "a pixel on the shared edge was produced twice"
        union = a | b
        for px in range(RES):
            for py in range(RES):
                ctr = (256 * px + 128, 256 * py + 128)
                d = [orient(s[i], s[(i + 1) % 4], ctr) for i in range(4)]
                inside = all(x > 0 for x in d) or all(x < 0 for x in d)
                outside = not (all(x >= 0 for x in d) or all(x <= 0 for x in d))
                if inside:
                    assert (px, py) in union, "hole inside the quad (on the diagonal?)"
                if outside:
                    assert (px, py) not in union
        checked += 1
    assert checked > 150


def test_conservative_is_the_closed_square_triangle_intersection():
    rng = np.random.default_rng(12)
    for k in range(250):
        if k % 3 == 0:  # vertices on pixel corners / edges: exact-touch ties
            pts = lattice_points(rng, 3, step=1.0, lo=4, hi=RES - 4)
        elif k % 3 == 1:  # slivers
            a = rng.uniform(6, RES - 6, 2)
            d = rng.uniform(-8, 8, 2)
            pts = [tuple(a), tuple(a + d), tuple(a + d * rng.uniform(0.2, 0.9) + rng.uniform(-0.02, 0.02, 2))]
        else:
            pts = [tuple(v) for v in rng.uniform(3, RES - 3, (3, 2))]
        t = snapped(pts)
        got = covered(pts, oracle.CONSERVATIVE_EXACT)
        xs, ys = [v[0] for v in t], [v[1] for v in t]
        exp = {(px, py) for px in range(max(0, min(xs) // 256 - 1), min(RES, max(xs) // 256 + 2))
               for py in range(max(0, min(ys) // 256 - 1), min(RES, max(ys) // 256 + 2)) if square_touches_triangle(px, py, t)}
        assert got == exp, (pts, sorted(got ^ exp))
        assert covered(pts, oracle.CENTER) <= got


def test_coverage_ignores_vertex_order_and_whole_pixel_translation():
    rng = np.random.default_rng(13)
    for _ in range(60):
        pts = lattice_points(rng, 3, step=0.25, lo=6, hi=RES - 14)
        for mode in (oracle.CENTER, oracle.CONSERVATIVE_EXACT):
            base = covered(pts, mode)
            for perm in itertools.permutations(range(3)):
                assert covered([pts[i] for i in perm], mode) == base
            dx, dy = int(rng.integers(1, 8)), int(rng.integers(1, 8))
            moved = covered([(x + dx, y + dy) for x, y in pts], mode)
            assert moved == {(x + dx, y + dy) for x, y in base}


@pytest.mark.parametrize("mode", [oracle.CENTER, oracle.CONSERVATIVE_EXACT])
def test_depth_is_the_plane_through_the_vertices(mode):
    """Depth at the pixel centre = the triangle's plane (extrapolated outside in conservative mode): compare with an
    exact rational evaluation on the snapped vertices."""
    from fractions import Fraction as Fr
    rng = np.random.default_rng(14)
    for _ in range(40):
        pts = [tuple(v) for v in rng.uniform(4, RES - 4, (3, 2))]
        zs = rng.uniform(-0.9, 0.9, 3).astype(np.float32)
        p = [np.array([ndc(x), ndc(y), z], np.float32) for (x, y), z in zip(pts, zs)]
        a, xy = oracle.debug_tri_setup(p[0], p[1], p[2], L)
        if a[0] != 2:
            continue
        px, py, z = oracle.debug_raster_pixels(p[0], p[1], p[2], L, mode)
        # snapped vertices in the ORIGINAL order (the oracle may swap two of them to normalise the winding)
        f32 = np.float32
        own = [(int(np.rint((v[0] + f32(1)) * f32(0.5) * f32(RES) * f32(256))), int(np.rint((v[1] + f32(1)) * f32(0.5) * f32(RES) * f32(256))))
               for v in p]
        assert sorted(own) == sorted((int(x), int(y)) for x, y in xy)
        s = [(Fr(x), Fr(y)) for x, y in own]
        zf = [Fr(float((np.float32(v) + np.float32(1.0)) * np.float32(0.5))) for v in zs]  # depth = (z+1)/2 in fp32
        det = (s[1][0] - s[0][0]) * (s[2][1] - s[0][1]) - (s[2][0] - s[0][0]) * (s[1][1] - s[0][1])
        if det == 0:
            continue
        for x, y, zz in zip(px.tolist(), py.tolist(), z.tolist()):
            cx, cy = Fr(256 * x + 128), Fr(256 * y + 128)
            l1 = ((cx - s[0][0]) * (s[2][1] - s[0][1]) - (s[2][0] - s[0][0]) * (cy - s[0][1])) / det
            l2 = ((s[1][0] - s[0][0]) * (cy - s[0][1]) - (cx - s[0][0]) * (s[1][1] - s[0][1])) / det
            exact = zf[0] + l1 * (zf[1] - zf[0]) + l2 * (zf[2] - zf[0])
            assert abs(float(exact) - zz) < 1e-12
