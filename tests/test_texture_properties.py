"""The pinned texture arithmetic against an independent float64 statement of the Vulkan
specification's ideal formulas (texel filtering, scale factor / LOD, sRGB transfer function): same mip chain to
+-1 code, same filtered colour to +-2/255 -- i.e. the pinning only fixes rounding, it does not change the operation
the reference asks its driver for (voxelizer.frag:27-36, Scene.cpp:262-296,409-411)."""
import numpy as np

from oracle import oracle
from sparsevoxeloctree_b200 import scenes

L = 7
RES = 1 << L


def srgb_decode(c):
    x = c / 255.0
    return np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)


def srgb_encode(l):
    return np.where(l <= 0.0031308, l * 12.92, 1.055 * np.maximum(l, 1e-12) ** (1 / 2.4) - 0.055) * 255.0


def ideal_mips(img):
    """vkCmdBlitImage(LINEAR) chain in float64: linear-space box / tent filter, re-encoded and rounded per level."""
    levels = [img.astype(np.float64)]
    h, w = img.shape[:2]
    while w > 1 or h > 1:
        src = levels[-1]
        lin = np.concatenate([srgb_decode(src[..., :3]), src[..., 3:] / 255.0], axis=-1)
        hs, ws = src.shape[:2]
        wd, hd = max(w // 2, 1), max(h // 2, 1)
        out = np.zeros((hd, wd, 4))
        for j in range(hd):
            for i in range(wd):
                u, v = (i + 0.5) * ws / wd - 0.5, (j + 0.5) * hs / hd - 0.5
                i0, j0 = int(np.floor(u)), int(np.floor(v))
                a, b = u - i0, v - j0
                px = lambda ii, jj: lin[min(max(jj, 0), hs - 1), min(max(ii, 0), ws - 1)]
                out[j, i] = (1 - b) * ((1 - a) * px(i0, j0) + a * px(i0 + 1, j0)) + b * ((1 - a) * px(i0, j0 + 1) + a * px(i0 + 1, j0 + 1))
        enc = np.concatenate([srgb_encode(out[..., :3]), out[..., 3:] * 255.0], axis=-1)
        levels.append(np.rint(np.clip(enc, 0, 255)))
        w, h = wd, hd
    return levels


def test_mip_chain_close_to_ideal():
    for t in scenes.procedural_textures(3) + [np.random.default_rng(1).integers(0, 256, (9, 5, 4), dtype=np.uint8)]:
        ts = oracle.TexSet([t])
        ref = ideal_mips(t)
        assert ts.level_count(0) == len(ref)
        for lv in range(1, len(ref)):
            got = ts.level(0, lv).astype(np.float64)
            # each level is built from the previous ROUNDED level: compare against the ideal filter of the oracle's own
            # previous level, so that roundings do not accumulate in the comparison
            prev = ideal_mips(ts.level(0, lv - 1))[1]
            assert np.abs(got - prev).max() <= 1.0, lv


def ideal_sample(levels, uvmap, px, py):
    """texture() at the centre of pixel (px,py) for an affine uv map: float64 trilinear, REPEAT, exact LOD."""
    (u0, v0, dudx, dudy, dvdx, dvdy, x0, y0) = uvmap
    cx, cy = px + 0.5, py + 0.5
    u = u0 + dudx * (cx - x0) + dudy * (cy - y0)
    v = v0 + dvdx * (cx - x0) + dvdy * (cy - y0)
    H, W = levels[0].shape[:2]
    rho = max(np.hypot(dudx * W, dvdx * H), np.hypot(dudy * W, dvdy * H))
    lam = np.log2(rho) if rho > 0 else -np.inf
    q = len(levels) - 1
    lam = min(max(lam, 0.0), q)
    hi = int(np.floor(lam))
    lo = min(hi + 1, q)
    d = lam - hi

    def bil(level):
        img = levels[level]
        h, w = img.shape[:2]
        lin = np.concatenate([srgb_decode(img[..., :3]), img[..., 3:] / 255.0], axis=-1)
        U, V = u * w - 0.5, v * h - 0.5
        i0, j0 = int(np.floor(U)), int(np.floor(V))
        a, b = U - i0, V - j0
        px_ = lambda ii, jj: lin[jj % h, ii % w]
        return (1 - b) * ((1 - a) * px_(i0, j0) + a * px_(i0 + 1, j0)) + b * ((1 - a) * px_(i0, j0 + 1) + a * px_(i0 + 1, j0 + 1))

    return (1 - d) * bil(hi) + d * bil(lo)


def test_sampling_close_to_ideal_trilinear():
    rng = np.random.default_rng(5)
    tex = scenes.procedural_textures(9)
    ts = oracle.TexSet(tex)
    levels = [[ts.level(t, lv).astype(np.float64) for lv in range(ts.level_count(t))] for t in range(len(tex))]
    checked = discarded = 0
    for k in range(400):
        t = k % len(tex)
        # z-facing triangle (axis 2) with window-space vertices and arbitrary uv; uv scale spans magnification .. deep mips
        w = rng.uniform(8, RES - 8, (3, 2))
        e1, e2 = w[1] - w[0], w[2] - w[0]
        if abs(e1[0] * e2[1] - e1[1] * e2[0]) < 20:
            continue
        scale = np.exp(rng.uniform(np.log(0.02), np.log(3.0)))
        uv = (rng.uniform(-1, 1, (3, 2)) * scale + rng.uniform(-2, 2, 2)).astype(np.float32)
        pos = np.array([[x / RES * 2 - 1, y / RES * 2 - 1, 0.3] for x, y in w], np.float32)
        # the affine uv map through the fp32 window coordinates the oracle uses
        xy = ((pos[:, :2].astype(np.float32) + np.float32(1)) * np.float32(0.5) * np.float32(RES)).astype(np.float64)
        A = np.array([[xy[1, 0] - xy[0, 0], xy[1, 1] - xy[0, 1]], [xy[2, 0] - xy[0, 0], xy[2, 1] - xy[0, 1]]])
        gu = np.linalg.solve(A, (uv[1:, 0] - uv[0, 0]).astype(np.float64))
        gv = np.linalg.solve(A, (uv[1:, 1] - uv[0, 1]).astype(np.float64))
        um = (float(uv[0, 0]), float(uv[0, 1]), gu[0], gu[1], gv[0], gv[1], xy[0, 0], xy[0, 1])
        px, py = int(w[:, 0].mean()), int(w[:, 1].mean())
        got, lod = ts.sample(t, pos, uv, L, px, py)
        ref = ideal_sample(levels[t], um, px, py)
        if abs(ref[3] - 0.5) < 0.01:
            continue  # alpha test on the fence
        assert (got is None) == (ref[3] < 0.5), (k, ref[3])
        if got is None:
            discarded += 1
            continue
        rgb = np.array([got & 255, (got >> 8) & 255, (got >> 16) & 255], np.float64)
        # LOD fraction is floored to 1/256 and weights are fp32: allow 2 codes
        assert np.abs(rgb - ref[:3] * 255.0).max() <= 2.0, (k, rgb, ref[:3] * 255, lod)
        checked += 1
    assert checked > 200 and discarded > 10
