"""Scene ingestion: Wavefront OBJ + MTL -> the mesh hand-off of the build path.

Restates what Scene::load_meshes does with tinyobjloader's output (src/Scene.cpp:36-143):
  * faces are triangulated (tinyobj::LoadObj's default); polygons as a fan around their first vertex (tinyobj
    ear-clips, dep/tiny_obj_loader.h:1394-1560: the same surface for convex polygons, possibly another diagonal);
  * one mesh per MATERIAL; a vertex is (position, (u, 1 - v)), (0, 0) without a texture coordinate (Scene.cpp:68-80);
  * positions are normalised to [-1,1]^3 in fp32: centre of the bounding box of the referenced vertices, largest
    half-extent scaled to 1 (Scene.cpp:90-99);
  * albedo = Kd, texture = map_Kd; texture ids are handed out in material order, skipping materials without faces
    (Scene.cpp:101-121); empty meshes are dropped and the rest sorted by size, largest first (Scene.cpp:123-132);
  * one draw per mesh: {first_index, index_count, texture_id, packUnorm4x8(vec4(albedo, 0))} (Scene.cpp:157-170).
The reference then de-duplicates vertices and reorders triangles inside each draw with meshoptimizer
(Scene.cpp:172-197); that changes neither the triangle set nor the draw a triangle belongs to and is not done here.
Images are decoded with Pillow to RGBA8 (stbi_load(..., 4), Scene.cpp:247) when it is installed.
"""
from __future__ import annotations

import os

import numpy as np

from .scenes import DRAW_DTYPE, Mesh


def _parse_mtl(path):
    mats, cur = {}, None
    if not os.path.exists(path):
        return mats
    with open(path, "r", errors="replace") as f:
        for line in f:
            t = line.split("#", 1)[0].split()
            if not t:
                continue
            if t[0] == "newmtl":
                cur = {"Kd": (0.0, 0.0, 0.0), "map_Kd": "", "has_kd": False}  # InitMaterial (dep/tiny_obj_loader.h:1299-1325)
                mats[" ".join(t[1:])] = cur
            elif cur is not None and t[0] == "Kd" and len(t) >= 4:
                cur["Kd"], cur["has_kd"] = tuple(float(x) for x in t[1:4]), True
            elif cur is not None and t[0] == "map_Kd" and len(t) >= 2:
                cur["map_Kd"] = t[-1]  # options (-s, -o, ...) precede the file name
                if not cur["has_kd"]:  # default diffuse when only a map is given (dep/tiny_obj_loader.h:1951-1962)
                    cur["Kd"] = (0.6, 0.6, 0.6)
    return mats


def pack_unorm4x8_rgb(rgb) -> int:
    """glm::packUnorm4x8(vec4(albedo, 0)) (Scene.cpp:164): round(clamp(c,0,1)*255), R in bits 0-7."""
    v = np.rint(np.clip(np.asarray(rgb, np.float32), 0, 1) * np.float32(255)).astype(np.uint32)
    return int(v[0] | (v[1] << 8) | (v[2] << 16))


def load_obj(filename: str, load_images: bool = True) -> Mesh:
    base_dir = os.path.dirname(os.path.abspath(filename))
    pos, uv = [], []
    materials, mat_order = {}, []
    faces = {}  # material name -> list of (vi, ti) triples
    cur = None
    with open(filename, "r", errors="replace") as f:
        for line in f:
            t = line.split("#", 1)[0].split()
            if not t:
                continue
            if t[0] == "v":
                pos.append([float(x) for x in t[1:4]])
            elif t[0] == "vt":
                uv.append([float(t[1]), float(t[2]) if len(t) > 2 else 0.0])
            elif t[0] == "mtllib":
                for k, v in _parse_mtl(os.path.join(base_dir, " ".join(t[1:]))).items():
                    if k not in materials:
                        materials[k] = v
                        mat_order.append(k)
            elif t[0] == "usemtl":
                cur = " ".join(t[1:])
            elif t[0] == "f":
                corners = []
                for c in t[1:]:
                    p = c.split("/")
                    vi = int(p[0])
                    vi = vi - 1 if vi > 0 else len(pos) + vi
                    ti = -1
                    if len(p) > 1 and p[1]:
                        ti = int(p[1])
                        ti = ti - 1 if ti > 0 else len(uv) + ti
                    corners.append((vi, ti))
                lst = faces.setdefault(cur, [])
                for k in range(1, len(corners) - 1):
                    lst.append((corners[0], corners[k], corners[k + 1]))
    if not materials:
        raise ValueError("No material found")  # Scene.cpp:48-51
    if any(m not in materials for m in faces):
        raise ValueError("a face uses an undefined material (the reference indexes meshes by material id, Scene.cpp:69)")
    P = np.asarray(pos, np.float32).reshape(-1, 3)
    UV = np.asarray(uv, np.float32).reshape(-1, 2)
    meshes = []  # (material index, positions [n,3], texcoords [n,2])
    for mi, name in enumerate(mat_order):
        tris = faces.get(name, [])
        if not tris:
            continue
        vi = np.array([[c[0] for c in tri] for tri in tris], np.int64).reshape(-1)
        ti = np.array([[c[1] for c in tri] for tri in tris], np.int64).reshape(-1)
        tc = np.zeros((len(vi), 2), np.float32)
        has = ti >= 0
        tc[has, 0] = UV[ti[has], 0]
        tc[has, 1] = np.float32(1.0) - UV[ti[has], 1]
        meshes.append((mi, P[vi], tc))
    if not meshes:
        raise ValueError("Empty mesh")  # Scene.cpp:134-137
    allp = np.concatenate([m[1] for m in meshes])
    pmin, pmax = allp.min(axis=0), allp.max(axis=0)
    extent = np.float32(max(pmax - pmin)) * np.float32(0.5)
    inv_extent = np.float32(1.0) / extent
    center = (pmax + pmin) * np.float32(0.5)
    tex_ids, tex_files = {}, []
    for mi, _, _ in meshes:  # texture ids in material order over non-empty meshes (Scene.cpp:101-121)
        name = materials[mat_order[mi]]["map_Kd"]
        if name and name not in tex_ids:
            tex_ids[name] = len(tex_files)
            tex_files.append(os.path.join(base_dir, name.replace("\\", "/")))
    meshes.sort(key=lambda m: -len(m[1]))  # Scene.cpp:128-130 (std::sort: order of equal sizes unspecified; stable here)
    positions, texcoords, draws, first = [], [], [], 0
    for mi, p, tc in meshes:
        mat = materials[mat_order[mi]]
        positions.append(((p - center) * inv_extent).astype(np.float32))
        texcoords.append(tc)
        draws.append((first, len(p), tex_ids.get(mat["map_Kd"], 0xFFFFFFFF) if mat["map_Kd"] else 0xFFFFFFFF,
                      pack_unorm4x8_rgb(mat["Kd"])))
        first += len(p)
    textures = None
    if tex_files and load_images:
        from PIL import Image  # stbi_load(path, ..., 4)
        textures = [np.asarray(Image.open(p).convert("RGBA"), dtype=np.uint8) for p in tex_files]
    mesh = Mesh(np.concatenate(positions), np.arange(first, dtype=np.uint32), np.array(draws, dtype=DRAW_DTYPE),
                os.path.basename(filename), texcoords=np.concatenate(texcoords), textures=textures)
    mesh.texture_files = tex_files
    if tex_files and textures is None:  # images not loaded: fall back to the albedo like a failed load would not -- be explicit
        mesh.draws["texture_id"] = 0xFFFFFFFF
    return mesh
