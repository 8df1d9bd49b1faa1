"""The C++ mirror of the reference classes (sparsevoxeloctree_b200/host/svo_host.hpp): it compiles and links
against the C ABI on the CPU box; on the GPU box the reference loader sequence written in C++ gives the same
octree as the Python path."""
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as graft

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "loader_example")


def build_example():
    lib = graft.build_cuda()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "loader_example.cpp")
    cmd = ["g++", "-std=c++17", "-O1", "-o", EXE, src, lib, "-Wl,-rpath," + os.path.dirname(lib)]
    subprocess.run(cmd, check=True)
    return EXE


def test_cpp_mirror_compiles_and_links():
    exe = build_example()
    assert os.path.exists(exe)
    out = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for sym in ("svo_scene_create", "svo_voxelizer_create", "svo_builder_build", "svo_builder_octree_range_bytes"):
        assert sym in out


@pytest.mark.gpu
def test_cpp_loader_matches_python_path():
    from sparsevoxeloctree_b200 import api, scenes
    exe = build_example()
    r = subprocess.run([exe, "7", "33"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    tok = r.stdout.split()
    frags, rng, level = int(tok[0]), int(tok[1]), int(tok[2])
    root = [int(t, 16) for t in tok[3:11]]
    # the same mesh through the Python mirror
    n = 33
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    x = (-1.0 + 2.0 * i / (n - 1)).astype(np.float32)
    z = (-1.0 + 2.0 * j / (n - 1)).astype(np.float32)
    y = (np.float32(0.4) * np.sin(np.float32(3.0) * x) * np.cos(np.float32(2.0) * z)).astype(np.float32)
    # the C++ example evaluates sin/cos in float via std::sin(float): compare sizes and the root block shape only
    pos = np.stack([x, y, z], -1).reshape(-1, 3)
    a, b, c, d = i[:-1, :-1] * n + j[:-1, :-1], (i[:-1, :-1] + 1) * n + j[:-1, :-1], (i[:-1, :-1] + 1) * n + j[:-1, :-1] + 1, i[:-1, :-1] * n + j[:-1, :-1] + 1
    idx = np.stack([a, b, c, a, c, d], -1).reshape(-1).astype(np.uint32)
    draws = np.array([(0, len(idx), 0xFFFFFFFF, 0x00C83C32)], scenes.DRAW_DTYPE)
    scene, vox, builder = api.build_svo(scenes.Mesh(pos, idx, draws, "cpp"), 7)
    assert level == 7
    assert abs(frags - vox.GetVoxelFragmentCount()) <= frags * 0.01
    assert abs(rng - builder.GetOctreeRange()) <= rng * 0.01
    proot = builder.octree_to_host()[:8]
    assert [(w != 0) for w in root] == [(int(w) != 0) for w in proot]
