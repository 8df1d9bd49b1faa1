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


def test_obj_ingestion_cpp_equals_python():
    """Scene::load_meshes restated twice (svo_host::LoadObj, obj_loader.load_obj): identical vertices (bit for bit),
    draws and texture list on a file with quads, a polygon, negative indices, shared/unused/map-only materials."""
    from sparsevoxeloctree_b200 import obj_loader
    exe = build_example()
    path = os.path.join(ROOT, "tests", "assets", "two_boxes.obj")
    r = subprocess.run([exe, "--obj", path], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().split("\n")
    nv, nd, nt = (int(x) for x in lines[0].split())
    draws = [tuple(int(x) for x in ln.split()[1:]) for ln in lines if ln.startswith("d ")]
    verts = np.array([[int(x, 16) for x in ln.split()[1:]] for ln in lines if ln.startswith("v ")], np.uint32).view(np.float32)
    tex = [ln[2:] for ln in lines if ln.startswith("t ")]
    m = obj_loader.load_obj(path)
    assert (nv, nd, nt) == (len(m.positions), len(m.draws), len(m.texture_files))
    assert draws == [tuple(int(x) for x in d) for d in m.draws]
    assert (verts[:, :3].view(np.uint32) == m.positions.view(np.uint32)).all()
    assert (verts[:, 3:].view(np.uint32) == m.texcoords.view(np.uint32)).all()
    assert tex == m.texture_files
    # what the reference guarantees downstream: positions in [-1,1]^3 with the largest extent exactly [-1,1] (Scene.cpp:90-99)
    assert m.positions.min() == -1.0 and m.positions.max() == 1.0
    assert m.draws["index_count"].tolist() == sorted(m.draws["index_count"].tolist(), reverse=True)
    assert m.textures[0].shape == (8, 8, 4)


@pytest.mark.gpu
def test_obj_scene_builds_like_the_oracle():
    from sparsevoxeloctree_b200 import api, obj_loader
    from tests.parity import check_against_oracle
    m = obj_loader.load_obj(os.path.join(ROOT, "tests", "assets", "two_boxes.obj"))
    info = check_against_oracle(api.get_library(), m, 7, api.CONSERVATIVE_EXACT)
    assert info["fragments"] > 10000


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_cpp_loader_matches_python_path(tmp_path, mode):
    """The reference's loader sequence written in C++ (tests/cpp/loader_example.cpp over host/svo_host.hpp) and the
    Python mirror, fed the same mesh bytes: identical fragment count, range and node buffer, word for word."""
    from sparsevoxeloctree_b200 import api, scenes
    exe = build_example()
    mesh = scenes.random_soup(700, 61, 0.004, 0.9)
    level = 8
    fin, fout = tmp_path / "mesh.bin", tmp_path / "tree.bin"
    with open(fin, "wb") as f:
        np.array([len(mesh.positions), len(mesh.indices), len(mesh.draws)], np.uint64).tofile(f)
        np.ascontiguousarray(mesh.positions, np.float32).tofile(f)
        np.ascontiguousarray(mesh.indices, np.uint32).tofile(f)
        np.ascontiguousarray(mesh.draws).tofile(f)
    r = subprocess.run([exe, "--mesh", str(fin), str(level), str(mode), str(fout)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    hdr = np.fromfile(fout, np.uint64, 2)
    words = np.fromfile(fout, np.uint32, offset=16)
    scene, vox, builder = api.build_svo(mesh, level, mode)
    assert int(hdr[0]) == vox.GetVoxelFragmentCount() and int(hdr[1]) == builder.GetOctreeRange() == 4 * len(words)
    assert (words == builder.octree_to_host()).all()


@pytest.mark.gpu
def test_cpp_host_builds_sharded_over_all_gpus(tmp_path):
    """A C++ host calls svo_build_sharded (one process, every GPU of the box -- or GPU 0 twice when there is only one):
    the stitched tree is canonically the single-GPU tree."""
    from sparsevoxeloctree_b200 import api, scenes
    from tests.parity import assert_same_tree
    exe = build_example()
    lib = api.get_library()
    n = lib.dll.svo_device_count()
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    devices = ",".join(str(k % n) for k in range(world))
    mesh = scenes.random_soup(900, 62, 0.004, 0.9)
    level = 9
    fin, fout = tmp_path / "mesh.bin", tmp_path / "tree.bin"
    with open(fin, "wb") as f:
        np.array([len(mesh.positions), len(mesh.indices), len(mesh.draws)], np.uint64).tofile(f)
        np.ascontiguousarray(mesh.positions, np.float32).tofile(f)
        np.ascontiguousarray(mesh.indices, np.uint32).tofile(f)
        np.ascontiguousarray(mesh.draws).tofile(f)
    env = dict(os.environ, SVO_DEVICES=devices)
    r = subprocess.run([exe, "--mesh", str(fin), str(level), "1", str(fout)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    hdr = np.fromfile(fout, np.uint64, 2)
    words = np.fromfile(fout, np.uint32, offset=16)
    scene, vox, builder = api.build_svo(mesh, level, 1)
    assert int(hdr[0]) == vox.GetVoxelFragmentCount() and int(hdr[1]) == builder.GetOctreeRange()
    assert_same_tree(words, builder.octree_to_host(), level)
