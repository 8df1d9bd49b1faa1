"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/svo.h declares.
No compute call is made (there is no GPU on the CPU test box)."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as graft
from sparsevoxeloctree_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    return graft.build_cuda()  # nvcc cross-compiles without a GPU


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "svo.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"SVO_API\s+[\w\s\*]+?\b(svo_\w+)\s*\(", text)


def test_header_declares_the_path():
    names = declared_symbols()
    for must in ("svo_scene_create", "svo_voxelizer_create", "svo_voxelizer_voxelize", "svo_voxelizer_fragment_count",
                 "svo_voxelizer_fragments", "svo_builder_create", "svo_builder_build", "svo_builder_octree_range_bytes",
                 "svo_builder_octree", "svo_sort_u64", "svo_builder_rebase_copy"):
        assert must in names
    assert len(names) == len(set(names))


def test_library_exports_every_declared_symbol(lib_path):
    dll = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(dll, name), f"{name} declared in include/svo.h but not exported"


def test_python_mirror_binds_every_declared_symbol(lib_path):
    bound = {n for n, _, _ in api.SYMBOLS}
    assert bound == set(declared_symbols())
    L = api.Library(lib_path)
    assert b"sm_100a" in L.dll.svo_version()


def test_no_cpu_fallback_in_product_package():
    # the product package must not reference the oracle or the emulation library
    pkg = os.path.join(ROOT, "sparsevoxeloctree_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "libsvo_emu" not in text and "libsvo_oracle" not in text, f


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(FileNotFoundError):
        api.Library(str(tmp_path / "libsvo_b200.so"))


def test_sass_is_sm100(lib_path):
    out = os.popen(f"cuobjdump -lelf {lib_path} 2>/dev/null").read()
    assert "sm_100a" in out
