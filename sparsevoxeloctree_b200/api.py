"""Host-side Python mirror of the reference's SVO construction classes, over the C ABI of include/svo.h.

Class and method names follow the reference (src/Scene.hpp, src/Voxelizer.hpp:43-52,
src/OctreeBuilder.hpp:36-48, src/Octree.hpp:18-33) so that callers -- and the parity tests -- read like
the reference's own loader (src/LoaderThread.cpp:51-89):

    scene     = Scene.Create(mesh)
    voxelizer = Voxelizer.Create(scene, octree_level)
    builder   = OctreeBuilder.Create(voxelizer)
    voxelizer.CmdVoxelize(stream); builder.CmdBuild(stream)
    octree.Update(builder)

All compute happens in libsvo_b200.so (hand-written sm_100a CUDA).  There is no Python or CPU fallback:
if the library is missing or no CUDA device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvo_b200.so")

CENTER, CONSERVATIVE_EXACT, CONSERVATIVE_DILATE = 0, 1, 2
PHASES = ("raster", "sort_hist", "sort_passes", "reduce", "levels", "emit")
# what the same six slots hold after a build on the brick path (svo_builder_build_path() == 1)
BRICK_PHASES = ("raster_small", "small_sort_reduce", "pairs", "bricks", "levels", "emit")


class SvoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"svo error {code}: {msg}")
        self.code = code


class svo_draw(C.Structure):
    _fields_ = [("first_index", C.c_uint32), ("index_count", C.c_uint32), ("texture_id", C.c_uint32),
                ("albedo_rgba8", C.c_uint32)]


class svo_texture(C.Structure):
    _fields_ = [("rgba8", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class svo_mesh(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("position_stride_bytes", C.c_uint32), ("on_device", C.c_uint32),
                ("indices", C.c_void_p), ("n_vertices", C.c_uint64), ("n_indices", C.c_uint64),
                ("draws", C.POINTER(svo_draw)), ("n_draws", C.c_uint32),
                ("n_textures", C.c_uint32), ("textures", C.POINTER(svo_texture)),
                ("texcoords", C.c_void_p), ("texcoord_stride_bytes", C.c_uint32)]


def expand_compact(lib, device: int, d_tables: int, plan, d_dst: int, stream=None):
    """The other half of OctreeBuilder.EmitCompactTo, on the GPU that owns d_dst: flat bricks' blocks and all pointer
    blocks of the two deepest windows, from the tables alone."""
    arr = (C.c_uint64 * 4)(*[int(v) for v in plan])
    lib.check(lib.dll.svo_expand_compact(device, d_tables, arr, d_dst, _stream_ptr(stream)))


class svo_shard(C.Structure):
    _fields_ = [("shard_level", C.c_uint32), ("cube_index", C.c_uint32 * 3)]


# every symbol include/svo.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("svo_last_error", C.c_char_p, []),
    ("svo_version", C.c_char_p, []),
    ("svo_device_count", C.c_int, []),
    ("svo_launch_count", C.c_uint64, []),
    ("svo_scene_create", C.c_int, [C.POINTER(svo_mesh), C.c_int, _P, C.POINTER(_P)]),
    ("svo_scene_destroy", None, [_P]),
    ("svo_scene_triangle_count", C.c_uint64, [_P]),
    ("svo_scene_texture_level", C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(_P)]),
    ("svo_voxelizer_create", C.c_int, [_P, C.c_uint32, C.c_int, C.POINTER(svo_shard), _P, C.POINTER(_P)]),
    ("svo_voxelizer_create_from_fragments", C.c_int, [C.c_int, C.c_uint32, _P, C.c_uint64, C.c_int, _P, C.POINTER(_P)]),
    ("svo_voxelizer_create_windowed", C.c_int, [_P, C.c_uint32, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), _P, C.POINTER(_P)]),
    ("svo_voxelizer_destroy", None, [_P]),
    ("svo_voxelizer_voxelize", C.c_int, [_P, _P]),
    ("svo_voxelizer_level", C.c_uint32, [_P]),
    ("svo_voxelizer_resolution", C.c_uint32, [_P]),
    ("svo_voxelizer_fragment_count", C.c_uint64, [_P]),
    ("svo_voxelizer_fragments", _P, [_P]),
    ("svo_voxelizer_export_reference_fragments", C.c_int, [_P, _P, _P]),
    ("svo_builder_create", C.c_int, [_P, _P, C.POINTER(_P)]),
    ("svo_builder_destroy", None, [_P]),
    ("svo_builder_build", C.c_int, [_P, _P]),
    ("svo_builder_prepare", C.c_int, [_P, _P]),
    ("svo_builder_emit_to", C.c_int, [_P, _P, C.c_uint32, C.c_int, _P]),
    ("svo_builder_compact_bytes", C.c_uint64, [_P]),
    ("svo_builder_push_tables", C.c_int, [_P, C.c_uint32, C.c_int, _P, C.POINTER(C.c_uint64), _P]),
    ("svo_builder_emit_compact_to", C.c_int, [_P, _P, C.c_uint32, C.c_int, _P, C.POINTER(C.c_uint64), _P]),
    ("svo_expand_compact", C.c_int, [C.c_int, _P, C.POINTER(C.c_uint64), _P, _P]),
    ("svo_builder_root_words", C.c_int, [_P, C.POINTER(C.c_uint32), _P]),
    ("svo_builder_top_words", C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), _P]),
    ("svo_builder_level", C.c_uint32, [_P]),
    ("svo_builder_octree_range_bytes", C.c_uint64, [_P]),
    ("svo_builder_octree", _P, [_P]),
    ("svo_builder_leaf_count", C.c_uint64, [_P]),
    ("svo_builder_level_counts", C.c_int, [_P, C.POINTER(C.c_uint64), C.c_uint32]),
    ("svo_builder_rebase_copy", C.c_int, [_P, _P, C.c_uint64, C.c_uint32, _P]),
    ("svo_voxelizer_last_ms", C.c_int, [_P, C.POINTER(C.c_float)]),
    ("svo_builder_last_ms", C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    ("svo_sort_u64", C.c_int, [_P, _P, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, _P]),
    ("svo_debug_force_wide_sort_state", None, [C.c_int]),
    ("svo_debug_profile_passes", None, [C.c_int]),
    ("svo_debug_set_build_path", None, [C.c_int]),
    ("svo_builder_build_path", C.c_int, [_P]),
    ("svo_builder_brick_stats", C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_float)]),
    ("svo_builder_sort_step_ms", C.c_int, [_P, C.POINTER(C.c_float), C.c_uint32]),
    ("svo_octree_raymarch_leaf", C.c_int, [C.c_int, _P, C.c_uint64, _P, _P, _P, _P]),
    ("svo_device_malloc", C.c_int, [C.c_int, C.c_uint64, C.POINTER(_P)]),
    ("svo_device_free", C.c_int, [C.c_int, _P]),
    ("svo_memcpy_h2d", C.c_int, [C.c_int, _P, _P, C.c_uint64, _P]),
    ("svo_memcpy_d2h", C.c_int, [C.c_int, _P, _P, C.c_uint64, _P]),
    ("svo_memcpy_d2d", C.c_int, [C.c_int, _P, _P, C.c_uint64, _P]),
    ("svo_stream_synchronize", C.c_int, [C.c_int, _P]),
    ("svo_ipc_export", C.c_int, [C.c_int, _P, C.c_char_p]),
    ("svo_ipc_open", C.c_int, [C.c_int, C.c_char_p, C.POINTER(_P)]),
    ("svo_ipc_close", C.c_int, [C.c_int, _P]),
    ("svo_build_sharded", C.c_int, [C.POINTER(svo_mesh), C.c_uint32, C.c_int, C.POINTER(C.c_int), C.c_uint32, C.POINTER(_P)]),
    ("svo_sharded_rebuild", C.c_int, [_P]),
    ("svo_sharded_octree", _P, [_P]),
    ("svo_sharded_octree_range_bytes", C.c_uint64, [_P]),
    ("svo_sharded_leaf_count", C.c_uint64, [_P]),
    ("svo_sharded_fragment_count", C.c_uint64, [_P]),
    ("svo_sharded_last_ms", C.c_float, [_P]),
    ("svo_sharded_destroy", None, [_P]),
    ("svo_builder_export_fd", C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(_P), _P]),
    ("svo_external_memory_import_fd", C.c_int, [C.c_int, C.c_int, C.c_uint64, C.POINTER(_P), C.POINTER(_P)]),
    ("svo_external_memory_release", C.c_int, [C.c_int, _P]),
]


class Library:
    """A loaded libsvo_b200.so with typed prototypes."""

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(path)
        for name, res, args in SYMBOLS:
            fn = getattr(self.dll, name)  # AttributeError = ABI symbol missing
            fn.restype, fn.argtypes = res, args

    def check(self, rc: int):
        if rc != 0:
            raise SvoError(rc, self.dll.svo_last_error().decode())

    # --- raw device memory helpers (tests / callers without a CUDA binding) ---
    def malloc(self, nbytes: int, device: int = 0) -> int:
        p = _P()
        self.check(self.dll.svo_device_malloc(device, nbytes, C.byref(p)))
        return p.value or 0

    def free(self, ptr: int, device: int = 0):
        self.check(self.dll.svo_device_free(device, ptr))

    def to_device(self, arr: np.ndarray, device: int = 0, stream: int = 0) -> int:
        arr = np.ascontiguousarray(arr)
        p = self.malloc(max(arr.nbytes, 1), device)
        if arr.nbytes:
            self.check(self.dll.svo_memcpy_h2d(device, p, arr.ctypes.data, arr.nbytes, stream))
            self.check(self.dll.svo_stream_synchronize(device, stream))
        return p

    def to_host(self, ptr: int, dtype, count: int, device: int = 0, stream: int = 0) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        if count:
            self.check(self.dll.svo_memcpy_d2h(device, out.ctypes.data, ptr, out.nbytes, stream))
        return out

    # --- CUDA IPC (multi-process stitch over NVLink) ---
    def ipc_export(self, ptr: int, device: int = 0) -> bytes:
        buf = C.create_string_buffer(64)
        self.check(self.dll.svo_ipc_export(device, ptr, buf))
        return buf.raw

    def ipc_open(self, handle: bytes, device: int = 0) -> int:
        p = _P()
        self.check(self.dll.svo_ipc_open(device, C.create_string_buffer(handle, 64), C.byref(p)))
        return p.value or 0

    def ipc_close(self, ptr: int, device: int = 0):
        self.check(self.dll.svo_ipc_close(device, ptr))

    def sort_u64(self, keys: np.ndarray, begin_bit: int, end_bit: int, device: int = 0) -> np.ndarray:
        """svo_sort_u64 on a host array (test helper)."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        d = self.to_device(keys, device)
        t = self.malloc(max(keys.nbytes, 8), device)
        try:
            self.check(self.dll.svo_sort_u64(d, t, len(keys), begin_bit, end_bit, device, 0))
            return self.to_host(d, np.uint64, len(keys), device)
        finally:
            self.free(d, device)
            self.free(t, device)


_default = None


def get_library() -> Library:
    """The in-tree CUDA library.  Raises when it has not been built -- the product never substitutes anything."""
    global _default
    if _default is None:
        _default = Library(LIB_PATH)
    return _default


def _stream_ptr(stream) -> int:
    if stream is None:
        return 0
    if hasattr(stream, "cuda_stream"):  # torch.cuda.Stream
        return int(stream.cuda_stream)
    return int(stream)


class Scene:
    """Mesh hand-off standing in for the reference Scene (src/Scene.hpp): vertex/index buffers + draw list on the
    device.  `positions` may be [V,3] (tight) or [V,5] (pos+uv, the reference Vertex of src/Scene.cpp:16-19).
    Textured materials (src/Scene.cpp:225-300): `textures` = list of uint8 [H,W,4] sRGB RGBA images (what stbi_load
    returns), `texcoords` = float32 [V,2] (taken from columns 3:5 of a [V,5] vertex array when omitted)."""

    def __init__(self):
        self._h = None

    @staticmethod
    def Create(mesh_or_positions, indices=None, draws=None, device: int = 0, stream=None, lib: Library | None = None,
               texcoords=None, textures=None):
        lib = lib or get_library()
        if indices is None:
            m = mesh_or_positions
            positions, indices, draws = m.positions, m.indices, m.draws
            texcoords = getattr(m, "texcoords", None) if texcoords is None else texcoords
            textures = getattr(m, "textures", None) if textures is None else textures
        else:
            positions = mesh_or_positions
        self = Scene()
        self.lib, self.device = lib, device
        self._positions = np.ascontiguousarray(positions, dtype=np.float32)
        self._indices = np.ascontiguousarray(indices, dtype=np.uint32)
        d = np.ascontiguousarray(draws)
        self._draws = (svo_draw * len(d))(*[svo_draw(int(r["first_index"]), int(r["index_count"]), int(r["texture_id"]),
                                                     int(r["albedo_rgba8"])) for r in d])
        if self._positions.ndim != 2 or self._positions.shape[1] < 3:
            raise ValueError("positions must be [V, >=3] float32")
        m = svo_mesh(self._positions.ctypes.data, self._positions.shape[1] * 4, 0, self._indices.ctypes.data,
                     len(self._positions), len(self._indices), self._draws, len(d))
        if textures:
            self._textures = [np.ascontiguousarray(t, dtype=np.uint8) for t in textures]
            for t in self._textures:
                if t.ndim != 3 or t.shape[2] != 4:
                    raise ValueError("textures must be uint8 [H, W, 4]")
            self._tex_arr = (svo_texture * len(self._textures))(*[svo_texture(t.ctypes.data, t.shape[1], t.shape[0])
                                                                   for t in self._textures])
            m.n_textures, m.textures = len(self._textures), self._tex_arr
            if texcoords is not None:
                self._texcoords = np.ascontiguousarray(texcoords, dtype=np.float32)
                if self._texcoords.shape != (len(self._positions), 2):
                    raise ValueError("texcoords must be [V, 2] float32")
                m.texcoords, m.texcoord_stride_bytes = self._texcoords.ctypes.data, 8
            elif self._positions.shape[1] >= 5:  # the reference's Vertex {vec3 pos; vec2 uv}
                m.texcoords, m.texcoord_stride_bytes = self._positions.ctypes.data + 12, self._positions.shape[1] * 4
        h = _P()
        lib.check(lib.dll.svo_scene_create(C.byref(m), device, _stream_ptr(stream), C.byref(h)))
        self._h = h
        return self

    @staticmethod
    def CreateFromDevice(d_positions: int, stride: int, n_vertices: int, d_indices: int, n_indices: int, draws,
                         device: int = 0, stream=None, lib: Library | None = None):
        """Borrow vertex/index buffers that already live on the device (e.g. torch tensors' data_ptr())."""
        lib = lib or get_library()
        self = Scene()
        self.lib, self.device = lib, device
        d = np.ascontiguousarray(draws)
        self._draws = (svo_draw * len(d))(*[svo_draw(int(r["first_index"]), int(r["index_count"]), int(r["texture_id"]),
                                                     int(r["albedo_rgba8"])) for r in d])
        m = svo_mesh(d_positions, stride, 1, d_indices, n_vertices, n_indices, self._draws, len(d))
        h = _P()
        lib.check(lib.dll.svo_scene_create(C.byref(m), device, _stream_ptr(stream), C.byref(h)))
        self._h = h
        return self

    def GetTriangleCount(self) -> int:
        return int(self.lib.dll.svo_scene_triangle_count(self._h))

    def texture_level_to_host(self, texture: int, level: int) -> np.ndarray:
        """uint8 [H,W,4] copy of one mip level as built on the device (Scene.cpp:290-295)."""
        w, h, d = C.c_uint32(), C.c_uint32(), _P()
        rc = self.lib.dll.svo_scene_texture_level(self._h, texture, level, C.byref(w), C.byref(h), C.byref(d))
        if rc < 0:
            self.lib.check(rc)
        return self.lib.to_host(d.value, np.uint8, w.value * h.value * 4, self.device).reshape(h.value, w.value, 4)

    def texture_level_count(self, texture: int) -> int:
        w, h, d = C.c_uint32(), C.c_uint32(), _P()
        rc = self.lib.dll.svo_scene_texture_level(self._h, texture, 0, C.byref(w), C.byref(h), C.byref(d))
        if rc < 0:
            self.lib.check(rc)
        return rc

    def Destroy(self):
        if self._h:
            self.lib.dll.svo_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.Destroy()
        except Exception:
            pass


class Voxelizer:
    """src/Voxelizer.hpp.  Create() runs the count pass and sizes the fragment list (Voxelizer.cpp:134-165);
    CmdVoxelize() enqueues the fragment emission (Voxelizer.cpp:167-179)."""

    def __init__(self):
        self._h = None

    @staticmethod
    def Create(scene: Scene, octree_level: int, mode: int = CONSERVATIVE_EXACT, shard=None, stream=None):
        self = Voxelizer()
        self.lib, self.device, self._scene = scene.lib, scene.device, scene
        sh = None
        if shard is not None:
            lvl, (cx, cy, cz) = shard
            sh = svo_shard(lvl, (C.c_uint32 * 3)(cx, cy, cz))
        h = _P()
        self.lib.check(self.lib.dll.svo_voxelizer_create(scene._h, octree_level, mode, C.byref(sh) if sh else None,
                                                         _stream_ptr(stream), C.byref(h)))
        self._h = h
        self.shard = shard
        return self

    @staticmethod
    def CreateFromFragments(fragments, octree_level: int, device: int = 0, stream=None, lib: Library | None = None):
        """A pre-voxelized source: `fragments` = uint64 array of morton << 24 | rgb (host), any order."""
        self = Voxelizer()
        self.lib, self.device, self._scene, self.shard = lib or get_library(), device, None, None
        self._ext = np.ascontiguousarray(fragments, dtype=np.uint64)
        h = _P()
        self.lib.check(self.lib.dll.svo_voxelizer_create_from_fragments(device, octree_level, self._ext.ctypes.data, len(self._ext), 0,
                                                                        _stream_ptr(stream), C.byref(h)))
        self._h = h
        return self

    @staticmethod
    def CreateWindowed(scene: Scene, octree_level: int, mode: int, window_lo, window_hi, stream=None):
        """Only the fragments inside the half-open voxel box [window_lo, window_hi), in global coordinates."""
        self = Voxelizer()
        self.lib, self.device, self._scene, self.shard = scene.lib, scene.device, scene, None
        lo, hi = (C.c_uint32 * 3)(*[int(v) for v in window_lo]), (C.c_uint32 * 3)(*[int(v) for v in window_hi])
        h = _P()
        self.lib.check(self.lib.dll.svo_voxelizer_create_windowed(scene._h, octree_level, mode, lo, hi, _stream_ptr(stream), C.byref(h)))
        self._h = h
        return self

    def GetScenePtr(self) -> Scene:
        return self._scene

    def GetLevel(self) -> int:
        return int(self.lib.dll.svo_voxelizer_level(self._h))

    def GetVoxelResolution(self) -> int:
        return int(self.lib.dll.svo_voxelizer_resolution(self._h))

    def GetVoxelFragmentCount(self) -> int:
        return int(self.lib.dll.svo_voxelizer_fragment_count(self._h))

    def GetVoxelFragmentList(self) -> int:
        """Device pointer to GetVoxelFragmentCount() 64-bit fragments: morton << 24 | rgb."""
        return int(self.lib.dll.svo_voxelizer_fragments(self._h) or 0)

    def CmdVoxelize(self, stream=None):
        self.lib.check(self.lib.dll.svo_voxelizer_voxelize(self._h, _stream_ptr(stream)))

    def LastMs(self) -> float:
        ms = C.c_float()
        self.lib.check(self.lib.dll.svo_voxelizer_last_ms(self._h, C.byref(ms)))
        return ms.value

    # --- test / debugging helpers -------------------------------------------------------------------------
    def fragments_to_host(self, stream=None) -> np.ndarray:
        return self.lib.to_host(self.GetVoxelFragmentList(), np.uint64, self.GetVoxelFragmentCount(), self.device,
                                _stream_ptr(stream))

    def reference_fragments_to_host(self, stream=None) -> np.ndarray:
        """The fragment list in the reference's uvec2 packing (voxelizer.frag:40-42): uint32 [F, 2]."""
        n = self.GetVoxelFragmentCount()
        d = self.lib.malloc(max(n * 8, 8), self.device)
        try:
            self.lib.check(self.lib.dll.svo_voxelizer_export_reference_fragments(self._h, d, _stream_ptr(stream)))
            return self.lib.to_host(d, np.uint32, n * 2, self.device, _stream_ptr(stream)).reshape(n, 2)
        finally:
            self.lib.free(d, self.device)

    def Destroy(self):
        if self._h:
            self.lib.dll.svo_voxelizer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.Destroy()
        except Exception:
            pass


class OctreeBuilder:
    """src/OctreeBuilder.hpp.  CmdBuild() enqueues sort + reduce + level build; GetOctreeRange() is the byte
    range of the node buffer (OctreeBuilder.cpp:212-214), GetOctree() the device pointer the tracer binds."""

    def __init__(self):
        self._h = None

    @staticmethod
    def Create(voxelizer: Voxelizer, stream=None):
        self = OctreeBuilder()
        self.lib, self.device, self._vox = voxelizer.lib, voxelizer.device, voxelizer
        h = _P()
        self.lib.check(self.lib.dll.svo_builder_create(voxelizer._h, _stream_ptr(stream), C.byref(h)))
        self._h = h
        return self

    def GetVoxelizerPtr(self) -> Voxelizer:
        return self._vox

    def GetLevel(self) -> int:
        return int(self.lib.dll.svo_builder_level(self._h))

    def CmdBuild(self, stream=None):
        self.lib.check(self.lib.dll.svo_builder_build(self._h, _stream_ptr(stream)))

    def Prepare(self, stream=None):
        """Phase 1 of CmdBuild: everything but the node-word emission; sizes are known afterwards."""
        self.lib.check(self.lib.dll.svo_builder_prepare(self._h, _stream_ptr(stream)))

    def EmitTo(self, d_dst: int, pointer_bias_words: int = 0, skip_root=False, stream=None):
        """Phase 2: write the node words into caller-provided device memory (possibly a peer GPU's).
        skip_root: False / True, or 2 to keep the root block and the depth-1 blocks aside (TopWords)."""
        self.lib.check(self.lib.dll.svo_builder_emit_to(self._h, d_dst, pointer_bias_words, int(skip_root), _stream_ptr(stream)))

    def CompactBytes(self) -> int:
        """Bytes of the per-brick tables of the compact gather; 0 when the build has no compact form (fragment-sort path)."""
        return int(self.lib.dll.svo_builder_compact_bytes(self._h))

    def EmitCompactTo(self, d_dst: int, pointer_bias_words: int, skip_root, d_tables, stream=None):
        """Phase 2 in compact form (brick path): upper windows and the rasterized bricks' leaf blocks go to d_dst, 32 bytes
        per brick to d_tables; returns the four plan words expand_compact() needs on the GPU that owns d_dst.
        d_tables = None: the tables have been sent already (PushTables); returns None."""
        plan = (C.c_uint64 * 4)()
        self.lib.check(self.lib.dll.svo_builder_emit_compact_to(self._h, d_dst, pointer_bias_words, int(skip_root), d_tables or None,
                                                                plan if d_tables else None, _stream_ptr(stream)))
        return [int(v) for v in plan] if d_tables else None

    def PushTables(self, pointer_bias_words: int, skip_root, d_tables: int, stream=None):
        """The table part of EmitCompactTo alone (32 bytes per brick to d_tables); returns the four plan words."""
        plan = (C.c_uint64 * 4)()
        self.lib.check(self.lib.dll.svo_builder_push_tables(self._h, pointer_bias_words, int(skip_root), d_tables, plan, _stream_ptr(stream)))
        return [int(v) for v in plan]

    def TopWords(self, stream=None) -> np.ndarray:
        """The blocks kept aside by EmitTo(skip_root=1 or 2): uint32 [n_blocks, 8], the root block first."""
        out, n = (C.c_uint32 * 72)(), C.c_uint32()
        self.lib.check(self.lib.dll.svo_builder_top_words(self._h, out, C.byref(n), _stream_ptr(stream)))
        return np.array(list(out), dtype=np.uint32)[: 8 * n.value].reshape(n.value, 8)

    def RootWords(self, stream=None) -> np.ndarray:
        out = (C.c_uint32 * 8)()
        self.lib.check(self.lib.dll.svo_builder_root_words(self._h, out, _stream_ptr(stream)))
        return np.array(list(out), dtype=np.uint32)

    def GetOctreeRange(self) -> int:
        return int(self.lib.dll.svo_builder_octree_range_bytes(self._h))

    def GetOctree(self) -> int:
        return int(self.lib.dll.svo_builder_octree(self._h) or 0)

    def CmdTransferOctreeOwnership(self, *args, **kwargs):
        """Queue-family ownership transfer (OctreeBuilder.cpp:215-220): nothing to do for a CUDA-produced buffer
        (external memory is acquired from VK_QUEUE_FAMILY_EXTERNAL on the Vulkan side)."""
        return None

    def GetLeafCount(self) -> int:
        return int(self.lib.dll.svo_builder_leaf_count(self._h))

    def GetLevelCounts(self):
        n = self.GetLevel() + 1
        out = (C.c_uint64 * n)()
        self.lib.check(self.lib.dll.svo_builder_level_counts(self._h, out, n))
        return [int(v) for v in out]

    def LastMs(self):
        ms = (C.c_float * len(PHASES))()
        np_ = C.c_uint32()
        self.lib.check(self.lib.dll.svo_builder_last_ms(self._h, ms, C.byref(np_)))
        return {k: float(ms[i]) for i, k in enumerate(PHASES)}, int(np_.value)

    def BuildPath(self) -> int:
        """0: every fragment emitted, sorted and reduced; 1: large triangles binned to bricks (brick.cuh)."""
        return int(self.lib.dll.svo_builder_build_path(self._h))

    def BrickStats(self):
        """({pairs, bricks, small_leaves, raster_bricks}, {raster, scans, keys, emit} in ms) of the last brick-path build."""
        c = (C.c_uint64 * 4)()
        ms = (C.c_float * 4)()
        self.lib.check(self.lib.dll.svo_builder_brick_stats(self._h, c, ms))
        return (dict(pairs=int(c[0]), bricks=int(c[1]), small_leaves=int(c[2]), raster_bricks=int(c[3])),
                dict(raster=float(ms[0]), scans=float(ms[1]), keys=float(ms[2]), emit=float(ms[3])))

    def SortStepMs(self):
        """Milliseconds of each kernel of the last sort (needs svo_debug_profile_passes(1) before the build)."""
        out = (C.c_float * 32)()
        n = self.lib.dll.svo_builder_sort_step_ms(self._h, out, 32)
        if n < 0:
            self.lib.check(n)
        return [float(out[i]) for i in range(n)]

    def ExportFd(self, stream=None):
        """The node buffer as a VK_KHR_external_memory_fd-importable file descriptor: (fd, allocation size, CUDA address)."""
        fd, size, ptr = C.c_int(-1), C.c_uint64(0), _P()
        self.lib.check(self.lib.dll.svo_builder_export_fd(self._h, C.byref(fd), C.byref(size), C.byref(ptr), _stream_ptr(stream)))
        return fd.value, int(size.value), int(ptr.value or 0)

    def RebaseCopy(self, d_dst: int, dst_word_offset: int, base_words: int, stream=None):
        self.lib.check(self.lib.dll.svo_builder_rebase_copy(self._h, d_dst, dst_word_offset, base_words, _stream_ptr(stream)))

    def octree_to_host(self, stream=None) -> np.ndarray:
        return self.lib.to_host(self.GetOctree(), np.uint32, self.GetOctreeRange() // 4, self.device, _stream_ptr(stream))

    def Destroy(self):
        if self._h:
            self.lib.dll.svo_builder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.Destroy()
        except Exception:
            pass


class ShardedBuild:
    """svo_build_sharded: the whole loader sequence on several GPUs of this process (include/svo.h); the stitched node
    buffer lives on devices[0]."""

    def __init__(self):
        self._h = None

    @staticmethod
    def Create(mesh, level: int, mode: int = CONSERVATIVE_EXACT, devices=(0,), lib: Library | None = None):
        self = ShardedBuild()
        self.lib, self.devices, self.level = lib or get_library(), list(devices), level
        pos = np.ascontiguousarray(mesh.positions, dtype=np.float32)
        idx = np.ascontiguousarray(mesh.indices, dtype=np.uint32)
        d = np.ascontiguousarray(mesh.draws)
        draws = (svo_draw * len(d))(*[svo_draw(int(r["first_index"]), int(r["index_count"]), int(r["texture_id"]), int(r["albedo_rgba8"]))
                                      for r in d])
        m = svo_mesh(pos.ctypes.data, pos.shape[1] * 4, 0, idx.ctypes.data, len(pos), len(idx), draws, len(d))
        devs = (C.c_int * len(self.devices))(*self.devices)
        h = _P()
        self.lib.check(self.lib.dll.svo_build_sharded(C.byref(m), level, mode, devs, len(self.devices), C.byref(h)))
        self._h = h
        return self

    def Rebuild(self):
        self.lib.check(self.lib.dll.svo_sharded_rebuild(self._h))

    def GetOctree(self) -> int:
        return int(self.lib.dll.svo_sharded_octree(self._h) or 0)

    def GetOctreeRange(self) -> int:
        return int(self.lib.dll.svo_sharded_octree_range_bytes(self._h))

    def GetLeafCount(self) -> int:
        return int(self.lib.dll.svo_sharded_leaf_count(self._h))

    def GetVoxelFragmentCount(self) -> int:
        return int(self.lib.dll.svo_sharded_fragment_count(self._h))

    def LastMs(self) -> float:
        return float(self.lib.dll.svo_sharded_last_ms(self._h))

    def octree_to_host(self) -> np.ndarray:
        return self.lib.to_host(self.GetOctree(), np.uint32, self.GetOctreeRange() // 4, self.devices[0])

    def Destroy(self):
        if self._h:
            self.lib.dll.svo_sharded_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.Destroy()
        except Exception:
            pass


class Octree:
    """src/Octree.hpp: the hand-off object the tracers read.  Update() takes the builder's buffer, range and level
    (Octree.cpp:22-35); the builder is kept alive because it owns the device allocation."""

    def __init__(self):
        self.m_buffer, self.m_range, self.m_level, self._builder = 0, 0, 0, None

    @staticmethod
    def Create():
        return Octree()

    def Update(self, builder: OctreeBuilder):
        self._builder = builder
        self.m_buffer = builder.GetOctree()
        self.m_level = builder.GetLevel()
        self.m_range = builder.GetOctreeRange()

    def Empty(self) -> bool:
        return self.m_buffer == 0

    def GetBuffer(self) -> int:
        return self.m_buffer

    def GetLevel(self) -> int:
        return self.m_level

    def GetRange(self) -> int:
        return self.m_range


RAY_HIT_DTYPE = np.dtype([("pos", "<f4", 3), ("colour", "<f4", 3), ("normal", "<f4", 3), ("hit", "<u4"), ("iter", "<u4")])


def raymarch_leaf(d_octree: int, origins, dirs, device: int = 0, stream=None, lib: Library | None = None) -> np.ndarray:
    """Octree_RayMarchLeaf (shader/octree.glsl:179-340) on a device node buffer; origins/dirs float32 [N,3] on the host.
    Returns a RAY_HIT_DTYPE array."""
    lib = lib or get_library()
    o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
    assert o.shape == d.shape
    n = len(o)
    d_o, d_d, d_h = lib.to_device(o, device), lib.to_device(d, device), lib.malloc(max(1, n) * RAY_HIT_DTYPE.itemsize, device)
    try:
        lib.check(lib.dll.svo_octree_raymarch_leaf(device, d_octree, n, d_o, d_d, d_h, _stream_ptr(stream)))
        lib.check(lib.dll.svo_stream_synchronize(device, _stream_ptr(stream)))
        raw = lib.to_host(d_h, np.uint8, n * RAY_HIT_DTYPE.itemsize, device)
    finally:
        lib.free(d_o, device), lib.free(d_d, device), lib.free(d_h, device)
    return raw.view(RAY_HIT_DTYPE)


def camera_rays(position, look, side, up, width: int, height: int):
    """octree_tracer.frag:24 + Camera_GenRay (shader/camera.glsl:12-15): one primary ray per pixel, row-major.
    (normalize() precision is the driver's; fp32 x / sqrt(dot) here.)"""
    px, py = np.meshgrid(np.arange(width, dtype=np.float32), np.arange(height, dtype=np.float32))
    cx = (px / np.float32(width)) * np.float32(2) - np.float32(1)
    cy = (py / np.float32(height)) * np.float32(2) - np.float32(1)
    look, side, up = (np.asarray(v, np.float32) for v in (look, side, up))
    d = look[None, None, :] - side[None, None, :] * cx[..., None] - up[None, None, :] * cy[..., None]
    d = (d / np.sqrt((d * d).sum(-1, keepdims=True, dtype=np.float32))).astype(np.float32)
    o = np.broadcast_to(np.asarray(position, np.float32), d.shape)
    return o.reshape(-1, 3).copy(), d.reshape(-1, 3)


def build_svo(mesh, level: int, mode: int = CONSERVATIVE_EXACT, device: int = 0, stream=None, lib: Library | None = None):
    """The loader sequence of src/LoaderThread.cpp:51-89 in one call. Returns (scene, voxelizer, builder)."""
    scene = Scene.Create(mesh, device=device, stream=stream, lib=lib)
    vox = Voxelizer.Create(scene, level, mode, stream=stream)
    builder = OctreeBuilder.Create(vox, stream=stream)
    vox.CmdVoxelize(stream)
    builder.CmdBuild(stream)
    return scene, vox, builder
