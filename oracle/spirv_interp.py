"""A small SPIR-V interpreter, just large enough to EXECUTE the reference's own checked-in shader binaries for
the SVO construction path (shader/include/spirv/{voxelizer.geom,voxelizer_conservative.geom,voxelizer.frag,
octree_init_node.comp,octree_tag_node.comp,octree_alloc_node.comp,octree_modify_arg.comp}.u32).

TEST INFRASTRUCTURE ONLY.  It exists so that the CPU oracle (oracle/svo_oracle.c) can be pinned against outputs
of the reference itself: the reference cannot run in this image (no Vulkan ICD), but its shaders are plain
SPIR-V 1.3 modules, and everything except the fixed-function rasterizer is defined by them.
tests/golden/make_spirv_golden.py drives this interpreter over /root/reference (where it is mounted) and
commits the resulting vectors; the tests then only need the committed vectors.

Semantics implemented: 32-bit int/uint/float scalars and vectors, structs / arrays / runtime arrays with
explicit Offset / ArrayStride layout for buffer blocks, the structured control flow that glslang emits
(Branch, BranchConditional, Switch, Phi), atomics, MatrixTimesVector, GLSL.std.450 {FAbs, FMin, FMax, UMin, UMax,
UClamp, Fma (one rounding), Cross, PackUnorm4x8}, geometry-shader EmitVertex, fragment Kill, and subgroup ops with subgroup size 1 (a legal
implementation choice: every invocation is its own subgroup).  Floats are numpy.float32, one rounding per
operation (no contraction) -- the same choice as the pinned arithmetic of DESIGN.md section 3.
"""
from __future__ import annotations

import re

import numpy as np

F32 = np.float32
M32 = 0xFFFFFFFF


class Discard(Exception):
    pass


class Ptr:
    """Pointer into interpreter memory: either a python container + path, or a buffer binding + byte offset."""
    __slots__ = ("kind", "obj", "path", "type_id")

    def __init__(self, kind, obj, path, type_id):
        self.kind, self.obj, self.path, self.type_id = kind, obj, path, type_id


def _f(v):
    return F32(v)


class Module:
    def __init__(self, words, spec=None):
        assert words[0] == 0x07230203, "not SPIR-V"
        self.words = words
        self.spec = spec or {}
        self.types, self.consts, self.decor, self.mdecor, self.names = {}, {}, {}, {}, {}
        self.vars, self.labels, self.ext = {}, {}, {}
        self.entry, self.exec_model = None, None
        self.code = []
        i = 5
        while i < len(words):
            op, n = words[i] & 0xFFFF, words[i] >> 16
            self.code.append((op, words[i + 1:i + n]))
            i += n
        self._scan()

    @staticmethod
    def from_u32_file(path, spec=None):
        words = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", open(path).read())]
        return Module(words, spec)

    # ---- static scan: types, constants, decorations, labels -------------------------------------------
    def _scan(self):
        T, Cn = self.types, self.consts
        for pc, (op, a) in enumerate(self.code):
            if op == 71:
                self.decor.setdefault(a[0], {})[a[1]] = a[2:] if len(a) > 2 else [1]
            elif op == 72:
                self.mdecor.setdefault((a[0], a[1]), {})[a[2]] = a[3:] if len(a) > 3 else [1]
            elif op == 11:
                self.ext[a[0]] = "glsl"
            elif op == 15:
                self.exec_model, self.entry = a[0], a[1]
        for pc, (op, a) in enumerate(self.code):
            if op == 19: T[a[0]] = ("void",)
            elif op == 20: T[a[0]] = ("bool",)
            elif op == 21: T[a[0]] = ("int", a[1], a[2])
            elif op == 22: T[a[0]] = ("float", a[1])
            elif op == 23: T[a[0]] = ("vector", a[1], a[2])
            elif op == 24: T[a[0]] = ("matrix", a[1], a[2])
            elif op in (25, 26, 27): T[a[0]] = ("opaque",)
            elif op == 28: T[a[0]] = ("array", a[1], a[2])
            elif op == 29: T[a[0]] = ("rtarray", a[1])
            elif op == 30: T[a[0]] = ("struct", list(a[1:]))
            elif op == 32: T[a[0]] = ("pointer", a[1], a[2])
            elif op == 33: T[a[0]] = ("function",)
            elif op in (43, 50):  # Constant / SpecConstant
                t = T[a[0]]
                raw = a[2]
                if op == 50:
                    sid = self.decor.get(a[1], {}).get(1)
                    if sid is not None and sid[0] in self.spec:
                        raw = int(self.spec[sid[0]]) & M32
                Cn[a[1]] = self._from_raw(t, raw)
            elif op in (41, 48): Cn[a[1]] = True
            elif op in (42, 49): Cn[a[1]] = False
            elif op in (44, 51): Cn[a[1]] = [Cn[x] for x in a[2:]]
            elif op == 46: Cn[a[1]] = self._null(a[0])
            elif op == 1: Cn[a[1]] = self._null(a[0])  # OpUndef: any value is valid
            elif op == 52:  # SpecConstantOp
                Cn[a[1]] = self._alu(a[2], a[0], [Cn[x] for x in a[3:]], a[3:])
            elif op == 59 and a[2] != 7:
                self.vars[a[1]] = (a[0], a[2])
            elif op == 248:
                self.labels[a[0]] = pc

    def _from_raw(self, t, raw):
        if t[0] == "float":
            return np.array([raw], dtype=np.uint32).view(np.float32)[0]
        return raw & M32

    def _null(self, tid):
        t = self.types[tid]
        if t[0] == "bool": return False
        if t[0] == "int": return 0
        if t[0] == "float": return F32(0)
        if t[0] in ("vector", "matrix"): return [self._null(t[1]) for _ in range(t[2])]
        if t[0] == "array": return [self._null(t[1]) for _ in range(self.consts[t[2]])]
        if t[0] == "struct": return [self._null(m) for m in t[1]]
        return None

    # ---- layout of buffer blocks -----------------------------------------------------------------------
    def _buf_offset(self, tid, indices):
        """Walk a type with constant/dynamic indices; return (byte offset, final type id)."""
        off = 0
        for ix in indices:
            t = self.types[tid]
            if t[0] == "struct":
                off += self.mdecor[(tid, ix)][35][0]
                tid = t[1][ix]
            elif t[0] in ("array", "rtarray"):
                off += ix * self.decor[tid][6][0]
                tid = t[1]
            elif t[0] == "vector":
                off += ix * 4
                tid = t[1]
            else:
                raise NotImplementedError(t)
        return off, tid

    def _buf_load(self, arr, off, tid):
        t = self.types[tid]
        if t[0] == "int": return int(arr[off // 4])
        if t[0] == "float": return arr[off // 4:off // 4 + 1].view(np.float32)[0]
        if t[0] == "vector": return [self._buf_load(arr, off + 4 * k, t[1]) for k in range(t[2])]
        raise NotImplementedError(t)

    def _buf_store(self, arr, off, tid, v):
        t = self.types[tid]
        if t[0] == "int": arr[off // 4] = v & M32
        elif t[0] == "float": arr[off // 4:off // 4 + 1].view(np.float32)[0] = v
        elif t[0] == "vector":
            for k in range(t[2]): self._buf_store(arr, off + 4 * k, t[1], v[k])
        else: raise NotImplementedError(t)

    # ---- ALU ---------------------------------------------------------------------------------------------
    def _signed(self, v): return v - (1 << 32) if v & 0x80000000 else v

    def _alu(self, op, rtid, v, ids=None):
        t = self.types[rtid]
        if t[0] == "vector" and op not in (148, 154, 155, 79, 80, 81, 82, 142):
            n = t[2]
            vv = [x if isinstance(x, list) else [x] * n for x in v]
            return [self._alu(op, t[1], [x[k] for x in vv]) for k in range(n)]
        a = v[0] if v else None
        b = v[1] if len(v) > 1 else None
        if op == 128: return (a + b) & M32
        if op == 130: return (a - b) & M32
        if op == 132: return (a * b) & M32
        if op == 134: return (a // b) & M32 if b else 0
        if op == 137: return (a % b) & M32 if b else 0
        if op == 126: return (-a) & M32
        if op == 129: return F32(a + b)
        if op == 131: return F32(a - b)
        if op == 133: return F32(a * b)
        if op == 136: return F32(a / b) if b != 0 else F32(np.inf if a > 0 else (-np.inf if a < 0 else np.nan))
        if op == 127: return F32(-a)
        if op == 194: return (a >> (b & 31)) & M32
        if op == 195: return (self._signed(a) >> (b & 31)) & M32
        if op == 196: return (a << (b & 31)) & M32
        if op == 197: return a | b
        if op == 198: return a ^ b
        if op == 199: return a & b
        if op == 200: return (~a) & M32
        if op == 164: return a == b
        if op == 165: return a != b
        if op == 166: return bool(a or b)
        if op == 167: return bool(a and b)
        if op == 168: return not a
        if op == 169: return v[1] if v[0] else v[2]
        if op == 170: return a == b
        if op == 171: return a != b
        if op == 172: return a > b
        if op == 174: return a >= b
        if op == 176: return a < b
        if op == 178: return a <= b
        if op == 173: return self._signed(a) > self._signed(b)
        if op == 175: return self._signed(a) >= self._signed(b)
        if op == 177: return self._signed(a) < self._signed(b)
        if op == 179: return self._signed(a) <= self._signed(b)
        if op == 180: return bool(a == b)
        if op == 182: return bool(a != b)
        if op == 184: return bool(a < b)
        if op == 186: return bool(a > b)
        if op == 183: return bool(a != b)   # FUnordNotEqual (true for NaN, like python !=)
        if op == 188: return bool(a <= b)
        if op == 190: return bool(a >= b)
        if op == 109:  # ConvertFToU: truncation; negative / NaN are undefined in SPIR-V -> 0
            return 0 if not (a > 0) else (M32 if a >= 4294967296.0 else int(a))
        if op == 110: return int(a) & M32
        if op == 111: return F32(self._signed(a))
        if op == 112: return F32(a)
        if op == 124:  # Bitcast (32-bit scalars)
            if t[0] == "float": return np.array([a], dtype=np.uint32).view(np.float32)[0]
            return int(np.array([a], dtype=np.float32).view(np.uint32)[0]) if isinstance(a, np.floating) else a
        raise NotImplementedError(f"opcode {op}")

    def _ext(self, inst, rtid, v):
        t = self.types[rtid]
        if inst == 68:  # Cross
            a, b = v
            return [F32(F32(a[1] * b[2]) - F32(a[2] * b[1])), F32(F32(a[2] * b[0]) - F32(a[0] * b[2])),
                    F32(F32(a[0] * b[1]) - F32(a[1] * b[0]))]
        if inst == 64:  # UnpackUnorm4x8
            return [F32(F32((v[0] >> (8 * k)) & 0xff) / F32(255.0)) for k in range(4)]
        if inst == 69:  # Normalize: x * inversesqrt(dot(x, x)) evaluated as x / sqrt(dot) (driver-defined precision)
            acc = F32(0)
            for c in v[0]: acc = F32(acc + F32(c * c))
            return [F32(c / np.sqrt(acc)) for c in v[0]]
        if inst == 75:  # FindUMsb
            return (int(v[0]).bit_length() - 1) & M32
        if inst == 55:  # PackUnorm4x8
            out = 0
            for k, c in enumerate(v[0]):
                out |= int(np.rint(np.clip(F32(c), 0, 1) * F32(255.0))) << (8 * k)
            return out
        if t[0] == "vector":
            n = t[2]
            vv = [x if isinstance(x, list) else [x] * n for x in v]
            return [self._ext(inst, t[1], [x[k] for x in vv]) for k in range(n)]
        if inst == 4: return F32(abs(v[0]))
        if inst == 37: return v[1] if v[1] < v[0] else v[0]       # FMin: y < x ? y : x
        if inst == 40: return v[1] if v[0] < v[1] else v[0]       # FMax: x < y ? y : x
        if inst == 38: return min(v[0], v[1])
        if inst == 41: return max(v[0], v[1])
        if inst == 44: return min(max(v[0], v[1]), v[2])          # UClamp
        if inst == 50: return F32(np.float64(v[0]) * np.float64(v[1]) + np.float64(v[2]))  # Fma: one rounding
        if inst == 43: return min(max(v[0], v[1]), v[2])          # FClamp
        if inst == 13: return F32(np.sin(np.float64(v[0])))       # Sin / Pow: driver-defined precision
        if inst == 26: return F32(np.power(np.float64(v[0]), np.float64(v[1])))
        raise NotImplementedError(f"GLSL.std.450 {inst}")

    # ---- execution ---------------------------------------------------------------------------------------
    def run(self, inputs=None, buffers=None, push=None, on_emit=None, sampler=None, on_ext=None):
        """Execute the entry point once.
        inputs : {variable id or BuiltIn number or ('loc', n): python value}
        buffers: {(set, binding): np.uint32 array}
        push   : python value (nested lists) of the push-constant block
        Returns {output variable id: value}; raises Discard on OpKill."""
        T, Cn = self.types, self.consts
        val = dict(Cn)
        mem = {}
        for vid, (ptid, sc) in self.vars.items():
            pointee = T[ptid][2]
            d = self.decor.get(vid, {})
            if sc in (12, 2) and (d.get(33) is not None):
                val[vid] = Ptr("buf", buffers[(d.get(34, [0])[0], d[33][0])], 0, pointee)
                continue
            if sc == 9:
                mem[vid] = [push]
            elif sc == 1:
                v = None
                if inputs is not None:
                    if vid in inputs: v = inputs[vid]
                    elif 11 in d and ("builtin", d[11][0]) in inputs: v = inputs[("builtin", d[11][0])]
                    elif 30 in d and ("loc", d[30][0]) in inputs: v = inputs[("loc", d[30][0])]
                    elif ("input", vid) in inputs: v = inputs[("input", vid)]
                    elif "gl_in" in inputs and T[pointee][0] == "array" and 30 not in d: v = inputs["gl_in"]
                mem[vid] = [v if v is not None else self._null(pointee)]
            else:
                mem[vid] = [self._null(pointee)]
            val[vid] = Ptr("mem", mem[vid], [0], pointee)
        outs = {vid: mem[vid] for vid, (ptid, sc) in self.vars.items() if sc == 3}

        def deref(p):
            o = p.obj
            for k in p.path: o = o[k]
            return o

        def load(p):
            if p.kind == "buf": return self._buf_load(p.obj, p.path, p.type_id)
            v = deref(p)
            return [x for x in v] if isinstance(v, list) and not any(isinstance(x, list) for x in v) else v

        def store(p, v):
            if p.kind == "buf":
                self._buf_store(p.obj, p.path, p.type_id, v)
                return
            o = p.obj
            for k in p.path[:-1]: o = o[k]
            o[p.path[-1]] = v

        # find the entry function body
        pc = next(i for i, (op, a) in enumerate(self.code) if op == 54 and a[1] == self.entry)
        prev_label = cur_label = None
        steps = 0
        while True:
            op, a = self.code[pc]
            pc += 1
            steps += 1
            if steps > 5_000_000: raise RuntimeError("interpreter step limit")
            if op in (54, 55, 246, 247): continue
            if op == 248:
                prev_label, cur_label = cur_label, a[0]
                continue
            if op == 59:  # function-local variable
                pointee = T[a[0]][2]
                cell = [val[a[3]] if len(a) > 3 else self._null(pointee)]
                val[a[1]] = Ptr("mem", cell, [0], pointee)
            elif op == 61:
                p = val[a[2]]
                if T[a[0]][0] == "opaque": val[a[1]] = ("tex", p.path[-1] if p.kind == "mem" else 0)  # a combined image sampler
                else: val[a[1]] = load(p)
            elif op == 87:  # ImageSampleImplicitLod: the driver's sampler -- supplied by the caller
                val[a[1]] = [F32(x) for x in sampler(val[a[2]][1], val[a[3]])]
            elif op == 62: store(val[a[0]], val[a[1]])
            elif op in (65, 66):
                base = val[a[2]]
                idx = [val[x] for x in a[3:]]
                if base.kind == "buf":
                    off, tid = self._buf_offset(base.type_id, idx)
                    val[a[1]] = Ptr("buf", base.obj, base.path + off, tid)
                else:
                    val[a[1]] = Ptr("mem", base.obj, base.path + idx, T[a[0]][2])
            elif op == 12:
                val[a[1]] = self._ext(a[3], a[0], [val[x] for x in a[4:]])
                if on_ext: on_ext(a[3], [val[x] for x in a[4:]], val[a[1]])
            elif op == 79:  # VectorShuffle
                both = list(val[a[2]]) + list(val[a[3]])
                val[a[1]] = [both[k] if k != M32 else self._null(T[a[0]][1]) for k in a[4:]]
            elif op == 80:
                out = []
                for x in a[2:]:
                    v = val[x]
                    if isinstance(v, list) and T[a[0]][0] == "vector": out.extend(v)
                    else: out.append(v)
                val[a[1]] = out
            elif op == 81:
                v = val[a[2]]
                for k in a[3:]: v = v[k]
                val[a[1]] = v
            elif op == 82:
                import copy
                comp = copy.deepcopy(val[a[3]])
                o = comp
                for k in a[4:-1]: o = o[k]
                o[a[-1]] = val[a[2]]
                val[a[1]] = comp
            elif op == 77: val[a[1]] = val[a[2]][val[a[3]]]
            elif op == 142: val[a[1]] = [F32(c * val[a[3]]) for c in val[a[2]]]
            elif op == 145:  # MatrixTimesVector: sum_j column_j * v[j], one rounding per operation
                cols, vec = val[a[2]], val[a[3]]
                out = []
                for r in range(len(cols[0])):
                    acc = F32(cols[0][r] * vec[0])
                    for j in range(1, len(cols)): acc = F32(acc + F32(cols[j][r] * vec[j]))
                    out.append(acc)
                val[a[1]] = out
            elif op == 148:
                acc = F32(0)
                for x, y in zip(val[a[2]], val[a[3]]): acc = F32(acc + F32(x * y))
                val[a[1]] = acc
            elif op == 154: val[a[1]] = any(val[a[2]])
            elif op == 155: val[a[1]] = all(val[a[2]])
            elif op == 245:
                for k in range(2, len(a), 2):
                    if a[k + 1] == prev_label:
                        val[a[1]] = val[a[k]]
                        break
                else:
                    raise RuntimeError("phi without matching predecessor")
            elif op == 249:
                pc = self.labels[a[0]]
            elif op == 250:
                pc = self.labels[a[1] if val[a[0]] else a[2]]
            elif op == 251:
                sel = val[a[0]]
                tgt = a[1]
                for k in range(2, len(a), 2):
                    if a[k] == sel: tgt = a[k + 1]
                pc = self.labels[tgt]
            elif op == 252: raise Discard()
            elif op in (253, 254): break
            elif op == 56: break
            elif op == 234:  # AtomicIAdd
                p = val[a[2]]
                old = load(p)
                store(p, (old + val[a[5]]) & M32)
                val[a[1]] = old
            elif op == 230:  # AtomicCompareExchange
                p = val[a[2]]
                old = load(p)
                if old == val[a[7]]: store(p, val[a[6]])
                val[a[1]] = old
            elif op == 333: val[a[1]] = True                                  # Elect (subgroup size 1)
            elif op == 339: val[a[1]] = [1 if val[a[3]] else 0, 0, 0, 0]      # Ballot
            elif op == 342:                                                    # BallotBitCount
                bits = bin(val[a[4]][0]).count("1")
                val[a[1]] = bits if a[3] in (0, 1) else 0                      # Reduce / InclusiveScan / ExclusiveScan
            elif op in (337, 338): val[a[1]] = val[a[3]]                       # Broadcast(First)
            elif op == 218:
                if on_emit: on_emit({vid: (list(c[0]) if isinstance(c[0], list) else c[0]) for vid, c in outs.items()})
            elif op == 219: pass
            elif op in (86, 87): raise NotImplementedError("texture sampling is outside the built path")
            else:
                val[a[1]] = self._alu(op, a[0], [val[x] for x in a[2:]], a[2:])
        return {vid: c[0] for vid, c in outs.items()}

    # helpers for drivers ---------------------------------------------------------------------------------
    def var_by_builtin(self, builtin, storage):
        for vid, (ptid, sc) in self.vars.items():
            if sc == storage and self.decor.get(vid, {}).get(11, [None])[0] == builtin: return vid
        return None

    def var_by_location(self, loc, storage):
        for vid, (ptid, sc) in self.vars.items():
            if sc == storage and self.decor.get(vid, {}).get(30, [None])[0] == loc: return vid
        return None
