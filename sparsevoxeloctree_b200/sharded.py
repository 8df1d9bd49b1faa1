"""Octant-sharded SVO build over several GPUs -- one process per GPU, torch.distributed
for the plumbing.

The grid is split by the top-level octant bits (child slot x | y<<1 | z<<2 of the root block,
shader/octree_tag_node.comp:24-25).  Octant o belongs to rank o % world_size, which is the x / xy / xyz
split for 2 / 4 / 8 ranks.  Every rank voxelizes and builds the (level-1)-deep subtree of each of its
octants in cube-local coordinates (svo_shard); the only exchange step is:

  1. all_gather of the 8 subtree sizes (words)                         -- tiny collective
  2. identical exclusive scan on every rank -> base word offset of each subtree
  3. every rank adds its base to the child pointers of its subtrees WHILE storing them into rank 0's
     buffer: k_rebase_copy writes straight through a CUDA-IPC mapping of that buffer (NVLink P2P), so the
     pointer fix-up is fused with the transfer; without P2P it stages locally and NCCL send/recv moves it
  4. rank 0 writes the 8 root words (0x80000000 | base_o, or 0 for an empty octant).

The union of the shards' fragment sets equals the single-GPU fragment set exactly (fragments are culled
against the half-open voxel window inside the raster kernels), so the stitched tree canonicalises to the
single-GPU tree bit for bit.
"""
from __future__ import annotations

import os

import numpy as np

ROOT_WORDS = 8


def octants_of_rank(rank: int, world: int):
    """Octant ids (x | y<<1 | z<<2) owned by `rank`."""
    if world not in (1, 2, 4, 8):
        raise ValueError("world size must be 1, 2, 4 or 8 (octant sharding)")
    return [o for o in range(8) if o % world == rank]


def octant_cube(o: int):
    return (o & 1, (o >> 1) & 1, (o >> 2) & 1)


def slab_window(rank: int, world: int, level: int):
    """Voxel window (lo, hi) of the octants owned by `rank`: rank bit 0 selects the x half, bit 1 the y half,
    bit 2 the z half (the x / xy / xyz split for 2 / 4 / 8 ranks)."""
    res, half = 1 << level, 1 << (level - 1)
    lo, hi = [0, 0, 0], [res, res, res]
    for axis in range({1: 0, 2: 1, 4: 2, 8: 3}[world]):
        bit = (rank >> axis) & 1
        lo[axis], hi[axis] = bit * half, bit * half + half
    return lo, hi


def plan_offsets(words_per_octant):
    """words_per_octant[o] = node words of octant o's subtree (0 when the octant is empty).
    Returns (base word offset per octant, total words).  Subtrees are laid out after the root block in
    octant order; every rank computes the same plan from the all-gathered sizes."""
    base, run = [], ROOT_WORDS
    for w in words_per_octant:
        w = int(w)
        if w % 8:
            raise ValueError("subtree sizes are multiples of 8 words")
        base.append(run if w else 0)
        run += w
    if run >= 1 << 30:
        raise OverflowError("stitched octree needs >= 2^30 words: 30-bit child pointers cannot address it")
    return base, run


def root_block(bases, words_per_octant) -> np.ndarray:
    """The 8 root words: an internal node pointing at each non-empty octant's subtree root block."""
    out = np.zeros(ROOT_WORDS, dtype=np.uint32)
    for o in range(8):
        if int(words_per_octant[o]):
            out[o] = np.uint32(0x80000000 | int(bases[o]))
    return out


HEADER_WORDS = 72  # pipelined slab mode: root block + the 8 depth-1 blocks at fixed places, the parts' bodies behind them


def sub_windows(rank: int, world: int, level: int, n_sub: int):
    """The slab of `rank` cut into n_sub (1, 2 or 4) boxes along y (and z) at multiples of res/4 -- depth-2 cell borders,
    so that parts built separately share nothing below the depth-1 blocks (svo_builder_emit_to, skip_root = 2)."""
    if n_sub not in (1, 2, 4):
        raise ValueError("n_sub must be 1, 2 or 4")
    lo, hi = slab_window(rank, world, level)
    boxes = [(list(lo), list(hi))]
    for axis in ((1,) if n_sub == 2 else (1, 2) if n_sub == 4 else ()):
        out = []
        for blo, bhi in boxes:
            mid = (blo[axis] + bhi[axis]) // 2
            if mid % max(1, (1 << level) // 4):
                raise ValueError("the cut is not on a depth-2 cell border (level too small)")
            a_hi, b_lo = list(bhi), list(blo)
            a_hi[axis], b_lo[axis] = mid, mid
            out += [(blo, a_hi), (b_lo, bhi)]
        boxes = out
    return boxes


def merge_top_blocks(tops) -> np.ndarray:
    """tops: per separately built part, the blocks svo_builder_top_words returns (uint32 [1 + n1, 8]: its root block, then
    its depth-1 blocks in the order of the root's non-empty slots; child pointers of the depth-1 blocks already final).
    Returns the 72 header words: root block (pointing at the fixed depth-1 block places 8, 16, ... 64) + 8 depth-1
    blocks, each the sum of the parts' blocks for that octant (the parts' depth-2 cells are disjoint)."""
    header = np.zeros(HEADER_WORDS, dtype=np.uint64)
    for t in tops:
        t = np.asarray(t, dtype=np.uint32).reshape(-1, 8)
        if len(t) == 0:
            continue
        octants = [o for o in range(8) if t[0][o]]
        if len(octants) != len(t) - 1:
            raise ValueError("top blocks do not match the root block")
        for k, o in enumerate(octants):
            header[8 * (1 + o): 8 * (2 + o)] += t[1 + k]
            header[o] = 0x80000000 | (8 * (1 + o))
    if (header >> np.uint64(32)).any():
        raise ValueError("overlapping parts: a child slot was written twice")
    return header.astype(np.uint32)


def merge_headers(per_rank: np.ndarray) -> np.ndarray:
    """per_rank: [world, 72] headers (merge_top_blocks of every rank's parts, all-gathered).  Depth-1 blocks are summed
    (disjoint depth-2 cells), the root block takes the word of whichever ranks have the octant."""
    h = np.asarray(per_rank, dtype=np.int64).reshape(-1, HEADER_WORDS)
    header = h.sum(axis=0)
    header[:8] = h[:, :8].max(axis=0)
    if (header >> 32).any():
        raise ValueError("overlapping parts: a child slot was written twice")
    return header.astype(np.uint32)


def rebase_words_numpy(words: np.ndarray, base: int) -> np.ndarray:
    """Host restatement of k_rebase_copy for the CPU (gloo) tests: add base to internal child pointers."""
    w = np.asarray(words, dtype=np.uint32).copy()
    internal = (w & np.uint32(0xC0000000)) == np.uint32(0x80000000)
    w[internal] += np.uint32(base)
    return w


def exchange_sizes(torch, dist, local_sizes, world: int, device):
    """all_gather of the per-octant subtree sizes; returns words[8] indexed by octant id (same on every rank)."""
    if world == 1:
        return [int(v) for v in local_sizes]
    local = torch.tensor([int(v) for v in local_sizes], dtype=torch.int64, device=device)
    out = torch.zeros(8, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, local)
    # all_gather orders by rank; octant o is slot (o // world) of rank (o % world)
    per_rank = out.view(world, 8 // world).cpu().numpy()
    return [int(per_rank[o % world, o // world]) for o in range(8)]


class ShardedSVO:
    """Multi-GPU build driver.  `dist` is torch.distributed (initialised, backend nccl on GPUs); `torch` is
    passed in so that this module imports without it.  With dist=None it runs all 8 octants on one GPU
    ("virtual shards"): that is how a level-14 grid, whose 42 Morton bits + 24 colour bits do not fit one
    64-bit fragment, is built on a single device."""

    def __init__(self, torch, dist, mesh, level: int, mode: int, device: int, lib=None, use_ipc: bool = True):
        from . import api
        self.torch, self.dist, self.api = torch, dist, api
        self.lib = lib or api.get_library()
        self.rank, self.world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
        self.device, self.level, self.mode = device, level, mode
        self.octants = octants_of_rank(self.rank, self.world)
        self.scene = api.Scene.Create(mesh, device=device, lib=self.lib)
        self.vox, self.builders = [], []
        # Slab mode (several ranks, and the full-level Morton code still fits a 64-bit fragment): each rank runs ONE
        # build over the window made of its octants, in global coordinates, and emits its node words straight into
        # rank 0's buffer.  Octant mode (level 14, or a single GPU): one cube-local build per octant.
        self.slab = self.world > 1 and 3 * level + 24 <= 64 and use_ipc
        # the slab is built as n_sub separate parts cut at depth-2 cell borders: the node words of part k cross NVLink
        # while part k + 1 is being voxelized and sorted
        # (measured on 2 B200, C4: 1 part 4.83 ms, 2 parts 4.81 ms, 4 parts 5.33 ms -- what the overlap wins, the smaller
        # sorts lose; one part stays the default)
        self.n_sub = int(os.environ.get("SVO_SUBSLABS", "1")) if self.slab and level >= 3 else 1
        if self.slab:
            for lo, hi in sub_windows(self.rank, self.world, level, self.n_sub):
                v = api.Voxelizer.CreateWindowed(self.scene, level, mode, lo, hi)
                self.vox.append(v)
                self.builders.append(api.OctreeBuilder.Create(v))
        else:
            for o in self.octants:
                v = api.Voxelizer.Create(self.scene, level, mode, shard=(1, octant_cube(o)))
                self.vox.append(v)
                self.builders.append(api.OctreeBuilder.Create(v))
        self.tdev = torch.device("cuda", device) if torch is not None else None
        self.final = None       # rank 0: the stitched node buffer (device pointer, cudaMalloc'ed)
        self.final_cap = 0
        self.peer_final = 0     # ranks > 0: rank 0's buffer mapped through CUDA IPC
        self.use_ipc = use_ipc and self.world > 1
        self.total_words = 0
        self.bases = [0] * 8
        self.words = [0] * 8
        self.stage = None
        self.push_stream = None
        # pipelined slab mode, ranks > 0: staging buffer per part (cached per process like the arena: cudaMalloc is slow)
        # how a part's node words reach rank 0: "copy" = emitted into local memory, then one asynchronous device-to-device
        # copy (the copy engine moves them while the SMs build the next part); "store" = the emit kernel stores them
        # over NVLink itself (no staging, but the kernel holds its SM slots for the duration of the transfer)
        # (2 B200, C4, one part: store 4.83 ms, copy 5.31 ms)
        self.push_mode = os.environ.get("SVO_PUSH", "store")
        # compact gather (slab mode, parts built on the brick path): SVO_COMPACT=0 sends finished node words instead
        self.compact = os.environ.get("SVO_COMPACT", "1") != "0"
        self.stage_base, self.stage_cap = 0, 0
        # SVO_SLAB_TRACE=1: host-side time stamps of every slab step (ms between: start, prepared, sizes exchanged, pushed,
        # [rank 0: tables of all ranks here,] plans exchanged (rank 0: expansions enqueued), own stores complete,
        # [rank 0: top blocks of all ranks here, header written,] device idle)
        self._trace = [] if os.environ.get("SVO_SLAB_TRACE") else None
        self._table_events = []  # one per part sent in compact form: its tables have arrived on rank 0
        # SVO_DEFER_REST = 1 / 0: launch the other stores of a compact part after / before the plan exchange
        # (default: after, when 4 or more ranks share rank 0's NVLink port)
        self.defer_rest = {"1": True, "0": False}.get(os.environ.get("SVO_DEFER_REST", ""), self.world >= 4)
        self._pending_rest = []  # parts whose tables are on their way and whose other stores have not been launched yet
        if self.slab:
            ar = ShardedSVO._STAGE_ARENA.get((self.device, self.world))
            if ar:
                self.stage_base, self.stage_cap = (ar["ptr"] if self.rank == 0 else ar["peer"]), ar["cap"]
        self.pipelined = True  # overlap the NVLink push of one octant with the build of the next (steady state)

    # -- rank 0 owns a grow-only arena for the stitched tree (like the reference's up-front octree buffer).
    #    It is cached per process and device, so a fresh ShardedSVO (the end-to-end path) reuses the allocation
    #    and the IPC mapping.  Every rank sees the same total, so all ranks take the same branch: no collective
    #    is needed unless the arena really has to grow.
    _ARENA = {}  # device -> dict(ptr, cap, peer)
    _STAGE = {}  # (device, part) -> (ptr, capacity in words)
    _STAGE_ARENA = {}  # (device, world) -> dict(ptr, cap, peer): rank 0's staging area of the compact gather
    _PUSH_STREAM = {}  # device -> torch stream
    _XCHG = {}  # (device, world, width) -> the exchange tensors of the slab mode

    def _ensure_final(self, words: int):
        dist = self.dist
        need = int(words)
        ar = ShardedSVO._ARENA.setdefault((self.device, self.world, self.use_ipc), dict(ptr=0, cap=0, peer=0))
        if need > ar["cap"]:
            cap = max(need + need // 4, 1 << 20)
            if self.rank == 0:
                if ar["ptr"]:
                    self.lib.free(ar["ptr"], self.device)
                ar["ptr"] = self.lib.malloc(cap * 4, self.device)
            if self.use_ipc:
                handle = [self.lib.ipc_export(ar["ptr"], self.device) if self.rank == 0 else None]
                dist.broadcast_object_list(handle, 0)
                if self.rank != 0:
                    if ar["peer"]:
                        self.lib.ipc_close(ar["peer"], self.device)
                    ar["peer"] = self.lib.ipc_open(handle[0], self.device)
            ar["cap"] = cap
        self.final, self.final_cap, self.peer_final = (ar["ptr"] if self.rank == 0 else None), ar["cap"], ar["peer"]

    def _ensure_stage(self, nbytes: int):
        """Rank 0's staging area for the per-brick tables of the compact gather (32 bytes per brick of every other rank),
        mapped into the other ranks through CUDA IPC; cached and grow-only like the arena.  Collective when it grows."""
        ar = ShardedSVO._STAGE_ARENA.setdefault((self.device, self.world), dict(ptr=0, cap=0, peer=0))
        if nbytes > ar["cap"]:
            cap = max(nbytes + nbytes // 4, 1 << 20)
            if self.rank == 0:
                if ar["ptr"]:
                    self.lib.free(ar["ptr"], self.device)
                ar["ptr"] = self.lib.malloc(cap, self.device)
            handle = [self.lib.ipc_export(ar["ptr"], self.device) if self.rank == 0 else None]
            self.dist.broadcast_object_list(handle, 0)
            if self.rank != 0:
                if ar["peer"]:
                    self.lib.ipc_close(ar["peer"], self.device)
                ar["peer"] = self.lib.ipc_open(handle[0], self.device)
            ar["cap"] = cap
        self.stage_base, self.stage_cap = (ar["ptr"] if self.rank == 0 else ar["peer"]), ar["cap"]

    def _step_slab(self, stream):
        """Slab mode: every rank builds its slab as n_sub parts; as soon as a part's sizes are known (one small
        all_gather) it is sent into rank 0's buffer over NVLink on a second stream -- while the rank voxelizes and sorts
        its next part.  Layout: 72-word header (root block + 8 depth-1 blocks, merged by rank 0 from the parts' top
        blocks), then the bodies part by part.

        What crosses the link: a part built on the brick path goes in COMPACT form (svo_builder_emit_compact_to) -- the
        windows above depth L-2 and the leaf blocks of the rasterized bricks to their final places, 32 bytes per brick
        (record + two ranks) to a staging area on rank 0 -- and rank 0 generates the flat bricks' blocks and all pointer
        blocks of the two deepest windows itself (svo_expand_compact) once the tables have arrived, at local HBM speed.
        A part built on the fragment-sort path stores its finished node words (svo_builder_emit_to).

        Three small exchanges per step: (1) per part, the sizes -> offsets; (2) the plan words of the compact parts, entered
        when a rank's tables have arrived -> rank 0 starts expanding; (3) the top blocks, entered when a rank's stores have
        completed -> rank 0 writes the merged header."""
        torch, dist = self.torch, self.dist
        import time
        trace = self._trace
        t = [time.perf_counter()] if trace is not None else None

        def mark():
            if t is not None:
                t.append(time.perf_counter())

        if self.push_stream is None:
            self.push_stream = ShardedSVO._PUSH_STREAM.setdefault(self.device, torch.cuda.Stream(self.tdev))
        width = HEADER_WORDS
        bufs = ShardedSVO._XCHG.get((self.device, self.world, self.n_sub))
        if bufs is None:  # exchange buffers, once per process: device tensors for the collectives, pinned host mirrors
            def dev(n):
                return torch.zeros(n, dtype=torch.int64, device=self.tdev)

            def pin(n):
                return torch.zeros(n, dtype=torch.int64).pin_memory()

            bufs = dict(mine=dev(1), gathered=dev(self.world), gathered_h=pin(self.world),
                        plans=dev(6 * self.n_sub), plans_h=pin(6 * self.n_sub),
                        plans_all=dev(self.world * 6 * self.n_sub), plans_all_h=pin(self.world * 6 * self.n_sub),
                        local=dev(width), local_h=pin(width), headers=dev(self.world * width), headers_h=pin(self.world * width))
            ShardedSVO._XCHG[(self.device, self.world, self.n_sub)] = bufs
        mine, gathered = bufs["mine"], bufs["gathered"]
        run, stage_run, placed = HEADER_WORDS, 0, []
        overflow = self.final_cap == 0
        plans = np.zeros((self.n_sub, 6), dtype=np.int64)  # per part: the four plan words, staging offset, base word
        for k, (v, b) in enumerate(zip(self.vox, self.builders)):
            v.CmdVoxelize(stream)
            b.Prepare(stream)  # ends with the size read-back: the stream is idle afterwards
            mark()
            body = b.GetOctreeRange() // 4 - 8 * (1 + b.GetLevelCounts()[1]) if b.GetLeafCount() else 0
            tables = (b.CompactBytes() + 255) // 256 * 256 if (self.compact and self.rank != 0 and body) else 0
            mine.fill_(body | (tables // 256) << 32)
            dist.all_gather_into_tensor(gathered, mine)
            bufs["gathered_h"].copy_(gathered)  # (device -> pinned host: returns when the values are there)
            g = [int(x) for x in bufs["gathered_h"].tolist()]
            mark()
            bodies, stages = [x & 0xFFFFFFFF for x in g], [(x >> 32) * 256 for x in g]
            base = run + sum(bodies[: self.rank])
            stage_off = stage_run + sum(stages[: self.rank])
            run += sum(bodies)
            stage_run += sum(stages)
            if run >= 1 << 30:
                raise OverflowError("stitched octree needs >= 2^30 words: 30-bit child pointers cannot address it")
            # (the same on every rank: all see the same sizes)
            overflow = overflow or run > self.final_cap or stage_run > self.stage_cap
            if body:
                placed.append((k, b, base, body, stage_off if tables else None))
                if not overflow:
                    plans[k] = self._push_part(b, base, body, stream, stage_off if tables else None)
        if overflow:  # first step, or the tree outgrew the arena: size it now and emit everything (again)
            self.push_stream.synchronize()
            self._table_events, self._pending_rest = [], []  # (what was sent before the arena turned out too small is sent again)
            self._ensure_final(run)
            if stage_run > self.stage_cap:
                self._ensure_stage(stage_run)
            for k, b, base, body, stage_off in placed:
                plans[k] = self._push_part(b, base, body, stream, stage_off)
        mark()
        # -- second exchange: the plan words of the parts sent in compact form.  A rank enters it when its TABLES have
        #    arrived on rank 0, and launches the kernels that store the upper windows and the rasterized bricks' leaf blocks
        #    right behind it: rank 0 expands the tables while the rest crosses.
        if self.rank != 0:
            for ev in self._table_events:
                ev.synchronize()
            bufs["plans_h"].copy_(torch.from_numpy(plans.reshape(-1)))
            with torch.cuda.stream(self.push_stream):  # (the push stream waits for the collective: what follows on it runs after)
                bufs["plans"].copy_(bufs["plans_h"], non_blocking=True)
                dist.all_gather_into_tensor(bufs["plans_all"], bufs["plans"])
            for b, dst, base in self._pending_rest:  # upper windows + the rasterized bricks' leaf blocks, to their final places
                b.EmitCompactTo(dst, base, 2, None, self.push_stream)
        else:
            dist.all_gather_into_tensor(bufs["plans_all"], bufs["plans"])
        self._table_events, self._pending_rest = [], []
        if self.rank == 0:
            bufs["plans_all_h"].copy_(bufs["plans_all"])
            mark()
            pl = bufs["plans_all_h"].numpy().reshape(self.world, self.n_sub, 6)
            for r in range(1, self.world):
                for k in range(self.n_sub):
                    plan = pl[r, k]
                    if plan[0]:
                        self.api.expand_compact(self.lib, self.device, self.stage_base + int(plan[4]), [int(x) for x in plan[:4]],
                                                self.final + int(plan[5]) * 4, stream)
        mark()
        # -- third exchange: the parts' top blocks (root block + depth-1 blocks).  A rank enters it when all its stores have
        #    completed, so it doubles as the completion signal; rank 0 contributes nothing (it merges its own top blocks
        #    below) and enters at once.
        if self.rank != 0:
            tops = [b.TopWords(self.push_stream) for _, b, _, _, _ in placed]  # (waits for this rank's stores)
            bufs["local_h"].copy_(torch.from_numpy(merge_top_blocks(tops).astype(np.int64)))
            bufs["local"].copy_(bufs["local_h"], non_blocking=True)
        mark()
        dist.all_gather_into_tensor(bufs["headers"], bufs["local"])
        self.total_words = run
        if self.rank == 0:
            bufs["headers_h"].copy_(bufs["headers"])
            mark()
            h = bufs["headers_h"].numpy().reshape(self.world, HEADER_WORDS).copy()
            tops = [b.TopWords(self.push_stream) for _, b, _, _, _ in placed]
            h[0] = merge_top_blocks(tops).astype(np.int64)
            header = merge_headers(h)
            self.lib.check(self.lib.dll.svo_memcpy_h2d(self.device, self.final, header.ctypes.data, header.nbytes, self.api._stream_ptr(stream)))
            mark()
        # No barrier at the end: a rank's next step cannot touch rank 0's memory before the next step's first all_gather,
        # which rank 0 joins only after this synchronize -- the other ranks are free to start voxelizing their next slab.
        torch.cuda.synchronize(self.tdev)
        mark()
        if trace is not None:
            trace.append([1e3 * (b_ - a_) for a_, b_ in zip(t[:-1], t[1:])])
        return run * 4

    def _push_part(self, b, base, body, stream, stage_off=None):
        """Node words of a prepared part -> words [base, base + body) of rank 0's buffer, child pointers final.
        stage_off: the part goes in compact form, its tables to that offset of rank 0's staging area.  Returns the part's
        six plan words (zeros when nothing is left for rank 0 to expand)."""
        torch = self.torch
        none = np.zeros(6, dtype=np.int64)
        if self.rank == 0:
            # local: straight into the stitched buffer, on the push stream (the part's stream is idle after Prepare), so that
            # the collectives on the caller's stream do not queue behind it
            b.EmitTo(self.final + base * 4, base, 2, self.push_stream)
            return none
        dst = self.peer_final + base * 4
        if stage_off is not None:
            plan = b.PushTables(base, 2, self.stage_base + stage_off, self.push_stream)  # the tables first ...
            ev = torch.cuda.Event()
            ev.record(self.push_stream)
            self._table_events.append(ev)
            if self.defer_rest:
                # ... the kernels that store the rest are launched after the plan exchange (_step_slab): the collective's own
                # messages would otherwise queue behind their stores on rank 0's NVLink port
                self._pending_rest.append((b, dst, base))
            else:
                b.EmitCompactTo(dst, base, 2, None, self.push_stream)  # ... then the kernels that store the rest
            return np.array(plan + [stage_off, base], dtype=np.uint64).astype(np.int64)
        if self.push_mode == "store":
            b.EmitTo(dst, base, 2, self.push_stream)
            return none
        key = (self.device, self.builders.index(b))
        buf, cap = ShardedSVO._STAGE.get(key, (0, 0))
        if cap < body:
            if buf:
                self.torch.cuda.synchronize(self.tdev)
                self.lib.free(buf, self.device)
            cap = body + body // 8
            buf = self.lib.malloc(cap * 4, self.device)
            ShardedSVO._STAGE[key] = (buf, cap)
        b.EmitTo(buf, base, 2, stream)  # fast local emit on the build stream ...
        self.push_stream.wait_stream(stream if stream is not None else torch.cuda.current_stream(self.tdev))
        self.lib.check(self.lib.dll.svo_memcpy_d2d(self.device, dst, buf, body * 4, int(self.push_stream.cuda_stream)))  # ... DMA over NVLink
        return none

    def step(self, stream=None):
        """One sharded build: local subtrees, size exchange, fused rebase + gather, root block on rank 0."""
        if self.slab:
            return self._step_slab(stream)
        if self.world > 1 and self.use_ipc and self.total_words and self.pipelined:
            done = self._step_pipelined(stream)
            if done is not None:
                return done
        torch, dist = self.torch, self.dist
        for v, b in zip(self.vox, self.builders):
            v.CmdVoxelize(stream)
            b.CmdBuild(stream)
        local = [b.GetOctreeRange() // 4 if b.GetLeafCount() else 0 for b in self.builders]
        words = exchange_sizes(torch, dist, local, self.world, self.tdev)
        self.words = words
        self.bases, self.total_words = plan_offsets(words)
        self._ensure_final(self.total_words)
        if self.rank == 0 or self.use_ipc:
            dst = self.final if self.rank == 0 else self.peer_final
            for o, b in zip(self.octants, self.builders):
                if words[o]:
                    b.RebaseCopy(dst, self.bases[o], self.bases[o], stream)
        if self.world > 1 and not self.use_ipc:
            self._gather_nccl(stream)
        if self.rank == 0:
            rb = root_block(self.bases, words)
            self.lib.check(self.lib.dll.svo_memcpy_h2d(self.device, self.final, rb.ctypes.data, rb.nbytes, 0))
        self.lib.check(self.lib.dll.svo_stream_synchronize(self.device, self.api._stream_ptr(stream)))
        if self.world > 1:
            dist.barrier()  # remote stores into rank 0's buffer are complete
        return self.total_words * 4

    def _step_pipelined(self, stream):
        """Steady-state variant (the stitched-tree arena already exists): octants are built in rounds -- round r is
        the r-th octant of every rank -- and as soon as a round's sizes are exchanged, each rank pushes that
        subtree (rebase fused with the P2P store) on a second stream while it builds its next octant.  The layout
        is the same as step()'s: subtrees in octant order after the root block.  Returns None (nothing pushed
        yet) if the arena turns out too small, so that step() can take the sizing path."""
        torch, dist = self.torch, self.dist
        if self.push_stream is None:
            self.push_stream = torch.cuda.Stream(self.tdev)
        build_stream = stream if stream is not None else torch.cuda.current_stream(self.tdev)
        n_rounds = 8 // self.world
        words, bases, run = [0] * 8, [0] * 8, ROOT_WORDS
        dst = self.final if self.rank == 0 else self.peer_final
        sizes = torch.zeros(self.world, dtype=torch.int64, device=self.tdev)
        mine = torch.zeros(1, dtype=torch.int64, device=self.tdev)
        pushed = False
        for r in range(n_rounds):
            v, b = self.vox[r], self.builders[r]
            v.CmdVoxelize(build_stream)
            b.CmdBuild(build_stream)
            built = torch.cuda.Event()
            built.record(build_stream)
            mine.fill_(b.GetOctreeRange() // 4 if b.GetLeafCount() else 0)
            dist.all_gather_into_tensor(sizes, mine)
            round_sizes = sizes.cpu().tolist()
            for k in range(self.world):
                o = k + r * self.world
                words[o] = int(round_sizes[k])
                bases[o] = run if words[o] else 0
                run += words[o]
            if run > self.final_cap or run >= 1 << 30:
                if pushed:
                    raise OverflowError("stitched octree outgrew its arena mid-step")
                return None
            o = self.octants[r]
            if words[o]:
                self.push_stream.wait_event(built)
                b.RebaseCopy(dst, bases[o], bases[o], self.push_stream)
                pushed = True
        self.words, self.bases, self.total_words = words, bases, run
        if self.rank == 0:
            rb = root_block(bases, words)
            self.lib.check(self.lib.dll.svo_memcpy_h2d(self.device, self.final, rb.ctypes.data, rb.nbytes, 0))
        torch.cuda.synchronize(self.tdev)
        dist.barrier()  # remote stores into rank 0's buffer are complete
        return self.total_words * 4

    def _gather_nccl(self, stream):
        """Fallback without P2P mapping: rebase into a local staging tensor, NCCL send/recv to rank 0."""
        torch, dist = self.torch, self.dist
        if self.rank != 0:
            n = sum(self.words[o] for o in self.octants)
            if self.stage is None or self.stage.numel() < n:
                self.stage = torch.empty(max(n, 8), dtype=torch.int32, device=self.tdev)
            off = 0
            for o, b in zip(self.octants, self.builders):
                if self.words[o]:
                    b.RebaseCopy(self.stage.data_ptr(), off, self.bases[o], stream)
                    off += self.words[o]
            torch.cuda.synchronize(self.tdev)
            off = 0
            for o in self.octants:
                if self.words[o]:
                    dist.send(self.stage[off:off + self.words[o]], 0)
                    off += self.words[o]
        else:
            for o in range(8):
                r = o % self.world
                if r != 0 and self.words[o]:
                    if self.stage is None or self.stage.numel() < self.words[o]:
                        self.stage = torch.empty(self.words[o], dtype=torch.int32, device=self.tdev)
                    buf = self.stage[: self.words[o]]
                    dist.recv(buf, r)
                    self.lib.check(self.lib.dll.svo_memcpy_d2d(self.device, self.final + self.bases[o] * 4, buf.data_ptr(),
                                                               self.words[o] * 4, 0))

    def leaf_count_local(self) -> int:
        return sum(b.GetLeafCount() for b in self.builders)

    def fragment_count_local(self) -> int:
        return sum(v.GetVoxelFragmentCount() for v in self.vox)

    def octree_to_host(self) -> np.ndarray:
        assert self.rank == 0
        return self.lib.to_host(self.final, np.uint32, self.total_words, self.device)

    def destroy(self):
        if self._trace and len(self._trace) > 3:
            rows = [r for r in self._trace[2:] if len(r) == len(self._trace[-1])]
            avg = [sum(c) / len(c) for c in zip(*rows)]
            print(f"[slab trace] rank {self.rank}: " + " ".join(f"{x:.3f}" for x in avg) + f" | total {sum(avg):.3f} ms over {len(rows)} steps",
                  file=__import__("sys").stderr, flush=True)
        for b in self.builders:
            b.Destroy()
        for v in self.vox:
            v.Destroy()
        self.scene.Destroy()
        self.final, self.peer_final = None, 0  # the stitched-tree arena is cached for the process (see _ARENA)
