"""Per-rank phase timing of the slab-mode sharded step (run under torchrun)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from sparsevoxeloctree_b200 import api, scenes, sharded

lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
mesh = scenes.living_room_like(); level, mode = 12, api.CONSERVATIVE_EXACT
sh = sharded.ShardedSVO(torch, dist, mesh, level, mode, lr)
st = torch.cuda.current_stream()
for _ in range(3): sh.step(st)
v, b = sh.vox[0], sh.builders[0]
def sync(): torch.cuda.synchronize()
for it in range(3):
    dist.barrier(); sync(); t0 = time.perf_counter()
    v.CmdVoxelize(st); b.Prepare(st); sync(); t1 = time.perf_counter()
    body = b.GetOctreeRange() // 4 - 8
    mine = torch.tensor([body], dtype=torch.int64, device="cuda"); bodies = torch.zeros(dist.get_world_size(), dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(bodies, mine); bl = bodies.cpu().tolist(); t2 = time.perf_counter()
    base = 8 + sum(bl[:dist.get_rank()])
    dst = (sh.final if dist.get_rank() == 0 else sh.peer_final) + base * 4
    b.EmitTo(dst, base, True, st); sync(); t3 = time.perf_counter()
    r = b.RootWords(st); t4 = time.perf_counter()
    dist.barrier(); sync(); t5 = time.perf_counter()
    print(f"rank {dist.get_rank()} it {it}: frags {v.GetVoxelFragmentCount()} body {body*4/1e6:.0f} MB | voxelize+prepare {1e3*(t1-t0):.2f} ms, "
          f"allgather {1e3*(t2-t1):.2f}, emit {1e3*(t3-t2):.2f} ({body*4/1e9/(t3-t2):.0f} GB/s), root {1e3*(t4-t3):.2f}, barrier {1e3*(t5-t4):.2f}, total {1e3*(t5-t0):.2f}", flush=True)
dist.barrier(); dist.destroy_process_group()
