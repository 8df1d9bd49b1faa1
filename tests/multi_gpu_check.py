"""torchrun helper for tests/test_gpu_parity.py::test_two_rank_nccl_stitch: octant-sharded build on N GPUs,
rank 0 compares the stitched tree with the oracle-checked single-GPU build."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparsevoxeloctree_b200 import api, scenes, sharded  # noqa: E402
from tests.parity import assert_same_tree  # noqa: E402


def main():
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    # --room: walls parallel to the grid (flat bricks), what the compact gather of the slab mode is for
    mesh = scenes.living_room_like(n_boxes=8, n_small=2000, level=9) if "--room" in sys.argv else scenes.random_soup(800, 41, 0.01, 1.2)
    level, mode = 9, api.CONSERVATIVE_EXACT
    sh = sharded.ShardedSVO(torch, dist, mesh, level, mode, lr, use_ipc="--no-ipc" not in sys.argv)
    for _ in range(2):
        nbytes = sh.step(torch.cuda.current_stream())
    if dist.get_rank() == 0:
        stitched = sh.octree_to_host()
        scene, vox, builder = api.build_svo(mesh, level, mode, device=lr)
        assert_same_tree(stitched, builder.octree_to_host(), level)
        assert nbytes == len(stitched) * 4
        print("STITCH_OK", nbytes, flush=True)
    dist.barrier()
    sh.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
