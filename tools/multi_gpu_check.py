#!/usr/bin/env python
"""Multi-GPU correctness check (run under torchrun, one rank per GPU, from the repo root):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_check.py

1. view sharding: every rank runs its share of the reference views (sharding.partition_views), the
   depth maps are gathered with the library's own NCCL path (sr_comm_allgather_views: one broadcast
   per view from its owner over NVLink), every rank cross-checks, and rank 0 compares the result
   with the same job computed by itself alone — bit for bit (sharding must be invisible);
2. row sharding of one two-view direction (sr_params.row_begin/row_end + sr_comm_allgather_rows).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from stereoreconstruction_b200 import capi, sharding, types as T  # noqa: E402
from scene_util import refractive_arc_scene  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cams, imgs, ms, _ = refractive_arc_scene(V=6, w=160, h=96, masks=True, arc_deg=35.0)
    V, h, w = len(cams), 96, 160
    P = T.default_params(True, 420.0, 580.0, 48)

    ctx = capi.Context(local)
    ctx.set_views(cams, imgs, ms)
    ctx.set_params(P)
    nb = ctx.select_neighbours(3)
    uid = [capi.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)

    # 1. view sharding + library gather + cross-check on every rank
    for v in sharding.partition_views(V, world)[rank]:
        ctx.run_view(v, nb[v])
    ctx.allgather_views(sharding.view_owner(V, world))
    ctx.cross_check(False, 12.0)
    got = [(ctx.depth(v).copy(), ctx.depth_index(v).copy()) for v in range(V)]

    ok = True
    if rank == 0:
        solo = capi.Context(local)
        solo.set_views(cams, imgs, ms)
        solo.set_params(P)
        for v in range(V):
            solo.run_view(v, nb[v])
        solo.cross_check(False, 12.0)
        for v in range(V):
            d, i = solo.depth(v), solo.depth_index(v)
            same = ((got[v][0] == d) | (np.isnan(got[v][0]) & np.isnan(d))).all() and (got[v][1] == i).all()
            ok &= bool(same)
        print(f"view sharding over {world} GPUs + sr_comm_allgather_views + cross-check == 1 GPU: {ok}")
        solo.close()

    # 2. row sharding of a two-view direction
    P2 = T.default_params(False, 420.0, 580.0, 32, radius=3)
    bands = sharding.row_bands(h, world)
    Q = T.SrParams.from_buffer_copy(P2)
    Q.row_begin, Q.row_end = bands[rank]
    ctx.set_views(cams, imgs, ms)
    ctx.set_params(Q)
    if bands[rank][1] > bands[rank][0]:
        ctx.run_view(1, [2])
    ctx.allgather_rows(1, [b[0] for b in bands], [b[1] for b in bands])
    gd, gi = ctx.depth(1).copy(), ctx.depth_index(1).copy()
    if rank == 0:
        solo = capi.Context(local)
        solo.set_views(cams, imgs, ms)
        solo.set_params(P2)
        solo.run_view(1, [2])
        d, i = solo.depth(1), solo.depth_index(1)
        same = bool(((gd == d) | (np.isnan(gd) & np.isnan(d))).all() and (gi == i).all())
        print(f"row sharding over {world} GPUs + sr_comm_allgather_rows == 1 GPU: {same}")
        ok &= same
        solo.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    ctx.close()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        raise SystemExit(1)
    if rank == 0:
        print("MULTI-GPU CHECK OK")


if __name__ == "__main__":
    main()
