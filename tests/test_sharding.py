"""Host-side multi-GPU logic on CPU: partitioning, and the end-of-run gather with world_size 2
over gloo (the data path itself has no collective)."""
import os
import socket

import numpy as np
import pytest

from stereoreconstruction_b200 import sharding


def test_partition_views_covers_everything_once():
    for V in (1, 2, 7, 8, 9):
        for G in (1, 2, 4, 8):
            parts = sharding.partition_views(V, G)
            assert sorted(v for p in parts for v in p) == list(range(V))
            own = sharding.view_owner(V, G)
            for r, p in enumerate(parts):
                assert all(own[v] == r for v in p)
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_row_bands_are_contiguous_and_balanced():
    for H in (1, 7, 64, 1080, 2160):
        for G in (1, 2, 3, 4, 8):
            b = sharding.row_bands(H, G)
            assert b[0][0] == 0 and b[-1][1] == H
            assert all(b[i][1] == b[i + 1][0] for i in range(G - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1
    b = sharding.row_bands(1080, 8, align=8)
    assert all(s % 8 == 0 for s, _ in b) and b[-1][1] == 1080


def test_plan_switches_to_row_bands_when_views_are_few():
    p = sharding.plan(8, 1080, 4)
    assert all(len(x) == 2 and all((b0, b1) == (0, 1080) for _, b0, b1 in x) for x in p)
    p = sharding.plan(2, 1080, 8)
    assert all(len(x) == 2 for x in p)
    rows = sorted((b0, b1) for (v, b0, b1) in (it for x in p for it in x) if v == 0)
    assert rows[0][0] == 0 and rows[-1][1] == 1080 and all(rows[i][1] == rows[i + 1][0] for i in range(7))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, V, h, w, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        work = sharding.plan(V, h, world)
        # stand-in for sr_run_view: a deterministic function of (view, row, column)
        yy, xx = np.mgrid[0:h, 0:w]
        local_i = {(v, b0, b1): ((v * 1000 + yy[b0:b1] * 7 + xx[b0:b1]) % 251).astype(np.int32) for (v, b0, b1) in work[rank]}
        local_d = {k: a.astype(np.float64) * 0.5 + 300.0 for k, a in local_i.items()}
        full_i = sharding.gather_maps(local_i, work, V, (h, w), dist, fill=-1)
        full_d = sharding.gather_maps(local_d, work, V, (h, w), dist, fill=np.nan)
        ok = True
        for v in range(V):
            want = ((v * 1000 + yy * 7 + xx) % 251).astype(np.int32)
            ok &= bool((full_i[v] == want).all()) and bool((full_d[v] == want * 0.5 + 300.0).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("V", [5, 1])  # 5 views over 2 ranks (view sharding); 1 view (row sharding)
def test_gather_world_size_2_gloo(V):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, V, 33, 40, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
