"""Multi-GPU plumbing without GPUs: the band map is integer arithmetic, and the
gather runs over gloo with two processes.  Shard CONTENTS come from the oracle
(as the checker's stand-in for a rank's render); the de-interleave here is a
numpy restatement of lol_deinterleave_kernel's index map."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def np_deinterleave(gathered, w, h, world):
    from loltracer_b200 import sharding

    frame = np.zeros((h, w), gathered.dtype)
    shards = gathered.reshape(world, -1, w)
    for y in range(h):
        rank, lrow = sharding.row_location(y, world)
        frame[y] = shards[rank, lrow]
    return frame


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("h", [1, 4, 7, 90, 240, 2160, 2161])
def test_band_map_is_a_partition(world, h):
    import loltracer_b200 as lb
    from loltracer_b200 import sharding

    seen = []
    for rank in range(world):
        rows = sharding.shard_rows(h, world, rank)
        assert len(rows) == sharding.local_bands(h, world, rank) * 4
        assert len(rows) <= sharding.padded_local_bands(h, world) * 4
        for lrow, y in enumerate(rows):
            if y >= 0:
                assert sharding.row_location(y, world) == (rank, lrow)
                seen.append(y)
    assert sorted(seen) == list(range(h))
    w = 37
    assert lb.shard_pixels(w, h, world) == sharding.padded_local_bands(h, world) * 4 * w


def test_cyclic_bands_balance_the_work(scenes_dir):
    """contiguous blocks are unbalanced, 4-row cyclic bands are not."""
    import loltracer_b200 as lb
    import oracle_lib as ol
    from loltracer_b200 import sharding

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene4.lol"))
    w, h, world = 160, 512, 8
    r = ol.port_render(scene, w, h, counts=True)
    work = (r["nprimary"].astype(np.int64) + r["nshadow"] + 4).sum(axis=1)
    cyc = [sum(work[y] for y in sharding.shard_rows(h, world, k) if y >= 0) for k in range(world)]
    blk = [work[k * h // world:(k + 1) * h // world].sum() for k in range(world)]
    assert max(cyc) / np.mean(cyc) < 1.10 < max(blk) / np.mean(blk)


def _worker(rank, world, port, w, h, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import loltracer_b200 as lb
    import oracle_lib as ol
    from loltracer_b200 import sharding

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    scene = lb.Scene.from_file(os.path.join(ROOT, "tests", "golden", "scenes", "scene2.lol"))
    full = ol.port_render(scene, w, h, threads=2)["rgba"]
    g = sharding.FrameGatherer(w, h, world, rank, torch.device("cpu"))
    local = g.local.numpy().view(np.uint32).reshape(-1, w)
    for lrow, y in enumerate(sharding.shard_rows(h, world, rank)):
        if y >= 0:
            local[lrow] = full[y]  # what this rank's kernel would have stored
    gathered = g.gather()
    ok = True
    if rank == 0:
        frame = np_deinterleave(gathered.numpy().view(np.uint32), w, h, world)
        ok = bool(np.array_equal(frame, full))
        try:
            g.assemble()
            ok = False  # must refuse on CPU tensors: no CPU fallback
        except lb.LolB200Error:
            pass
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("size", [(64, 36), (50, 31)])
def test_gather_world2_gloo(size):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    w, h = size
    procs = [ctx.Process(target=_worker, args=(r, 2, port, w, h, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok in res), res
