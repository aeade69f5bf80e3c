"""Sharded rendering on ONE GPU: every rank's launch is issued in turn on the same
device, the gather is a concatenation, the de-interleave is the CUDA kernel.  The
NCCL leg itself runs under `gpurun --gpus N` through bench.py."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("size", [(640, 360), (333, 129)])
def test_shards_reassemble_to_the_full_frame(world, size, scenes_dir):
    import loltracer_b200 as lb
    from loltracer_b200 import sharding

    w, h = size
    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene4.lol"))
    r = lb.Renderer(scene)
    st = torch.cuda.current_stream().cuda_stream
    full = torch.zeros((h, w), dtype=torch.int32, device="cuda:0")
    r.render_device(full.data_ptr(), w, h, stream=st)
    shard_px = lb.shard_pixels(w, h, world)
    gathered = torch.full((world, shard_px), 0x55, dtype=torch.int32, device="cuda:0")
    for rank in range(world):
        r.render_device(gathered[rank].data_ptr(), w, h, pitch_px=w,
                        shard=lb.Shard(rank=rank, world=world), stream=st)
    frame = torch.zeros((h, w), dtype=torch.int32, device="cuda:0")
    lb.deinterleave(gathered.data_ptr(), frame.data_ptr(), w, h, world, shard_px, stream=st)
    torch.cuda.synchronize()
    assert torch.equal(frame, full)
    # and the compact layout is exactly what sharding.shard_rows says
    g = gathered.cpu().numpy().reshape(world, -1, w)
    f = full.cpu().numpy()
    for rank in range(world):
        for lrow, y in enumerate(sharding.shard_rows(h, world, rank)):
            if y >= 0:
                assert np.array_equal(g[rank, lrow], f[y])
    r.close()


@pytest.mark.parametrize("world", [2, 8])
def test_full_frame_destination(world, scenes_dir):
    """dst_full_frame: each rank stores its bands at their final place (what the
    peer-store gather does into rank 0's frame over NVLink)."""
    import loltracer_b200 as lb

    w, h = 500, 203
    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene3.lol"))
    r = lb.Renderer(scene)
    st = torch.cuda.current_stream().cuda_stream
    full = torch.zeros((h, w), dtype=torch.int32, device="cuda:0")
    r.render_device(full.data_ptr(), w, h, stream=st)
    frame = torch.zeros((h, w), dtype=torch.int32, device="cuda:0")
    for rank in range(world):
        r.render_device(frame.data_ptr(), w, h, pitch_px=w,
                        shard=lb.Shard(rank=rank, world=world, dst_full_frame=1), stream=st)
    torch.cuda.synchronize()
    assert torch.equal(frame, full)
    r.close()


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("size,pad", [((640, 360), 0), ((333, 129), 5), ((64, 7), 0), ((1920, 1080), 0)])
def test_ranks_copy_their_own_bands_into_one_host_surface(world, size, pad, scenes_dir):
    """lolb200_render_host_shard: each rank renders its cyclic bands and ITS copy engine
    writes them to their final rows of the full-frame host surface (strided 2-D copies,
    ragged last band, foreign pitch).  All ranks run on this one GPU here; on an 8-GPU box
    the same calls come from eight processes with a shared-memory surface (bench.py)."""
    import loltracer_b200 as lb

    w, h = size
    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene3.lol"))
    r = lb.Renderer(scene, device=0)
    want = np.zeros((h, w), np.uint32)
    r.render_host(want.ctypes.data, w, h)
    host = np.full((h, w + pad), 0xDEADBEEF, np.uint32)
    for rank in range(world):
        r.render_host_shard(host.ctypes.data, w, h, lb.Shard(rank=rank, world=world),
                            pitch_bytes=(w + pad) * 4)
    assert np.array_equal(host[:, :w], want)
    assert (host[:, w:] == 0xDEADBEEF).all()
    r.close()
