"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle.

Tolerances are BASELINE.json's: hit/miss mask agreement >= 99.99 %, RGB within
1/255 per channel on agreeing hits.  In the exact arithmetic mode the march is
expected to be bit-identical (distance and id of every pixel); only powf differs
between glibc and CUDA, hence the 1/255.
"""
import os

import numpy as np
import pytest

import oracle_lib as ol
from conftest import EXAMPLES

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _render(lb, scene, w, h, camera=None, options=None, shard=None, counts=False):
    r = lb.Renderer(scene, options, device=0)
    dev = "cuda:0"
    frame = torch.zeros((h, w), dtype=torch.int32, device=dev)
    dist = torch.zeros((h, w), dtype=torch.float32, device=dev)
    ids = torch.zeros((h, w), dtype=torch.int32, device=dev)
    npr = torch.zeros((h, w), dtype=torch.int16, device=dev)
    nsh = torch.zeros((h, w), dtype=torch.int16, device=dev)
    aux = lb.Aux(dist=dist.data_ptr(), id=ids.data_ptr(), primary_steps=npr.data_ptr(),
                 shadow_steps=nsh.data_ptr())
    r.render_device(frame.data_ptr(), w, h, camera=camera, aux=aux, shard=shard,
                    stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = dict(rgba=frame.cpu().numpy().view(np.uint32), dist=dist.cpu().numpy(),
               id=ids.cpu().numpy().view(np.uint32), nprimary=npr.cpu().numpy().view(np.uint16),
               nshadow=nsh.cpu().numpy().view(np.uint16), renderer=r)
    return out


def _check(got, want, exact=True):
    cmp = ol.compare_frames(got["rgba"], got["id"], want["rgba"], want["id"])
    assert cmp["mask_agree"] >= 0.9999, cmp
    assert cmp["max_rgb_err"] <= 1, cmp          # 1/255 per channel on agreeing hits
    assert cmp["max_miss_rgb_err"] == 0, cmp     # misses are exactly the background
    if exact:
        assert cmp["n_mask_off"] == 0, cmp
        assert np.array_equal(got["id"], want["id"])
        assert np.array_equal(got["dist"].view(np.uint32), want["dist"].view(np.uint32))
    return cmp


@pytest.mark.parametrize("name", EXAMPLES)
@pytest.mark.parametrize("size", [(320, 240), (1920, 1080)])
def test_examples_match_oracle(name, size, scenes_dir):
    import loltracer_b200 as lb

    w, h = size
    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    got = _render(lb, scene, w, h)
    want = ol.port_render(scene, w, h, counts=True)
    _check(got, want)
    # the primary march takes exactly the reference's number of steps per pixel
    assert np.array_equal(got["nprimary"], want["nprimary"])
