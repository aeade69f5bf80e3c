"""The renderer.h drop-in (b200_renderer.c) under the headless twin of main.c."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from conftest import ROOT

pytestmark = pytest.mark.gpu

HOST = os.path.join(ROOT, "loltracer_b200", "backend", "build", "lol_headless_b200")


@pytest.mark.parametrize("name,threads,size", [("scene", 1, (320, 240)), ("scene4", 8, (320, 240)),
                                               ("scene3", 3, (1283, 721))])
def test_backend_under_main_protocol(name, threads, size, scenes_dir, tmp_path):
    import loltracer_b200 as lb

    if not os.path.exists(HOST):
        pytest.skip("headless host not built (needs the reference headers at build time)")
    w, h = size
    raw = tmp_path / "frame.bin"
    path = os.path.join(scenes_dir, name + ".lol")
    out = subprocess.run([HOST, str(threads), path, "--size", f"{w}x{h}", "--frames", "3", "--warmup", "1",
                          "--raw", str(raw), "--verbose"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = np.fromfile(raw, np.uint32).reshape(h, w)
    want = ol.port_render(lb.Scene.from_file(path), w, h)
    hit = want["id"] != 0
    assert ((got & 0xFFFFFF) != 0).sum() > 0
    err = np.zeros(got.shape, np.int32)
    for s in (16, 8, 0):
        err = np.maximum(err, np.abs(((got >> s) & 0xFF).astype(np.int32) - ((want["rgba"] >> s) & 0xFF).astype(np.int32)))
    assert err.max() <= 1 and err[~hit].max() == 0
    assert "Frame 3" in out.stdout and "Cerrando" in out.stdout


def test_backend_dumps_generated_code(scenes_dir, tmp_path):
    """-j / --jitdump keeps its meaning: leave the generated code for a profiler."""
    if not os.path.exists(HOST):
        pytest.skip("headless host not built")
    out = subprocess.run([HOST, "2", os.path.join(scenes_dir, "scene2.lol"), "-j", "--size", "64x48"],
                         capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "lol_sdf" in (tmp_path / "lol-b200-kernel.cu").read_text()
    assert (tmp_path / "lol-b200-kernel.cubin").read_bytes()[:4] == b"\x7fELF"


@pytest.mark.parametrize("gather", ["nccl", "peer", "host"])
def test_backend_on_all_gpus_of_the_box(gather, scenes_dir, tmp_path):
    """--gpus N: one host process drives every GPU; same frame as one GPU.  Needs >= 2 GPUs
    (gpurun --gpus 2); on a single-GPU box it is skipped."""
    import loltracer_b200 as lb

    n = lb.device_count()
    if n < 2:
        pytest.skip("one GPU on this box")
    if not os.path.exists(HOST):
        pytest.skip("headless host not built")
    w, h = 1283, 721
    path = os.path.join(scenes_dir, "scene4.lol")
    frames = []
    for gpus in (1, n):
        raw = tmp_path / f"frame{gpus}.bin"
        out = subprocess.run([HOST, "4", path, "--size", f"{w}x{h}", "--frames", "3", "--gpus", str(gpus),
                              "--gather", gather, "--raw", str(raw)], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr
        frames.append(np.fromfile(raw, np.uint32))
    assert np.array_equal(frames[0], frames[1])


def test_group_api_matches_single_gpu(scenes_dir):
    import loltracer_b200 as lb

    n = lb.device_count()
    if n < 2:
        pytest.skip("one GPU on this box")
    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene3.lol"))
    w, h = 1000, 563
    one = np.zeros((h, w), np.uint32)
    r = lb.Renderer(scene)
    r.render_host(one.ctypes.data, w, h)
    for gather in ("nccl", "peer", "host"):
        g = lb.Group(scene, n, gather)
        got = np.zeros((h, w), np.uint32)
        for _ in range(2):
            ms = g.render_host(got.ctypes.data, w, h)
        assert np.array_equal(got, one), gather
        assert ms > 0
        g.close()
    r.close()


@pytest.mark.parametrize("threads,gpus", [(1, 4), (3, 2), (8, 2), (2, 5)])
@pytest.mark.parametrize("pin", [False, True])
def test_worker_threads_pull_gpu_shares(threads, gpus, pin, scenes_dir, tmp_path):
    """main.c's N worker threads each take a GPU's share of the frame (b200_renderer.c: shares are
    pulled through current_line like scanlines): fewer threads than GPUs, more threads than GPUs,
    pageable and pinned surfaces -- the same frame as one GPU.  --wrap-devices lets the N-share
    protocol run on a box with fewer GPUs (share i on GPU i mod count), so this runs everywhere."""
    if not os.path.exists(HOST):
        pytest.skip("headless host not built")
    w, h = 1283, 721
    path = os.path.join(scenes_dir, "scene4.lol")
    one = tmp_path / "one.bin"
    out = subprocess.run([HOST, "2", path, "--size", f"{w}x{h}", "--frames", "2", "--raw", str(one)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    raw = tmp_path / "many.bin"
    cmd = [HOST, str(threads), path, "--size", f"{w}x{h}", "--frames", "4", "--warmup", "1", "--gpus", str(gpus),
           "--gather", "host", "--wrap-devices", "--raw", str(raw)] + (["--pin-surface"] if pin else [])
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    assert np.array_equal(np.fromfile(raw, np.uint32), np.fromfile(one, np.uint32))


def test_group_shares_driven_by_threads_and_resizes(scenes_dir):
    """lolb200_group_share_enqueue / _wait: every share of a frame driven by its own host thread,
    frame after frame with the surface changing size in between; equals the single-GPU frame.
    Two shares on ONE device when the box has a single GPU."""
    import threading

    import loltracer_b200 as lb

    n = max(2, lb.device_count())
    devices = [i % lb.device_count() for i in range(n)]
    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene3.lol"))
    r = lb.Renderer(scene)
    g = lb.Group(scene, n, "host", devices=devices)
    assert g.size == n
    for (w, h) in [(1000, 563), (320, 240), (1000, 563), (37, 9)]:
        one = np.zeros((h, w), np.uint32)
        r.render_host(one.ctypes.data, w, h)
        got = np.zeros((h, w), np.uint32)
        errors = []

        def drive(share):
            try:
                g.share_enqueue(share, got.ctypes.data, w, h)
                g.share_wait(share)
            except Exception as e:  # surfaced below
                errors.append(e)

        ts = [threading.Thread(target=drive, args=(i,)) for i in range(n)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert not errors, errors
        assert np.array_equal(got, one), (w, h)
        # and the one-thread entry point over the same group
        got[:] = 0
        g.render_host(got.ctypes.data, w, h)
        assert np.array_equal(got, one), (w, h)
    g.close()
    r.close()
