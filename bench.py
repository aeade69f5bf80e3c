#!/usr/bin/env python
"""bench.py -- Mrays/s and ms/frame of the sphere-tracing hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--scene scene4] [--size 3840x2160]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      the reference's CPU renderer, host cores

A step is one frame of the workload (default: scene4.lol at 3840x2160, BASELINE
config C3).  With N > 1 the frame is sharded in 4-row bands, band b -> rank b % N,
and every step ends with the complete frame in rank 0's HBM:

  --gather peer (default)  every rank's render kernel stores its bands straight into rank
                           0's frame over NVLink (CUDA IPC mapping) and its last CTA
                           publishes a frame number in a flag word next to the frame; rank
                           0's stream waits for the flags with stream memory operations
                           (cuStreamWaitValue32).  No collective on the data path.
  --gather nccl            compact shards, dist.gather over NVLink, de-interleave kernel.
                           Its step time is reported beside the default as gather_nccl_ms.

Prints ONE JSON line (rank 0).  `value` is whole-job Mrays/s with the frame left
in HBM; `e2e` is the same metric through the host-surface entry point
(lolb200_render_host: camera in, pixels copied into a host buffer).  Extra keys
(GPU-only, seconds each): per_config = every BASELINE config at this N,
moving_camera_ms, tail_us / barrier_us per rank, gather_nccl_ms, frame_latency_ms,
inprocess_group_ms (the single-process renderer.h drop-in on all N GPUs).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "Mrays/s"
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.45, SURVEY.md 8d
FLAG_BYTES = 4096  # completion flags behind rank 0's frame (one word per rank, spaced 128 B)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="scene4")
    ap.add_argument("--size", default=None, help="default 3840x2160 (frame) / 7680x4320 (orbit)")
    ap.add_argument("--gather", default="peer", choices=["nccl", "peer", "peer-allreduce"],
                    help="peer: stores into rank 0's frame + completion flags (default); peer-allreduce: the same "
                         "stores with a 4-byte NCCL all-reduce as the barrier (round 1's path); nccl: gather + "
                         "de-interleave")
    ap.add_argument("--e2e-path", default="host-shards", choices=["host-shards", "gather-then-copy"],
                    help="N > 1, e2e leg: every rank copies its own bands into one shared-memory host frame over "
                         "its own PCIe link (default), or the frame is gathered on rank 0 and copied from there")
    ap.add_argument("--workload", default="frame", choices=["frame", "orbit"],
                    help="frame: one frame per step, sharded by bands over the GPUs (configs C1-C4); "
                         "orbit: one step = 64 camera-orbit frames, whole frames dealt to the GPUs (C5)")
    ap.add_argument("--orbit-frames", type=int, default=64)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--arith", default="exact", choices=["exact", "fast"])
    ap.add_argument("--opts", default="", help="extra lowering options, k=v,k=v (A/B runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-slabs", default="", help="N > 1: also time the e2e path with these slab counts per shard "
                    "(LOLB200_SHARD_SLABS), e.g. 2,4 -> extra key e2e_slab_sweep (A/B inside one run)")
    ap.add_argument("--no-extras", action="store_true", help="skip per_config / moving camera / in-process group")
    ap.add_argument("--all-scenes", action="store_true",
                    help="also time the other example scenes at this size (extra key per_scene)")
    return ap.parse_args()


def scene_path(name):
    return name if os.path.exists(name) else os.path.join(ROOT, "tests", "golden", "scenes", name + ".lol")


def load_scene(lb, name):
    if name in ("synthetic", "synthetic_csg"):
        from loltracer_b200 import scenegen
        return lb.Scene.from_string(scenegen.synthetic_scene_text(csg=name == "synthetic_csg"))
    return lb.Scene.from_file(scene_path(name))


def workload_name(scene, w, h):
    """The same string in both arms (the driver compares them)."""
    return f"{scene}.lol at {w}x{h}, one primary ray per pixel"


# ------------------------------------------------------------------ clocks --

REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
           0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}


class ClockSampler:
    """Polls NVML for SM clock and clock-event reasons while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------ CPU baseline --
# The only place bench.py touches oracle/: the reference's own naive_renderer.c compiled
# unmodified (oracle/_ref/liblolref.so), or the oracle port when that library did not travel.


def build_checkers():
    """oracle/ only (plain C, make): the reference arm must not map the product library."""
    for target in ("port", "ref"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), target], check=True,
                       stdout=subprocess.DEVNULL)


def scene_text(name):
    if name in ("synthetic", "synthetic_csg"):
        from loltracer_b200 import scenegen  # pure Python text generator: loads no native code
        return scenegen.synthetic_scene_text(csg=name == "synthetic_csg")
    return open(scene_path(name)).read()


def ref_protocol_frames(scene_name, w, h, frames, threads=None):
    """`frames` whole frames through the reference's UNMODIFIED render_thread() under main.c's semaphore
    protocol (oracle/ref_harness.c: lolref_render_protocol), worker threads persistent over the frames,
    scene loaded by the reference's own scene.c.  Returns the per-frame wall times in ms."""
    import numpy as np
    import oracle_lib as ol

    rs = ol.RefScene(text=scene_text(scene_name))
    _, ms = rs.render_protocol(w, h, threads=threads or ol.nthreads(), frames=frames)
    return list(ms)


def ref_probe_sample(scene_name, w, h, ystride, repeats=1, threads=None):
    """Every `ystride`-th scanline of the w x h frame through the reference's static pipeline functions
    (lolref_probe) -- the bounded sample for workloads whose whole frame takes minutes on the CPU."""
    import oracle_lib as ol

    rs = ol.RefScene(text=scene_text(scene_name))
    best = min(rs.probe(w, h, ystride=ystride, threads=threads)["ms"] for _ in range(repeats))
    rows = (h + ystride - 1) // ystride
    return best, rows * w


def port_probe_sample(scene_name, w, h, ystride, repeats=1):
    """The oracle port on the same sample, parsed by the port library's own copy of the front-end (the
    product library stays unloaded)."""
    import numpy as np
    import oracle_lib as ol

    L = ol.port()
    raw = scene_text(scene_name).encode()
    sp = C.c_void_p()
    L.lolb200_scene_parse_string.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]
    assert L.lolb200_scene_parse_string(raw, len(raw), C.byref(sp)) == 0
    rows = (h + ystride - 1) // ystride
    bufs = [np.zeros((rows, w), t) for t in (np.float32, np.uint32, np.uint32)]
    tot = (C.c_uint64 * 4)()
    best = None
    for _ in range(repeats):
        ms = L.lolo_render(sp, None, 0, w, h, 0, h, ystride, ol.nthreads(), bufs[0].ctypes.data, bufs[1].ctypes.data,
                           bufs[2].ctypes.data, None, None, C.cast(tot, C.c_void_p))
        best = ms if best is None else min(best, ms)
    return best, rows * w


def cpu_sample(scene_name, w, h, ystride, lb=None, want_totals=True):
    """cpu_baseline leg of the b200 arm: the reference (or the port) on every `ystride`-th scanline.
    Returns (ms, rays, kind, cores, totals); totals = the oracle port's evaluation counts of the sample."""
    import oracle_lib as ol

    cores = ol.nthreads()
    rows = (h + ystride - 1) // ystride
    rays = rows * w
    totals = None
    scene = load_scene(lb, scene_name) if lb is not None else None
    if want_totals and scene is not None:
        totals = ol.port_render(scene, w, h, ystride=ystride)["totals"]  # also warms the threads up
    if ol.have_ref() and scene_name != "synthetic_csg":  # the CSG scene uses extension nodes
        ms, _ = ref_probe_sample(scene_name, w, h, ystride)
        return ms, rays, "reference", cores, totals
    ms = ol.port_render(scene, w, h, ystride=ystride)["ms"]
    return ms, rays, "port", cores, totals


def cpu_baseline_variants(scene_name, w, h):
    """The reference's naive renderer on ONE thread, and built without optimisation (its Makefile passes no
    -O flag, Makefile:3) on all threads: small samples of the same frame, a second or so each."""
    import oracle_lib as ol

    if not ol.have_ref() or scene_name.startswith("synthetic"):
        return None
    res = {}
    rs = ol.RefScene(path=scene_path(scene_name))
    stride = max(1, h // 24)
    rows = (h + stride - 1) // stride
    ms = rs.probe(w, h, ystride=stride, threads=1)["ms"]
    res["one_thread_O2"] = {"value": rows * w / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "cores": 1,
                            "sample": f"every {stride}th scanline ({rows * w} rays, {ms:.0f} ms)"}
    if os.path.exists(ol.REF_O0_PATH):
        r0 = ol.RefScene(path=scene_path(scene_name), lib=ol.ref(ol.REF_O0_PATH))
        stride = max(1, h // 96)
        rows = (h + stride - 1) // stride
        r0.probe(w, h, ystride=stride * 4)  # warm the threads
        ms = r0.probe(w, h, ystride=stride)["ms"]
        res["all_threads_O0"] = {"value": rows * w / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "cores": ol.nthreads(),
                                 "sample": f"every {stride}th scanline ({rows * w} rays, {ms:.0f} ms), "
                                           "gcc -O0 as the reference Makefile builds it"}
    return res


def cpu_jit_equivalent(lb, scene_name, w, h, ystride):
    """JIT-equivalent CPU renderer (stand-in for tracing_jit_renderer.dasc) on the sample."""
    import tempfile

    import oracle_lib as ol

    scene = load_scene(lb, scene_name)
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.perf_counter()
        keep = ol.specialised_sdf(scene, tmp)
        compile_ms = (time.perf_counter() - t0) * 1e3
        ms = min(ol.port_render(scene, w, h, ystride=ystride, mode=2)["ms"] for _ in range(3))
        del keep
    rows = (h + ystride - 1) // ystride
    rays = rows * w
    return {"value": rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "cores": ol.nthreads(), "kind": "port",
            "what": "oracle pipeline + per-scene straight-line sdf from the lowering, g++ -O2 "
                    "(stand-in for the DynASM JIT, which needs Lua to build)",
            "sample": f"every {ystride}th scanline of the same {w}x{h} frame ({rays} rays, {ms:.0f} ms wall)",
            "ms_per_frame_extrapolated": ms * (w * h / rays), "specialise_and_compile_ms": compile_ms}


def run_reference(args, w, h):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.

    Stock code path: the UNMODIFIED render_thread() of naive_renderer.c under main.c's frame protocol, worker
    threads created once and kept over all W + K frames, scene built by the reference's scene.c.  The product
    library is never loaded in this arm.  Only when whole frames would take longer than the budget (the
    1024-primitive scene: ~1 h per 4K frame) does a step become a bounded sample of scanlines through the
    same static functions (lolref_probe)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    build_checkers()
    import oracle_lib as ol

    cores = ol.nthreads()
    frames = args.steps + args.warmup
    budget_ms = 240e3
    kind = "reference"
    if not ol.have_ref():
        # oracle/_ref did not travel: the oracle port (plain-C restatement, bit-pinned to the reference by
        # tests/test_oracle_pin.py) stands in, parsed by the port library's own copy of the front-end
        kind = "port"
        global ref_probe_sample

        def ref_probe_sample(scene_name, w, h, ystride, repeats=1, threads=None):  # noqa: F811
            return port_probe_sample(scene_name, w, h, ystride, repeats)
    probe_stride = 64
    ms, rays = ref_probe_sample(args.scene, w, h, probe_stride)
    est_frame_ms = ms * (w * h / rays)
    if kind == "reference" and est_frame_ms * frames <= budget_ms:
        per_frame = ref_protocol_frames(args.scene, w, h, frames)
        timed = per_frame[args.warmup:]
        ms_per_step = sum(timed) / len(timed)
        rays = w * h
        sample = (f"every scanline of the {w}x{h} frame ({rays} primary rays per step) through the unmodified "
                  f"render_thread() of naive_renderer.c under main.c's semaphore protocol, {cores} worker threads "
                  f"kept over all {frames} frames, gcc -O2 (the reference Makefile sets no -O level)")
        path = "render_thread"
    else:
        stride = 1
        while stride < h and est_frame_ms / stride * frames > budget_ms:
            stride *= 2
        for _ in range(args.warmup):
            ref_probe_sample(args.scene, w, h, stride)
        total = 0.0
        for _ in range(args.steps):
            ms, rays = ref_probe_sample(args.scene, w, h, stride)
            total += ms
        ms_per_step = total / args.steps
        what = ("naive_renderer.c's own static pipeline functions compiled unmodified" if kind == "reference"
                else "the oracle port (oracle/_ref did not travel)")
        sample = (f"every {stride}th scanline of the {w}x{h} frame ({rays} primary rays per step; a whole frame "
                  f"would take ~{est_frame_ms / 1e3:.0f} s), {what}, {cores} threads pulling scanlines from one "
                  f"atomic counter")
        path = "probe"
    value = rays / (ms_per_step * 1e-3) / 1e6
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "ms_per_frame_extrapolated": ms_per_step * (w * h / rays),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(args.scene, w, h), "sample": sample,
                                        "reference_path": path},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


# --------------------------------------------------------------- FLOP model --


# ncu --set full summary of the default scene4 4K launch (tools/ncu_summary.py), committed per round
NCU_SUMMARY = "r02_v1_scene4_4k_ball_test.txt"


def _ncu_applies(args, w, h, world):
    return world == 1 and args.scene == "scene4" and (w, h) == (3840, 2160) and args.workload == "frame" \
        and not args.opts and args.variant == 0 and args.arith == "exact"


def ncu_dram_traffic(args, w, h, world):
    """DRAM bytes of one launch from the committed ncu summary of this very workload, else None."""
    if not _ncu_applies(args, w, h, world):
        return None
    try:
        total, launches, scale = 0.0, 0, {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for line in open(os.path.join(ROOT, "profiles", NCU_SUMMARY)):
            launches += line.startswith("kernel ")
            if line.strip().startswith("DRAM bytes"):
                val, unit = line.split("[")[0].split()[-2:]
                total += float(val) * scale[unit]
        return total / max(launches, 1) or None  # per launch: the summary holds one block per captured launch
    except Exception:
        return None


def _ncu_line(args, w, h, world, marker, pick):
    if not _ncu_applies(args, w, h, world):
        return None
    try:
        for line in open(os.path.join(ROOT, "profiles", NCU_SUMMARY)):
            if marker in line:
                return pick(line)
    except Exception:
        pass
    return None


def ncu_hw_flop_frac(args, w, h, world):
    """FP32 FLOPs as the hardware counts them (ncu: fadd + fmul + 2 ffma thread-instructions per cycle, of the
    chip's peak) from the committed capture: every executed FP32 instruction, including the Newton steps of
    sqrt, the box tests and the range guards that the algorithmic model does not credit."""
    return _ncu_line(args, w, h, world, "hardware FP32 FLOP/cycle", lambda l: float(l.split("=")[1].split("%")[0]) / 100.0)


def ncu_issue_utilisation(args, w, h, world):
    """smsp__issue_active % of the same workload from the committed ncu capture: the kernel is bound by
    instruction issue (scalar FP32 that cannot fuse in exact mode), which is what FLOP fractions miss."""
    return _ncu_line(args, w, h, world, "SM issue-slot utilisation", lambda l: float(l.split("%")[1].split()[0]) / 100.0)


def flops_model(f_sdf, n_lights, pixels, primary, normal, shadow, shaded, rays_marched, rays_culled):
    """SURVEY.md 8d convention: F = E*F_sdf + 9*n_primary + 12*n_shadow + 56 (normal
    assembly, per shaded pixel) + 95 per light shaded (+20 for a culled one: L-p,
    normalise, n.l) + 50 per pixel (camera ray, ambient, clamp, gamma, pack)."""
    return (f_sdf * (primary + normal + shadow) + 9 * primary + 12 * shadow + 56 * shaded +
            95 * rays_marched + 20 * rays_culled + 50 * pixels)


# ------------------------------------------------------------ the GPU job --


class Job:
    """One process per GPU.  Holds what every workload of this run shares: rank 0's frame (big enough for
    every config, exported to the other ranks through CUDA IPC, completion flags behind it), the compact shard
    buffers of the NCCL path, the L2 flush buffer."""

    def __init__(self, args, torch, dist, lb, max_pixels):
        self.args, self.torch, self.dist, self.lb = args, torch, dist, lb
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = torch.device("cuda", self.local_rank)
        self.stream = torch.cuda.current_stream().cuda_stream
        self.max_pixels = max_pixels
        self.seq = 0  # frame number published through the completion flags (all ranks count alike)
        self.launches = 0
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.cpu_group = dist.new_group(backend="gloo") if self.world > 1 else None
        # rank 0's frame + flags; other ranks map it
        words = max_pixels + FLAG_BYTES // 4
        self.frame_store = torch.zeros(words, dtype=torch.int32, device=self.dev) if self.rank == 0 else None
        self.frame_base = self.frame_store.data_ptr() if self.rank == 0 else None
        self.peer_base = self.frame_base
        if self.world > 1:
            handle = torch.zeros(64, dtype=torch.uint8)
            if self.rank == 0:
                hb = (C.c_uint8 * 64)()
                rc = lb.lib().lolb200_ipc_export(self.frame_base, C.byref(hb))
                assert rc == 0, lb.lib().lolb200_last_error()
                handle = torch.tensor(list(hb), dtype=torch.uint8)
            hdev = handle.to(self.dev)
            dist.broadcast(hdev, 0)
            if self.rank != 0:
                hb = (C.c_uint8 * 64)(*hdev.cpu().tolist())
                p = C.c_void_p()
                rc = lb.lib().lolb200_ipc_open(C.byref(hb), C.byref(p))
                assert rc == 0, lb.lib().lolb200_last_error()
                self.peer_base = p.value
            self.token = torch.zeros(1, dtype=torch.int32, device=self.dev)
            # a compact shard of any w x h frame with w * h <= max_pixels, w <= 8192: the rank's bands, padded
            shard_max = max_pixels // self.world + 2 * 4 * 8192
            self.shard_store = torch.zeros(shard_max, dtype=torch.int32, device=self.dev)
            self.gathered_store = (torch.zeros(self.world * shard_max, dtype=torch.int32, device=self.dev)
                                   if self.rank == 0 else None)

    def flag_addr(self, base, r):
        return base + self.max_pixels * 4 + 128 * r

    def frame_view(self, w, h):
        return self.frame_store[: w * h].view(h, w)

    def sync_all(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        if self.world == 1:
            return [float(v) for v in values]
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def gather_list(self, values):
        """Per-rank lists of floats on rank 0 (None elsewhere)."""
        if self.world == 1:
            return [list(map(float, values))]
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [[float(x) for x in o] for o in out]

    # -- one frame, complete in rank 0's HBM when rank 0's stream drains --
    def step(self, renderer, w, h, cam=None, gather=None, aux=None):
        lb, world, rank = self.lb, self.world, self.rank
        gather = gather or self.args.gather
        if world == 1:
            renderer.render_device(self.frame_base, w, h, camera=cam, aux=aux, stream=self.stream)
            self.launches += 1
        elif gather == "peer":
            self.seq += 1
            if rank == 0:
                renderer.render_device(self.frame_base, w, h, camera=cam, pitch_px=w, aux=aux, stream=self.stream,
                                       shard=lb.Shard(rank=0, world=world, dst_full_frame=1))
                for r in range(1, world):  # stream memory operations: no kernel, no collective
                    lb.stream_wait_value32(self.stream, self.flag_addr(self.frame_base, r), self.seq)
            else:
                renderer.render_device(self.peer_base, w, h, camera=cam, pitch_px=w, aux=aux, stream=self.stream,
                                       shard=lb.Shard(rank=rank, world=world, dst_full_frame=1,
                                                      done_flag=self.flag_addr(self.peer_base, rank),
                                                      done_value=self.seq & 0xFFFFFFFF))
            self.launches += 1
        elif gather == "peer-allreduce":
            renderer.render_device(self.peer_base, w, h, camera=cam, pitch_px=w, aux=aux, stream=self.stream,
                                   shard=lb.Shard(rank=rank, world=world, dst_full_frame=1))
            self.launches += 1
            self.dist.all_reduce(self.token)  # every rank's stores are done before rank 0 goes on
        else:  # nccl
            shard_px = lb.shard_pixels(w, h, world)
            local = self.gathered_store[:shard_px] if rank == 0 else self.shard_store[:shard_px]
            renderer.render_device(local.data_ptr(), w, h, camera=cam, pitch_px=w, aux=aux, stream=self.stream,
                                   shard=lb.Shard(rank=rank, world=world))
            self.launches += 1
            glist = ([self.gathered_store[i * shard_px:(i + 1) * shard_px] for i in range(world)]
                     if rank == 0 else None)
            self.dist.gather(local, glist, dst=0)
            if rank == 0:
                lb.deinterleave(self.gathered_store.data_ptr(), self.frame_base, w, h, world, shard_px,
                                stream=self.stream)
                self.launches += 1

    def kernel_only(self, renderer, w, h, cam=None, aux=None):
        """The render kernel alone (this rank's shard), no completion wait."""
        lb = self.lb
        if self.world == 1:
            renderer.render_device(self.frame_base, w, h, camera=cam, aux=aux, stream=self.stream)
        else:
            renderer.render_device(self.shard_store.data_ptr(), w, h, camera=cam, pitch_px=w, aux=aux,
                                   stream=self.stream, shard=lb.Shard(rank=self.rank, world=self.world))

    def time_steps(self, fn, n, flush=True):
        """n calls of fn(i) timed with CUDA events on the launching stream, L2 flushed (untimed) between
        them, bracketed by barrier + synchronize; returns this rank's total ms."""
        torch = self.torch
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        self.sync_all()
        for i in range(n):
            if flush:
                self.flush.fill_(i & 0xFF)
            ev[i][0].record()
            fn(i)
            ev[i][1].record()
        self.sync_all()
        return float(sum(a.elapsed_time(b) for a, b in ev))


def time_config(job, renderer, w, h, cams, n_frames, warmup=3):
    """ms per frame (max over ranks) of n_frames frames with cameras cycling through `cams`."""
    cams = cams or [None]
    for i in range(warmup):
        job.step(renderer, w, h, cam=cams[i % len(cams)])
    total = job.time_steps(lambda i: job.step(renderer, w, h, cam=cams[i % len(cams)]), n_frames)
    return job.max_over_ranks([total])[0] / n_frames


def orbit_throughput(job, lb, scene, renderer, w, h, n_frames, steps, warmup=1):
    """Config C5, throughput mode: one step = n_frames orbit frames; whole frames are dealt to the ranks
    (frame k -> rank k % N) and stay in that rank's HBM: no data-path collective.  Returns ms per step."""
    from loltracer_b200 import scenegen

    torch = job.torch
    cams = [scenegen.orbit_camera(scene.camera, k, n_frames) for k in range(n_frames)]
    mine = [k for k in range(n_frames) if k % job.world == job.rank]
    frames = torch.zeros((max(1, len(mine)), h, w), dtype=torch.int32, device=job.dev)  # all resident

    def step(_):
        for i, k in enumerate(mine):
            renderer.render_device(frames[i].data_ptr(), w, h, camera=cams[k], stream=job.stream)
        job.launches += len(mine)

    for _ in range(warmup):
        step(0)
    total = job.time_steps(step, steps)
    ms = job.max_over_ranks([total])[0] / steps
    del frames
    return ms, cams, mine


def launch_probes(job, renderer, w, h, n=8):
    """tail_us (work queue dry -> last warp's exit) and span_us (first CTA's start -> last warp's exit) of this
    rank's render kernel, from the kernel's own global-timer probes (lolb200_aux.launch_timing), mean of n."""
    import numpy as np
    torch, lb = job.torch, job.lb
    init = torch.tensor([-1, 0, -1], dtype=torch.int64, device=job.dev)  # ~0, 0, ~0 as u64
    probes = torch.zeros((n, 3), dtype=torch.int64, device=job.dev)
    for i in range(n):
        probes[i].copy_(init)
        aux = lb.Aux(launch_timing=probes[i].data_ptr())
        job.flush.fill_(i)
        job.kernel_only(renderer, w, h, aux=aux)
    job.sync_all()
    p = probes.cpu().numpy().view(np.uint64).astype(np.float64)
    return float(np.mean(p[:, 1] - p[:, 0]) / 1e3), float(np.mean(p[:, 1] - p[:, 2]) / 1e3)


def run_orbit(args, job, lb, scene, renderer, w, h):
    """--workload orbit: config C5 as the headline line."""
    F, K, W = args.orbit_frames, args.steps, max(args.warmup, 3)
    torch, dist = job.torch, job.dist
    job.launches = 0
    with ClockSampler(job.local_rank) as clocks:
        ms_per_step, cams, mine = orbit_throughput(job, lb, scene, renderer, w, h, F, K, warmup=W)
    n_launches = K * len(mine)
    host = torch.empty((h, w), dtype=torch.int32).pin_memory()
    for k in mine[:2]:
        renderer.render_host(host.data_ptr(), w, h, camera=cams[k])
    job.sync_all()
    reps = max(1, K // 4)
    t0 = time.perf_counter()
    for _ in range(reps):
        for k in mine:
            renderer.render_host(host.data_ptr(), w, h, camera=cams[k])
    job.sync_all()
    e2e_ms = job.max_over_ranks([(time.perf_counter() - t0) / reps * 1e3])[0]
    if job.rank == 0:
        rays = F * w * h
        print(json.dumps({
            "metric": METRIC, "value": rays / (ms_per_step * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": job.world,
            "steps": K, "warmup": W, "ms_per_step": ms_per_step, "ms_per_frame": ms_per_step / F,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{F}-frame camera orbit of {args.scene}.lol at {w}x{h} (BASELINE config C5), "
                                   f"whole frames dealt to ranks, frames left in each rank's HBM",
                       "arith": args.arith, "variant": args.variant,
                       "l2": "256 MB write between timed steps (untimed)", "kernel": renderer.kernel_info()},
            "e2e": {"value": rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_frame": e2e_ms / F,
                    "h2d_bytes_per_step": 192 * F, "d2h_bytes_per_step": rays * 4},
            "gpu_launches": n_launches,
            "clocks": clocks.summary(),
        }))
    if job.world > 1:
        dist.destroy_process_group()
    return 0


def per_config(job, lb, opt, main_scene, main_renderer, main_ms):
    """Every BASELINE.json config at this N (GPU only; the CPU column is the reference arm's business):
    C1 scene.lol 320x240, C2 scene2/scene3 1920x1080, C3 = the headline, C4 synthetic 1024 spheres 4K,
    C5 64-frame orbit of scene4 at 7680x4320 (throughput mode)."""
    out = {}

    def one(key, scene_name, w, h, frames):
        scene = main_scene if scene_name == "scene4" else load_scene(lb, scene_name)
        r = main_renderer if scene_name == "scene4" else lb.Renderer(scene, opt, device=job.local_rank)
        ms = time_config(job, r, w, h, None, frames)
        out[key] = {"ms_per_frame": ms, "mrays_s": w * h / ms / 1e3}
        if r is not main_renderer:
            r.close()

    one("c1_scene_320x240", "scene", 320, 240, 30)
    one("c2_scene2_1920x1080", "scene2", 1920, 1080, 30)
    one("c2_scene3_1920x1080", "scene3", 1920, 1080, 30)
    out["c3_scene4_3840x2160"] = {"ms_per_frame": main_ms, "mrays_s": 3840 * 2160 / main_ms / 1e3,
                                  "note": "the headline line"}
    one("c4_synthetic1024_3840x2160", "synthetic", 3840, 2160, 4)
    ms, _, _ = orbit_throughput(job, lb, main_scene, main_renderer, 7680, 4320, 64, 2)
    out["c5_orbit64_scene4_7680x4320"] = {"ms_per_step": ms, "ms_per_frame": ms / 64,
                                          "mrays_s": 64 * 7680 * 4320 / ms / 1e3,
                                          "mode": "whole frames dealt to ranks, no data-path collective"}
    return out


def inprocess_group(job, scene_name, w, h):
    """The single-process renderer.h drop-in (b200_renderer.c under the headless twin of main.c) driving all N
    GPUs of the box: N worker threads, each pulling one GPU's share (lolb200_group_share_*), frame in host
    memory.  Run by rank 0 as a subprocess while the other ranks wait on a CPU (gloo) barrier."""
    host = os.path.join(ROOT, "loltracer_b200", "backend", "build", "lol_headless_b200")
    res = None
    if job.rank == 0:
        if not os.path.exists(host):
            res = {"unavailable": "loltracer_b200/backend/build/lol_headless_b200 did not travel"}
        else:
            res = {}
            for tag, extra in (("pinned_surface", ["--pin-surface"]), ("pageable_surface", [])):
                try:
                    cmd = [host, str(max(job.world, 2)), scene_path(scene_name), "--gpus", str(job.world), "--gather",
                           "host", "--size", f"{w}x{h}", "--frames", "30", "--warmup", "8"] + extra
                    env = {k: v for k, v in os.environ.items() if not k.startswith(("RANK", "LOCAL_RANK", "WORLD_SIZE"))}
                    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
                    line = [l for l in out.stdout.splitlines() if "min " in l and "avg" in l]
                    if out.returncode != 0 or not line:
                        res[tag] = {"unavailable": (out.stderr or out.stdout)[-300:]}
                        continue
                    f = line[-1].replace("\t", " ").split()
                    res[tag] = {"min_ms": float(f[f.index("min") + 1]), "avg_ms": float(f[f.index("avg") + 1]),
                                "max_ms": float(f[f.index("max") + 1])}
                except Exception as e:
                    res[tag] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
            res["command"] = "lol_headless_b200 <N threads> scene4.lol --gpus N --gather host [--pin-surface]"
    if job.world > 1:
        job.dist.barrier(group=job.cpu_group)
    return res


# ------------------------------------------------------------------- main --


def main():
    args = parse_args()
    if args.size is None:
        args.size = "7680x4320" if args.workload == "orbit" else "3840x2160"
    w, h = (int(x) for x in args.size.lower().split("x"))
    if args.impl == "reference":
        return run_reference(args, w, h)

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry
    import loltracer_b200 as lb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(1, args.gpus) and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: liblolb200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()

    extras = not args.no_extras and args.workload == "frame" and args.scene == "scene4" and (w, h) == (3840, 2160)
    job = Job(args, torch, dist, lb, max(w * h, 7680 * 4320 if extras else 0))
    scene = load_scene(lb, args.scene)
    kw = {k: int(v) for k, v in (kv.split("=") for kv in args.opts.split(",") if kv)}
    opt = lb.Options.default(variant=args.variant, arith=1 if args.arith == "fast" else 0, **kw)
    renderer = lb.Renderer(scene, opt, device=local_rank)
    stream = job.stream
    K, W = args.steps, max(args.warmup, 3)
    if args.workload == "orbit":
        return run_orbit(args, job, lb, scene, renderer, w, h)

    frame = job.frame_view(w, h) if rank == 0 else None
    for _ in range(W):
        job.step(renderer, w, h)
    job.sync_all()

    # ---- timed region: K steps, device time, L2 flushed between steps (untimed) ----
    job.launches = 0
    with ClockSampler(local_rank) as clocks:
        t_wall0 = time.perf_counter()
        total_ms = job.time_steps(lambda i: job.step(renderer, w, h), K)
        t_wall = time.perf_counter() - t_wall0
    n_launches = job.launches
    # kernel-only duration (this rank's render kernel alone), for the roofline
    kernel_ms = job.time_steps(lambda i: job.kernel_only(renderer, w, h), K) / K
    total_ms, kernel_ms = job.max_over_ranks([total_ms, kernel_ms])
    ms_per_step = total_ms / K
    value = w * h / (ms_per_step * 1e-3) / 1e6

    # ---- the sharded frame must be the single-GPU frame, bit for bit ----
    verified = None
    if world > 1:
        job.step(renderer, w, h)
        job.sync_all()
        if rank == 0:
            check = torch.zeros((h, w), dtype=torch.int32, device=dev)
            renderer.render_device(check.data_ptr(), w, h, stream=stream)
            torch.cuda.synchronize()
            verified = bool(torch.equal(check, frame))
            del check

    # ---- where a step's time goes (instrumented passes, untimed for the headline) ----
    tail_us, span_us = launch_probes(job, renderer, w, h)
    per_rank = job.gather_list([tail_us, span_us])
    scaling_extras = {}
    if world > 1:
        # completion barrier on rank 0: own kernel's end -> every rank's flag seen
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        waits = []
        for i in range(10):
            job.sync_all()
            job.flush.fill_(i)
            job.seq += 1
            sh = lb.Shard(rank=rank, world=world, dst_full_frame=1)
            if rank != 0:
                sh.done_flag, sh.done_value = job.flag_addr(job.peer_base, rank), job.seq & 0xFFFFFFFF
            renderer.render_device(job.peer_base, w, h, pitch_px=w, shard=sh, stream=stream)
            if rank == 0:
                e0.record()
                for r in range(1, world):
                    lb.stream_wait_value32(stream, job.flag_addr(job.frame_base, r), job.seq)
                e1.record()
                torch.cuda.synchronize()
                waits.append(e0.elapsed_time(e1) * 1e3)
        job.sync_all()
        # one frame's latency: all ranks start together (barrier), launch -> complete on rank 0
        lat = []
        for i in range(10):
            job.sync_all()
            e0.record()
            job.step(renderer, w, h, gather="peer")
            e1.record()
            torch.cuda.synchronize()
            lat.append(e0.elapsed_time(e1))
        job.sync_all()
        lat_ms = job.max_over_ranks([sum(lat) / len(lat)])[0]
        # the other gathers, same K steps
        others = {}
        for g in ("nccl", "peer-allreduce", "peer"):
            if g == args.gather:
                continue
            for _ in range(3):
                job.step(renderer, w, h, gather=g)
            t = job.time_steps(lambda i: job.step(renderer, w, h, gather=g), K)
            others[g] = job.max_over_ranks([t])[0] / K
        # round 1's completion barrier for comparison: kernel end -> 4-byte NCCL all-reduce done
        ar = []
        for i in range(10):
            job.sync_all()
            job.kernel_only(renderer, w, h)
            e0.record()
            dist.all_reduce(job.token)
            e1.record()
            torch.cuda.synchronize()
            ar.append(e0.elapsed_time(e1) * 1e3)
        ar_us = job.max_over_ranks([sum(ar) / len(ar)])[0]
        scaling_extras = {
            "barrier_us": {"flags_rank0_mean": (sum(waits) / len(waits)) if rank == 0 else None,
                           "what": "rank 0's stream: own kernel's end -> the frame numbers of all other ranks seen "
                                   "(cuStreamWaitValue32 on flag words the ranks' last CTAs store over NVLink); "
                                   "includes waiting for the slowest rank",
                           "nccl_allreduce_4B_mean_max_over_ranks": ar_us},
            "frame_latency_ms": lat_ms,
            "gather_nccl_ms": others.get("nccl", ms_per_step if args.gather == "nccl" else None),
            "gather_peer_allreduce_ms": others.get("peer-allreduce"),
            "gather_peer_flags_ms": others.get("peer", ms_per_step if args.gather == "peer" else None),
        }

    # ---- e2e: host buffers, D2H inside the timed region ----
    # N = 1: lolb200_render_host, the call b200_renderer.c makes.  N > 1: the frame has to end
    # up in HOST memory, so by default no GPU gathers anything: every rank renders its bands and
    # its own copy engine writes them into one POSIX shared-memory frame (rank 0 created it, all
    # ranks map and pin it): 1/N of the bytes per PCIe link instead of all of them through rank
    # 0's (lolb200_render_host_shard; the in-process twin is `--gather host` of the C backend).
    # A frame is complete when every rank has written its frame number into its slot of a shared
    # page, which rank 0 polls: no NCCL call inside the timed region.
    host = torch.empty((h, w), dtype=torch.int32).pin_memory() if rank == 0 else None
    shared, shared_path, e2e_verified, slots = None, None, None, None
    if world > 1 and args.e2e_path == "host-shards":
        shared_path = f"/dev/shm/lolb200_bench_{os.environ.get('MASTER_PORT', '0')}_{w}x{h}"
        if rank == 0:
            np.memmap(shared_path, dtype=np.uint32, mode="w+", shape=(h * w + 1024,)).flush()
        dist.barrier()
        whole = np.memmap(shared_path, dtype=np.uint32, mode="r+", shape=(h * w + 1024,))
        shared = whole[: h * w].reshape(h, w)
        slots = whole[h * w:]
        lb.surface_pin(shared.ctypes.data, h * w * 4)  # this process owns the mapping for the run
    e2e_seq = [0]
    shard = lb.Shard(rank=rank, world=world, band_rows=0, dst_full_frame=0)

    def e2e_step():
        if world == 1:
            renderer.render_host(host.data_ptr(), w, h)
        elif shared is not None:
            renderer.render_host_shard(shared.ctypes.data, w, h, shard)  # returns when this rank's rows are in
            e2e_seq[0] += 1
            slots[rank * 16] = e2e_seq[0]
            if rank == 0:                                                # the frame is complete for its consumer
                while int(slots[: world * 16: 16].min()) < e2e_seq[0]:
                    pass
        else:
            job.step(renderer, w, h)
            if rank == 0:
                host.copy_(frame, non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(3):
        e2e_step()
    job.sync_all()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    job.sync_all()
    e2e_ms = job.max_over_ranks([(time.perf_counter() - t0) / K * 1e3])[0]
    e2e_value = w * h / (e2e_ms * 1e-3) / 1e6
    e2e_sweep = {}
    if shared is not None and args.e2e_slabs:
        for v in args.e2e_slabs.split(","):
            os.environ["LOLB200_SHARD_SLABS"] = v.strip()
            for _ in range(3):
                e2e_step()
            job.sync_all()
            t0 = time.perf_counter()
            for _ in range(K):
                e2e_step()
            job.sync_all()
            e2e_sweep[v.strip()] = job.max_over_ranks([(time.perf_counter() - t0) / K * 1e3])[0]
        os.environ.pop("LOLB200_SHARD_SLABS", None)
    if shared is not None:
        if rank == 0:
            e2e_verified = bool(np.array_equal(np.asarray(shared), frame.cpu().numpy().view(np.uint32)))
        dist.barrier()
        lb.surface_unpin(shared.ctypes.data)
        del shared, slots, whole
        if rank == 0:
            os.unlink(shared_path)

    # ---- executed work (instrumented twin of the kernel, untimed) ----
    copt = lb.Options.default(variant=args.variant, arith=1 if args.arith == "fast" else 0, counters=1, **kw)
    crend = lb.Renderer(scene, copt, device=local_rank)
    shard_px = lb.shard_pixels(w, h, world)
    scratch = torch.zeros((h, w) if world == 1 else (shard_px,), dtype=torch.int32, device=dev)
    crend.render_device(scratch.data_ptr(), w, h, pitch_px=w,
                        shard=None if world == 1 else lb.Shard(rank=rank, world=world), stream=stream)
    torch.cuda.synchronize()
    cnt = crend.read_counters()
    crend.close()
    del scratch
    f_sdf = scene.flops_per_eval()
    n_lights = scene.struct.n_lights
    exec_flops = flops_model(f_sdf, n_lights, cnt["pixels"], cnt["primary_evals"], cnt["normal_evals"],
                             cnt["shadow_evals"], cnt["normal_evals"] // 4, cnt["shadow_rays"],
                             cnt["shadow_rays_culled"])
    # objects that a box test skipped inside an evaluation were not executed
    exec_flops -= cnt.get("skipped_flops", 0)
    exec_flops = job.max_over_ranks([exec_flops])[0]  # the slowest rank bounds the frame

    # ---- GPU-only extras: every BASELINE config at this N, a moving camera, the in-process drop-in ----
    extra = {}
    if extras:
        from loltracer_b200 import scenegen
        extra["per_config"] = per_config(job, lb, opt, scene, renderer, ms_per_step)
        cams = [scenegen.orbit_camera(scene.camera, k, 64) for k in range(64)]
        extra["moving_camera_ms"] = time_config(job, renderer, w, h, cams, 64, warmup=8)
        # the same 64 cameras, each one rendered three times in a row and the third timed: what the orbit
        # costs when every frame repeats its predecessor (the bench's own best case)
        rep_total = 0.0
        for cam in cams[::4]:
            for _ in range(2):
                job.step(renderer, w, h, cam=cam)
            rep_total += job.time_steps(lambda i: job.step(renderer, w, h, cam=cam), 1)
        extra["moving_camera_repeated_frames_ms"] = job.max_over_ranks([rep_total])[0] / len(cams[::4])
        extra["moving_camera_note"] = ("moving_camera_ms: mean ms/frame over the 64-frame orbit at this size through "
                                       "the same path as the headline, a new camera every frame (the longest-first "
                                       "chunk order is one to eight frames stale; the lowering's box tests were chosen "
                                       "for the FILE camera).  moving_camera_repeated_frames_ms: every 4th of those "
                                       "cameras rendered three times in a row, the third timed (the order is re-sorted every "
                                       "eighth launch whatever the camera does).  ms_per_step is the file camera alone")
        if world > 1:
            extra["inprocess_group_ms"] = inprocess_group(job, args.scene, w, h)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak_tf, _ = lb.measure_fp32_peak(local_rank)
    achieved_tf = exec_flops / (kernel_ms * 1e-3) / 1e12
    roofline = {
        "bound": "fp32", "kernel": "lol_render (NVRTC, per scene)",
        "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
        "peak_source": "measured here: lolb200_measure_fp32_peak (independent FFMA chains); "
                       "MEASURED_PEAKS.json has no FP32 entry; nominal 74.45",
        "frac_of_nominal": achieved_tf / FP32_NOMINAL_TFLOPS,
        "flop_per_launch_executed": exec_flops, "kernel_ms": kernel_ms,
        "traffic": ncu_dram_traffic(args, w, h, world),
        "issue_slot_utilisation_ncu": ncu_issue_utilisation(args, w, h, world),
        "hw_fp32_frac_ncu": ncu_hw_flop_frac(args, w, h, world),
        "ncu_summary": f"profiles/{NCU_SUMMARY}",
        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one lol_render launch from the committed "
                        "ncu --set full capture (profiles/); the 33 MB frame stays in the 126 MB L2, so DRAM sees "
                        "only KBs -- algorithmic HBM bytes are 4 per pixel",
        "algorithmic_hbm_bytes": w * h * 4,
        "hbm_write_gbs": (w * h * 4 / world) / (kernel_ms * 1e-3) / 1e9,
    }
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        roofline["hbm_frac_of_measured"] = roofline["hbm_write_gbs"] / peaks["hbm_gbs"]
    except Exception:
        pass

    out = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "ms_per_frame": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.scene, w, h),
                   "result": "frame complete in rank 0's HBM", "arith": args.arith, "variant": args.variant,
                   "opts": args.opts or None,
                   "sharding": "single GPU" if world == 1 else f"4-row bands cyclic over {world} ranks, "
                               f"gather={args.gather}",
                   "l2": "256 MB write between timed steps (untimed); the kernel reads no global inputs",
                   "kernel": renderer.kernel_info()},
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_frame": e2e_ms,
                "h2d_bytes_per_step": 192 * world, "d2h_bytes_per_step": w * h * 4,
                "path": ("lolb200_render_host: slab launches overlapped with the read-back into a pinned host frame"
                         if world == 1 else
                         "lolb200_render_host_shard on every rank: own bands over own PCIe link into one "
                         "shared-memory host frame; completion through frame numbers in a shared page"
                         if args.e2e_path == "host-shards" else
                         "frame gathered on rank 0 over NVLink, then one D2H copy from rank 0"),
                "host_frame_equals_single_gpu": e2e_verified,
                **({"slab_sweep_ms_per_frame": e2e_sweep} if e2e_sweep else {})},
        "gpu_launches": n_launches,
        "sharded_frame_equals_single_gpu": verified,
        "roofline": roofline,
        "clocks": clocks.summary(),
        "wall_ms_per_step_incl_flush": t_wall / K * 1e3,
        "executed_evals_per_pixel": {k: cnt[k] / max(1, cnt["pixels"]) for k in
                                     ("primary_evals", "normal_evals", "shadow_evals")},
        "tail_us": {"per_rank": [p[0] for p in per_rank],
                    "what": "render kernel: first moment a warp finds the work queue dry -> last warp's exit "
                            "(global-timer probes inside the kernel, instrumented pass)"},
        "kernel_span_us": {"per_rank": [p[1] for p in per_rank]},
    }
    out.update(scaling_extras)
    out.update(extra)

    if world == 1 and not args.no_cpu_baseline:
        # ~20 core-seconds on scene4: every 4th scanline of the same frame; the
        # 1024-primitive scene costs ~1000x more per ray, so only a few scanlines
        stride = 4 if not args.scene.startswith("synthetic") else max(4, h // 3)
        ms, rays, kind, cores, totals = cpu_sample(args.scene, w, h, stride, lb=lb)
        cpu_value = rays / (ms * 1e-3) / 1e6
        out["cpu_baseline"] = {
            "value": cpu_value, "unit": "Mrays/s", "cores": cores, "kind": kind,
            "sample": f"every {stride}th scanline of the same {w}x{h} frame ({rays} rays, {ms:.0f} ms "
                      f"wall), {cores} threads",
            "ms_per_frame_extrapolated": ms * (w * h / rays),
        }
        scale = w * h / rays
        ref_flops = flops_model(f_sdf, n_lights, w * h, totals["primary"] * scale, totals["normal"] * scale,
                                totals["shadow"] * scale, w * h, w * h * n_lights, 0)
        # The reference's own work for this frame per second of GPU time.  NOT a fraction of any peak (the
        # exact skips and box tests mean the GPU never executes these FLOPs): a useful-work rate.
        roofline["reference_work_tflops_equiv"] = ref_flops / (kernel_ms * 1e-3) / 1e12
        roofline["flop_per_launch_reference"] = ref_flops
        # one thread, and the reference as its own Makefile builds it (no -O flag)
        try:
            out["cpu_baseline"]["variants"] = cpu_baseline_variants(args.scene, w, h)
        except Exception as e:
            out["cpu_baseline"]["variants"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        # The reference's second CPU renderer, the DynASM tracing JIT, cannot be built in
        # this image (no Lua for the .dasc preprocessor).  Its stand-in: the same lowering's
        # straight-line distance code with baked constants, compiled by g++ and driven by the
        # oracle's pipeline on the same sample (bit-identical frame, tests/test_oracle_pin.py).
        try:
            out["cpu_jit_equivalent"] = cpu_jit_equivalent(lb, args.scene, w, h, stride)
        except Exception as e:  # a missing host compiler must not cost the GPU numbers
            out["cpu_jit_equivalent"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    if args.all_scenes and world == 1:
        per = {}
        for name in ("scene", "scene2", "scene3", "scene4"):
            r2 = lb.Renderer(load_scene(lb, name), opt, device=local_rank)
            ms = time_config(job, r2, w, h, None, 20)
            per[name] = {"ms_per_frame": ms, "mrays_s": w * h / ms / 1e3}
            r2.close()
        out["per_scene"] = per

    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
