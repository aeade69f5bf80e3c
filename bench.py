#!/usr/bin/env python
"""bench.py -- Mrays/s and ms/frame of the sphere-tracing hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--scene scene4] [--size 3840x2160]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      the reference's CPU renderer, host cores

A step is one frame of the workload (default: scene4.lol at 3840x2160, BASELINE
config C3).  With N > 1 the frame is sharded in 4-row bands, band b -> rank b % N;
every step ends with the complete frame on rank 0 (NCCL gather + de-interleave,
or --gather peer: ranks store straight into rank 0's frame over NVLink).

Prints ONE JSON line (rank 0).  `value` is whole-job Mrays/s with the frame left
in HBM; `e2e` is the same metric through the host-surface entry point
(lolb200_render_host: camera in, pixels copied into a host buffer).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "Mrays/s"
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.45, SURVEY.md 8d


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="scene4")
    ap.add_argument("--size", default=None, help="default 3840x2160 (frame) / 7680x4320 (orbit)")
    ap.add_argument("--gather", default="peer", choices=["nccl", "peer"])
    ap.add_argument("--e2e-path", default="host-shards", choices=["host-shards", "gather-then-copy"],
                    help="N > 1, e2e leg: every rank copies its own bands into one shared-memory host frame over "
                         "its own PCIe link (default), or the frame is gathered on rank 0 and copied from there")
    ap.add_argument("--workload", default="frame", choices=["frame", "orbit"],
                    help="frame: one frame per step, sharded by bands over the GPUs (configs C1-C4); "
                         "orbit: one step = 64 camera-orbit frames, whole frames dealt to the GPUs (C5)")
    ap.add_argument("--orbit-frames", type=int, default=64)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--arith", default="exact", choices=["exact", "fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--all-scenes", action="store_true",
                    help="also time the other example scenes (extra keys, same JSON line)")
    return ap.parse_args()


def load_scene(lb, name):
    if name in ("synthetic", "synthetic_csg"):
        from loltracer_b200 import scenegen
        return lb.Scene.from_string(scenegen.synthetic_scene_text(csg=name == "synthetic_csg"))
    path = name if os.path.exists(name) else os.path.join(ROOT, "tests", "golden", "scenes", name + ".lol")
    return lb.Scene.from_file(path)


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


def cpu_sample(scene_name, w, h, ystride, repeats=1, force_port=False, want_totals=True):
    """The reference's naive renderer (oracle/_ref, built from its own sources) -- or the
    oracle port when that library did not travel -- on every `ystride`-th scanline of the
    w x h frame, all host threads.  Returns (best_ms, rays, kind, cores, totals)."""
    import oracle_lib as ol
    import loltracer_b200 as lb

    cores = ol.nthreads()
    rows = (h + ystride - 1) // ystride
    rays = rows * w
    scene = load_scene(lb, scene_name)
    # evaluation counts of the sample (oracle port; also warms the threads up)
    totals = ol.port_render(scene, w, h, ystride=ystride)["totals"] if want_totals else None
    best = None
    # the CSG scene uses extension nodes the reference cannot hold: oracle port only
    if ol.have_ref() and not force_port and scene_name != "synthetic_csg":
        kind = "reference"
        if scene_name == "synthetic":
            from loltracer_b200 import scenegen
            rs = ol.RefScene(text=scenegen.synthetic_scene_text())
        else:
            path = scene_name if os.path.exists(scene_name) else os.path.join(
                ROOT, "tests", "golden", "scenes", scene_name + ".lol")
            rs = ol.RefScene(path=path)
        for _ in range(repeats):
            ms = rs.probe(w, h, ystride=ystride)["ms"]
            best = ms if best is None else min(best, ms)
    else:
        kind = "port"
        for _ in range(repeats):
            ms = ol.port_render(scene, w, h, ystride=ystride)["ms"]
            best = ms if best is None else min(best, ms)
    return best, rays, kind, cores, totals


def cpu_baseline_variants(scene_name, w, h):
    """The reference's naive renderer on ONE thread, and built without optimisation (its Makefile passes no
    -O flag, Makefile:3) on all threads: small samples of the same frame, a second or so each."""
    import oracle_lib as ol

    if not ol.have_ref() or scene_name.startswith("synthetic"):
        return None
    path = scene_name if os.path.exists(scene_name) else os.path.join(ROOT, "tests", "golden", "scenes", scene_name + ".lol")
    res = {}
    rs = ol.RefScene(path=path)
    stride = max(1, h // 24)
    rows = (h + stride - 1) // stride
    ms = rs.probe(w, h, ystride=stride, threads=1)["ms"]
    res["one_thread_O2"] = {"value": rows * w / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "cores": 1,
                            "sample": f"every {stride}th scanline ({rows * w} rays, {ms:.0f} ms)"}
    if os.path.exists(ol.REF_O0_PATH):
        r0 = ol.RefScene(path=path, lib=ol.ref(ol.REF_O0_PATH))
        stride = max(1, h // 96)
        rows = (h + stride - 1) // stride
        r0.probe(w, h, ystride=stride * 4)  # warm the threads
        ms = r0.probe(w, h, ystride=stride)["ms"]
        res["all_threads_O0"] = {"value": rows * w / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "cores": ol.nthreads(),
                                 "sample": f"every {stride}th scanline ({rows * w} rays, {ms:.0f} ms), "
                                           "gcc -O0 as the reference Makefile builds it"}
    return res


def cpu_jit_equivalent(scene_name, w, h, ystride):
    """JIT-equivalent CPU renderer (stand-in for tracing_jit_renderer.dasc) on the sample."""
    import tempfile

    import oracle_lib as ol
    import loltracer_b200 as lb

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
    """--impl reference: the reference's CPU implementation of the path, host cores only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as entry
    entry.build()
    # size the sample so that (steps + warmup) samples end within ~2 minutes
    probe_stride = 64
    ms, rays, kind, cores, _ = cpu_sample(args.scene, w, h, probe_stride)
    per_row_ms = ms / ((h + probe_stride - 1) // probe_stride)
    budget_ms = 100e3 / max(1, args.steps + args.warmup)
    stride = 1
    while stride < 64 and per_row_ms * ((h + stride - 1) // stride) > budget_ms:
        stride *= 2
    for _ in range(args.warmup):
        cpu_sample(args.scene, w, h, stride, want_totals=False)
    t_total, rays = 0.0, 0
    for _ in range(args.steps):
        ms, rays, kind, cores, _ = cpu_sample(args.scene, w, h, stride, want_totals=False)
        t_total += ms
    ms_per_step = t_total / args.steps
    value = rays / (ms_per_step * 1e-3) / 1e6
    sample = (f"{'every scanline' if stride == 1 else f'every {stride}th scanline'} of the {w}x{h} frame "
              f"({rays} primary rays per step), "
              f"{'naive_renderer.c compiled unmodified' if kind == 'reference' else 'oracle port'}, "
              f"{cores} threads pulling scanlines from one atomic counter")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "ms_per_frame_extrapolated": ms_per_step * (w * h / rays),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": f"{args.scene}.lol at {w}x{h}, one primary ray per pixel (the frame of the b200 arm, "
                                           f"rendered by the reference's CPU renderer)", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


# --------------------------------------------------------------- FLOP model --


# ncu --set full summary of the default scene4 4K launch (tools/ncu_summary.py), committed per round
NCU_SUMMARY = "r01_v1_final_scene4_4k.txt"


def ncu_dram_traffic(args, w, h, world):
    """DRAM bytes of one launch from the committed ncu summary of this very workload, else None."""
    if world != 1 or args.scene != "scene4" or (w, h) != (3840, 2160) or args.workload != "frame":
        return None
    path = os.path.join(ROOT, "profiles", NCU_SUMMARY)  # the kernel as it is benched today
    try:
        total, scale = 0.0, {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for line in open(path):
            if line.strip().startswith("DRAM bytes"):
                val, unit = line.split("[")[0].split()[-2:]
                total += float(val) * scale[unit]
        return total or None
    except Exception:
        return None


def ncu_hw_flop_frac(args, w, h, world):
    """FP32 FLOPs as the hardware counts them (ncu: fadd + fmul + 2 ffma thread-instructions per cycle, of the
    chip's peak) for the same workload, from the committed capture: every executed FP32 instruction, including
    the Newton steps of sqrt, the box tests and the range guards that the algorithmic model does not credit."""
    if world != 1 or args.scene != "scene4" or (w, h) != (3840, 2160) or args.workload != "frame":
        return None
    try:
        for line in open(os.path.join(ROOT, "profiles", NCU_SUMMARY)):
            if "hardware FP32 FLOP/cycle" in line:
                return float(line.split("=")[1].split("%")[0]) / 100.0
    except Exception:
        pass
    return None


def ncu_issue_utilisation(args, w, h, world):
    """smsp__issue_active % of the same workload from the committed ncu capture: the kernel is bound by
    instruction issue (scalar FP32 that cannot fuse in exact mode), which is what FLOP fractions miss."""
    if world != 1 or args.scene != "scene4" or (w, h) != (3840, 2160) or args.workload != "frame":
        return None
    try:
        for line in open(os.path.join(ROOT, "profiles", NCU_SUMMARY)):
            if "SM issue-slot utilisation" in line:
                return float(line.split("%")[1].split()[0]) / 100.0
    except Exception:
        pass
    return None


def flops_model(f_sdf, n_lights, pixels, primary, normal, shadow, shaded, rays_marched, rays_culled):
    """SURVEY.md 8d convention: F = E*F_sdf + 9*n_primary + 12*n_shadow + 56 (normal
    assembly, per shaded pixel) + 95 per light shaded (+20 for a culled one: L-p,
    normalise, n.l) + 50 per pixel (camera ray, ambient, clamp, gamma, pack)."""
    return (f_sdf * (primary + normal + shadow) + 9 * primary + 12 * shadow + 56 * shaded +
            95 * rays_marched + 20 * rays_culled + 50 * pixels)



def run_orbit(args, lb, torch, dist, scene, renderer, world, rank, local_rank, dev, w, h):
    """Config C5, throughput mode: one step = F camera-orbit frames of the scene; whole
    frames are dealt to the ranks (frame k -> rank k % N) and stay in that rank's HBM: no
    data-path collective.  value = F*W*H rays / step time (max over ranks)."""
    import time as _time
    from loltracer_b200 import scenegen

    F, K, W = args.orbit_frames, args.steps, max(args.warmup, 3)
    cams = [scenegen.orbit_camera(scene.camera, k, F) for k in range(F)]
    mine = [k for k in range(F) if k % world == rank]
    stream = torch.cuda.current_stream().cuda_stream
    frames = torch.zeros((max(1, len(mine)), h, w), dtype=torch.int32, device=dev)  # all resident
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        for i, k in enumerate(mine):
            renderer.render_device(frames[i].data_ptr(), w, h, camera=cams[k], stream=stream)

    for _ in range(W):
        step()
    sync_all()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    with ClockSampler(local_rank) as clocks:
        sync_all()
        for i in range(K):
            flush.fill_(i & 0xFF)
            ev[i][0].record()
            step()
            ev[i][1].record()
        sync_all()
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))

    host = torch.empty((h, w), dtype=torch.int32).pin_memory()
    for k in mine[:2]:
        renderer.render_host(host.data_ptr(), w, h, camera=cams[k])
    sync_all()
    t0 = _time.perf_counter()
    for _ in range(max(1, K // 4)):
        for k in mine:
            renderer.render_host(host.data_ptr(), w, h, camera=cams[k])
    sync_all()
    e2e_ms = (_time.perf_counter() - t0) / max(1, K // 4) * 1e3

    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        ms_per_step = total_ms / K
        rays = F * w * h
        print(json.dumps({
            "metric": METRIC, "value": rays / (ms_per_step * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_per_step, "ms_per_frame": ms_per_step / F,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{F}-frame camera orbit of {args.scene}.lol at {w}x{h} (BASELINE config C5), "
                                   f"whole frames dealt to ranks, frames left in each rank's HBM",
                       "arith": args.arith, "variant": args.variant,
                       "l2": "256 MB write between timed steps (untimed)", "kernel": renderer.kernel_info()},
            "e2e": {"value": rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_frame": e2e_ms / F,
                    "h2d_bytes_per_step": 192 * F, "d2h_bytes_per_step": rays * 4},
            "gpu_launches": K * len(mine),
            "clocks": clocks.summary(),
        }))
    if world > 1:
        dist.destroy_process_group()
    return 0

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

    scene = load_scene(lb, args.scene)
    opt = lb.Options.default(variant=args.variant, arith=1 if args.arith == "fast" else 0)
    renderer = lb.Renderer(scene, opt, device=local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    K, W = args.steps, args.warmup
    if args.workload == "orbit":
        return run_orbit(args, lb, torch, dist, scene, renderer, world, rank, local_rank, dev, w, h)

    shard_px = lb.shard_pixels(w, h, world)
    frame = torch.zeros((h, w), dtype=torch.int32, device=dev) if rank == 0 else None
    if world > 1:
        shard = lb.Shard(rank=rank, world=world, band_rows=0, dst_full_frame=0)
        if rank == 0:
            gathered = torch.zeros((world, shard_px), dtype=torch.int32, device=dev)
            local = gathered[0]
            gather_list = [gathered[i] for i in range(world)]
        else:
            local = torch.zeros((shard_px,), dtype=torch.int32, device=dev)
            gather_list = None
        peer_frame_ptr = None
        if args.gather == "peer":
            handle = torch.zeros(64, dtype=torch.uint8)
            if rank == 0:
                import ctypes as C
                hb = (C.c_uint8 * 64)()
                rc = lb.lib().lolb200_ipc_export(frame.data_ptr(), C.byref(hb))
                assert rc == 0, lb.lib().lolb200_last_error()
                handle = torch.tensor(list(hb), dtype=torch.uint8)
            hdev = handle.to(dev)
            dist.broadcast(hdev, 0)
            if rank == 0:
                peer_frame_ptr = frame.data_ptr()
            else:
                import ctypes as C
                hb = (C.c_uint8 * 64)(*hdev.cpu().tolist())
                p = C.c_void_p()
                rc = lb.lib().lolb200_ipc_open(C.byref(hb), C.byref(p))
                assert rc == 0, lb.lib().lolb200_last_error()
                peer_frame_ptr = p.value
            shard = lb.Shard(rank=rank, world=world, band_rows=0, dst_full_frame=1)
            token = torch.zeros(1, dtype=torch.int32, device=dev)

    launches = [0]

    def step():
        """One frame, complete on rank 0's HBM when the stream drains."""
        if world == 1:
            renderer.render_device(frame.data_ptr(), w, h, stream=stream)
            launches[0] += 1
        elif args.gather == "peer":
            renderer.render_device(peer_frame_ptr, w, h, shard=shard, pitch_px=w, stream=stream)
            launches[0] += 1
            dist.all_reduce(token)  # every rank's stores are done before rank 0 goes on
        else:
            renderer.render_device(local.data_ptr(), w, h, shard=shard, pitch_px=w, stream=stream)
            launches[0] += 1
            dist.gather(local, gather_list, dst=0)
            if rank == 0:
                lb.deinterleave(gathered.data_ptr(), frame.data_ptr(), w, h, world, shard_px,
                                stream=stream)
                launches[0] += 1

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(W, 3)):
        step()
    sync_all()

    # ---- timed region: K steps, device time, L2 flushed between steps (untimed) ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches[0] = 0
    with ClockSampler(local_rank) as clocks:
        sync_all()
        t_wall0 = time.perf_counter()
        for i in range(K):
            flush.fill_(i & 0xFF)
            ev[i][0].record()
            step()
            ev[i][1].record()
        sync_all()
        t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    n_launches = launches[0]

    # kernel-only duration (the render kernel alone), for the roofline
    for i in range(K):
        flush.fill_(i & 0xFF)
        kev[i][0].record()
        if world == 1:
            renderer.render_device(frame.data_ptr(), w, h, stream=stream)
        else:
            renderer.render_device(local.data_ptr() if args.gather != "peer" else peer_frame_ptr,
                                   w, h, shard=shard, pitch_px=w, stream=stream)
        kev[i][1].record()
    sync_all()
    kernel_ms = float(sum(a.elapsed_time(b) for a, b in kev)) / K

    if world > 1:
        t = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, kernel_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / K
    value = w * h / (ms_per_step * 1e-3) / 1e6

    # ---- the sharded frame must be the single-GPU frame, bit for bit ----
    verified = None
    if world > 1 and rank == 0:
        check = torch.zeros((h, w), dtype=torch.int32, device=dev)
        renderer.render_device(check.data_ptr(), w, h, stream=stream)
        torch.cuda.synchronize()
        verified = bool(torch.equal(check, frame))
        del check

    # ---- e2e: host buffers, D2H inside the timed region ----
    # N = 1: lolb200_render_host, the call b200_renderer.c makes.  N > 1: the frame has to end
    # up in HOST memory, so by default no GPU gathers anything: every rank renders its bands and
    # its own copy engine writes them into one POSIX shared-memory frame (rank 0 created it, all
    # ranks map and pin it): 1/N of the bytes per PCIe link instead of all of them through rank
    # 0's (lolb200_render_host_shard; the in-process twin is `--gather host` of the C backend).
    host = torch.empty((h, w), dtype=torch.int32).pin_memory() if rank == 0 else None
    shared, shared_path, e2e_verified = None, None, None
    if world > 1 and args.e2e_path == "host-shards":
        import numpy as np
        shared_path = f"/dev/shm/lolb200_bench_{os.environ.get('MASTER_PORT', '0')}_{w}x{h}"
        if rank == 0:
            np.memmap(shared_path, dtype=np.uint32, mode="w+", shape=(h, w)).flush()
        dist.barrier()
        shared = np.memmap(shared_path, dtype=np.uint32, mode="r+", shape=(h, w))

    def e2e_step():
        if world == 1:
            renderer.render_host(host.data_ptr(), w, h)
        elif shared is not None:
            renderer.render_host_shard(shared.ctypes.data, w, h, shard)  # returns when this rank's rows are in
            dist.barrier()                                               # the frame is complete for everyone
        else:
            step()
            if rank == 0:
                host.copy_(frame, non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(3):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    sync_all()
    e2e_ms = (time.perf_counter() - t0) / K * 1e3
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
    e2e_value = w * h / (e2e_ms * 1e-3) / 1e6
    if shared is not None:
        if rank == 0:
            import numpy as np
            e2e_verified = bool(np.array_equal(np.asarray(shared), frame.cpu().numpy().view(np.uint32)))
        dist.barrier()
        del shared
        if rank == 0:
            os.unlink(shared_path)

    # ---- executed work (instrumented twin of the kernel, untimed) ----
    copt = lb.Options.default(variant=args.variant, arith=1 if args.arith == "fast" else 0, counters=1)
    crend = lb.Renderer(scene, copt, device=local_rank)
    scratch = torch.zeros((h, w) if world == 1 else (shard_px,), dtype=torch.int32, device=dev)
    crend.render_device(scratch.data_ptr(), w, h, pitch_px=w,
                        shard=None if world == 1 else lb.Shard(rank=rank, world=world), stream=stream)
    torch.cuda.synchronize()
    cnt = crend.read_counters()
    f_sdf = scene.flops_per_eval()
    n_lights = scene.struct.n_lights
    exec_flops = flops_model(f_sdf, n_lights, cnt["pixels"], cnt["primary_evals"], cnt["normal_evals"],
                             cnt["shadow_evals"], cnt["normal_evals"] // 4, cnt["shadow_rays"],
                             cnt["shadow_rays_culled"])
    # objects that a box test skipped inside an evaluation were not executed
    exec_flops -= cnt.get("skipped_flops", 0)
    if world > 1:
        t = torch.tensor([exec_flops], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the slowest rank bounds the frame
        exec_flops = float(t[0])

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
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
        "ms_per_step": ms_per_step, "ms_per_frame": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.scene}.lol at {w}x{h}, one primary ray per pixel, frame complete "
                               f"in rank 0's HBM", "arith": args.arith, "variant": args.variant,
                   "sharding": "single GPU" if world == 1 else f"4-row bands cyclic over {world} ranks, "
                               f"gather={args.gather}",
                   "l2": "256 MB write between timed steps (untimed); the kernel reads no global inputs",
                   "kernel": renderer.kernel_info()},
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_frame": e2e_ms,
                "h2d_bytes_per_step": 192 * world, "d2h_bytes_per_step": w * h * 4,
                "path": ("lolb200_render_host: slab launches overlapped with the read-back into a pinned host frame"
                         if world == 1 else
                         "lolb200_render_host_shard on every rank: own bands over own PCIe link into one "
                         "shared-memory host frame, then a barrier" if args.e2e_path == "host-shards" else
                         "frame gathered on rank 0 over NVLink, then one D2H copy from rank 0"),
                "host_frame_equals_single_gpu": e2e_verified},
        "gpu_launches": n_launches,
        "sharded_frame_equals_single_gpu": verified,
        "roofline": roofline,
        "clocks": clocks.summary(),
        "wall_ms_per_step_incl_flush": t_wall / K * 1e3,
        "executed_evals_per_pixel": {k: cnt[k] / max(1, cnt["pixels"]) for k in
                                     ("primary_evals", "normal_evals", "shadow_evals")},
    }

    if world == 1 and not args.no_cpu_baseline:
        # ~20 core-seconds on scene4: every 4th scanline of the same frame; the
        # 1024-primitive scene costs ~1000x more per ray, so only a few scanlines
        stride = 4 if not args.scene.startswith("synthetic") else max(4, h // 3)
        ms, rays, kind, cores, totals = cpu_sample(args.scene, w, h, stride)
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
        roofline["achieved_reference_work"] = ref_flops / (kernel_ms * 1e-3) / 1e12
        roofline["flop_per_launch_reference"] = ref_flops
        # the reference's own work for this frame per second of GPU time, against the same peak: what
        # the exact skips and box tests buy on top of the hardware rate `frac`
        roofline["frac_reference_work"] = roofline["achieved_reference_work"] / peak_tf
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
            out["cpu_jit_equivalent"] = cpu_jit_equivalent(args.scene, w, h, stride)
        except Exception as e:  # a missing host compiler must not cost the GPU numbers
            out["cpu_jit_equivalent"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    if args.all_scenes and world == 1:
        per = {}
        for name in ("scene", "scene2", "scene3", "scene4"):
            r2 = lb.Renderer(load_scene(lb, name), opt, device=local_rank)
            for _ in range(3):
                r2.render_device(frame.data_ptr(), w, h, stream=stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                r2.render_device(frame.data_ptr(), w, h, stream=stream)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            per[name] = {"ms_per_frame": ms, "mrays_s": w * h / ms / 1e3}
            r2.close()
        out["per_scene"] = per

    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
