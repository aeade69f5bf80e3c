"""Dumps the per-chunk costs the longest-first order is sorted by (LOLB200_LPT_DUMP) for several option sets and
compares them with the work the chunks really hold (per-pixel step counts from the aux probes):
    python tools/lpt_probe.py scene4 3840x2160 "guard_out=0" "guard_out=3" """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import loltracer_b200 as lb

name, size = sys.argv[1], sys.argv[2]
configs = sys.argv[3:] or [""]
w, h = (int(x) for x in size.split("x"))
scene = lb.Scene.from_file(os.path.join(ROOT, "tests", "golden", "scenes", name + ".lol"))
st = torch.cuda.current_stream().cuda_stream
frame = torch.zeros((h, w), dtype=torch.int32, device="cuda")
npri = torch.zeros((h, w), dtype=torch.int16, device="cuda")
nsh = torch.zeros((h, w), dtype=torch.int16, device="cuda")
costs = {}
for cfg in configs:
    kw = {k: int(v) for k, v in (kv.split("=") for kv in cfg.split(",") if kv)}
    path = f"/tmp/lpt_{cfg.replace('=', '_').replace(',', '_') or 'default'}.bin"
    os.environ["LOLB200_LPT_DUMP"] = path
    r = lb.Renderer(scene, lb.Options.default(**kw))
    for _ in range(20):
        r.render_device(frame.data_ptr(), w, h, stream=st)
    torch.cuda.synchronize()
    costs[cfg] = np.fromfile(path, np.uint32).astype(np.float64)
    r.close()
del os.environ["LOLB200_LPT_DUMP"]
r = lb.Renderer(scene)
r.render_device(frame.data_ptr(), w, h, stream=st, aux=lb.Aux(primary_steps=npri.data_ptr(), shadow_steps=nsh.data_ptr()))
torch.cuda.synchronize()
work = (npri.cpu().numpy().astype(np.int64) + nsh.cpu().numpy().astype(np.int64))
n = len(next(iter(costs.values())))
bands = (h + 3) // 4
cpb = n // bands
cw = -(-w // cpb)
cw = (cw + 7) // 8 * 8
print(f"{n} chunks, {bands} bands x {cpb} chunks of {cw}x4 pixels")
# a chunk's serial work: per 8x4 sub-tile the slowest lane's evaluations, summed over its sub-tiles
pad = np.zeros((bands * 4, cpb * cw), np.int64)
pad[:h, :w] = work[:, :cpb * cw] if cpb * cw <= w else np.pad(work, ((0, 0), (0, cpb * cw - w)))[:, :cpb * cw]
t = pad.reshape(bands, 4, cpb, cw // 8, 8).max(axis=(1, 4)).sum(axis=2).reshape(-1)
for cfg, c in costs.items():
    order = np.argsort(-c, kind="stable")
    last = order[-4736 * 2:]  # what is in flight when the queue runs dry
    print(f"[{cfg}] cost: mean {c.mean():.0f} max {c.max():.0f} clocks; corr(cost, slowest-lane evaluations) {np.corrcoef(c, t)[0, 1]:.3f}; "
          f"zeros {int((c == 0).sum())}; among the last {len(last)} of the order: max evaluations {t[last].max()}, "
          f"mean {t[last].mean():.1f} (all chunks: max {t.max()}, mean {t.mean():.1f})")
    worst = last[np.argsort(-t[last])[:5]]
    print("   heaviest chunks sorted to the end:", [(int(i), int(t[i]), int(c[i])) for i in worst])
keys = list(costs)
if len(keys) >= 2:
    a, b = costs[keys[0]], costs[keys[1]]
    print(f"cost ratio [{keys[1]}]/[{keys[0]}]: median {np.median(b / np.maximum(a, 1)):.3f}, "
          f"p1 {np.percentile(b / np.maximum(a, 1), 1):.3f}, p99 {np.percentile(b / np.maximum(a, 1), 99):.3f}")
