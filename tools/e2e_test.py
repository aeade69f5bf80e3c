"""Times lolb200_render_host (host surface in, pixels out) per scene:  python tools/e2e_test.py [WxH]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import loltracer_b200 as lb
w, h = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "3840x2160").split("x"))
for name in ("scene", "scene4"):
    scene = lb.Scene.from_file(os.path.join(ROOT, "tests", "golden", "scenes", name + ".lol"))
    r = lb.Renderer(scene)
    host = np.zeros((h, w), np.uint32)
    for _ in range(3):
        r.render_host(host.ctypes.data, w, h)
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        r.render_host(host.ctypes.data, w, h)
    ms = (time.perf_counter() - t0) / n * 1e3
    print(f"{name} {w}x{h} mode={os.environ.get('LOLB200_HOST_MODE','default')}: e2e {ms:.3f} ms/frame  checksum {int(host.sum()) & 0xffffffff:08x}")
    r.close()
