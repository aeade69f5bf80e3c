"""What the per-ray candidate memory of a pruned table loop does (lol_kernel.cuh: struct lol_near), counted on
the CPU: the generated program's per-pixel pipeline compiled for the host with -DLOL_NEAR_STATS.
    python tools/near_stats.py [synthetic|synthetic_csg] [WxH]"""
import ctypes as C, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import loltracer_b200 as lb
import oracle_lib as ol
from loltracer_b200 import scenegen
name = sys.argv[1] if len(sys.argv) > 1 else "synthetic"
w, h = (int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "160x90").split("x"))
scene = lb.Scene.from_string(scenegen.synthetic_scene_text(csg=name.endswith("csg")))
near = int(sys.argv[3]) if len(sys.argv) > 3 else -1
grid = int(sys.argv[4]) if len(sys.argv) > 4 else 0
src = lb.lower_cuda(scene, lb.Options.default(variant=1, near_cache=near, grid_cells=grid))
with tempfile.TemporaryDirectory() as tmp:
    cu = os.path.join(tmp, "p.cpp")
    open(cu, "w").write("#define LOL_NEAR_STATS 1\n" + ol.HOST_SHIM + src + ol.PIPELINE_WRAPPER)
    so = os.path.join(tmp, "p.so")
    subprocess.check_call(["g++", "-O2", "-msse4.2", "-mavx2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, cu])
    L = C.CDLL(so)
    L.lol_near_stats_ptr.restype = C.POINTER(C.c_ulonglong * 24)
    ol.cpu_pipeline_render(L, lb, scene, w, h)
    st = list(L.lol_near_stats_ptr().contents)
calls, collects, slow, rows = st[:4]
print(f"{name} {w}x{h}: {calls} sdf calls, rows looked at again in {collects / calls:.1%} of them, the long way in "
      f"{slow / calls:.1%}, {rows / calls:.2f} rows evaluated per call (+ the long way's)")
print("rows that could not be skipped per look:", " ".join(f"{n}:{v / max(collects, 1):.1%}" for n, v in enumerate(st[4:20])))
print(f"looks answered by the candidate grid: {st[20] / max(collects, 1):.1%}; point outside the grid: {st[21] / max(collects, 1):.1%}; "
      f"the cell's list too short: {st[22] / max(collects, 1):.1%}")
