"""Lower + NVRTC-compile a scene here (no GPU needed) and report registers, spills and the SASS loops:
    python tools/cubin_info.py scene4 [k=v,k=v] [min_loop_len]
Writes /tmp/lol_<scene>.cu and /tmp/lol_<scene>.cubin."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import loltracer_b200 as lb
from loltracer_b200 import scenegen

name = sys.argv[1]
kw = dict(kv.split("=") for kv in (sys.argv[2] if len(sys.argv) > 2 else "").split(",") if kv)
scene = (lb.Scene.from_string(scenegen.synthetic_scene_text(csg=name.endswith("csg"))) if name.startswith("synthetic")
         else lb.Scene.from_file(os.path.join(ROOT, "tests", "golden", "scenes", name + ".lol")))
opt = lb.Options.default(**{k: int(v) for k, v in kw.items()})
src = lb.lower_cuda(scene, opt)
img = lb.compile_cubin(src, opt)
open(f"/tmp/lol_{name}.cu", "w").write(src)
open(f"/tmp/lol_{name}.cubin", "wb").write(img)
res = subprocess.run(["cuobjdump", "-res-usage", f"/tmp/lol_{name}.cubin"], capture_output=True, text=True).stdout
for line in res.splitlines():
    if "REG" in line or "Function" in line:
        print(line.strip())
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_loops.py"), f"/tmp/lol_{name}.cubin",
                sys.argv[3] if len(sys.argv) > 3 else "40"])
