"""The C-ABI library loads and exports every symbol include/lolb200.h declares;
device entry points fail loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lolb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lolb200_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    import loltracer_b200 as lb

    raw = C.CDLL(lb.library_path())
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in lolb200.h but not exported"
    assert lb.lib().lolb200_abi_version() == 3


def test_bindings_cover_the_header():
    import loltracer_b200 as lb

    L = lb.lib()
    for s in declared_symbols():
        assert getattr(L, s).argtypes is not None, f"no ctypes signature for {s}"


def test_struct_layouts_match_c(tmp_path):
    """sizeof of every POD and the offset of its last field, as gcc sees include/lolb200.h, against the
    ctypes mirrors in loltracer_b200/api.py."""
    import subprocess

    from loltracer_b200 import api

    pods = {"lolb200_material": (api.Material, "ambient"), "lolb200_light": (api.Light, "specular_intensity"),
            "lolb200_object": (api.Object, "b"), "lolb200_camera": (api.Camera, "fov"),
            "lolb200_scene": (api.SceneStruct, "camera"), "lolb200_camera_basis": (api.CameraBasis, "height"),
            "lolb200_options": (api.Options, "child_materials"), "lolb200_pixfmt": (api.PixFmt, "amask"),
            "lolb200_shard": (api.Shard, "done_value"), "lolb200_aux": (api.Aux, "launch_timing")}
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lolb200.h"\nint main(void) {\n' + "".join(
        f'  printf("{name} %zu %zu\\n", sizeof({name}), offsetof({name}, {last}));\n'
        for name, (_, last) in pods.items()) + "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    seen = {}
    for line in subprocess.check_output([str(exe)], text=True).splitlines():
        name, size, off = line.split()
        seen[name] = (int(size), int(off))
    for name, (cls, last) in pods.items():
        assert seen[name] == (C.sizeof(cls), getattr(cls, last).offset), name
    assert C.sizeof(api.Options) == 96 and C.sizeof(api.Shard) == 32


def test_no_cpu_fallback(scenes_dir):
    """Without a GPU the device layer must refuse, not emulate."""
    import torch
    import loltracer_b200 as lb

    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less box")
    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene.lol"))
    assert lb.device_count() == 0
    with pytest.raises(lb.LolB200Error) as e:
        lb.Renderer(scene)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    with pytest.raises(lb.LolB200Error):
        lb.measure_fp32_peak(0)


def test_product_does_not_import_the_oracle():
    """Nothing under loltracer_b200/ or include/ may mention the checker."""
    bad = []
    for base in ("loltracer_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in d or "__pycache__" in d:
                continue
            for f in files:
                if f.endswith((".so", ".o", ".pyc")):
                    continue
                text = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"oracle_lib|liblol_oracle|liblolref|lolo_render|lolref_", text):
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_ptx_and_sass_dumps_work_without_a_gpu(scenes_dir):
    """--dump-ptx / --dump-sass (the jitdump analogue): the program's PTX targets sm_100a
    and carries the kernel; the SASS listing of the compiled image comes from the toolkit's
    disassembler and shows the FP32 pipeline the kernel is made of."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene4.lol"))
    opt = lb.Options.default()
    src = lb.lower_cuda(scene, opt)
    ptx = lb.compile_ptx(src, opt)
    assert ".target sm_100a" in ptx and ".entry lol_render" in ptx
    assert "fma.rn.f32" in ptx and "mul.rn.f32" in ptx       # the guarded forms' explicit FMAs, plain products
    sass = lb.disassemble(lb.compile_cubin(src, opt))
    assert "lol_render" in sass and "FADD" in sass and "MUFU.RSQ" in sass
    assert "HMMA" not in sass and "UTCMMA" not in sass         # no tensor-core instruction: FP32 ALU work


def test_library_has_no_link_time_dependency_on_the_driver_or_nvtx():
    """libcuda (stream memory operations) and a profiler's NVTX injection library are bound at run
    time, NCCL through dlopen: the library loads on a box that has none of them."""
    import subprocess

    import loltracer_b200 as lb

    needed = subprocess.check_output(["readelf", "-d", lb.library_path()], text=True)
    libs = re.findall(r"NEEDED.*\[(.*?)\]", needed)
    assert not [x for x in libs if x.startswith(("libcuda.so", "libnccl", "libnvToolsExt"))], libs
