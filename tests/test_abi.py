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
    assert lb.lib().lolb200_abi_version() == 1


def test_bindings_cover_the_header():
    import loltracer_b200 as lb

    L = lb.lib()
    for s in declared_symbols():
        assert getattr(L, s).argtypes is not None, f"no ctypes signature for {s}"


def test_struct_layouts_match_c():
    """sizeof of the PODs as the C compiler sees them (checked through the parser)."""
    from loltracer_b200 import api

    assert C.sizeof(api.Material) == 40
    assert C.sizeof(api.Light) == 36
    assert C.sizeof(api.Object) == 48
    assert C.sizeof(api.Camera) == 28
    assert C.sizeof(api.Options) == 64
    assert C.sizeof(api.PixFmt) == 12
    assert C.sizeof(api.Shard) == 16


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
