"""loltracer_b200 -- B200 (sm_100a) backend for loltracer's per-pixel sphere-tracing path.

The product is the C library ``liblolb200.so`` (C front-end + code generator +
thin CUDA layer, ``include/lolb200.h``) and the ``renderer.h`` drop-in built on it
(``loltracer_b200/backend/b200_renderer.c``).  This package is the ctypes view of
that C ABI used by the tests and ``bench.py``; it adds no compute of its own and
has no CPU fallback: without the built library, or without a GPU for the device
calls, it raises.
"""
from .api import (  # noqa: F401
    Aux,
    Camera,
    Group,
    LolB200Error,
    Options,
    PixFmt,
    Renderer,
    Scene,
    Shard,
    camera_basis,
    compile_cubin,
    compile_ptx,
    disassemble,
    deinterleave,
    device_count,
    lib,
    library_path,
    lower_cuda,
    measure_fp32_peak,
    shard_pixels,
    stream_wait_value32,
    stream_write_value32,
    surface_pin,
    surface_unpin,
)
