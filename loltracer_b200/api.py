"""ctypes bindings for include/lolb200.h.

Names follow the C ABI one to one; the docstrings cite the reference interface
each call replaces (paths relative to the reference tree).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))


def library_path() -> str:
    return os.path.join(_HERE, "liblolb200.so")


class LolB200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"lolb200 error {code}: {message}")
        self.code = code


# ----------------------------------------------------------------- PODs --


class Material(C.Structure):  # scene.h:44-49
    _fields_ = [("shininess", C.c_float), ("diffuse", C.c_float * 3),
                ("specular", C.c_float * 3), ("ambient", C.c_float * 3)]


class Light(C.Structure):  # scene.h:52-56
    _fields_ = [("point", C.c_float * 3), ("diffuse_intensity", C.c_float * 3),
                ("specular_intensity", C.c_float * 3)]


class Object(C.Structure):  # scene.h:58-82
    _fields_ = [("type", C.c_int32), ("material", C.c_uint32), ("point", C.c_float * 3),
                ("radius", C.c_float), ("point2", C.c_float * 3), ("smoothness", C.c_float),
                ("a", C.c_int32), ("b", C.c_int32)]


class Camera(C.Structure):  # scene.h:84-88
    _fields_ = [("point", C.c_float * 3), ("direction", C.c_float * 3), ("fov", C.c_float)]

    @classmethod
    def make(cls, point: Sequence[float], direction: Sequence[float], fov: float) -> "Camera":
        c = cls()
        c.point[:] = list(point)
        c.direction[:] = list(direction)
        c.fov = fov
        return c


class SceneStruct(C.Structure):  # scene.h:90-96, flattened
    _fields_ = [("n_materials", C.c_uint32), ("materials", C.POINTER(Material)),
                ("ambient_color", C.c_float * 3),
                ("n_lights", C.c_uint32), ("lights", C.POINTER(Light)),
                ("n_nodes", C.c_uint32), ("nodes", C.POINTER(Object)),
                ("n_objects", C.c_uint32), ("objects", C.POINTER(C.c_uint32)),
                ("camera", Camera)]


class CameraBasis(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("dir", C.c_float * 3), ("right", C.c_float * 3),
                ("up", C.c_float * 3), ("width", C.c_float), ("height", C.c_float)]


class Options(C.Structure):
    _fields_ = [("arith", C.c_int32), ("skip_black_miss", C.c_int32),
                ("cull_backfacing", C.c_int32), ("shadow_early_out", C.c_int32),
                ("counters", C.c_int32), ("variant", C.c_int32), ("loop_threshold", C.c_int32),
                ("guarded_fastpath", C.c_int32), ("prune_bounds", C.c_int32),
                ("block_threads", C.c_int32), ("min_blocks", C.c_int32), ("roll_phases", C.c_int32),
                ("prune_group", C.c_int32), ("pack_pairs", C.c_int32), ("share_first_step", C.c_int32),
                ("shadow_div_pretest", C.c_int32), ("defer_cap_primary", C.c_int32),
                ("defer_cap_shadow", C.c_int32), ("roll_v1", C.c_int32), ("loop_worklist", C.c_int32),
                ("near_cache", C.c_int32), ("guard_out", C.c_int32), ("grid_cells", C.c_int32), ("child_materials", C.c_int32)]

    @classmethod
    def default(cls, **kw) -> "Options":
        o = cls()
        lib().lolb200_options_default(C.byref(o))
        for k, v in kw.items():
            if not hasattr(o, k):
                raise AttributeError(k)
            setattr(o, k, int(v))
        return o


class PixFmt(C.Structure):
    _fields_ = [("rshift", C.c_uint8), ("gshift", C.c_uint8), ("bshift", C.c_uint8),
                ("rloss", C.c_uint8), ("gloss", C.c_uint8), ("bloss", C.c_uint8),
                ("pad", C.c_uint16), ("amask", C.c_uint32)]

    @classmethod
    def default(cls) -> "PixFmt":
        f = cls()
        lib().lolb200_pixfmt_default(C.byref(f))
        return f


class Shard(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("band_rows", C.c_int32),
                ("dst_full_frame", C.c_int32), ("done_flag", C.c_void_p), ("done_value", C.c_uint32),
                ("reserved", C.c_uint32)]


class Aux(C.Structure):
    _fields_ = [("dist", C.c_void_p), ("id", C.c_void_p), ("primary_steps", C.c_void_p),
                ("shadow_steps", C.c_void_p), ("launch_timing", C.c_void_p)]


# -------------------------------------------------------------- library --

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Loads liblolb200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise LolB200Error(-3, f"{path} is missing: run `python -c 'import __graft_entry__ as g; "
                               "g.build()'` or `make -C loltracer_b200/csrc`")
    L = C.CDLL(path)
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    sig = {
        "lolb200_scene_parse_file": (i32, [C.c_char_p, C.POINTER(C.POINTER(SceneStruct))]),
        "lolb200_scene_parse_string": (i32, [C.c_char_p, sz, C.POINTER(C.POINTER(SceneStruct))]),
        "lolb200_scene_clone": (C.POINTER(SceneStruct), [C.POINTER(SceneStruct)]),
        "lolb200_scene_free": (None, [C.POINTER(SceneStruct)]),
        "lolb200_render_host_shard": (i32, [vp, C.POINTER(Camera), i32, i32, C.POINTER(PixFmt), C.POINTER(Shard),
                                            vp, sz]),
        "lolb200_camera_basis_compute": (None, [C.POINTER(Camera), i32, i32, C.POINTER(CameraBasis)]),
        "lolb200_options_default": (None, [C.POINTER(Options)]),
        "lolb200_lower_cuda": (vp, [C.POINTER(SceneStruct), C.POINTER(Options), C.POINTER(sz)]),
        "lolb200_free": (None, [vp]),
        "lolb200_scene_flops_per_eval": (C.c_uint64, [C.POINTER(SceneStruct)]),
        "lolb200_compile_cubin": (i32, [C.c_char_p, C.POINTER(Options), C.POINTER(vp),
                                         C.POINTER(sz), C.POINTER(vp)]),
        "lolb200_compile_ptx": (i32, [C.c_char_p, C.POINTER(Options), C.POINTER(vp), C.POINTER(sz)]),
        "lolb200_disassemble": (i32, [C.c_char_p, sz, C.POINTER(vp), C.POINTER(sz)]),
        "lolb200_device_count": (i32, []),
        "lolb200_renderer_create": (i32, [C.POINTER(SceneStruct), C.POINTER(Options), i32,
                                           C.POINTER(vp)]),
        "lolb200_renderer_destroy": (None, [vp]),
        "lolb200_renderer_source": (C.c_char_p, [vp]),
        "lolb200_renderer_image": (vp, [vp, C.POINTER(sz)]),
        "lolb200_renderer_kernel_info": (i32, [vp] + [C.POINTER(i32)] * 4),
        "lolb200_pixfmt_default": (None, [C.POINTER(PixFmt)]),
        "lolb200_render_device": (i32, [vp, C.POINTER(Camera), i32, i32, C.POINTER(PixFmt),
                                         C.POINTER(Shard), vp, sz, C.POINTER(Aux), vp]),
        "lolb200_render_host": (i32, [vp, C.POINTER(Camera), i32, i32, C.POINTER(PixFmt), vp, sz]),
        "lolb200_read_counters": (i32, [vp, C.POINTER(C.c_uint64 * 8)]),
        "lolb200_deinterleave_device": (i32, [vp, vp, i32, i32, i32, i32, sz, sz, vp]),
        "lolb200_shard_pixels": (sz, [i32, i32, i32, i32]),
        "lolb200_group_create": (i32, [C.POINTER(SceneStruct), C.POINTER(Options), C.POINTER(i32), i32, i32,
                                        C.POINTER(vp)]),
        "lolb200_group_destroy": (None, [vp]),
        "lolb200_group_render_host": (i32, [vp, C.POINTER(Camera), i32, i32, C.POINTER(PixFmt), vp, sz]),
        "lolb200_group_last_frame_ms": (C.c_double, [vp]),
        "lolb200_group_size": (i32, [vp]),
        "lolb200_group_share_enqueue": (i32, [vp, i32, C.POINTER(Camera), i32, i32, C.POINTER(PixFmt), vp, sz]),
        "lolb200_group_share_wait": (i32, [vp, i32]),
        "lolb200_surface_pin": (i32, [vp, sz]),
        "lolb200_surface_unpin": (i32, [vp]),
        "lolb200_stream_wait_value32": (i32, [vp, vp, C.c_uint32]),
        "lolb200_stream_write_value32": (i32, [vp, vp, C.c_uint32]),
        "lolb200_ipc_export": (i32, [vp, C.POINTER(C.c_uint8 * 64)]),
        "lolb200_ipc_open": (i32, [C.POINTER(C.c_uint8 * 64), C.POINTER(vp)]),
        "lolb200_ipc_close": (i32, [vp]),
        "lolb200_measure_fp32_peak": (C.c_double, [i32, i32, C.POINTER(C.c_double)]),
        "lolb200_last_error": (C.c_char_p, []),
        "lolb200_abi_version": (i32, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


ABI_SYMBOLS = None  # filled lazily by tests from include/lolb200.h


def _check(rc: int) -> None:
    if rc != 0:
        raise LolB200Error(rc, lib().lolb200_last_error().decode(errors="replace"))


# ---------------------------------------------------------------- scene --


class Scene:
    """A parsed .lol scene (replaces scene_parse(), scene-parser.y:197-214)."""

    def __init__(self, ptr):
        self._ptr = ptr

    @classmethod
    def from_file(cls, path: str) -> "Scene":
        p = C.POINTER(SceneStruct)()
        _check(lib().lolb200_scene_parse_file(os.fsencode(path), C.byref(p)))
        return cls(p)

    @classmethod
    def from_string(cls, text: str) -> "Scene":
        raw = text.encode()
        p = C.POINTER(SceneStruct)()
        _check(lib().lolb200_scene_parse_string(raw, len(raw), C.byref(p)))
        return cls(p)

    @property
    def struct(self) -> SceneStruct:
        return self._ptr.contents

    @property
    def camera(self) -> Camera:
        c = Camera()
        C.memmove(C.byref(c), C.byref(self._ptr.contents.camera), C.sizeof(Camera))
        return c

    def flops_per_eval(self) -> int:
        return int(lib().lolb200_scene_flops_per_eval(self._ptr))

    def __del__(self):
        if getattr(self, "_ptr", None) and _lib is not None:
            _lib.lolb200_scene_free(self._ptr)
            self._ptr = None


def camera_basis(cam: Camera, w: int, h: int) -> CameraBasis:
    """The per-frame half of get_camera_ray() (naive_renderer.c:178-188)."""
    out = CameraBasis()
    lib().lolb200_camera_basis_compute(C.byref(cam), w, h, C.byref(out))
    return out


def lower_cuda(scene: Scene, options: Optional[Options] = None) -> str:
    """Scene -> specialised CUDA C (replaces generate_sdf(), tracing_jit_renderer.dasc:76-143)."""
    n = C.c_size_t()
    p = lib().lolb200_lower_cuda(scene._ptr, C.byref(options) if options else None, C.byref(n))
    if not p:
        raise LolB200Error(-2, lib().lolb200_last_error().decode(errors="replace"))
    try:
        return C.string_at(p, n.value).decode()
    finally:
        lib().lolb200_free(p)


def compile_cubin(src: str, options: Optional[Options] = None) -> bytes:
    """NVRTC for sm_100a; needs no GPU (replaces link_and_encode(), tracing_jit_renderer.dasc:60-74)."""
    img, n, log = C.c_void_p(), C.c_size_t(), C.c_void_p()
    rc = lib().lolb200_compile_cubin(src.encode(), C.byref(options) if options else None,
                                     C.byref(img), C.byref(n), C.byref(log))
    if log.value:
        lib().lolb200_free(log)
    _check(rc)
    try:
        return C.string_at(img, n.value)
    finally:
        lib().lolb200_free(img)


def compile_ptx(src: str, options: Optional[Options] = None) -> str:
    """The program's PTX (compute_100a): --dump-ptx of the backend."""
    out, n = C.c_void_p(), C.c_size_t()
    _check(lib().lolb200_compile_ptx(src.encode(), C.byref(options) if options else None, C.byref(out), C.byref(n)))
    try:
        return C.string_at(out, n.value).decode()
    finally:
        lib().lolb200_free(out)


def disassemble(image: bytes) -> str:
    """SASS listing of a compiled image (cuobjdump / nvdisasm): --dump-sass of the backend."""
    out, n = C.c_void_p(), C.c_size_t()
    _check(lib().lolb200_disassemble(image, len(image), C.byref(out), C.byref(n)))
    try:
        return C.string_at(out, n.value).decode(errors="replace")
    finally:
        lib().lolb200_free(out)


def device_count() -> int:
    return int(lib().lolb200_device_count())


def shard_pixels(w: int, h: int, world: int, band_rows: int = 0) -> int:
    return int(lib().lolb200_shard_pixels(w, h, world, band_rows))


def deinterleave(gathered_ptr: int, frame_ptr: int, w: int, h: int, world: int,
                 shard_px: int, pitch_px: Optional[int] = None, stream: int = 0) -> None:
    _check(lib().lolb200_deinterleave_device(gathered_ptr, frame_ptr, w, h, world, 0, shard_px,
                                             pitch_px or w, stream))


def stream_wait_value32(stream: int, dev_addr: int, value: int) -> None:
    """`stream` waits until the word at dev_addr is >= value (cyclic): a stream memory operation."""
    _check(lib().lolb200_stream_wait_value32(stream, dev_addr, value & 0xFFFFFFFF))


def stream_write_value32(stream: int, dev_addr: int, value: int) -> None:
    _check(lib().lolb200_stream_write_value32(stream, dev_addr, value & 0xFFFFFFFF))


def measure_fp32_peak(device: int = 0, iters: int = 0) -> tuple[float, float]:
    ms = C.c_double()
    tf = lib().lolb200_measure_fp32_peak(device, iters, C.byref(ms))
    if tf < 0:
        raise LolB200Error(-5, lib().lolb200_last_error().decode(errors="replace"))
    return float(tf), float(ms.value)


# ------------------------------------------------------------- renderer --


class Renderer:
    """render_prepare / render_thread's pixel loop / render_destroy (renderer.h:24-26)
    for one scene on one GPU."""

    def __init__(self, scene: Scene, options: Optional[Options] = None, device: int = 0):
        self._h = C.c_void_p()
        self.scene = scene
        self.device = device
        _check(lib().lolb200_renderer_create(scene._ptr, C.byref(options) if options else None,
                                             device, C.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value and _lib is not None:
            _lib.lolb200_renderer_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    @property
    def source(self) -> str:
        return lib().lolb200_renderer_source(self._h).decode()

    @property
    def image(self) -> bytes:
        n = C.c_size_t()
        p = lib().lolb200_renderer_image(self._h, C.byref(n))
        return C.string_at(p, n.value)

    def kernel_info(self) -> dict:
        v = [C.c_int() for _ in range(4)]
        _check(lib().lolb200_renderer_kernel_info(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("regs", "smem_bytes", "local_bytes", "max_threads"), (x.value for x in v)))

    def render_device(self, dst_ptr: int, w: int, h: int, camera: Optional[Camera] = None,
                      pitch_px: Optional[int] = None, shard: Optional[Shard] = None,
                      fmt: Optional[PixFmt] = None, aux: Optional[Aux] = None,
                      stream: int = 0) -> None:
        """Asynchronous: one frame (or one shard of it) into device memory."""
        _check(lib().lolb200_render_device(
            self._h, C.byref(camera) if camera else None, w, h, C.byref(fmt) if fmt else None,
            C.byref(shard) if shard else None, dst_ptr, pitch_px or w,
            C.byref(aux) if aux else None, stream))

    def render_host(self, pixels_ptr: int, w: int, h: int, camera: Optional[Camera] = None,
                    pitch_bytes: Optional[int] = None, fmt: Optional[PixFmt] = None) -> None:
        """Synchronous, end to end into a host surface (what the renderer.h leader calls)."""
        _check(lib().lolb200_render_host(self._h, C.byref(camera) if camera else None, w, h,
                                         C.byref(fmt) if fmt else None, pixels_ptr,
                                         pitch_bytes or w * 4))

    def render_host_shard(self, pixels_ptr: int, w: int, h: int, shard: Shard,
                          camera: Optional[Camera] = None, pitch_bytes: Optional[int] = None,
                          fmt: Optional[PixFmt] = None) -> None:
        """This rank's bands, end to end into the full-frame host surface (own PCIe link)."""
        _check(lib().lolb200_render_host_shard(self._h, C.byref(camera) if camera else None, w, h,
                                               C.byref(fmt) if fmt else None, C.byref(shard), pixels_ptr,
                                               pitch_bytes or w * 4))

    def read_counters(self) -> dict:
        out = (C.c_uint64 * 8)()
        _check(lib().lolb200_read_counters(self._h, C.byref(out)))
        names = ("primary_evals", "normal_evals", "shadow_evals", "pixels", "hit_pixels",
                 "shadow_rays", "shadow_rays_culled", "skipped_flops")
        return dict(zip(names, (int(x) for x in out)))


class Group:
    """Several GPUs driven by one process (what b200_renderer.c --gpus N uses)."""

    GATHER = {"nccl": 0, "peer": 1, "host": 2}

    def __init__(self, scene: Scene, n_devices: int, gather: str = "nccl",
                 options: Optional[Options] = None, devices: Optional[Sequence[int]] = None):
        self._h = C.c_void_p()
        devs = (C.c_int * n_devices)(*(devices or range(n_devices)))
        _check(lib().lolb200_group_create(scene._ptr, C.byref(options) if options else None, devs,
                                          n_devices, self.GATHER[gather], C.byref(self._h)))

    def render_host(self, pixels_ptr: int, w: int, h: int, camera: Optional[Camera] = None,
                    pitch_bytes: Optional[int] = None, fmt: Optional[PixFmt] = None) -> float:
        _check(lib().lolb200_group_render_host(self._h, C.byref(camera) if camera else None, w, h,
                                               C.byref(fmt) if fmt else None, pixels_ptr,
                                               pitch_bytes or w * 4))
        return float(lib().lolb200_group_last_frame_ms(self._h))

    @property
    def size(self) -> int:
        return int(lib().lolb200_group_size(self._h))

    def share_enqueue(self, share: int, pixels_ptr: int, w: int, h: int, camera: Optional[Camera] = None,
                      pitch_bytes: Optional[int] = None, fmt: Optional[PixFmt] = None) -> None:
        """One device's bands of a frame (gather 'host'): asynchronous; any thread may drive a share."""
        _check(lib().lolb200_group_share_enqueue(self._h, share, C.byref(camera) if camera else None, w, h,
                                                 C.byref(fmt) if fmt else None, pixels_ptr, pitch_bytes or w * 4))

    def share_wait(self, share: int) -> None:
        _check(lib().lolb200_group_share_wait(self._h, share))

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value and _lib is not None:
            _lib.lolb200_group_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close


def surface_pin(pixels_ptr: int, nbytes: int) -> None:
    """Page-lock a surface its caller owns (frames are then DMA-ed straight into it)."""
    _check(lib().lolb200_surface_pin(pixels_ptr, nbytes))


def surface_unpin(pixels_ptr: int) -> None:
    _check(lib().lolb200_surface_unpin(pixels_ptr))
