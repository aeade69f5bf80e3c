"""ctypes views of the CPU checkers (TEST INFRASTRUCTURE).

  liblol_oracle.so        oracle/lol_oracle.c, our restatement ("port")
  _ref/liblolref.so       the reference's own naive_renderer.c ("reference");
                          present when it was built in the container that has
                          /root/reference and travelled with the snapshot.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_PATH = os.path.join(ROOT, "oracle", "liblol_oracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "liblolref.so")
REF_O0_PATH = os.path.join(ROOT, "oracle", "_ref", "liblolref_O0.so")  # as the reference Makefile builds it: no -O

_port = None
_ref = None


def port():
    global _port
    if _port is None:
        L = C.CDLL(PORT_PATH)
        L.lolo_render.restype = C.c_double
        L.lolo_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_int] * 6 + [C.c_void_p] * 6
        L.lolo_sdf.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float * 3),
                               C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
        L.lolo_set_specialised_sdf.argtypes = [C.c_void_p]
        _port = L
    return _port


def have_ref() -> bool:
    return os.path.exists(REF_PATH)


def ref(path=None):
    """The compiled reference; path=REF_O0_PATH: the same sources without optimisation."""
    global _ref
    if path is not None:
        return _bind_ref(C.CDLL(path))
    if _ref is None:
        _ref = _bind_ref(C.CDLL(REF_PATH))
    return _ref


def _bind_ref(L):
    if True:
        L.lolref_scene_load.restype = C.c_void_p
        L.lolref_scene_load.argtypes = [C.c_char_p]
        L.lolref_scene_load_string.restype = C.c_void_p
        L.lolref_scene_load_string.argtypes = [C.c_char_p, C.c_size_t]
        L.lolref_scene_free.argtypes = [C.c_void_p]
        L.lolref_scene_set_camera.argtypes = [C.c_void_p, C.POINTER(C.c_float * 3),
                                              C.POINTER(C.c_float * 3)]
        L.lolref_scene_flatten.restype = C.c_void_p
        L.lolref_scene_flatten.argtypes = [C.c_void_p]
        L.lolref_render_protocol.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p]
        L.lolref_probe.restype = C.c_double
        L.lolref_probe.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p] * 3
        L.lolref_sdf.argtypes = [C.c_void_p, C.POINTER(C.c_float * 3), C.POINTER(C.c_float),
                                 C.POINTER(C.c_uint32)]
    return L


def nthreads() -> int:
    return max(1, len(os.sched_getaffinity(0)))


def _rows(h, y0, y1, ystride):
    y1 = h if y1 is None else y1
    return y1, (y1 - y0 + ystride - 1) // ystride


def port_render(scene, w, h, camera=None, mode=0, y0=0, y1=None, ystride=1, threads=None,
                counts=False):
    """scene: loltracer_b200.Scene.  Returns dict(dist, id, rgba[, nprimary, nshadow], ms, totals)."""
    y1, rows = _rows(h, y0, y1, ystride)
    dist = np.zeros((rows, w), np.float32)
    ids = np.zeros((rows, w), np.uint32)
    rgba = np.zeros((rows, w), np.uint32)
    npr = np.zeros((rows, w), np.uint16) if counts else None
    nsh = np.zeros((rows, w), np.uint16) if counts else None
    tot = (C.c_uint64 * 4)()
    ms = port().lolo_render(
        C.cast(scene._ptr, C.c_void_p), C.cast(C.byref(camera), C.c_void_p) if camera else None,
        mode, w, h, y0, y1, ystride, threads or nthreads(), dist.ctypes.data, ids.ctypes.data,
        rgba.ctypes.data, npr.ctypes.data if counts else None, nsh.ctypes.data if counts else None,
        C.cast(tot, C.c_void_p))
    out = dict(dist=dist, id=ids, rgba=rgba, ms=ms,
               totals=dict(primary=int(tot[0]), normal=int(tot[1]), shadow=int(tot[2]), hits=int(tot[3])))
    if counts:
        out.update(nprimary=npr, nshadow=nsh)
    return out


class RefScene:
    """A scene held by the compiled reference (struct scene*, built by its own scene.c)."""

    def __init__(self, path=None, text=None, lib=None):
        self.lib = lib or ref()
        if path is not None:
            self.ptr = self.lib.lolref_scene_load(os.fsencode(path))
        else:
            raw = text.encode()
            self.ptr = self.lib.lolref_scene_load_string(raw, len(raw))
        if not self.ptr:
            raise RuntimeError("reference failed to load the scene")

    def set_camera(self, point, direction):
        p = (C.c_float * 3)(*point)
        d = (C.c_float * 3)(*direction)
        self.lib.lolref_scene_set_camera(self.ptr, C.byref(p), C.byref(d))

    def flatten(self):
        """The reference's structs through the backend's translation -> SceneStruct pointer."""
        from loltracer_b200.api import SceneStruct, Scene

        p = self.lib.lolref_scene_flatten(self.ptr)
        return Scene(C.cast(p, C.POINTER(SceneStruct)))

    def probe(self, w, h, y0=0, y1=None, ystride=1, threads=None):
        y1, rows = _rows(h, y0, y1, ystride)
        dist = np.zeros((rows, w), np.float32)
        ids = np.zeros((rows, w), np.uint32)
        rgba = np.zeros((rows, w), np.uint32)
        ms = self.lib.lolref_probe(self.ptr, w, h, y0, y1, ystride, threads or nthreads(),
                                dist.ctypes.data, ids.ctypes.data, rgba.ctypes.data)
        return dict(dist=dist, id=ids, rgba=rgba, ms=ms)

    def render_protocol(self, w, h, threads=None, frames=1):
        """Through the unmodified render_thread() and main.c's semaphore protocol."""
        px = np.zeros((h, w), np.uint32)
        ms = (C.c_double * frames)()
        self.lib.lolref_render_protocol(self.ptr, w, h, threads or nthreads(), frames,
                                     px.ctypes.data, C.cast(ms, C.c_void_p))
        return px, list(ms)

    def __del__(self):
        if getattr(self, "ptr", None) and _ref is not None:
            _ref.lolref_scene_free(self.ptr)
            self.ptr = None


def frame_hash(px: np.ndarray) -> str:
    """h = h*1000003 + pixel over row-major pixels, 64-bit wrap."""
    v = np.ascontiguousarray(px, dtype=np.uint32).ravel().astype(np.uint64)
    n = v.size
    mult = np.uint64(1000003)
    # Horner in blocks: h = sum v[i] * m^(n-1-i)  (mod 2^64)
    with np.errstate(over="ignore"):
        pw = np.empty(n, np.uint64)
        # powers by repeated doubling
        pw[0] = 1
        filled = 1
        while filled < n:
            step = min(filled, n - filled)
            pw[filled:filled + step] = pw[:step] * (pw[filled - 1] * mult)
            filled += step
        h = np.sum(v * pw[::-1], dtype=np.uint64)
    return f"{int(h):016x}"


def compare_frames(got_rgba, got_id, want_rgba, want_id):
    """North-star tolerances: hit/miss mask agreement and per-channel RGB error on
    agreeing hits.  Returns dict(mask_agree, max_rgb_err, n_rgb_off, n_mask_off)."""
    got_hit = got_id != 0
    want_hit = want_id != 0
    agree = got_hit == want_hit
    both = got_hit & want_hit
    def ch(a, s):
        return ((a >> s) & 0xFF).astype(np.int32)
    err = np.zeros(got_rgba.shape, np.int32)
    for s in (16, 8, 0):
        err = np.maximum(err, np.abs(ch(got_rgba, s) - ch(want_rgba, s)))
    max_err = int(err[both].max()) if both.any() else 0
    miss_both = (~got_hit) & (~want_hit)
    miss_err = int(err[miss_both].max()) if miss_both.any() else 0
    return dict(mask_agree=float(agree.mean()), n_mask_off=int((~agree).sum()),
                max_rgb_err=max_err, n_rgb_off=int((err[both] > 0).sum()),
                max_miss_rgb_err=miss_err, id_agree=float((got_id == want_id).mean()))


# ---- the generated distance code, compiled for the host -------------------------

HOST_SHIM = r"""
#define LOL_HOST_SHIM 1
#include <cmath>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define __noinline__
typedef unsigned long long lol_u64_shim;
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
#define __constant__ static const
#define __align__(n) __attribute__((aligned(n)))
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __saturatef(float v) { return (v > 0.f) ? ((v < 1.f) ? v : 1.f) : 0.f; }
static inline float lol_sqrt_fast(float x) { return std::sqrt(x); }
static inline float lol_min_nan(float a, float b) { return (a != a || b != b) ? NAN : std::fmin(a, b); }
static inline float lol_max_nan(float a, float b) { return (a != a || b != b) ? NAN : std::fmax(a, b); }
static inline float lol_fma(float a, float b, float c) { return std::fma(a, b, c); }
#define __fmaf_rn lol_fma
static inline int __float2int_rz(float f) { return (int)f; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
"""


def cpu_sdf(tmp_path, src, tag):
    """Compiles everything above the pipeline (helpers + generated lol_sdf) for the host:
    eval/eval2 for tests, lol_spec_sdf for the oracle's mode 2 (the JIT-equivalent CPU baseline)."""
    import pathlib
    import subprocess
    tmp_path = pathlib.Path(tmp_path)
    head = src.split("//@@SCENE@@")[0]
    cu = tmp_path / f"sdf_{tag}.cpp"
    pair = """
// the two-rays-per-call form (variant 3): points 2i and 2i+1 share one evaluation
// hints: whatever won at the previous pair -- arbitrary for random points, and the result
// must not depend on them
extern "C" void eval2(const float* p, int n, float* d, unsigned* id) {
  unsigned ha = 0, hb = 0;
  for (int i = 0; i + 1 < n; i += 2) {
    const lol_f2 r = lol_sdf2(lol_pk(p[3*i], p[3*i+3]), lol_pk(p[3*i+1], p[3*i+4]),
                              lol_pk(p[3*i+2], p[3*i+5]), ha, hb, id[i], id[i+1]);
    d[i] = lol_lo(r); d[i+1] = lol_hi(r);
    ha = id[i + 1]; hb = (i % 6 == 0) ? id[i] : 0u;   // crossed over, sometimes none
  }
}
""" if "lol_sdf2(" in head else ""
    cu.write_text(HOST_SHIM + head + """
extern "C" void eval(const float* p, int n, float* d, unsigned* id) {
  unsigned hint = 0;  // the previous point's winner: arbitrary here, must not matter
  for (int i = 0; i < n; ++i) { d[i] = lol_sdf(p[3*i], p[3*i+1], p[3*i+2], hint, id[i]); hint = (i % 5) ? id[i] : 0u; }
}
extern "C" float lol_spec_sdf(float x, float y, float z, unsigned* id) { return lol_sdf(x, y, z, 0u, *id); }
""" + pair)
    so = tmp_path / f"sdf_{tag}.so"
    subprocess.check_call(["g++", "-O2", "-msse4.2", "-mavx2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(cu)])
    return C.CDLL(str(so))


def specialised_sdf(scene, workdir, tag="spec"):
    """JIT-equivalent: the lowering's straight-line distance function for `scene`
    (IEEE forms, constants baked in), compiled by g++ and registered with the oracle
    as mode 2.  Returns the CDLL (keep it alive while mode 2 is in use)."""
    import loltracer_b200 as lb

    src = lb.lower_cuda(scene, lb.Options.default(variant=1, guarded_fastpath=0, prune_bounds=0))
    L = cpu_sdf(workdir, src, tag)
    L.lol_spec_sdf.restype = C.c_float
    port().lolo_set_specialised_sdf(C.cast(L.lol_spec_sdf, C.c_void_p))
    return L


# ---- the whole per-pixel pipeline of variant 1, compiled for the host ------------------

PIPELINE_WRAPPER = r"""
#include <vector>
static int lol_host_cap_primary = 256, lol_host_cap_shadow = 128;
static long lol_host_deferrals = 0;
extern "C" void lol_host_set_caps(int primary, int shadow) { lol_host_cap_primary = primary; lol_host_cap_shadow = shadow; }
extern "C" long lol_host_get_deferrals(void) { return lol_host_deferrals; }
// cb: origin[3] dir[3] right[3] up[3] width height (lolb200_camera_basis)
extern "C" void lol_host_render(const float* cb, int w, int h, unsigned* rgba, float* dist, unsigned* id,
                                unsigned short* nprimary, unsigned short* nshadow) {
  lol_params P;
  std::memset(&P, 0, sizeof P);
  P.ox = cb[0]; P.oy = cb[1]; P.oz = cb[2];
  P.dx = cb[3]; P.dy = cb[4]; P.dz = cb[5];
  P.rx = cb[6]; P.ry = cb[7]; P.rz = cb[8];
  P.ux = cb[9]; P.uy = cb[10]; P.uz = cb[11];
  P.cw = cb[12]; P.ch = cb[13];
  P.fw = (float)w; P.fh = (float)h; P.w = w; P.h = h; P.world = 1;
  P.rshift = 16; P.gshift = 8; P.bshift = 0; P.amask = 0xFF000000u;  // XRGB8888, alpha forced
#if LOL_VARIANT == 1
  lol_host_prologue(P);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      lol_pixel_out o;
      lol_shade_pixel(P, x, y, o);
      const size_t i = (size_t)y * w + x;
      rgba[i] = o.pixel; dist[i] = o.dist; id[i] = o.id;
      nprimary[i] = (unsigned short)o.n_primary; nshadow[i] = (unsigned short)o.n_shadow;
    }
#elif LOL_VARIANT == 4  // staged: the resumable pipeline -- cap every march, put unfinished pixels aside, resume
  std::vector<lol_cont> queue;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      lol_cont c;
      lol_cont_begin(c, x, y);
      queue.push_back(c);
    }
  lol_host_deferrals = 0;
  while (!queue.empty()) {  // pass after pass, like launch after launch
    std::vector<lol_cont> next;
    for (lol_cont& c : queue) {
      lol_pixel_out o;
      if (!lol_pixel_run(P, c, o, lol_host_cap_primary, lol_host_cap_shadow)) {
        next.push_back(c);
        ++lol_host_deferrals;
        continue;
      }
      const size_t i = (size_t)(c.xy >> 16) * w + (c.xy & 0xffffu);
      rgba[i] = o.pixel; dist[i] = o.dist; id[i] = o.id;
      nprimary[i] = (unsigned short)o.n_primary; nshadow[i] = (unsigned short)o.n_shadow;
    }
    queue.swap(next);
  }
#else  // variant 3: two horizontally adjacent pixels per call, the second one absent at an odd right edge
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; x += 2) {
      lol_pixel_out o[2];
      const bool actB = x + 1 < w;
      lol_shade_pair(P, x, y, actB, o[0], o[1]);
      for (int k = 0; k < (actB ? 2 : 1); ++k) {
        const size_t i = (size_t)y * w + x + k;
        rgba[i] = o[k].pixel; dist[i] = o[k].dist; id[i] = o[k].id;
        nprimary[i] = (unsigned short)o[k].n_primary; nshadow[i] = (unsigned short)o[k].n_shadow;
      }
    }
#endif
}
"""


def cpu_pipeline(tmp_path, src, tag):
    """Compiles a generated program (variant 1, or variant 3: two rays per call on the shim's plain pairs)
    WHOLE for the host -- helpers, the generated distance code and lol_kernel.cuh's per-pixel pipeline
    (lol_shade_pixel / lol_shade_pair: camera ray, marches, normal, shadows, Phong, gamma, pack), everything
    but the kernel around it -- so that kernel-side logic is checked against the oracle without a GPU.
    Returns a CDLL with lol_host_render."""
    import pathlib
    import subprocess
    assert "#define LOL_VARIANT 2" not in src, "variant 2 lives in shared-memory queues"
    tmp_path = pathlib.Path(tmp_path)
    cu = tmp_path / f"pipe_{tag}.cpp"
    cu.write_text(HOST_SHIM + src + PIPELINE_WRAPPER)
    so = tmp_path / f"pipe_{tag}.so"
    subprocess.check_call(["g++", "-O2", "-msse4.2", "-mavx2", "-ffp-contract=off", "-shared", "-fPIC",
                           "-o", str(so), str(cu)])
    return C.CDLL(str(so))


def cpu_pipeline_render(L, lb, scene, w, h, camera=None):
    cb = lb.camera_basis(camera or scene.camera, w, h)
    flat = (C.c_float * 14)(*list(cb.origin), *list(cb.dir), *list(cb.right), *list(cb.up), cb.width, cb.height)
    rgba = np.zeros((h, w), np.uint32)
    dist = np.zeros((h, w), np.float32)
    ids = np.zeros((h, w), np.uint32)
    npr = np.zeros((h, w), np.uint16)
    nsh = np.zeros((h, w), np.uint16)
    L.lol_host_render(flat, w, h, rgba.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p),
                      ids.ctypes.data_as(C.c_void_p), npr.ctypes.data_as(C.c_void_p), nsh.ctypes.data_as(C.c_void_p))
    return dict(rgba=rgba, dist=dist, id=ids, nprimary=npr, nshadow=nsh)
