/*
 * lolb200.h -- C ABI of the B200 (sm_100a) sphere-tracing backend for loltracer.
 *
 * Plain C: pointers, sizes and PODs only; no CUDA, torch or SDL type crosses this
 * boundary.  Every entry point names the reference interface it stands in for
 * (paths relative to the reference tree, iglosiggio/loltracer).
 *
 * Layering (bottom-up):
 *   1. scene description  (PODs mirroring scene.h:44-96, flattened: no pointers
 *                          into the reference's `struct vector`)
 *   2. .lol front-end     (replaces scene_parse(), scene-parser.y:197-214)
 *   3. lowering           (scene -> specialised CUDA C; replaces generate_sdf(),
 *                          tracing_jit_renderer.dasc:76-143)
 *   4. device layer       (NVRTC sm_100a + launch; replaces link_and_encode(),
 *                          tracing_jit_renderer.dasc:60-74, and the pixel loop of
 *                          render_thread(), naive_renderer.c:195-240)
 *
 * The renderer.h drop-in itself (render_thread / render_prepare / render_destroy,
 * renderer.h:24-26) lives in loltracer_b200/backend/b200_renderer.c and is built
 * on this ABI only.
 *
 * Error convention: functions returning int return 0 on success and a negative
 * LOLB200_E* code on failure; lolb200_last_error() gives the message for the
 * calling thread.  There is no CPU fallback anywhere: device functions fail with
 * LOLB200_ENODEVICE when no CUDA device is usable.
 */
#ifndef LOLB200_H
#define LOLB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LOLB200_ABI_VERSION 3 /* 3: lolb200_options grew (guard_out, grid_cells) */

enum {
	LOLB200_OK = 0,
	LOLB200_EPARSE = -1,    /* .lol syntax / semantic error                    */
	LOLB200_EINVAL = -2,    /* bad argument                                    */
	LOLB200_ENODEVICE = -3, /* no usable CUDA device / driver                  */
	LOLB200_ECOMPILE = -4,  /* NVRTC rejected the generated source             */
	LOLB200_ECUDA = -5,     /* a CUDA call failed                              */
	LOLB200_ENOMEM = -6
};

/* ---------------------------------------------------------------- 1. scene -- */

/* Values equal enum components in scene.h:27-35 so a reference `struct object`
 * translates with a plain copy of `type`. */
enum lolb200_object_type {
	LOLB200_OBJ_SPHERE = 3,
	LOLB200_OBJ_BOX = 4,
	LOLB200_OBJ_PLANE = 5,
	LOLB200_OBJ_SMOOTH_UNION = 6,
	/* Extensions beyond the reference grammar (scene.h:27-35 ends at 6; BASELINE
	 * config C4 asks for "unioned/intersected primitives"): hard CSG nodes with
	 * children a, b like a smooth union.  union = minf(a, b), intersection =
	 * maxf(a, b), difference = maxf(a, -b), with float.h's MINSS/MAXSS operand
	 * rules.  The reference cannot render them, so their oracle is our own
	 * restatement only (oracle/lol_oracle.c; DESIGN.md "extensions"). */
	LOLB200_OBJ_UNION = 7,
	LOLB200_OBJ_INTERSECTION = 8,
	LOLB200_OBJ_DIFFERENCE = 9
};
#define LOLB200_OBJ_HAS_CHILDREN(t) ((t) >= LOLB200_OBJ_SMOOTH_UNION && (t) <= LOLB200_OBJ_DIFFERENCE)

/* scene.h:44-49 */
typedef struct lolb200_material {
	float shininess;
	float diffuse[3];
	float specular[3];
	float ambient[3];
} lolb200_material;

/* scene.h:52-56 */
typedef struct lolb200_light {
	float point[3];
	float diffuse_intensity[3];
	float specular_intensity[3];
} lolb200_light;

/* scene.h:58-82.  Children of a smooth union are indices into the node array
 * (the reference holds malloc'ed pointers, scene.c:18-29). */
typedef struct lolb200_object {
	int32_t type;      /* enum lolb200_object_type                        */
	uint32_t material; /* only read for top-level objects (naive_renderer.c:102-112) */
	float point[3];    /* plane: (0, y, 0) as scene.c:215                 */
	float radius;      /* sphere radius / box rounding radius             */
	float point2[3];   /* box half extents                                */
	float smoothness;  /* smooth union k                                  */
	int32_t a, b;      /* smooth union children (node indices), else -1   */
} lolb200_object;

/* scene.h:84-88.  direction is unit length and fov is in radians, exactly as
 * camera_from_definition_list leaves them (scene.c:173-174). */
typedef struct lolb200_camera {
	float point[3];
	float direction[3];
	float fov;
} lolb200_camera;

/* scene.h:90-96, flattened.  `objects[i]` is the node index of top-level object
 * with id i+1 (id 0 = nothing, naive_renderer.c:30-44). */
typedef struct lolb200_scene {
	uint32_t n_materials;
	lolb200_material* materials;
	float ambient_color[3];
	uint32_t n_lights;
	lolb200_light* lights;
	uint32_t n_nodes;
	lolb200_object* nodes;
	uint32_t n_objects;
	uint32_t* objects;
	lolb200_camera camera;
} lolb200_scene;

/* ------------------------------------------------------------ 2. front-end -- */

/* Replaces scene_parse() (scene-parser.y:197-214) + the property extractors of
 * scene.c:140-281 + scene_validate_materials() (scene.c:284-292, main.c:235).
 * Accepts exactly the token set of scene-lexer.l:10-50 and the grammar of
 * scene-parser.y:73-189.  Where the reference calls exit(1)/assert (unknown
 * property, wrong value type, bad material index) this returns LOLB200_EPARSE. */
int lolb200_scene_parse_file(const char* path, lolb200_scene** out);
int lolb200_scene_parse_string(const char* text, size_t len, lolb200_scene** out);
/* Deep copy / free (scene_free(), scene.c:60-65). */
lolb200_scene* lolb200_scene_clone(const lolb200_scene* s);
void lolb200_scene_free(lolb200_scene* s);

/* Camera basis exactly as get_camera_ray() derives it per pixel
 * (naive_renderer.c:178-193), hoisted to once per frame: this is the only place
 * the path calls atanf, so it stays on the host (glibc) for parity.
 * rd(x,y) = normalize((right*(vx*width) + up*(vy*height)) + dir). */
typedef struct lolb200_camera_basis {
	float origin[3];
	float dir[3];
	float right[3];
	float up[3];
	float width;  /* aspect * atanf(fov/2)  */
	float height; /* atanf(fov/2)           */
} lolb200_camera_basis;
void lolb200_camera_basis_compute(const lolb200_camera* cam, int w, int h,
                                  lolb200_camera_basis* out);

/* ------------------------------------------------------------- 3. lowering -- */

enum lolb200_arith {
	/* Operation order and rounding of the reference (no FMA contraction, IEEE
	 * div/sqrt, MINSS/MAXSS NaN rules): the parity mode. */
	LOLB200_ARITH_EXACT = 0,
	/* --fmad=true, reciprocal multiplies, FMNMX min/max: not bit-faithful. */
	LOLB200_ARITH_FAST = 1
};

typedef struct lolb200_options {
	int32_t arith;           /* enum lolb200_arith                              */
	int32_t skip_black_miss; /* 1: allow the exact miss-pixel shortcut when
	                            material 0 is all-zero; 0: always shade misses  */
	int32_t cull_backfacing; /* 1: skip a light's shadow march when n.l <= 0
	                            (its contribution is exactly +0)                */
	int32_t shadow_early_out;/* 1: leave the shadow march once res <= 0 (the
	                            returned value is then exactly 0)               */
	int32_t counters;        /* 1: kernel also accumulates per-phase SDF
	                            evaluation counts (instrumented build)          */
	int32_t variant;         /* kernel structure: 0 = chosen per scene,
	                            1 = phase-sequential, 2 = megaloop + lane refill,
	                            3 = two rays per thread in packed FP32 registers
	                                (FADD2/FMUL2/FFMA2; needs the guarded forms),
	                            4 = deferred long rays: a march that is not over
	                                after defer_cap_* evaluations puts its pixel
	                                aside in a global queue; a second launch
	                                (lol_resume) finishes those pixels, one per
	                                lane (exact: same operations per ray)         */
	int32_t loop_threshold;  /* top-level runs of >= this many same-shape
	                            objects become a loop over __constant__ tables;
	                            0 = default (16)                                */
	int32_t guarded_fastpath;/* exact mode only.  1: sqrt and the division by a
	                            smoothness constant run without their per-
	                            operation special-case branches, under one range
	                            guard per sdf() evaluation that falls back to
	                            the IEEE forms; results are bit-identical.
	                            1 = where it pays (>= 3 spheres or a smooth
	                            union), 2 = always, 0 = never                   */
	int32_t prune_bounds;    /* 1: an object (or a group of them) is skipped when
	                            its bounding box proves it cannot beat the
	                            running minimum (exact): always
	                            in table loops, in straight-line code where a
	                            sampled estimate says the test pays -- a box, or
	                            a ball around one of the object's own sphere
	                            centres (four instructions instead of sixteen);
	                            2: every straight-line box test on; 3: as 1, but
	                            boxes only; 4: as 2, with the ball in place of
	                            the box wherever an object has one (tests);
	                            0: off                                          */
	int32_t block_threads;   /* tuning: threads per CTA (multiple of 32);
	                            0 = the variant's default                       */
	int32_t min_blocks;      /* tuning: __launch_bounds__ second argument (caps
	                            registers so that this many CTAs fit on an SM);
	                            0 = the variant's default                       */
	int32_t roll_phases;     /* tuning, variant 3: 1 = the normal taps and the
	                            lights are loops around ONE copy of the distance
	                            code each (instruction-cache footprint), 0 =
	                            unrolled; -1... not used; default 1             */
	int32_t prune_group;     /* tuning: objects per group of the two-level box
	                            pruning in table loops; 0 = default (8)         */
	int32_t pack_pairs;      /* exact mode with the guarded forms: two subtrees of one
	                            shape inside an object (two spheres, two smooth
	                            unions of spheres, ...) are evaluated as ONE packed
	                            FP32 instruction stream (FADD2 / FMUL2 / FFMA2, one
	                            subtree per half; each half rounds like the scalar
	                            instruction, so results are bit-identical).
	                            1 = where it pays (inside table loops), 2 =
	                            everywhere, 3 = everywhere, leaves only, 0 = never */
	int32_t share_first_step;/* step 1 of every primary ray evaluates sdf() at the
	                            camera position (ro + rd * 0): it is evaluated once
	                            per CTA and shared (variant 1; exact, see
	                            lol_kernel.cuh).  1 = where it pays (scenes without
	                            table loops), 2 = always, 0 = never               */
	int32_t shadow_div_pretest; /* 1: the shadow march divides (50 * d) / t only when
	                            the quotient can lower res; a multiplication with a
	                            2^-21 margin proves the other steps (exact; variant
	                            1, exact arithmetic, with shadow_early_out).
	                            Measured slower on B200 (scene4 4K 2.149 -> 2.171
	                            ms: the divergent branch costs more than the six
	                            divisions in seven it saves), so off by default    */
	int32_t defer_cap_primary;  /* variant 4: evaluations of a primary / shadow march  */
	int32_t defer_cap_shadow;   /* before the pixel is put aside; 0 = default (48, 24) */
	int32_t roll_v1;            /* variants 1 and 4, instruction-cache footprint: 1 = the four
	                            normal taps are a loop around ONE copy of the distance
	                            code, 2 = the lights as well; 0 = everything unrolled;
	                            -1 = default                                        */
	int32_t loop_worklist;      /* pruned table loops: 1 = every lane collects the rows it
	                            cannot skip and the warp drains the lists together, each
	                            lane its own row per round (rounds = the longest list,
	                            not the union of all lanes' rows); 0 = the plain loops;
	                            -1 = default (0: measured 11 % slower on B200, the box
	                            tests dominate, not the rows).  Exact either way     */
	int32_t near_cache;         /* variant 1, scenes with one pruned table loop: a ray remembers
	                            the (up to four) rows it could not skip and how far it may
	                            move before the others must be looked at again, so most
	                            march steps evaluate the candidates and walk no group box
	                            at all (lol_kernel.cuh: struct lol_near; exact).
	                            1 = on, 2 = on, and the warp looks at all rows again
	                            whenever one of its lanes has to (B200, 4K: 1024 spheres
	                            31.5 -> 29.9 ms), 3 = and a look reads the point's cell
	                            of a candidate grid (two levels of 64^3 cells, each with
	                            its eight nearest rows, built on the device when the
	                            renderer is created) instead of walking every group,
	                            and up to eight rows are evaluated on the spot before
	                            the plain loop is called in, 0 = off, -1 = default (3) */
	int32_t guard_out;          /* variant 1 with the guarded forms: the march loops run the guarded
	                            arithmetic alone and leave the loop when the range guard
	                            fails; that one step is taken with the IEEE forms and the
	                            loop entered again -- the fall-back call and its
	                            reconvergence point are no longer part of every step
	                            (exact: same evaluations, same order).  Bits: 1 = the march
	                            loops, 2 = the four normal taps (one test of their four
	                            guards); 0 = off, -1 = default (3)                   */
	int32_t grid_cells;         /* tuning, near_cache = 3: cells per axis of the candidate grid
	                            (16, 32, 48 or 64); 0 = default (64: 2 x 64^3 cells of 48
	                            bytes, 25 MB of device memory per renderer)         */
	int32_t child_materials;    /* EXTENSION, off by default (the reference ignores the
	                            materials of a composite's children,
	                            naive_renderer.c:102-112): 1 = a hit on a composite
	                            object takes the material of the child that decides
	                            the node's distance at the hit point -- the nearer
	                            child of a (smooth) union, the farther of an
	                            intersection, a or the carved-out b of a difference,
	                            ties keep a -- evaluated once per hit pixel at
	                            p = ro + rd * dist with the IEEE forms.  A node whose
	                            material is #0 (the field's default) inherits its
	                            parent's.  Distances, ids and the hit mask do not
	                            change.  Checked against the oracle's own restatement
	                            only (oracle/lol_oracle.c: child_material)          */
} lolb200_options;
void lolb200_options_default(lolb200_options* o);

/* Replaces generate_sdf()/generate_obj_dist() (tracing_jit_renderer.dasc:76-216):
 * emits CUDA C with every scene constant baked in.  Pure host code, no device
 * needed.  Returns a malloc'ed NUL-terminated string (free with lolb200_free). */
char* lolb200_lower_cuda(const lolb200_scene* s, const lolb200_options* o, size_t* len);
void lolb200_free(void* p);

/* Algorithmic FLOPs of one sdf() evaluation with the convention of DESIGN.md
 * (sphere 10, round box 20, plane 1, smooth node 13, +1 per top-level object). */
uint64_t lolb200_scene_flops_per_eval(const lolb200_scene* s);

/* ---------------------------------------------------------- 4. device layer -- */

/* NVRTC for sm_100a; works without a GPU (cross-compile).  Replaces
 * link_and_encode() (tracing_jit_renderer.dasc:60-74).  *image is malloc'ed.
 * Environment: LOLB200_CACHE_DIR=<dir> keeps compiled programs as
 * <dir>/lol-<hash of program text, arithmetic mode, NVRTC version>.cubin and
 * reuses them (NVRTC costs 0.4-1.1 s per scene); LOLB200_DUMP_DIR=<dir> writes the
 * program as <dir>/lol-<hash>.cu and compiles it under that name so that
 * -lineinfo points at a file a profiler can import (the jitdump analogue,
 * jitdump.c:69-129). */
int lolb200_compile_cubin(const char* cuda_src, const lolb200_options* o,
                          void** image, size_t* image_size, char** log);

/* The jitdump analogues beyond the generated source (tracing_jit_renderer.dasc:424-433,
 * jitdump.c:93-120): the program's PTX (nvrtcGetPTX for compute_100a) and the SASS listing of a
 * compiled image (the toolkit's cuobjdump / nvdisasm run on a temporary copy).  Both return
 * malloc'ed NUL-terminated text (free with lolb200_free).  Host-side phases (lowering, NVRTC,
 * module load, launches, gathers, read-back) are also bracketed by NVTX ranges named
 * "lolb200: ..." for timeline profilers. */
int lolb200_compile_ptx(const char* cuda_src, const lolb200_options* o, char** ptx, size_t* len);
int lolb200_disassemble(const void* image, size_t image_size, char** sass, size_t* len);

int lolb200_device_count(void);

typedef struct lolb200_renderer lolb200_renderer; /* opaque */

/* render_prepare() analogue (tracing_jit_renderer.dasc:416-434): lower, compile,
 * load the module on `device`, allocate the work counter. */
int lolb200_renderer_create(const lolb200_scene* s, const lolb200_options* o,
                            int device, lolb200_renderer** out);
/* render_destroy() analogue. */
void lolb200_renderer_destroy(lolb200_renderer* r);
/* Generated source / SASS-bearing image, the analogue of `--jitdump`
 * (tracing_jit_renderer.dasc:424-433, jitdump.c:93-120). */
const char* lolb200_renderer_source(const lolb200_renderer* r);
const void* lolb200_renderer_image(const lolb200_renderer* r, size_t* size);
int lolb200_renderer_kernel_info(const lolb200_renderer* r, int* regs, int* smem_bytes,
                                 int* local_bytes, int* max_threads);

/* How 32-bit pixels are packed: what SDL_MapRGB(surf->format, r, g, b) does for
 * a non-palettised 32-bit format (renderer.h:17-22). */
typedef struct lolb200_pixfmt {
	uint8_t rshift, gshift, bshift;
	uint8_t rloss, gloss, bloss;
	uint16_t pad;
	uint32_t amask; /* OR-ed into every pixel */
} lolb200_pixfmt;
/* XRGB8888 with alpha forced to 0xFF: 0xFF000000 | r<<16 | g<<8 | b. */
void lolb200_pixfmt_default(lolb200_pixfmt* f);

/* Which rows of the frame this launch renders.  Row band b (band_rows rows)
 * belongs to rank b % world; a rank stores its bands compactly, band after band
 * ([local_band][band_rows][W] pixels), or -- with dst_full_frame -- at their
 * final position in a full W x H frame (its own or a peer-mapped one). */
typedef struct lolb200_shard {
	int32_t rank, world;
	int32_t band_rows;      /* multiple of 4; 0 = default (4)                  */
	int32_t dst_full_frame; /* 0: compact local buffer, 1: full-frame indexing */
	/* Completion flag of the peer-store gather (optional, NULL = none): a 32-bit word in
	 * device memory -- typically the CONSUMER's, mapped here through CUDA IPC or peer access --
	 * that receives done_value (release, system scope) once every pixel store of the launch is
	 * visible system-wide.  The consumer's stream waits for it with
	 * lolb200_stream_wait_value32: a stream memory operation over NVLink instead of a
	 * collective as the "all ranks have stored their bands" barrier.  Use one word per
	 * producer and a value that grows by one per frame. */
	void* done_flag;
	uint32_t done_value;
	uint32_t reserved;
} lolb200_shard;

/* Optional per-pixel auxiliaries for parity tests (not part of the product
 * frame): final march distance, object id (0 = miss), primary step count and
 * total shadow-march evaluations.  Any pointer may be NULL. Full-frame indexed. */
typedef struct lolb200_aux {
	float* dist;
	uint32_t* id;
	uint16_t* primary_steps;
	uint16_t* shadow_steps;
	/* launch probes (3 words of device memory, preset to {~0, 0, ~0}): nanoseconds of the GPU's
	 * global timer at [0] the first moment a warp found the work queue dry, [1] the last warp's
	 * exit, [2] the first CTA's start.  [1] - [0] is the tail of the launch. */
	uint64_t* launch_timing;
} lolb200_aux;

/* The pixel loop of render_thread() (naive_renderer.c:216-236) for one frame or
 * one shard of it, asynchronously on `stream` (a cudaStream_t passed as void*,
 * NULL = default stream).  dst_dev is DEVICE memory, pitch_px in pixels.
 * A renderer has one work queue: launches of one renderer are serialised on the
 * device (a launch on another stream than the previous one first waits for it),
 * and calls into one renderer must come from one host thread at a time.  Use one
 * renderer per stream to overlap frames. */
int lolb200_render_device(lolb200_renderer* r, const lolb200_camera* cam, int w, int h,
                          const lolb200_pixfmt* fmt, const lolb200_shard* shard,
                          void* dst_dev, size_t pitch_px, const lolb200_aux* aux_dev,
                          void* stream);

/* Same, end to end for a host surface (surf->pixels, surf->pitch in BYTES):
 * uploads the camera, renders on the renderer's device, copies the frame into
 * `pixels` honouring `pitch_bytes`, and returns when the pixels are visible to
 * the host.  This is what b200_renderer.c's frame leader calls.
 * The surface belongs to the caller (SDL frees and reallocates a window surface on
 * resize), so the library never page-locks it behind the owner's back.  On every
 * call it asks whether `pixels` is page-locked CUDA host memory right now: if so
 * the copy engine writes straight into it; if not, slabs arrive in a pinned frame
 * the renderer owns and the calling thread copies them on (overlapped with the
 * slabs still rendering). */
int lolb200_render_host(lolb200_renderer* r, const lolb200_camera* cam, int w, int h,
                        const lolb200_pixfmt* fmt, void* pixels, size_t pitch_bytes);

/* For owners of a surface: page-lock [pixels, pixels + bytes) so that frames are
 * DMA-ed straight into it (no staging copy).  The range must stay allocated until
 * lolb200_surface_unpin(pixels); never pin memory somebody else may free. */
int lolb200_surface_pin(void* pixels, size_t bytes);
int lolb200_surface_unpin(void* pixels);

/* Sum of the instrumented counters since the last call (options.counters = 1):
 * [0] primary evals, [1] normal-tap evals, [2] shadow evals, [3] pixels,
 * [4] hit pixels, [5] shadow rays marched, [6] shadow rays culled. */
int lolb200_read_counters(lolb200_renderer* r, uint64_t out[8]);

/* Rank 0's de-interleave after a gather: `gathered` holds world compact shards
 * back to back (each padded to shard_pixels), `frame` is the W x H result. */
int lolb200_deinterleave_device(const void* gathered_dev, void* frame_dev, int w, int h,
                                int world, int band_rows, size_t shard_pixels,
                                size_t pitch_px, void* stream);
/* One rank's bands of the frame, end to end into the FULL-FRAME host surface
 * `pixels` (row y of the frame at pixels + y * pitch_bytes): render, then the
 * rank's own copy engine writes each 4-row band to its final rows.  Ranks of a
 * multi-process job hand in the same shared-memory surface; nothing crosses
 * NVLink.  Synchronous for this rank. */
int lolb200_render_host_shard(lolb200_renderer* r, const lolb200_camera* cam, int w, int h,
                              const lolb200_pixfmt* fmt, const lolb200_shard* shard, void* pixels,
                              size_t pitch_bytes);
/* Pixels a rank's compact buffer needs (bands padded so every rank is equal). */
size_t lolb200_shard_pixels(int w, int h, int world, int band_rows);

/* ---- several GPUs driven by ONE process (the reference's main.c is one) -------
 * One renderer per device; the frame is cut into cyclic 4-row bands, band b on
 * devices[b % n]; devices[0] ends up with the complete frame.
 *   LOLB200_GATHER_NCCL: compact shards, ncclSend/ncclRecv in one group over
 *                        NVLink (ncclCommInitAll), de-interleave on devices[0];
 *   LOLB200_GATHER_PEER: every device's kernel stores its bands straight into
 *                        devices[0]'s frame through peer access.
 * render_host then copies the frame into the caller's surface like
 * lolb200_render_host does.
 *   LOLB200_GATHER_HOST: no device-side gather at all: every device copies its
 *                        own bands into the caller's (pinned, portable) surface
 *                        over its own PCIe link -- the fastest way to a frame
 *                        in HOST memory, which is what renderer.h asks for.
 * lolb200_group_render_host drives every device from the calling thread; calls
 * of it must come from one thread at a time.  A LOLB200_GATHER_HOST frame can also
 * be driven by several host threads, one SHARE (= one device's bands) each:
 * this is how main.c's N worker threads (main.c:147-149) each take a GPU in
 * b200_renderer.c instead of one leader issuing every launch and copy. */
enum { LOLB200_GATHER_NCCL = 0, LOLB200_GATHER_PEER = 1, LOLB200_GATHER_HOST = 2 };
typedef struct lolb200_group lolb200_group; /* opaque */
int lolb200_group_create(const lolb200_scene* s, const lolb200_options* o, const int* devices,
                         int n_devices, int gather, lolb200_group** out);
void lolb200_group_destroy(lolb200_group* g);
int lolb200_group_render_host(lolb200_group* g, const lolb200_camera* cam, int w, int h,
                              const lolb200_pixfmt* fmt, void* pixels, size_t pitch_bytes);
/* Share `share` (0 .. lolb200_group_size - 1) of one frame, from launch to its rows in
 * `pixels`: enqueue returns at once, wait returns when the rows are in host memory.
 * Different shares may be driven by different threads concurrently (all with the same
 * camera, size, format and surface); one share by one thread at a time. */
int lolb200_group_size(const lolb200_group* g);
int lolb200_group_share_enqueue(lolb200_group* g, int share, const lolb200_camera* cam, int w, int h,
                                const lolb200_pixfmt* fmt, void* pixels, size_t pitch_bytes);
int lolb200_group_share_wait(lolb200_group* g, int share);
/* Milliseconds (CUDA events on devices[0]) of the last lolb200_group_render_host frame:
 * render + gather, without the copy to the host. */
double lolb200_group_last_frame_ms(const lolb200_group* g);

/* Makes `stream` wait (a stream memory operation, no kernel, no host thread) until the 32-bit word
 * at dev_addr is >= value in the cyclic sense ((int32_t)(*dev_addr - value) >= 0): the consumer's
 * side of lolb200_shard.done_flag.  lolb200_stream_write_value32 stores a word in stream order
 * (e.g. to reset the flags). */
int lolb200_stream_wait_value32(void* stream, void* dev_addr, uint32_t value);
int lolb200_stream_write_value32(void* stream, void* dev_addr, uint32_t value);

/* CUDA-IPC plumbing for the peer-store variant (render fused with its gather):
 * export a 64-byte handle for a device allocation / map a peer's handle. */
int lolb200_ipc_export(void* dev_ptr, uint8_t handle[64]);
int lolb200_ipc_open(const uint8_t handle[64], void** dev_ptr);
int lolb200_ipc_close(void* dev_ptr);

/* Independent-chain FFMA microbenchmark: the measured FP32 denominator that
 * MEASURED_PEAKS.json lacks.  Returns TFLOP/s (2 FLOP per FFMA), <0 on error. */
double lolb200_measure_fp32_peak(int device, int iters, double* ms_out);

const char* lolb200_last_error(void);
int lolb200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LOLB200_H */
