/*
 * b200_renderer.c -- the renderer.h backend for NVIDIA B200: a third translation
 * unit beside naive_renderer.c and tracing_jit_renderer.dasc.
 *
 * Link exactly one backend with main.c (reference Makefile:10-13); this one also
 * needs liblolb200 (see INTEGRATION.md for the Makefile target).  It defines the
 * three functions of renderer.h:24-26 and uses the four globals main.c owns
 * (renderer.h:6-9).  Everything GPU-side goes through the C ABI of
 * include/lolb200.h; no CUDA type appears here.
 *
 * Protocol (main.c:139-149,161,182-194,166-170,213):
 *  - N worker threads exist BEFORE render_prepare(); they block on
 *    frame_entry_barrier and read data->private only after a wake-up.
 *  - Per frame main sets data->surf, zeroes current_line, posts `entry` N times
 *    and waits for N posts on `exit`.  The reference's workers share the frame by
 *    pulling scanlines from current_line (naive_renderer.c:215-216).  A GPU
 *    wants the whole frame in one launch, so the worker that moves current_line
 *    off zero is the frame LEADER: it renders the frame and copies it into
 *    surf->pixels; every other wake-up finds current_line >= height (exactly
 *    what a late worker of the naive renderer sees) and answers at once.  Every
 *    wake-up posts `exit` exactly once, the leader only after the pixels are in
 *    host memory.
 *  - Flags for the backend sit in argv[3..] (tracing_jit_renderer.dasc:424-428).
 *  - No error channel: message on stderr and exit(1) (main.c:114-118).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "renderer.h" /* the reference's */

#include "lolb200.h"
#include "scene_translate.h"

struct b200_state {
	lolb200_renderer* renderer; /* one GPU */
	lolb200_group* group;       /* --gpus N > 1: all of them, driven by the frame leader */
	lolb200_options options;
	int device;
	int gpus;
	int gather;
	int verbose;
};

static void b200_die(const char* what) {
	fprintf(stderr, "b200_renderer: %s: %s\n", what, lolb200_last_error());
	exit(1);
}

static void b200_dump(const char* path, const void* data, size_t len) {
	FILE* f = fopen(path, "wb");
	if (!f || fwrite(data, 1, len, f) != len) {
		perror(path);
		exit(1);
	}
	fclose(f);
}

void render_prepare(struct render_data* data, int argc, const char* argv[]) {
	struct b200_state* st = calloc(1, sizeof *st);
	const char* dump_cuda = NULL;
	const char* dump_cubin = NULL;
	lolb200_scene* flat;

	lolb200_options_default(&st->options);
	st->gather = LOLB200_GATHER_HOST;
	for (int i = 3; i < argc; i++) {
		if (!strcmp(argv[i], "--exact"))
			st->options.arith = LOLB200_ARITH_EXACT;
		else if (!strcmp(argv[i], "--fast"))
			st->options.arith = LOLB200_ARITH_FAST;
		else if (!strcmp(argv[i], "--no-skips"))
			st->options.skip_black_miss = st->options.cull_backfacing =
				st->options.shadow_early_out = 0;
		else if (!strcmp(argv[i], "--variant") && i + 1 < argc)
			st->options.variant = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--pack-pairs") && i + 1 < argc)
			st->options.pack_pairs = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--device") && i + 1 < argc)
			st->device = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--gpus") && i + 1 < argc)
			st->gpus = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--gather") && i + 1 < argc) {
			/* host (default): every GPU copies its own bands into surf->pixels over its own
			 * PCIe link; nccl / peer: the frame is first completed on the first GPU over NVLink */
			const char* m = argv[++i];
			st->gather = !strcmp(m, "peer") ? LOLB200_GATHER_PEER :
			             !strcmp(m, "nccl") ? LOLB200_GATHER_NCCL : LOLB200_GATHER_HOST;
		}
		else if (!strcmp(argv[i], "--dump-cuda") && i + 1 < argc)
			dump_cuda = argv[++i];
		else if (!strcmp(argv[i], "--dump-cubin") && i + 1 < argc)
			dump_cubin = argv[++i];
		else if (!strcmp(argv[i], "-j") || !strcmp(argv[i], "--jitdump")) {
			/* the JIT backend's introspection switch, kept with its meaning:
			 * leave the generated code where a profiler can find it */
			dump_cuda = "lol-b200-kernel.cu";
			dump_cubin = "lol-b200-kernel.cubin";
		} else if (!strcmp(argv[i], "--cache") && i + 1 < argc)
			/* compiled kernels are kept here, keyed by a hash of the generated program:
			 * the next start on the same scene skips NVRTC (lolb200.h, LOLB200_CACHE_DIR) */
			setenv("LOLB200_CACHE_DIR", argv[++i], 1);
		else if (!strcmp(argv[i], "--no-cache"))
			unsetenv("LOLB200_CACHE_DIR");
		else if (!strcmp(argv[i], "--verbose"))
			st->verbose = 1;
		/* anything else belongs to someone else (the JIT backend ignores
		 * unknown flags too) */
	}

	/* Materials, lights, objects and fov never change after the parse, so they
	 * are baked into the kernel here; the camera is re-read every frame. */
	flat = lolb200_scene_from_reference(data->scene);
	if (st->gpus > 1) {
		/* the image is sharded in cyclic 4-row bands over GPUs device..device+N-1;
		 * st->gather says how the bands reach surf->pixels */
		int devices[64];
		if (st->gpus > 64)
			st->gpus = 64;
		for (int i = 0; i < st->gpus; i++)
			devices[i] = st->device + i;
		if (lolb200_group_create(flat, &st->options, devices, st->gpus, st->gather, &st->group) !=
		    LOLB200_OK)
			b200_die("render_prepare");
	}
	/* also built with --gpus N: its source and image are what -j dumps */
	if (lolb200_renderer_create(flat, &st->options, st->device, &st->renderer) != LOLB200_OK)
		b200_die("render_prepare");
	lolb200_scene_free(flat);

	if (dump_cuda) {
		const char* src = lolb200_renderer_source(st->renderer);
		b200_dump(dump_cuda, src, strlen(src));
	}
	if (dump_cubin) {
		size_t n = 0;
		const void* img = lolb200_renderer_image(st->renderer, &n);
		b200_dump(dump_cubin, img, n);
	}
	if (st->verbose) {
		int regs = 0, smem = 0, local = 0, threads = 0;
		lolb200_renderer_kernel_info(st->renderer, &regs, &smem, &local, &threads);
		fprintf(stderr, "b200_renderer: device %d, kernel %d regs, %d B smem, %d B local\n",
		        st->device, regs, smem, local);
	}
	data->private = st;
}

void render_destroy(struct render_data* data) {
	struct b200_state* st = data->private;
	if (!st)
		return;
	lolb200_group_destroy(st->group);
	lolb200_renderer_destroy(st->renderer);
	free(st);
	data->private = NULL;
}

static void b200_render_frame(struct render_data* data) {
	struct b200_state* st = data->private;
	SDL_Surface* surf = data->surf;
	const SDL_PixelFormat* f = surf->format;
	lolb200_camera cam;
	lolb200_pixfmt fmt;

	if (f->BytesPerPixel != 4) {
		/* the reference stores a Uint32 per pixel whatever the format says
		 * (naive_renderer.c:233-235); only 32-bit surfaces make sense */
		fprintf(stderr, "b200_renderer: %d bytes per pixel; only 32-bit surfaces are supported\n",
		        f->BytesPerPixel);
		exit(1);
	}
	memset(&fmt, 0, sizeof fmt);
	fmt.rshift = f->Rshift;
	fmt.gshift = f->Gshift;
	fmt.bshift = f->Bshift;
	fmt.rloss = f->Rloss;
	fmt.gloss = f->Gloss;
	fmt.bloss = f->Bloss;
	fmt.amask = f->Amask;
	lolb200__camera_out(&data->scene->camera, &cam); /* main.c:180 moves it every frame */

	if (st->group) {
		if (lolb200_group_render_host(st->group, &cam, surf->w, surf->h, &fmt, surf->pixels,
		                              (size_t)surf->pitch) != LOLB200_OK)
			b200_die("render_thread");
	} else if (lolb200_render_host(st->renderer, &cam, surf->w, surf->h, &fmt, surf->pixels,
	                               (size_t)surf->pitch) != LOLB200_OK)
		b200_die("render_thread");
}

int render_thread(void* ptr) {
	struct render_data* data = ptr;

	while (true) {
		SDL_SemWait(frame_entry_barrier);
		if (SDL_AtomicGet(&exiting))
			return 0;

		/* Claim all scanlines at once; whoever saw 0 owns the frame.  (A
		 * minimised window has no rows: claim one anyway so only one thread
		 * leads, and render nothing.) */
		int rows = data->surf->h > 0 ? data->surf->h : 1;
		if (SDL_AtomicAdd(&current_line, rows) == 0 && data->surf->w > 0 && data->surf->h > 0)
			b200_render_frame(data);

		SDL_SemPost(frame_exit_barrier);
	}
}
