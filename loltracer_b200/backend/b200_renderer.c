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
 *    pulling scanlines from current_line (naive_renderer.c:215-216).  Here the
 *    unit a worker pulls is a GPU's SHARE of the frame, not a scanline:
 *      one GPU (or an NVLink gather, --gather nccl|peer): one share.  The worker
 *        that pulls it is the frame leader: it renders the frame and copies it into
 *        surf->pixels.
 *      --gpus N --gather host: N shares, share d = the cyclic 4-row bands of GPU d.
 *        Each worker pulls shares with SDL_AtomicAdd(&current_line, 1) until none
 *        is left, enqueues the launches and copies of the GPUs it pulled, then waits
 *        for them: with N workers every GPU has its own host thread issuing its API
 *        calls (one leader issuing all of them was the limit at 8 GPUs), with fewer
 *        workers each drives several GPUs, with more the spare ones answer at once.
 *    A wake-up that finds no share left (exactly what a late worker of the naive
 *    renderer sees) answers at once.  Every wake-up posts `exit` exactly once, and
 *    only after the rows it is responsible for are in host memory.
 *  - Flags for the backend sit in argv[3..] (tracing_jit_renderer.dasc:424-428).
 *  - No error channel: message on stderr and exit(1) (main.c:114-118).
 */
#include <pthread.h>
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
	int shares;          /* units of work per frame pulled through current_line */
	int pin_surface;     /* --pin-surface: page-lock surf->pixels (see render_prepare) */
	int wrap_devices;    /* --wrap-devices: --gpus N on fewer GPUs (shares wrap around) */
	void* pinned;        /* the surface pinned by us */
	size_t pinned_bytes;
	pthread_mutex_t pin_lock;
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
	const char* dump_ptx = NULL;
	const char* dump_sass = NULL;
	lolb200_scene* flat;

	lolb200_options_default(&st->options);
	pthread_mutex_init(&st->pin_lock, NULL);
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
		else if (!strcmp(argv[i], "--dump-ptx") && i + 1 < argc)
			dump_ptx = argv[++i];
		else if (!strcmp(argv[i], "--dump-sass") && i + 1 < argc)
			dump_sass = argv[++i];
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
		else if (!strcmp(argv[i], "--wrap-devices"))
			/* --gpus N with fewer than N GPUs in the box: share i runs on GPU i mod count (for
			 * trying the N-share protocol on a small box; NCCL refuses a GPU twice) */
			st->wrap_devices = 1;
		else if (!strcmp(argv[i], "--pin-surface"))
			/* Page-lock surf->pixels so that the GPUs' copy engines write straight into it
			 * (saves the staging copy: matters from about 1080p on).  The surface is SDL's,
			 * not ours: only for hosts that keep a surface allocated until the frame after
			 * they stop handing it in (a fixed-size window, the headless host).  Without the
			 * flag frames go through pinned staging memory the backend owns. */
			st->pin_surface = 1;
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
		for (int i = 0; i < st->gpus; i++) {
			devices[i] = st->device + i;
			if (st->wrap_devices && lolb200_device_count() > 0)
				devices[i] %= lolb200_device_count();
		}
		if (lolb200_group_create(flat, &st->options, devices, st->gpus, st->gather, &st->group) !=
		    LOLB200_OK)
			b200_die("render_prepare");
	}
	st->shares = (st->group && st->gather == LOLB200_GATHER_HOST) ? lolb200_group_size(st->group) : 1;
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
	if (dump_ptx) {
		char* ptx = NULL;
		size_t n = 0;
		if (lolb200_compile_ptx(lolb200_renderer_source(st->renderer), &st->options, &ptx, &n) != LOLB200_OK)
			b200_die("--dump-ptx");
		b200_dump(dump_ptx, ptx, n);
		lolb200_free(ptx);
	}
	if (dump_sass) {
		/* what the GPU executes: the analogue of the JIT's machine code in the jitdump file */
		size_t n = 0, len = 0;
		const void* img = lolb200_renderer_image(st->renderer, &n);
		char* sass = NULL;
		if (lolb200_disassemble(img, n, &sass, &len) != LOLB200_OK)
			b200_die("--dump-sass");
		b200_dump(dump_sass, sass, len);
		lolb200_free(sass);
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
	if (st->pinned)
		lolb200_surface_unpin(st->pinned);
	lolb200_group_destroy(st->group);
	lolb200_renderer_destroy(st->renderer);
	free(st);
	data->private = NULL;
}

/* What every share of a frame needs: the surface's packing and this frame's camera. */
static void b200_frame_args(struct render_data* data, lolb200_camera* cam, lolb200_pixfmt* fmt) {
	const SDL_PixelFormat* f = data->surf->format;

	if (f->BytesPerPixel != 4) {
		/* the reference stores a Uint32 per pixel whatever the format says
		 * (naive_renderer.c:233-235); only 32-bit surfaces make sense */
		fprintf(stderr, "b200_renderer: %d bytes per pixel; only 32-bit surfaces are supported\n",
		        f->BytesPerPixel);
		exit(1);
	}
	memset(fmt, 0, sizeof *fmt);
	fmt->rshift = f->Rshift;
	fmt->gshift = f->Gshift;
	fmt->bshift = f->Bshift;
	fmt->rloss = f->Rloss;
	fmt->gloss = f->Gloss;
	fmt->bloss = f->Bloss;
	fmt->amask = f->Amask;
	lolb200__camera_out(&data->scene->camera, cam); /* main.c:180 moves it every frame */
}

/* --pin-surface: follow the surface main.c hands in (main.c:182 re-fetches it every frame).
 * Every worker passes here before it enqueues anything of a frame; the first one to notice a
 * new surface re-pins it.  At that moment nothing is in flight: main posts a frame's tokens
 * only after every wake-up of the previous frame has answered (main.c:189-194), and no worker
 * of this frame enqueues before it has been through this lock. */
static void b200_follow_surface(struct b200_state* st, SDL_Surface* surf) {
	const size_t bytes = (size_t)surf->pitch * (size_t)surf->h;
	if (!st->pin_surface)
		return;
	pthread_mutex_lock(&st->pin_lock);
	if (st->pinned != surf->pixels || st->pinned_bytes != bytes) {
		if (st->pinned)
			lolb200_surface_unpin(st->pinned);
		st->pinned = NULL;
		if (lolb200_surface_pin(surf->pixels, bytes) == LOLB200_OK) {
			st->pinned = surf->pixels;
			st->pinned_bytes = bytes;
		} /* else: somebody else pinned it, or it cannot be pinned: staging still works */
	}
	pthread_mutex_unlock(&st->pin_lock);
}

int render_thread(void* ptr) {
	struct render_data* data = ptr;

	while (true) {
		SDL_SemWait(frame_entry_barrier);
		if (SDL_AtomicGet(&exiting))
			return 0;

		struct b200_state* st = data->private;
		SDL_Surface* surf = data->surf;
		const int drawable = surf->w > 0 && surf->h > 0; /* a minimised window has no rows */
		int mine[64], n = 0, share;
		lolb200_camera cam;
		lolb200_pixfmt fmt;

		/* Pull shares like the reference pulls scanlines (naive_renderer.c:215-216). */
		while ((share = SDL_AtomicAdd(&current_line, 1)) < st->shares) {
			if (!drawable)
				continue;
			if (n == 0) {
				b200_frame_args(data, &cam, &fmt);
				b200_follow_surface(st, surf);
			}
			if (st->shares == 1) {
				/* the frame leader: one GPU, or an NVLink gather driven by one thread */
				int rc = st->group
					? lolb200_group_render_host(st->group, &cam, surf->w, surf->h, &fmt, surf->pixels,
					                            (size_t)surf->pitch)
					: lolb200_render_host(st->renderer, &cam, surf->w, surf->h, &fmt, surf->pixels,
					                      (size_t)surf->pitch);
				if (rc != LOLB200_OK)
					b200_die("render_thread");
				continue;
			}
			if (lolb200_group_share_enqueue(st->group, share, &cam, surf->w, surf->h, &fmt, surf->pixels,
			                                (size_t)surf->pitch) != LOLB200_OK)
				b200_die("render_thread");
			mine[n++] = share;
		}
		for (int i = 0; i < n; i++)
			if (lolb200_group_share_wait(st->group, mine[i]) != LOLB200_OK)
				b200_die("render_thread");

		SDL_SemPost(frame_exit_barrier);
	}
}
