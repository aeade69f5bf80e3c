/*
 * ref_harness.c -- the reference's own naive renderer, compiled as a library.
 *
 * TEST INFRASTRUCTURE, never on the product path.  Only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load the result.
 *
 * This file holds NO copy of the reference: it #includes
 * /root/reference/naive_renderer.c where it lies (so its `static` functions are
 * reachable) and links /root/reference/scene.c unmodified; see oracle/Makefile.
 * The output goes to oracle/_ref/ (git-ignored, but it travels to the GPU box).
 *
 * What it adds around the reference:
 *   - the four globals main.c defines (main.c:21-24);
 *   - scene_parse() from loltracer_b200/backend/scene_parse_shim.c: our syntax
 *     tree replayed through the reference's scene.c API in the order the bison
 *     actions call it (scene-parser.y:73-145); flex/bison are not installed, so
 *     the reference's own scene_parse() cannot be generated;
 *   - a headless replay of main.c's frame protocol (main.c:139-149,161,189-194,
 *     166-170,213) around the unmodified render_thread();
 *   - a per-pixel probe that calls the reference's static pipeline functions in
 *     the order render_thread's loop body does (naive_renderer.c:218-235) and
 *     also returns (dist, id), which the image alone does not carry.
 */
#include <pthread.h>
#include <time.h>

#include "naive_renderer.c" /* the reference's, via -I/root/reference */

#include "../loltracer_b200/backend/scene_translate.h"

SDL_atomic_t exiting;
SDL_atomic_t current_line;
SDL_sem* frame_entry_barrier;
SDL_sem* frame_exit_barrier;

/* ---------------------------------------------------------------- loading -- */

/* scene_parse() / scene_parse_text(): our parser replayed through the
 * reference's scene.c API (loltracer_b200/backend/scene_parse_shim.c). */
struct scene* scene_parse(const char* filename);
struct scene* scene_parse_text(const char* text, size_t len);

static void* checked(struct scene* scene) {
	if (scene && !scene_validate_materials(scene)) { /* main.c:235 */
		fprintf(stderr, "lolref: material index out of range\n");
		scene_free(scene);
		return NULL;
	}
	return scene;
}

void* lolref_scene_load_string(const char* text, size_t len) { return checked(scene_parse_text(text, len)); }
void* lolref_scene_load(const char* path) { return checked(scene_parse(path)); }

void lolref_scene_free(void* scene) {
	if (scene)
		scene_free(scene);
}

/* The camera is the one thing main.c mutates between frames (main.c:71-112). */
void lolref_scene_set_camera(void* scene_, const float point[3], const float direction[3]) {
	struct scene* scene = scene_;
	scene->camera.point = (v3){point[0], point[1], point[2]};
	scene->camera.direction = (v3){direction[0], direction[1], direction[2]};
}

/* The reference's structs, translated with the code the renderer.h backend uses. */
lolb200_scene* lolref_scene_flatten(void* scene) { return lolb200_scene_from_reference(scene); }

/* ---------------------------------------------- main.c's protocol, headless -- */

static double now_ms(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void* thread_entry(void* p) {
	render_thread(p);
	return NULL;
}

/* Renders `frames` frames of w x h XRGB8888 into `pixels` through the unmodified
 * render_thread(); frame_ms[i] receives each frame's wall time.  Returns 0. */
int lolref_render_protocol(void* scene, int w, int h, int nthreads, int frames,
                           uint32_t* pixels, double* frame_ms) {
	SDL_PixelFormat fmt;
	SDL_Surface surf;
	struct render_data data = {.scene = scene};
	pthread_t* threads = malloc(sizeof *threads * nthreads);
	const char* argv[] = {"lolref", "0", "-"};

	lolb200_stub_format_xrgb8888(&fmt);
	memset(&surf, 0, sizeof surf);
	surf.format = &fmt;
	surf.w = w;
	surf.h = h;
	surf.pitch = w * 4;
	surf.pixels = pixels;

	SDL_AtomicSet(&exiting, 0);
	frame_entry_barrier = SDL_CreateSemaphore(0);
	frame_exit_barrier = SDL_CreateSemaphore(0);
	for (int i = 0; i < nthreads; i++) /* main.c:147-149: threads first */
		pthread_create(&threads[i], NULL, thread_entry, &data);
	render_prepare(&data, 3, argv); /* main.c:161 */

	for (int f = 0; f < frames; f++) {
		double t0;
		data.surf = &surf;               /* main.c:182 */
		SDL_AtomicSet(&current_line, 0); /* main.c:189 */
		t0 = now_ms();
		for (int i = 0; i < nthreads; i++)
			SDL_SemPost(frame_entry_barrier);
		for (int i = 0; i < nthreads; i++)
			SDL_SemWait(frame_exit_barrier);
		if (frame_ms)
			frame_ms[f] = now_ms() - t0;
	}

	SDL_AtomicSet(&exiting, 1); /* main.c:166-170 */
	for (int i = 0; i < nthreads; i++)
		SDL_SemPost(frame_entry_barrier);
	for (int i = 0; i < nthreads; i++)
		pthread_join(threads[i], NULL);
	render_destroy(&data); /* main.c:213 */
	SDL_DestroySemaphore(frame_entry_barrier);
	SDL_DestroySemaphore(frame_exit_barrier);
	free(threads);
	return 0;
}

/* ------------------------------------------------------- per-pixel probe -- */

struct probe_job {
	const struct scene* scene;
	int w, h, y0, y1, ystride;
	float* dist;
	uint32_t* id;
	uint32_t* rgba;
	SDL_atomic_t next;
	SDL_PixelFormat fmt;
};

/* The body of render_thread's pixel loop (naive_renderer.c:207-235), calling the
 * reference's own static functions; buffers are compact over the sampled rows. */
static void* probe_worker(void* p) {
	struct probe_job* job = p;
	const struct scene* scene = job->scene;
	float fwidth = job->w, fheight = job->h;
	v3 ro = scene->camera.point;
	float aspect_ratio = fwidth / fheight;
	int nrows = (job->y1 - job->y0 + job->ystride - 1) / job->ystride;
	int r;

	while ((r = SDL_AtomicAdd(&job->next, 1)) < nrows) {
		int y = job->y0 + r * job->ystride;
		for (int x = 0; x < job->w; x++) {
			v2 view_pos = (v2){
				(x + .5f) / fwidth * 2.f - 1.f,
				1.f - (y + .5f) / fheight * 2.f,
			};
			v3 rd = get_camera_ray(scene->camera, view_pos, aspect_ratio);
			struct world_dist hit = get_intersection(scene, ro, rd);
			v3 pt = v3add(ro, v3scale(rd, hit.dist));
			v3 n = get_normal(scene, pt, hit.dist);
			v3 colorf = get_light(scene, pt, n, hit.id);
			size_t o = (size_t)r * job->w + x;
			colorf = v3pow(colorf, 1.f / 2.2f);
			if (job->dist)
				job->dist[o] = hit.dist;
			if (job->id)
				job->id[o] = hit.id;
			if (job->rgba)
				job->rgba[o] = colorf_to_pixfmt(colorf, &job->fmt);
		}
	}
	return NULL;
}

/* Rows y0, y0+ystride, ... < y1 of the w x h frame.  Returns elapsed ms. */
double lolref_probe(void* scene, int w, int h, int y0, int y1, int ystride, int nthreads,
                    float* dist, uint32_t* id, uint32_t* rgba) {
	struct probe_job job = {.scene = scene, .w = w, .h = h, .y0 = y0, .y1 = y1,
	                        .ystride = ystride, .dist = dist, .id = id, .rgba = rgba};
	pthread_t* threads = malloc(sizeof *threads * nthreads);
	double t0;

	lolb200_stub_format_xrgb8888(&job.fmt);
	t0 = now_ms();
	for (int i = 0; i < nthreads; i++)
		pthread_create(&threads[i], NULL, probe_worker, &job);
	for (int i = 0; i < nthreads; i++)
		pthread_join(threads[i], NULL);
	t0 = now_ms() - t0;
	free(threads);
	return t0;
}

/* One sdf() evaluation of the reference at an arbitrary point (unit tests of the
 * generated distance code). */
void lolref_sdf(void* scene, const float p[3], float* dist, uint32_t* id) {
	struct world_dist d = sdf(scene, (v3){p[0], p[1], p[2]});
	*dist = d.dist;
	*id = d.id;
}
