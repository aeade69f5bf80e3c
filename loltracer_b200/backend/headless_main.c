/*
 * headless_main.c -- main.c's render_scene() without a window.
 *
 * SDL2 is not installed on the build or GPU boxes, so the reference's main.c
 * cannot be built there.  This host replays the same protocol against whichever
 * renderer.h backend it is linked with (naive_renderer.c or b200_renderer.c):
 * threads first, then render_prepare, then per frame {surf, current_line = 0,
 * post N, wait N}, then exiting = 1, post N, join, render_destroy
 * (main.c:139-149,161,182-194,166-170,213).  Frame times use clock_gettime, not
 * SDL_GetTicks' millisecond counter.
 *
 *   lol_headless_<backend> <threads> <scene.lol> [backend flags]
 *        [--size WxH] [--frames N] [--warmup N] [--ppm out.ppm] [--raw out.bin]
 *        [--orbit N]   camera orbit of N frames about (0,1,-6) (BASELINE config C5)
 */
#include <assert.h>
#include <math.h>
#include <pthread.h>
#include <time.h>

#include <SDL.h>

#include "renderer.h"

#define LOG(format, ...) printf("[" __FILE__ ":%d] " format "\n", __LINE__ __VA_OPT__(, ) __VA_ARGS__)

SDL_atomic_t exiting;
SDL_atomic_t current_line;
SDL_sem* frame_entry_barrier;
SDL_sem* frame_exit_barrier;

struct scene* scene_parse(const char* filename); /* scene_parse_shim.c */

static double now_ms(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void* thread_entry(void* p) {
	render_thread(p);
	return NULL;
}

static uint64_t frame_hash(const uint32_t* px, size_t n) {
	uint64_t h = 0;
	for (size_t i = 0; i < n; i++)
		h = h * 1000003u + px[i];
	return h;
}

int main(int argc, const char* argv[]) {
	size_t num_threads = argc > 1 ? (size_t)atoi(argv[1]) : 1;
	const char* filename = argc > 2 ? argv[2] : NULL;
	int width = 320, height = 240; /* main.c:136-137 */
	int frames = 1, warmup = 0, orbit = 0;
	const char* ppm = NULL;
	const char* raw = NULL;
	struct scene* scene;

	for (int i = 3; i < argc; i++) {
		if (!strcmp(argv[i], "--size") && i + 1 < argc)
			sscanf(argv[++i], "%dx%d", &width, &height);
		else if (!strcmp(argv[i], "--frames") && i + 1 < argc)
			frames = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--warmup") && i + 1 < argc)
			warmup = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--orbit") && i + 1 < argc)
			orbit = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--ppm") && i + 1 < argc)
			ppm = argv[++i];
		else if (!strcmp(argv[i], "--raw") && i + 1 < argc)
			raw = argv[++i];
	}
	if (num_threads < 1)
		num_threads = 1;

	scene = scene_parse(filename);
	assert(scene && scene_validate_materials(scene)); /* main.c:234-235 */

	SDL_PixelFormat fmt;
	SDL_Surface surf;
	struct render_data data = {.scene = scene};
	pthread_t* threads = malloc(sizeof *threads * num_threads);
	uint32_t* pixels = calloc((size_t)width * height, sizeof *pixels);
	double tmin = 1e30, tmax = 0, ttotal = 0;
	const struct camera cam0 = scene->camera;

	lolb200_stub_format_xrgb8888(&fmt);
	memset(&surf, 0, sizeof surf);
	surf.format = &fmt;
	surf.w = width;
	surf.h = height;
	surf.pitch = width * 4;
	surf.pixels = pixels;

	LOG("Inicializando threads = %zu", num_threads);
	frame_entry_barrier = SDL_CreateSemaphore(0);
	frame_exit_barrier = SDL_CreateSemaphore(0);
	for (size_t i = 0; i < num_threads; i++)
		pthread_create(&threads[i], NULL, thread_entry, &data);

	render_prepare(&data, argc, argv);

	for (int f = 0; f < warmup + frames; f++) {
		if (orbit > 0) { /* update_camera()'s stand-in: mutate scene->camera in place */
			const double th = 2.0 * M_PI * (f % orbit) / orbit, c = cos(th), s = sin(th);
			const double T[3] = {0.0, 1.0, -6.0};
			double rx = cam0.point.x - T[0], rz = cam0.point.z - T[2];
			scene->camera.point = (v3){(float)(T[0] + c * rx + s * rz), cam0.point.y,
			                           (float)(T[2] - s * rx + c * rz)};
			scene->camera.direction =
				(v3){(float)(c * cam0.direction.x + s * cam0.direction.z), cam0.direction.y,
			         (float)(-s * cam0.direction.x + c * cam0.direction.z)};
			if (f % orbit == 0) {
				scene->camera.point = cam0.point;
				scene->camera.direction = cam0.direction;
			}
		}
		data.surf = &surf;
		SDL_AtomicSet(&current_line, 0);
		double t0 = now_ms();
		for (size_t i = 0; i < num_threads; i++)
			SDL_SemPost(frame_entry_barrier);
		for (size_t i = 0; i < num_threads; i++)
			SDL_SemWait(frame_exit_barrier);
		double dt = now_ms() - t0;
		if (f >= warmup) {
			ttotal += dt;
			if (dt < tmin) tmin = dt;
			if (dt > tmax) tmax = dt;
			LOG("Frame %d\ttime %.3f", f - warmup + 1, dt);
		}
	}
	LOG("min %.3f\tmax %.3f\tavg %.3f ms\t%.3f Mrays/s (best)", tmin, tmax, ttotal / frames,
	    (double)width * height / tmin / 1e3);
	LOG("hash %016llx", (unsigned long long)frame_hash(pixels, (size_t)width * height));

	SDL_AtomicSet(&exiting, 1);
	for (size_t i = 0; i < num_threads; i++)
		SDL_SemPost(frame_entry_barrier);
	for (size_t i = 0; i < num_threads; i++)
		pthread_join(threads[i], NULL);
	render_destroy(&data);
	LOG("Cerrando");

	if (ppm) {
		FILE* f = fopen(ppm, "wb");
		fprintf(f, "P6\n%d %d\n255\n", width, height);
		for (size_t i = 0; i < (size_t)width * height; i++) {
			unsigned char rgb[3] = {pixels[i] >> 16, pixels[i] >> 8, pixels[i]};
			fwrite(rgb, 1, 3, f);
		}
		fclose(f);
	}
	if (raw) {
		FILE* f = fopen(raw, "wb");
		fwrite(pixels, 4, (size_t)width * height, f);
		fclose(f);
	}
	free(threads);
	free(pixels);
	SDL_DestroySemaphore(frame_entry_barrier);
	SDL_DestroySemaphore(frame_exit_barrier);
	scene_free(scene);
	return 0;
}
