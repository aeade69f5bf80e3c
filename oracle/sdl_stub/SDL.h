/*
 * SDL.h -- headless stand-in for the slice of SDL2 the loltracer renderers touch.
 *
 * TEST INFRASTRUCTURE.  SDL2 is not installed in this image; this header lets the
 * reference's naive_renderer.c (and our renderer.h backend) compile unmodified.
 * Field names and SDL_MapRGB's arithmetic follow SDL2's public headers
 * (SDL_pixels.h, SDL_surface.h, SDL_atomic.h, SDL_mutex.h), reduced to
 * non-palettised 32-bit formats.  Semaphores map to POSIX sem_t, atomics to
 * __atomic builtins.  The reference relies on SDL.h for <stdio.h>, <stdbool.h>
 * and <string.h>, so they are pulled in here as SDL_stdinc.h would.
 */
#ifndef LOLB200_SDL_STUB_H
#define LOLB200_SDL_STUB_H

#include <math.h>
#include <semaphore.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define LOLB200_SDL_STUB 1

typedef uint8_t Uint8;
typedef uint16_t Uint16;
typedef uint32_t Uint32;
typedef int32_t Sint32;

typedef struct SDL_atomic_t {
	int value;
} SDL_atomic_t;

typedef struct SDL_semaphore {
	sem_t sem;
} SDL_sem;

typedef struct SDL_PixelFormat {
	Uint32 format;
	void* palette;
	Uint8 BitsPerPixel;
	Uint8 BytesPerPixel;
	Uint8 padding[2];
	Uint32 Rmask, Gmask, Bmask, Amask;
	Uint8 Rloss, Gloss, Bloss, Aloss;
	Uint8 Rshift, Gshift, Bshift, Ashift;
	int refcount;
	struct SDL_PixelFormat* next;
} SDL_PixelFormat;

typedef struct SDL_Surface {
	Uint32 flags;
	SDL_PixelFormat* format;
	int w, h;
	int pitch;
	void* pixels;
} SDL_Surface;

static inline int SDL_AtomicGet(SDL_atomic_t* a) { return __atomic_load_n(&a->value, __ATOMIC_SEQ_CST); }
static inline int SDL_AtomicSet(SDL_atomic_t* a, int v) {
	return __atomic_exchange_n(&a->value, v, __ATOMIC_SEQ_CST);
}
static inline int SDL_AtomicAdd(SDL_atomic_t* a, int v) {
	return __atomic_fetch_add(&a->value, v, __ATOMIC_SEQ_CST);
}

static inline SDL_sem* SDL_CreateSemaphore(Uint32 initial) {
	SDL_sem* s = (SDL_sem*)malloc(sizeof *s);
	sem_init(&s->sem, 0, initial);
	return s;
}
static inline void SDL_DestroySemaphore(SDL_sem* s) {
	sem_destroy(&s->sem);
	free(s);
}
static inline int SDL_SemWait(SDL_sem* s) {
	int rc;
	do
		rc = sem_wait(&s->sem);
	while (rc != 0);
	return 0;
}
static inline int SDL_SemPost(SDL_sem* s) { return sem_post(&s->sem); }

/* SDL_MapRGB for a format without palette (SDL_pixels.c):
 * (r >> Rloss) << Rshift | (g >> Gloss) << Gshift | (b >> Bloss) << Bshift | Amask */
static inline Uint32 SDL_MapRGB(const SDL_PixelFormat* f, Uint8 r, Uint8 g, Uint8 b) {
	return ((Uint32)(r >> f->Rloss) << f->Rshift) | ((Uint32)(g >> f->Gloss) << f->Gshift) |
	       ((Uint32)(b >> f->Bloss) << f->Bshift) | f->Amask;
}

/* XRGB8888 whose mapped pixels carry alpha 0xFF: 0xFF000000 | r<<16 | g<<8 | b. */
static inline void lolb200_stub_format_xrgb8888(SDL_PixelFormat* f) {
	memset(f, 0, sizeof *f);
	f->BitsPerPixel = 32;
	f->BytesPerPixel = 4;
	f->Rmask = 0x00FF0000u;
	f->Gmask = 0x0000FF00u;
	f->Bmask = 0x000000FFu;
	f->Amask = 0xFF000000u;
	f->Rshift = 16;
	f->Gshift = 8;
	f->Bshift = 0;
	f->Ashift = 24;
}

#define SDL_MUSTLOCK(s) 0

#endif
