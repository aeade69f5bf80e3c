// lol_params.h -- kernel argument block shared by the host layer (lol_cuda.cu
// includes it) and the generated kernels (lol_lower.c splices its text into
// every program it emits), so the two can never drift apart.
// Plain C types only: it is compiled by gcc/nvcc and by NVRTC (no headers).
#ifndef LOL_PARAMS_H
#define LOL_PARAMS_H

typedef unsigned int lol_u32;
typedef unsigned short lol_u16;
typedef unsigned long long lol_u64;

// Everything that may change between frames.  Scene constants never appear
// here: they are immediates / tables of the generated code.
struct lol_params {
	float ox, oy, oz;    // camera.point
	float dx, dy, dz;    // camera.direction
	float rx, ry, rz;    // right = normalize(cross(dir, (0,1,0)))
	float ux, uy, uz;    // up = cross(right, dir)
	float cw, ch;        // aspect*atanf(fov/2), atanf(fov/2)
	float fw, fh;        // (float)w, (float)h
	int w, h;
	int rank, world;     // band b (4 rows) belongs to rank b % world
	int dst_full;        // 1: dst indexed as a full frame, 0: compact local bands
	lol_u32 pitch;       // dst row pitch in pixels
	lol_u32 chunk_w;     // pixels per work chunk along x (multiple of 8)
	lol_u32 chunks_per_band;
	lol_u32 n_chunks;
	lol_u32 band_begin;  // first local band of this launch (slabs of one frame)
	lol_u32 rshift, gshift, bshift, rloss, gloss, bloss, amask;
	lol_u32* counter;    // [0] next chunk, [1] finished CTAs
	lol_u32* dst;
	float* aux_dist;     // optional per-pixel probes (full-frame, pitch = w)
	lol_u32* aux_id;
	lol_u16* aux_primary;
	lol_u16* aux_shadow;
	lol_u64* stats;      // LOL_COUNTERS: 8 accumulators
	const lol_u32* order; // optional: n-th pull from the queue -> chunk (longest first)
	lol_u32* cost;        // optional: clocks each chunk took, by chunk
	lol_u64* timing;      // optional probes: [0] queue dry, [1] last exit, [2] first start (global timer, ns)
	lol_u32* done_flag;   // optional: word (any GPU's memory) that receives done_value when the launch is complete
	lol_u32 done_value;
	// variant 4 (deferred long rays): a march that is not over after cap_* evaluations puts its pixel aside
	lol_u32 cap_primary, cap_shadow;
	lol_u32 q_cap;        // slots per continuation queue (one queue per class: lol_kernel.cuh)
	lol_u32* q;           // records of 16 words, queue c at q + c * q_cap * 16
	lol_u32* q_ctl;       // per class [2c] records pushed, [2c + 1] next to resume; then finished CTAs of lol_resume
};

#define LOL_BAND_ROWS 4

#endif
