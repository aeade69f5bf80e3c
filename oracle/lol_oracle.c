/*
 * lol_oracle.c -- CPU restatement of loltracer's per-pixel sphere-tracing path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product path never
 * does, and fails loudly without its CUDA library.
 *
 * Parity status: PINNED.  tests/test_oracle_pin.py checks this file bit for bit
 * (distance, object id and packed pixel of every pixel) against the reference's
 * own naive_renderer.c compiled unmodified (oracle/_ref/liblolref.so, built by
 * oracle/Makefile from /root/reference), and against the frame hashes committed
 * in tests/golden/.  The reference ships no tests or golden vectors of its own.
 *
 * Plain scalar C, one rounding per operation (-ffp-contract=off, no -mfma), in
 * the operation order of the reference's SSE code.  mode 0 = naive_renderer.c
 * semantics; mode 1 = the deltas of tracing_jit_renderer.dasc (see sdf_jit);
 * mode 2 = the pipeline of mode 0 around a per-scene SPECIALISED distance
 * function handed in by the caller (lolo_set_specialised_sdf): the CPU timing
 * stand-in for the DynASM JIT, which cannot be built here (no Lua).
 * mode | 0x100 = per-child materials, an extension of ours (child_material below;
 * parity for it UNPINNED: the reference has no such feature).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "lolb200.h"

typedef struct { float x, y, z; } V3;

/* float.h:6-14 -- MAXSS/MINSS: the second operand comes back when unordered. */
static inline float maxf_(float a, float b) { return a > b ? a : b; }
static inline float minf_(float a, float b) { return a < b ? a : b; }
/* float.h:16-22 */
static inline float clamp_(float v, float lo, float hi) { return minf_(maxf_(v, lo), hi); }

/* vec.h:42-49 */
static inline V3 add(V3 a, V3 b) { return (V3){a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 sub(V3 a, V3 b) { return (V3){a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 mul(V3 a, V3 b) { return (V3){a.x * b.x, a.y * b.y, a.z * b.z}; }
/* vec.h:56-57 */
static inline V3 scale(V3 v, float f) { return (V3){v.x * f, v.y * f, v.z * f}; }
/* vec.h:50-51: dpps 0x71 = (x*x' + y*y') + (z*z' + 0) */
static inline float dot(V3 a, V3 b) {
	float xx = a.x * b.x, yy = a.y * b.y, zz = a.z * b.z;
	float s = xx + yy;
	return s + zz;
}
/* vec.h:52-53 */
static inline float len(V3 a) { return sqrtf(dot(a, a)); }
/* vec.h:58-59: reciprocal first, then three multiplies */
static inline V3 normalize(V3 v) { return scale(v, 1.0f / len(v)); }
static inline V3 from3(const float p[3]) { return (V3){p[0], p[1], p[2]}; }

struct world_dist { float dist; uint32_t id; };

/* float.h:24-33 */
static inline float sminf_(float a, float b, float k) {
	float h = clamp_(.5f + .5f * (b - a) / k, 0.f, 1.f);
	float l = b + (a - b) * h; /* lerp(b, a, h) */
	return l - k * h * (1.f - h);
}

/* Extension nodes (lolb200.h: union / intersection / difference).  The reference
 * has no such nodes: this is OUR definition, restated once here and once in the
 * lowering -- parity for them is against this file only ("unpinned"). */
static float csg_dist(const lolb200_scene* s, const lolb200_object* o, V3 p,
                      float (*rec)(const lolb200_scene*, const lolb200_object*, V3)) {
	float a = rec(s, &s->nodes[o->a], p);
	float b = rec(s, &s->nodes[o->b], p);
	switch (o->type) {
	case LOLB200_OBJ_UNION: return minf_(a, b);
	case LOLB200_OBJ_INTERSECTION: return maxf_(a, b);
	default: return maxf_(a, -b); /* difference */
	}
}

/* get_obj_dist (naive_renderer.c:10-28) with sdSphere / sdRoundBox (sdf.h:8-22).
 * Children of a smooth union see p, not p - point. */
static float obj_dist(const lolb200_scene* s, const lolb200_object* o, V3 p) {
	V3 q = sub(p, from3(o->point));
	switch (o->type) {
	case LOLB200_OBJ_SPHERE: return len(q) - o->radius;
	case LOLB200_OBJ_BOX: {
		V3 d = {fabsf(q.x) - o->point2[0], fabsf(q.y) - o->point2[1], fabsf(q.z) - o->point2[2]};
		V3 c = {maxf_(d.x, 0.f), maxf_(d.y, 0.f), maxf_(d.z, 0.f)};
		return len(c) + minf_(maxf_(d.x, maxf_(d.y, d.z)), 0.f) - o->radius;
	}
	case LOLB200_OBJ_PLANE: return q.y;
	case LOLB200_OBJ_UNION:
	case LOLB200_OBJ_INTERSECTION:
	case LOLB200_OBJ_DIFFERENCE: return csg_dist(s, o, p, obj_dist);
	default: {
		float a = obj_dist(s, &s->nodes[o->a], p);
		float b = obj_dist(s, &s->nodes[o->b], p);
		return sminf_(a, b, o->smoothness);
	}
	}
}

/* sdf (naive_renderer.c:30-44): strict < keeps the first of equal distances. */
static struct world_dist sdf_naive(const lolb200_scene* s, V3 p) {
	struct world_dist r = {INFINITY, 0};
	for (uint32_t i = 0; i < s->n_objects; i++) {
		float d = obj_dist(s, &s->nodes[s->objects[i]], p);
		if (d < r.dist)
			r = (struct world_dist){d, i + 1};
	}
	return r;
}

/* generate_obj_dist (tracing_jit_renderer.dasc:148-216): boxes are +INF
 * (:168-174); sminf is max(min(h,1),0) (:197-198) and its tail (h*(1-h))*k
 * (:204-208). */
static float obj_dist_jit(const lolb200_scene* s, const lolb200_object* o, V3 p) {
	V3 q = sub(p, from3(o->point));
	switch (o->type) {
	case LOLB200_OBJ_SPHERE: return len(q) - o->radius;
	case LOLB200_OBJ_BOX: return INFINITY;
	case LOLB200_OBJ_PLANE: return q.y;
	case LOLB200_OBJ_UNION:
	case LOLB200_OBJ_INTERSECTION:
	case LOLB200_OBJ_DIFFERENCE: return csg_dist(s, o, p, obj_dist_jit);
	default: {
		float a = obj_dist_jit(s, &s->nodes[o->a], p);
		float b = obj_dist_jit(s, &s->nodes[o->b], p);
		float k = o->smoothness;
		float h = ((b - a) * .5f) / k + .5f;
		h = maxf_(minf_(h, 1.f), 0.f);
		float r = b + (a - b) * h;
		return r - (h * (1.f - h)) * k;
	}
	}
}

/* sdf_main (tracing_jit_renderer.dasc:113-133): cmpps LE, so later objects win ties. */
static struct world_dist sdf_jit(const lolb200_scene* s, V3 p) {
	struct world_dist r = {INFINITY, 0};
	for (uint32_t i = 0; i < s->n_objects; i++) {
		float d = obj_dist_jit(s, &s->nodes[s->objects[i]], p);
		if (d <= r.dist)
			r = (struct world_dist){d, i + 1};
	}
	return r;
}

/* EXTENSION (lolb200_options.child_materials).  The reference ignores the materials of a
 * composite's children (naive_renderer.c:102-112): this is OUR definition, restated here and in the lowering
 * (lol_lower.c: emit_mat_node) -- parity for it is against this file only ("unpinned").
 * The material of a hit on a composite object is the material of the child that decides the node's distance
 * at the hit point: the nearer child of a (smooth) union, the farther one of an intersection, a -- or b
 * where b carves -- of a difference; ties keep a.  A node whose material is #0 (the field's default after
 * scene.c:124's memset) inherits its parent's.  Distances are the naive renderer's (obj_dist). */
static float child_material(const lolb200_scene* s, const lolb200_object* o, V3 p, uint32_t inherited,
                            uint32_t* mat) {
	const uint32_t eff = o->material ? o->material : inherited;
	uint32_t ma, mb;
	float a, b;
	if (!LOLB200_OBJ_HAS_CHILDREN(o->type)) {
		*mat = eff;
		return obj_dist(s, o, p);
	}
	a = child_material(s, &s->nodes[o->a], p, eff, &ma);
	b = child_material(s, &s->nodes[o->b], p, eff, &mb);
	switch (o->type) {
	case LOLB200_OBJ_UNION: *mat = (b < a) ? mb : ma; return minf_(a, b);
	case LOLB200_OBJ_INTERSECTION: *mat = (b > a) ? mb : ma; return maxf_(a, b);
	case LOLB200_OBJ_DIFFERENCE: *mat = (-b > a) ? mb : ma; return maxf_(a, -b);
	default: *mat = (b < a) ? mb : ma; return sminf_(a, b, o->smoothness);
	}
}

#define LOLO_MODE_CHILD_MATERIALS 0x100 /* OR-ed into `mode` */

struct ctx {
	const lolb200_scene* s;
	int mode;
	int child_materials;
	uint32_t n_primary, n_normal, n_shadow;
};

/* mode 2: straight-line code with baked constants, generated by the same lowering
 * that feeds NVRTC and compiled by gcc (tests/oracle_lib.py: specialised_sdf) --
 * what generate_sdf() (tracing_jit_renderer.dasc:76-216) does with DynASM. */
typedef float (*lolo_spec_fn)(float x, float y, float z, uint32_t* id);
static lolo_spec_fn g_spec;
void lolo_set_specialised_sdf(lolo_spec_fn fn) { g_spec = fn; }

static inline struct world_dist sdf(struct ctx* c, V3 p) {
	if (c->mode == 2) {
		struct world_dist r;
		r.dist = g_spec(p.x, p.y, p.z, &r.id);
		return r;
	}
	return c->mode == 1 ? sdf_jit(c->s, p) : sdf_naive(c->s, p);
}

/* get_intersection (naive_renderer.c:47-69) */
static struct world_dist intersect(struct ctx* c, V3 ro, V3 rd) {
	uint32_t id = 0;
	float dist = 0.f;
	for (int i = 0; i < 256; i++) {
		V3 p = add(ro, scale(rd, dist));
		struct world_dist d = sdf(c, p);
		c->n_primary++;
		dist += d.dist;
		id = d.id;
		if (d.dist < 0.001f || dist > 100.f)
			break;
	}
	if (dist >= 100.f)
		id = 0;
	return (struct world_dist){dist, id};
}

/* softshadow (naive_renderer.c:72-90): no epsilon exit; first step divides by 0.
 * JIT copy uses libm fminf/fmaxf (tracing_jit_renderer.dasc:256,261). */
static float softshadow(struct ctx* c, V3 ro, V3 rd, int max_steps, float max_dist, float w) {
	float res = 1.f, dist = 0.f;
	for (int i = 0; i < max_steps; i++) {
		V3 p = add(ro, scale(rd, dist));
		float d = sdf(c, p).dist;
		c->n_shadow++;
		res = c->mode == 1 ? fminf(res, w * d / dist) : minf_(res, w * d / dist);
		dist += d;
		if (res < -1 || dist > max_dist)
			break;
	}
	return c->mode == 1 ? fmaxf(res, 0.f) : maxf_(res, 0.f);
}

/* in_shadow (naive_renderer.c:92-100) */
static float in_shadow(struct ctx* c, const lolb200_light* l, V3 p) {
	V3 lp = sub(from3(l->point), p);
	float light_dist = len(lp);
	V3 dir = normalize(lp);
	return softshadow(c, add(p, dir), dir, 128, light_dist, 50.f);
}

/* get_normal (naive_renderer.c:114-125) */
static V3 get_normal(struct ctx* c, V3 p, float dist) {
	static const V3 k[4] = {{1.f, -1.f, -1.f}, {-1.f, -1.f, 1.f}, {-1.f, 1.f, -1.f}, {1.f, 1.f, 1.f}};
	const float h = dist / 100.f;
	V3 t[4];
	for (int i = 0; i < 4; i++) {
		t[i] = scale(k[i], sdf(c, add(p, scale(k[i], h))).dist);
		c->n_normal++;
	}
	return normalize(add(t[0], add(t[1], add(t[2], t[3]))));
}

/* get_light (naive_renderer.c:128-175) with get_material (:102-112) */
static V3 get_light(struct ctx* c, V3 p, V3 n, uint32_t id) {
	const lolb200_scene* s = c->s;
	uint32_t mi = id ? s->nodes[s->objects[id - 1]].material : 0;
	if (id && c->child_materials) {
		const lolb200_object* top = &s->nodes[s->objects[id - 1]];
		child_material(s, top, p, top->material, &mi);
	}
	const lolb200_material* mat = &s->materials[mi];
	V3 total = {0.f, 0.f, 0.f};
	V3 cam_pos = from3(s->camera.point);

	for (uint32_t i = 0; i < s->n_lights; i++) {
		const lolb200_light* l = &s->lights[i];
		float shadow = in_shadow(c, l, p);
		V3 light_dir = normalize(sub(from3(l->point), p));
		V3 reflected = sub(scale(n, 2.f * dot(light_dir, n)), light_dir);
		V3 camera_dir = normalize(sub(cam_pos, p));
		float diffuse_incidence = clamp_(dot(n, light_dir), 0.f, 1.f);
		V3 ld = scale(from3(l->diffuse_intensity), shadow * diffuse_incidence);
		ld = mul(ld, from3(mat->diffuse));
		total = add(total, ld);
		float specular_incidence =
			diffuse_incidence * powf(clamp_(dot(reflected, camera_dir), 0.f, 1.f), mat->shininess);
		V3 ls = scale(from3(l->specular_intensity), shadow * specular_incidence);
		ls = mul(ls, from3(mat->specular));
		total = add(total, ls);
	}
	total = add(total, mul(from3(s->ambient_color), from3(mat->ambient)));
	/* v3clamp (vec.h:63-65): max_ps(min_ps(v, 1), 0) */
	total.x = maxf_(minf_(total.x, 1.f), 0.f);
	total.y = maxf_(minf_(total.y, 1.f), 0.f);
	total.z = maxf_(minf_(total.z, 1.f), 0.f);
	return total;
}

struct job {
	const lolb200_scene* s;
	lolb200_camera_basis cb;
	int mode, w, h, y0, y1, ystride;
	float* dist;
	uint32_t* id;
	uint32_t* rgba;
	uint16_t* nprimary;
	uint16_t* nshadow;
	int next;
	uint64_t tot_primary, tot_normal, tot_shadow, tot_hits;
	pthread_mutex_t mu;
};

/* render_thread's pixel loop (naive_renderer.c:216-236); the camera basis is
 * hoisted (get_camera_ray recomputes the same values per pixel, :178-188). */
static void* worker(void* arg) {
	struct job* j = arg;
	struct ctx c = {.s = j->s, .mode = j->mode & 0xff, .child_materials = (j->mode & LOLO_MODE_CHILD_MATERIALS) != 0};
	float fwidth = j->w, fheight = j->h;
	V3 ro = from3(j->cb.origin), dir = from3(j->cb.dir);
	V3 right = from3(j->cb.right), up = from3(j->cb.up);
	int nrows = (j->y1 - j->y0 + j->ystride - 1) / j->ystride;
	uint64_t hits = 0;
	int r;

	while ((r = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED)) < nrows) {
		int y = j->y0 + r * j->ystride;
		for (int x = 0; x < j->w; x++) {
			float vx = (x + .5f) / fwidth * 2.f - 1.f;
			float vy = 1.f - (y + .5f) / fheight * 2.f;
			V3 rd = add(scale(right, vx * j->cb.width), scale(up, vy * j->cb.height));
			rd = normalize(add(rd, dir));
			uint32_t p0 = c.n_primary, s0 = c.n_shadow;
			struct world_dist hit = intersect(&c, ro, rd);
			V3 p = add(ro, scale(rd, hit.dist));
			V3 n = get_normal(&c, p, hit.dist);
			V3 col = get_light(&c, p, n, hit.id);
			col = (V3){powf(col.x, 1.f / 2.2f), powf(col.y, 1.f / 2.2f), powf(col.z, 1.f / 2.2f)};
			/* colorf_to_pixfmt (renderer.h:17-22) + SDL_MapRGB, XRGB8888 */
			uint8_t cr = col.x * 255, cg = col.y * 255, cb = col.z * 255;
			size_t o = (size_t)r * j->w + x;
			hits += hit.id != 0;
			if (j->dist) j->dist[o] = hit.dist;
			if (j->id) j->id[o] = hit.id;
			if (j->rgba)
				j->rgba[o] = 0xFF000000u | ((uint32_t)cr << 16) | ((uint32_t)cg << 8) | cb;
			if (j->nprimary) j->nprimary[o] = (uint16_t)(c.n_primary - p0);
			if (j->nshadow) j->nshadow[o] = (uint16_t)(c.n_shadow - s0);
		}
	}
	pthread_mutex_lock(&j->mu);
	j->tot_primary += c.n_primary;
	j->tot_normal += c.n_normal;
	j->tot_shadow += c.n_shadow;
	j->tot_hits += hits;
	pthread_mutex_unlock(&j->mu);
	return NULL;
}

static double now_ms(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* Rows y0, y0+ystride, ... < y1 of the w x h frame; buffers are compact over
 * the sampled rows; any of them may be NULL.  cam NULL = the scene's camera.
 * totals (may be NULL): [0] primary, [1] normal, [2] shadow evals, [3] hit
 * pixels.  Returns the elapsed wall time in ms. */
double lolo_render(const lolb200_scene* s, const lolb200_camera* cam, int mode, int w, int h,
                   int y0, int y1, int ystride, int nthreads, float* dist, uint32_t* id,
                   uint32_t* rgba, uint16_t* nprimary, uint16_t* nshadow, uint64_t totals[4]) {
	lolb200_scene local = *s;
	struct job j = {.mode = mode, .w = w, .h = h, .y0 = y0, .y1 = y1, .ystride = ystride,
	                .dist = dist, .id = id, .rgba = rgba, .nprimary = nprimary, .nshadow = nshadow};
	pthread_t* th;
	double t0;

	if (cam)
		local.camera = *cam; /* get_light reads scene->camera.point (naive_renderer.c:131) */
	j.s = &local;
	lolb200_camera_basis_compute(&local.camera, w, h, &j.cb);
	pthread_mutex_init(&j.mu, NULL);
	if (nthreads < 1)
		nthreads = 1;
	th = malloc(sizeof *th * nthreads);
	t0 = now_ms();
	for (int i = 0; i < nthreads; i++)
		pthread_create(&th[i], NULL, worker, &j);
	for (int i = 0; i < nthreads; i++)
		pthread_join(th[i], NULL);
	t0 = now_ms() - t0;
	free(th);
	if (totals) {
		totals[0] = j.tot_primary;
		totals[1] = j.tot_normal;
		totals[2] = j.tot_shadow;
		totals[3] = j.tot_hits;
	}
	return t0;
}

void lolo_sdf(const lolb200_scene* s, int mode, const float p[3], float* dist, uint32_t* id) {
	struct ctx c = {.s = s, .mode = mode};
	struct world_dist d = sdf(&c, from3(p));
	*dist = d.dist;
	*id = d.id;
}
