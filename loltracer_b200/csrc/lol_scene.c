/*
 * lol_scene.c -- gives meaning to a .lol syntax tree: the flat lolb200_scene.
 *
 * Restates, for the product path, what the reference's property extractors do
 * (scene.c:104-281) and the checks main.c:235 / scene.c:284-292 make.  Where
 * the reference aborts (assert / exit(1)) this returns LOLB200_EPARSE.
 *
 * Built with -ffp-contract=off and without -mfma: the few float operations
 * here (camera normalisation, fov conversion, camera basis) must round exactly
 * like the reference's SSE code (vec.h:50-59, scene.c:173-174).
 */
#include "lolb200.h"
#include "lol_ast.h"
#include "lol_internal.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static _Thread_local char g_err[4096];

void lolb200_set_error(const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
}

const char* lolb200_last_error(void) { return g_err; }
int lolb200_abi_version(void) { return LOLB200_ABI_VERSION; }
void lolb200_free(void* p) { free(p); }

/* ---- vec.h restated on plain floats (each operation rounded separately) ---- */

/* v3dot: _mm_dp_ps(a, b, 0x71) = (ax*bx + ay*by) + (az*bz + 0)  (vec.h:50-51) */
static float dot3(const float a[3], const float b[3]) {
	float xx = a[0] * b[0], yy = a[1] * b[1], zz = a[2] * b[2];
	float s = xx + yy;
	return s + zz;
}

/* v3normalize: v * (1 / len)  (vec.h:58-59) */
static void normalize3(float v[3]) {
	float inv = 1.0f / sqrtf(dot3(v, v));
	v[0] *= inv;
	v[1] *= inv;
	v[2] *= inv;
}

/* v3cross (vec.h:68-71) */
static void cross3(const float a[3], const float b[3], float out[3]) {
	float t0 = a[1] * b[2], t1 = a[2] * b[1];
	float t2 = a[2] * b[0], t3 = a[0] * b[2];
	float t4 = a[0] * b[1], t5 = a[1] * b[0];
	out[0] = t0 - t1;
	out[1] = t2 - t3;
	out[2] = t4 - t5;
}

/* get_camera_ray()'s per-frame part (naive_renderer.c:178-188). */
void lolb200_camera_basis_compute(const lolb200_camera* cam, int w, int h,
                                  lolb200_camera_basis* out) {
	const float up_guide[3] = {0.f, 1.f, 0.f};
	float fw = (float)w, fh = (float)h;
	float aspect = fw / fh; /* naive_renderer.c:213 */
	float half_fov = cam->fov / 2.f;
	float height = atanf(half_fov);
	float width = aspect * height;
	float right[3], up[3];

	cross3(cam->direction, up_guide, right);
	normalize3(right);
	cross3(right, cam->direction, up);
	memcpy(out->origin, cam->point, sizeof out->origin);
	memcpy(out->dir, cam->direction, sizeof out->dir);
	memcpy(out->right, right, sizeof right);
	memcpy(out->up, up, sizeof up);
	out->width = width;
	out->height = height;
}

void lolb200_pixfmt_default(lolb200_pixfmt* f) {
	memset(f, 0, sizeof *f);
	f->rshift = 16;
	f->gshift = 8;
	f->bshift = 0;
	f->amask = 0xFF000000u;
}

void lolb200_options_default(lolb200_options* o) {
	memset(o, 0, sizeof *o);
	o->arith = LOLB200_ARITH_EXACT;
	o->skip_black_miss = 1;
	o->cull_backfacing = 1;
	o->shadow_early_out = 1;
	o->guarded_fastpath = 1;
	o->prune_bounds = 1;
	o->roll_phases = 1;
	o->pack_pairs = 1;
	o->share_first_step = 1;
	o->roll_v1 = -1;
	o->loop_worklist = -1;
	o->near_cache = -1;
	o->guard_out = -1;
}

/* ------------------------------------------------------- tree -> flat scene -- */

struct builder {
	lolb200_scene* s;
	size_t cap_nodes, cap_objects, cap_lights;
	int failed;
};

static const char* prop_names[] = {
	"shininess", "diffuse", "specular", "ambient", "color", "point", "direction",
	"fov", "diffuse_intensity", "specular_intensity", "radius", "material", "point2",
	"y", "smoothness", "a", "b"};

static int want_num(struct builder* b, const struct lol_def* d, float* out) {
	if (d->value.kind != LOL_V_NUM) { /* prop_check_num assert, scene.c:68-71 */
		lolb200_set_error("property '%s' must be a number", prop_names[d->prop]);
		return b->failed = 1;
	}
	*out = d->value.num;
	return 0;
}

static int want_v3(struct builder* b, const struct lol_def* d, float out[3]) {
	if (d->value.kind != LOL_V_LIST || d->value.nlist != 3) { /* scene.c:67-74,80-83 */
		lolb200_set_error("property '%s' must be a list of 3 numbers", prop_names[d->prop]);
		return b->failed = 1;
	}
	memcpy(out, d->value.list, 3 * sizeof(float));
	return 0;
}

static int want_id(struct builder* b, const struct lol_def* d, uint32_t* out) {
	if (d->value.kind != LOL_V_ID) { /* scene.c:86-89 */
		lolb200_set_error("property '%s' must be a #id", prop_names[d->prop]);
		return b->failed = 1;
	}
	*out = (uint32_t)d->value.id;
	return 0;
}

static int unknown_prop(struct builder* b, const char* what, const struct lol_def* d) {
	/* SWITCH_END: "Unknown %s property" + exit(1), scene.c:130-134 */
	lolb200_set_error("Unknown %s property '%s'", what, prop_names[d->prop]);
	return b->failed = 1;
}

static int32_t add_node(struct builder* b, const lolb200_object* o) {
	lolb200_scene* s = b->s;
	if (s->n_nodes == b->cap_nodes) {
		b->cap_nodes = b->cap_nodes ? b->cap_nodes * 2 : 32;
		s->nodes = realloc(s->nodes, b->cap_nodes * sizeof *s->nodes);
	}
	s->nodes[s->n_nodes] = *o;
	return (int32_t)s->n_nodes++;
}

/* sphere/box/plane/smooth_union_from_definition_list (scene.c:185-227) and
 * object_from_definition_list (scene.c:266-281).  Children first, so a node's
 * index is always larger than its children's. */
static int32_t build_object(struct builder* b, const struct lol_node* n) {
	lolb200_object o;
	memset(&o, 0, sizeof o);
	o.a = o.b = -1;
	float plane_y = 0.f;

	switch (n->type) {
	case LOL_T_SPHERE: o.type = LOLB200_OBJ_SPHERE; break;
	case LOL_T_BOX: o.type = LOLB200_OBJ_BOX; break;
	case LOL_T_PLANE: o.type = LOLB200_OBJ_PLANE; break;
	case LOL_T_SMOOTH_UNION: o.type = LOLB200_OBJ_SMOOTH_UNION; break;
	case LOL_T_UNION: o.type = LOLB200_OBJ_UNION; break;
	case LOL_T_INTERSECTION: o.type = LOLB200_OBJ_INTERSECTION; break;
	case LOL_T_DIFFERENCE: o.type = LOLB200_OBJ_DIFFERENCE; break;
	default: /* scene.c:277-280 */
		lolb200_set_error("Unknown scene object (only sphere, box, plane and "
		                  "smooth_union can be nested)");
		b->failed = 1;
		return -1;
	}

	for (size_t i = 0; i < n->ndefs && !b->failed; i++) {
		const struct lol_def* d = &n->defs[i];
		int ok = 0;
		switch (d->prop) {
		case LOL_PROP_MATERIAL:
			want_id(b, d, &o.material);
			ok = 1;
			break;
		case LOL_PROP_POINT:
			if (o.type == LOLB200_OBJ_SPHERE || o.type == LOLB200_OBJ_BOX) {
				want_v3(b, d, o.point);
				ok = 1;
			}
			break;
		case LOL_PROP_RADIUS:
			if (o.type == LOLB200_OBJ_SPHERE || o.type == LOLB200_OBJ_BOX) {
				want_num(b, d, &o.radius);
				ok = 1;
			}
			break;
		case LOL_PROP_POINT2:
			if (o.type == LOLB200_OBJ_BOX) {
				want_v3(b, d, o.point2);
				ok = 1;
			}
			break;
		case LOL_PROP_Y:
			if (o.type == LOLB200_OBJ_PLANE) {
				want_num(b, d, &plane_y);
				ok = 1;
			}
			break;
		case LOL_PROP_SMOOTHNESS:
			if (o.type == LOLB200_OBJ_SMOOTH_UNION) {
				want_num(b, d, &o.smoothness);
				ok = 1;
			}
			break;
		case LOL_PROP_A:
		case LOL_PROP_B:
			if (LOLB200_OBJ_HAS_CHILDREN(o.type)) {
				ok = 1;
				if (d->value.kind != LOL_V_OBJ) { /* scene.c:92-97 */
					lolb200_set_error("property '%s' must be an object",
					                  prop_names[d->prop]);
					b->failed = 1;
					break;
				}
				int32_t child = build_object(b, d->value.obj);
				if (d->prop == LOL_PROP_A)
					o.a = child; /* a repeated property: the last one wins */
				else
					o.b = child;
			}
			break;
		default: break;
		}
		if (!ok && !b->failed)
			unknown_prop(b, "object", d);
	}
	if (b->failed)
		return -1;
	if (o.type == LOLB200_OBJ_PLANE) { /* scene.c:215 */
		o.point[0] = 0.f;
		o.point[1] = plane_y;
		o.point[2] = 0.f;
	}
	if (LOLB200_OBJ_HAS_CHILDREN(o.type) && (o.a < 0 || o.b < 0)) {
		/* the reference would dereference a NULL child at render time */
		lolb200_set_error("%s needs both 'a' and 'b'",
		                  o.type == LOLB200_OBJ_SMOOTH_UNION ? "smooth_union" : "a CSG node");
		b->failed = 1;
		return -1;
	}
	return add_node(b, &o);
}

/* material_from_definition_list (scene.c:140-147) */
static void build_material(struct builder* b, const struct lol_node* n, lolb200_material* m) {
	memset(m, 0, sizeof *m);
	for (size_t i = 0; i < n->ndefs && !b->failed; i++) {
		const struct lol_def* d = &n->defs[i];
		switch (d->prop) {
		case LOL_PROP_SHININESS: want_num(b, d, &m->shininess); break;
		case LOL_PROP_DIFFUSE: want_v3(b, d, m->diffuse); break;
		case LOL_PROP_SPECULAR: want_v3(b, d, m->specular); break;
		case LOL_PROP_AMBIENT: want_v3(b, d, m->ambient); break;
		default: unknown_prop(b, "material", d);
		}
	}
}

static void build_component(struct builder* b, const struct lol_node* n) {
	lolb200_scene* s = b->s;
	switch (n->type) {
	case LOL_T_AMBIENT: /* ambient_from_definition_list, scene.c:149-165 */
		for (size_t i = 0; i < n->ndefs && !b->failed; i++) {
			if (n->defs[i].prop == LOL_PROP_COLOR)
				want_v3(b, &n->defs[i], s->ambient_color);
			else
				unknown_prop(b, "ambient", &n->defs[i]);
		}
		break;
	case LOL_T_CAMERA: { /* camera_from_definition_list, scene.c:167-175 */
		lolb200_camera c;
		memset(&c, 0, sizeof c);
		for (size_t i = 0; i < n->ndefs && !b->failed; i++) {
			const struct lol_def* d = &n->defs[i];
			switch (d->prop) {
			case LOL_PROP_POINT: want_v3(b, d, c.point); break;
			case LOL_PROP_DIRECTION: want_v3(b, d, c.direction); break;
			case LOL_PROP_FOV: want_num(b, d, &c.fov); break;
			default: unknown_prop(b, "camera", d);
			}
		}
		normalize3(c.direction);                        /* scene.c:173 */
		c.fov = (float)((double)(c.fov / 180) * M_PI);  /* scene.c:174 */
		s->camera = c;
		break;
	}
	case LOL_T_POINT_LIGHT: { /* light_from_definition_list, scene.c:177-183 */
		lolb200_light l;
		memset(&l, 0, sizeof l);
		for (size_t i = 0; i < n->ndefs && !b->failed; i++) {
			const struct lol_def* d = &n->defs[i];
			switch (d->prop) {
			case LOL_PROP_POINT: want_v3(b, d, l.point); break;
			case LOL_PROP_DIFFUSE_INTENSITY: want_v3(b, d, l.diffuse_intensity); break;
			case LOL_PROP_SPECULAR_INTENSITY: want_v3(b, d, l.specular_intensity); break;
			default: unknown_prop(b, "light", d);
			}
		}
		if (s->n_lights == b->cap_lights) {
			b->cap_lights = b->cap_lights ? b->cap_lights * 2 : 16;
			s->lights = realloc(s->lights, b->cap_lights * sizeof *s->lights);
		}
		s->lights[s->n_lights++] = l;
		break;
	}
	default: {
		int32_t idx = build_object(b, n);
		if (b->failed)
			return;
		if (s->n_objects == b->cap_objects) {
			b->cap_objects = b->cap_objects ? b->cap_objects * 2 : 16;
			s->objects = realloc(s->objects, b->cap_objects * sizeof *s->objects);
		}
		s->objects[s->n_objects++] = (uint32_t)idx;
	}
	}
}

int lolb200_scene_parse_string(const char* text, size_t len, lolb200_scene** out) {
	char err[512];
	struct lol_doc* doc;
	struct builder b;

	if (!text || !out) {
		lolb200_set_error("lolb200_scene_parse_string: NULL argument");
		return LOLB200_EINVAL;
	}
	*out = NULL;
	doc = lol_parse_text(text, len, err, sizeof err);
	if (!doc) {
		lolb200_set_error("%s", err);
		return LOLB200_EPARSE;
	}

	memset(&b, 0, sizeof b);
	b.s = calloc(1, sizeof *b.s);
	/* scene_new() defaults (scene.c:44-58) */
	b.s->camera.direction[2] = 1.f;
	b.s->camera.fov = (float)(M_PI / 2);

	b.s->n_materials = (uint32_t)doc->nmaterials;
	b.s->materials = calloc(doc->nmaterials ? doc->nmaterials : 1, sizeof *b.s->materials);
	for (size_t i = 0; i < doc->nmaterials && !b.failed; i++)
		build_material(&b, &doc->materials[i], &b.s->materials[i]);
	for (size_t i = 0; i < doc->ncomponents && !b.failed; i++)
		build_component(&b, &doc->components[i]);
	lol_doc_free(doc);

	/* scene_validate_materials (scene.c:284-292): top-level objects only */
	for (uint32_t i = 0; i < b.s->n_objects && !b.failed; i++)
		if (b.s->nodes[b.s->objects[i]].material >= b.s->n_materials) {
			lolb200_set_error("object %u uses material #%u but only %u are defined", i + 1,
			                  b.s->nodes[b.s->objects[i]].material, b.s->n_materials);
			b.failed = 1;
		}
	if (b.failed) {
		lolb200_scene_free(b.s);
		return LOLB200_EPARSE;
	}
	*out = b.s;
	return LOLB200_OK;
}

int lolb200_scene_parse_file(const char* path, lolb200_scene** out) {
	FILE* f;
	char* buf;
	long n;
	int rc;

	if (!path || !out) {
		lolb200_set_error("lolb200_scene_parse_file: NULL argument");
		return LOLB200_EINVAL;
	}
	f = fopen(path, "rb");
	if (!f) { /* scene_parse returns NULL (scene-parser.y:205-206) */
		lolb200_set_error("cannot open '%s'", path);
		return LOLB200_EPARSE;
	}
	fseek(f, 0, SEEK_END);
	n = ftell(f);
	fseek(f, 0, SEEK_SET);
	buf = malloc((size_t)n + 1);
	if (fread(buf, 1, (size_t)n, f) != (size_t)n) {
		fclose(f);
		free(buf);
		lolb200_set_error("cannot read '%s'", path);
		return LOLB200_EPARSE;
	}
	fclose(f);
	buf[n] = 0;
	rc = lolb200_scene_parse_string(buf, (size_t)n, out);
	free(buf);
	return rc;
}

static void* dup_mem(const void* p, size_t n) {
	void* q = malloc(n ? n : 1);
	if (n)
		memcpy(q, p, n);
	return q;
}

lolb200_scene* lolb200_scene_clone(const lolb200_scene* s) {
	lolb200_scene* c;
	if (!s)
		return NULL;
	c = malloc(sizeof *c);
	*c = *s;
	c->materials = dup_mem(s->materials, s->n_materials * sizeof *s->materials);
	c->lights = dup_mem(s->lights, s->n_lights * sizeof *s->lights);
	c->nodes = dup_mem(s->nodes, s->n_nodes * sizeof *s->nodes);
	c->objects = dup_mem(s->objects, s->n_objects * sizeof *s->objects);
	return c;
}

void lolb200_scene_free(lolb200_scene* s) {
	if (!s)
		return;
	free(s->materials);
	free(s->lights);
	free(s->nodes);
	free(s->objects);
	free(s);
}

/* Basic structural validation shared by the lowering and the device layer:
 * hand-built scenes (the renderer.h backend translates the reference's structs)
 * come through here too. */
int lolb200_scene_check(const lolb200_scene* s) {
	if (!s) {
		lolb200_set_error("NULL scene");
		return LOLB200_EINVAL;
	}
	if (s->n_materials == 0) {
		lolb200_set_error("scene has no materials (material 0 shades misses)");
		return LOLB200_EINVAL;
	}
	for (uint32_t i = 0; i < s->n_nodes; i++) {
		const lolb200_object* o = &s->nodes[i];
		switch (o->type) {
		case LOLB200_OBJ_SPHERE:
		case LOLB200_OBJ_BOX:
		case LOLB200_OBJ_PLANE: break;
		case LOLB200_OBJ_SMOOTH_UNION:
		case LOLB200_OBJ_UNION:
		case LOLB200_OBJ_INTERSECTION:
		case LOLB200_OBJ_DIFFERENCE:
			if (o->a < 0 || o->b < 0 || (uint32_t)o->a >= i || (uint32_t)o->b >= i) {
				lolb200_set_error("node %u: children must precede their parent", i);
				return LOLB200_EINVAL;
			}
			break;
		default:
			lolb200_set_error("node %u: unknown object type %d", i, o->type);
			return LOLB200_EINVAL;
		}
	}
	if (s->n_nodes) {
		/* children precede their parents, so one forward pass gives every node's height */
		uint32_t* height = calloc(s->n_nodes, sizeof *height);
		int too_deep = 0;
		for (uint32_t i = 0; i < s->n_nodes && height; i++) {
			const lolb200_object* o = &s->nodes[i];
			height[i] = 1;
			if (LOLB200_OBJ_HAS_CHILDREN(o->type)) {
				const uint32_t ha = height[o->a], hb = height[o->b];
				height[i] = 1 + (ha > hb ? ha : hb);
			}
			too_deep |= height[i] > LOLB200_MAX_NESTING;
		}
		free(height);
		if (too_deep) {
			lolb200_set_error("objects nested deeper than %d levels", LOLB200_MAX_NESTING);
			return LOLB200_EINVAL;
		}
	}
	for (uint32_t i = 0; i < s->n_objects; i++) {
		if (s->objects[i] >= s->n_nodes) {
			lolb200_set_error("object %u: node index out of range", i + 1);
			return LOLB200_EINVAL;
		}
		if (s->nodes[s->objects[i]].material >= s->n_materials) {
			lolb200_set_error("object %u: material out of range", i + 1);
			return LOLB200_EINVAL;
		}
	}
	return LOLB200_OK;
}

static uint64_t node_flops(const lolb200_scene* s, uint32_t i) {
	const lolb200_object* o = &s->nodes[i];
	switch (o->type) {
	case LOLB200_OBJ_SPHERE: return 10;
	case LOLB200_OBJ_BOX: return 20;
	case LOLB200_OBJ_PLANE: return 1;
	case LOLB200_OBJ_SMOOTH_UNION:
		return 13 + node_flops(s, (uint32_t)o->a) + node_flops(s, (uint32_t)o->b);
	case LOLB200_OBJ_UNION:
	case LOLB200_OBJ_INTERSECTION:
	case LOLB200_OBJ_DIFFERENCE: /* one min/max; the negation is free */
		return 1 + node_flops(s, (uint32_t)o->a) + node_flops(s, (uint32_t)o->b);
	}
	return 0;
}

uint64_t lolb200_scene_flops_per_eval(const lolb200_scene* s) {
	uint64_t f = 0;
	if (lolb200_scene_check(s) != LOLB200_OK)
		return 0;
	for (uint32_t i = 0; i < s->n_objects; i++)
		f += node_flops(s, s->objects[i]) + 1;
	return f;
}
