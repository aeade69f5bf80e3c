/*
 * scene_translate.h -- reference `struct scene` (scene.h:90-96) -> lolb200_scene.
 *
 * Compiled only inside translation units that already see the reference's
 * scene.h (the renderer.h backend, and the oracle harness that checks the
 * translation against our own front-end).  Plain copies: no value is
 * recomputed, so the flat scene carries the reference's bits.
 */
#ifndef LOLB200_SCENE_TRANSLATE_H
#define LOLB200_SCENE_TRANSLATE_H

#include <stdlib.h>
#include <string.h>

#include "lolb200.h"
#include "scene.h" /* the reference's */

static inline void lolb200__v3_out(v3 v, float out[3]) {
	out[0] = v.x;
	out[1] = v.y;
	out[2] = v.z;
}

static size_t lolb200__count_nodes(const struct object* o) {
	if (o->type == OBJ_SMOOTH_UNION)
		return 1 + lolb200__count_nodes(o->smooth_op.a) + lolb200__count_nodes(o->smooth_op.b);
	return 1;
}

/* Children first (same order as our front-end: a, b, then the node). */
static int32_t lolb200__flatten(const struct object* o, lolb200_scene* s) {
	lolb200_object n;
	memset(&n, 0, sizeof n);
	n.type = (int32_t)o->type;
	n.material = (uint32_t)o->material;
	n.a = n.b = -1;
	lolb200__v3_out(o->point, n.point);
	switch (o->type) {
	case OBJ_SPHERE: n.radius = o->sphere.radius; break;
	case OBJ_BOX:
		n.radius = o->box.radius;
		lolb200__v3_out(o->box.point2, n.point2);
		break;
	case OBJ_PLANE: break;
	case OBJ_SMOOTH_UNION:
		n.smoothness = o->smooth_op.smoothness;
		n.a = lolb200__flatten(o->smooth_op.a, s);
		n.b = lolb200__flatten(o->smooth_op.b, s);
		break;
	default: break;
	}
	s->nodes[s->n_nodes] = n;
	return (int32_t)s->n_nodes++;
}

static inline void lolb200__camera_out(const struct camera* c, lolb200_camera* out) {
	lolb200__v3_out(c->point, out->point);
	lolb200__v3_out(c->direction, out->direction);
	out->fov = c->fov;
}

/* Returns a scene to release with lolb200_scene_free(). */
static lolb200_scene* lolb200_scene_from_reference(const struct scene* ref) {
	lolb200_scene* s = (lolb200_scene*)calloc(1, sizeof *s);
	size_t nodes = 0, i = 0;

	s->n_materials = (uint32_t)ref->materials->size;
	s->materials = (lolb200_material*)calloc(s->n_materials ? s->n_materials : 1,
	                                         sizeof *s->materials);
	vector_foreach(struct material, ref->materials, m) {
		lolb200_material* d = &s->materials[i++];
		d->shininess = m->shininess;
		lolb200__v3_out(m->diffuse, d->diffuse);
		lolb200__v3_out(m->specular, d->specular);
		lolb200__v3_out(m->ambient, d->ambient);
	}
	lolb200__v3_out(ref->ambient_color, s->ambient_color);

	s->n_lights = (uint32_t)ref->lights->size;
	s->lights = (lolb200_light*)calloc(s->n_lights ? s->n_lights : 1, sizeof *s->lights);
	i = 0;
	vector_foreach(struct light, ref->lights, l) {
		lolb200_light* d = &s->lights[i++];
		lolb200__v3_out(l->point, d->point);
		lolb200__v3_out(l->diffuse_intensity, d->diffuse_intensity);
		lolb200__v3_out(l->specular_intensity, d->specular_intensity);
	}

	vector_foreach(struct object, ref->objects, o) nodes += lolb200__count_nodes(o);
	s->nodes = (lolb200_object*)calloc(nodes ? nodes : 1, sizeof *s->nodes);
	s->objects = (uint32_t*)calloc(ref->objects->size ? ref->objects->size : 1,
	                               sizeof *s->objects);
	vector_foreach(struct object, ref->objects, o)
		s->objects[s->n_objects++] = (uint32_t)lolb200__flatten(o, s);

	lolb200__camera_out(&ref->camera, &s->camera);
	return s;
}

#endif
