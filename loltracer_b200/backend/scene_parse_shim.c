/*
 * scene_parse_shim.c -- scene_parse() for hosts without flex/bison.
 *
 * The reference generates scene_parse() from scene-lexer.l / scene-parser.y
 * (scene-parser.y:197-214); neither tool is installed here.  This file gives
 * the same function on top of our hand-written parser (csrc/lol_parse.c) by
 * replaying the syntax tree through the reference's OWN scene.c API in the
 * order the bison actions call it (scene-parser.y:73-145), so the resulting
 * `struct scene` is built by the reference's code, not ours.
 *
 * Compiled only next to the reference's headers (headless host, oracle/_ref).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "scene.h" /* the reference's */

#include "../csrc/lol_ast.h"

static struct vector* shim_defs_from_node(const struct lol_node* n);

static struct definition_value shim_value(const struct lol_value* v) {
	struct definition_value out;
	memset(&out, 0, sizeof out);
	switch (v->kind) {
	case LOL_V_NUM: /* scene-parser.y:129-132 */
		out.type = VAL_NUM;
		out.num = v->num;
		break;
	case LOL_V_LIST: /* scene-parser.y:133-136,148-160 */
		out.type = VAL_LIST;
		out.list = vector_new(float, 4);
		for (size_t i = 0; i < v->nlist; i++)
			vector_add(float, out.list) = v->list[i];
		break;
	case LOL_V_ID: /* scene-parser.y:137-140 */
		out.type = VAL_ID;
		out.id = v->id;
		break;
	case LOL_V_OBJ: { /* scene-parser.y:141-145 */
		struct vector* defs = shim_defs_from_node(v->obj);
		out.type = VAL_OBJ;
		out.obj = object_from_definition_list(v->obj->type, defs);
		vector_free(defs, definition_free);
		break;
	}
	}
	return out;
}

static struct vector* shim_defs_from_node(const struct lol_node* n) {
	struct vector* defs = vector_new(struct definition, 16);
	for (size_t i = 0; i < n->ndefs; i++) {
		struct definition d;
		d.prop = (enum property)n->defs[i].prop;
		d.value = shim_value(&n->defs[i].value);
		vector_add(struct definition, defs) = d;
	}
	return defs;
}

/* union / intersection / difference are extensions of our front-end; the
 * reference's struct object (scene.h:58-82) has no place for them. */
static int shim_has_extension(const struct lol_node* n) {
	if (n->type == LOL_T_UNION || n->type == LOL_T_INTERSECTION || n->type == LOL_T_DIFFERENCE)
		return 1;
	for (size_t i = 0; i < n->ndefs; i++)
		if (n->defs[i].value.kind == LOL_V_OBJ && shim_has_extension(n->defs[i].value.obj))
			return 1;
	return 0;
}

struct scene* scene_parse_text(const char* text, size_t len) {
	char err[256];
	struct lol_doc* doc = lol_parse_text(text, len, err, sizeof err);
	struct vector* materials;
	struct scene* scene = NULL;

	if (!doc) {
		fprintf(stderr, "Error: %s\n", err); /* yyerror, scene-parser.y:193-195 */
		return NULL;
	}
	for (size_t i = 0; i < doc->ncomponents; i++)
		if (shim_has_extension(&doc->components[i])) {
			fprintf(stderr, "Error: union/intersection/difference nodes are a lolb200 extension; "
			                "the reference's scene.c cannot represent them\n");
			lol_doc_free(doc);
			return NULL;
		}
	materials = vector_new(struct material, 16);
	for (size_t i = 0; i < doc->nmaterials; i++) { /* scene-parser.y:89-103 */
		struct vector* defs = shim_defs_from_node(&doc->materials[i]);
		vector_add(struct material, materials) = material_from_definition_list(defs);
		vector_free(defs, definition_free);
	}
	for (size_t i = 0; i < doc->ncomponents; i++) { /* scene-parser.y:105-114 */
		struct vector* defs = shim_defs_from_node(&doc->components[i]);
		if (!scene)
			scene = scene_new();
		scene_add_component_from_definition_list(scene, doc->components[i].type, defs);
		vector_free(defs, definition_free);
	}
	lol_doc_free(doc);
	if (!scene) {
		vector_free(materials, NULL);
		return NULL;
	}
	vector_free(scene->materials, NULL); /* scene-parser.y:74-77 */
	scene->materials = materials;
	return scene;
}

/* Same signature as the generated parser's entry point (scene-parser.y:197). */
struct scene* scene_parse(const char* filename) {
	FILE* f = filename ? fopen(filename, "rb") : stdin;
	size_t cap = 1 << 16, len = 0;
	char* buf;
	struct scene* scene;

	if (!f)
		return NULL;
	buf = malloc(cap);
	for (;;) {
		size_t n = fread(buf + len, 1, cap - len - 1, f);
		len += n;
		if (n == 0)
			break;
		if (len + 1 >= cap)
			buf = realloc(buf, cap *= 2);
	}
	buf[len] = 0;
	if (f != stdin)
		fclose(f);
	scene = scene_parse_text(buf, len);
	free(buf);
	return scene;
}
