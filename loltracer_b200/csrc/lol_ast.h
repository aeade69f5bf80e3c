/*
 * lol_ast.h -- syntax tree of a .lol scene file (internal to the front-end).
 *
 * The tree is what the reference's bison actions see before they hand
 * definition lists to scene.c (scene-parser.y:73-189).  Two consumers:
 *   - lol_scene.c lowers it to the flat lolb200_scene (product path);
 *   - oracle/ref_harness.c replays it through the reference's own scene.c API
 *     so the compiled reference renders the very same input.
 */
#ifndef LOL_AST_H
#define LOL_AST_H

#include <stddef.h>

/* Numbering equals enum property (scene.h:7-25) and enum components
 * (scene.h:27-35): the harness passes these ints straight to scene.c. */
enum lol_prop {
	LOL_PROP_SHININESS, LOL_PROP_DIFFUSE, LOL_PROP_SPECULAR, LOL_PROP_AMBIENT,
	LOL_PROP_COLOR, LOL_PROP_POINT, LOL_PROP_DIRECTION, LOL_PROP_FOV,
	LOL_PROP_DIFFUSE_INTENSITY, LOL_PROP_SPECULAR_INTENSITY, LOL_PROP_RADIUS,
	LOL_PROP_MATERIAL, LOL_PROP_POINT2, LOL_PROP_Y, LOL_PROP_SMOOTHNESS,
	LOL_PROP_A, LOL_PROP_B
};

enum lol_type {
	LOL_T_AMBIENT, LOL_T_CAMERA, LOL_T_POINT_LIGHT, LOL_T_SPHERE, LOL_T_BOX,
	LOL_T_PLANE, LOL_T_SMOOTH_UNION,
	LOL_T_UNION, LOL_T_INTERSECTION, LOL_T_DIFFERENCE, /* extensions (lolb200.h) */
	LOL_T_MATERIAL = 100 /* a `{ ... }` entry of the materials section */
};

enum lol_value_kind { LOL_V_NUM, LOL_V_LIST, LOL_V_ID, LOL_V_OBJ };

struct lol_node;

struct lol_value {
	int kind;
	float num;
	size_t nlist;
	float* list;
	size_t id;
	struct lol_node* obj;
};

struct lol_def {
	int prop;
	struct lol_value value;
};

struct lol_node {
	int type;
	size_t ndefs;
	struct lol_def* defs;
};

struct lol_doc {
	size_t nmaterials;
	struct lol_node* materials;
	size_t ncomponents;
	struct lol_node* components;
};

/* Returns NULL and fills err (if given) on a syntax error.  The reference
 * prints "Error: ... on line N" and goes on with an unset scene pointer
 * (scene-parser.y:193-195); refusing the file is the only sane restatement. */
struct lol_doc* lol_parse_text(const char* text, size_t len, char* err, size_t errlen);
void lol_doc_free(struct lol_doc* doc);

/* Deepest object nesting the front-end and the lowering accept: both walk the tree
 * recursively, one frame per level (the reference's bison parser gives up with "memory
 * exhausted" at YYMAXDEPTH). */
#define LOLB200_MAX_NESTING 1000

#endif
