/*
 * lol_parse.c -- hand-written front-end for .lol scene files.
 *
 * Token set:  scene-lexer.l:10-50   (flex: longest match, first rule on ties,
 *                                    any unknown character silently dropped)
 * Grammar:    scene-parser.y:73-189 (no trailing commas, no comments)
 *
 * Produces the syntax tree of lol_ast.h; meaning is given to it elsewhere
 * (lol_scene.c for the product, scene.c itself for the compiled reference).
 */
#include "lol_ast.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum tok {
	T_EOF, T_NUM, T_ID, T_MATERIALS, T_SCENE, T_TYPE_AMBIENT, T_TYPE, T_PROP,
	T_PUNCT
};

struct keyword {
	const char* text;
	int tok;
	int value;
};

/* scene-lexer.l:15-46.  "ambient" is one token that the grammar accepts both as
 * a component type and as a material property (scene-parser.y:162,175). */
static const struct keyword keywords[] = {
	{"materials", T_MATERIALS, 0},
	{"scene", T_SCENE, 0},
	{"ambient", T_TYPE_AMBIENT, LOL_T_AMBIENT},
	{"camera", T_TYPE, LOL_T_CAMERA},
	{"point-light", T_TYPE, LOL_T_POINT_LIGHT},
	{"point_light", T_TYPE, LOL_T_POINT_LIGHT},
	{"sphere", T_TYPE, LOL_T_SPHERE},
	{"box", T_TYPE, LOL_T_BOX},
	{"plane", T_TYPE, LOL_T_PLANE},
	{"smooth_union", T_TYPE, LOL_T_SMOOTH_UNION},
	{"smooth-union", T_TYPE, LOL_T_SMOOTH_UNION},
	/* extensions: not in scene-lexer.l */
	{"union", T_TYPE, LOL_T_UNION},
	{"intersection", T_TYPE, LOL_T_INTERSECTION},
	{"difference", T_TYPE, LOL_T_DIFFERENCE},
	{"shininess", T_PROP, LOL_PROP_SHININESS},
	{"diffuse", T_PROP, LOL_PROP_DIFFUSE},
	{"specular", T_PROP, LOL_PROP_SPECULAR},
	{"color", T_PROP, LOL_PROP_COLOR},
	{"point", T_PROP, LOL_PROP_POINT},
	{"direction", T_PROP, LOL_PROP_DIRECTION},
	{"fov", T_PROP, LOL_PROP_FOV},
	{"diffuse_intensity", T_PROP, LOL_PROP_DIFFUSE_INTENSITY},
	{"diffuse-intensity", T_PROP, LOL_PROP_DIFFUSE_INTENSITY},
	{"specular_intensity", T_PROP, LOL_PROP_SPECULAR_INTENSITY},
	{"specular-intensity", T_PROP, LOL_PROP_SPECULAR_INTENSITY},
	{"radius", T_PROP, LOL_PROP_RADIUS},
	{"material", T_PROP, LOL_PROP_MATERIAL},
	{"point2", T_PROP, LOL_PROP_POINT2},
	{"y", T_PROP, LOL_PROP_Y},
	{"smoothness", T_PROP, LOL_PROP_SMOOTHNESS},
	{"a", T_PROP, LOL_PROP_A},
	{"b", T_PROP, LOL_PROP_B},
};

struct parser {
	const char* s;
	size_t len, pos;
	size_t line;
	int tok;     /* current token kind          */
	int tokval;  /* type / property / punct char */
	float num;   /* yylval.NUM: keeps its last value when sscanf fails */
	size_t id;   /* yylval.ID                                           */
	char* err;
	size_t errlen;
	int failed;
	int depth;   /* objects open around the current token */
};

/* An object value nests one level per `type { ... }`.  The reference's bison parser
 * stops with "memory exhausted" when its stack passes YYMAXDEPTH (10000 symbols, about
 * six per nesting level); this parser, and the tree walks behind it (build_object, the
 * lowering's bound_node / emit_node), recurse once per level, so the depth is capped
 * well inside a default thread stack instead of overflowing it. */
#define LOL_MAX_NESTING LOLB200_MAX_NESTING

static int is_numch(char c) { return c == '-' || c == '.' || (c >= '0' && c <= '9'); }

static void fail(struct parser* p, const char* fmt, ...) {
	if (p->failed)
		return;
	p->failed = 1;
	if (p->err && p->errlen) {
		va_list ap;
		int n = snprintf(p->err, p->errlen, "line %zu: ", p->line);
		va_start(ap, fmt);
		if (n >= 0 && (size_t)n < p->errlen)
			vsnprintf(p->err + n, p->errlen - n, fmt, ap);
		va_end(ap);
	}
}

static void next(struct parser* p) {
	for (;;) {
		if (p->pos >= p->len) {
			p->tok = T_EOF;
			return;
		}
		char c = p->s[p->pos];
		if (c == '\n') {
			p->line++;
			p->pos++;
			continue;
		}
		if (c == ' ' || c == '\r' || c == '\t') {
			p->pos++;
			continue;
		}
		if (is_numch(c)) { /* [-.0-9]+ -> sscanf("%f") (scene-lexer.l:12) */
			size_t e = p->pos;
			char buf[64];
			while (e < p->len && is_numch(p->s[e]))
				e++;
			size_t n = e - p->pos;
			if (n >= sizeof buf)
				n = sizeof buf - 1;
			memcpy(buf, p->s + p->pos, n);
			buf[n] = 0;
			sscanf(buf, "%f", &p->num);
			p->pos = e;
			p->tok = T_NUM;
			return;
		}
		if (c == '#' && p->pos + 1 < p->len && p->s[p->pos + 1] >= '0' &&
		    p->s[p->pos + 1] <= '9') { /* #[0-9]+ -> sscanf("%d") (scene-lexer.l:13) */
			size_t e = p->pos + 1;
			char buf[32];
			while (e < p->len && p->s[e] >= '0' && p->s[e] <= '9')
				e++;
			size_t n = e - (p->pos + 1);
			if (n >= sizeof buf)
				n = sizeof buf - 1;
			memcpy(buf, p->s + p->pos + 1, n);
			buf[n] = 0;
			int v = 0;
			sscanf(buf, "%d", &v);
			p->id = (size_t)(unsigned)v;
			p->pos = e;
			p->tok = T_ID;
			return;
		}
		/* keywords: longest match wins */
		const struct keyword* best = NULL;
		size_t bestlen = 0;
		for (size_t i = 0; i < sizeof keywords / sizeof *keywords; i++) {
			size_t kl = strlen(keywords[i].text);
			if (kl > bestlen && p->pos + kl <= p->len &&
			    memcmp(p->s + p->pos, keywords[i].text, kl) == 0) {
				best = &keywords[i];
				bestlen = kl;
			}
		}
		if (best) {
			p->pos += bestlen;
			p->tok = best->tok;
			p->tokval = best->value;
			return;
		}
		if (strchr(",(){}=", c)) {
			p->pos++;
			p->tok = T_PUNCT;
			p->tokval = c;
			return;
		}
		p->pos++; /* scene-lexer.l:50: ignore unexpected characters */
	}
}

static int accept_punct(struct parser* p, int c) {
	if (p->tok == T_PUNCT && p->tokval == c) {
		next(p);
		return 1;
	}
	return 0;
}

static void expect_punct(struct parser* p, int c) {
	if (!accept_punct(p, c))
		fail(p, "syntax error, expected '%c'", c);
}

static void node_free(struct lol_node* n);

static void value_free(struct lol_value* v) {
	free(v->list);
	if (v->obj) {
		node_free(v->obj);
		free(v->obj);
	}
}

static void node_free(struct lol_node* n) {
	for (size_t i = 0; i < n->ndefs; i++)
		value_free(&n->defs[i].value);
	free(n->defs);
	n->defs = NULL;
	n->ndefs = 0;
}

static int is_type_tok(const struct parser* p) {
	return p->tok == T_TYPE || p->tok == T_TYPE_AMBIENT;
}

static void parse_deflist(struct parser* p, struct lol_node* node);

/* value (scene-parser.y:128-146) */
static void parse_value(struct parser* p, struct lol_value* v) {
	memset(v, 0, sizeof *v);
	if (p->tok == T_NUM) {
		v->kind = LOL_V_NUM;
		v->num = p->num;
		next(p);
	} else if (p->tok == T_ID) {
		v->kind = LOL_V_ID;
		v->id = p->id;
		next(p);
	} else if (p->tok == T_PUNCT && p->tokval == '(') {
		size_t cap = 4;
		next(p);
		v->kind = LOL_V_LIST;
		v->list = malloc(cap * sizeof(float));
		do {
			if (p->tok != T_NUM) {
				fail(p, "syntax error, expected a number");
				return;
			}
			if (v->nlist == cap)
				v->list = realloc(v->list, (cap *= 2) * sizeof(float));
			v->list[v->nlist++] = p->num;
			next(p);
		} while (accept_punct(p, ','));
		expect_punct(p, ')');
	} else if (is_type_tok(p)) {
		if (p->depth >= LOL_MAX_NESTING) {
			fail(p, "objects nested deeper than %d levels", LOL_MAX_NESTING);
			return;
		}
		v->kind = LOL_V_OBJ;
		v->obj = calloc(1, sizeof *v->obj);
		v->obj->type = p->tokval;
		next(p);
		expect_punct(p, '{');
		p->depth++;
		parse_deflist(p, v->obj);
		p->depth--;
		expect_punct(p, '}');
	} else {
		fail(p, "syntax error, expected a value");
	}
}

/* definition_list (scene-parser.y:116-126) */
static void parse_deflist(struct parser* p, struct lol_node* node) {
	size_t cap = 8;
	node->defs = malloc(cap * sizeof *node->defs);
	node->ndefs = 0;
	do {
		int prop;
		if (p->failed)
			return;
		if (p->tok == T_PROP)
			prop = p->tokval;
		else if (p->tok == T_TYPE_AMBIENT)
			prop = LOL_PROP_AMBIENT;
		else {
			fail(p, "syntax error, expected a property name");
			return;
		}
		next(p);
		expect_punct(p, '=');
		if (node->ndefs == cap)
			node->defs = realloc(node->defs, (cap *= 2) * sizeof *node->defs);
		struct lol_def* d = &node->defs[node->ndefs++];
		d->prop = prop;
		parse_value(p, &d->value);
	} while (!p->failed && accept_punct(p, ','));
}

struct lol_doc* lol_parse_text(const char* text, size_t len, char* err, size_t errlen) {
	struct parser P = {.s = text, .len = len, .line = 1, .err = err, .errlen = errlen};
	struct parser* p = &P;
	struct lol_doc* doc = calloc(1, sizeof *doc);
	size_t cap;

	if (err && errlen)
		err[0] = 0;
	next(p);

	/* materials (scene-parser.y:80-103) */
	if (p->tok != T_MATERIALS)
		fail(p, "syntax error, expected 'materials'");
	else
		next(p);
	expect_punct(p, '{');
	cap = 16;
	doc->materials = calloc(cap, sizeof *doc->materials);
	while (!p->failed) {
		if (doc->nmaterials == cap) {
			doc->materials = realloc(doc->materials, (cap *= 2) * sizeof *doc->materials);
			memset(doc->materials + doc->nmaterials, 0,
			       (cap - doc->nmaterials) * sizeof *doc->materials);
		}
		struct lol_node* m = &doc->materials[doc->nmaterials++];
		m->type = LOL_T_MATERIAL;
		expect_punct(p, '{');
		if (!p->failed)
			parse_deflist(p, m);
		expect_punct(p, '}');
		if (!accept_punct(p, ','))
			break;
	}
	expect_punct(p, '}');

	/* scene (scene-parser.y:85-113) */
	if (!p->failed) {
		if (p->tok != T_SCENE)
			fail(p, "syntax error, expected 'scene'");
		else
			next(p);
	}
	expect_punct(p, '{');
	cap = 16;
	doc->components = calloc(cap, sizeof *doc->components);
	while (!p->failed) {
		if (!is_type_tok(p)) {
			fail(p, "syntax error, expected a component type");
			break;
		}
		if (doc->ncomponents == cap) {
			doc->components = realloc(doc->components, (cap *= 2) * sizeof *doc->components);
			memset(doc->components + doc->ncomponents, 0,
			       (cap - doc->ncomponents) * sizeof *doc->components);
		}
		struct lol_node* c = &doc->components[doc->ncomponents++];
		c->type = p->tokval;
		next(p);
		expect_punct(p, '{');
		if (!p->failed)
			parse_deflist(p, c);
		expect_punct(p, '}');
		if (!accept_punct(p, ','))
			break;
	}
	expect_punct(p, '}');
	if (!p->failed && p->tok != T_EOF)
		fail(p, "syntax error, trailing input");

	if (p->failed) {
		lol_doc_free(doc);
		return NULL;
	}
	return doc;
}

void lol_doc_free(struct lol_doc* doc) {
	if (!doc)
		return;
	for (size_t i = 0; i < doc->nmaterials; i++)
		node_free(&doc->materials[i]);
	for (size_t i = 0; i < doc->ncomponents; i++)
		node_free(&doc->components[i]);
	free(doc->materials);
	free(doc->components);
	free(doc);
}
