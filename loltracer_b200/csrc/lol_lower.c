/*
 * lol_lower.c -- lowers a scene to specialised CUDA C: the code generator.
 *
 * GPU analogue of generate_sdf() / generate_obj_dist()
 * (tracing_jit_renderer.dasc:76-216): the SDF tree becomes straight-line code
 * with every constant baked in as an immediate, and the scene-independent
 * pipeline text (lol_kernel.cuh) is spliced around it.  Long runs of top-level
 * objects of one shape (the 1024-primitive synthetic scene) become a loop over a
 * __constant__ parameter table instead of 25k unrolled instructions per call
 * site.  Pure host C: testable without a device.
 */
#include "lol_internal.h"
#include "lolb200.h"

#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

extern const char lol_kernel_text[]; /* lol_kernel.cuh, embedded by lol_kernel_text.c */
extern const char lol_params_text[]; /* lol_params.h, likewise */

/* ------------------------------------------------------------ string builder */

struct sb {
	char* p;
	size_t len, cap;
};

static void sb_reserve(struct sb* b, size_t more) {
	if (b->len + more + 1 > b->cap) {
		while (b->len + more + 1 > b->cap)
			b->cap = b->cap ? b->cap * 2 : 1 << 14;
		b->p = realloc(b->p, b->cap);
	}
}

static void sb_putn(struct sb* b, const char* s, size_t n) {
	sb_reserve(b, n);
	memcpy(b->p + b->len, s, n);
	b->len += n;
	b->p[b->len] = 0;
}

static void sb_printf(struct sb* b, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
static void sb_printf(struct sb* b, const char* fmt, ...) {
	va_list ap;
	char tmp[512];
	int n;
	va_start(ap, fmt);
	n = vsnprintf(tmp, sizeof tmp, fmt, ap);
	va_end(ap);
	if (n < 0)
		return;
	if ((size_t)n < sizeof tmp) {
		sb_putn(b, tmp, (size_t)n);
		return;
	}
	sb_reserve(b, (size_t)n);
	va_start(ap, fmt);
	vsnprintf(b->p + b->len, (size_t)n + 1, fmt, ap);
	va_end(ap);
	b->len += (size_t)n;
}

static uint32_t f2u(float f) {
	uint32_t u;
	memcpy(&u, &f, 4);
	return u;
}

/* A float as the kernel sees it: exact bits, with the decimal value as a comment. */
static void sb_float(struct sb* b, float f) {
	sb_printf(b, "LOL_F(0x%08x /*%.9g*/)", f2u(f), (double)f);
}

/* The same inside a table initialiser: tables hold raw bits (static
 * initialisers cannot call __int_as_float) and are read through LOL_TF(). */
static void sb_bits(struct sb* b, float f) {
	sb_printf(b, "0x%08xu /*%.9g*/", f2u(f), (double)f);
}

/* ------------------------------------------------------------ distance code */

/* Where a constant comes from: an immediate, or a slot of the current row of a
 * run table (loop mode).  A row is laid out before its code is written: the box
 * and the object id first (cst / cst_raw, in emission order), then every node's
 * constants in post-order (layout_object), so that the reference form, the
 * guarded form and the packed pairs -- which read them in different orders --
 * agree on where each one lives. */
struct pair_plan {
	uint32_t a, b; /* roots of two disjoint subtrees of one shape: low / high half */
	int tmp;       /* the lol_f2 temporary that holds both results, -1 = not emitted yet */
};
struct pairc { /* the constant pairs of one function: (low, high) bits */
	uint32_t (*w)[2];
	size_t n, cap;
};
#define LOL_GRID_OUTER_F 4.f /* the outer candidate grid's cells are this many times as large (lol_near_grid_text) */

struct cgen {
	const lolb200_scene* s;
	struct sb* out;
	int tmp;      /* next temporary number                      */
	int in_loop;  /* 1: constants are c[j], gathered into row[] */
	int fast;     /* 1: guarded fast forms (lol_sqrt_fast, lol_smin_c) */
	int div_ok;   /* 1: every smoothness passed the constant-division proof */
	int two;      /* 1: two rays per evaluation, packed FP32 (lol_f2, variant 3) */
	int pack;     /* 1: same-shaped subtrees of one ray in pairs, packed FP32 (plan_pairs) */
	float* row;
	size_t nrow, caprow;
	size_t* slot;        /* loop mode: first row slot of each node's constants */
	unsigned char* stride; /* ... and the distance between its fields: 2 inside a pair (a.f, b.f adjacent) */
	int lay_pack, lay_div_ok; /* the pairing of the GUARDED form: every form of a program reads one row layout */
	uint32_t* pair_of;   /* per node: 0, or 1 + 2 * pair + half for the roots of a planned pair */
	struct pair_plan* pairs;
	size_t npairs;
	struct pairc* pc;    /* where constant pairs of straight-line code go */
	const char* indent;
	int first_free;      /* guarded single-ray form: nothing was evaluated yet, `best` is still +INF (emit_straight_object) */
};

static void row_push(struct cgen* g, float v) {
	if (g->nrow == g->caprow) {
		g->caprow = g->caprow ? g->caprow * 2 : 16;
		g->row = realloc(g->row, g->caprow * sizeof(float));
	}
	g->row[g->nrow++] = v;
}

/* a constant that is not a node's (a bounding box): next slot of the row */
static void cst(struct cgen* g, float v) {
	if (!g->in_loop) {
		sb_float(g->out, v);
		return;
	}
	sb_printf(g->out, "LOL_TF(c[%zu])", g->nrow);
	row_push(g, v);
}

/* An integer constant (an object id): raw bits in the table, read without LOL_TF. */
static void cst_raw(struct cgen* g, uint32_t v) {
	float f;
	if (!g->in_loop) {
		sb_printf(g->out, "%uu", v);
		return;
	}
	sb_printf(g->out, "c[%zu]", g->nrow);
	memcpy(&f, &v, 4);
	row_push(g, f);
}

/* A node's constants: sphere (centre, radius), rounded box (centre and extent per
 * axis, radius), plane (y), smooth union (k, RN(1/k) / 2, 2k); CSG nodes have none. */
enum { F_PX, F_PY, F_PZ, F_R, F_EX, F_EY, F_EZ, F_K, F_RKH, F_K2 };

static int node_slots(const lolb200_object* o) {
	switch (o->type) {
	case LOLB200_OBJ_SPHERE: return 4;
	case LOLB200_OBJ_BOX: return 7;
	case LOLB200_OBJ_PLANE: return 1;
	case LOLB200_OBJ_SMOOTH_UNION: return 3;
	default: return 0;
	}
}

static int field_slot(const lolb200_object* o, int f) {
	switch (o->type) {
	case LOLB200_OBJ_SPHERE: return f == F_R ? 3 : f - F_PX;
	case LOLB200_OBJ_BOX: return f == F_R ? 6 : f <= F_PZ ? 2 * (f - F_PX) : 2 * (f - F_EX) + 1;
	case LOLB200_OBJ_PLANE: return 0;
	default: return f - F_K;
	}
}

static float field_value(const lolb200_object* o, int f) {
	switch (f) {
	case F_PX: case F_PY: case F_PZ: return o->point[f - F_PX];
	case F_R: return o->radius;
	case F_EX: case F_EY: case F_EZ: return o->point2[f - F_EX];
	case F_K: return o->smoothness;
	case F_RKH: return (1.0f / o->smoothness) * 0.5f; /* RN(1/k) / 2, see lol_smin_c */
	default: return o->smoothness * 2.0f;
	}
}

static const int lol_field_order[4][7] = {{F_PX, F_PY, F_PZ, F_R}, {F_PX, F_EX, F_PY, F_EY, F_PZ, F_EZ, F_R}, {F_PY}, {F_K, F_RKH, F_K2}};
static const int* node_fields(const lolb200_object* o) {
	return lol_field_order[o->type == LOLB200_OBJ_SPHERE ? 0 : o->type == LOLB200_OBJ_BOX ? 1 : o->type == LOLB200_OBJ_PLANE ? 2 : 3];
}

/* the constants of the paired subtrees (a, b): field by field, a's next to b's, at an
 * even slot -- one 64-bit load hands the packed code its (low, high) operand */
static void layout_pair(struct cgen* g, uint32_t a, uint32_t b) {
	const lolb200_object *oa = &g->s->nodes[a], *ob = &g->s->nodes[b];
	const int* fields = node_fields(oa);
	if (oa->type != LOLB200_OBJ_SPHERE) {
		layout_pair(g, (uint32_t)oa->a, (uint32_t)ob->a);
		layout_pair(g, (uint32_t)oa->b, (uint32_t)ob->b);
	}
	if (node_slots(oa) && (g->nrow & 1u))
		row_push(g, 0.f);
	g->slot[a] = g->nrow;
	g->slot[b] = g->nrow + 1;
	g->stride[a] = g->stride[b] = 2;
	for (int q = 0; q < node_slots(oa); q++) {
		row_push(g, field_value(oa, fields[q]));
		row_push(g, field_value(ob, fields[q]));
	}
}

/* loop mode: the constants of the object rooted at idx go into the row, post-order */
static void layout_node(struct cgen* g, uint32_t idx) {
	const lolb200_object* o = &g->s->nodes[idx];
	const int* fields = node_fields(o);
	if (g->npairs && g->pair_of[idx]) {
		const struct pair_plan* p = &g->pairs[(g->pair_of[idx] - 1) / 2];
		if (idx == p->a) /* (b comes later in the tree; its slots exist by then) */
			layout_pair(g, p->a, p->b);
		return;
	}
	if (o->type != LOLB200_OBJ_SPHERE && o->type != LOLB200_OBJ_BOX && o->type != LOLB200_OBJ_PLANE) {
		layout_node(g, (uint32_t)o->a);
		layout_node(g, (uint32_t)o->b);
	}
	g->slot[idx] = g->nrow;
	g->stride[idx] = 1;
	for (int q = 0; q < node_slots(o); q++)
		row_push(g, field_value(o, fields[q]));
}

/* field f of node idx as a float expression */
static void ncst(struct cgen* g, uint32_t idx, int f) {
	const lolb200_object* o = &g->s->nodes[idx];
	if (!g->in_loop)
		sb_float(g->out, field_value(o, f));
	else
		sb_printf(g->out, "LOL_TF(c[%zu])", g->slot[idx] + (size_t)g->stride[idx] * (size_t)field_slot(o, f));
}

static int is_pos_zero(float v) { return f2u(v) == 0u; }

/* `x - c`; x - (+0) is x for every x, so it is dropped outside loops. */
static void coord_minus(struct cgen* g, const char* x, uint32_t idx, int f) {
	if (!g->in_loop && is_pos_zero(field_value(&g->s->nodes[idx], f))) {
		sb_printf(g->out, "%s", x);
		return;
	}
	sb_printf(g->out, "%s - ", x);
	ncst(g, idx, f);
}

/* ---- pairs: two subtrees of one shape, one packed instruction stream ------------
 * lol_render is bound by instruction issue (ncu: 94 % of the slots), and sm_100a's
 * two-wide FP32 instructions do two IEEE operations per slot.  Variant 3 fills the
 * halves with two RAYS and pays for it in registers and bookkeeping; here the halves
 * are two SUBTREES of the same ray: spheres (0, 4) and (1, 5) of scene4's blob, the
 * two halves of a balanced smooth-union tree.  Same operations, same order, same
 * rounding per half -- only the instruction count changes (scene4's march loop:
 * 174 -> 146).  Chosen by shape alone, so every row of a table loop pairs alike. */
static int packable(const struct cgen* g, uint32_t idx, int div_ok) {
	const lolb200_object* o = &g->s->nodes[idx];
	switch (o->type) {
	case LOLB200_OBJ_SPHERE: return 1;
	case LOLB200_OBJ_SMOOTH_UNION:
		return div_ok && packable(g, (uint32_t)o->a, div_ok) && packable(g, (uint32_t)o->b, div_ok);
	case LOLB200_OBJ_UNION:
	case LOLB200_OBJ_INTERSECTION:
	case LOLB200_OBJ_DIFFERENCE: return packable(g, (uint32_t)o->a, div_ok) && packable(g, (uint32_t)o->b, div_ok);
	default: return 0;
	}
}

static void signature(const lolb200_scene* s, uint32_t idx, struct sb* out);

struct pair_cand {
	uint32_t idx, size, order;
	char* sig;
};

static uint32_t subtree_size(const lolb200_scene* s, uint32_t idx) {
	const lolb200_object* o = &s->nodes[idx];
	if (o->type == LOLB200_OBJ_SPHERE || o->type == LOLB200_OBJ_BOX || o->type == LOLB200_OBJ_PLANE)
		return 1;
	return 1 + subtree_size(s, (uint32_t)o->a) + subtree_size(s, (uint32_t)o->b);
}

static void collect_cands(const struct cgen* g, uint32_t idx, int div_ok, struct pair_cand* c, uint32_t* n) {
	const lolb200_object* o = &g->s->nodes[idx];
	if (packable(g, idx, div_ok)) {
		struct sb sig = {0};
		signature(g->s, idx, &sig);
		c[*n].idx = idx;
		c[*n].size = subtree_size(g->s, idx);
		c[*n].order = *n;
		c[*n].sig = sig.p;
		++*n;
	}
	if (o->type != LOLB200_OBJ_SPHERE && o->type != LOLB200_OBJ_BOX && o->type != LOLB200_OBJ_PLANE) {
		collect_cands(g, (uint32_t)o->a, div_ok, c, n);
		collect_cands(g, (uint32_t)o->b, div_ok, c, n);
	}
}

static int cand_cmp(const void* a, const void* b) {
	const struct pair_cand *x = a, *y = b;
	return x->size != y->size ? (x->size < y->size) - (x->size > y->size) : (x->order > y->order) - (x->order < y->order);
}

static void mark_subtree(const lolb200_scene* s, uint32_t idx, unsigned char* covered) {
	const lolb200_object* o = &s->nodes[idx];
	covered[idx] = 1;
	if (o->type != LOLB200_OBJ_SPHERE && o->type != LOLB200_OBJ_BOX && o->type != LOLB200_OBJ_PLANE) {
		mark_subtree(s, (uint32_t)o->a, covered);
		mark_subtree(s, (uint32_t)o->b, covered);
	}
}

/* Largest subtrees first: each packable subtree not yet inside a pair takes the next
 * free subtree of the same shape (in tree order) as its partner. */
static void plan_pairs(struct cgen* g, uint32_t root, int enabled) {
	const uint32_t total = subtree_size(g->s, root);
	struct pair_cand* c;
	unsigned char* covered;
	uint32_t n = 0;
	g->npairs = 0;
	if (!enabled || total < 2)
		return;
	c = malloc(sizeof *c * total);
	collect_cands(g, root, enabled > 1 ? g->lay_div_ok : g->div_ok, c, &n);
	if ((enabled > 1 ? g->lay_pack : g->pack) == 3) { /* pack_pairs = 3: leaves only (an A/B point) */
		uint32_t m = 0;
		for (uint32_t i = 0; i < n; i++)
			if (c[i].size == 1)
				c[m++] = c[i];
			else
				free(c[i].sig);
		n = m;
	}
	if (n >= 2) {
		covered = calloc(g->s->n_nodes, 1);
		if (!g->pair_of)
			g->pair_of = calloc(g->s->n_nodes, sizeof *g->pair_of);
		g->pairs = realloc(g->pairs, sizeof *g->pairs * (n / 2));
		qsort(c, n, sizeof *c, cand_cmp);
		for (uint32_t i = 0; i < n; i++) {
			if (covered[c[i].idx])
				continue;
			for (uint32_t j = i + 1; j < n && c[j].size == c[i].size; j++) {
				if (covered[c[j].idx] || strcmp(c[i].sig, c[j].sig) != 0)
					continue;
				g->pairs[g->npairs].a = c[i].idx;
				g->pairs[g->npairs].b = c[j].idx;
				g->pairs[g->npairs].tmp = -1;
				g->pair_of[c[i].idx] = 1 + 2 * (uint32_t)g->npairs;
				g->pair_of[c[j].idx] = 2 + 2 * (uint32_t)g->npairs;
				g->npairs++;
				mark_subtree(g->s, c[i].idx, covered);
				mark_subtree(g->s, c[j].idx, covered);
				break;
			}
		}
		free(covered);
	}
	for (uint32_t i = 0; i < n; i++)
		free(c[i].sig);
	free(c);
}

static void unplan_pairs(struct cgen* g) {
	for (size_t i = 0; i < g->npairs; i++)
		g->pair_of[g->pairs[i].a] = g->pair_of[g->pairs[i].b] = 0;
	g->npairs = 0;
}

static void cgen_release(struct cgen* g) {
	free(g->row);
	free(g->slot);
	free(g->stride);
	free(g->pair_of);
	free(g->pairs);
	g->row = NULL;
	g->slot = NULL;
	g->stride = NULL;
	g->pair_of = NULL;
	g->pairs = NULL;
	g->nrow = g->caprow = g->npairs = 0;
}

/* LOL_PC(i): the i-th constant pair of this function (lol_pairc[], emitted in front of it) */
static size_t pair_const(struct cgen* g, float lo, float hi) {
	struct pairc* pc = g->pc;
	for (size_t i = 0; i < pc->n; i++)
		if (pc->w[i][0] == f2u(lo) && pc->w[i][1] == f2u(hi))
			return i;
	if (pc->n == pc->cap) {
		pc->cap = pc->cap ? pc->cap * 2 : 16;
		pc->w = realloc(pc->w, pc->cap * sizeof *pc->w);
	}
	pc->w[pc->n][0] = f2u(lo);
	pc->w[pc->n][1] = f2u(hi);
	return pc->n++;
}

/* field f of nodes (a, b) as one lol_f2 expression */
static void pcst(struct cgen* g, uint32_t a, uint32_t b, int f) {
	if (g->in_loop) {
		/* layout_pair put them side by side at an even slot */
		sb_printf(g->out, "lol_ld2_(c + %zu)", g->slot[a] + 2u * (size_t)field_slot(&g->s->nodes[a], f));
	} else {
		const float va = field_value(&g->s->nodes[a], f), vb = field_value(&g->s->nodes[b], f);
		sb_printf(g->out, "LOL_PC(%zu) /*%.9g, %.9g*/", pair_const(g, va, vb), (double)va, (double)vb);
	}
}

/* `lhs - (field f of a, of b)`, both halves at once.  Outside loops: nothing when both
 * are +0, a broadcast immediate when they are equal, else lhs + (-va, -vb) from the
 * constant pairs (x - c and x + (-c) are the same IEEE operation). */
static void pair_minus(struct cgen* g, const char* lhs, uint32_t a, uint32_t b, int f) {
	const float va = field_value(&g->s->nodes[a], f), vb = field_value(&g->s->nodes[b], f);
	if (g->in_loop) {
		sb_printf(g->out, "%s - ", lhs);
		pcst(g, a, b, f);
	} else if (is_pos_zero(va) && is_pos_zero(vb)) {
		sb_printf(g->out, "%s", lhs);
	} else if (f2u(va) == f2u(vb)) {
		sb_printf(g->out, "%s - ", lhs);
		sb_float(g->out, va);
	} else {
		sb_printf(g->out, "%s + LOL_PC(%zu) /*-(%.9g, %.9g)*/", lhs, pair_const(g, -va, -vb), (double)va, (double)vb);
	}
}

/* The subtrees rooted at a and b (one shape) evaluated together:
 * `const lol_f2 pN = (dist(a, p), dist(b, p));`.  Returns N. */
static int emit_pair(struct cgen* g, uint32_t a, uint32_t b) {
	const lolb200_object *oa = &g->s->nodes[a], *ob = &g->s->nodes[b];
	int me;
	if (oa->type == LOLB200_OBJ_SPHERE) { /* sdSphere, sdf.h:8-10, twice */
		char lhs[48];
		me = g->tmp++;
		sb_printf(g->out, "%sconst lol_f2 qx%d = ", g->indent, me);
		pair_minus(g, "lol_bc(x)", a, b, F_PX);
		sb_printf(g->out, ", qy%d = ", me);
		pair_minus(g, "lol_bc(y)", a, b, F_PY);
		sb_printf(g->out, ", qz%d = ", me);
		pair_minus(g, "lol_bc(z)", a, b, F_PZ);
		sb_printf(g->out, ";\n%sconst lol_f2 s%d = lol_dot2(qx%d, qy%d, qz%d, qx%d, qy%d, qz%d);\n", g->indent, me, me,
		          me, me, me, me, me);
		sb_printf(g->out, "%slo = lol_min_halves(lo, s%d);\n", g->indent, me);
		snprintf(lhs, sizeof lhs, "lol_sqrt_fast2(s%d)", me);
		sb_printf(g->out, "%sconst lol_f2 p%d = ", g->indent, me);
		pair_minus(g, lhs, a, b, F_R);
		sb_printf(g->out, ";\n");
		return me;
	}
	{
		const int pa = emit_pair(g, (uint32_t)oa->a, (uint32_t)ob->a);
		const int pb = emit_pair(g, (uint32_t)oa->b, (uint32_t)ob->b);
		me = g->tmp++;
		if (oa->type != LOLB200_OBJ_SMOOTH_UNION) {
			sb_printf(g->out, "%sconst lol_f2 p%d = lol_csg_%s(p%d, p%d);\n", g->indent, me,
			          oa->type == LOLB200_OBJ_UNION ? "union" : oa->type == LOLB200_OBJ_INTERSECTION ? "inter" : "diff",
			          pa, pb);
		} else if (!g->in_loop && f2u(oa->smoothness) == f2u(ob->smoothness)) {
			sb_printf(g->out, "%sconst lol_f2 p%d = lol_smin_c2(p%d, p%d, ", g->indent, me, pa, pb);
			ncst(g, a, F_K);
			sb_printf(g->out, ", ");
			ncst(g, a, F_RKH);
			sb_printf(g->out, ", ");
			ncst(g, a, F_K2);
			sb_printf(g->out, ");\n");
		} else {
			sb_printf(g->out, "%sconst lol_f2 p%d = lol_smin_c2v(p%d, p%d, ", g->indent, me, pa, pb);
			pcst(g, a, b, F_K);
			sb_printf(g->out, ", ");
			pcst(g, a, b, F_RKH);
			sb_printf(g->out, ", ");
			pcst(g, a, b, F_K2);
			sb_printf(g->out, ");\n");
		}
	}
	return me;
}

/* get_obj_dist (naive_renderer.c:10-28), one object -> `const float tN = ...;`.
 * Returns N.  Children of a smooth union see the same (x, y, z). */
static int emit_node(struct cgen* g, uint32_t idx) {
	const lolb200_object* o = &g->s->nodes[idx];
	const char* T = g->two ? "lol_f2" : "float";
	int me;

	if (g->npairs && g->pair_of[idx]) {
		/* one half of a planned pair: both subtrees are evaluated where the first is needed */
		struct pair_plan* p = &g->pairs[(g->pair_of[idx] - 1) / 2];
		if (p->tmp < 0)
			p->tmp = emit_pair(g, p->a, p->b);
		me = g->tmp++;
		sb_printf(g->out, "%sconst float t%d = lol_%s(p%d);\n", g->indent, me,
		          ((g->pair_of[idx] - 1) & 1) ? "hi" : "lo", p->tmp);
		return me;
	}
	switch (o->type) {
	case LOLB200_OBJ_SPHERE: /* sdSphere, sdf.h:8-10 */
		me = g->tmp++;
		if (g->fast) {
			/* same operations; the squared length also feeds the range guard */
			sb_printf(g->out, "%sconst %s qx%d = ", g->indent, T, me);
			coord_minus(g, "x", idx, F_PX);
			sb_printf(g->out, ", qy%d = ", me);
			coord_minus(g, "y", idx, F_PY);
			sb_printf(g->out, ", qz%d = ", me);
			coord_minus(g, "z", idx, F_PZ);
			sb_printf(g->out, ";\n%sconst %s s%d = lol_dot%s(qx%d, qy%d, qz%d, qx%d, qy%d, qz%d);\n",
			          g->indent, T, me, g->two ? "2" : "", me, me, me, me, me, me);
			if (g->two)
				sb_printf(g->out, "%slo = lol_min_halves(lo, s%d);\n", g->indent, me);
			else
				sb_printf(g->out, "%slo = lol_min_nan(lo, s%d);\n", g->indent, me);
			sb_printf(g->out, "%sconst %s t%d = lol_sqrt_fast%s(s%d) - ", g->indent, T, me,
			          g->two ? "2" : "", me);
			ncst(g, idx, F_R);
			sb_printf(g->out, ";\n");
			return me;
		}
		sb_printf(g->out, "%sconst float t%d = lol_len(", g->indent, me);
		coord_minus(g, "x", idx, F_PX);
		sb_printf(g->out, ", ");
		coord_minus(g, "y", idx, F_PY);
		sb_printf(g->out, ", ");
		coord_minus(g, "z", idx, F_PZ);
		sb_printf(g->out, ") - ");
		ncst(g, idx, F_R);
		sb_printf(g->out, ";\n");
		return me;
	case LOLB200_OBJ_BOX: /* sdRoundBox, sdf.h:18-22 */
		me = g->tmp++;
		if (g->two) {
			/* no two-wide abs/max: the box runs per half on the packed coordinates */
			sb_printf(g->out, "%sconst lol_f2 t%d = lol_roundbox2(", g->indent, me);
			for (int k = 0; k < 3; k++) {
				coord_minus(g, k == 0 ? "x" : k == 1 ? "y" : "z", idx, F_PX + k);
				sb_printf(g->out, ", ");
				ncst(g, idx, F_EX + k);
				sb_printf(g->out, ", ");
			}
			ncst(g, idx, F_R);
			sb_printf(g->out, ");\n");
			return me;
		}
		sb_printf(g->out, "%sconst float t%d = lol_roundbox(fabsf(", g->indent, me);
		coord_minus(g, "x", idx, F_PX);
		sb_printf(g->out, ") - ");
		ncst(g, idx, F_EX);
		sb_printf(g->out, ", fabsf(");
		coord_minus(g, "y", idx, F_PY);
		sb_printf(g->out, ") - ");
		ncst(g, idx, F_EY);
		sb_printf(g->out, ", fabsf(");
		coord_minus(g, "z", idx, F_PZ);
		sb_printf(g->out, ") - ");
		ncst(g, idx, F_EZ);
		sb_printf(g->out, ", ");
		ncst(g, idx, F_R);
		sb_printf(g->out, ");\n");
		return me;
	case LOLB200_OBJ_PLANE: /* point.y, naive_renderer.c:19-20 */
		me = g->tmp++;
		sb_printf(g->out, "%sconst %s t%d = ", g->indent, T, me);
		coord_minus(g, "y", idx, F_PY);
		sb_printf(g->out, ";\n");
		return me;
	case LOLB200_OBJ_UNION:
	case LOLB200_OBJ_INTERSECTION:
	case LOLB200_OBJ_DIFFERENCE: { /* extensions: minf(a,b) / maxf(a,b) / maxf(a,-b) */
		int a = emit_node(g, (uint32_t)o->a);
		int b = emit_node(g, (uint32_t)o->b);
		me = g->tmp++;
		sb_printf(g->out, "%sconst %s t%d = lol_csg_%s(t%d, t%d);\n", g->indent, T, me,
		          o->type == LOLB200_OBJ_UNION ? "union" :
		          o->type == LOLB200_OBJ_INTERSECTION ? "inter" : "diff", a, b);
		return me;
	}
	default: { /* sminf(a, b, k), naive_renderer.c:21-24 */
		int a = emit_node(g, (uint32_t)o->a);
		int b = emit_node(g, (uint32_t)o->b);
		me = g->tmp++;
		if (g->fast && g->div_ok) {
			sb_printf(g->out, "%sconst %s t%d = lol_smin_c%s(t%d, t%d, ", g->indent, T, me,
			          g->two ? "2" : "", a, b);
			ncst(g, idx, F_K);
			sb_printf(g->out, ", ");
			ncst(g, idx, F_RKH);
			sb_printf(g->out, ", ");
			ncst(g, idx, F_K2);
			sb_printf(g->out, ");\n");
		} else {
			sb_printf(g->out, "%sconst %s t%d = lol_smin%s(t%d, t%d, ", g->indent, T, me,
			          g->two ? "2" : "", a, b);
			ncst(g, idx, F_K);
			sb_printf(g->out, ");\n");
		}
		return me;
	}
	}
}

/* One top-level object: its row layout (loop mode), its pairs, its code. */
static int emit_object(struct cgen* g, uint32_t root) {
	const int pairing = g->pack && g->fast && !g->two;
	int t;
	if (g->in_loop) {
		if (!g->slot) {
			g->slot = calloc(g->s->n_nodes ? g->s->n_nodes : 1, sizeof *g->slot);
			g->stride = calloc(g->s->n_nodes ? g->s->n_nodes : 1, 1);
		}
		/* the layout follows the guarded form's pairs whichever form is being written */
		plan_pairs(g, root, g->lay_pack ? 2 : 0);
		layout_node(g, root);
		while (g->nrow % 4) /* rows of 16-byte multiples: the loads of neighbouring slots merge */
			row_push(g, 0.f);
		unplan_pairs(g);
	}
	plan_pairs(g, root, pairing);
	t = emit_node(g, root);
	unplan_pairs(g);
	return t;
}

/* Shape of a subtree without its constants: objects with equal signatures run
 * the same instructions on different table rows. */
static void signature(const lolb200_scene* s, uint32_t idx, struct sb* out) {
	const lolb200_object* o = &s->nodes[idx];
	switch (o->type) {
	case LOLB200_OBJ_SPHERE: sb_putn(out, "S", 1); break;
	case LOLB200_OBJ_BOX: sb_putn(out, "B", 1); break;
	case LOLB200_OBJ_PLANE: sb_putn(out, "P", 1); break;
	default:
		sb_putn(out, o->type == LOLB200_OBJ_UNION ? "N(" : o->type == LOLB200_OBJ_INTERSECTION ? "I(" :
		             o->type == LOLB200_OBJ_DIFFERENCE ? "D(" : "U(", 2);
		signature(s, (uint32_t)o->a, out);
		sb_putn(out, ",", 1);
		signature(s, (uint32_t)o->b, out);
		sb_putn(out, ")", 1);
	}
}

/* Conservative bounding BOX of an object's distance field: with
 *     dbox(p) = | max(|p - C| - H, 0) |      (distance to the box, 0 inside)
 *     dist(obj, p) >= dbox(p) - M            for every p (in exact arithmetic).
 * sphere: the cube around it, M = 0 (outside the sphere |p-c| - r >= dbox; inside
 * it p is inside the cube, dbox = 0 and the test below never skips).  rounded box:
 * half extents + radius.  (smooth) union: the box around both children's boxes,
 * M = max(children) + k/4, because sminf(a,b,k) >= min(a,b) - k/4 (float.h:29-33)
 * and min_i dbox_i >= dbox of the union box.  intersection, max(a,b): either
 * child's bound holds, the smaller box is kept.  difference, max(a,-b): a's.  A
 * plane has no box (returns 0).  Computed in double.
 *
 * A box instead of a ball because objects are rarely round: a smooth-union chain
 * of eight spheres along x has a ball of radius 4.3 around a 1 x 1 x 8 body.
 * On the 1024-sphere scene the running minimum leaves 7.7 of 128 objects per
 * evaluation to compute with boxes against 18.0 with balls. */
static int bound_node(const lolb200_scene* s, uint32_t idx, double lo[3], double hi[3], double* M) {
	const lolb200_object* o = &s->nodes[idx];
	switch (o->type) {
	case LOLB200_OBJ_SPHERE:
		for (int k = 0; k < 3; k++) {
			lo[k] = (double)o->point[k] - fabs((double)o->radius);
			hi[k] = (double)o->point[k] + fabs((double)o->radius);
		}
		*M = 0;
		return isfinite(lo[0] + lo[1] + lo[2] + hi[0] + hi[1] + hi[2]);
	case LOLB200_OBJ_BOX:
		for (int k = 0; k < 3; k++) {
			const double e = fabs((double)o->point2[k]) + fabs((double)o->radius);
			lo[k] = (double)o->point[k] - e;
			hi[k] = (double)o->point[k] + e;
		}
		*M = 0;
		/* sdRoundBox with a negative extent or radius is not the distance to a box */
		return o->radius >= 0.f && o->point2[0] >= 0.f && o->point2[1] >= 0.f && o->point2[2] >= 0.f &&
		       isfinite(lo[0] + lo[1] + lo[2] + hi[0] + hi[1] + hi[2]);
	case LOLB200_OBJ_PLANE: return 0;
	case LOLB200_OBJ_DIFFERENCE: /* maxf(a, -b) >= a */
		return bound_node(s, (uint32_t)o->a, lo, hi, M);
	case LOLB200_OBJ_INTERSECTION: { /* maxf(a, b) >= a and >= b: the smaller box */
		double lb[3], hb[3], Mb;
		const int ha = bound_node(s, (uint32_t)o->a, lo, hi, M);
		const int hb_ = bound_node(s, (uint32_t)o->b, lb, hb, &Mb);
		if (hb_) {
			const double va = ha ? (hi[0] - lo[0]) + (hi[1] - lo[1]) + (hi[2] - lo[2]) + *M : INFINITY;
			const double vb = (hb[0] - lb[0]) + (hb[1] - lb[1]) + (hb[2] - lb[2]) + Mb;
			if (vb < va) {
				memcpy(lo, lb, sizeof lb);
				memcpy(hi, hb, sizeof hb);
				*M = Mb;
			}
		}
		return ha || hb_;
	}
	default: { /* (smooth) union */
		double la[3], ha[3], lb[3], hb[3], Ma, Mb;
		const double k = o->type == LOLB200_OBJ_UNION ? 0.0 : (double)o->smoothness;
		if (!(k >= 0.0) || !isfinite(k))
			return 0;
		if (!bound_node(s, (uint32_t)o->a, la, ha, &Ma) || !bound_node(s, (uint32_t)o->b, lb, hb, &Mb))
			return 0;
		for (int c = 0; c < 3; c++) {
			lo[c] = la[c] < lb[c] ? la[c] : lb[c];
			hi[c] = ha[c] > hb[c] ? ha[c] : hb[c];
		}
		*M = (Ma > Mb ? Ma : Mb) + 0.25 * k;
		return 1;
	}
	}
}

/* The box as the kernel uses it -- centre, half extents, margin: extents padded by
 * 0.2 % plus 0.002 * (|C|_1 + 1), the margin by 0.2 % -- three orders of magnitude
 * above the rounding error of the evaluated distance (a few ulps of the
 * coordinates per tree level) -- and the kernel adds a relative 0.4 % on the
 * distance side, so the test stays conservative at any scene scale.  Unboundable
 * objects get infinite extents: dbox = 0, never skipped. */
#define LOL_BOUND_SLOTS 7
static void bound_row(const lolb200_scene* s, uint32_t idx, float row[LOL_BOUND_SLOTS]) {
	double lo[3], hi[3], M;
	if (!bound_node(s, idx, lo, hi, &M)) {
		row[0] = row[1] = row[2] = 0.f;
		row[3] = row[4] = row[5] = INFINITY;
		row[6] = 0.f;
		return;
	}
	const double l1 = fabs(lo[0] + hi[0]) * 0.5 + fabs(lo[1] + hi[1]) * 0.5 + fabs(lo[2] + hi[2]) * 0.5;
	for (int k = 0; k < 3; k++) {
		row[k] = (float)((lo[k] + hi[k]) * 0.5);
		row[3 + k] = (float)((hi[k] - lo[k]) * 0.5 * 1.002 + 0.002 * (l1 + 1.0) + 1e-6);
	}
	/* slot 6 is the margin times the kernel's 1.004 (lol_box_skips: u = best * 1.004 + m1) */
	row[6] = nextafterf((float)(M * 1.002 + 1e-6) * 1.004f, INFINITY);
}

/* A bounding BALL around one of the object's own sphere centres (straight-line objects, emit_sdf_fn).
 * The box test costs 16 instructions on every evaluation; the squared distance to a sphere's centre is
 * computed by the object anyway, so a ball around that centre is tested with four.  Every leaf lies within
 * R of the centre c (sphere: |c_i - c| + r_i; round box: |c_i - c| + |half extents| + r_i), and composites
 * only shrink or smooth what their leaves span -- dist(object, p) >= (|p - c| - R) - M with the margin M of
 * bound_node, exactly as for the box.  ball[0..2] = c, ball[3] = 1.004 * (R + M) with the box row's
 * allowances (lol_ball_skips: u = best * 1.004 + ball[3]).  Returns the leaf's node index, or -1. */
static void ball_leaves(const lolb200_scene* s, uint32_t idx, const double c[3], double* R, int* ok) {
	const lolb200_object* o = &s->nodes[idx];
	double ext;
	switch (o->type) {
	case LOLB200_OBJ_SPHERE: ext = fabs((double)o->radius); break;
	case LOLB200_OBJ_BOX:
		ext = sqrt((double)o->point2[0] * o->point2[0] + (double)o->point2[1] * o->point2[1] +
		           (double)o->point2[2] * o->point2[2]) + fabs((double)o->radius);
		break;
	case LOLB200_OBJ_PLANE: *ok = 0; return;
	default:
		ball_leaves(s, (uint32_t)o->a, c, R, ok);
		ball_leaves(s, (uint32_t)o->b, c, R, ok);
		return;
	}
	const double d = sqrt(((double)o->point[0] - c[0]) * ((double)o->point[0] - c[0]) +
	                      ((double)o->point[1] - c[1]) * ((double)o->point[1] - c[1]) +
	                      ((double)o->point[2] - c[2]) * ((double)o->point[2] - c[2])) + ext;
	if (!(d <= 1e30))
		*ok = 0;
	if (d > *R)
		*R = d;
}

static void ball_centres(const lolb200_scene* s, uint32_t root, uint32_t idx, double* bestR, int* best_leaf) {
	const lolb200_object* o = &s->nodes[idx];
	if (o->type == LOLB200_OBJ_SPHERE) {
		const double c[3] = {o->point[0], o->point[1], o->point[2]};
		double R = 0;
		int ok = 1;
		ball_leaves(s, root, c, &R, &ok);
		if (ok && R < *bestR) {
			*bestR = R;
			*best_leaf = (int)idx;
		}
	} else if (o->type != LOLB200_OBJ_BOX && o->type != LOLB200_OBJ_PLANE) {
		ball_centres(s, root, (uint32_t)o->a, bestR, best_leaf);
		ball_centres(s, root, (uint32_t)o->b, bestR, best_leaf);
	}
}

static int ball_row(const lolb200_scene* s, uint32_t idx, float ball[4]) {
	double lo[3], hi[3], M, R = INFINITY;
	int leaf = -1;
	if (s->nodes[idx].type == LOLB200_OBJ_SPHERE || !bound_node(s, idx, lo, hi, &M))
		return -1; /* (a lone sphere IS its ball: nothing to skip) */
	ball_centres(s, idx, idx, &R, &leaf);
	if (leaf < 0)
		return -1;
	const lolb200_object* o = &s->nodes[leaf];
	const double l1 = fabs((double)o->point[0]) + fabs((double)o->point[1]) + fabs((double)o->point[2]);
	for (int k = 0; k < 3; k++)
		ball[k] = o->point[k];
	ball[3] = nextafterf((float)((R * 1.002 + 0.002 * (l1 + 1.0) + 1e-6) + (M * 1.002 + 1e-6)) * 1.004f, INFINITY);
	return isfinite(ball[3]) ? leaf : -1;
}

/* ---- two-level pruning: rows sorted along a Morton curve, groups of neighbours ---- */
/* objects per group; options.prune_group, set by lolb200_lower_cuda for the calling thread */
static _Thread_local uint32_t lol_group = 8;
#define LOL_GROUP lol_group
/* pruned table loops as per-lane work lists (emit_sdf_fn; options.loop_worklist) */
static _Thread_local int lol_worklist = 0;
/* cells per axis of the candidate grid (options.grid_cells) */
static _Thread_local int lol_grid_n = 64;
/* options.prune_bounds = 3: straight-line tests are boxes only, no balls (A/B) */
static _Thread_local int lol_no_balls = 0;
/* options.prune_bounds = 4: as 2, and every straight-line object that has a ball is tested with it (tests) */
static _Thread_local int lol_force_balls = 0;
/* emit_sdf_fn writes lol_sdf_nr: the pruned loop with the per-ray candidate memory (lol_kernel.cuh: struct lol_near) */
static _Thread_local int lol_emit_near = 0;
/* Rows that are read with per-lane addresses (work lists, candidate memory) get a stride of 4 x odd
 * words: two rows then fall into different 16-byte bank groups of shared memory unless their numbers
 * agree modulo 8 (Morton neighbours do not).  Every form of a program reads ONE table, so this is
 * decided per program, not per function. */
static _Thread_local int lol_pad_rows = 0;

struct morton_key {
	uint32_t key, idx;
};

static int morton_cmp(const void* a, const void* b) {
	const struct morton_key *x = a, *y = b;
	return x->key < y->key ? -1 : x->key > y->key ? 1 : (x->idx > y->idx) - (x->idx < y->idx);
}

/* order[q] = index of the q-th box along a 30-bit Morton curve through the box
 * centres (unboundable boxes, H = inf, go last). */
static void morton_order(float (*boxes)[LOL_BOUND_SLOTS], uint32_t n, uint32_t* order) {
	struct morton_key* keys = malloc(sizeof *keys * (n ? n : 1));
	float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
	for (uint32_t k = 0; k < n; k++)
		if (isfinite(boxes[k][3]))
			for (int c = 0; c < 3; c++) {
				lo[c] = boxes[k][c] < lo[c] ? boxes[k][c] : lo[c];
				hi[c] = boxes[k][c] > hi[c] ? boxes[k][c] : hi[c];
			}
	for (uint32_t k = 0; k < n; k++) {
		uint32_t key = 0xffffffffu;
		if (isfinite(boxes[k][3])) {
			key = 0;
			for (int c = 0; c < 3; c++) {
				const double span = (double)hi[c] - lo[c];
				uint32_t q = span > 0 ? (uint32_t)(((double)boxes[k][c] - lo[c]) / span * 1023.0) : 0;
				for (int b = 0; b < 10; b++)
					key |= ((q >> b) & 1u) << (3 * b + c);
			}
		}
		keys[k].key = key;
		keys[k].idx = k;
	}
	qsort(keys, n, sizeof *keys, morton_cmp);
	for (uint32_t k = 0; k < n; k++)
		order[k] = keys[k].idx;
	free(keys);
}

/* The box around `count` member boxes (already padded), margin = the largest. */
static void group_box(float (*boxes)[LOL_BOUND_SLOTS], const uint32_t* members, uint32_t count,
                      float out[LOL_BOUND_SLOTS]) {
	double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, m = 0;
	int unbounded = 0;
	for (uint32_t q = 0; q < count; q++) {
		const float* b = boxes[members[q]];
		if (!isfinite(b[3]) || !isfinite(b[4]) || !isfinite(b[5]))
			unbounded = 1;
		for (int c = 0; c < 3; c++) {
			lo[c] = (double)b[c] - b[3 + c] < lo[c] ? (double)b[c] - b[3 + c] : lo[c];
			hi[c] = (double)b[c] + b[3 + c] > hi[c] ? (double)b[c] + b[3 + c] : hi[c];
		}
		m = b[6] > m ? b[6] : m;
	}
	for (int c = 0; c < 3; c++) {
		out[c] = unbounded ? 0.f : (float)((lo[c] + hi[c]) * 0.5);
		/* rounded up: the group box must contain every member box */
		out[3 + c] = unbounded ? INFINITY
		                       : nextafterf((float)((hi[c] - lo[c]) * 0.5 * 1.0001 + 2e-7 * (fabs(lo[c]) + fabs(hi[c]))),
		                                    INFINITY);
	}
	out[6] = (float)m;
}

/* Cost of one object in the FLOP convention of DESIGN.md (sphere 10, round box 20,
 * plane 1, smooth node 13, CSG node 1): what a skipped evaluation saves. */
static unsigned node_cost(const lolb200_scene* s, uint32_t idx) {
	const lolb200_object* o = &s->nodes[idx];
	switch (o->type) {
	case LOLB200_OBJ_SPHERE: return 10;
	case LOLB200_OBJ_BOX: return 20;
	case LOLB200_OBJ_PLANE: return 1;
	case LOLB200_OBJ_SMOOTH_UNION:
		return 13 + node_cost(s, (uint32_t)o->a) + node_cost(s, (uint32_t)o->b);
	default: return 1 + node_cost(s, (uint32_t)o->a) + node_cost(s, (uint32_t)o->b);
	}
}
/* a box test is about 20 instructions: worth it from two spheres' worth of work on */
#define LOL_TEST_PAYS 20u

/* ---- does a box test pay?  a sampled estimate ----------------------------------
 * A test costs about 20 instructions on every evaluation and saves the object only
 * where it fires; whether that is often depends on where rays actually go (scene4:
 * 42 % of all evaluations skip the blob; scene3, whose blob fills the view: almost
 * none, and the tests cost 21 %).  So the lowering marches a coarse frame of the
 * scene's own camera on the CPU -- primary rays and the shadow rays of every hit,
 * the steps render_thread would take -- and counts how often each test would fire.
 * Plain float arithmetic: this steers an optimisation that is exact either way, it
 * need not round like the reference. */
static float est_node(const lolb200_scene* s, uint32_t idx, const float p[3]) {
	const lolb200_object* o = &s->nodes[idx];
	const float q[3] = {p[0] - o->point[0], p[1] - o->point[1], p[2] - o->point[2]};
	switch (o->type) {
	case LOLB200_OBJ_SPHERE: return sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]) - o->radius;
	case LOLB200_OBJ_BOX: {
		const float d[3] = {fabsf(q[0]) - o->point2[0], fabsf(q[1]) - o->point2[1], fabsf(q[2]) - o->point2[2]};
		const float c[3] = {fmaxf(d[0], 0.f), fmaxf(d[1], 0.f), fmaxf(d[2], 0.f)};
		return sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) + fminf(fmaxf(d[0], fmaxf(d[1], d[2])), 0.f) -
		       o->radius;
	}
	case LOLB200_OBJ_PLANE: return q[1];
	default: {
		const float a = est_node(s, (uint32_t)o->a, p), b = est_node(s, (uint32_t)o->b, p);
		if (o->type == LOLB200_OBJ_UNION)
			return fminf(a, b);
		if (o->type == LOLB200_OBJ_INTERSECTION)
			return fmaxf(a, b);
		if (o->type == LOLB200_OBJ_DIFFERENCE)
			return fmaxf(a, -b);
		const float h = fminf(fmaxf(.5f + .5f * (b - a) / o->smoothness, 0.f), 1.f);
		return (b + (a - b) * h) - o->smoothness * h * (1.f - h);
	}
	}
}

static int est_box_skips(const float p[3], const float b[7], float best) {
	const float q[3] = {fmaxf(fabsf(p[0] - b[0]) - b[3], 0.f), fmaxf(fabsf(p[1] - b[1]) - b[4], 0.f),
	                    fmaxf(fabsf(p[2] - b[2]) - b[5], 0.f)};
	const float u = best * 1.004f + b[6];
	return u > 0.f && q[0] * q[0] + q[1] * q[1] + q[2] * q[2] > u * u;
}

static int est_ball_skips(const float p[3], const float b[4], float best) {
	const float q[3] = {p[0] - b[0], p[1] - b[1], p[2] - b[2]};
	const float u = best * 1.004f + b[3];
	return u > 0.f && q[0] * q[0] + q[1] * q[1] + q[2] * q[2] > u * u;
}

struct est {
	const lolb200_scene* s;
	const unsigned char* straight; /* per object: 1 = straight-line code, 2 = straight-line and bounded */
	float (*boxes)[LOL_BOUND_SLOTS];
	const uint32_t* bounded;       /* the bounded straight-line objects */
	uint32_t nb;
	const float* all;              /* the box around them */
	unsigned long points, all_fires, *reached, *fires; /* per bounded object */
	float (*balls)[4];             /* per bounded object: its ball (ball_row), radius slot NaN = none */
	unsigned long* ball_fires;
};

/* sdf() at p as the generated code would run it with every test on, counting. */
static float est_sdf(struct est* e, const float p[3]) {
	const lolb200_scene* s = e->s;
	float best = INFINITY;
	for (uint32_t k = 0; k < s->n_objects; k++) /* the objects without a box come first */
		if (e->straight[k] == 1)
			best = fminf(best, est_node(s, s->objects[k], p));
	e->points++;
	if (est_box_skips(p, e->all, best))
		e->all_fires++;
	for (uint32_t q = 0; q < e->nb; q++) { /* own tests see the running minimum */
		const uint32_t k = e->bounded[q];
		e->reached[q]++;
		if (est_box_skips(p, e->boxes[k], best))
			e->fires[q]++;
		if (e->balls && e->balls[q][3] == e->balls[q][3] && est_ball_skips(p, e->balls[q], best))
			e->ball_fires[q]++;
		best = fminf(best, est_node(s, s->objects[k], p));
	}
	for (uint32_t k = 0; k < s->n_objects; k++) /* table loops */
		if (!e->straight[k])
			best = fminf(best, est_node(s, s->objects[k], p));
	return best;
}

static void est_march(struct est* e) {
	const lolb200_scene* s = e->s;
	const int w = s->n_nodes > 256 ? 24 : 64, h = s->n_nodes > 256 ? 14 : 36;
	lolb200_camera_basis cb;
	lolb200_camera_basis_compute(&s->camera, w, h, &cb);
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++) {
			/* a bounded effort: about 4e7 node evaluations (a fraction of a second), in
			 * an order that still covers the frame when it is cut short */
			if ((double)e->points * (double)s->n_nodes > 4e7)
				return;
			const int px = (x * 7) % w, py = (y * 5) % h; /* 7, 5 coprime to 64x36 and 24x14 */
			const float vx = ((float)px + .5f) / (float)w * 2.f - 1.f, vy = 1.f - ((float)py + .5f) / (float)h * 2.f;
			float rd[3], t = 0.f;
			for (int c = 0; c < 3; c++)
				rd[c] = cb.right[c] * (vx * cb.width) + cb.up[c] * (vy * cb.height) + cb.dir[c];
			const float inv = 1.f / sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
			for (int c = 0; c < 3; c++)
				rd[c] *= inv;
			for (int i = 0; i < 256; i++) { /* get_intersection, naive_renderer.c:47-69 */
				const float p[3] = {cb.origin[0] + rd[0] * t, cb.origin[1] + rd[1] * t, cb.origin[2] + rd[2] * t};
				const float d = est_sdf(e, p);
				t += d;
				if (!(d >= 0.001f) || !(t <= 100.f))
					break;
			}
			if (!(t < 100.f))
				continue;
			const float hit[3] = {cb.origin[0] + rd[0] * t, cb.origin[1] + rd[1] * t, cb.origin[2] + rd[2] * t};
			for (uint32_t l = 0; l < s->n_lights; l++) { /* softshadow, naive_renderer.c:72-100 */
				float dir[3] = {s->lights[l].point[0] - hit[0], s->lights[l].point[1] - hit[1],
				                s->lights[l].point[2] - hit[2]};
				const float dist = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
				float st = 0.f, res = 1.f;
				if (!(dist > 0.f))
					continue;
				for (int c = 0; c < 3; c++)
					dir[c] /= dist;
				for (int i = 0; i < 128; i++) {
					const float p[3] = {hit[0] + dir[0] * (1.f + st), hit[1] + dir[1] * (1.f + st),
					                    hit[2] + dir[2] * (1.f + st)};
					const float d = est_sdf(e, p);
					if (st > 0.f)
						res = fminf(res, 50.f * d / st);
					st += d;
					if (!(res > 0.f) || !(st <= dist) || !(d == d))
						break;
				}
			}
		}
}

/* `lol_box_skips(x, y, z, <box as immediates>, best...)` */
static void emit_box_test(struct cgen* g, const float box[7], int two) {
	sb_printf(g->out, "lol_box_skips%s(x, y, z", two ? "2" : "");
	for (int q = 0; q < 7; q++) {
		sb_printf(g->out, ", ");
		cst(g, box[q]);
	}
	sb_printf(g->out, two ? ", bestA, bestB)" : ", best)");
}

/* One straight-line top-level object: `{ evaluate; update the running minimum }`.
 * tie_aware: the update also takes an equal distance from a smaller id (needed when
 * something with a larger id may have been evaluated before).  own_box: the object
 * is skipped when its bounding box proves it cannot win. */
static void emit_straight_object(struct sb* body, struct cgen* g, uint32_t k, const char* sig,
                                 int two, int tie_aware, const float* own_box, const char* tabs,
                                 const float* own_ball, int ball_leaf) {
	char ind[16];
	snprintf(ind, sizeof ind, "%s\t", tabs);
	g->indent = ind;
	g->tmp = 0;
	sb_printf(body, "%s{ // object %u: %s\n", tabs, k + 1, sig);
	if (own_ball && !two) {
		/* the squared distance to one of the object's own sphere centres: the object computes the very same
		 * value (same operations, same order: the compiler keeps one copy), so the test costs four instructions */
		sb_printf(body,
		          "%s\t// every leaf of the object lies within R of this sphere's centre: dist(object, p) >= |p - c| - R - M >= best\n"
		          "%s\tconst float bqx = ", tabs, tabs);
		coord_minus(g, "x", (uint32_t)ball_leaf, F_PX);
		sb_printf(body, ", bqy = ");
		coord_minus(g, "y", (uint32_t)ball_leaf, F_PY);
		sb_printf(body, ", bqz = ");
		coord_minus(g, "z", (uint32_t)ball_leaf, F_PZ);
		sb_printf(body, ";\n%s\tif (!lol_ball_skips(lol_dot(bqx, bqy, bqz, bqx, bqy, bqz), ", tabs);
		cst(g, own_ball[3]);
		sb_printf(body, ", best)) {\n");
		snprintf(ind, sizeof ind, "%s\t\t", tabs);
		own_box = own_ball; /* (closes like a box test below) */
	} else if (own_box) {
		sb_printf(body, "%s\t// dist(object, p) >= dbox(p) - M >= best: cannot win\n%s\tif (!", tabs, tabs);
		emit_box_test(g, own_box, two);
		sb_printf(body, ") {\n");
		snprintf(ind, sizeof ind, "%s\t\t", tabs);
	}
	int t = emit_object(g, g->s->objects[k]);
	if (two) {
		char tA[64] = "", tB[64] = "";
		if (tie_aware) {
			snprintf(tA, sizeof tA, " || (a_ == bestA && %uu < bidA)", k + 1);
			snprintf(tB, sizeof tB, " || (b_ == bestB && %uu < bidB)", k + 1);
		}
		sb_printf(body,
		          "%sconst float a_ = lol_lo(t%d), b_ = lol_hi(t%d);\n"
		          "%sif (a_ < bestA%s) {\n%s\tbestA = a_;\n%s\tbidA = %uu;\n%s}\n"
		          "%sif (b_ < bestB%s) {\n%s\tbestB = b_;\n%s\tbidB = %uu;\n%s}\n",
		          ind, t, t, ind, tA, ind, ind, k + 1, ind, ind, tB, ind, ind, k + 1, ind);
	} else if (g->first_free && g->fast && g->div_ok && !own_box && !own_ball && tabs[1] == '\0')
		/* The first object of the guarded form, outside every test: `best` is still +INF, and the
		 * result of this function only counts when its range guard passes -- coordinates and every
		 * scene constant at most 2^60, no NaN -- where an object's distance is finite (sums,
		 * products and square roots of values below 2^123; quotients by a proved constant), so
		 * `t < +INF` holds.  The update is unconditional, and the compiler then knows `bid`. */
		sb_printf(body, "%sbest = t%d; // first object: t < +INF under the range guard\n%sbid = %uu;\n", ind, t, ind, k + 1);
	else if (tie_aware)
		sb_printf(body, "%sif (t%d < best || (t%d == best && %uu < bid)) {\n%s\tbest = t%d;\n%s\tbid = %uu;\n%s}\n",
		          ind, t, t, k + 1, ind, t, ind, k + 1, ind);
	else
		sb_printf(body, "%sif (t%d < best) {\n%s\tbest = t%d;\n%s\tbid = %uu;\n%s}\n", ind, t, ind, t, ind,
		          k + 1, ind);
	if (own_box)
		sb_printf(body, "%s\t} else\n%s\t\tlol_count_skip(%uu);\n", tabs, tabs,
		          (node_cost(g->s, g->s->objects[k]) + 1u) * (two ? 2u : 1u));
	sb_printf(body, "%s}\n", tabs);
	g->first_free = 0;
}

/* The decisions of the sampled estimate, kept for the other distance functions of
 * the same program (reference form, guarded form, two-ray form): one march each for
 * the IEEE and the guarded cost model instead of one per function. */
struct est_memo {
	int valid[2], wrap[2];
	unsigned char* own[2];
};

/* ---- the constant tables of table loops: ONE array of words ---------------------
 * Every table (rows, group boxes, id -> row) is a range of lol_tables[]; its name
 * is a macro `(LOL_TAB + offset)`.  LOL_TAB is the array itself (__constant__, or
 * global memory above 56 KB), or -- on the GPU, when it fits -- a copy in SHARED
 * memory that every CTA makes once (lol_kernel.cuh): the loops read their rows with
 * register-indexed loads, which cost a constant-cache lookup each (ncu: short-
 * scoreboard stalls 4.3 per issue, 20 % lookup misses on the 31 KB table of the
 * 1024-sphere scene) and a plain LDS from shared memory. */
struct tabs {
	struct sb words, defs;
	size_t n;
};

static void tab_start(struct tabs* T, const char* fmt, int run_no) {
	char name[64];
	while (T->n % 4) { /* 16-byte aligned: vector loads */
		sb_printf(&T->words, "0u, ");
		T->n++;
	}
	snprintf(name, sizeof name, fmt, run_no);
	sb_printf(&T->defs, "#define %s (LOL_TAB + %zu)\n#define %s_offset %zuu\n", name, T->n, name, T->n);
	sb_printf(&T->words, "\n\t/* %s */\n", name);
}

static void tab_float(struct tabs* T, float f) {
	sb_bits(&T->words, f);
	sb_printf(&T->words, ", ");
	T->n++;
}

static void tab_u32(struct tabs* T, uint32_t v) {
	sb_printf(&T->words, "%uu, ", v);
	T->n++;
}

/* One distance function.  fast = 0: the reference form (IEEE sqrt and division as
 * the compiler emits them), named `name`.  fast = 1: the guarded form, which
 * runs the same arithmetic without the per-operation special-case branches and
 * hands the whole evaluation to `fallback` when its one range check fails. */
static void emit_sdf_fn(struct sb* out, struct sb* tables_out, const lolb200_scene* s,
                        int loop_threshold, const char* name, const char* attrs, int fast,
                        int div_ok, const char* fallback, int prune, int two, int smem_ok,
                        struct est_memo* memo, int pack, int lay_pack, int lay_div_ok) {
	struct sb body = {0};
	struct tabs tables = {{0}, {0}, 0};
	struct pairc pc = {0};
	/* pack_pairs = 1 packs where it was measured to pay: inside table loops (B200, 4K: 1024
	 * spheres 48.6 -> 45.9 ms, CSG 45.8 -> 43.5 ms).  Straight-line scenes lose a few
	 * percent (scene4 2.15 -> 2.26 ms with 13 % fewer instructions: the packed instructions
	 * keep the FMA pipe as busy as the scalar ones did and their dependent chains issue
	 * at 0.3-0.45 per clock, profiles/r01_ubench_f32x2.txt), so they pack only on request. */
	struct cgen g = {.s = s, .out = &body, .fast = fast, .div_ok = div_ok, .two = two, .pack = pack >= 2 ? pack : 0, .pc = &pc,
		                 .lay_pack = lay_pack, .lay_div_ok = lay_div_ok};
	char** sigs = calloc(s->n_objects ? s->n_objects : 1, sizeof *sigs);
	int run_no = 0;

	for (uint32_t i = 0; i < s->n_objects; i++) {
		struct sb sig = {0};
		signature(s, s->objects[i], &sig);
		sigs[i] = sig.p;
	}

	/* The out-of-line fallback hands (distance, id) back in one 64-bit register
	 * pair; a reference parameter would force the caller's id onto the stack. */
	const int packed_ret = strcmp(name, "lol_sdf") != 0 && strcmp(name, "lol_sdf_nr") != 0 && !two; /* (distance, id) in one 64-bit value */
	/* the guarded single-ray function comes in two pieces: the arithmetic (lol_sdf_try) and the fall-back around it */
	const int split_guard = fast && !two && !packed_ret && !lol_emit_near;
	sb_printf(&body,
	          "// sdf (naive_renderer.c:30-44): running strict-< minimum over the top-level\n"
	          "// objects, ids 1..n in file order, (INF, 0) when nothing is closer.\n");
	if (two)
		sb_printf(&body,
		          "// Two rays per call: x, y, z hold ray A in the low and ray B in the high half.\n"
		          "// hintA / hintB: object ids worth evaluating first (the last winners), 0 = none.\n"
		          "__device__ %s lol_f2 %s(const lol_f2 x, const lol_f2 y, const lol_f2 z,\n"
		          "                                         const lol_u32 hintA, const lol_u32 hintB,\n"
		          "                                         lol_u32& idA, lol_u32& idB) {\n"
		          "\t(void)hintA;\n\t(void)hintB;\n"
		          "\tfloat bestA = LOL_INF, bestB = LOL_INF;\n\tlol_u32 bidA = 0u, bidB = 0u;\n"
		          "\t// One range guard per evaluation, over both rays.\n"
		          "\tfloat lo = LOL_COORD_MAX - lol_max_abs_halves(x, y, z);\n",
		          attrs, name);
	else if (packed_ret)
		sb_printf(&body, "__device__ %s lol_u64 %s(const float x, const float y, const float z) {\n",
		          attrs, name);
	else if (lol_emit_near)
		sb_printf(&body,
		          "// nr: what the ray remembers of the pruned table loop (struct lol_near); move: an upper bound of\n"
		          "// the distance between this point and the one of the ray's previous call.  Neither changes the result.\n"
		          "__device__ %s float %s(const float x, const float y, const float z,\n"
		          "                                         lol_near& nr, const float move, lol_u32& id) {\n"
		          "\tconst lol_u32 hint = 0u;\n",
		          attrs, name);
	else if (split_guard)
		sb_printf(&body,
		          "// The guarded form WITHOUT its fall-back: the caller is told whether the guard passed (`ok`), and\n"
		          "// the result only counts when it did.  lol_sdf() below is the complete function; the march\n"
		          "// loops of variant 1 call this one and leave the loop when the guard fails (LOL_GUARD_OUT), so\n"
		          "// the out-of-line call and its reconvergence point are not part of every step.\n"
		          "// hint: an object id worth evaluating first (the ray's last winner), 0 = none; only\n"
		          "// table loops use it, and it never changes the result.\n"
		          "__device__ %s float %s_try(const float x, const float y, const float z,\n"
		          "                                         const lol_u32 hint, lol_u32& id, bool& ok) {\n",
		          attrs, name);
	else
		sb_printf(&body,
		          "// hint: an object id worth evaluating first (the ray's last winner), 0 = none; only\n"
		          "// table loops use it, and it never changes the result.\n"
		          "__device__ %s float %s(const float x, const float y, const float z,\n"
		          "                                         const lol_u32 hint, lol_u32& id) {\n",
		          attrs, name);
	if (!two && !packed_ret)
		sb_printf(&body, "\t(void)hint;\n");
	if (!two)
		sb_printf(&body, "\tfloat best = LOL_INF;\n\tlol_u32 bid = 0u;\n");
	if (fast && !two)
		sb_printf(&body,
		          "\t// One range guard per evaluation: lo = min(every sqrt argument, 2^60 - max|p|), with a minimum\n"
		          "\t// and a maximum that hand a NaN on (min.NaN / max.NaN): a NaN or infinite coordinate fails\n"
		          "\t// the guard, so inside it every distance below is a finite number.\n"
		          "\tfloat lo = LOL_COORD_MAX - lol_max_nan(lol_max_nan(fabsf(x), fabsf(y)), fabsf(z));\n");

	/* Segments: maximal runs of same-shaped neighbours.  Long runs become table
	 * loops.  With pruning the straight-line objects are evaluated FIRST (they
	 * tighten `best`, which is what every box test compares against); the update
	 * rules then break ties by object id, so the result is the reference's "first of
	 * equal distances" whatever the order. */
	int straight_in_file_order = !prune;
	if (prune) {
		/* Straight-line objects, reordered: first the ones without a box (planes: one
		 * subtraction, and they give `best` a finite value), then the bounded ones --
		 * all of them behind ONE test of the box around them when that pays, and each
		 * expensive one behind its own.  On scene4 the five-sphere blob (105 FLOP) is
		 * skipped in 42 % of all evaluations: wherever a ray is nearer to the floor
		 * than to the blob's box (simulated on the march points of a 192x108 frame). */
		const uint32_t no = s->n_objects;
		unsigned char* straight = calloc(no ? no : 1, 1);
		float(*boxes)[LOL_BOUND_SLOTS] = malloc(sizeof *boxes * (no ? no : 1));
		uint32_t* bounded = malloc(sizeof *bounded * (no ? no : 1));
		uint32_t nb = 0;
		unsigned total = 0;
		for (uint32_t i = 0; i < no;) {
			uint32_t j = i + 1;
			while (j < no && strcmp(sigs[j], sigs[i]) == 0)
				j++;
			if ((int)(j - i) < loop_threshold)
				memset(straight + i, 1, j - i);
			i = j;
		}
		for (uint32_t k = 0; k < no; k++) {
			if (!straight[k])
				continue;
			bound_row(s, s->objects[k], boxes[k]);
			if (isfinite(boxes[k][3]) && isfinite(boxes[k][4]) && isfinite(boxes[k][5])) {
				bounded[nb++] = k;
				straight[k] = 2;
				total += node_cost(s, s->objects[k]);
			}
		}
		struct sb reordered = {0}; /* kept only if some test is switched on */
		struct sb* const real_body = g.out;
		int any_test = 0;
		g.out = &reordered;
		g.first_free = fast && !two;
		for (uint32_t k = 0; k < no; k++)
			if (straight[k] == 1)
				emit_straight_object(&reordered, &g, k, sigs[k], two, 0, NULL, "\t", NULL, -1);
		if (nb) {
			float all[LOL_BOUND_SLOTS];
			unsigned char* own = calloc(nb, 1); /* per bounded object: 0 no test, 1 its box, 2 its ball (ball_row) */
			float(*balls)[4] = malloc(sizeof *balls * nb);
			int* ball_leaf = malloc(sizeof *ball_leaf * nb);
			int wrap = 0;
			group_box(boxes, bounded, nb, all);
			for (uint32_t q = 0; q < nb; q++) {
				ball_leaf[q] = (two || (prune >= 2 && !lol_force_balls) || lol_no_balls) ? -1 : ball_row(s, s->objects[bounded[q]], balls[q]);
				if (ball_leaf[q] < 0)
					balls[q][0] = balls[q][1] = balls[q][2] = balls[q][3] = NAN;
			}
			if (total >= LOL_TEST_PAYS && memo->valid[fast != 0]) {
				wrap = memo->wrap[fast != 0];
				memcpy(own, memo->own[fast != 0], nb);
			} else if (total >= LOL_TEST_PAYS) {
				/* instructions saved where a test fires (about 1.5 per FLOP of the convention),
				 * discounted because a warp only saves what ALL its lanes skip, against the
				 * ~20 instructions the test costs everywhere else */
				struct est e = {.s = s, .straight = straight, .boxes = boxes, .bounded = bounded, .nb = nb, .all = all,
				                .balls = balls};
				double inside = 0, ball_gain0 = 0;
				e.reached = calloc(nb, sizeof *e.reached);
				e.fires = calloc(nb, sizeof *e.fires);
				e.ball_fires = calloc(nb, sizeof *e.ball_fires);
				est_march(&e);
				/* model (calibrated on B200 against the four example scenes): a test is 24
				 * instructions, reordering costs a tie-aware update (3) per bounded object, and
				 * only 70 % of what single lanes could skip is skipped by whole warps */
				for (uint32_t q = 0; q < nb; q++) {
					/* instructions per FLOP of the convention: the IEEE forms carry their special-case code */
					const double saves = (fast ? 1.5 : 2.3) * node_cost(s, s->objects[bounded[q]]);
					const double rate = e.reached[q] ? (double)e.fires[q] / (double)e.reached[q] : 0.0;
					own[q] = nb >= 2 && (prune >= 2 ? node_cost(s, s->objects[bounded[q]]) >= LOL_TEST_PAYS
					                               : 0.7 * rate * saves > 24.0 + 3.0); /* the test is paid on every evaluation */
					{
						/* the ball: 4 instructions + 3 for the reordering; where it fires the object is saved but for
						 * the squared distance that went in front of the test (about 12 instructions) */
						const double brate = e.reached[q] ? (double)e.ball_fires[q] / (double)e.reached[q] : 0.0;
						const double ball_gain = ball_leaf[q] >= 0 ? 0.7 * brate * (saves - 12.0) - (4.0 + 3.0) : -1.0;
						const double box_gain = 0.7 * rate * saves - (24.0 + 3.0);
						if (getenv("LOLB200_DEBUG_EST"))
							fprintf(stderr, "lolb200 estimate: object %u: its ball fires %.1f %% (gain %.1f, its box %.1f)\n",
							        bounded[q] + 1, brate * 100, ball_gain, box_gain);
						if (prune == 1 && ball_gain > 0.0 && ball_gain > box_gain && (nb >= 2 || q == 0)) {
							if (nb >= 2)
								own[q] = 2;
							else
								ball_gain0 = ball_gain; /* one bounded object: against the box around "all" below */
						}
					}
					inside += saves + (own[q] == 1 ? 24.0 : own[q] == 2 ? 4.0 : 0.0);
					if (getenv("LOLB200_DEBUG_EST"))
						fprintf(stderr, "lolb200 estimate: object %u: own box test fires %.1f %% (saves %.0f) -> %s\n",
						        bounded[q] + 1, rate * 100, saves, own[q] == 1 ? "on" : own[q] == 2 ? "its ball instead" : "off");
				}
				{
					const double rate = e.points ? (double)e.all_fires / (double)e.points : 0.0;
					wrap = prune >= 2 || 0.7 * rate * inside > 24.0 + 3.0 * nb;
					if (nb == 1 && ball_gain0 > 0.0 && ball_gain0 > 0.7 * rate * inside - (24.0 + 3.0)) {
						wrap = 0; /* the one bounded object behind its ball instead of its box */
						own[0] = 2;
					}
					if (getenv("LOLB200_DEBUG_EST"))
						fprintf(stderr, "lolb200 estimate: %lu points, the box around all %u fires %.1f %% (saves %.0f) -> %s\n",
						        e.points, nb, rate * 100, inside, wrap ? "on" : "off");
				}
				free(e.reached);
				free(e.fires);
				free(e.ball_fires);
				memo->valid[fast != 0] = 1;
				memo->wrap[fast != 0] = wrap;
				memo->own[fast != 0] = malloc(nb);
				memcpy(memo->own[fast != 0], own, nb);
			}
			if (lol_force_balls) /* every object that has a ball behind it, whatever an estimate would say */
				for (uint32_t q = 0; q < nb; q++)
					if (ball_leaf[q] >= 0)
						own[q] = 2;
			if (wrap) {
				sb_printf(&reordered, "\t// none of the %u bounded objects can win: dist >= dbox(p) - M >= best\n\tif (!", nb);
				emit_box_test(&g, all, two);
				sb_printf(&reordered, ") {\n");
			}
			for (uint32_t q = 0; q < nb; q++) {
				const uint32_t k = bounded[q];
				/* (the two-ray form has no ball test: where the single-ray form chose the ball, it keeps the box) */
				const int ball = own[q] == 2 && ball_leaf[q] >= 0 && !two;
				emit_straight_object(&reordered, &g, k, sigs[k], two, 1, (own[q] == 1 || (own[q] == 2 && !ball)) ? boxes[k] : NULL,
				                     wrap ? "\t\t" : "\t", ball ? balls[q] : NULL, ball_leaf[q]);
				any_test |= own[q];
			}
			if (wrap)
				sb_printf(&reordered, "\t} else\n\t\tlol_count_skip(%uu);\n", (total + nb) * (two ? 2u : 1u));
			any_test |= wrap;
			free(own);
			free(balls);
			free(ball_leaf);
		}
		g.out = real_body;
		if (any_test)
			sb_putn(&body, reordered.p, reordered.len);
		else
			straight_in_file_order = 1; /* nothing pays: the plain code, no reordering, no tie rule */
		free(reordered.p);

		free(straight);
		free(boxes);
		free(bounded);
	}
	g.first_free = fast && !two && straight_in_file_order;
	for (int pass = 0; pass < 2; pass++)
	for (uint32_t i = 0; i < s->n_objects;) {
		uint32_t j = i + 1;
		while (j < s->n_objects && strcmp(sigs[j], sigs[i]) == 0)
			j++;
		const int is_loop = (int)(j - i) >= loop_threshold;
		/* without pruning: everything in file order (pass 0); with it the straight-line
		 * objects are already out (above) and pass 1 adds the loops */
		const int now = prune ? (pass == 0 ? (!is_loop && straight_in_file_order) : is_loop) : (pass == 0);
		if (!now) {
			i = j;
			continue;
		}
		if (is_loop) {
			/* objects i .. j-1 share one shape: loop over a parameter table */
			size_t per_row = 0;
			g.first_free = 0;
			const uint32_t n = j - i;
			sb_printf(&body, "\t// objects %u..%u: %u x %s\n", i + 1, j, n, sigs[i]);
			if (!prune) {
				/* plain loop in file order; row r belongs to object id i + 1 + r */
				tab_start(&tables, "lol_run%d", run_no);
				for (uint32_t k = i; k < j; k++) {
					struct sb scratch = {0};
					struct cgen r = {.s = s, .out = (k == i) ? &body : &scratch, .in_loop = 1,
					                 .indent = "\t\t", .fast = fast, .div_ok = div_ok, .two = two, .pack = pack, .pc = &pc,
		                 .lay_pack = lay_pack, .lay_div_ok = lay_div_ok};
					if (k == i)
						sb_printf(&body,
						          "\t{\n#pragma unroll 1\n\tfor (int i = 0; i < %u; ++i) {\n"
						          "\t\tconst lol_u32* c = lol_run%d + i * LOL_RUN%d_STRIDE;\n",
						          n, run_no, run_no);
					int t = emit_object(&r, s->objects[k]);
					if (k == i && two)
						sb_printf(&body,
						          "\t\tconst float a_ = lol_lo(t%d), b_ = lol_hi(t%d);\n"
						          "\t\tif (a_ < bestA) {\n\t\t\tbestA = a_;\n\t\t\tbidA = %uu + (lol_u32)i;\n\t\t}\n"
						          "\t\tif (b_ < bestB) {\n\t\t\tbestB = b_;\n\t\t\tbidB = %uu + (lol_u32)i;\n\t\t}\n\t}\n\t}\n",
						          t, t, i + 1, i + 1);
					else if (k == i)
						sb_printf(&body,
						          "\t\tif (t%d < best) {\n\t\t\tbest = t%d;\n\t\t\tbid = %uu + "
						          "(lol_u32)i;\n\t\t}\n\t}\n\t}\n",
						          t, t, i + 1);
					if (k == i)
						per_row = r.nrow;
					sb_printf(&tables.words, "\n\t");
					for (size_t q = 0; q < r.nrow; q++)
						tab_float(&tables, r.row[q]);
					cgen_release(&r);
					free(scratch.p);
				}
				sb_printf(&tables.defs, "#define LOL_RUN%d_STRIDE %zu\n", run_no, per_row);
				run_no++;
			} else {
				/* Pruned loop.  The rows are sorted along a Morton curve through their box
				 * centres and cut into groups of LOL_GROUP neighbours; a group whose
				 * box cannot win is skipped with ONE test, and inside a surviving group
				 * every object has its own test.  The ray's last winner (hint) is
				 * evaluated before anything else, so `best` is tight from the start.
				 * Any order is legal: ties are broken by object id, the reference's
				 * "first of equal distances" (naive_renderer.c:39). */
				const int use_hint = !packed_ret;
				/* Work lists (single ray, the program's main form): see the comment at the
				 * generated code below.  128 rows per block, LOL_GROUP must divide it. */
				const int worklist = lol_worklist && !two && !packed_ret && 128u % LOL_GROUP == 0u;
				float(*boxes)[LOL_BOUND_SLOTS] = malloc(sizeof *boxes * n);
				uint32_t* order = malloc(sizeof *order * n);
				uint32_t* rowof = malloc(sizeof *rowof * n);
				const uint32_t ngroups = (n + LOL_GROUP - 1) / LOL_GROUP;
				for (uint32_t k = 0; k < n; k++)
					bound_row(s, s->objects[i + k], boxes[k]);
				morton_order(boxes, n, order);
				for (uint32_t q = 0; q < n; q++)
					rowof[order[q]] = q;
				/* tables: groups, id -> row, rows */
				tab_start(&tables, "lol_run%d_groups", run_no);
				for (uint32_t g0 = 0; g0 < n; g0 += LOL_GROUP) {
					float gb[LOL_BOUND_SLOTS];
					group_box(boxes, order + g0, g0 + LOL_GROUP <= n ? LOL_GROUP : n - g0, gb);
					sb_printf(&tables.words, "\n\t");
					for (int q = 0; q < LOL_BOUND_SLOTS; q++)
						tab_float(&tables, gb[q]);
				}
				tab_start(&tables, "lol_run%d_rowof", run_no);
				for (uint32_t k = 0; k < n; k++) {
					if (k % 16 == 0)
						sb_printf(&tables.words, "\n\t");
					tab_u32(&tables, rowof[k]);
				}
				tab_start(&tables, "lol_run%d", run_no);

				for (uint32_t q = 0; q < n; q++) {
					const uint32_t k = i + order[q];
					struct sb scratch = {0};
					struct cgen r = {.s = s, .out = (q == 0) ? &body : &scratch, .in_loop = 1,
					                 .indent = "\t\t\t", .fast = fast, .div_ok = div_ok, .two = two, .pack = pack, .pc = &pc,
		                 .lay_pack = lay_pack, .lay_div_ok = lay_div_ok};
				const char* best_args = two ? "bestA, bestB" : "best";
				const char* sfx = two ? "2" : "";
				if (lol_emit_near && !two && !packed_ret) {
					/* The per-ray candidate memory (lol_kernel.cuh: struct lol_near).  The remembered
					 * candidates are evaluated first (the last winner untested, the others behind their
					 * box test).  If the room the ray had is used up, every row is looked at again --
					 * tests only, out of line (lol_near_collect) -- which gives the rows that cannot be
					 * skipped now and the distance to the nearest skipped one; the new candidates that
					 * were not evaluated yet are evaluated by the SAME copy of the object's code.  More
					 * than four survivors (a ray's first call knows only the floor's distance, or a
					 * crowd): the evaluation goes the long way (lol_sdf_slow) and the rows are looked at
					 * once more with its exact result, which leaves the few that really matter. */
					const unsigned cost = node_cost(s, s->objects[i]) + 1u;
					struct sb* const real_out = r.out;
					struct sb sink = {0};
					if (q == 0) {
						sb_printf(&body,
						          "\t{\n"
						          "\tLOL_NEAR_STAT(0, 1);\n"
						          "\tnr.room -= lol_fma(move, LOL_F(0x3f800347 /*1.0001*/), LOL_NEAR_PAD);\n"
						          "\tconst bool near_ok = fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z)) <= LOL_NEAR_COORD;\n"
						          "\tconst lol_u32 list0 = nr.cand;\n"
						          "\tlol_u64 list = 0xffffffff00000000ull | list0;  // rows to evaluate in this round (0xff ends the list)\n"
						          "\tbool wide = false;     // the look found five to eight rows: all are evaluated, none remembered\n"
						          "\tlol_u32 wrow = 0xffu;  // the row that holds `best`, if a row does\n"
						          "\tlol_u32 nev = 0u;\n"
						          "\tbool retest = false; // round 0: the last winner is evaluated untested, the others re-tested\n"
						          "\tbool slow = false;\n"
						          "\tfor (int round = 0;; ++round) {\n"
						          "#pragma unroll 1\n"
						          "\t\tfor (int k = 0; k < 8; ++k) {\n"
						          "\t\t\tconst lol_u32 row = (lol_u32)(list >> (8 * k)) & 0xffu;\n"
						          "\t\t\tif (row == 0xffu)\n\t\t\t\tbreak;\n"
						          "\t\t\tconst lol_u32* c = lol_run%d + row * LOL_RUN%d_STRIDE;\n"
						          "\t\t\tif (retest && lol_box_skips(x, y, z, LOL_TF(c[0]), LOL_TF(c[1]), LOL_TF(c[2]), LOL_TF(c[3]), "
						          "LOL_TF(c[4]), LOL_TF(c[5]), LOL_TF(c[6]), best))\n\t\t\t\tcontinue;\n"
						          "\t\t\tretest = round == 0 || wide;\n"
						          "\t\t\t++nev;\n",
						          run_no, run_no);
					}
					/* row = box (C, H, M), object id, then the object's own constants */
					r.out = &sink;
					for (int b = 0; b < LOL_BOUND_SLOTS; b++)
						cst(&r, boxes[order[q]][b]);
					cst_raw(&r, k + 1);
					r.out = real_out;
					free(sink.p);
					if (q == 0)
						sb_printf(&body, "\t\t\tconst lol_u32 oid = c[%d];\n", LOL_BOUND_SLOTS);
					int t = emit_object(&r, s->objects[k]);
					if (q == 0)
						sb_printf(&body,
						          "\t\t\tif (t%d < best || (t%d == best && oid < bid)) {\n"
						          "\t\t\t\tbest = t%d;\n\t\t\t\tbid = oid;\n\t\t\t\twrow = row;\n\t\t\t}\n\t\t}\n"
						          "\t\tif (round == 1)\n\t\t\tbreak;\n"
						          "\t\t// every row outside the candidates still fails its box test here: done\n"
						          "\t\tconst bool need = !(near_ok && nr.room > LOL_F(0x3f808312 /*1.004*/) * fabsf(best));\n"
						          "#if LOL_NEAR >= 2\n"
						          "\t\t// The warp walks the rows whenever ONE of its lanes has to: the others look again too (a look is\n"
						          "\t\t// legal at any time), which costs the warp nothing and renews their room.\n"
						          "\t\tif (!lol_any(need))\n\t\t\tbreak;\n"
						          "#else\n"
						          "\t\tif (!need)\n\t\t\tbreak;\n"
						          "#endif\n"
						          "\t\t// look at every row again (tests only, out of line)\n"
						          "\t\tLOL_NEAR_STAT(1, 1);\n"
						          "\t\tconst lol_look seen = lol_near_collect(x, y, z, best);\n"
						          "\t\tif (!need && seen.n > 4u)\n\t\t\tbreak; // a look the ray did not need found more than four rows: what it knew still holds\n"
						          "\t\tif (seen.n > LOL_NEAR_WIDE) { // more rows cannot be skipped than a look lists: the plain loop, out of line\n"
						          "\t\t\tslow = true;\n\t\t\tbreak;\n\t\t}\n"
						          "\t\twide = seen.n > 4u;\n"
						          "\t\tif (!wide) {\n\t\t\tnr.cand = seen.lo;\n\t\t\tnr.room = seen.room;\n\t\t}\n"
						          "\t\t// the rows of the look that round 0 did not evaluate\n"
						          "\t\tconst lol_u64 rows = (lol_u64)seen.lo | ((lol_u64)seen.hi << 32);\n"
						          "\t\tlist = ~0ull;\n"
						          "\t\tlol_u32 nl = 0u;\n"
						          "#pragma unroll 1\n"
						          "\t\tfor (lol_u32 k = 0; k < seen.n; ++k) {\n"
						          "\t\t\tconst lol_u32 row = (lol_u32)(rows >> (8u * k)) & 0xffu;\n"
						          "\t\t\tif (!lol_near_has(list0, row)) {\n"
						          "\t\t\t\tlist = (list & ~(0xffull << (8u * nl))) | ((lol_u64)row << (8u * nl));\n"
						          "\t\t\t\t++nl;\n\t\t\t}\n\t\t}\n"
						          "\t\tif (nl == 0u)\n\t\t\tbreak;\n"
						          "\t\tretest = wide; // (many rows: `best` tightens as they are evaluated, and a row is tested again before it is)\n"
						          "\t}\n"
						          "\tLOL_NEAR_STAT(3, nev);\n"
						          "\tif (slow) {\n"
						          "\t\tLOL_NEAR_STAT(2, 1);\n"
						          "\t\t// many rows matter here (open space beside a crowd): the plain pruned loop handles that best --\n"
						          "\t\t// it tightens `best` as it goes.  The ray keeps the winner and looks again at its next point.\n"
						          "\t\tconst lol_u64 r = lol_sdf_slow(x, y, z, bid);\n"
						          "\t\tid = (lol_u32)(r >> 32);\n"
						          "\t\tnr.room = -LOL_INF;\n"
						          "\t\tnr.cand = (id >= %uu && id < %uu) ? (0xffffff00u | lol_run%d_rowof[id - %uu]) : 0xffffffffu;\n"
						          "\t\treturn __uint_as_float((lol_u32)r);\n"
						          "\t}\n"
						          "\tif (wide) { // nothing is remembered but the winner: the ray looks again at its next point\n"
						          "\t\tnr.room = -LOL_INF;\n"
						          "\t\tnr.cand = 0xffffff00u | wrow; // (0xff: no row holds `best`)\n"
						          "\t} else if (wrow != 0xffu)\n\t\tnr.cand = lol_near_front(nr.cand, wrow);\n"
						          "\tlol_count_skip((%uu - nev) * %uu);\n"
						          "\t}\n",
						          t, t, t, i + 1, j + 1, run_no, i + 1, n, cost);
				} else if (worklist) {
					/* Per-lane work lists.  The plain loops below walk groups and members with a
					 * warp-uniform index: every lane tests, and the warp then evaluates the UNION of
					 * what its lanes could not skip (21 of 32 lanes per instruction on the
					 * 1024-sphere scene).  Here every lane first COLLECTS the rows it cannot skip as
					 * bits (group tests, member tests: cheap, uniform), then the warp drains the lists
					 * together: in each round every lane evaluates ITS next row -- different rows in
					 * one instruction stream, read from shared memory with per-lane addresses -- so
					 * the rounds number the longest list, not the union.  Any order is legal (ties go
					 * to the smaller object id); a row is re-tested against the current `best` when
					 * it is popped, because earlier rows may have tightened it. */
					const uint32_t nblk = (n + 127u) / 128u, gpb = 128u / LOL_GROUP;
					const unsigned cost = node_cost(s, s->objects[i]) + 1u;
					struct sb* const real_out = r.out;
					struct sb sink = {0};
					if (q == 0) {
						sb_printf(&body, "\t{\n");
						if (use_hint)
							sb_printf(&body,
							          "\tconst int hrow = (hint >= %uu && hint < %uu) ? (int)lol_run%d_rowof[hint - %uu] : -1;\n",
							          i + 1, j + 1, run_no, i + 1);
						else
							sb_printf(&body, "\tconst int hrow = -1;\n");
						sb_printf(&body,
						          "\tlol_u32 m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u; // the rows of the current block this lane cannot skip\n"
						          "\tint blk = -1, i = hrow;\n"
						          "\tbool found = hrow >= 0; // the ray's last winner first: it sets a tight `best`\n"
						          "\tfor (;;) {\n"
						          "\t\twhile (!found) { // this lane's next row: cheap, per lane\n"
						          "\t\t\tif ((m0 | m1 | m2 | m3) == 0u) {\n"
						          "\t\t\t\tif (++blk >= %d)\n\t\t\t\t\tbreak;\n"
						          "#pragma unroll 1\n"
						          "\t\t\t\tfor (int g = 0; g < %u; ++g) { // collect: group tests, then member tests\n"
						          "\t\t\t\t\tconst int gi = blk * %u + g;\n"
						          "\t\t\t\t\tif (gi >= %u)\n\t\t\t\t\t\tbreak;\n"
						          "\t\t\t\t\tconst lol_u32* gc = lol_run%d_groups + gi * %d;\n"
						          "\t\t\t\t\tconst int rows = (gi + 1) * %u <= %u ? %u : %u - gi * %u;\n"
						          "\t\t\t\t\tif (lol_box_skips(x, y, z, LOL_TF(gc[0]), LOL_TF(gc[1]), LOL_TF(gc[2]), LOL_TF(gc[3]), "
						          "LOL_TF(gc[4]), LOL_TF(gc[5]), LOL_TF(gc[6]), best)) {\n"
						          "\t\t\t\t\t\tlol_count_skip((lol_u32)(rows - (hrow >= gi * %u && hrow < gi * %u + rows)) * %uu);\n"
						          "\t\t\t\t\t\tcontinue;\n\t\t\t\t\t}\n"
						          "\t\t\t\t\tlol_u32 bits = 0u;\n"
						          "#pragma unroll 1\n"
						          "\t\t\t\t\tfor (int k = 0; k < rows; ++k) {\n"
						          "\t\t\t\t\t\tconst lol_u32* ct = lol_run%d + (gi * %u + k) * LOL_RUN%d_STRIDE;\n"
						          "\t\t\t\t\t\tif (gi * %u + k == hrow)\n\t\t\t\t\t\t\tcontinue;\n"
						          "\t\t\t\t\t\tif (lol_box_skips(x, y, z, LOL_TF(ct[0]), LOL_TF(ct[1]), LOL_TF(ct[2]), LOL_TF(ct[3]), "
						          "LOL_TF(ct[4]), LOL_TF(ct[5]), LOL_TF(ct[6]), best))\n"
						          "\t\t\t\t\t\t\tlol_count_skip(%uu);\n"
						          "\t\t\t\t\t\telse\n\t\t\t\t\t\t\tbits |= 1u << k;\n"
						          "\t\t\t\t\t}\n"
						          "\t\t\t\t\tconst int pos = g * %u;\n"
						          "\t\t\t\t\tif (pos < 32) m0 |= bits << pos;\n"
						          "\t\t\t\t\telse if (pos < 64) m1 |= bits << (pos - 32);\n"
						          "\t\t\t\t\telse if (pos < 96) m2 |= bits << (pos - 64);\n"
						          "\t\t\t\t\telse m3 |= bits << (pos - 96);\n"
						          "\t\t\t\t}\n"
						          "\t\t\t\tcontinue;\n"
						          "\t\t\t}\n"
						          "\t\t\tif (m0) { i = __ffs((int)m0) - 1; m0 &= m0 - 1u; }\n"
						          "\t\t\telse if (m1) { i = 31 + __ffs((int)m1); m1 &= m1 - 1u; }\n"
						          "\t\t\telse if (m2) { i = 63 + __ffs((int)m2); m2 &= m2 - 1u; }\n"
						          "\t\t\telse { i = 95 + __ffs((int)m3); m3 &= m3 - 1u; }\n"
						          "\t\t\ti += blk * 128;\n"
						          "\t\t\tconst lol_u32* ct = lol_run%d + i * LOL_RUN%d_STRIDE;\n"
						          "\t\t\t// `best` may be tighter than when the row was collected\n"
						          "\t\t\tif (lol_box_skips(x, y, z, LOL_TF(ct[0]), LOL_TF(ct[1]), LOL_TF(ct[2]), LOL_TF(ct[3]), "
						          "LOL_TF(ct[4]), LOL_TF(ct[5]), LOL_TF(ct[6]), best))\n"
						          "\t\t\t\tlol_count_skip(%uu);\n"
						          "\t\t\telse\n\t\t\t\tfound = true;\n"
						          "\t\t}\n"
						          "\t\tif (!found)\n\t\t\tbreak;\n"
						          "\t\tfound = false;\n"
						          "\t\t{ // one round: every lane that has a row evaluates it\n"
						          "\t\t\tconst lol_u32* c = lol_run%d + i * LOL_RUN%d_STRIDE;\n",
						          (int)nblk, gpb, gpb, ngroups, run_no, LOL_BOUND_SLOTS,
						          LOL_GROUP, n, LOL_GROUP, n, LOL_GROUP,
						          LOL_GROUP, LOL_GROUP, cost,
						          run_no, LOL_GROUP, run_no, LOL_GROUP, cost, LOL_GROUP,
						          run_no, run_no, cost, run_no, run_no);
					}
					/* row = box (C, H, M), object id, then the object's own constants */
					r.out = &sink;
					for (int b = 0; b < LOL_BOUND_SLOTS; b++)
						cst(&r, boxes[order[q]][b]);
					cst_raw(&r, k + 1);
					r.out = real_out;
					free(sink.p);
					if (q == 0)
						sb_printf(&body, "\t\t\tconst lol_u32 oid = c[%d];\n", LOL_BOUND_SLOTS);
					int t = emit_object(&r, s->objects[k]);
					if (q == 0)
						sb_printf(&body,
						          "\t\t\tif (t%d < best || (t%d == best && oid < bid)) {\n"
						          "\t\t\t\tbest = t%d;\n\t\t\t\tbid = oid;\n\t\t\t}\n\t\t}\n\t}\n\t}\n",
						          t, t, t);
				} else {
				if (q == 0) {
						sb_printf(&body, "\t{\n");
						if (use_hint && two)
							sb_printf(&body,
							          "\tconst int hrowA = (hintA >= %uu && hintA < %uu) ? (int)lol_run%d_rowof[hintA - %uu] : -1;\n"
							          "\tconst int hrowB = (hintB >= %uu && hintB < %uu) ? (int)lol_run%d_rowof[hintB - %uu] : -1;\n",
							          i + 1, j + 1, run_no, i + 1, i + 1, j + 1, run_no, i + 1);
						else if (use_hint)
							sb_printf(&body,
							          "\tconst int hrow = (hint >= %uu && hint < %uu) ? (int)lol_run%d_rowof[hint - %uu] : -1;\n",
							          i + 1, j + 1, run_no, i + 1);
						/* group loop; groups -2 / -1 are the hinted rows on their own */
						sb_printf(&body, "#pragma unroll 1\n\tfor (int g = %d; g < %u; ++g) {\n",
						          use_hint ? (two ? -2 : -1) : 0, ngroups);
						sb_printf(&body, "\t\tint first = g * %u, last = first + %u;\n", LOL_GROUP, LOL_GROUP);
						if (n % LOL_GROUP)
							sb_printf(&body, "\t\tif (last > %u)\n\t\t\tlast = %u;\n", n, n);
						if (use_hint && two)
							sb_printf(&body,
							          "\t\tif (g == -2) { // both rays' last winners first: they set a tight `best`\n"
							          "\t\t\tif (hrowA < 0)\n\t\t\t\tcontinue;\n\t\t\tfirst = hrowA;\n\t\t\tlast = first + 1;\n"
							          "\t\t} else if (g == -1) {\n"
							          "\t\t\tif (hrowB < 0 || hrowB == hrowA)\n\t\t\t\tcontinue;\n\t\t\tfirst = hrowB;\n\t\t\tlast = first + 1;\n"
							          "\t\t} else {\n");
						else if (use_hint)
							sb_printf(&body,
							          "\t\tif (g < 0) { // the ray's last winner first: it sets a tight `best`\n"
							          "\t\t\tif (hrow < 0)\n\t\t\t\tcontinue;\n\t\t\tfirst = hrow;\n\t\t\tlast = first + 1;\n"
							          "\t\t} else {\n");
						else
							sb_printf(&body, "\t\t{\n");
						sb_printf(&body,
						          "\t\t\t// the whole group: dist(any member, p) >= dbox(p) - M >= best: none can win%s\n"
						          "\t\t\tconst lol_u32* gc = lol_run%d_groups + g * %d;\n"
						          "\t\t\tif (lol_box_skips%s(x, y, z, LOL_TF(gc[0]), LOL_TF(gc[1]), LOL_TF(gc[2]), LOL_TF(gc[3]), "
						          "LOL_TF(gc[4]), LOL_TF(gc[5]), LOL_TF(gc[6]), %s)) {\n"
						          "\t\t\t\tlol_count_skip((lol_u32)(last - first) * %uu);\n\t\t\t\tcontinue;\n\t\t\t}\n\t\t}\n",
						          two ? "\n\t\t\t// (two rays: skipped only when NEITHER can win; evaluating an object one ray could\n"
						                "\t\t\t// have skipped does not change that ray's result)" : "",
						          run_no, LOL_BOUND_SLOTS, sfx, best_args,
						          (node_cost(s, s->objects[i]) + 1u) * (two ? 2u : 1u));
						sb_printf(&body, "#pragma unroll 1\n\t\tfor (int i = first; i < last; ++i) {\n");
						if (use_hint && two)
							sb_printf(&body, "\t\t\tif (g >= 0 && (i == hrowA || i == hrowB))\n\t\t\t\tcontinue;\n");
						else if (use_hint)
							sb_printf(&body, "\t\t\tif (g >= 0 && i == hrow)\n\t\t\t\tcontinue;\n");
						sb_printf(&body, "\t\t\tconst lol_u32* c = lol_run%d + i * LOL_RUN%d_STRIDE;\n", run_no, run_no);
						sb_printf(&body, "\t\t\tif (%slol_box_skips%s(x, y, z", use_hint ? "g >= 0 && " : "", sfx);
					}
					/* row = box (C, H, M), object id, then the object's own constants */
					for (int b = 0; b < LOL_BOUND_SLOTS; b++) {
						if (q == 0)
							sb_printf(&body, ", ");
						cst(&r, boxes[order[q]][b]);
					}
					if (q == 0)
						sb_printf(&body, ", %s)) {\n\t\t\t\tlol_count_skip(%uu);\n\t\t\t\tcontinue;\n\t\t\t}\n"
						          "\t\t\tconst lol_u32 oid = ", best_args,
						          (node_cost(s, s->objects[i]) + 1u) * (two ? 2u : 1u));
					cst_raw(&r, k + 1);
					if (q == 0)
						sb_printf(&body, ";\n");
					int t = emit_object(&r, s->objects[k]);
					if (q == 0 && two)
						sb_printf(&body,
						          "\t\t\tconst float a_ = lol_lo(t%d), b_ = lol_hi(t%d);\n"
						          "\t\t\tif (a_ < bestA || (a_ == bestA && oid < bidA)) {\n\t\t\t\tbestA = a_;\n\t\t\t\tbidA = oid;\n\t\t\t}\n"
						          "\t\t\tif (b_ < bestB || (b_ == bestB && oid < bidB)) {\n\t\t\t\tbestB = b_;\n\t\t\t\tbidB = oid;\n\t\t\t}\n"
						          "\t\t}\n\t}\n\t}\n",
						          t, t);
					else if (q == 0)
						sb_printf(&body,
						          "\t\t\tif (t%d < best || (t%d == best && oid < bid)) {\n"
						          "\t\t\t\tbest = t%d;\n\t\t\t\tbid = oid;\n\t\t\t}\n\t\t}\n\t}\n\t}\n",
						          t, t, t);
				}
					if (lol_pad_rows)
						while ((r.nrow / 4) % 2 == 0)
							row_push(&r, 0.f);
					if (q == 0)
						per_row = r.nrow;
					sb_printf(&tables.words, "\n\t");
					for (size_t w = 0; w < r.nrow; w++)
						tab_float(&tables, r.row[w]);
					cgen_release(&r);
					free(scratch.p);
				}
				sb_printf(&tables.defs, "#define LOL_RUN%d_STRIDE %zu\n#define LOL_RUN%d_ROWS %u\n#define LOL_RUN%d_GROUP %u\n",
				          run_no, per_row, run_no, n, run_no, LOL_GROUP);
				if (run_no == 0) {
					/* the domain of the candidate grid (lol_near_collect_text): the box around every row's box, a
					 * tenth of its size wider on every side, cut into LOL_GRID_N^3 cells */
					float lo3[3] = {INFINITY, INFINITY, INFINITY}, hi3[3] = {-INFINITY, -INFINITY, -INFINITY};
					int ok = 1;
					for (uint32_t k = 0; k < n; k++)
						for (int a = 0; a < 3; a++) {
							ok &= isfinite(boxes[k][a]) && isfinite(boxes[k][3 + a]) && isfinite(boxes[k][6]);
							lo3[a] = fminf(lo3[a], boxes[k][a] - boxes[k][3 + a] - boxes[k][6]);
							hi3[a] = fmaxf(hi3[a], boxes[k][a] + boxes[k][3 + a] + boxes[k][6]);
						}
					sb_printf(&tables.defs, "#define LOL_GRID_OK %d\n#define LOL_GRID_N %d\n#define LOL_GRID_OUTER %.1ff\n", ok, lol_grid_n, LOL_GRID_OUTER_F);
					for (int a = 0; a < 3 && ok; a++) {
						const float size = (hi3[a] - lo3[a]) * 1.2f + 1e-3f, x0 = lo3[a] - (hi3[a] - lo3[a]) * 0.1f - 5e-4f;
						sb_printf(&tables.defs, "#define LOL_GRID_%c0 ", "XYZ"[a]);
						sb_float(&tables.defs, x0);
						sb_printf(&tables.defs, "\n#define LOL_GRID_S%c ", "XYZ"[a]);
						sb_float(&tables.defs, size / (float)lol_grid_n);
						sb_printf(&tables.defs, "\n#define LOL_GRID_I%c ", "XYZ"[a]);
						sb_float(&tables.defs, (float)lol_grid_n / size);
						/* the outer grid: cells LOL_GRID_OUTER times as large, around the same centre */
						sb_printf(&tables.defs, "\n#define LOL_GRID_%c1 ", "XYZ"[a]);
						sb_float(&tables.defs, x0 - size * (0.5f * (LOL_GRID_OUTER_F - 1.f)));
						sb_printf(&tables.defs, "\n");
					}
				}
				run_no++;
				free(boxes);
				free(order);
				free(rowof);
			}
		} else {
			for (uint32_t k = i; k < j; k++)
				emit_straight_object(&body, &g, k, sigs[k], two, 0, NULL, "\t", NULL, -1);
		}
		i = j;
	}
	if (two)
		sb_printf(&body,
		          "\t// Outside the fast forms' ranges (DESIGN.md, guarded fast path) both rays are\n"
		          "\t// redone the long way, one at a time.\n"
		          "\tif (!(lo >= LOL_SQRT_FAST_MIN)) {\n"
		          "\t\tconst lol_u64 rA = %s(lol_lo(x), lol_lo(y), lol_lo(z));\n"
		          "\t\tconst lol_u64 rB = %s(lol_hi(x), lol_hi(y), lol_hi(z));\n"
		          "\t\tidA = (lol_u32)(rA >> 32);\n\t\tidB = (lol_u32)(rB >> 32);\n"
		          "\t\treturn lol_pk(__uint_as_float((lol_u32)rA), __uint_as_float((lol_u32)rB));\n"
		          "\t}\n\tidA = bidA;\n\tidB = bidB;\n\treturn lol_pk(bestA, bestB);\n}\n",
		          fallback, fallback);
	else if (split_guard)
		sb_printf(&body,
		          "\tok = lo >= LOL_SQRT_FAST_MIN;\n\tid = bid;\n\treturn best;\n}\n"
		          "// The complete guarded function.\n"
		          "__device__ %s float %s(const float x, const float y, const float z,\n"
		          "                                         const lol_u32 hint, lol_u32& id) {\n"
		          "\tbool ok;\n"
		          "\tconst float best = %s_try(x, y, z, hint, id, ok);\n"
		          "\t// The fast forms are bit-identical to IEEE sqrt / division only inside\n"
		          "\t// these ranges (DESIGN.md, guarded fast path); outside, redo it the long way.\n"
		          "\tif (!ok) {\n"
		          "\t\tconst lol_u64 r = %s(x, y, z);\n"
		          "\t\tid = (lol_u32)(r >> 32);\n"
		          "\t\treturn __uint_as_float((lol_u32)r);\n"
		          "\t}\n"
		          "\treturn best;\n}\n",
		          attrs, name, name, fallback);
	else if (fast)
		sb_printf(&body,
		          "\t// The fast forms are bit-identical to IEEE sqrt / division only inside\n"
		          "\t// these ranges (DESIGN.md, guarded fast path); outside, redo it the long way.\n"
		          "\tif (!(lo >= LOL_SQRT_FAST_MIN)) {\n%s"
		          "\t\tconst lol_u64 r = %s(x, y, z);\n"
		          "\t\tid = (lol_u32)(r >> 32);\n"
		          "\t\treturn __uint_as_float((lol_u32)r);\n"
		          "\t}\n",
		          lol_emit_near ? "\t\tlol_near_reset(nr); // whatever was decided with out-of-range values\n" : "", fallback);
	if (two || split_guard)
		;
	else if (packed_ret)
		sb_printf(&body, "\treturn ((lol_u64)bid << 32) | (lol_u64)__float_as_uint(best);\n}\n");
	else
		sb_printf(&body, "\tid = bid;\n\treturn best;\n}\n");

	if (tables_out) {
		const size_t table_bytes = tables.n * 4;
		/* 64 KB of __constant__ space; keep a margin for the kernel parameters. */
		sb_printf(tables_out, "#define LOL_TABLE_SPACE %s\n#define LOL_TAB_WORDS %zu\n",
		          table_bytes <= 56 * 1024 ? "__constant__" : "__device__ const", tables.n);
		if (tables.n) {
			sb_printf(tables_out, "LOL_TABLE_SPACE __align__(16) lol_u32 lol_tables[] = {");
			sb_putn(tables_out, tables.words.p, tables.words.len);
			sb_printf(tables_out, "\n};\n");
			/* up to 96 KB of shared memory per CTA for the copy (two CTAs per SM still fit) */
			if (smem_ok && table_bytes <= 96 * 1024)
				sb_printf(tables_out,
				          "#ifdef LOL_HOST_SHIM\n#define LOL_TAB lol_tables\n#else\n"
				          "// every CTA copies the tables into shared memory once (lol_render's prologue)\n"
				          "extern __shared__ __align__(16) lol_u32 lol_tab_smem[];\n"
				          "#define LOL_TAB lol_tab_smem\n#define LOL_TAB_IN_SMEM 1\n#endif\n");
			else
				sb_printf(tables_out, "#define LOL_TAB lol_tables\n");
			sb_putn(tables_out, tables.defs.p, tables.defs.len);
		}
	}
	if (pc.n) {
		/* the constant pairs of the packed code: ptxas keeps them in uniform register pairs */
		sb_printf(out, "__constant__ __align__(8) lol_u32 lol_pairc[] = {");
		for (size_t i = 0; i < pc.n; i++)
			sb_printf(out, "%s0x%08xu, 0x%08xu,", i % 4 ? " " : "\n\t", pc.w[i][0], pc.w[i][1]);
		sb_printf(out, "\n};\n");
	}
	free(pc.w);
	cgen_release(&g);
	sb_putn(out, body.p, body.len);

	for (uint32_t i = 0; i < s->n_objects; i++)
		free(sigs[i]);
	free(sigs);
	free(body.p);
	free(tables.words.p);
	free(tables.defs.p);
}

/* ------------------------------------------------- guarded fast path proofs */

/* Division by a scene constant without the divider's special-case branch:
 *     q0 = n * rk;  r = fma(-k, q0, n);  q1 = fma(r, rk, q0);     rk = RN(1 / k)
 * is the correctly rounded n / k for "almost all" k (Markstein); instead of
 * relying on the theorem's side conditions the lowering PROVES it for the k at
 * hand by trying all 2^23 significands of n.  Scaling n by a power of two scales
 * q0, r and q1 exactly, so one binade covers every n whose intermediates stay
 * normal; the kernel's range guard and the 0.5 + q rounding cover
 * the rest. */
__attribute__((target("fma"))) static int div_const_proof_fma(float k, float rk) {
	for (uint32_t m = 0; m < (1u << 23); m++) {
		uint32_t bits = 0x3f800000u | m;
		float n, q0, r, q1;
		memcpy(&n, &bits, 4);
		q0 = n * rk;
		r = __builtin_fmaf(-k, q0, n);
		q1 = __builtin_fmaf(r, rk, q0);
		if (q1 != n / k)
			return 0;
	}
	return 1;
}

static int div_const_proof_soft(float k, float rk) {
	for (uint32_t m = 0; m < (1u << 23); m++) {
		uint32_t bits = 0x3f800000u | m;
		float n, q0, r, q1;
		memcpy(&n, &bits, 4);
		q0 = n * rk;
		r = fmaf(-k, q0, n);
		q1 = fmaf(r, rk, q0);
		if (q1 != n / k)
			return 0;
	}
	return 1;
}

/* The proof walks all 2^23 significands (about 14 ms per distinct k): results are kept for
 * the life of the process, shared by all threads -- a group of N devices lowers the same
 * scene N times, and a scene may use any number of distinct smoothness values. */
int lolb200_div_const_is_exact(float k) {
	static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
	static struct { uint32_t kbits; int ok; }* cache;
	static size_t cache_n, cache_cap;
	const uint32_t kbits = f2u(k);
	float rk;
	int ok = -1;
	/* k in [2^-20, 2^20]: quotients of in-range numerators neither overflow nor
	 * lose bits to underflow where 0.5 + q could still notice */
	if (!(k >= 0x1p-20f && k <= 0x1p20f))
		return 0;
	pthread_mutex_lock(&mu);
	for (size_t i = 0; i < cache_n; i++)
		if (cache[i].kbits == kbits) {
			ok = cache[i].ok;
			break;
		}
	pthread_mutex_unlock(&mu);
	if (ok >= 0)
		return ok;
	rk = 1.0f / k;
	ok = __builtin_cpu_supports("fma") ? div_const_proof_fma(k, rk) : div_const_proof_soft(k, rk);
	pthread_mutex_lock(&mu);
	if (cache_n == cache_cap) {
		const size_t cap = cache_cap ? cache_cap * 2 : 16;
		void* grown = realloc(cache, cap * sizeof *cache);
		if (grown) {
			cache = grown;
			cache_cap = cap;
		}
	}
	if (cache_n < cache_cap) {
		cache[cache_n].kbits = kbits;
		cache[cache_n++].ok = ok;
	}
	pthread_mutex_unlock(&mu);
	return ok;
}

/* The guard checks |x|,|y|,|z| <= 2^60 at run time; with every centre, extent and
 * radius below 2^60 too, each squared length stays finite (< 3 * 2^122) and every
 * numerator of a smooth union stays far from overflow. */
static int constants_in_range(const lolb200_scene* s) {
	for (uint32_t i = 0; i < s->n_nodes; i++) {
		const lolb200_object* o = &s->nodes[i];
		const float v[8] = {o->point[0], o->point[1], o->point[2], o->radius,
		                    o->point2[0], o->point2[1], o->point2[2], o->smoothness};
		for (int k = 0; k < 8; k++)
			if (!(fabsf(v[k]) <= 0x1p60f))
				return 0;
	}
	return 1;
}

static int all_divisions_provable(const lolb200_scene* s) {
	for (uint32_t i = 0; i < s->n_nodes; i++)
		if (s->nodes[i].type == LOLB200_OBJ_SMOOTH_UNION &&
		    !lolb200_div_const_is_exact(s->nodes[i].smoothness))
			return 0;
	return 1;
}

/* Measured on B200: the guard costs about as much as it removes when
 * a scene has one or two spheres and no smooth union; from three spheres or one
 * smooth union (a division) on it wins.  guarded_fastpath = 2 forces it. */
static int guard_pays(const lolb200_scene* s) {
	uint32_t spheres = 0, unions = 0;
	for (uint32_t i = 0; i < s->n_nodes; i++) {
		spheres += s->nodes[i].type == LOLB200_OBJ_SPHERE;
		unions += s->nodes[i].type == LOLB200_OBJ_SMOOTH_UNION;
	}
	return unions >= 1 || spheres >= 3;
}

static const char lol_sdf_slow_text[] =
	"// the candidate memory's way out when more rows matter than it holds: ONE out-of-line copy of the plain function\n"
	"__device__ __noinline__ lol_u64 lol_sdf_slow(const float x, const float y, const float z, const lol_u32 hint) {\n"
	"\tlol_u32 id;\n\tconst float d = lol_sdf(x, y, z, hint, id);\n"
	"\treturn ((lol_u64)id << 32) | (lol_u64)__float_as_uint(d);\n}\n";

/* The candidate grid of LOL_NEAR programs (options.near_cache = 3).  Looking at every row again -- 16 group tests
 * and the member tests of the groups that survive -- is a third of what the 1024-primitive scenes execute.  Which
 * rows can matter at a point is mostly a matter of WHERE the point is, and the scene does not move: the box around
 * all rows is cut into LOL_GRID_N^3 cells, and every cell knows the eight rows nearest to it, sorted by a lower
 * bound w of  dbox_row(p) - m1_row  over every point p of the cell (distance between the two boxes), and the bound
 * of the ninth.  A look then reads its cell: the rows with w <= 1.004 |best| are the only ones that can fail to be
 * skipped (their own box test decides, as before); the first w above it is the room.  A list that ends before
 * such a w is found (a crowd, or a large `best`), or a point outside the grid: every row is looked at, as before.
 * The grid is built on the device by lol_grid_build (one thread per cell, once per renderer) from the same table
 * the loops read; the bounds are lowered by 0.01 % and LOL_NEAR_PAD, the cells widened by 0.1 %. */
static const char lol_near_grid_text[] =
	"#if LOL_NEAR_GRID\n"
	"struct __align__(16) lol_cell {\n"
	"\tfloat w[8];       // lower bounds of dbox(p) - m1 over the cell, ascending (+INF: no row)\n"
	"\tlol_u32 rows[2];  // their rows, one per byte\n"
	"\tfloat rest;       // the bound of every row that is not listed\n"
	"\tlol_u32 pad;\n"
	"};\n"
	"#ifdef LOL_HOST_SHIM\n"
	"static lol_cell lol_grid[2 * LOL_GRID_N * LOL_GRID_N * LOL_GRID_N];\n"
	"#else\n"
	"__device__ lol_cell lol_grid[2 * LOL_GRID_N * LOL_GRID_N * LOL_GRID_N];\n"
	"#endif\n"
	"// level 0: the fine grid around the rows; level 1: the same number of cells, LOL_GRID_OUTER times as large each,\n"
	"// around the same centre -- for the points of rays on their way in and out\n"
	"__device__ void lol_grid_build_cell(const int cell) {\n"
	"\tconst int level = cell / (LOL_GRID_N * LOL_GRID_N * LOL_GRID_N), ci = cell % (LOL_GRID_N * LOL_GRID_N * LOL_GRID_N);\n"
	"\tconst int ix = ci % LOL_GRID_N, iy = (ci / LOL_GRID_N) % LOL_GRID_N, iz = ci / (LOL_GRID_N * LOL_GRID_N);\n"
	"\tconst float sx = level ? LOL_GRID_SX * LOL_GRID_OUTER : LOL_GRID_SX, sy = level ? LOL_GRID_SY * LOL_GRID_OUTER : LOL_GRID_SY,\n"
	"\t            sz = level ? LOL_GRID_SZ * LOL_GRID_OUTER : LOL_GRID_SZ;\n"
	"\tconst float x0 = level ? LOL_GRID_X1 : LOL_GRID_X0, y0 = level ? LOL_GRID_Y1 : LOL_GRID_Y0, z0 = level ? LOL_GRID_Z1 : LOL_GRID_Z0;\n"
	"\tconst float cx = x0 + ((float)ix + .5f) * sx, cy = y0 + ((float)iy + .5f) * sy, cz = z0 + ((float)iz + .5f) * sz;\n"
	"\tconst float hx = .5005f * sx + LOL_NEAR_PAD, hy = .5005f * sy + LOL_NEAR_PAD, hz = .5005f * sz + LOL_NEAR_PAD;\n"
	"\tfloat w[9];\n"
	"\tlol_u32 r[9];\n"
	"\tfor (int k = 0; k < 9; ++k) {\n\t\tw[k] = LOL_INF;\n\t\tr[k] = 0xffu;\n\t}\n"
	"\tfor (int i = 0; i < LOL_RUN0_ROWS; ++i) {\n"
	"\t\tconst lol_u32* ct = lol_tables + lol_run0_offset + i * LOL_RUN0_STRIDE;\n"
	"\t\tconst float gx = fmaxf(fabsf(cx - LOL_TF(ct[0])) - (hx + LOL_TF(ct[3])), 0.f);\n"
	"\t\tconst float gy = fmaxf(fabsf(cy - LOL_TF(ct[1])) - (hy + LOL_TF(ct[4])), 0.f);\n"
	"\t\tconst float gz = fmaxf(fabsf(cz - LOL_TF(ct[2])) - (hz + LOL_TF(ct[5])), 0.f);\n"
	"\t\tfloat v = sqrtf(gx * gx + gy * gy + gz * gz) * 0.9999f - LOL_TF(ct[6]) - LOL_NEAR_PAD;\n"
	"\t\tlol_u32 vr = (lol_u32)i;\n"
	"\t\tfor (int k = 0; k < 9; ++k) // insertion into the nine smallest\n"
	"\t\t\tif (v < w[k]) {\n"
	"\t\t\t\tconst float tw = w[k];\n\t\t\t\tconst lol_u32 tr = r[k];\n"
	"\t\t\t\tw[k] = v;\n\t\t\t\tr[k] = vr;\n\t\t\t\tv = tw;\n\t\t\t\tvr = tr;\n"
	"\t\t\t}\n"
	"\t}\n"
	"\tlol_cell c;\n"
	"\tfor (int k = 0; k < 8; ++k)\n\t\tc.w[k] = w[k];\n"
	"\tc.rows[0] = r[0] | (r[1] << 8) | (r[2] << 16) | (r[3] << 24);\n"
	"\tc.rows[1] = r[4] | (r[5] << 8) | (r[6] << 16) | (r[7] << 24);\n"
	"\tc.rest = w[8];\n"
	"\tc.pad = 0u;\n"
	"\tlol_grid[cell] = c;\n"
	"}\n"
	"#ifdef LOL_HOST_SHIM\n"
	"// host builds of the pipeline (the CPU test suite) build a cell when a look first reads it\n"
	"static unsigned char lol_grid_built[2 * LOL_GRID_N * LOL_GRID_N * LOL_GRID_N];\n"
	"static inline const lol_cell& lol_grid_at(const int cell) {\n"
	"\tif (!lol_grid_built[cell]) {\n\t\tlol_grid_build_cell(cell);\n\t\tlol_grid_built[cell] = 1;\n\t}\n"
	"\treturn lol_grid[cell];\n}\n"
	"#else\n"
	"#define lol_grid_at(cell) lol_grid[cell]\n"
	"#endif\n"
	"#ifndef LOL_HOST_SHIM\n"
	"extern \"C\" __global__ void lol_grid_build() {\n"
	"\tconst int ci = (int)(blockIdx.x * blockDim.x + threadIdx.x);\n"
	"\tif (ci < 2 * LOL_GRID_N * LOL_GRID_N * LOL_GRID_N)\n\t\tlol_grid_build_cell(ci);\n"
	"}\n"
	"#endif\n"
	"#endif // LOL_NEAR_GRID\n";

static const char lol_near_collect_text[] =
	"// Every row of the pruned table loop against `best`, tests only (lol_kernel.cuh: struct lol_near, lol_look): the\n"
	"// rows that cannot be skipped (the first eight, one per byte), how many there are, and a lower bound of min over\n"
	"// the skipped rows of  dbox(p) - m1  -- how far the point may move before a skipped row has to be looked at\n"
	"// again (+INF: none skipped).\n"
	"__device__ __noinline__ lol_look lol_near_collect(const float x, const float y, const float z, const float best) {\n"
	"\tfloat room = LOL_INF;\n"
	"\tlol_u64 nc = ~0ull;\n"
	"\tlol_u32 nn = 0u;\n"
	"#if LOL_NEAR_GRID\n"
	"\t{ // the point's cell knows the rows that can matter here, nearest first (lol_grid_build_cell)\n"
	"\t\tfloat fx = (x - LOL_GRID_X0) * LOL_GRID_IX, fy = (y - LOL_GRID_Y0) * LOL_GRID_IY, fz = (z - LOL_GRID_Z0) * LOL_GRID_IZ;\n"
	"\t\tconst float top = (float)LOL_GRID_N;\n"
	"\t\tint base = 0;\n"
	"\t\tif (!(fx >= 0.f && fx < top && fy >= 0.f && fy < top && fz >= 0.f && fz < top)) { // not in the fine grid: the outer one\n"
	"\t\t\tfx = (x - LOL_GRID_X1) * (LOL_GRID_IX / LOL_GRID_OUTER);\n"
	"\t\t\tfy = (y - LOL_GRID_Y1) * (LOL_GRID_IY / LOL_GRID_OUTER);\n"
	"\t\t\tfz = (z - LOL_GRID_Z1) * (LOL_GRID_IZ / LOL_GRID_OUTER);\n"
	"\t\t\tbase = LOL_GRID_N * LOL_GRID_N * LOL_GRID_N;\n"
	"\t\t}\n"
	"\t\tif (fx >= 0.f && fx < top && fy >= 0.f && fy < top && fz >= 0.f && fz < top) { // (a NaN is outside)\n"
	"\t\t\tconst lol_cell cell = lol_grid_at(base + ((int)fz * LOL_GRID_N + (int)fy) * LOL_GRID_N + (int)fx);\n"
	"\t\t\tconst float thr = fabsf(best) * LOL_F(0x3f808366 /*1.00401*/); // a row with w > 1.004 |best| is skipped wherever the point is in the cell\n"
	"\t\t\tfloat next = cell.rest;\n"
	"\t\t\tint c = 0;\n"
	"#pragma unroll\n"
	"\t\t\tfor (int k = 7; k >= 0; --k) {\n"
	"\t\t\t\tif (cell.w[k] > thr)\n\t\t\t\t\tnext = cell.w[k];\n\t\t\t\telse\n\t\t\t\t\t++c;\n"
	"\t\t\t}\n"
	"\t\t\tif (next > thr) { // the list is complete for this `best`: its first c rows, and `next` bounds all others\n"
	"\t\t\t\tLOL_NEAR_STAT(20, 1);\n"
	"\t\t\t\troom = next;\n"
	"\t\t\t\tconst lol_u64 rows = (lol_u64)cell.rows[0] | ((lol_u64)cell.rows[1] << 32);\n"
	"#pragma unroll 1\n"
	"\t\t\t\tfor (int k = 0; k < c; ++k) {\n"
	"\t\t\t\t\tconst lol_u32 i = (lol_u32)(rows >> (8 * k)) & 0xffu;\n"
	"\t\t\t\t\tconst lol_u32* ct = lol_run0 + i * LOL_RUN0_STRIDE;\n"
	"\t\t\t\t\tconst float q2 = lol_box_q2(x, y, z, LOL_TF(ct[0]), LOL_TF(ct[1]), LOL_TF(ct[2]), LOL_TF(ct[3]), LOL_TF(ct[4]), LOL_TF(ct[5]));\n"
	"\t\t\t\t\tif (lol_q2_skips(q2, LOL_TF(ct[6]), best)) {\n"
	"\t\t\t\t\t\troom = fminf(room, lol_box_gap(q2, LOL_TF(ct[6])));\n"
	"\t\t\t\t\t\tcontinue;\n"
	"\t\t\t\t\t}\n"
	"\t\t\t\t\tnc = (nc & ~(0xffull << (8u * nn))) | ((lol_u64)i << (8u * nn)); // (c <= 8: nn < 8 here)\n"
	"\t\t\t\t\t++nn;\n"
	"\t\t\t\t}\n"
	"\t\t\t\tLOL_NEAR_STAT(4 + (nn < 15u ? nn : 15u), 1);\n"
	"\t\t\t\treturn lol_look_make(nc, room, nn);\n"
	"\t\t\t}\n"
	"\t\t\tLOL_NEAR_STAT(22, 1); // the cell's list ends before a bound above the threshold\n"
	"\t\t} else\n"
	"\t\t\tLOL_NEAR_STAT(21, 1); // outside the grid\n"
	"\t}\n"
	"#endif\n"
	"#pragma unroll 1\n"
	"\tfor (int g = 0; g * LOL_RUN0_GROUP < LOL_RUN0_ROWS; ++g) {\n"
	"\t\tconst lol_u32* gc = lol_run0_groups + g * 7;\n"
	"\t\tconst float gq2 = lol_box_q2(x, y, z, LOL_TF(gc[0]), LOL_TF(gc[1]), LOL_TF(gc[2]), LOL_TF(gc[3]), LOL_TF(gc[4]), LOL_TF(gc[5]));\n"
	"\t\tif (lol_q2_skips(gq2, LOL_TF(gc[6]), best)) {\n"
	"\t\t\troom = fminf(room, lol_box_gap(gq2, LOL_TF(gc[6])));\n"
	"\t\t\tcontinue;\n"
	"\t\t}\n"
	"\t\tconst int last = (g + 1) * LOL_RUN0_GROUP <= LOL_RUN0_ROWS ? (g + 1) * LOL_RUN0_GROUP : LOL_RUN0_ROWS;\n"
	"#pragma unroll 1\n"
	"\t\tfor (int i = g * LOL_RUN0_GROUP; i < last; ++i) {\n"
	"\t\t\tconst lol_u32* ct = lol_run0 + i * LOL_RUN0_STRIDE;\n"
	"\t\t\tconst float q2 = lol_box_q2(x, y, z, LOL_TF(ct[0]), LOL_TF(ct[1]), LOL_TF(ct[2]), LOL_TF(ct[3]), LOL_TF(ct[4]), LOL_TF(ct[5]));\n"
	"\t\t\tif (lol_q2_skips(q2, LOL_TF(ct[6]), best)) {\n"
	"\t\t\t\troom = fminf(room, lol_box_gap(q2, LOL_TF(ct[6])));\n"
	"\t\t\t\tcontinue;\n"
	"\t\t\t}\n"
	"\t\t\tif (nn < 8u)\n"
	"\t\t\t\tnc = (nc & ~(0xffull << (8u * nn))) | ((lol_u64)(lol_u32)i << (8u * nn));\n"
	"\t\t\t++nn;\n"
	"\t\t}\n"
	"\t}\n"
	"\tLOL_NEAR_STAT(4 + (nn < 15u ? nn : 15u), 1);\n"
	"\treturn lol_look_make(nc, room, nn);\n"
	"}\n";

/* One pruned table loop of at most 254 rows, and nothing else looped: the shape the per-ray candidate
 * memory handles (row numbers are bytes, 0xff = none). */
static int single_pruned_run(const lolb200_scene* s, int threshold) {
	int runs = 0, ok = 1;
	char** sigs = calloc(s->n_objects ? s->n_objects : 1, sizeof *sigs);
	for (uint32_t i = 0; i < s->n_objects; i++) {
		struct sb sig = {0};
		signature(s, s->objects[i], &sig);
		sigs[i] = sig.p;
	}
	for (uint32_t i = 0; i < s->n_objects;) {
		uint32_t j = i + 1;
		while (j < s->n_objects && strcmp(sigs[j], sigs[i]) == 0)
			j++;
		if ((int)(j - i) >= threshold) {
			runs++;
			ok &= j - i <= 254u;
		}
		i = j;
	}
	for (uint32_t i = 0; i < s->n_objects; i++)
		free(sigs[i]);
	free(sigs);
	return runs == 1 && ok;
}

static void emit_sdf(struct sb* out, const lolb200_scene* s, int loop_threshold, int guarded,
                     int prune, int two, int smem_ok, int pack, int near, int guard_out) {
	struct sb tables = {0};
	struct est_memo memo = {{0, 0}, {0, 0}, {NULL, NULL}};
	near = (prune && !two && single_pruned_run(s, loop_threshold)) ? near : 0;
	lol_pad_rows = near || lol_worklist;
	sb_printf(out, "#define LOL_NEAR %d\n", near > 2 ? 2 : near);
	if (guarded == 1 && !guard_pays(s) && !two)
		guarded = 0;
	if (guarded && constants_in_range(s)) {
		struct sb ref = {0};
		int div_ok = all_divisions_provable(s);
		sb_printf(out, "#define LOL_GUARDED 1\n#define LOL_DIV_CONST %d\n", div_ok);
		/* variant 1's march loops call lol_sdf_try and keep the guard's fall-back outside the loop */
		sb_printf(out, "#define LOL_GUARD_OUT %d\n", (!near && !two) ? guard_out : 0);
		emit_sdf_fn(&ref, &tables, s, loop_threshold, "lol_sdf_ref", "__noinline__", 0, 0, NULL, prune, 0, smem_ok, &memo, 0, pack, div_ok);
		sb_putn(out, tables.p, tables.len);
		sb_putn(out, ref.p, ref.len);
		emit_sdf_fn(out, NULL, s, loop_threshold, "lol_sdf", "__forceinline__", 1, div_ok,
		            "lol_sdf_ref", prune, 0, smem_ok, &memo, pack, pack, div_ok);
		if (two)
			emit_sdf_fn(out, NULL, s, loop_threshold, "lol_sdf2", "__forceinline__", 1, div_ok,
			            "lol_sdf_ref", prune, 1, smem_ok, &memo, 0, pack, div_ok);
		if (near) {
			sb_printf(out, "%s", lol_sdf_slow_text);
			sb_printf(out, "#define LOL_NEAR_GRID (%d && LOL_GRID_OK)\n#define LOL_NEAR_WIDE %du\n", near >= 3, near >= 3 ? 8 : 4);
			sb_putn(out, lol_near_grid_text, strlen(lol_near_grid_text));
			sb_putn(out, lol_near_collect_text, strlen(lol_near_collect_text));
			lol_emit_near = 1;
			sb_printf(out, "#define lol_pairc lol_pairc_nr // this form's own pair constants\n");
			emit_sdf_fn(out, NULL, s, loop_threshold, "lol_sdf_nr", "__forceinline__", 1, div_ok,
			            "lol_sdf_ref", prune, 0, smem_ok, &memo, pack, pack, div_ok);
			sb_printf(out, "#undef lol_pairc\n");
			lol_emit_near = 0;
		}
		free(ref.p);
	} else {
		struct sb fn = {0};
		sb_printf(out, "#define LOL_GUARDED 0\n#define LOL_DIV_CONST 0\n#define LOL_GUARD_OUT 0\n");
		emit_sdf_fn(&fn, &tables, s, loop_threshold, "lol_sdf", "__forceinline__", 0, 0, NULL, prune, 0, smem_ok, &memo, 0, 0, 0);
		sb_putn(out, tables.p, tables.len);
		sb_putn(out, fn.p, fn.len);
		free(fn.p);
		if (near) {
			sb_printf(out, "%s", lol_sdf_slow_text);
			sb_printf(out, "#define LOL_NEAR_GRID (%d && LOL_GRID_OK)\n#define LOL_NEAR_WIDE %du\n", near >= 3, near >= 3 ? 8 : 4);
			sb_putn(out, lol_near_grid_text, strlen(lol_near_grid_text));
			sb_putn(out, lol_near_collect_text, strlen(lol_near_collect_text));
			lol_emit_near = 1;
			emit_sdf_fn(out, NULL, s, loop_threshold, "lol_sdf_nr", "__forceinline__", 0, 0, NULL, prune, 0, smem_ok, &memo, 0, 0, 0);
			lol_emit_near = 0;
		}
	}
	lol_pad_rows = 0;
	free(tables.p);
	free(memo.own[0]);
	free(memo.own[1]);
}

/* ----------------------------------------------------- exact-skip conditions */

static int finite3(const float v[3]) { return isfinite(v[0]) && isfinite(v[1]) && isfinite(v[2]); }
static int zero3(const float v[3]) { return v[0] == 0.f && v[1] == 0.f && v[2] == 0.f; }

/* All light intensities and the ambient colour finite: `finite * 0 == 0` holds. */
static int lights_finite(const lolb200_scene* s) {
	for (uint32_t i = 0; i < s->n_lights; i++)
		if (!finite3(s->lights[i].diffuse_intensity) || !finite3(s->lights[i].specular_intensity))
			return 0;
	return finite3(s->ambient_color);
}

/* powf(x in [0,1], shininess) stays finite iff shininess >= 0 (powf(0,0) = 1). */
static int shininess_ok(const lolb200_material* m) {
	return isfinite(m->shininess) && m->shininess >= 0.f;
}

/* Miss pixels take material 0 (naive_renderer.c:105-109).  With kd = ks = ka = 0
 * and everything else finite each term of get_light is +-0 and the pixel is
 * exactly black. */
int lolb200_can_skip_black_miss(const lolb200_scene* s) {
	const lolb200_material* m = &s->materials[0];
	return lights_finite(s) && shininess_ok(m) && zero3(m->diffuse) && zero3(m->specular) &&
	       zero3(m->ambient);
}

/* A light with clamp(n.l) == 0 adds (I*(shadow*0))*k = +-0 twice: exact for
 * finite I, k and finite powf, i.e. for every material a hit can select. */
int lolb200_can_cull_backfacing(const lolb200_scene* s) {
	if (!lights_finite(s))
		return 0;
	for (uint32_t i = 0; i < s->n_materials; i++) {
		const lolb200_material* m = &s->materials[i];
		if (!shininess_ok(m) || !finite3(m->diffuse) || !finite3(m->specular))
			return 0;
	}
	return 1;
}

/* The shadow march may stop once res <= 0 (softshadow then returns maxf(res, 0) = 0) if
 * nothing can RAISE res afterwards.  res = minf(res, q) only falls while q is ordered; a NaN q
 * replaces res (MINSS hands back its second operand) and the next ordered q could then lift it
 * above 0.  q = (50 * d) / t is NaN only for 0/0, inf/inf or a NaN operand.  With every scene
 * constant and light position finite, sdf() of a finite point is finite (sums of squares, sqrt,
 * a division by the constant k whose result goes through clamp), so what remains is 0/0, t == 0:
 *   - t starts at 0 and the first step either leaves the loop (d < 0: q = -inf < -1), makes res
 *     NaN (d == 0: not <= 0, no early-out), or moves to t = d > 0 with q = +inf;
 *   - from t > 0 on, a step that does not leave the loop has q >= -1, i.e. d >= -t/50, so
 *     t + d >= 0.98 t > 0 (for a denormal t the only d in (-t/50, 0) is -0): t stays positive.
 * Hence at the moment res <= 0 is seen t is positive and finite and stays so: no later q is NaN.
 * A scene with a non-finite constant, or a NaN camera (then d is NaN, res NaN, never <= 0 -- safe
 * anyway), is the only way out of this argument, so the licence is per scene like the others. */
int lolb200_can_shadow_early(const lolb200_scene* s) {
	for (uint32_t i = 0; i < s->n_lights; i++)
		if (!finite3(s->lights[i].point))
			return 0;
	for (uint32_t i = 0; i < s->n_nodes; i++) {
		const lolb200_object* o = &s->nodes[i];
		if (!finite3(o->point) || !finite3(o->point2) || !isfinite(o->radius) || !isfinite(o->smoothness))
			return 0;
	}
	return 1;
}

/* Does the scene lower to a table loop (a run of >= threshold same-shaped
 * top-level objects)?  A BRUTE-FORCE loop over a big scene is where the two-rays-
 * per-thread kernel wins (measured on B200, 1024 spheres, pruning off: 1.38x).
 * With the pruned loops (boxes, groups, hints) a ray computes only a handful of
 * objects and the pair loses again: both rays of a thread must agree to skip, and
 * there are half as many warps to hide the table loads (4K: 50.6 vs 47.8 ms,
 * 1080p: 25.4 vs 13.8 ms).  The small example scenes are 10-20 % slower with it
 * (DESIGN.md).  So: variant 3 only for unpruned table loops. */
static int has_table_loop(const lolb200_scene* s, int threshold) {
	int found = 0;
	char** sigs = calloc(s->n_objects ? s->n_objects : 1, sizeof *sigs);
	for (uint32_t i = 0; i < s->n_objects; i++) {
		struct sb sig = {0};
		signature(s, s->objects[i], &sig);
		sigs[i] = sig.p;
	}
	for (uint32_t i = 0; i < s->n_objects && !found;) {
		uint32_t j = i + 1;
		while (j < s->n_objects && strcmp(sigs[j], sigs[i]) == 0)
			j++;
		if ((int)(j - i) >= threshold)
			found = 1;
		i = j;
	}
	for (uint32_t i = 0; i < s->n_objects; i++)
		free(sigs[i]);
	free(sigs);
	return found;
}

/* ---------------------------------------------- per-child materials (extension) */

/* material of a node with #0 meaning "my parent's" */
static uint32_t effective_material(const lolb200_object* o, uint32_t inherited) {
	return o->material ? o->material : inherited;
}

/* does any leaf below idx end up with a material other than `m`? */
static int subtree_has_other_material(const lolb200_scene* s, uint32_t idx, uint32_t inherited, uint32_t m) {
	const lolb200_object* o = &s->nodes[idx];
	const uint32_t eff = effective_material(o, inherited);
	if (!LOLB200_OBJ_HAS_CHILDREN(o->type))
		return eff != m;
	return subtree_has_other_material(s, (uint32_t)o->a, eff, m) ||
	       subtree_has_other_material(s, (uint32_t)o->b, eff, m);
}

/* distance t<N> (IEEE forms, the reference's operation order) and material m<N> of a subtree */
static int emit_mat_node(struct cgen* g, uint32_t idx, uint32_t inherited) {
	const lolb200_object* o = &g->s->nodes[idx];
	const uint32_t eff = effective_material(o, inherited);
	int me;
	if (!LOLB200_OBJ_HAS_CHILDREN(o->type)) {
		me = emit_node(g, idx);
		sb_printf(g->out, "%sconst lol_u32 m%d = %uu;\n", g->indent, me, eff);
		return me;
	}
	const int a = emit_mat_node(g, (uint32_t)o->a, eff);
	const int b = emit_mat_node(g, (uint32_t)o->b, eff);
	me = g->tmp++;
	switch (o->type) {
	case LOLB200_OBJ_UNION:
		sb_printf(g->out, "%sconst float t%d = lol_csg_union(t%d, t%d);\n", g->indent, me, a, b);
		sb_printf(g->out, "%sconst lol_u32 m%d = (t%d < t%d) ? m%d : m%d;\n", g->indent, me, b, a, b, a);
		break;
	case LOLB200_OBJ_INTERSECTION:
		sb_printf(g->out, "%sconst float t%d = lol_csg_inter(t%d, t%d);\n", g->indent, me, a, b);
		sb_printf(g->out, "%sconst lol_u32 m%d = (t%d > t%d) ? m%d : m%d;\n", g->indent, me, b, a, b, a);
		break;
	case LOLB200_OBJ_DIFFERENCE:
		sb_printf(g->out, "%sconst float t%d = lol_csg_diff(t%d, t%d);\n", g->indent, me, a, b);
		sb_printf(g->out, "%sconst lol_u32 m%d = (-t%d > t%d) ? m%d : m%d;\n", g->indent, me, b, a, b, a);
		break;
	default:
		sb_printf(g->out, "%sconst float t%d = lol_smin(t%d, t%d, ", g->indent, me, a, b);
		sb_float(g->out, o->smoothness);
		sb_printf(g->out, ");\n");
		sb_printf(g->out, "%sconst lol_u32 m%d = (t%d < t%d) ? m%d : m%d;\n", g->indent, me, b, a, b, a);
		break;
	}
	return me;
}

/* lol_scene_materials[] (by material index) and lol_child_material(): one case per top-level
 * object whose leaves do not all share its material; every other id keeps its object's. */
static void emit_child_materials(struct sb* out, const lolb200_scene* s) {
	sb_printf(out, "__device__ const lol_u32 lol_scene_materials[] = {\n");
	for (uint32_t mi = 0; mi < s->n_materials; mi++) {
		const lolb200_material* m = &s->materials[mi];
		const float row[10] = {m->shininess,   m->diffuse[0],  m->diffuse[1], m->diffuse[2],
		                       m->specular[0], m->specular[1], m->specular[2], m->ambient[0],
		                       m->ambient[1],  m->ambient[2]};
		sb_printf(out, "\t/* material #%u */ ", mi);
		for (int k = 0; k < 10; k++) {
			sb_bits(out, row[k]);
			sb_printf(out, ", ");
		}
		sb_printf(out, "0u, 0u,\n");
	}
	sb_printf(out, "};\n__device__ const lol_u32 lol_object_material[] = {0u");
	for (uint32_t i = 0; i < s->n_objects; i++)
		sb_printf(out, ", %uu", s->nodes[s->objects[i]].material);
	sb_printf(out, "};\n"
	               "// The material at a hit point of object `id` (extension: options.child_materials).\n"
	               "__device__ __noinline__ lol_u32 lol_child_material(float x, float y, float z, lol_u32 id) {\n"
	               "\tswitch (id) {\n");
	for (uint32_t i = 0; i < s->n_objects; i++) {
		const uint32_t root = s->objects[i];
		const lolb200_object* o = &s->nodes[root];
		struct cgen g = {.s = s, .out = out, .indent = "\t\t"};
		int t;
		if (!LOLB200_OBJ_HAS_CHILDREN(o->type) || !subtree_has_other_material(s, root, o->material, o->material))
			continue;
		sb_printf(out, "\tcase %uu: {\n", i + 1);
		t = emit_mat_node(&g, root, o->material);
		sb_printf(out, "\t\treturn m%d;\n\t}\n", t);
		cgen_release(&g);
	}
	sb_printf(out, "\tdefault: return lol_object_material[id];\n\t}\n}\n");
}

/* ------------------------------------------------------------------- driver */

static void emit_tables(struct sb* out, const lolb200_scene* s) {
	sb_printf(out, "#define LOL_NLIGHTS %u\n#define LOL_NOBJECTS %u\n", s->n_lights, s->n_objects);
	sb_printf(out, "#define LOL_AMBIENT_R ");
	sb_float(out, s->ambient_color[0]);
	sb_printf(out, "\n#define LOL_AMBIENT_G ");
	sb_float(out, s->ambient_color[1]);
	sb_printf(out, "\n#define LOL_AMBIENT_B ");
	sb_float(out, s->ambient_color[2]);
	sb_printf(out, "\n");

	/* get_material (naive_renderer.c:102-112) resolved per object id at lowering
	 * time: row 0 = material 0 (misses), row i = material of top-level object i.
	 * Row: shininess, diffuse[3], specular[3], ambient[3], 2 pad. */
	sb_printf(out, "__device__ const lol_u32 lol_materials[] = {\n");
	for (uint32_t id = 0; id <= s->n_objects; id++) {
		uint32_t mi = id ? s->nodes[s->objects[id - 1]].material : 0;
		const lolb200_material* m = &s->materials[mi];
		const float row[10] = {m->shininess,   m->diffuse[0],  m->diffuse[1], m->diffuse[2],
		                       m->specular[0], m->specular[1], m->specular[2], m->ambient[0],
		                       m->ambient[1],  m->ambient[2]};
		sb_printf(out, "\t/* id %u -> material #%u */ ", id, mi);
		for (int k = 0; k < 10; k++) {
			sb_bits(out, row[k]);
			sb_printf(out, ", ");
		}
		sb_printf(out, "0u, 0u,\n");
	}
	sb_printf(out, "};\n");

	sb_printf(out,
	          "// struct light (scene.h:52-56) by index; `i` is a compile-time constant at\n"
	          "// every call site, so the switch folds to immediates.\n"
	          "__device__ __forceinline__ void lol_light(const int i, float& px, float& py, "
	          "float& pz,\n\t\tfloat& dr, float& dg, float& db, float& sr, float& sg, float& sb) "
	          "{\n\tswitch (i) {\n");
	for (uint32_t i = 0; i < s->n_lights; i++) {
		const lolb200_light* l = &s->lights[i];
		const float v[9] = {l->point[0], l->point[1], l->point[2],
		                    l->diffuse_intensity[0], l->diffuse_intensity[1], l->diffuse_intensity[2],
		                    l->specular_intensity[0], l->specular_intensity[1], l->specular_intensity[2]};
		static const char* names[9] = {"px", "py", "pz", "dr", "dg", "db", "sr", "sg", "sb"};
		sb_printf(out, "\t%s %u:\n", i + 1 == s->n_lights ? "default: // light" : "case", i);
		for (int k = 0; k < 9; k++) {
			sb_printf(out, "\t\t%s = ", names[k]);
			sb_float(out, v[k]);
			sb_printf(out, ";\n");
		}
		sb_printf(out, "\t\tbreak;\n");
	}
	if (s->n_lights == 0)
		sb_printf(out, "\tdefault: px = py = pz = dr = dg = db = sr = sg = sb = 0.f;\n");
	sb_printf(out, "\t}\n}\n");
}

char* lolb200_lower_cuda(const lolb200_scene* s, const lolb200_options* opt, size_t* len) {
	lolb200_options o;
	struct sb out = {0};
	const char* marker;
	int variant, threshold;

	if (lolb200_scene_check(s) != LOLB200_OK)
		return NULL;
	if (opt)
		o = *opt;
	else
		lolb200_options_default(&o);
	threshold = o.loop_threshold > 0 ? o.loop_threshold : 16;
	lol_group = o.prune_group > 0 ? (uint32_t)o.prune_group : 8u;
	lol_worklist = o.loop_worklist < 0 ? 0 : o.loop_worklist; /* measured slower on B200: off */
	lol_no_balls = o.prune_bounds == 3;
	lol_force_balls = o.prune_bounds == 4;
	if (o.prune_bounds == 4)
		o.prune_bounds = 2;
	/* measured on B200, 1024 spheres at 4K: 16 cells per axis 25.31 ms, 32: 24.33, 48: 24.00, 64: 23.75 (25 MB of cells) */
	lol_grid_n = (o.grid_cells == 16 || o.grid_cells == 32 || o.grid_cells == 48 || o.grid_cells == 64) ? o.grid_cells : 64;
	if (o.prune_bounds == 3)
		o.prune_bounds = 1;
	variant = o.variant;
	if (variant == 0) /* chosen per scene */
		variant = (has_table_loop(s, threshold) && !o.prune_bounds) ? 3 : LOLB200_DEFAULT_VARIANT;
	if (variant == 2 && s->n_objects > 65535u)
		variant = 1; /* variant 2 keeps object ids in 16 bits */
	if (variant == 4 && s->n_lights > 30u)
		variant = 1; /* variant 4: one continuation queue per light, counters packed in bytes */
	if (variant < 1 || variant > 4) {
		lolb200_set_error("unknown kernel variant %d", variant);
		return NULL;
	}
	/* Variant 3 (two rays per thread, packed FP32) is built from the guarded fast
	 * forms: it needs exact arithmetic, the guard enabled and every scene constant
	 * inside the guard's range.  Otherwise variant 1 is what runs. */
	if (variant == 3 && (o.arith != LOLB200_ARITH_EXACT || !o.guarded_fastpath || !constants_in_range(s)))
		variant = 1;

	sb_printf(&out, "// Generated by lolb200_lower_cuda (ABI %d) -- do not edit.\n",
	          LOLB200_ABI_VERSION);
	sb_printf(&out, "// scene: %u objects (%u nodes), %u lights, %u materials, %llu FLOP per sdf()\n",
	          s->n_objects, s->n_nodes, s->n_lights, s->n_materials,
	          (unsigned long long)lolb200_scene_flops_per_eval(s));
	sb_printf(&out, "#define LOL_EXACT %d\n", o.arith == LOLB200_ARITH_EXACT);
	sb_printf(&out, "#define LOL_SKIP_MISS %d\n", o.skip_black_miss && lolb200_can_skip_black_miss(s));
	sb_printf(&out, "#define LOL_CULL %d\n", o.cull_backfacing && lolb200_can_cull_backfacing(s));
	sb_printf(&out, "#define LOL_SHADOW_EARLY %d\n", o.shadow_early_out && lolb200_can_shadow_early(s));
	sb_printf(&out, "#define LOL_COUNTERS %d\n", o.counters != 0);
	sb_printf(&out, "#define LOL_VARIANT %d\n", variant);
	sb_printf(&out, "#define LOL_DIV_PRETEST %d\n", variant == 1 && o.arith == LOLB200_ARITH_EXACT && o.shadow_early_out != 0 &&
	                                                  lolb200_can_shadow_early(s) && o.shadow_div_pretest != 0);
	/* share_first_step = 1: where it was measured to pay (B200, 4K: scene 0.694 -> 0.679 ms, scene4
	 * 2.149 -> 2.129 ms; scene2 loses 0.8 %, the table-loop scenes more); 2 forces it */
	sb_printf(&out, "#define LOL_SHARE_FIRST %d\n#define LOL_SDF_FLOPS %lluu\n",
	          variant == 1 && (o.share_first_step >= 2 || (o.share_first_step == 1 && !has_table_loop(s, threshold))),
	          (unsigned long long)lolb200_scene_flops_per_eval(s));
	{
		/* CTA shape.  Variant 3 holds two rays per thread (about twice the
		 * registers): 128-thread CTAs let the register file be divided more finely. */
		int threads = o.block_threads > 0 ? o.block_threads : (variant == 3 ? 128 : LOLB200_KERNEL_THREADS);
		if (threads % 32 || threads > 1024) {
			lolb200_set_error("block_threads = %d: must be a multiple of 32, at most 1024", threads);
			free(out.p);
			return NULL;
		}
		/* min_blocks caps the registers; measured best for variant 3: 5 CTAs of 128 */
		/* variant 1: four CTAs of 256 (64 registers) -- without the cap ptxas takes 68 on
		 * scene4 and a quarter of the warps is gone (measured 2.39 vs 2.30 ms) */
		const int min_blocks = o.min_blocks > 0 ? o.min_blocks
		                       : (variant == 3 && threads == 128) ? 5
		                       : ((variant == 1 || variant == 4) && threads == 256) ? 4 : 0;
		sb_printf(&out, "#define LOL_THREADS %d\n", threads);
		if (min_blocks)
			sb_printf(&out, "#define LOL_LAUNCH_BOUNDS __launch_bounds__(%d, %d)\n", threads, min_blocks);
		else
			sb_printf(&out, "#define LOL_LAUNCH_BOUNDS __launch_bounds__(%d)\n", threads);
		sb_printf(&out, "#define LOL_ROLL_PHASES %d\n", o.roll_phases != 0);
		/* default: straight-line scenes unrolled (rolled: scene4 +0.8 % / +2.5 %), table-loop scenes rolled
		 * (their distance code is long: 1024 spheres 43.8 -> 43.2 ms, and the per-ray candidate memory
		 * adds to every copy) -- measured on B200, DESIGN.md */
		sb_printf(&out, "#define LOL_ROLL_V1 %d\n",
		          o.roll_v1 < 0 ? (has_table_loop(s, threshold) ? 2 : LOLB200_DEFAULT_ROLL_V1) : o.roll_v1);
	}
	if (variant == 2) {
		/* struct lol_warp_smem (lol_kernel.cuh): p, n, t|px, dir, sh[lights], id,
		 * hits, task (+ nsh in instrumented builds), 128 pixels per warp */
		const unsigned lights = s->n_lights ? s->n_lights : 1;
		sb_printf(&out, "#define LOL_SMEM_PER_WARP %u\n",
		          4u * 128u * (3u + 3u + 1u + 4u + lights) + 128u * 4u + (o.counters ? 256u : 0u));
	}

	marker = strstr(lol_kernel_text, "//@@SCENE@@");
	if (!marker) {
		lolb200_set_error("kernel text has no scene marker");
		free(out.p);
		return NULL;
	}
	sb_putn(&out, lol_params_text, strlen(lol_params_text));
	sb_putn(&out, lol_kernel_text, (size_t)(marker - lol_kernel_text));
	emit_tables(&out, s);
	sb_printf(&out, "#define LOL_CHILD_MATERIALS %d\n", o.child_materials != 0);
	if (o.child_materials)
		emit_child_materials(&out, s);
	emit_sdf(&out, s, threshold, o.arith == LOLB200_ARITH_EXACT ? o.guarded_fastpath : 0,
	         o.prune_bounds, variant == 3, variant != 2 /* variant 2's dynamic smem holds its queues */,
	         o.pack_pairs, variant == 1 ? (o.near_cache < 0 ? 3 : o.near_cache) : 0,
	         (variant == 1 && !(o.shadow_div_pretest != 0)) ? (o.guard_out < 0 ? 3 : (o.guard_out & 3)) : 0);
	sb_putn(&out, marker, strlen(marker));

	if (len)
		*len = out.len;
	return out.p;
}
