/* oracle/er_port.c -- CPU RESTATEMENT ORACLE ("port").  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's Extremal-Region detect+classify hot path, written
 * from the behaviour of HsiehYiChia/Scene-text-recognition @1025d09 (each function cites the
 * reference file:line it follows).  It is array based (no per-node heap objects, no OpenCV) but
 * reproduces the reference's arithmetic AND traversal order, so node sets, child order, pool
 * order, histograms and scores are comparable bit for bit.
 *
 * Parity pin: tests/test_oracle.py and tests/test_port_next.py check every function here against
 * oracle/_ref/libref_oracle.so (the reference's own code, built by oracle/build_ref.sh) on the
 * ICDAR fixtures and on synthetic planes, and tests/golden/ holds the outputs of that
 * reference build for the GPU box (where /root/reference does not exist).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this file's
 * shared object.  The product path (libertext.so) never links or calls it.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <float.h>

#define PORT_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * a1  ERFilter::compute_channels  (src/ER.cpp:114-128) -- cvtColor(BGR2YCrCb) + split + 255-x.
 * OpenCV's 8-bit path is fixed point with 14 fractional bits; coefficients are
 * round(0.299*2^14)=4899, 9617, 1868 and round(0.713*2^14)=11682, round(0.564*2^14)=9241.
 * ------------------------------------------------------------------------------------------ */
PORT_API void port_channels(const uint8_t *bgr, int w, int h, int stride, uint8_t *planes6)
{
	const size_t n = (size_t)w * h;
	for (int y = 0; y < h; y++) {
		const uint8_t *row = bgr + (size_t)y * stride;
		for (int x = 0; x < w; x++) {
			const int b = row[3 * x], g = row[3 * x + 1], r = row[3 * x + 2];
			const int lum = (4899 * r + 9617 * g + 1868 * b + 8192) >> 14;
			int cr = ((r - lum) * 11682 + (128 << 14) + 8192) >> 14;
			int cb = ((b - lum) * 9241 + (128 << 14) + 8192) >> 14;
			if (cr < 0) cr = 0; else if (cr > 255) cr = 255;
			if (cb < 0) cb = 0; else if (cb > 255) cb = 255;
			const size_t i = (size_t)y * w + x;
			planes6[i] = (uint8_t)lum;
			planes6[n + i] = (uint8_t)cr;
			planes6[2 * n + i] = (uint8_t)cb;
			planes6[3 * n + i] = (uint8_t)(255 - lum);
			planes6[4 * n + i] = (uint8_t)(255 - cr);
			planes6[5 * n + i] = (uint8_t)(255 - cb);
		}
	}
}

/* `input_clone /= THRESH_STEP` (src/ER.cpp:250): cv::Mat /= s is convertTo with a FLOAT scale
 * 1/s, rounded half-to-even (cvRound) and saturated. */
static inline uint8_t quant_level(uint8_t v, int step)
{
	const float a = (float)(1.0 / (double)step);
	long q = lrintf((float)v * a);
	return (uint8_t)(q > 255 ? 255 : q);
}

PORT_API void port_quantize(const uint8_t *src, int n, int step, uint8_t *dst)
{
	for (int i = 0; i < n; i++) dst[i] = quant_level(src[i], step);
}

/* ------------------------------------------------------------------------------------------
 * a2/a3  component tree.  Node record mirrors the fields of `struct ER` that the hot path
 * uses (inc/ER.h:42-80): level, area (starts at 1: src/ER.cpp:6), bound, parent/child/next.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
	int level, area;
	int x0, y0, x1, y1;        /* inclusive pixel bounding box */
	int seed_pix;              /* pixel the component was opened on (ER::pixel) */
	int parent, child, next;   /* indices, -1 = null */
	int alive;
	/* NMS state */
	int done;
	double stability;
} PNode;

typedef struct {
	PNode *nd;
	int n, cap;
	int root;
	int w, h;
	/* flattened view (DFS pre-order, reference child order) */
	int *flat;      /* flat index -> node id */
	int *flat_of;   /* node id -> flat index */
	int nflat;
	int *pool; int npool;   /* node ids, reference push order */
} PTree;

static int node_new(PTree *t, int level, int pix, int x, int y)
{
	if (t->n == t->cap) {
		t->cap = t->cap ? t->cap * 2 : 1024;
		t->nd = (PNode *)realloc(t->nd, sizeof(PNode) * (size_t)t->cap);
	}
	PNode *e = &t->nd[t->n];
	e->level = level; e->area = 1; e->x0 = e->x1 = x; e->y0 = e->y1 = y; e->seed_pix = pix;
	e->parent = e->child = e->next = -1; e->alive = 1; e->done = 0; e->stability = 0.0;
	return t->n++;
}

/* er_accumulate (src/ER.cpp:131-151): one more pixel + bbox grow */
static inline void node_add_pixel(PNode *e, int x, int y)
{
	e->area++;
	if (x < e->x0) e->x0 = x;
	if (x > e->x1) e->x1 = x;
	if (y < e->y0) e->y0 = y;
	if (y > e->y1) e->y1 = y;
}

/* er_merge (src/ER.cpp:153-191): fold child into parent; children whose area <= min_area are
 * dropped (their own children, if any, are spliced in front of the parent's list -- unreachable in
 * practice because area is monotone up the tree, restated anyway); survivors are PREPENDED. */
static void node_merge(PTree *t, int pi, int ci, int min_area)
{
	PNode *p = &t->nd[pi], *c = &t->nd[ci];
	p->area += c->area;
	if (c->x0 < p->x0) p->x0 = c->x0;
	if (c->x1 > p->x1) p->x1 = c->x1;
	if (c->y0 < p->y0) p->y0 = c->y0;
	if (c->y1 > p->y1) p->y1 = c->y1;
	if (c->area <= min_area) {
		int g = c->child;
		if (g >= 0) {
			int last = g;
			while (t->nd[last].next >= 0) last = t->nd[last].next;
			t->nd[last].next = p->child;
			p->child = g;
			t->nd[g].parent = pi;
		}
		c->alive = 0;
	} else {
		c->next = p->child;
		p->child = ci;
		c->parent = pi;
	}
}

/* ER tree extraction: Nister & Stewenius linear-time flood as used by
 * ERFilter::er_tree_extract + process_stack (src/ER.cpp:240-413).
 * Reference behaviours kept on purpose:
 *  - flood source is pixel 0; neighbour order right, bottom, left, top (inc/ER.h:147)
 *  - highest_level = 255/step + 1; boundary pixels at a level >= highest_level are queued but
 *    never popped (they act as walls), the flood ends when the priority reaches highest_level
 *  - per-level boundary queues are LIFO
 *  - the returned root is the component on top of the stack when the flood ends.
 */
typedef struct { int pix, edge, next; } QEnt;

PORT_API PTree *port_tree_extract(const uint8_t *plane, int w, int h, int stride, int step, int min_area)
{
	const int npx = w * h;
	const int hi = 255 / step + 1;
	PTree *t = (PTree *)calloc(1, sizeof(PTree));
	t->w = w; t->h = h;
	uint8_t *lev = (uint8_t *)malloc((size_t)npx);
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++) lev[y * w + x] = quant_level(plane[(size_t)y * stride + x], step);
	uint8_t *seen = (uint8_t *)calloc((size_t)npx, 1);
	/* boundary queues: one LIFO per level, entries from a shared pool (<= 5 pushes per pixel) */
	QEnt *q = (QEnt *)malloc(sizeof(QEnt) * ((size_t)npx * 5 + 8));
	int qn = 0;
	int head[257];
	for (int i = 0; i < 257; i++) head[i] = -1;
	int *stack = (int *)malloc(sizeof(int) * (size_t)(258 + 8));
	int sp = 0;
	int prio = hi;

	stack[sp++] = node_new(t, 256, 0, 0, 0);   /* sentinel above every real level */

	int cur = 0, edge = 0, cur_level = lev[0];
	seen[0] = 1;
	for (;;) {
		/* open a component for the current pixel (reference label step_3) */
		int x = cur % w, y = cur / w;
		stack[sp++] = node_new(t, cur_level, cur, x, y);
		int descended;
		for (;;) {
			descended = 0;
			for (; edge < 4; edge++) {
				int nb = cur;
				if (edge == 0) { if (x + 1 < w) nb = cur + 1; }
				else if (edge == 1) { if (y + 1 < h) nb = cur + w; }
				else if (edge == 2) { if (x > 0) nb = cur - 1; }
				else { if (y > 0) nb = cur - w; }
				if (nb == cur || seen[nb]) continue;
				seen[nb] = 1;
				const int nl = lev[nb];
				if (nl >= cur_level) {
					q[qn].pix = nb; q[qn].edge = 0; q[qn].next = head[nl]; head[nl] = qn++;
					if (nl < prio) prio = nl;
				} else {
					q[qn].pix = cur; q[qn].edge = edge + 1; q[qn].next = head[cur_level]; head[cur_level] = qn++;
					if (cur_level < prio) prio = cur_level;
					cur = nb; cur_level = nl; edge = 0;
					descended = 1;
					break;
				}
			}
			if (descended) break;
			/* water saturates the current pixel */
			node_add_pixel(&t->nd[stack[sp - 1]], x, y);
			if (prio == hi) goto finished;
			const int e = head[prio];
			const int np = q[e].pix, ne = q[e].edge;
			head[prio] = q[e].next;
			const int nlev = lev[np];
			while (prio < hi && head[prio] < 0) prio++;
			cur = np; edge = ne; x = cur % w; y = cur / w;
			if (nlev != cur_level) {
				cur_level = nlev;
				/* process_stack (src/ER.cpp:377-413) */
				do {
					const int top = stack[sp - 1], second = stack[sp - 2];
					sp--;
					if (nlev < t->nd[second].level) {
						const int wrap = node_new(t, nlev, t->nd[top].seed_pix, t->nd[top].seed_pix % w, t->nd[top].seed_pix / w);
						stack[sp++] = wrap;
						node_merge(t, wrap, top, min_area);
						break;
					}
					node_merge(t, second, top, min_area);
				} while (nlev > t->nd[stack[sp - 1]].level);
			}
		}
	}
finished:
	t->root = stack[sp - 1];
	free(lev); free(seen); free(q); free(stack);
	return t;
}

static void tree_flatten(PTree *t)
{
	if (t->flat) return;
	t->flat = (int *)malloc(sizeof(int) * (size_t)t->n);
	t->flat_of = (int *)malloc(sizeof(int) * (size_t)t->n);
	for (int i = 0; i < t->n; i++) t->flat_of[i] = -1;
	int *st = (int *)malloc(sizeof(int) * (size_t)(t->n + 1));
	int *tmp = (int *)malloc(sizeof(int) * (size_t)(t->n + 1));
	int sp = 0;
	t->nflat = 0;
	st[sp++] = t->root;
	while (sp) {
		const int id = st[--sp];
		t->flat_of[id] = t->nflat;
		t->flat[t->nflat++] = id;
		int k = 0;
		for (int c = t->nd[id].child; c >= 0; c = t->nd[c].next) tmp[k++] = c;
		while (k) st[sp++] = tmp[--k];
	}
	free(st); free(tmp);
}

PORT_API int port_tree_size(PTree *t) { tree_flatten(t); return t->nflat; }

/* out: n x 8 int32 = level, area, x, y, w, h, parent (flat index, -1 root), n_children */
PORT_API void port_tree_dump(PTree *t, int32_t *out)
{
	tree_flatten(t);
	for (int i = 0; i < t->nflat; i++) {
		const PNode *e = &t->nd[t->flat[i]];
		int nc = 0;
		for (int c = e->child; c >= 0; c = t->nd[c].next) nc++;
		int32_t *o = out + 8 * (size_t)i;
		o[0] = e->level; o[1] = e->area; o[2] = e->x0; o[3] = e->y0;
		o[4] = e->x1 - e->x0 + 1; o[5] = e->y1 - e->y0 + 1;
		o[6] = (t->flat[i] == t->root) ? -1 : t->flat_of[e->parent];
		o[7] = nc;
	}
}

/* ------------------------------------------------------------------------------------------
 * a4  ERFilter::non_maximum_supression (src/ER.cpp:416-505).
 * ------------------------------------------------------------------------------------------ */
static inline int bb_w(const PNode *e) { return e->x1 - e->x0 + 1; }
static inline int bb_h(const PNode *e) { return e->y1 - e->y0 + 1; }
static inline int bb_area(const PNode *e) { return bb_w(e) * bb_h(e); }
static inline int bb_inter_area(const PNode *a, const PNode *b)
{
	const int x0 = a->x0 > b->x0 ? a->x0 : b->x0, y0 = a->y0 > b->y0 ? a->y0 : b->y0;
	const int x1 = a->x1 < b->x1 ? a->x1 : b->x1, y1 = a->y1 < b->y1 ? a->y1 : b->y1;
	if (x1 < x0 || y1 < y0) return 0;
	return (x1 - x0 + 1) * (y1 - y0 + 1);
}

PORT_API int port_nms(PTree *t, int min_area, int max_area, int stability_t, double overlap_coef)
{
	PNode *nd = t->nd;
	free(t->pool);
	t->pool = (int *)malloc(sizeof(int) * (size_t)(t->n + 1));
	t->npool = 0;
	int *st = (int *)malloc(sizeof(int) * (size_t)(t->n + 2));
	int *chain = (int *)malloc(sizeof(int) * (size_t)(t->n + 2));
	int sp = 0;
	nd[t->root].parent = t->root;   /* src/ER.cpp:424 */
	int cur = t->root;
	for (;;) {
		for (; cur >= 0; cur = nd[cur].child) st[sp++] = cur;
		if (sp == 0) break;
		cur = st[--sp];
		if (!nd[cur].done) {
			int len = 0, p = cur;
			while ((double)bb_inter_area(&nd[cur], &nd[p]) / (double)bb_area(&nd[p]) > overlap_coef && !nd[p].done) {
				nd[p].done = 1;
				chain[len++] = p;
				p = nd[p].parent;
			}
			if (len >= 1 + stability_t) {
				for (int i = 0; i < len - stability_t; i++)
					nd[chain[i]].stability = (double)bb_area(&nd[chain[i]]) /
					                         (double)(bb_area(&nd[chain[i + stability_t]]) - bb_area(&nd[chain[i]]));
				int best = 0;
				for (int i = 1; i < len - stability_t; i++) {
					if (nd[chain[i]].stability > nd[chain[best]].stability) best = i;
					else if (nd[chain[i]].stability == nd[chain[best]].stability)
						best = (bb_area(&nd[chain[i]]) < bb_area(&nd[chain[best]])) ? i : best;
				}
				const PNode *b = &nd[chain[best]];
				const double ar = (double)bb_w(b) / (double)bb_h(b);
				if (ar < 2.0 && ar > 0.10 && b->area < max_area && b->area > min_area &&
				    bb_h(b) < t->h * 0.8 && bb_w(b) < t->w * 0.8)
					t->pool[t->npool++] = chain[best];
			}
		}
		cur = nd[cur].next;
	}
	free(st); free(chain);
	return t->npool;
}


/* Re-order every child list by a canonical key (experiments on how well a traversal-independent
 * sibling order reproduces the reference's NMS output).  mode:
 *  1: descending (y0, x0)   2: descending (y1, x1)   3: descending seed pixel   4: ascending (y0,x0)
 *  5: descending (y0,x0) after ascending level   6: reverse the reference order */
typedef struct { long long k; int id; } SortEnt;
static int sortent_cmp(const void *a, const void *b)
{
	const SortEnt *x = (const SortEnt *)a, *y = (const SortEnt *)b;
	if (x->k != y->k) return x->k < y->k ? -1 : 1;
	return x->id < y->id ? -1 : (x->id > y->id);
}
static PTree *g_sort_tree;
static int sortent_cmp9(const void *a, const void *b)
{
	const SortEnt *x = (const SortEnt *)a, *y = (const SortEnt *)b;
	if (x->k != y->k) return x->k < y->k ? -1 : 1;
	const PNode *p = &g_sort_tree->nd[x->id], *q = &g_sort_tree->nd[y->id];
	if (p->area != q->area) return p->area > q->area ? -1 : 1;
	if (p->x1 != q->x1) return p->x1 > q->x1 ? -1 : 1;
	if (p->y1 != q->y1) return p->y1 > q->y1 ? -1 : 1;
	return x->id < y->id ? -1 : (x->id > y->id);
}
PORT_API void port_sort_children(PTree *t, int mode)
{
	g_sort_tree = t;
	SortEnt *tmp = (SortEnt *)malloc(sizeof(SortEnt) * (size_t)(t->n + 1));
	for (int i = 0; i < t->n; i++) {
		if (!t->nd[i].alive || t->nd[i].child < 0) continue;
		int k = 0;
		for (int c = t->nd[i].child; c >= 0; c = t->nd[c].next) {
			const PNode *e = &t->nd[c];
			long long key = 0;
			switch (mode) {
			case 1: key = -((long long)e->y0 * 100000 + e->x0); break;
			case 2: key = -((long long)e->y1 * 100000 + e->x1); break;
			case 3: key = -(long long)e->seed_pix; break;
			case 4: key = ((long long)e->y0 * 100000 + e->x0); break;
			case 5: key = (long long)e->level * 10000000000LL - ((long long)e->y0 * 100000 + e->x0); break;
			case 6: key = -(long long)k; break;
			case 7: key = -((long long)e->y1 * 100000 + e->x0); break;
			case 8: key = -((long long)e->y0 * 100000 + e->x1); break;
			case 9: /* the GPU path's canonical order: descending (y0, x0, level), then area, x1, y1 (see sortent_cmp9) */
				key = -(((long long)e->y0 * 8192 + e->x0) * 64 + e->level); break;
			default: key = k; break;
			}
			tmp[k].k = key; tmp[k].id = c; k++;
		}
		qsort(tmp, (size_t)k, sizeof(SortEnt), mode == 9 ? sortent_cmp9 : sortent_cmp);
		t->nd[i].child = tmp[0].id;
		for (int j = 0; j < k; j++) t->nd[tmp[j].id].next = (j + 1 < k) ? tmp[j + 1].id : -1;
	}
	free(tmp);
	free(t->flat); free(t->flat_of); t->flat = NULL; t->flat_of = NULL;
}

PORT_API void port_pool_indices(PTree *t, int32_t *out)
{
	tree_flatten(t);
	for (int i = 0; i < t->npool; i++) out[i] = t->flat_of[t->pool[i]];
}

PORT_API void port_tree_free(PTree *t)
{
	if (!t) return;
	free(t->nd); free(t->flat); free(t->flat_of); free(t->pool); free(t);
}

/* ------------------------------------------------------------------------------------------
 * a5  OCR::ARAN (src/OCR.cpp:394-430) on top of cv::resize (INTER_LINEAR, 8UC1).
 * The resize restatement follows OpenCV 4.x resizeGeneric_/HResizeLinear/VResizeLinear
 * fixed-point arithmetic (INTER_RESIZE_COEF_BITS = 11) -- pinned against cv2 in the tests.
 * ------------------------------------------------------------------------------------------ */
PORT_API void port_resize(const uint8_t *src, int sw, int sh, int stride, int dw, int dh, uint8_t *dst)
{
	if (sw == 2 * dw && sh == 2 * dh) {   /* INTER_LINEAR -> INTER_AREA for exact 2x decimation */
		for (int y = 0; y < dh; y++) {
			const uint8_t *a = src + (size_t)(2 * y) * stride, *b = a + stride;
			for (int x = 0; x < dw; x++)
				dst[y * dw + x] = (uint8_t)((a[2 * x] + a[2 * x + 1] + b[2 * x] + b[2 * x + 1] + 2) >> 2);
		}
		return;
	}
	const double sx = 1.0 / ((double)dw / sw), sy = 1.0 / ((double)dh / sh);
	for (int y = 0; y < dh; y++) {
		float fy = (float)((y + 0.5) * sy - 0.5);
		int iy = (int)floorf(fy);
		fy -= (float)iy;
		const int wy0 = (int)lrintf((1.f - fy) * 2048.f), wy1 = (int)lrintf(fy * 2048.f);
		int r0 = iy, r1 = iy + 1;
		if (r0 < 0) r0 = 0; if (r0 > sh - 1) r0 = sh - 1;
		if (r1 < 0) r1 = 0; if (r1 > sh - 1) r1 = sh - 1;
		const uint8_t *p0 = src + (size_t)r0 * stride, *p1 = src + (size_t)r1 * stride;
		for (int x = 0; x < dw; x++) {
			float fx = (float)((x + 0.5) * sx - 0.5);
			int ix = (int)floorf(fx);
			fx -= (float)ix;
			if (ix < 0) { ix = 0; fx = 0.f; }
			if (ix >= sw - 1) { ix = sw - 1; fx = 0.f; }
			const int wx0 = (int)lrintf((1.f - fx) * 2048.f), wx1 = (int)lrintf(fx * 2048.f);
			const int ix1 = ix + 1 < sw ? ix + 1 : sw - 1;
			const int h0 = p0[ix] * wx0 + p0[ix1] * wx1;
			const int h1 = p1[ix] * wx0 + p1[ix1] * wx1;
			int v = (((wy0 * (h0 >> 4)) >> 16) + ((wy1 * (h1 >> 4)) >> 16) + 2) >> 2;
			dst[y * dw + x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
		}
	}
}

/* minor side of the ARAN target: (int)(L * pow(min/max, para)) with glibc pow, para = 0.5 */
PORT_API int port_aran_minor(int w, int h, int L)
{
	const double r1 = (w > h) ? (double)h / w : (double)w / h;
	return (int)(L * pow(r1, 0.5));
}

PORT_API void port_aran(const uint8_t *src, int w, int h, int stride, int L, uint8_t *dstLL)
{
	const int minor = port_aran_minor(w, h, L);
	const int dw = (w > h) ? L : minor, dh = (w > h) ? minor : L;
	uint8_t *tmp = (uint8_t *)malloc((size_t)(dw * dh > 0 ? dw * dh : 1));
	memset(dstLL, 0, (size_t)L * L);
	if (dw > 0 && dh > 0) {
		port_resize(src, w, h, stride, dw, dh, tmp);
		if (dw > dh) {
			const int off = (int)round((double)((L - dh) / 2));
			for (int i = 0; i < dh; i++) memcpy(dstLL + (size_t)(i + off) * L, tmp + (size_t)i * dw, (size_t)dw);
		} else {
			const int off = (int)round((double)((L - dw) / 2));
			for (int i = 0; i < dh; i++) memcpy(dstLL + (size_t)i * L + off, tmp + (size_t)i * dw, (size_t)dw);
		}
	}
	free(tmp);
}

/* a6/a7  calc_LBP + make_LBP_hist (src/ER.cpp:789-845).  The reference indexes the 26-stride
 * buffer with 24-stride neighbour offsets; restated literally: offsets relative to
 * base = (i+1)*26 + 1 + j are {-25,-24,-23,+1,+25,+24,+23,-1}; bit k set iff 8*v_k > sum. */
PORT_API void port_lbp_hist(const uint8_t *crop, int w, int h, int stride, double *hist1024)
{
	uint8_t buf[26 * 26 + 64];
	static const int off[8] = { -25, -24, -23, 1, 25, 24, 23, -1 };
	memset(buf, 0, sizeof buf);
	port_aran(crop, w, h, stride, 26, buf);
	for (int i = 0; i < 1024; i++) hist1024[i] = 0.0;
	for (int i = 0; i < 24; i++)
		for (int j = 0; j < 24; j++) {
			const int base = (i + 1) * 26 + 1 + j;
			int s = 0, code = 0;
			for (int k = 0; k < 8; k++) s += buf[base + off[k]];
			for (int k = 0; k < 8; k++) if (8 * buf[base + off[k]] > s) code |= 1 << k;
			hist1024[(i / 12) * 512 + (j / 12) * 256 + code] += 1.0;
		}
}

/* ------------------------------------------------------------------------------------------
 * a9  CascadeBoost (REAL, decision stumps): loader src/adaboost.cpp:873-951, predict :507-542.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
	int n_stages;
	int *stage_len;
	int *stage_thr;           /* (int)stod(...)  (src/adaboost.cpp:919) */
	int n_stumps;
	int *dim;
	double *thr, *cp, *cn, *weight;
	int is_real;
} PCascade;

PORT_API PCascade *port_cascade_load(const char *path)
{
	FILE *f = fopen(path, "rb");
	if (!f) return NULL;
	fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
	char *txt = (char *)malloc((size_t)sz + 1);
	if (fread(txt, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(txt); return NULL; }
	txt[sz] = 0; fclose(f);
	PCascade *c = (PCascade *)calloc(1, sizeof(PCascade));
	c->is_real = 1;
	char *p = txt;
	int cap_st = 0;
	/* header: four whitespace separated records */
	char *line_end;
	for (int ln = 0; ln < 4; ln++) {
		line_end = strchr(p, '\n');
		if (!line_end) line_end = txt + sz;
		char save = *line_end; *line_end = 0;
		char *tok = strtok(p, " \t\r");
		if (tok && !strcmp(tok, "boost_type")) { tok = strtok(NULL, " \t\r"); c->is_real = !(tok && !strcmp(tok, "DISCRETE")); }
		else if (tok && !strcmp(tok, "num_of_iter")) {
			while ((tok = strtok(NULL, " \t\r"))) {
				if (c->n_stages == cap_st) { cap_st = cap_st ? 2 * cap_st : 8; c->stage_len = (int *)realloc(c->stage_len, sizeof(int) * cap_st); }
				c->stage_len[c->n_stages++] = (int)strtod(tok, NULL);
			}
		} else if (tok && !strcmp(tok, "threshold")) {
			c->stage_thr = (int *)calloc((size_t)c->n_stages + 1, sizeof(int));
			for (int j = 0; j < c->n_stages && (tok = strtok(NULL, " \t\r")); j++) c->stage_thr[j] = (int)strtod(tok, NULL);
		}
		*line_end = save;
		p = (save == 0) ? line_end : line_end + 1;
	}
	int cap = 0;
	while (*p) {
		line_end = strchr(p, '\n');
		if (!line_end) line_end = p + strlen(p);
		char save = *line_end; *line_end = 0;
		double v[5]; int k = 0;
		char *s = p, *e;
		while (k < 5) { double d = strtod(s, &e); if (e == s) break; v[k++] = d; s = e; }
		if (k == 5) {
			if (c->n_stumps == cap) {
				cap = cap ? cap * 2 : 4096;
				c->dim = (int *)realloc(c->dim, sizeof(int) * cap);
				c->thr = (double *)realloc(c->thr, sizeof(double) * cap);
				c->cp = (double *)realloc(c->cp, sizeof(double) * cap);
				c->cn = (double *)realloc(c->cn, sizeof(double) * cap);
				c->weight = (double *)realloc(c->weight, sizeof(double) * cap);
			}
			const int i = c->n_stumps++;
			c->weight[i] = v[0]; c->dim[i] = (int)v[1]; c->thr[i] = v[2]; c->cp[i] = v[3]; c->cn[i] = v[4];
		}
		*line_end = save;
		p = (save == 0) ? line_end : line_end + 1;
	}
	free(txt);
	return c;
}

PORT_API int port_cascade_info(const PCascade *c, int *n_stages, int *stage_len, int *stage_thr)
{
	if (n_stages) *n_stages = c->n_stages;
	for (int i = 0; i < c->n_stages; i++) { if (stage_len) stage_len[i] = c->stage_len[i]; if (stage_thr) stage_thr[i] = c->stage_thr[i]; }
	return c->n_stumps;
}

/* REAL branch of CascadeBoost::predict: per stage, in-order FP64 sum of cp/cn; reject (-DBL_MAX)
 * when the stage sum is below the stage's integer threshold; else the LAST stage's sum. */
PORT_API double port_cascade_predict(const PCascade *c, const double *fv)
{
	double score = 0;
	int off = 0;
	for (int s = 0; s < c->n_stages; s++) {
		score = 0;
		for (int j = off; j < off + c->stage_len[s]; j++) score += (fv[c->dim[j]] < c->thr[j]) ? c->cp[j] : c->cn[j];
		if (score < c->stage_thr[s]) return -DBL_MAX;
		off += c->stage_len[s];
	}
	return score;
}

PORT_API void port_cascade_predict_batch(const PCascade *c, const double *fv, int n, int dims, double *score)
{
	for (int i = 0; i < n; i++) score[i] = port_cascade_predict(c, fv + (size_t)i * dims);
}

PORT_API void port_cascade_free(PCascade *c)
{
	if (!c) return;
	free(c->stage_len); free(c->stage_thr); free(c->dim); free(c->thr); free(c->cp); free(c->cn); free(c->weight); free(c);
}

/* a8  ERFilter::classify (src/ER.cpp:507-528): strong first, weak only if strong rejected. */
PORT_API void port_classify(PTree *t, const uint8_t *plane, int stride, const PCascade *strong, const PCascade *weak,
                            int32_t *label, double *strong_score, double *weak_score)
{
	double fv[1024];
	for (int i = 0; i < t->npool; i++) {
		const PNode *e = &t->nd[t->pool[i]];
		port_lbp_hist(plane + (size_t)e->y0 * stride + e->x0, bb_w(e), bb_h(e), stride, fv);
		const double s = port_cascade_predict(strong, fv);
		const double wv = port_cascade_predict(weak, fv);
		label[i] = (s > -DBL_MAX) ? 2 : ((wv > -DBL_MAX) ? 1 : 0);
		if (strong_score) strong_score[i] = s;
		if (weak_score) weak_score[i] = wv;
	}
}

/* ------------------------------------------------------------------------------------------
 * a10  libsvm 3.21 C-SVC / RBF probability prediction.
 * loader src/svm.cpp:2767-2982, k_function RBF :325-365, svm_predict_values :2501-2575,
 * sigmoid_predict :1818-1826, multiclass_probability :1829-1890, svm_predict_probability :2592-2629.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
	int nr_class, l;
	double gamma;
	double *rho, *probA, *probB;
	int *label, *nsv;
	double *coef;        /* (nr_class-1) x l */
	int *sv_start;       /* l+1 offsets into sv_idx / sv_val */
	int *sv_idx; double *sv_val;
} PSvm;

static char *next_line(char *p, char **line)
{
	*line = p;
	char *e = strchr(p, '\n');
	if (!e) return p + strlen(p);
	*e = 0;
	return e + 1;
}

PORT_API PSvm *port_svm_load(const char *path)
{
	FILE *f = fopen(path, "rb");
	if (!f) return NULL;
	fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
	char *txt = (char *)malloc((size_t)sz + 1);
	if (fread(txt, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(txt); return NULL; }
	txt[sz] = 0; fclose(f);
	PSvm *m = (PSvm *)calloc(1, sizeof(PSvm));
	char *p = txt, *line;
	for (;;) {
		p = next_line(p, &line);
		char *key = strtok(line, " \t\r");
		if (!key) { if (!*p) break; continue; }
		if (!strcmp(key, "SV")) break;
		if (!strcmp(key, "gamma")) m->gamma = strtod(strtok(NULL, " \t\r"), NULL);
		else if (!strcmp(key, "nr_class")) m->nr_class = atoi(strtok(NULL, " \t\r"));
		else if (!strcmp(key, "total_sv")) m->l = atoi(strtok(NULL, " \t\r"));
		else if (!strcmp(key, "rho") || !strcmp(key, "probA") || !strcmp(key, "probB")) {
			const int n = m->nr_class * (m->nr_class - 1) / 2;
			double *a = (double *)calloc((size_t)n, sizeof(double));
			for (int i = 0; i < n; i++) { char *tk = strtok(NULL, " \t\r"); a[i] = tk ? strtod(tk, NULL) : 0; }
			if (key[0] == 'r') m->rho = a; else if (key[4] == 'A') m->probA = a; else m->probB = a;
		} else if (!strcmp(key, "label") || !strcmp(key, "nr_sv")) {
			int *a = (int *)calloc((size_t)m->nr_class, sizeof(int));
			for (int i = 0; i < m->nr_class; i++) { char *tk = strtok(NULL, " \t\r"); a[i] = tk ? atoi(tk) : 0; }
			if (key[0] == 'l') m->label = a; else m->nsv = a;
		}
		/* svm_type / kernel_type: this restatement covers c_svc + rbf only */
	}
	const int k1 = m->nr_class - 1;
	m->coef = (double *)calloc((size_t)k1 * m->l, sizeof(double));
	m->sv_start = (int *)calloc((size_t)m->l + 1, sizeof(int));
	size_t cap = 1 << 20, nnz = 0;
	m->sv_idx = (int *)malloc(sizeof(int) * cap);
	m->sv_val = (double *)malloc(sizeof(double) * cap);
	for (int i = 0; i < m->l; i++) {
		p = next_line(p, &line);
		char *s = line, *e;
		for (int j = 0; j < k1; j++) { m->coef[(size_t)j * m->l + i] = strtod(s, &e); s = e; }
		m->sv_start[i] = (int)nnz;
		for (;;) {
			while (*s == ' ' || *s == '\t' || *s == '\r') s++;
			if (!*s) break;
			long idx = strtol(s, &e, 10);
			if (e == s || *e != ':') break;
			s = e + 1;
			double v = strtod(s, &e);
			s = e;
			if (nnz == cap) { cap *= 2; m->sv_idx = (int *)realloc(m->sv_idx, sizeof(int) * cap); m->sv_val = (double *)realloc(m->sv_val, sizeof(double) * cap); }
			m->sv_idx[nnz] = (int)idx; m->sv_val[nnz] = v; nnz++;
		}
	}
	m->sv_start[m->l] = (int)nnz;
	free(txt);
	return m;
}

PORT_API void port_svm_info(const PSvm *m, int *nr_class, int *l, double *gamma, int *nnz)
{
	if (nr_class) *nr_class = m->nr_class;
	if (l) *l = m->l;
	if (gamma) *gamma = m->gamma;
	if (nnz) *nnz = m->sv_start[m->l];
}

/* densified SV matrix (l x dims doubles), coefficient matrix ((nr_class-1) x l) etc. for tests */
PORT_API void port_svm_dense_sv(const PSvm *m, int dims, double *out)
{
	memset(out, 0, sizeof(double) * (size_t)m->l * dims);
	for (int i = 0; i < m->l; i++)
		for (int k = m->sv_start[i]; k < m->sv_start[i + 1]; k++)
			if (m->sv_idx[k] >= 0 && m->sv_idx[k] < dims) out[(size_t)i * dims + m->sv_idx[k]] = m->sv_val[k];
}

static double sigmoid_of(double dec, double A, double B)
{
	const double f = dec * A + B;
	if (f >= 0) return exp(-f) / (1.0 + exp(-f));
	return 1.0 / (1 + exp(f));
}

/* Wu, Lin & Weng "method 2" pairwise coupling, as libsvm's multiclass_probability */
static void couple(int k, const double *r /*k x k*/, double *p)
{
	int max_iter = k > 100 ? k : 100, iter;
	double *Q = (double *)malloc(sizeof(double) * (size_t)k * k);
	double *Qp = (double *)malloc(sizeof(double) * (size_t)k);
	const double eps = 0.005 / k;
	for (int t = 0; t < k; t++) {
		p[t] = 1.0 / k;
		Q[t * k + t] = 0;
		for (int j = 0; j < t; j++) { Q[t * k + t] += r[j * k + t] * r[j * k + t]; Q[t * k + j] = Q[j * k + t]; }
		for (int j = t + 1; j < k; j++) { Q[t * k + t] += r[j * k + t] * r[j * k + t]; Q[t * k + j] = -r[j * k + t] * r[t * k + j]; }
	}
	for (iter = 0; iter < max_iter; iter++) {
		double pQp = 0;
		for (int t = 0; t < k; t++) {
			Qp[t] = 0;
			for (int j = 0; j < k; j++) Qp[t] += Q[t * k + j] * p[j];
			pQp += p[t] * Qp[t];
		}
		double max_err = 0;
		for (int t = 0; t < k; t++) { const double er = fabs(Qp[t] - pQp); if (er > max_err) max_err = er; }
		if (max_err < eps) break;
		for (int t = 0; t < k; t++) {
			const double diff = (-Qp[t] + pQp) / Q[t * k + t];
			p[t] += diff;
			pQp = (pQp + diff * (diff * Q[t * k + t] + 2 * Qp[t])) / (1 + diff) / (1 + diff);
			for (int j = 0; j < k; j++) { Qp[j] = (Qp[j] + diff * Q[t * k + j]) / (1 + diff); p[j] /= (1 + diff); }
		}
	}
	free(Q); free(Qp);
}

/* x: dense doubles, zeros omitted (OCR::extract_feature emits only non-zero entries, 0-based
 * index = position, src/OCR.cpp:203-218).  Returns the label; prob[nr_class]; optional
 * kvalue[l] and dec[nr_class*(nr_class-1)/2] for intermediate-stage parity tests. */
PORT_API double port_svm_predict_probability(const PSvm *m, const double *x, int dims, double *prob, double *kvalue_out, double *dec_out)
{
	const int k = m->nr_class, l = m->l;
	int *xi = (int *)malloc(sizeof(int) * (size_t)(dims + 1));
	double *xv = (double *)malloc(sizeof(double) * (size_t)(dims + 1));
	int xn = 0;
	for (int d = 0; d < dims; d++) if (x[d] != 0) { xi[xn] = d; xv[xn] = x[d]; xn++; }
	double *kv = (double *)malloc(sizeof(double) * (size_t)l);
	for (int i = 0; i < l; i++) {
		/* RBF: sparse merge of squared differences, then exp(-gamma*sum) */
		double sum = 0;
		int a = 0, b = m->sv_start[i];
		const int be = m->sv_start[i + 1];
		while (a < xn && b < be) {
			if (xi[a] == m->sv_idx[b]) { const double d = xv[a] - m->sv_val[b]; sum += d * d; a++; b++; }
			else if (xi[a] > m->sv_idx[b]) { sum += m->sv_val[b] * m->sv_val[b]; b++; }
			else { sum += xv[a] * xv[a]; a++; }
		}
		while (a < xn) { sum += xv[a] * xv[a]; a++; }
		while (b < be) { sum += m->sv_val[b] * m->sv_val[b]; b++; }
		kv[i] = exp(-m->gamma * sum);
	}
	int *start = (int *)malloc(sizeof(int) * (size_t)k);
	start[0] = 0;
	for (int i = 1; i < k; i++) start[i] = start[i - 1] + m->nsv[i - 1];
	double *r = (double *)calloc((size_t)k * k, sizeof(double));
	int pidx = 0;
	for (int i = 0; i < k; i++)
		for (int j = i + 1; j < k; j++) {
			double sum = 0;
			const double *c1 = m->coef + (size_t)(j - 1) * l, *c2 = m->coef + (size_t)i * l;
			for (int q = 0; q < m->nsv[i]; q++) sum += c1[start[i] + q] * kv[start[i] + q];
			for (int q = 0; q < m->nsv[j]; q++) sum += c2[start[j] + q] * kv[start[j] + q];
			sum -= m->rho[pidx];
			if (dec_out) dec_out[pidx] = sum;
			double pr = sigmoid_of(sum, m->probA[pidx], m->probB[pidx]);
			const double lo = 1e-7;
			if (pr < lo) pr = lo;
			if (pr > 1 - lo) pr = 1 - lo;
			r[i * k + j] = pr;
			r[j * k + i] = 1 - pr;
			pidx++;
		}
	couple(k, r, prob);
	int best = 0;
	for (int i = 1; i < k; i++) if (prob[i] > prob[best]) best = i;
	if (kvalue_out) memcpy(kvalue_out, kv, sizeof(double) * (size_t)l);
	const double lab = (double)m->label[best];
	free(xi); free(xv); free(kv); free(start); free(r);
	return lab;
}

PORT_API void port_svm_free(PSvm *m)
{
	if (!m) return;
	free(m->rho); free(m->probA); free(m->probB); free(m->label); free(m->nsv);
	free(m->coef); free(m->sv_start); free(m->sv_idx); free(m->sv_val); free(m);
}

/* ------------------------------------------------------------------------------------------
 * Canonical (order-free) statement of the node set the flood produces, used to pin the GPU
 * kernel's node multiset independently of traversal order:
 *   node (L, C)  <=>  C is a 4-connected component of {level <= L} inside the reach set and C
 *   holds at least one pixel of level exactly L;   area = |C| + #nodes in the subtree.
 * Computed here by a plain union-find over pixels sorted by level (Kruskal style).
 * out rows: level, area, x, y, w, h ; returns number of rows written (<= cap), kept rule
 * area > min_area or root.  Rows are emitted in no particular order (compare as multisets).
 * ------------------------------------------------------------------------------------------ */
static int uf_find(int *par, int x)
{
	while (par[x] != x) { par[x] = par[par[x]]; x = par[x]; }
	return x;
}

PORT_API int port_canonical_nodes(const uint8_t *plane, int w, int h, int stride, int step, int min_area, int32_t *out, int cap)
{
	const int npx = w * h, hi = 255 / step + 1;
	uint8_t *lev = (uint8_t *)malloc((size_t)npx);
	for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) lev[y * w + x] = quant_level(plane[(size_t)y * stride + x], step);
	/* reach: component of {lev < hi} entered by the flood from pixel 0 (src/ER.cpp:267-341) */
	int start = -1;
	if (lev[0] < hi) start = 0;
	else if (w > 1 && lev[1] < hi) start = 1;
	else if (h > 1 && lev[w] < hi) start = w;
	int nout = 0;
	if (start < 0) {
		if (cap > 0) { out[0] = lev[0]; out[1] = 2; out[2] = 0; out[3] = 0; out[4] = 1; out[5] = 1; }
		free(lev);
		return 1;
	}
	uint8_t *reach = (uint8_t *)calloc((size_t)npx, 1);
	int *stk = (int *)malloc(sizeof(int) * (size_t)npx);
	int sp = 0;
	stk[sp++] = start; reach[start] = 1;
	while (sp) {
		const int p = stk[--sp], x = p % w, y = p / w;
		const int nb[4] = { x + 1 < w ? p + 1 : -1, y + 1 < h ? p + w : -1, x > 0 ? p - 1 : -1, y > 0 ? p - w : -1 };
		for (int k = 0; k < 4; k++) if (nb[k] >= 0 && !reach[nb[k]] && lev[nb[k]] < hi) { reach[nb[k]] = 1; stk[sp++] = nb[k]; }
	}
	/* counting sort of reach pixels by level */
	int cnt[257]; memset(cnt, 0, sizeof cnt);
	for (int p = 0; p < npx; p++) if (reach[p]) cnt[lev[p] + 1]++;
	for (int i = 0; i < 256; i++) cnt[i + 1] += cnt[i];
	int *order = stk; /* reuse */
	int pos[257]; memcpy(pos, cnt, sizeof pos);
	int total = 0;
	for (int p = 0; p < npx; p++) if (reach[p]) { order[pos[lev[p]]++] = p; total++; }
	int *par = (int *)malloc(sizeof(int) * (size_t)npx);
	int *npix = (int *)calloc((size_t)npx, sizeof(int));
	int *nnode = (int *)calloc((size_t)npx, sizeof(int));
	int *bx0 = (int *)malloc(sizeof(int) * (size_t)npx), *bx1 = (int *)malloc(sizeof(int) * (size_t)npx);
	int *by0 = (int *)malloc(sizeof(int) * (size_t)npx), *by1 = (int *)malloc(sizeof(int) * (size_t)npx);
	int *stamp = (int *)malloc(sizeof(int) * (size_t)npx);
	uint8_t *act = (uint8_t *)calloc((size_t)npx, 1);
	for (int p = 0; p < npx; p++) { par[p] = p; stamp[p] = -1; }
	int maxlev = 0;
	for (int i = 0; i < total; ) {
		const int L = lev[order[i]];
		int j = i;
		while (j < total && lev[order[j]] == L) j++;
		maxlev = L;
		for (int q = i; q < j; q++) {
			const int p = order[q], x = p % w, y = p / w;
			act[p] = 1; npix[p] = 1; bx0[p] = bx1[p] = x; by0[p] = by1[p] = y;
			const int nb[4] = { x + 1 < w ? p + 1 : -1, y + 1 < h ? p + w : -1, x > 0 ? p - 1 : -1, y > 0 ? p - w : -1 };
			for (int k = 0; k < 4; k++) {
				if (nb[k] < 0 || !act[nb[k]]) continue;
				int a = uf_find(par, p), b = uf_find(par, nb[k]);
				if (a == b) continue;
				par[b] = a;
				npix[a] += npix[b]; nnode[a] += nnode[b];
				if (bx0[b] < bx0[a]) bx0[a] = bx0[b];
				if (bx1[b] > bx1[a]) bx1[a] = bx1[b];
				if (by0[b] < by0[a]) by0[a] = by0[b];
				if (by1[b] > by1[a]) by1[a] = by1[b];
			}
		}
		/* every component that received a level-L pixel is a node at level L */
		for (int q = i; q < j; q++) {
			const int r = uf_find(par, order[q]);
			if (stamp[r] == L) continue;
			stamp[r] = L;
			nnode[r] += 1;
		}
		const int is_top = (j == total);
		for (int q = i; q < j; q++) {
			const int r = uf_find(par, order[q]);
			if (stamp[r] != L) continue;
			stamp[r] = -2 - L;  /* emitted */
			const int area = npix[r] + nnode[r];
			if (area > min_area || is_top) {
				if (nout < cap) {
					int32_t *o = out + 6 * (size_t)nout;
					o[0] = L; o[1] = area; o[2] = bx0[r]; o[3] = by0[r]; o[4] = bx1[r] - bx0[r] + 1; o[5] = by1[r] - by0[r] + 1;
				}
				nout++;
			}
		}
		i = j;
	}
	(void)maxlev;
	free(lev); free(reach); free(stk); free(par); free(npix); free(nnode);
	free(bx0); free(bx1); free(by0); free(by1); free(stamp); free(act);
	return nout;
}

/* ==========================================================================================
 * Rows AFTER the detect path: er_track + calc_color, OCR::chain_run's feature path.
 * Restated from the reference's behaviour; the OpenCV primitives it calls (threshold OTSU, findContours,
 * GaussianBlur, normalize, resize) are restated from OpenCV 4.x and pinned against cv2 in tests/test_oracle_next.py
 * through the reference-backed oracle, against which every function below is checked (tests/test_port_next.py).
 * ========================================================================================== */

/* cv::threshold(..., THRESH_OTSU) -> getThreshVal_Otsu_8u: running-mean recurrence in double, strict '>' */
static int otsu_from_hist(const int *h, int total)
{
	double mu = 0, scale = 1. / total;
	for (int i = 0; i < 256; i++) mu += i * (double)h[i];
	mu *= scale;
	double mu1 = 0, q1 = 0, max_sigma = 0;
	int max_val = 0;
	for (int i = 0; i < 256; i++) {
		const double p_i = h[i] * scale;
		mu1 *= q1;
		q1 += p_i;
		const double q2 = 1. - q1;
		if (fmin(q1, q2) < FLT_EPSILON || fmax(q1, q2) > 1. - FLT_EPSILON) continue;
		mu1 = (mu1 + i * p_i) / q1;
		const double mu2 = (mu - q1 * mu1) / q2;
		const double sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
		if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
	}
	return max_val;
}

/* OTSU threshold of (255 - crop), the image both calc_color (src/ER.cpp:1395) and chain_run (src/OCR.cpp:72) binarise */
static int otsu_of_inverted(const uint8_t *crop, int w, int h, int stride)
{
	int hist[256];
	memset(hist, 0, sizeof hist);
	for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) hist[255 - crop[(size_t)y * stride + x]]++;
	return otsu_from_hist(hist, w * h);
}

/* calc_color (src/ER.cpp:1391-1437): mean YCrCb under the OTSU mask of 255 - channel(bound).  The colour rows and
 * columns are counted from the IMAGE origin, not from the bound (color_img.ptr(i), src/ER.cpp:1404).  0/0 -> NaN. */
PORT_API void port_calc_color(const uint8_t *plane, const uint8_t *ycrcb, int W, int H, const int32_t *rects, int n, double *color3)
{
	(void)H;
	for (int i = 0; i < n; i++) {
		const int x0 = rects[4 * i], y0 = rects[4 * i + 1], w = rects[4 * i + 2], h = rects[4 * i + 3];
		const uint8_t *crop = plane + (size_t)y0 * W + x0;
		const int thr = otsu_of_inverted(crop, w, h, W);
		int count = 0;
		double c1 = 0, c2 = 0, c3 = 0;
		for (int r = 0; r < h; r++) {
			const uint8_t *cp = ycrcb + (size_t)r * W * 3;
			for (int j = 0; j < w; j++)
				if (255 - crop[(size_t)r * W + j] > thr) { ++count; c1 += cp[3 * j]; c2 += cp[3 * j + 1]; c3 += cp[3 * j + 2]; }
		}
		color3[3 * i] = c1 / count; color3[3 * i + 1] = c2 / count; color3[3 * i + 2] = c3 / count;
	}
}

/* ERFilter::er_track (src/ER.cpp:532-609).  strong / weak: rows (ch, x, y, w, h, area), channel-major.  all_er starts as
 * the strong rows; every entry, in order, pulls in the not-yet-tracked weak rows that pass the geometry / colour / area
 * test (src/ER.cpp:579-590), in (channel, index) order.  tracked_out: (kind 0 strong / 1 weak, row).  Returns its length. */
PORT_API int port_er_track(const uint8_t *planes6, const uint8_t *ycrcb, int W, int H, const int32_t *strong, int ns, const int32_t *weak, int nw,
                           int32_t *tracked_out, double *strong_color, double *weak_color, int32_t *strong_center, int32_t *weak_center)
{
	const size_t np = (size_t)W * H;
	for (int pass = 0; pass < 2; pass++) {
		const int32_t *rows = pass ? weak : strong;
		const int n = pass ? nw : ns;
		double *col = pass ? weak_color : strong_color;
		int32_t *cen = pass ? weak_center : strong_center;
		for (int i = 0; i < n; i++) {
			const int32_t *r = rows + 6 * i;
			port_calc_color(planes6 + (size_t)r[0] * np, ycrcb, W, H, r + 1, 1, col + 3 * i);
			cen[2 * i] = r[1] + r[3] / 2; cen[2 * i + 1] = r[2] + r[4] / 2;          /* src/ER.cpp:545 */
		}
	}
	int len = 0;
	for (int i = 0; i < ns; i++) { tracked_out[2 * len] = 0; tracked_out[2 * len + 1] = i; len++; }
	uint8_t *taken = (uint8_t *)calloc((size_t)(nw > 0 ? nw : 1), 1);
	for (int i = 0; i < len; i++) {
		const int kind = tracked_out[2 * i], idx = tracked_out[2 * i + 1];
		const int32_t *s = (kind ? weak : strong) + 6 * idx;
		const double *sc = (kind ? weak_color : strong_color) + 3 * idx;
		const int32_t *scen = (kind ? weak_center : strong_center) + 2 * idx;
		for (int n = 0; n < nw; n++) {       /* rows are channel-major: this IS the (m, n) loop order of the reference */
			if (taken[n]) continue;
			const int32_t *w = weak + 6 * n;
			const double *wc = weak_color + 3 * n;
			const int32_t *wcen = weak_center + 2 * n;
			const int sw = s[3], sh = s[4], ww = w[3], wh = w[4];
			if (abs(scen[0] - wcen[0]) + abs(scen[1] - wcen[1]) < ((sw > sh ? sw : sh) << 1) &&
			    abs(sh - wh) < (sh < wh ? sh : wh) &&
			    abs(sw - ww) < ((sw + ww) >> 1) &&
			    fabs(sc[0] - wc[0]) < 25 && fabs(sc[1] - wc[1]) < 25 && fabs(sc[2] - wc[2]) < 25 &&
			    abs(s[5] - w[5]) < (s[5] < w[5] ? s[5] : w[5]) * 3) {
				taken[n] = 1;
				tracked_out[2 * len] = 1; tracked_out[2 * len + 1] = n; len++;
			}
		}
	}
	free(taken);
	return len;
}

/* OCR::rotate_mat(src, dst, rad, crop) (src/OCR.cpp:254-360) on a contiguous image; returns a malloc'ed image */
static uint8_t *rotate_image(const uint8_t *src, int cols, int rows, double rad, int crop, int *out_w, int *out_h)
{
	const int x0 = (int)((cols - 1) / 2.0), y0 = (int)((rows - 1) / 2.0);
	const int cx[4] = { 0 - x0, (cols - 1) - x0, (cols - 1) - x0, 0 - x0 };
	const int cy[4] = { 0 - y0, 0 - y0, (rows - 1) - y0, (rows - 1) - y0 };
	int nx[4], ny[4];
	for (int k = 0; k < 4; k++) {
		nx[k] = (int)round(cx[k] * cos(rad) - cy[k] * sin(rad));
		ny[k] = (int)round(cx[k] * sin(rad) + cy[k] * cos(rad));
	}
	int max_x = nx[0], max_y = ny[0], min_x = nx[0], min_y = ny[0];
	for (int k = 1; k < 4; k++) {
		if (nx[k] > max_x) max_x = nx[k]; if (nx[k] < min_x) min_x = nx[k];
		if (ny[k] > max_y) max_y = ny[k]; if (ny[k] < min_y) min_y = ny[k];
	}
	int crop_h = 0;
	if (crop) {
		crop_h = (int)((nx[1] - nx[0]) * tan(rad) * 0.5);
		if (max_y - min_y + 1 - 2 * crop_h <= 0) { crop = 0; crop_h = 0; }          /* falls back to the un-cropped form */
	}
	const int ow = max_x - min_x + 1, oh = max_y - min_y + 1 - 2 * crop_h;
	uint8_t *out = (uint8_t *)calloc((size_t)(ow > 0 && oh > 0 ? ow * oh : 1), 1);
	*out_w = ow; *out_h = oh;
	if (ow <= 0 || oh <= 0) return out;
	for (int i = min_y + crop_h; i < max_y - crop_h; i++) {
		uint8_t *t = out + (size_t)(i - min_y - crop_h) * ow;
		for (int j = min_x; j < max_x; j++) {
			const double ii = (double)(i - crop_h);                                    /* crop_h is 0 in the un-cropped form */
			const double new_j = cos(rad) * j - sin(rad) * ii + x0;
			const double new_i = sin(rad) * j + cos(rad) * ii + y0;
			if (!(new_i > 0 && new_j > 0 && new_i < rows - 1 && new_j < cols - 1)) continue;
			if (crop && !(i > min_y + crop_h)) continue;                               /* first row stays empty in the cropped form */
			const uint8_t *s = src + (size_t)(int)new_i * cols + (int)new_j;
			if (new_i == floor(new_i) && new_j == floor(new_j)) t[j - min_x] = s[0];
			else {
				const double alpha = new_i - floor(new_i), beta = new_j - floor(new_j);
				const uint8_t A = s[0], B = s[1], C = s[cols], D = s[cols + 1];
				t[j - min_x] = (uint8_t)round((1 - alpha) * (1 - beta) * A + (1 - alpha) * beta * B + alpha * (1 - beta) * C + alpha * beta * D);
			}
		}
	}
	return out;
}

/* the eight chain-code bitmaps of OCR::extract_feature (src/OCR.cpp:153-169): every step (pixel -> next pixel) of every
 * border that cv::findContours(RETR_LIST, CHAIN_APPROX_NONE) traces sets the pixel in the bitmap of its direction
 * (chain_code_direction(next, current), src/OCR.cpp:602-622 = (4 - s) & 7 for OpenCV's step code s).  Border following is
 * Suzuki-Abe as OpenCV runs it, on a zero-framed copy: 1 untouched, 2 visited, 2|-128 visited with the border to its east. */
static void chain_code_bitmaps(const uint8_t *img, int L, uint8_t *f /* 8 x L x L */)
{
	const int W = L + 2;
	signed char *b = (signed char *)calloc((size_t)W * W, 1);
	for (int y = 0; y < L; y++) for (int x = 0; x < L; x++) b[(y + 1) * W + x + 1] = img[y * L + x] ? 1 : 0;
	const int dx[8] = { 1, 1, 0, -1, -1, -1, 0, 1 }, dy[8] = { 0, -1, -1, -1, 0, 1, 1, 1 };
	int delta[16];
	for (int k = 0; k < 16; k++) delta[k] = dy[k & 7] * W + dx[k & 7];
	memset(f, 0, (size_t)8 * L * L);
	for (int y = 1; y <= L; y++) {
		int prev = 0;
		for (int x = 1; x <= L + 1; x++) {
			int p = b[y * W + x];
			if (p == prev) continue;
			int hole = 0, start = 0;
			if (prev == 0 && p == 1) start = 1;
			else if (p == 0 && prev >= 1) { start = 1; hole = 1; }
			if (start) {
				const int i0 = y * W + x - hole;
				int s_end = hole ? 0 : 4, s = s_end, i1;
				do { s = (s - 1) & 7; i1 = i0 + delta[s]; } while (b[i1] == 0 && s != s_end);
				if (s == s_end) b[i0] = (signed char)(2 | -128);                  /* isolated pixel: a 1-point contour, skipped (src/OCR.cpp:159) */
				else {
					int i3 = i0, i4 = i0;
					for (;;) {
						s_end = s;
						while (s < 15) { i4 = i3 + delta[++s]; if (b[i4] != 0) break; }
						s &= 7;
						if ((unsigned)(s - 1) < (unsigned)s_end) b[i3] = (signed char)(2 | -128);
						else if (b[i3] == 1) b[i3] = 2;
						f[((4 - s) & 7) * L * L + (i3 / W - 1) * L + (i3 % W - 1)] = 255;
						if (i4 == i0 && i3 == i1) break;
						i3 = i4;
						s = (s + 4) & 7;
					}
				}
				p = b[y * W + x];
			}
			prev = p;
		}
	}
	free(b);
}

/* chain_run's pre-processing (src/OCR.cpp:72-79) + extract_feature (src/OCR.cpp:144-218): returns -1 where the reference's
 * cv::resize would throw (empty target).  img30: the 30x30 image extract_feature receives; feat1800: value * 255. */
PORT_API int port_ocr_features(const uint8_t *crop, int w, int h, int stride, double slope, uint8_t *img30, uint8_t *feat1800)
{
	const int L = 30, FL = 15;
	const int thr = otsu_of_inverted(crop, w, h, stride);
	uint8_t *bin = (uint8_t *)malloc((size_t)w * h);
	for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) bin[(size_t)y * w + x] = (255 - crop[(size_t)y * stride + x] > thr) ? 255 : 0;
	int iw = w, ih = h;
	if (fabs(slope) > 0.01) {
		uint8_t *rot = rotate_image(bin, w, h, atan2(slope, 1), 1, &iw, &ih);
		free(bin);
		bin = rot;
	}
	if (iw < 1 || ih < 1 || port_aran_minor(iw, ih, L) < 1) { free(bin); return -1; }
	port_aran(bin, iw, ih, iw, L, img30);
	free(bin);
	uint8_t f[8 * 30 * 30];
	chain_code_bitmaps(img30, L, f);
	static const int K[7] = { 8, 28, 56, 72, 56, 28, 8 };      /* GaussianBlur(7x7, sigma 0) on 8U: fixed point, reflect-101 */
	int nnz = 0;
	for (int c = 0; c < 8; c++) {
		const uint8_t *s = f + c * L * L;
		int hp[30 * 30];
		uint8_t g[30 * 30];
		for (int y = 0; y < L; y++) for (int x = 0; x < L; x++) {
			int a = 0;
			for (int k = 0; k < 7; k++) { int xi = x + k - 3; if (xi < 0) xi = -xi; if (xi >= L) xi = 2 * L - 2 - xi; a += K[k] * s[y * L + xi]; }
			hp[y * L + x] = a;
		}
		int mn = 255, mx = 0;
		for (int y = 0; y < L; y++) for (int x = 0; x < L; x++) {
			int a = 0;
			for (int k = 0; k < 7; k++) { int yi = y + k - 3; if (yi < 0) yi = -yi; if (yi >= L) yi = 2 * L - 2 - yi; a += K[k] * hp[yi * L + x]; }
			const int v = (a + 32768) >> 16;
			g[y * L + x] = (uint8_t)v;
			if (v < mn) mn = v; if (v > mx) mx = v;
		}
		/* normalize(0, 255, NORM_MINMAX, CV_8U): double scale / shift, cast to float, ONE fused multiply-add, round half even */
		const double scale = 255.0 * ((double)(mx - mn) > DBL_EPSILON ? 1. / (double)(mx - mn) : 0);
		const double shift = 0.0 - mn * scale;
		const float fa = (float)scale, fb = (float)shift;
		for (int i = 0; i < L * L; i++) { long v = lrintf(fmaf((float)g[i], fa, fb)); g[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
		for (int y = 0; y < FL; y++) for (int x = 0; x < FL; x++) {            /* resize 30 -> 15: the exact-2x box */
			const uint8_t *a = g + (2 * y) * L + 2 * x;
			const uint8_t v = (uint8_t)((a[0] + a[1] + a[L] + a[L + 1] + 2) >> 2);
			feat1800[c * FL * FL + y * FL + x] = v;
			nnz += v != 0;
		}
	}
	return nnz;
}
