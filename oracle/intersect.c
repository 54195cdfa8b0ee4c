/*
 * oracle/intersect.c -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement of the closest-hit query the reference delegates to Mitsuba/OptiX
 * (reference utils/path_tracing.py:17-48: scene.ray_intersect_preliminary +
 * compute_surface_interaction + double_sided).  Mitsuba 3.5.0 / Dr.Jit 0.4.4 are
 * un-vendored third-party dependencies (environment.yml:14-15) and absent here, so the
 * semantics below are DEFINED by this file -- parity
 * with OptiX itself is unpinned:
 *
 *   - exact closest hit over all triangles of one mesh, prim index = face order;
 *   - Moller-Trumbore barycentrics in fp32, one rounding per operation (compile with
 *     -ffp-contract=off), fixed operation order shared bit-for-bit with the CUDA kernel
 *     (iris_b200/csrc/traverse.cuh:tri_test);
 *   - det = e1.(d x e2) must clear a noise floor: |det| > 1e-6 * (|e1x px|+|e1y py|+|e1z pz|)
 *     (rays within ~1e-6 rad of the triangle's plane count as parallel); triangles whose
 *     e1 x e2 is exactly zero never hit;
 *   - hit iff 0<=u<=1, v>=0, u+v<=1, 0 < t < inf where p = fma(v,e2,fma(u,e1,v0)) is the
 *     barycentric point and t = ((p-o).d) * (1/(d.d)) its projection on the ray -- unlike
 *     Moller-Trumbore's own t this stays well conditioned at grazing angles, so the hit always
 *     lies inside the triangle's bounding box and BVH culling agrees with the brute-force loop;
 *   - equal t -> lowest prim index, and the ray is flagged in `tie`;
 *   - n = normalize(e1 x e2) flipped so that dot(n,-d) >= 0 (utils/ops.py:85-96), uv = (u,v);
 *   - miss -> t=inf, prim=-1, p=n=uv=0.
 *
 * Two drivers share the triangle test: a brute-force loop (ground truth for small
 * scenes) and a median-split BVH2 (so 1M-triangle scenes finish in seconds; it is
 * validated against the brute-force loop in tests/test_oracle_cpu.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float v0[3], e1[3], e2[3]; } Tri;

static inline float dot3(const float a[3], const float b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static inline void cross3(const float a[3], const float b[3], float r[3])
{
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}

/* returns 1 on hit and writes t,u,v; rdd = 1/(d.d) */
static inline int tri_test(const float o[3], const float d[3], float rdd, const Tri *tr, float *t, float *u, float *v)
{
    float p[3], s[3], q[3];
    cross3(d, tr->e2, p);
    float det = dot3(tr->e1, p);
    float mag = (fabsf(tr->e1[0] * p[0]) + fabsf(tr->e1[1] * p[1])) + fabsf(tr->e1[2] * p[2]);
    if (!(fabsf(det) > 1e-6f * mag)) return 0;
    float inv = 1.0f / det;
    s[0] = o[0] - tr->v0[0]; s[1] = o[1] - tr->v0[1]; s[2] = o[2] - tr->v0[2];
    float uu = dot3(s, p) * inv;
    if (!(uu >= 0.0f && uu <= 1.0f)) return 0;
    cross3(s, tr->e1, q);
    float vv = dot3(d, q) * inv;
    if (!(vv >= 0.0f && uu + vv <= 1.0f)) return 0;
    float h[3];
    for (int k = 0; k < 3; ++k) h[k] = fmaf(vv, tr->e2[k], fmaf(uu, tr->e1[k], tr->v0[k])) - o[k];
    float tt = dot3(h, d) * rdd;
    if (!(tt > 0.0f && tt < INFINITY)) return 0;
    *t = tt; *u = uu; *v = vv;
    return 1;
}

/* ------------------------------------------------------------------------- BVH2 */
typedef struct { float lo[3], hi[3]; int32_t left, count; } Node; /* leaf: count>0,left=first; inner: count=0,left=child (right=left+1) */

typedef struct {
    int64_t nf;
    Tri *tris;        /* in face order */
    int32_t *order;   /* BVH leaf order -> face index */
    Node *nodes;
    int64_t n_nodes;
    float pad;
} OracleScene;

static void tri_bounds(const Tri *t, float lo[3], float hi[3])
{
    for (int k = 0; k < 3; ++k) {
        float a = t->v0[k], b = t->v0[k] + t->e1[k], c = t->v0[k] + t->e2[k];
        lo[k] = fminf(a, fminf(b, c));
        hi[k] = fmaxf(a, fmaxf(b, c));
    }
}

static void build_rec(OracleScene *S, float *cent, int32_t node, int32_t first, int32_t count)
{
    Node *n = &S->nodes[node];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int32_t i = first; i < first + count; ++i) {
        float a[3], b[3];
        tri_bounds(&S->tris[S->order[i]], a, b);
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], a[k]); hi[k] = fmaxf(hi[k], b[k]);
            float c = cent[3 * (int64_t)S->order[i] + k];
            clo[k] = fminf(clo[k], c); chi[k] = fmaxf(chi[k], c);
        }
    }
    for (int k = 0; k < 3; ++k) { n->lo[k] = lo[k] - S->pad; n->hi[k] = hi[k] + S->pad; }
    if (count <= 4) { n->left = first; n->count = count; return; }
    int ax = 0;
    if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
    if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
    float mid = 0.5f * (clo[ax] + chi[ax]);
    int32_t i = first, j = first + count - 1;
    while (i <= j) {
        if (cent[3 * (int64_t)S->order[i] + ax] < mid) ++i;
        else { int32_t tmp = S->order[i]; S->order[i] = S->order[j]; S->order[j] = tmp; --j; }
    }
    int32_t nl = i - first;
    if (nl == 0 || nl == count) nl = count / 2; /* coincident centroids: split by position in the list */
    int32_t child = (int32_t)S->n_nodes;
    S->n_nodes += 2;
    n->left = child; n->count = 0;
    build_rec(S, cent, child, first, nl);
    build_rec(S, cent, child + 1, first + nl, count - nl);
}

OracleScene *oracle_scene_create(const float *verts, int64_t nv, const int32_t *faces, int64_t nf)
{
    (void)nv;
    OracleScene *S = (OracleScene *)calloc(1, sizeof(OracleScene));
    S->nf = nf;
    S->tris = (Tri *)malloc(sizeof(Tri) * (size_t)(nf > 0 ? nf : 1));
    S->order = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nf > 0 ? nf : 1));
    S->nodes = (Node *)malloc(sizeof(Node) * (size_t)(2 * nf + 2));
    float *cent = (float *)malloc(sizeof(float) * 3 * (size_t)(nf > 0 ? nf : 1));
    float glo = INFINITY, ghi = -INFINITY;
    for (int64_t f = 0; f < nf; ++f) {
        const float *a = verts + 3 * (int64_t)faces[3 * f], *b = verts + 3 * (int64_t)faces[3 * f + 1], *c = verts + 3 * (int64_t)faces[3 * f + 2];
        for (int k = 0; k < 3; ++k) {
            S->tris[f].v0[k] = a[k];
            S->tris[f].e1[k] = b[k] - a[k];
            S->tris[f].e2[k] = c[k] - a[k];
            cent[3 * f + k] = (a[k] + b[k] + c[k]) * (1.0f / 3.0f);
            glo = fminf(glo, fminf(a[k], fminf(b[k], c[k])));
            ghi = fmaxf(ghi, fmaxf(a[k], fmaxf(b[k], c[k])));
        }
        S->order[f] = (int32_t)f;
        float cr[3];
        cross3(S->tris[f].e1, S->tris[f].e2, cr);
        if (cr[0] == 0.0f && cr[1] == 0.0f && cr[2] == 0.0f)      /* zero-area triangle: never hit */
            for (int k = 0; k < 3; ++k) S->tris[f].e1[k] = S->tris[f].e2[k] = 0.0f;
    }
    S->pad = nf > 0 ? 1e-5f * (ghi - glo) + 1e-30f : 0.0f;
    S->n_nodes = 1;
    if (nf > 0) build_rec(S, cent, 0, 0, (int32_t)nf);
    free(cent);
    return S;
}

void oracle_scene_destroy(OracleScene *S)
{
    if (!S) return;
    free(S->tris); free(S->order); free(S->nodes); free(S);
}

static inline void consider(const OracleScene *S, int32_t f, const float o[3], const float d[3], float rdd,
                            float *bt, int32_t *bp, float *bu, float *bv, uint8_t *tie)
{
    float t, u, v;
    if (!tri_test(o, d, rdd, &S->tris[f], &t, &u, &v)) return;
    if (t < *bt) { *bt = t; *bp = f; *bu = u; *bv = v; *tie = 0; }
    else if (t == *bt) { *tie = 1; if (f < *bp) { *bp = f; *bu = u; *bv = v; } }
}

static void finish(const OracleScene *S, const float o[3], const float d[3], float bt, int32_t bp, float bu, float bv,
                   float *t, int32_t *prim, float *uv, float *p, float *nrm)
{
    (void)o;
    if (bp < 0) {
        *t = INFINITY; *prim = -1; uv[0] = uv[1] = 0.0f;
        p[0] = p[1] = p[2] = 0.0f; nrm[0] = nrm[1] = nrm[2] = 0.0f;
        return;
    }
    const Tri *tr = &S->tris[bp];
    *t = bt; *prim = bp; uv[0] = bu; uv[1] = bv;
    for (int k = 0; k < 3; ++k) p[k] = fmaf(bv, tr->e2[k], fmaf(bu, tr->e1[k], tr->v0[k]));
    float c[3];
    cross3(tr->e1, tr->e2, c);
    float len = sqrtf(dot3(c, c));
    float n[3] = {c[0] / len, c[1] / len, c[2] / len};
    float nd = dot3(n, d);            /* dot(n,-d) < 0  <=>  dot(n,d) > 0 */
    if (nd > 0.0f) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
    nrm[0] = n[0]; nrm[1] = n[1]; nrm[2] = n[2];
}

/* mode 0 = brute force, 1 = BVH2 */
void oracle_intersect(const OracleScene *S, int mode, const float *os, const float *ds, int64_t n,
                      float *t, int32_t *prim, float *uv, float *p, float *nrm, uint8_t *tie)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < n; ++r) {
        const float *o = os + 3 * r, *d = ds + 3 * r;
        float bt = INFINITY, bu = 0, bv = 0;
        int32_t bp = -1;
        uint8_t tflag = 0;
        const float rdd = 1.0f / dot3(d, d);
        if (mode == 0 || S->nf == 0) {
            for (int64_t f = 0; f < S->nf; ++f) consider(S, (int32_t)f, o, d, rdd, &bt, &bp, &bu, &bv, &tflag);
        } else {
            float id[3];
            for (int k = 0; k < 3; ++k) {
                float dk = d[k];
                if (fabsf(dk) < 1e-30f) dk = copysignf(1e-30f, dk);
                id[k] = 1.0f / dk;
            }
            int32_t stack[256];
            int sp = 0;
            stack[sp++] = 0;
            while (sp) {
                const Node *nd = &S->nodes[stack[--sp]];
                float tn = 0.0f, tf = bt;
                for (int k = 0; k < 3; ++k) {
                    float a = (nd->lo[k] - o[k]) * id[k], b = (nd->hi[k] - o[k]) * id[k];
                    float t0 = fminf(a, b), t1 = fmaxf(a, b);
                    t0 = t0 - fabsf(t0) * 1e-6f; t1 = t1 + fabsf(t1) * 1e-6f;   /* conservative slabs */
                    if (t0 > tn) tn = t0;
                    if (t1 < tf) tf = t1;
                }
                if (tn > tf) continue;
                if (nd->count > 0) {
                    for (int32_t i = nd->left; i < nd->left + nd->count; ++i)
                        consider(S, S->order[i], o, d, rdd, &bt, &bp, &bu, &bv, &tflag);
                } else {
                    if (sp + 2 > 256) abort();
                    stack[sp++] = nd->left;
                    stack[sp++] = nd->left + 1;
                }
            }
        }
        finish(S, o, d, bt, bp, bu, bv, t + r, prim + r, uv + 2 * r, p + 3 * r, nrm + 3 * r);
        if (tie) tie[r] = tflag;
    }
}
