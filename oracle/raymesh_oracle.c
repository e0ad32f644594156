/*
 * raymesh_oracle.c — CPU ORACLE for the ray/mesh intersection hot path.  TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; nothing under trimesh-ray-optix_b200/ does.  It is the checker, never the
 * product.
 *
 * PARITY STATUS: "parity unpinned" beyond the reference's hand-derivable known answers.
 * The arithmetic of the reference lives in the closed-source NVIDIA OptiX runtime (SDK >= 7.7,
 * unpinned; README.md:8 of the reference), reached through optixAccelBuild
 * (triro/backend/ray.cpp:79) and optixTrace (triro/backend/shaders.cu:86,112,163,191,238).
 * It cannot be built or run here, and the reference's own tests (test/test.py) contain no
 * assertions.  This oracle therefore restates the SEMANTICS the reference's programs define
 * around those calls, and is pinned against the known answers K1-K7 of test/test.py, against the
 * reference's published README figure assets/location.png and against a fixture produced by executing
 * the reference's own Python host logic (see tests/test_oracle_known_answers.py, tests/golden/):
 *   - ray interval: tmin = 0, tmax = 1e7, open on both sides        shaders.cu:86,112,163,191,238
 *   - closest hit: min t; tri index; front = CCW seen from origin;
 *     loc = u*v1 + v*v2 + (1-u-v)*v0; uv = (1-u-v, u)                shaders.cu:137-153
 *   - miss: hit 0, front 0, tri -1, loc 0, uv 0                      shaders.cu:128-135
 *   - count: every triangle hit in the interval, once                shaders.cu:176-181, ray.cpp:60-62
 *   - all hits: (tri, loc) per hit                                   shaders.cu:207-224
 * Two evaluators share one driver:
 *   TRUTH  (mode 0): binary64, scalar-triple-product edge functions, plus a per-ray "grazing"
 *                    classifier saying when a binary32 implementation may legitimately differ
 *                    (the documented tie set of BASELINE.json's north_star).
 *   MIRROR (mode 1): binary32, the exact IEEE operation sequence of the product's watertight
 *                    test (Woop/Benthin/Wald 2013) — results must be bit-identical to the GPU
 *                    for every ray; it proves that BVH build/quantisation/traversal never drop
 *                    a candidate.  Compile with -ffp-contract=off.
 * Candidates are either all triangles (brute force) or the triangles of a padded binned-SAH
 * BVH2 (same answers, used for the large configurations and as the CPU baseline).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_MAX_LIST 64

/* grazing flags (truth mode) */
#define GZ_EDGE_CLOSEST 1   /* a triangle whose hit/miss decision is uncertain could be (or beat) the closest hit */
#define GZ_EDGE_ANY 2       /* some triangle in the interval has an uncertain hit/miss decision (count / all-hits) */
#define GZ_TIE 4            /* two different triangles within 1e-5 relative of the nearest t */
#define GZ_TNEAR 8          /* a candidate hit has t within tolerance of 0 or tmax */
#define GZ_EDGEON 16        /* closest triangle nearly parallel to the ray (front flag / t unstable) */

typedef struct {
    const float* verts;
    const int32_t* faces;
    int64_t nv, nf;
} Mesh;

typedef struct {
    double t;
    int32_t tri;
} HitRec;

typedef struct {
    int mode;
    double tmax;
    /* ray */
    double o[3], d[3], dlen;
    float of[3], df[3];
    /* mirror set-up (product: rt_core.cuh ray_setup) */
    int kz;
    float Sx, Sy, Sz, okx, oky, okz;
    /* accumulators */
    double best_t, second_t;
    int32_t best_tri;
    int32_t count;
    uint32_t flags;
    int nlist, list_cap;
    HitRec* list;
    double uncertain_tmin; /* smallest t of an uncertain triangle */
} RayState;

static inline float sel3f(int k, float x, float y, float z) { return k == 0 ? x : (k == 1 ? y : z); }

static void ray_init(RayState* rs, int mode, const float o[3], const float d[3], double tmax, HitRec* list, int cap) {
    rs->mode = mode;
    rs->tmax = tmax;
    for (int a = 0; a < 3; ++a) { rs->o[a] = o[a]; rs->d[a] = d[a]; rs->of[a] = o[a]; rs->df[a] = d[a]; }
    rs->dlen = sqrt(rs->d[0] * rs->d[0] + rs->d[1] * rs->d[1] + rs->d[2] * rs->d[2]);
    int kz = 0;
    float m = fabsf(d[0]);
    if (fabsf(d[1]) > m) { kz = 1; m = fabsf(d[1]); }
    if (fabsf(d[2]) > m) { kz = 2; }
    const int kx = kz == 2 ? 0 : kz + 1, ky = kx == 2 ? 0 : kx + 1;
    rs->kz = kz;
    const float dkz = sel3f(kz, d[0], d[1], d[2]);
    rs->Sx = sel3f(kx, d[0], d[1], d[2]) / dkz;
    rs->Sy = sel3f(ky, d[0], d[1], d[2]) / dkz;
    rs->Sz = 1.0f / dkz;
    rs->okx = sel3f(kx, o[0], o[1], o[2]);
    rs->oky = sel3f(ky, o[0], o[1], o[2]);
    rs->okz = sel3f(kz, o[0], o[1], o[2]);
    rs->best_t = tmax; rs->second_t = INFINITY; rs->best_tri = -1; rs->count = 0; rs->flags = 0;
    rs->nlist = 0; rs->list_cap = cap; rs->list = list; rs->uncertain_tmin = INFINITY;
}

static void list_insert(RayState* rs, double t, int32_t tri) {
    /* keep the list_cap hits with the smallest (t, tri) */
    if (rs->list_cap <= 0) return;
    int n = rs->nlist;
    if (n == rs->list_cap) {
        const HitRec* w = &rs->list[n - 1];
        if (t > w->t || (t == w->t && tri > w->tri)) return;
        --n;
    }
    int i = n;
    while (i > 0 && (rs->list[i - 1].t > t || (rs->list[i - 1].t == t && rs->list[i - 1].tri > tri))) {
        rs->list[i] = rs->list[i - 1];
        --i;
    }
    rs->list[i].t = t; rs->list[i].tri = tri;
    rs->nlist = n + 1;
}

typedef struct { double t, bu, bv, det_sign_front; int hit; } TriEval;

/* ---------------------------------------------------------------- TRUTH (binary64) */
static void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double norm3(const double a[3]) { return sqrt(dot3(a, a)); }
static double perp_len(const double a[3], const double d[3], double dlen) {
    /* length of the component of a perpendicular to d */
    double c[3]; cross3(a, d, c);
    return dlen > 0 ? norm3(c) / dlen : norm3(a);
}

/* returns 1 on hit; fills t, barycentrics (OptiX u,v), front; sets *uncertain when a binary32
 * evaluation could decide differently. */
static int tri_truth(const RayState* rs, const float* v0, const float* v1, const float* v2, double* t, double* bu,
                     double* bv, int* front, int* uncertain, int* edgeon) {
    double A[3], B[3], C[3];
    for (int a = 0; a < 3; ++a) { A[a] = (double)v0[a] - rs->o[a]; B[a] = (double)v1[a] - rs->o[a]; C[a] = (double)v2[a] - rs->o[a]; }
    double CxB[3], AxC[3], BxA[3];
    cross3(C, B, CxB); cross3(A, C, AxC); cross3(B, A, BxA);
    /* edge functions: signed volumes; all of one sign <=> the ray passes inside the triangle */
    const double U = dot3(rs->d, CxB), V = dot3(rs->d, AxC), W = dot3(rs->d, BxA);
    const double det = U + V + W;
    *uncertain = 0; *edgeon = 0; *front = 0; *t = 0; *bu = 0; *bv = 0;
    /* binary32 error model: coordinates of A,B,C carry ~eps32*|.| after shearing, so an edge
     * function carries ~eps32*(|B||C_perp| + |C||B_perp|)*|d| */
    const double eps = 16.0 * 5.9604644775390625e-08;
    const double la = norm3(A), lb = norm3(B), lc = norm3(C);
    const double pa = perp_len(A, rs->d, rs->dlen), pb = perp_len(B, rs->d, rs->dlen), pc = perp_len(C, rs->d, rs->dlen);
    const double tolU = eps * rs->dlen * (lb * pc + lc * pb);
    const double tolV = eps * rs->dlen * (la * pc + lc * pa);
    const double tolW = eps * rs->dlen * (la * pb + lb * pa);
    const int neg = (U < 0) || (V < 0) || (W < 0), pos = (U > 0) || (V > 0) || (W > 0);
    const int inside = !(neg && pos) && det != 0.0;
    /* could a perturbation within tolerance make it inside / outside? */
    const int near_edge = fabs(U) <= tolU || fabs(V) <= tolV || fabs(W) <= tolW;
    int maybe_inside = 0;
    if (near_edge) {
        /* inside-or-nearly for one orientation */
        const int okp = (U >= -tolU) && (V >= -tolV) && (W >= -tolW);
        const int okn = (U <= tolU) && (V <= tolV) && (W <= tolW);
        maybe_inside = okp || okn;
    }
    double tt = 0.0;
    if (det != 0.0 && rs->dlen > 0) {
        double P[3];
        for (int a = 0; a < 3; ++a) P[a] = (U * A[a] + V * B[a] + W * C[a]) / det;
        tt = dot3(P, rs->d) / (rs->dlen * rs->dlen);
    }
    /* edge-on: |det| small against its own error bound */
    const double tolDet = tolU + tolV + tolW;
    if (fabs(det) <= 4.0 * tolDet) *edgeon = 1;
    if (maybe_inside || (inside && *edgeon)) *uncertain = 1;
    *t = tt;
    if (!inside) return 0;
    *bu = V / det; *bv = W / det;
    /* front: dot(d, (v1-v0) x (v2-v0)) < 0 */
    double e1[3], e2[3], n[3];
    for (int a = 0; a < 3; ++a) { e1[a] = (double)v1[a] - (double)v0[a]; e2[a] = (double)v2[a] - (double)v0[a]; }
    cross3(e1, e2, n);
    *front = dot3(rs->d, n) < 0.0;
    return 1;
}

/* ---------------------------------------------------------------- MIRROR (binary32, product arithmetic) */
typedef struct { float t, U, V, W, det; } TriHitF;

static int tri_mirror(const RayState* rs, const float* v0, const float* v1, const float* v2, TriHitF* h) {
    const int kz = rs->kz, kx = kz == 2 ? 0 : kz + 1, ky = kx == 2 ? 0 : kx + 1;
    const float Akx = v0[kx] - rs->okx, Aky = v0[ky] - rs->oky, Akz = v0[kz] - rs->okz;
    const float Bkx = v1[kx] - rs->okx, Bky = v1[ky] - rs->oky, Bkz = v1[kz] - rs->okz;
    const float Ckx = v2[kx] - rs->okx, Cky = v2[ky] - rs->oky, Ckz = v2[kz] - rs->okz;
    const float Ax = fmaf(-rs->Sx, Akz, Akx), Ay = fmaf(-rs->Sy, Akz, Aky);
    const float Bx = fmaf(-rs->Sx, Bkz, Bkx), By = fmaf(-rs->Sy, Bkz, Bky);
    const float Cx = fmaf(-rs->Sx, Ckz, Ckx), Cy = fmaf(-rs->Sy, Ckz, Cky);
    float p, q;
    p = Cx * By; q = Cy * Bx; float U = p - q;
    p = Ax * Cy; q = Ay * Cx; float V = p - q;
    p = Bx * Ay; q = By * Ax; float W = p - q;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        double pd, qd;
        pd = (double)Cx * (double)By; qd = (double)Cy * (double)Bx; const double Ud = pd - qd;
        pd = (double)Ax * (double)Cy; qd = (double)Ay * (double)Cx; const double Vd = pd - qd;
        pd = (double)Bx * (double)Ay; qd = (double)By * (double)Ax; const double Wd = pd - qd;
        if ((Ud < 0.0 || Vd < 0.0 || Wd < 0.0) && (Ud > 0.0 || Vd > 0.0 || Wd > 0.0)) return 0;
        U = (float)Ud; V = (float)Vd; W = (float)Wd;
    } else {
        if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return 0;
    }
    float det = U + V;
    det = det + W;
    if (!(det != 0.0f)) return 0;
    const float Az = rs->Sz * Akz, Bz = rs->Sz * Bkz, Cz = rs->Sz * Ckz;
    float T = W * Cz;
    T = fmaf(V, Bz, T);
    T = fmaf(U, Az, T);
    h->t = T / det; h->U = U; h->V = V; h->W = W; h->det = det;
    return 1;
}

static void attr_mirror(const TriHitF* h, const float* v0, const float* v1, const float* v2, float loc[3], float uv[2]) {
    const float bu = h->V / h->det, bv = h->W / h->det;
    float w0 = 1.0f - bu;
    w0 = w0 - bv;
    for (int a = 0; a < 3; ++a) {
        float x = w0 * v0[a];
        x = fmaf(bv, v2[a], x);
        x = fmaf(bu, v1[a], x);
        loc[a] = x;
    }
    uv[0] = w0; uv[1] = bu;
}

static int front_mirror(const RayState* rs, const TriHitF* h) {
    const float dkz = sel3f(rs->kz, rs->df[0], rs->df[1], rs->df[2]);
    return (h->det > 0.0f) == (dkz > 0.0f);
}

/* ---------------------------------------------------------------- per (ray, triangle) visit */
static inline const float* vert(const Mesh* m, int64_t tri, int c) {
    int32_t i = m->faces[3 * tri + c];
    if (i < 0) i = 0;
    if (i >= m->nv) i = (int32_t)(m->nv - 1);
    return m->verts + 3 * (size_t)i;
}

static void visit_tri(RayState* rs, const Mesh* m, int32_t tri) {
    const float *v0 = vert(m, tri, 0), *v1 = vert(m, tri, 1), *v2 = vert(m, tri, 2);
    double t;
    int hit;
    if (rs->mode == 0) {
        double bu, bv; int front, uncertain, edgeon;
        hit = tri_truth(rs, v0, v1, v2, &t, &bu, &bv, &front, &uncertain, &edgeon);
        const double ttol = 1e-5 * fabs(t) + 1e-30;
        const int in_interval_loose = t > -ttol && t < rs->tmax * (1.0 + 1e-5);
        if (uncertain && in_interval_loose) {
            rs->flags |= GZ_EDGE_ANY;
            if (t < rs->uncertain_tmin) rs->uncertain_tmin = t;
        }
        if (hit && (fabs(t) <= 1e-6 * (fabs(rs->o[0]) + fabs(rs->o[1]) + fabs(rs->o[2]) + 1e-30) / (rs->dlen > 0 ? rs->dlen : 1) ||
                    fabs(t - rs->tmax) <= 1e-5 * rs->tmax))
            rs->flags |= GZ_TNEAR;
        if (!(hit && t > 0.0 && t < rs->tmax)) return;
    } else {
        TriHitF h;
        hit = tri_mirror(rs, v0, v1, v2, &h);
        if (!hit) return;
        if (!(h.t > 0.0f && h.t < (float)rs->tmax)) return;
        t = (double)h.t;
    }
    rs->count += 1;
    list_insert(rs, t, tri);
    if (t < rs->best_t || (t == rs->best_t && tri < rs->best_tri)) {
        if (rs->best_tri >= 0) rs->second_t = rs->best_t;
        rs->best_t = t; rs->best_tri = tri;
    } else if (t < rs->second_t) {
        rs->second_t = t;
    }
}

/* ---------------------------------------------------------------- binned-SAH BVH2 (candidate culling only) */
typedef struct {
    double lo[3], hi[3];
    int32_t left, right;   /* children, or -1 */
    int32_t first, count;  /* leaf range in `order` */
} BNode;

typedef struct {
    BNode* nodes;
    int32_t n_nodes;
    int32_t* order;
    Mesh mesh;
} OracleBvh;

typedef struct { double lo[3], hi[3], c[3]; } TBox;

static void tbox_of(const Mesh* m, int32_t tri, TBox* b) {
    for (int a = 0; a < 3; ++a) { b->lo[a] = INFINITY; b->hi[a] = -INFINITY; }
    for (int c = 0; c < 3; ++c) {
        const float* v = vert(m, tri, c);
        for (int a = 0; a < 3; ++a) {
            if (v[a] < b->lo[a]) b->lo[a] = v[a];
            if (v[a] > b->hi[a]) b->hi[a] = v[a];
        }
    }
    for (int a = 0; a < 3; ++a) b->c[a] = 0.5 * (b->lo[a] + b->hi[a]);
}

static double half_area(const double lo[3], const double hi[3]) {
    const double x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
    if (!(x >= 0 && y >= 0 && z >= 0)) return 0.0;
    return x * y + y * z + z * x;
}

#define NBINS 16
#define LEAF_TRIS 4

static int32_t build_rec(OracleBvh* bvh, const TBox* tb, int32_t first, int32_t count, int depth) {
    const int32_t id = bvh->n_nodes++;
    BNode* nd = &bvh->nodes[id];
    double clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int a = 0; a < 3; ++a) { nd->lo[a] = INFINITY; nd->hi[a] = -INFINITY; }
    for (int32_t i = first; i < first + count; ++i) {
        const TBox* b = &tb[bvh->order[i]];
        for (int a = 0; a < 3; ++a) {
            if (b->lo[a] < nd->lo[a]) nd->lo[a] = b->lo[a];
            if (b->hi[a] > nd->hi[a]) nd->hi[a] = b->hi[a];
            if (b->c[a] < clo[a]) clo[a] = b->c[a];
            if (b->c[a] > chi[a]) chi[a] = b->c[a];
        }
    }
    nd->left = nd->right = -1; nd->first = first; nd->count = count;
    if (count <= LEAF_TRIS || depth > 60) return id;
    /* binned SAH over the three axes */
    int best_axis = -1, best_split = -1;
    double best_cost = INFINITY;
    for (int a = 0; a < 3; ++a) {
        const double ext = chi[a] - clo[a];
        if (!(ext > 0)) continue;
        int cnt[NBINS]; double blo[NBINS][3], bhi[NBINS][3];
        for (int b = 0; b < NBINS; ++b) { cnt[b] = 0; for (int k = 0; k < 3; ++k) { blo[b][k] = INFINITY; bhi[b][k] = -INFINITY; } }
        const double sc = NBINS / ext;
        for (int32_t i = first; i < first + count; ++i) {
            const TBox* t = &tb[bvh->order[i]];
            int b = (int)((t->c[a] - clo[a]) * sc);
            if (b < 0) b = 0;
            if (b >= NBINS) b = NBINS - 1;
            cnt[b]++;
            for (int k = 0; k < 3; ++k) { if (t->lo[k] < blo[b][k]) blo[b][k] = t->lo[k]; if (t->hi[k] > bhi[b][k]) bhi[b][k] = t->hi[k]; }
        }
        double ra[NBINS]; int rc[NBINS];
        double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        int c = 0;
        for (int b = NBINS - 1; b > 0; --b) {
            for (int k = 0; k < 3; ++k) { if (blo[b][k] < lo[k]) lo[k] = blo[b][k]; if (bhi[b][k] > hi[k]) hi[k] = bhi[b][k]; }
            c += cnt[b]; ra[b] = half_area(lo, hi); rc[b] = c;
        }
        for (int k = 0; k < 3; ++k) { lo[k] = INFINITY; hi[k] = -INFINITY; }
        c = 0;
        for (int b = 0; b < NBINS - 1; ++b) {
            for (int k = 0; k < 3; ++k) { if (blo[b][k] < lo[k]) lo[k] = blo[b][k]; if (bhi[b][k] > hi[k]) hi[k] = bhi[b][k]; }
            c += cnt[b];
            if (c == 0 || rc[b + 1] == 0) continue;
            const double cost = half_area(lo, hi) * c + ra[b + 1] * rc[b + 1];
            if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = b; }
        }
    }
    int32_t mid;
    if (best_axis < 0) {
        mid = first + count / 2;   /* all centroids coincide */
    } else {
        const double ext = chi[best_axis] - clo[best_axis], sc = NBINS / ext;
        int32_t i = first, j = first + count - 1;
        while (i <= j) {
            int b = (int)((tb[bvh->order[i]].c[best_axis] - clo[best_axis]) * sc);
            if (b < 0) b = 0;
            if (b >= NBINS) b = NBINS - 1;
            if (b <= best_split) ++i;
            else { const int32_t t = bvh->order[i]; bvh->order[i] = bvh->order[j]; bvh->order[j] = t; --j; }
        }
        mid = i;
        if (mid == first || mid == first + count) mid = first + count / 2;
    }
    const int32_t l = build_rec(bvh, tb, first, mid - first, depth + 1);
    const int32_t r = build_rec(bvh, tb, mid, first + count - mid, depth + 1);
    nd = &bvh->nodes[id];
    nd->left = l; nd->right = r;
    return id;
}

void* oracle_bvh_build(const float* verts, int64_t nv, const int32_t* faces, int64_t nf) {
    OracleBvh* bvh = (OracleBvh*)calloc(1, sizeof(OracleBvh));
    bvh->mesh.verts = verts; bvh->mesh.faces = faces; bvh->mesh.nv = nv; bvh->mesh.nf = nf;
    if (nf == 0) return bvh;
    bvh->nodes = (BNode*)malloc(sizeof(BNode) * (size_t)(2 * nf));
    bvh->order = (int32_t*)malloc(sizeof(int32_t) * (size_t)nf);
    TBox* tb = (TBox*)malloc(sizeof(TBox) * (size_t)nf);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nf; ++i) { tbox_of(&bvh->mesh, (int32_t)i, &tb[i]); bvh->order[i] = (int32_t)i; }
    build_rec(bvh, tb, 0, (int32_t)nf, 0);
    free(tb);
    return bvh;
}

void oracle_bvh_free(void* p) {
    OracleBvh* bvh = (OracleBvh*)p;
    if (!bvh) return;
    free(bvh->nodes); free(bvh->order); free(bvh);
}

/* slab test in binary64 on a box padded proportionally to its distance from the origin, so
 * that every triangle either evaluator could accept is visited */
static int box_hit(const RayState* rs, const BNode* n, const double inv[3], double tlimit, double* tnear) {
    double far = 0.0;
    for (int a = 0; a < 3; ++a) {
        const double x = fabs(n->lo[a] - rs->o[a]), y = fabs(n->hi[a] - rs->o[a]);
        if (x > far) far = x;
        if (y > far) far = y;
    }
    const double pad = 1e-5 * far + 1e-30;
    double t0 = -INFINITY, t1 = INFINITY;
    for (int a = 0; a < 3; ++a) {
        const double lo = n->lo[a] - pad - rs->o[a], hi = n->hi[a] + pad - rs->o[a];
        if (rs->d[a] == 0.0) {
            if (lo > 0.0 || hi < 0.0) return 0;
            continue;
        }
        double a0 = lo * inv[a], a1 = hi * inv[a];
        if (a0 > a1) { const double s = a0; a0 = a1; a1 = s; }
        if (a0 > t0) t0 = a0;
        if (a1 < t1) t1 = a1;
    }
    *tnear = t0;
    if (t0 > t1) return 0;
    if (t1 < -1e-5 * fabs(t1) - 1e-30) return 0;
    if (t0 > tlimit) return 0;
    return 1;
}

static void trace_bvh(RayState* rs, const OracleBvh* bvh, int closest_only) {
    if (bvh->mesh.nf == 0) return;
    double inv[3];
    for (int a = 0; a < 3; ++a) inv[a] = rs->d[a] != 0.0 ? 1.0 / rs->d[a] : 0.0;
    int32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const BNode* n = &bvh->nodes[stack[--sp]];
        double tn;
        const double tlimit = closest_only ? rs->best_t * (1.0 + 1e-5) + 1e-30 : rs->tmax * (1.0 + 1e-5);
        if (!box_hit(rs, n, inv, tlimit, &tn)) continue;
        if (n->left < 0) {
            for (int32_t i = 0; i < n->count; ++i) visit_tri(rs, &bvh->mesh, bvh->order[n->first + i]);
        } else {
            double tl, tr;
            const int hl = box_hit(rs, &bvh->nodes[n->left], inv, tlimit, &tl);
            const int hr = box_hit(rs, &bvh->nodes[n->right], inv, tlimit, &tr);
            if (hl && hr) {
                if (tl < tr) { stack[sp++] = n->right; stack[sp++] = n->left; }
                else { stack[sp++] = n->left; stack[sp++] = n->right; }
            } else if (hl) stack[sp++] = n->left;
            else if (hr) stack[sp++] = n->right;
        }
        if (sp > 120) sp = 120; /* cannot happen: depth <= 61 */
    }
}

/* ---------------------------------------------------------------- public query
 * One pass produces every per-ray quantity the API can return.  Any output pointer may be NULL.
 *   mode: 0 truth, 1 mirror.   bvh: handle from oracle_bvh_build, or NULL for brute force.
 *   closest_only != 0 lets the BVH path prune beyond the nearest hit (count / list / flags are
 *   then not meaningful).
 *   list_cap (<= ORACLE_MAX_LIST): per ray, the list_cap nearest hits by (t, tri):
 *   list_tri[r*list_cap + j], list_t[...], list_loc[(r*list_cap + j)*3 ...].
 */
int oracle_query(const float* verts, int64_t nv, const int32_t* faces, int64_t nf, void* bvh_handle, int mode,
                 int closest_only, int64_t nray, const float* origins, const float* dirs, double tmax,
                 uint8_t* hit, uint8_t* front, int32_t* tri, float* loc, float* uv, double* t_out, int32_t* count,
                 uint32_t* flags, int list_cap, int32_t* list_n, int32_t* list_tri, double* list_t, float* list_loc) {
    if (list_cap > ORACLE_MAX_LIST) return -1;
    Mesh mesh; mesh.verts = verts; mesh.faces = faces; mesh.nv = nv; mesh.nf = nf;
    const OracleBvh* bvh = (const OracleBvh*)bvh_handle;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < nray; ++r) {
        HitRec list[ORACLE_MAX_LIST];
        RayState rs;
        ray_init(&rs, mode, origins + 3 * r, dirs + 3 * r, tmax, list, list_cap);
        if (bvh) trace_bvh(&rs, bvh, closest_only);
        else for (int64_t i = 0; i < nf; ++i) visit_tri(&rs, &mesh, (int32_t)i);
        const int h = rs.best_tri >= 0;
        float l[3] = {0, 0, 0}, u[2] = {0, 0};
        int fr = 0;
        if (h) {
            const float *v0 = vert(&mesh, rs.best_tri, 0), *v1 = vert(&mesh, rs.best_tri, 1), *v2 = vert(&mesh, rs.best_tri, 2);
            if (mode == 0) {
                double t, bu, bv; int uncertain, edgeon;
                tri_truth(&rs, v0, v1, v2, &t, &bu, &bv, &fr, &uncertain, &edgeon);
                const double w0 = 1.0 - bu - bv;
                for (int a = 0; a < 3; ++a) l[a] = (float)(bu * v1[a] + bv * v2[a] + w0 * v0[a]);
                u[0] = (float)w0; u[1] = (float)bu;
                if (edgeon) rs.flags |= GZ_EDGEON;
                if (rs.second_t - rs.best_t <= 1e-5 * fabs(rs.best_t)) rs.flags |= GZ_TIE;
            } else {
                TriHitF th;
                tri_mirror(&rs, v0, v1, v2, &th);
                attr_mirror(&th, v0, v1, v2, l, u);
                fr = front_mirror(&rs, &th);
            }
        }
        if (mode == 0 && rs.uncertain_tmin <= rs.best_t * (1.0 + 1e-5)) rs.flags |= GZ_EDGE_CLOSEST;
        if (hit) hit[r] = (uint8_t)h;
        if (front) front[r] = (uint8_t)fr;
        if (tri) tri[r] = rs.best_tri;
        if (loc) { loc[3 * r] = l[0]; loc[3 * r + 1] = l[1]; loc[3 * r + 2] = l[2]; }
        if (uv) { uv[2 * r] = u[0]; uv[2 * r + 1] = u[1]; }
        if (t_out) t_out[r] = h ? rs.best_t : INFINITY;
        if (count) count[r] = rs.count;
        if (flags) flags[r] = rs.flags;
        if (list_n) list_n[r] = rs.nlist;
        for (int j = 0; j < rs.nlist; ++j) {
            const int32_t ti = list[j].tri;
            if (list_tri) list_tri[r * list_cap + j] = ti;
            if (list_t) list_t[r * list_cap + j] = list[j].t;
            if (list_loc) {
                const float *v0 = vert(&mesh, ti, 0), *v1 = vert(&mesh, ti, 1), *v2 = vert(&mesh, ti, 2);
                float ll[3], uu[2];
                if (mode == 0) {
                    double t, bu, bv; int f2, un, eo;
                    tri_truth(&rs, v0, v1, v2, &t, &bu, &bv, &f2, &un, &eo);
                    const double w0 = 1.0 - bu - bv;
                    for (int a = 0; a < 3; ++a) ll[a] = (float)(bu * v1[a] + bv * v2[a] + w0 * v0[a]);
                } else {
                    TriHitF th;
                    tri_mirror(&rs, v0, v1, v2, &th);
                    attr_mirror(&th, v0, v1, v2, ll, uu);
                }
                for (int a = 0; a < 3; ++a) list_loc[(r * list_cap + j) * 3 + a] = ll[a];
            }
        }
    }
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Worker threads of the following oracle_query calls.  bench.py's CPU legs use it to take every host core even when
 * a launcher (torch.distributed.run) exported OMP_NUM_THREADS=1 into the process. */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
