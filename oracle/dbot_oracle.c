/*
 * oracle/dbot_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.   *** PARITY UNPINNED ***
 *
 * Plain-C, Eigen-free CPU restatement of dbot's depth observation model
 * (see dbot_oracle.h for provenance).  Every function names the piece of the
 * reference pipeline it follows; `dbot:` paths are in the un-vendored dbot
 * library (restated from SURVEY.md Appendix A), `ref:` paths are under
 * /root/reference.
 *
 * Build with -ffp-contract=off: the float32 renderer below spells out every
 * fused multiply-add with fmaf() so that its results are bit-identical to the
 * CUDA rasteriser's __fmaf_rn/__fmul_rn sequence.
 */
#include "dbot_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SUBPIX 256          /* fixed-point sub-pixel resolution (SURVEY.md A.6 new-build decision) */
#define SNAP_LIM 4194304.0f /* |snapped coordinate| <= 2^22 (16384 px) */
#define VERT_INVALID INT32_MIN

struct oracle_sensor
{
    oracle_config cfg;
    int P;
    float fx, sk, cx, fy, cy;          /* K rounded to float */
    float k00, k01, k02, k11, k12;     /* K^-1 (computed in double) rounded to float */
    int n_obj;
    float** verts;
    int* n_verts;
    int32_t** tris;
    int* n_tris;
    int max_verts;
    int* cull_sign;  /* per object, see oracle_cull_sign */
    /* pixel model constants */
    double lambda, wD, c, log_c;
    /* state */
    int n_slots;
    float* occ;        /* [n_slots][P] */
    double* occ_time;  /* [n_slots][P] */
    float* obs;        /* [P] */
    double t_obs;
};

/* ------------------------------------------------------------------------- */
/* dbot: model/occlusion_model.h  OcclusionModel::Condition + MapStandardGaussian
 * (SURVEY.md A.4):  o' = 1 - (c^t (1-o) + (1-p_oo)(c^t-1)/(c-1)),  c = p_oo - p_ov */
static double propagate(const oracle_sensor* s, double o, double dt)
{
    double ct = exp(dt * s->log_c);
    return 1.0 - (ct * (1.0 - o) +
                  (1.0 - s->cfg.p_occluded_occluded) * (ct - 1.0) / (s->c - 1.0));
}

/* dbot: model/kinect_pixel_model.h  KinectPixelModel::Probability, visible branch */
static double p_visible(const oracle_sensor* s, double y, double d)
{
    double sigma = s->cfg.model_sigma + s->cfg.sigma_factor * y * y;
    double sigma_sq = sigma * sigma;
    if (isinf(d)) return s->wD;
    return s->wD + (1.0 - s->cfg.tail_weight) *
                       exp(-((d - y) * (d - y) / (2.0 * sigma_sq))) /
                       sqrt(2.0 * M_PI * sigma_sq);
}

/* dbot: model/kinect_pixel_model.h  occluded branch, prediction = +inf */
static double p_occluded_inf(const oracle_sensor* s, double y)
{
    double sigma = s->cfg.model_sigma + s->cfg.sigma_factor * y * y;
    double sigma_sq = sigma * sigma;
    double l = s->lambda;
    return s->wD + (1.0 - s->cfg.tail_weight) * l * exp(0.5 * l * (-2.0 * y + l * sigma_sq));
}

/* dbot: model/kinect_pixel_model.h  occluded branch, finite prediction:
 * truncated exponential (rate lambda on [0,d]) convolved with the Gaussian noise, closed form */
static double p_occluded(const oracle_sensor* s, double y, double d)
{
    if (isinf(d)) return p_occluded_inf(s, y);
    double sigma = s->cfg.model_sigma + s->cfg.sigma_factor * y * y;
    double sigma_sq = sigma * sigma;
    double l = s->lambda;
    return s->wD + (1.0 - s->cfg.tail_weight) * l *
                       exp(0.5 * l * (2.0 * d - 2.0 * y + l * sigma_sq)) *
                       (1.0 + erf((d - y + l * sigma_sq) / sqrt(2.0 * sigma_sq))) /
                       (2.0 * (exp(d * l) - 1.0));
}

double oracle_p_visible(const oracle_sensor* s, double y, double d) { return p_visible(s, y, d); }
double oracle_p_occluded(const oracle_sensor* s, double y, double d) { return p_occluded(s, y, d); }
double oracle_p_occluded_inf(const oracle_sensor* s, double y) { return p_occluded_inf(s, y); }
double oracle_propagate(const oracle_sensor* s, double o, double dt) { return propagate(s, o, dt); }

/* ------------------------------------------------------------------------- */
/* dbot: rigid_body_renderer.cpp  RigidBodyRenderer::Render, float32.
 *
 * Follows upstream in: v' = R v + t; u = (K v'/v'_z).xy; integer pixel sample
 * points (col,row) with no half-pixel offset; closed (edge-inclusive) triangles;
 * plane n.x = 1 with n = (b-a)x(c-a)/((b-a)x(c-a).a); z = 1/(n.K^-1 [c,r,1]^T);
 * min-z; no back-face culling; depth = z along the optical axis.
 * Deviates (documented new-build decisions, identical in the CUDA path):
 *   - coverage is decided by int64 edge functions on vertices snapped to
 *     1/256 px instead of float span intersections (bit-exact CPU/GPU);
 *   - triangles with a vertex at z' <= 0 are skipped;
 *   - the z-buffer holds the inverse depth 1/z = n.K^-1 [c,r,1]^T (linear in the
 *     image plane) and keeps its maximum; the depth 1/(max 1/z) is formed once per
 *     pixel -- the same pixel wins as with min z, 1/x being monotonic;
 *   - 1/z is clamped to the triangle's own vertex range (no-op in exact
 *     arithmetic; keeps edge-on triangles from producing wild depths).
 */
typedef struct
{
    int32_t* X;
    int32_t* Y;
    float* x;
    float* y;
    float* z;
} vert_scratch;

static inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

static void transform_object(const oracle_sensor* s, int o, const float* pose, vert_scratch* vs)
{
    const float* R = pose;
    const float* t = pose + 9;
    const float* v = s->verts[o];
    for (int i = 0; i < s->n_verts[o]; ++i)
    {
        float vx = v[3 * i], vy = v[3 * i + 1], vz = v[3 * i + 2];
        float x = fmaf(R[0], vx, fmaf(R[1], vy, fmaf(R[2], vz, t[0])));
        float y = fmaf(R[3], vx, fmaf(R[4], vy, fmaf(R[5], vz, t[1])));
        float z = fmaf(R[6], vx, fmaf(R[7], vy, fmaf(R[8], vz, t[2])));
        vs->x[i] = x;
        vs->y[i] = y;
        vs->z[i] = z;
        if (z > 0.0f)
        {
            float iz = 1.0f / z;
            float xn = x * iz;
            float yn = y * iz;
            float ux = fmaf(s->fx, xn, fmaf(s->sk, yn, s->cx));
            float uy = fmaf(s->fy, yn, s->cy);
            vs->X[i] = (int32_t)lrintf(clampf(ux * (float)SUBPIX, -SNAP_LIM, SNAP_LIM));
            vs->Y[i] = (int32_t)lrintf(clampf(uy * (float)SUBPIX, -SNAP_LIM, SNAP_LIM));
        }
        else
        {
            vs->X[i] = VERT_INVALID;
            vs->Y[i] = VERT_INVALID;
        }
    }
}

static inline int32_t ceil_div_subpix(int32_t v) { return (v + (SUBPIX - 1)) >> 8; } /* arithmetic shift = floor */
static inline int32_t floor_div_subpix(int32_t v) { return v >> 8; }
static inline int32_t imin3(int32_t a, int32_t b, int32_t c) { int32_t m = a < b ? a : b; return m < c ? m : c; }
static inline int32_t imax3(int32_t a, int32_t b, int32_t c) { int32_t m = a > b ? a : b; return m > c ? m : c; }

static void raster_object(const oracle_sensor* s, int o, const vert_scratch* vs, float* wbuf)
{
    const int rows = s->cfg.rows, cols = s->cfg.cols;
    const int32_t* tri = s->tris[o];
    for (int ti = 0; ti < s->n_tris[o]; ++ti)
    {
        int ia = tri[3 * ti], ib = tri[3 * ti + 1], ic = tri[3 * ti + 2];
        int32_t Xa = vs->X[ia], Ya = vs->Y[ia];
        int32_t Xb = vs->X[ib], Yb = vs->Y[ib];
        int32_t Xc = vs->X[ic], Yc = vs->Y[ic];
        if (Xa == VERT_INVALID || Xb == VERT_INVALID || Xc == VERT_INVALID) continue;

        int64_t area2 = (int64_t)(Xb - Xa) * (int64_t)(Yc - Ya) - (int64_t)(Yb - Ya) * (int64_t)(Xc - Xa);
        if (area2 == 0) continue;
        int64_t sgn = area2 > 0 ? 1 : -1;
        if (s->cfg.cull_backfaces && s->cull_sign[o] != 0 && sgn != s->cull_sign[o]) continue;

        int32_t cmin = ceil_div_subpix(imin3(Xa, Xb, Xc));
        int32_t cmax = floor_div_subpix(imax3(Xa, Xb, Xc));
        int32_t rmin = ceil_div_subpix(imin3(Ya, Yb, Yc));
        int32_t rmax = floor_div_subpix(imax3(Ya, Yb, Yc));
        if (cmin < 0) cmin = 0;
        if (rmin < 0) rmin = 0;
        if (cmax > cols - 1) cmax = cols - 1;
        if (rmax > rows - 1) rmax = rows - 1;
        if (cmin > cmax || rmin > rmax) continue;

        float ax = vs->x[ia], ay = vs->y[ia], az = vs->z[ia];
        float e1x = vs->x[ib] - ax, e1y = vs->y[ib] - ay, e1z = vs->z[ib] - az;
        float e2x = vs->x[ic] - ax, e2y = vs->y[ic] - ay, e2z = vs->z[ic] - az;
        float nx = fmaf(e1y, e2z, -(e1z * e2y));
        float ny = fmaf(e1z, e2x, -(e1x * e2z));
        float nz = fmaf(e1x, e2y, -(e1y * e2x));
        float off = fmaf(nx, ax, fmaf(ny, ay, nz * az));
        if (off == 0.0f || !isfinite(off)) continue;
        float inv = 1.0f / off;
        nx = nx * inv;
        ny = ny * inv;
        nz = nz * inv;
        float m0 = nx * s->k00;
        float m1 = fmaf(nx, s->k01, ny * s->k11);
        float m2 = fmaf(nx, s->k02, fmaf(ny, s->k12, nz));
        float zlo = fminf(az, fminf(vs->z[ib], vs->z[ic]));
        float zhi = fmaxf(az, fmaxf(vs->z[ib], vs->z[ic]));
        float whi = 1.0f / zlo;
        float wlo = 1.0f / zhi;

        for (int32_t r = rmin; r <= rmax; ++r)
        {
            int64_t Py = (int64_t)r * SUBPIX;
            for (int32_t c = cmin; c <= cmax; ++c)
            {
                int64_t Px = (int64_t)c * SUBPIX;
                int64_t Eab = (int64_t)(Xb - Xa) * (Py - Ya) - (int64_t)(Yb - Ya) * (Px - Xa);
                int64_t Ebc = (int64_t)(Xc - Xb) * (Py - Yb) - (int64_t)(Yc - Yb) * (Px - Xb);
                int64_t Eca = (int64_t)(Xa - Xc) * (Py - Yc) - (int64_t)(Ya - Yc) * (Px - Xc);
                if (sgn * Eab < 0 || sgn * Ebc < 0 || sgn * Eca < 0) continue;
                float w = fmaf(m0, (float)c, fmaf(m1, (float)r, m2));
                w = fminf(fmaxf(w, wlo), whi);
                float* dst = &wbuf[(size_t)r * cols + c];
                if (w > *dst) *dst = w;
            }
        }
    }
}

/* full render of one particle: all objects into one z-buffer */
static void render_particle(const oracle_sensor* s, const float* pose, vert_scratch* vs, float* zbuf)
{
    for (int p = 0; p < s->P; ++p) zbuf[p] = 0.0f; /* inverse depth 0: nothing hit */
    for (int o = 0; o < s->n_obj; ++o)
    {
        transform_object(s, o, pose + 12 * o, vs);
        raster_object(s, o, vs, zbuf);
    }
    for (int p = 0; p < s->P; ++p) zbuf[p] = zbuf[p] > 0.0f ? 1.0f / zbuf[p] : INFINITY;
}

static int alloc_scratch(const oracle_sensor* s, vert_scratch* vs)
{
    size_t n = (size_t)(s->max_verts > 0 ? s->max_verts : 1);
    vs->X = (int32_t*)malloc(n * sizeof(int32_t));
    vs->Y = (int32_t*)malloc(n * sizeof(int32_t));
    vs->x = (float*)malloc(n * sizeof(float));
    vs->y = (float*)malloc(n * sizeof(float));
    vs->z = (float*)malloc(n * sizeof(float));
    return (vs->X && vs->Y && vs->x && vs->y && vs->z) ? 0 : -1;
}
static void free_scratch(vert_scratch* vs)
{
    free(vs->X); free(vs->Y); free(vs->x); free(vs->y); free(vs->z);
}

/* ------------------------------------------------------------------------- */
/* Closed-mesh proof for back-face culling (NOT in upstream dbot, whose renderer draws every
 * triangle; new-build decision documented in DESIGN.md).  Vertices with bit-identical
 * positions are welded (meshes exported with per-face vertices are still closed surfaces);
 * the mesh is closed and consistently oriented when every directed edge u->v occurs once and
 * its opposite v->u occurs once.  The orientation of each connected component is the sign of
 * its signed volume; all components must agree.  Outward-oriented (counter-clockwise seen from
 * outside, positive volume) triangles that face the camera project to negative area in an
 * image whose x grows to the right and y downwards (fx fy > 0). */
typedef struct { int32_t u, v, tri; } dedge;
static int cmp_dedge(const void* a, const void* b)
{
    const dedge* x = (const dedge*)a; const dedge* y = (const dedge*)b;
    if (x->u != y->u) return x->u < y->u ? -1 : 1;
    if (x->v != y->v) return x->v < y->v ? -1 : 1;
    return 0;
}
static const float* g_weld_v; /* qsort context (oracle_create is not re-entrant; test infrastructure) */
static int cmp_weld(const void* a, const void* b)
{
    return memcmp(g_weld_v + 3 * *(const int32_t*)a, g_weld_v + 3 * *(const int32_t*)b, 3 * sizeof(float));
}
static int32_t uf_find(int32_t* p, int32_t i)
{
    while (p[i] != i) { p[i] = p[p[i]]; i = p[i]; }
    return i;
}
static int compute_cull_sign(const float* V, int nv, const int32_t* T, int nt, double fxfy)
{
    if (nt < 4 || nv < 4) return 0;
    int result = 0;
    int32_t* order = (int32_t*)malloc(sizeof(int32_t) * nv);
    int32_t* canon = (int32_t*)malloc(sizeof(int32_t) * nv);
    dedge* E = (dedge*)malloc(sizeof(dedge) * 3 * (size_t)nt);
    int32_t* comp = (int32_t*)malloc(sizeof(int32_t) * nt);
    double* vol = (double*)calloc((size_t)nt, sizeof(double));
    if (!order || !canon || !E || !comp || !vol) goto done;
    for (int i = 0; i < nv; ++i) order[i] = i;
    g_weld_v = V;
    qsort(order, nv, sizeof(int32_t), cmp_weld);
    for (int i = 0; i < nv; ++i)
        canon[order[i]] = (i > 0 && cmp_weld(&order[i], &order[i - 1]) == 0) ? canon[order[i - 1]] : order[i];
    for (int t = 0; t < nt; ++t)
    {
        int32_t a = canon[T[3 * t]], b = canon[T[3 * t + 1]], c = canon[T[3 * t + 2]];
        if (a == b || b == c || c == a) goto done; /* collapsed triangle: not provably closed */
        E[3 * t] = (dedge){a, b, t}; E[3 * t + 1] = (dedge){b, c, t}; E[3 * t + 2] = (dedge){c, a, t};
        comp[t] = t;
    }
    qsort(E, 3 * (size_t)nt, sizeof(dedge), cmp_dedge);
    for (size_t i = 1; i < 3 * (size_t)nt; ++i)
        if (cmp_dedge(&E[i], &E[i - 1]) == 0) goto done; /* a directed edge used twice */
    for (size_t i = 0; i < 3 * (size_t)nt; ++i)
    {
        dedge key = {E[i].v, E[i].u, 0};
        const dedge* m = (const dedge*)bsearch(&key, E, 3 * (size_t)nt, sizeof(dedge), cmp_dedge);
        if (!m) goto done; /* boundary edge: open mesh */
        int32_t ra = uf_find(comp, E[i].tri), rb = uf_find(comp, m->tri);
        if (ra != rb) comp[ra] = rb;
    }
    for (int t = 0; t < nt; ++t)
    {
        const float* a = V + 3 * T[3 * t]; const float* b = V + 3 * T[3 * t + 1]; const float* c = V + 3 * T[3 * t + 2];
        double cx = (double)b[1] * c[2] - (double)b[2] * c[1];
        double cy = (double)b[2] * c[0] - (double)b[0] * c[2];
        double cz = (double)b[0] * c[1] - (double)b[1] * c[0];
        vol[uf_find(comp, t)] += a[0] * cx + a[1] * cy + a[2] * cz;
    }
    {
        int sgn = 0;
        for (int t = 0; t < nt; ++t)
        {
            if (uf_find(comp, t) != t) continue;
            int sv = vol[t] > 0.0 ? 1 : (vol[t] < 0.0 ? -1 : 0);
            if (sv == 0 || (sgn != 0 && sv != sgn)) goto done;
            sgn = sv;
        }
        result = (fxfy > 0.0) ? -sgn : sgn;
    }
done:
    free(order); free(canon); free(E); free(comp); free(vol);
    return result;
}

int oracle_cull_sign(const oracle_sensor* s, int32_t o)
{
    return (s && o >= 0 && o < s->n_obj) ? s->cull_sign[o] : 0;
}

/* ------------------------------------------------------------------------- */
int oracle_create(const oracle_config* cfg, oracle_sensor** out)
{
    if (!cfg || !out) return -1;
    if (cfg->rows <= 0 || cfg->cols <= 0 || cfg->n_objects <= 0) return -2;
    if (cfg->K[3] != 0.0 || cfg->K[6] != 0.0 || cfg->K[7] != 0.0 || cfg->K[8] != 1.0) return -3;
    if (cfg->K[0] == 0.0 || cfg->K[4] == 0.0) return -3;
    double c = cfg->p_occluded_occluded - cfg->p_occluded_visible;
    if (!(c > 0.0) || c == 1.0) return -4;
    oracle_sensor* s = (oracle_sensor*)calloc(1, sizeof(oracle_sensor));
    if (!s) return -5;
    s->cfg = *cfg;
    s->P = cfg->rows * cfg->cols;
    double fx = cfg->K[0], sk = cfg->K[1], cx = cfg->K[2], fy = cfg->K[4], cy = cfg->K[5];
    s->fx = (float)fx; s->sk = (float)sk; s->cx = (float)cx; s->fy = (float)fy; s->cy = (float)cy;
    s->k00 = (float)(1.0 / fx);
    s->k01 = (float)(-sk / (fx * fy));
    s->k02 = (float)((sk * cy - cx * fy) / (fx * fy));
    s->k11 = (float)(1.0 / fy);
    s->k12 = (float)(-cy / fy);
    s->n_obj = cfg->n_objects;
    s->verts = (float**)calloc(s->n_obj, sizeof(float*));
    s->tris = (int32_t**)calloc(s->n_obj, sizeof(int32_t*));
    s->n_verts = (int*)calloc(s->n_obj, sizeof(int));
    s->n_tris = (int*)calloc(s->n_obj, sizeof(int));
    s->cull_sign = (int*)calloc(s->n_obj, sizeof(int));
    for (int o = 0; o < s->n_obj; ++o)
    {
        int nv = cfg->n_vertices[o], nt = cfg->n_triangles[o];
        if (nv < 0 || nt < 0) return -2;
        s->n_verts[o] = nv;
        s->n_tris[o] = nt;
        if (nv > s->max_verts) s->max_verts = nv;
        s->verts[o] = (float*)malloc(sizeof(float) * 3 * (nv > 0 ? nv : 1));
        s->tris[o] = (int32_t*)malloc(sizeof(int32_t) * 3 * (nt > 0 ? nt : 1));
        memcpy(s->verts[o], cfg->vertices[o], sizeof(float) * 3 * nv);
        memcpy(s->tris[o], cfg->triangles[o], sizeof(int32_t) * 3 * nt);
        for (int i = 0; i < 3 * nt; ++i)
            if (s->tris[o][i] < 0 || s->tris[o][i] >= nv) return -6;
        s->cull_sign[o] = compute_cull_sign(s->verts[o], nv, s->tris[o], nt, fx * fy);
    }
    s->cfg.vertices = NULL; s->cfg.triangles = NULL; s->cfg.n_vertices = NULL; s->cfg.n_triangles = NULL;
    /* dbot KinectPixelModel: exponential_rate = -log(0.5)/half_life_depth */
    s->lambda = -log(0.5) / cfg->half_life_depth;
    s->wD = cfg->tail_weight / cfg->max_depth;
    s->c = c;
    s->log_c = log(c);
    s->obs = (float*)malloc(sizeof(float) * s->P);
    for (int p = 0; p < s->P; ++p) s->obs[p] = NAN;
    *out = s;
    return oracle_reset(s);
}

int oracle_destroy(oracle_sensor* s)
{
    if (!s) return -1;
    for (int o = 0; o < s->n_obj; ++o) { free(s->verts[o]); free(s->tris[o]); }
    free(s->verts); free(s->tris); free(s->n_verts); free(s->n_tris); free(s->cull_sign);
    free(s->occ); free(s->occ_time); free(s->obs);
    free(s);
    return 0;
}

/* dbot: KinectImageModel::reset: one slot, occ = initial prob, times 0, t_obs 0 */
int oracle_reset(oracle_sensor* s)
{
    if (!s) return -1;
    free(s->occ); free(s->occ_time);
    s->n_slots = 1;
    s->occ = (float*)malloc(sizeof(float) * s->P);
    s->occ_time = (double*)malloc(sizeof(double) * s->P);
    if (!s->occ || !s->occ_time) return -5;
    for (int p = 0; p < s->P; ++p)
    {
        s->occ[p] = (float)s->cfg.initial_occlusion_prob;
        s->occ_time[p] = 0.0;
    }
    s->t_obs = 0.0;
    return 0;
}

/* dbot: KinectImageModel::set_observation; input layout row*cols+col, metres,
 * NaN invalid (ref:source/dbot_ros/util/ros_interface.h:152-168) */
int oracle_set_observation(oracle_sensor* s, const float* depth_m)
{
    if (!s || !depth_m) return -1;
    memcpy(s->obs, depth_m, sizeof(float) * s->P);
    s->t_obs += s->cfg.delta_time;
    return 0;
}

/* dbot: KinectImageModel::loglikes */
int oracle_loglikes(oracle_sensor* s, const float* poses, int32_t* indices, int32_t n,
                    int32_t update, double* out)
{
    if (!s || !poses || !indices || !out || n < 0) return -1;
    for (int i = 0; i < n; ++i)
        if (indices[i] < 0 || indices[i] >= s->n_slots) return -7;
    const int P = s->P;
    float* new_occ = NULL;
    double* new_time = NULL;
    if (update && n > 0)
    {
        new_occ = (float*)malloc(sizeof(float) * (size_t)P * n);
        new_time = (double*)malloc(sizeof(double) * (size_t)P * n);
        if (!new_occ || !new_time) { free(new_occ); free(new_time); return -5; }
    }
    int fail = 0;
#pragma omp parallel
    {
        vert_scratch vs;
        float* zbuf = (float*)malloc(sizeof(float) * P);
        int ok = (alloc_scratch(s, &vs) == 0) && zbuf;
        if (!ok)
        {
#pragma omp atomic write
            fail = 1;
        }
#pragma omp for schedule(dynamic, 1)
        for (int i = 0; i < n; ++i)
        {
            if (!ok) continue;
            const int slot = indices[i];
            const float* occ = s->occ + (size_t)slot * P;
            const double* occ_t = s->occ_time + (size_t)slot * P;
            float* nocc = update ? new_occ + (size_t)i * P : NULL;
            double* ntime = update ? new_time + (size_t)i * P : NULL;
            if (update)
            {
                if (s->cfg.eager)
                {
                    /* dbot GPU evaluator: the whole image is propagated to t_obs on update */
                    for (int p = 0; p < P; ++p)
                    {
                        nocc[p] = (float)propagate(s, (double)occ[p], s->t_obs - occ_t[p]);
                        ntime[p] = s->t_obs;
                    }
                }
                else
                {
                    memcpy(nocc, occ, sizeof(float) * P);
                    memcpy(ntime, occ_t, sizeof(double) * P);
                }
            }
            render_particle(s, poses + (size_t)i * 12 * s->n_obj, &vs, zbuf);
            double L = 0.0;
            for (int p = 0; p < P; ++p)
            {
                float d = zbuf[p];
                if (!(d < INFINITY)) continue;
                float yf = s->obs[p];
                if (isnan(yf)) continue; /* no data: occlusion left as is */
                double y = (double)yf;
                double dt = s->t_obs - occ_t[p];
                double o = propagate(s, (double)occ[p], dt);
                if (s->cfg.float_intermediates || s->cfg.eager) o = (double)(float)o;
                double pv = p_visible(s, y, (double)d) * (1.0 - o);
                double po = p_occluded(s, y, (double)d) * o;
                double pinf = p_occluded_inf(s, y);
                if (s->cfg.float_intermediates)
                {
                    pv = (double)(float)pv; po = (double)(float)po; pinf = (double)(float)pinf;
                }
                L += log((pv + po) / pinf);
                if (update)
                {
                    nocc[p] = (float)(po / (pv + po));
                    ntime[p] = s->t_obs;
                }
            }
            out[i] = L;
        }
        free(zbuf);
        if (ok) free_scratch(&vs);
    }
    if (fail) { free(new_occ); free(new_time); return -5; }
    if (update && n > 0)
    {
        free(s->occ); free(s->occ_time);
        s->occ = new_occ;
        s->occ_time = new_time;
        s->n_slots = n;
        for (int i = 0; i < n; ++i) indices[i] = i;
    }
    return 0;
}

int oracle_render(oracle_sensor* s, const float* pose, float* out_depth)
{
    if (!s || !pose || !out_depth) return -1;
    vert_scratch vs;
    if (alloc_scratch(s, &vs) != 0) return -5;
    render_particle(s, pose, &vs, out_depth);
    free_scratch(&vs);
    return 0;
}

int oracle_n_slots(const oracle_sensor* s) { return s ? s->n_slots : -1; }
double oracle_observation_time(const oracle_sensor* s) { return s ? s->t_obs : NAN; }

int oracle_get_occlusion(const oracle_sensor* s, int32_t slot, float* out_occ)
{
    if (!s || !out_occ || slot < 0 || slot >= s->n_slots) return -1;
    const float* occ = s->occ + (size_t)slot * s->P;
    const double* t = s->occ_time + (size_t)slot * s->P;
    for (int p = 0; p < s->P; ++p)
        out_occ[p] = (t[p] == s->t_obs) ? occ[p] : (float)propagate(s, (double)occ[p], s->t_obs - t[p]);
    return 0;
}

int oracle_get_occlusion_raw(const oracle_sensor* s, int32_t slot, float* out_occ, double* out_time)
{
    if (!s || slot < 0 || slot >= s->n_slots) return -1;
    if (out_occ) memcpy(out_occ, s->occ + (size_t)slot * s->P, sizeof(float) * s->P);
    if (out_time) memcpy(out_time, s->occ_time + (size_t)slot * s->P, sizeof(double) * s->P);
    return 0;
}

/* TEST SUPPORT for the particle-sharded evaluator: append a slot holding an occlusion image (and its
 * per-pixel time stamps) exported from another oracle instance with oracle_get_occlusion_raw.
 * Returns the new slot index. */
int oracle_import_slot(oracle_sensor* s, const float* occ, const double* occ_time)
{
    if (!s || !occ || !occ_time) return -1;
    const size_t P = (size_t)s->P, n = (size_t)s->n_slots;
    float* no = (float*)realloc(s->occ, sizeof(float) * P * (n + 1));
    if (!no) return -5;
    s->occ = no;
    double* nt = (double*)realloc(s->occ_time, sizeof(double) * P * (n + 1));
    if (!nt) return -5;
    s->occ_time = nt;
    memcpy(s->occ + P * n, occ, sizeof(float) * P);
    memcpy(s->occ_time + P * n, occ_time, sizeof(double) * P);
    s->n_slots = (int)n + 1;
    return (int)n;
}

void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int oracle_get_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
