/*
 * tests/emu/emu_core.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Compiles dbot_ros_b200/csrc/particle_core.cuh (the per-particle phase functions of the CUDA
 * kernel) as plain C++ and walks one particle through them sequentially, the way one CTA does
 * through them sequentially.  It lets the CPU-only test suite check the kernel's integer raster
 * arithmetic, window/band logic and float32 mixture algebra against oracle/ without a GPU.  It is
 * built by tests/conftest.py into tests/emu/, is never imported by the package, and is not a
 * fallback: dbot_ros_b200 refuses to run without libdbot_b200.so and a CUDA device.
 */
#define DBOT_CORE_HOST 1
#include "../../dbot_ros_b200/csrc/particle_core.cuh"

#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

extern "C" {

struct emu_config
{
    int32_t rows, cols;
    double K[9];
    int32_t n_objects;
    const float* const* vertices;
    const int32_t* n_vertices;
    const int32_t* const* triangles;
    const int32_t* n_triangles;
    double tail_weight, model_sigma, sigma_factor, half_life_depth, max_depth;
    double p_occluded_visible, p_occluded_occluded;
    int32_t zcap; /* z tile capacity in pixels: small values force the multi-band path */
    int32_t cull_sign[8]; /* per object: sign of the snapped area of the triangles that are kept, 0 = all */
    int32_t force_blocks; /* 1: every triangle through the int64 block walk (dbot_covers), as the
                             kernel does for DBOT_TRI_WIDE triangles */
};

/* observation table exactly as dbot_observation_kernel computes it */
void emu_observation_table(const emu_config* cfg, const float* depth, float* tab /* [P][4] */)
{
    const double lambda = -log(0.5) / cfg->half_life_depth;
    const double wD = cfg->tail_weight / cfg->max_depth;
    const int P = cfg->rows * cfg->cols;
    for (int i = 0; i < P; ++i)
    {
        float* t = tab + 4 * i;
        t[0] = NAN; t[1] = t[2] = t[3] = 0.f;
        const float yf = depth[i];
        if (yf != yf) continue;
        const double y = yf;
        const double sigma = cfg->model_sigma + cfg->sigma_factor * y * y;
        const double g = 1.0 / (1.4142135623730951 * sigma);
        const double cv = (1.0 - cfg->tail_weight) / (2.5066282746310002 * sigma);
        const double co = 0.5 * (1.0 - cfg->tail_weight) * lambda *
                          exp(0.5 * lambda * (-2.0 * y + lambda * sigma * sigma));
        const double pinf = wD + 2.0 * co;
        t[0] = yf; t[1] = (float)g; t[2] = (float)(cv / pinf); t[3] = (float)(wD / pinf);
    }
}

/* one particle: returns the log-weight; if occ_out != NULL also writes the child's occlusion
 * image (update).  depth_out (optional) receives the rendered depth (+inf = empty). */
double emu_particle(const emu_config* cfg, const float* pose, const float* tab, const float* occ_in,
                    float prop_a, float prop_b, float* occ_out, float* depth_out)
{
    const int rows = cfg->rows, cols = cfg->cols, P = rows * cols;
    dbot_camera K;
    const double fx = cfg->K[0], sk = cfg->K[1], cx = cfg->K[2], fy = cfg->K[4], cy = cfg->K[5];
    K.fx = (float)fx; K.sk = (float)sk; K.cx = (float)cx; K.fy = (float)fy; K.cy = (float)cy;
    K.k00 = (float)(1.0 / fx);
    K.k01 = (float)(-sk / (fx * fy));
    K.k02 = (float)((sk * cy - cx * fy) / (fx * fy));
    K.k11 = (float)(1.0 / fy);
    K.k12 = (float)(-cy / fy);
    const double lambda = -log(0.5) / cfg->half_life_depth;
    dbot_pixel_model M;
    M.lam_rsqrt2 = (float)(lambda / 1.4142135623730951);
    M.nlam_log2e = (float)(-lambda * 1.4426950408889634);
    M.model_sigma = (float)cfg->model_sigma;
    M.sigma_factor = (float)cfg->sigma_factor;
    M.prop_a = prop_a;
    M.prop_b = prop_b;

    /* per-object transformed vertices (the kernel re-stages chunks; here each object is one) */
    std::vector<std::vector<dbot_vertex>> objs(cfg->n_objects);
    int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
    for (int o = 0; o < cfg->n_objects; ++o)
    {
        const int nv = cfg->n_vertices[o];
        objs[o].resize(nv);
        for (int i = 0; i < nv; ++i)
        {
            const float* v = cfg->vertices[o] + 3 * i;
            const dbot_vertex t = dbot_transform_vertex(v[0], v[1], v[2], pose + 12 * o, K);
            objs[o][i] = t;
            if (t.X != DBOT_VERT_INVALID)
            {
                mnx = t.X < mnx ? t.X : mnx; mxx = t.X > mxx ? t.X : mxx;
                mny = t.Y < mny ? t.Y : mny; mxy = t.Y > mxy ? t.Y : mxy;
            }
        }
    }
    if (depth_out)
        for (int i = 0; i < P; ++i) depth_out[i] = INFINITY;
    int c0 = 0, c1 = -1, r0 = 0, r1 = -1;
    if (mnx <= mxx)
    {
        c0 = dbot_ceil_subpix(mnx); if (c0 < 0) c0 = 0;
        c1 = dbot_floor_subpix(mxx); if (c1 > cols - 1) c1 = cols - 1;
        r0 = dbot_ceil_subpix(mny); if (r0 < 0) r0 = 0;
        r1 = dbot_floor_subpix(mxy); if (r1 > rows - 1) r1 = rows - 1;
    }
    const bool have_window = c0 <= c1 && r0 <= r1;
    const int ww = c1 - c0 + 1;
    double acc = 0.0;
    /* update: the streaming warps write the propagated parent image everywhere first */
    if (occ_out)
        for (int i = 0; i < P; ++i) occ_out[i] = dbot_fmaf(prop_a, occ_in[i], prop_b);
    if (have_window)
    {
        int band_rows = cfg->zcap / ww; if (band_rows < 1) band_rows = 1;
        std::vector<uint32_t> wt((size_t)band_rows * ww);
        for (int rb0 = r0; rb0 <= r1; rb0 += band_rows)
        {
            const int rb1 = rb0 + band_rows - 1 < r1 ? rb0 + band_rows - 1 : r1;
            const int npx = (rb1 - rb0 + 1) * ww;
            for (int q = 0; q < npx; ++q) wt[q] = DBOT_W_EMPTY;
            for (int o = 0; o < cfg->n_objects; ++o)
            {
                const std::vector<dbot_vertex>& tv = objs[o];
                for (int t = 0; t < cfg->n_triangles[o]; ++t)
                {
                    const int32_t* tri = cfg->triangles[o] + 3 * t;
                    const dbot_vertex &va = tv[tri[0]], &vb = tv[tri[1]], &vc = tv[tri[2]];
                    int cmin, cmax, rmin, rmax;
                    int kind = dbot_prepass_triangle(va.X, va.Y, vb.X, vb.Y, vc.X, vc.Y, cfg->cull_sign[o],
                                                     rb0, rb1, c0, c1, cmin, cmax, rmin, rmax);
                    if (kind == DBOT_TRI_REJECT) continue;
                    if (kind == DBOT_TRI_WIDE || cfg->force_blocks)
                    {
                        dbot_tri s;
                        if (!dbot_setup_triangle(va, vb, vc, K, rb0, rb1, c0, c1, s)) continue;
                        for (int r = s.rmin; r <= s.rmax; ++r)
                            for (int c = s.cmin; c <= s.cmax; ++c)
                                if (dbot_covers(s, c, r))
                                {
                                    const uint32_t w = dbot_fragment_w(s, (float)c, (float)r);
                                    uint32_t* dst = &wt[(size_t)(r - rb0) * ww + (c - c0)];
                                    if (w > *dst) *dst = w;
                                }
                        continue;
                    }
                    /* (triangle, row) span units, as the kernel's lanes walk them */
                    dbot_span_rec s;
                    dbot_setup_spans(va, vb, vc, K, cmin, cmax, rmin, rmax, s);
                    const int co = s.org & 0xffff, ro = s.org >> 16;
                    const int bwm1 = s.aux & 0xff, split = (s.aux >> 8) & 0xff, long_left = (s.aux >> 16) & 1;
                    for (int j = 0; j <= rmax - rmin; ++j)
                    {
                        const dbot_edge& B = j >= split ? s.B2 : s.B1;
                        int first;
                        const int n = long_left ? dbot_row_span(s.A, B, j, bwm1, first)
                                                : dbot_row_span(B, s.A, j, bwm1, first);
                        const int r = ro + j;
                        const float wr = dbot_fmaf(s.m1, (float)r, s.m2);
                        for (int k = 0; k < n; ++k)
                        {
                            const int c = co + first + k;
                            const uint32_t w = dbot_span_w(s, (float)c, wr);
                            uint32_t* dst = &wt[(size_t)(r - rb0) * ww + (c - c0)];
                            if (w > *dst) *dst = w;
                        }
                    }
                }
            }
            float facc = 0.f; /* the kernel accumulates float partials per thread */
            int cnt = 0;
            for (int q = 0; q < npx; ++q)
            {
                const int rr = q / ww, cc = q - rr * ww;
                const int gp = (rb0 + rr) * cols + c0 + cc;
                const uint32_t wb = wt[q];
                if (depth_out) depth_out[gp] = dbot_depth_from_w(wb);
                if (!tab || wb == DBOT_W_EMPTY) continue;
                float4 T = {tab[4 * gp], tab[4 * gp + 1], tab[4 * gp + 2], tab[4 * gp + 3]};
                if (T.x != T.x) continue;
                float o_new;
                facc += dbot_pixel_term(dbot_rcp(dbot_u2f(wb)), T, occ_in[gp], M, o_new);
                if (++cnt == 48) { acc += facc; facc = 0.f; cnt = 0; }
                if (occ_out) occ_out[gp] = o_new;
            }
            acc += facc;
        }
    }
    return acc;
}

/* coverage of one snapped triangle inside rows [row0,row1] x columns [col0,col1], two ways: the
 * span units (dbot_setup_spans + dbot_row_span) and the direct int64 edge functions
 * (dbot_covers).  Returns the number of pixels on which they disagree (-1: the triangle is not
 * one the span units take); *n_cov receives the number of covered pixels. */
int emu_span_mismatches(const int32_t* XY /* Xa Ya Xb Yb Xc Yc */, int row0, int row1, int col0, int col1,
                        int32_t* n_cov)
{
    dbot_vertex v[3];
    for (int k = 0; k < 3; ++k)
    {
        v[k].x = (float)k; v[k].y = (float)(k * k); v[k].z = 1.0f + k; v[k].iz = 1.0f / v[k].z;
        v[k].X = XY[2 * k]; v[k].Y = XY[2 * k + 1];
    }
    int cmin, cmax, rmin, rmax;
    *n_cov = 0;
    const int kind = dbot_prepass_triangle(v[0].X, v[0].Y, v[1].X, v[1].Y, v[2].X, v[2].Y, 0, row0, row1, col0,
                                           col1, cmin, cmax, rmin, rmax);
    if (kind != DBOT_TRI_SPANS && kind != DBOT_TRI_TALL) return -1;
    dbot_camera K = {1.f, 0.f, 0.f, 1.f, 0.f, 1.f, 0.f, 0.f, 1.f, 0.f};
    dbot_tri t;
    t.Xa = v[0].X; t.Ya = v[0].Y; t.Xb = v[1].X; t.Yb = v[1].Y; t.Xc = v[2].X; t.Yc = v[2].Y;
    const long long a2 = (long long)(t.Xb - t.Xa) * (t.Yc - t.Ya) - (long long)(t.Yb - t.Ya) * (t.Xc - t.Xa);
    t.flip = a2 < 0;
    dbot_span_rec s;
    dbot_setup_spans(v[0], v[1], v[2], K, cmin, cmax, rmin, rmax, s);
    const int bwm1 = s.aux & 0xff, split = (s.aux >> 8) & 0xff, long_left = (s.aux >> 16) & 1;
    int bad = 0;
    for (int j = 0; j <= rmax - rmin; ++j)
    {
        const dbot_edge& B = j >= split ? s.B2 : s.B1;
        int first;
        const int n = long_left ? dbot_row_span(s.A, B, j, bwm1, first) : dbot_row_span(B, s.A, j, bwm1, first);
        for (int i = 0; i <= cmax - cmin; ++i)
        {
            const bool in_span = n > 0 && i >= first && i < first + n;
            const bool cov = dbot_covers(t, cmin + i, rmin + j);
            bad += in_span != cov;
            *n_cov += cov;
        }
    }
    return bad;
}

/* the scalar mixture term alone, for tests/test_pixel_math.py */
float emu_pixel_term(const emu_config* cfg, float d, float y, float o_in, float prop_a, float prop_b,
                     float* o_post)
{
    emu_config c1 = *cfg;
    c1.rows = 1; c1.cols = 1;
    float tab[4];
    emu_observation_table(&c1, &y, tab);
    const double lambda = -log(0.5) / cfg->half_life_depth;
    dbot_pixel_model M;
    M.lam_rsqrt2 = (float)(lambda / 1.4142135623730951);
    M.nlam_log2e = (float)(-lambda * 1.4426950408889634);
    M.model_sigma = (float)cfg->model_sigma;
    M.sigma_factor = (float)cfg->sigma_factor;
    M.prop_a = prop_a; M.prop_b = prop_b;
    float4 T = {tab[0], tab[1], tab[2], tab[3]};
    return dbot_pixel_term(d, T, o_in, M, *o_post);
}
}
