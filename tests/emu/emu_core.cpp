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
                    dbot_tri s;
                    if (!dbot_setup_triangle(tv[tri[0]], tv[tri[1]], tv[tri[2]], K, rb0, rb1, c0, c1, s))
                        continue;
                    for (int r = s.rmin; r <= s.rmax; ++r)
                        for (int c = s.cmin; c <= s.cmax; ++c)
                            if (dbot_covers(s, c, r))
                            {
                                const uint32_t w = dbot_fragment_w(s, (float)c, (float)r);
                                uint32_t* dst = &wt[(size_t)(r - rb0) * ww + (c - c0)];
                                if (w > *dst) *dst = w;
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
