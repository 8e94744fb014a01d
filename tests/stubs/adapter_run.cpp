// TEST: drive include/dbot_b200_adapter.hpp exactly as dbot's filter drives an RbSensor -- reset,
// integrated_poses(), set_observation(Observation), loglikes(deltas, indices, update) -- on a scene file
// written by tests/test_adapter_compiles.py, and dump the log-weights for comparison with the oracle.
// usage: adapter_run <scene.bin> <out.bin> [gpu ordinal ...]
#include <array>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "dbot_b200_adapter.hpp"

template <typename T> static void rd(FILE* f, T* dst, size_t n)
{
    if (fread(dst, sizeof(T), n, f) != n) { std::fprintf(stderr, "short read\n"); std::exit(2); }
}

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    typedef dbot::FreeFloatingRigidBodiesState<-1> State;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t hdr[5]; /* rows, cols, n_obj, n, frames */
    rd(f, hdr, 5);
    const int rows = hdr[0], cols = hdr[1], n_obj = hdr[2], n = hdr[3], frames = hdr[4];
    Eigen::Matrix3 K;
    rd(f, K.m, 9);
    std::vector<std::vector<std::array<double, 3> > > verts(n_obj);
    std::vector<std::vector<std::array<int, 3> > > tris(n_obj);
    for (int o = 0; o < n_obj; ++o)
    {
        int32_t cnt[2];
        rd(f, cnt, 2);
        verts[o].resize(cnt[0]);
        tris[o].resize(cnt[1]);
        for (auto& v : verts[o]) rd(f, v.data(), 3);
        for (auto& t : tris[o]) { int32_t i3[3]; rd(f, i3, 3); t = {{i3[0], i3[1], i3[2]}}; }
    }
    std::vector<int> devices;
    for (int a = 3; a < argc; ++a) devices.push_back(std::atoi(argv[a]));
    try
    {
        dbot::B200KinectImageModel<State> sensor(K, rows, cols, n, verts, tris, 0.01, 0.003, 0.0014247, 0.1, 0.7, 0.1,
                                                 1.0 / 30.0, devices.size() == 1 ? devices[0] : 0, 1.0, 6.0,
                                                 devices.size() > 1 ? devices : std::vector<int>());
        dbot::RbSensor<State>& rb = sensor;   /* the filter only ever sees the base class */
        rb.integrated_poses().resize(n_obj);
        for (int o = 0; o < n_obj; ++o)
        {
            rd(f, rb.integrated_poses()[o].p.m, 3);
            rd(f, rb.integrated_poses()[o].e.R.m, 9);
        }
        rb.reset();
        FILE* out = std::fopen(argv[2], "wb");
        dbot::RbSensor<State>::IntArray indices;
        indices.resize(n);
        for (int fr = 0; fr < frames; ++fr)
        {
            dbot::RbSensor<State>::Observation image;
            image.resize((size_t)rows * cols);
            std::vector<float> px((size_t)rows * cols);
            rd(f, px.data(), px.size());
            for (size_t i = 0; i < px.size(); ++i) image[i] = px[i];   /* fl::Real holding float32 pixels */
            dbot::RbSensor<State>::StateArray deltas;
            deltas.resize(n);
            for (int i = 0; i < n; ++i)
            {
                deltas[i].bodies.resize(n_obj);
                for (int o = 0; o < n_obj; ++o)
                {
                    rd(f, deltas[i].bodies[o].p.m, 3);
                    rd(f, deltas[i].bodies[o].e.R.m, 9);
                }
            }
            std::vector<int32_t> par(n);
            rd(f, par.data(), n);
            for (int i = 0; i < n; ++i) indices[i] = par[i];
            rb.set_observation(image);
            dbot::RbSensor<State>::RealArray ev = rb.loglikes(deltas, indices, false);
            for (int i = 0; i < n; ++i)
                if (indices[i] != par[i]) { std::fprintf(stderr, "update=false changed indices\n"); return 3; }
            dbot::RbSensor<State>::RealArray up = rb.loglikes(deltas, indices, true);
            for (int i = 0; i < n; ++i)
                if (indices[i] != i) { std::fprintf(stderr, "update=true must leave identity indices\n"); return 3; }
            std::fwrite(ev.v.data(), sizeof(double), n, out);
            std::fwrite(up.v.data(), sizeof(double), n, out);
        }
        std::fclose(out);
        /* errors arrive as exceptions, as with dbot's own GPU model */
        dbot::RbSensor<State>::StateArray too_many;
        too_many.resize(n + 1);
        for (int i = 0; i <= n; ++i) too_many[i].bodies.resize(n_obj);
        dbot::RbSensor<State>::IntArray idx2;
        idx2.resize(n + 1);
        bool thrown = false;
        try { rb.loglikes(too_many, idx2, false); } catch (const std::runtime_error&) { thrown = true; }
        if (!thrown) { std::fprintf(stderr, "no exception for n > max_sample_count\n"); return 3; }
    }
    catch (const std::exception& e)
    {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
    std::fclose(f);
    std::printf("ok\n");
    return 0;
}
