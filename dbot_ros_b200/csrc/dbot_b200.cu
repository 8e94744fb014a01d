/*
 * dbot_b200.cu -- libdbot_b200.so: the B200 (sm_100a) depth observation model of the
 * dbot / dbot_ros particle tracker behind the C-ABI of include/dbot_b200.h.
 *
 * Replaces dbot's KinectImageModelGPU (OpenGL rasteriser + CUDA evaluator), which dbot_ros
 * selects with use_gpu:true (ref:config/particle_tracker.yaml:3,
 * ref:source/dbot_ros/tracker/particle_tracker_node.cpp:164-203).  Nothing here is derived
 * from that code: one persistent CTA per SM walks the particles; per particle it
 *   1. stages the mesh chunk's uint16 triangle list in shared memory with a TMA bulk copy
 *      (cp.async.bulk + mbarrier) -- resident across particles when the mesh is a single chunk,
 *   2. transforms + projects + snaps the vertices (shared memory SoA),
 *   3. rasterises all triangles into a shared-memory tile of inverse depths covering the
 *      particle's screen window: exact integer edge functions, branch-free shared-memory max on
 *      the bits of the positive float,
 *   4. evaluates the per-pixel visible/occluded mixture against the per-frame observation
 *      table and the parent's stored occlusion image, and on update writes the posterior of the
 *      covered pixels into the child's slot,
 *   5. reduces to one double log-weight with a fixed shuffle tree (deterministic).
 * The occlusion state lives in a pool of physical slots: propagation is composed lazily into two
 * numbers, the last child of every parent adopts the parent's slot in place and only the other
 * children get a copy, streamed by four extra warps while the compute warps rasterise
 * (dbot_plan_update_kernel).  dbot_filter.cuh adds the device-resident filter step around it.
 * There is no CPU fallback in this library.
 */
#include "../../include/dbot_b200.h"
#include "particle_core.cuh"

#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <unordered_map>
#include <vector>

#ifndef DBOT_CW
#define DBOT_CW 24                 /* compute warps per CTA: the particle kernel needs well under 80 registers per
                                      thread, so 24 warps (+4 streaming) fit the register file; measured 7 % faster
                                      than 16 warps at 110 registers (issue slots are the limiter), 28 no better */
#endif
#define DBOT_CT (DBOT_CW * 32)     /* compute threads */
#ifndef DBOT_SW
#define DBOT_SW 4                  /* streaming warps per CTA (update launches only) */
#endif
#define DBOT_ST (DBOT_SW * 32)
#ifndef DBOT_STREAM_UNROLL
#define DBOT_STREAM_UNROLL 4       /* 256-bit loads in flight per streaming thread */
#endif
/* update launches start every thread with 72 registers (896 threads); the streaming warp group
 * gives back what it does not need (setmaxnreg) */
#ifndef DBOT_PREFETCH_AHEAD
#define DBOT_PREFETCH_AHEAD 8      /* chunks of DBOT_ST x DBOT_STREAM_UNROLL x 32 B (16 KB) */
#endif
#ifndef DBOT_STREAM_REGS
#define DBOT_STREAM_REGS 64
#define DBOT_COMPUTE_REGS 72
#endif
#define DBOT_STAGERS 2             /* host threads beside the caller's that stage a pageable depth frame */
#define DBOT_STAGE_PIECES 8        /* ... in this many pieces (150 KB each at 640x480) */
#define DBOT_CHUNK_MAX_VERTS 3072
#define DBOT_CHUNK_MAX_TRIS 6144
#define DBOT_FULL 0xffffffffu
#ifndef DBOT_UNIT_PX
#define DBOT_UNIT_PX 4             /* pixels per lane and unit of the mixture pass */
#endif

enum
{
    DBOT_MODE_EVAL = 0,
    DBOT_MODE_UPDATE = 1,
    DBOT_MODE_RENDER = 2
};

struct dbot_chunk
{
    int32_t obj, n_verts, n_tris;
    int32_t cull_sign; /* back-face culling: sign of the snapped area of the triangles to keep, 0 = all */
    uint32_t vert_off, vert_bytes; /* byte offsets into the packed device mesh, 16 B aligned */
    uint32_t tri_off, tri_bytes;
};

struct dbot_kparams
{
    int rows, cols, pitch; /* pitch: floats per state / table row, multiple of 8 */
    int n_obj, n_chunks;
    int n, n_slots;
    dbot_camera K;
    dbot_pixel_model M;
    const dbot_chunk* chunks;
    const unsigned char* mesh;
    const float4* tab;      /* [rows*pitch] observation table */
    float* pool;            /* [2 * max_particles][slot_stride] physical occlusion slots */
    const int32_t* loc;     /* [n_slots] logical slot -> physical slot (current generation) */
    const int4* plan;       /* update: per particle { source slot, destination slot, parent, 0 } (physical) */
    const int32_t* order;   /* update: processing order, copying children before adopting ones */
    int32_t* pending;       /* update: per parent, copies of its image still to be read */
    int4* slot_rect;        /* [2 * max_particles] per PHYSICAL slot: { first row, last row, first word, last word } of the
                               rectangle in which the image may differ from the background value (words = columns
                               / 8, the 256-bit units of the copy; empty: first > last).  Every stored image equals
                               the background outside it, so a copy moves only the parent's rectangle. */
    float background;       /* stored value of a never-updated pixel (lazy encoding: constant between rebases) */
    size_t slot_stride;     /* rows*pitch */
    float post_b, post_inv_a; /* update: the posterior o is stored as (o - post_b) * post_inv_a */
    const float* poses;     /* [n][n_obj][12] */
    int32_t* indices;       /* [n] */
    double* out;            /* [n] */
    double* peer_out[DBOT_B200_MAX_DEVICES]; /* fused gather: the same log-weight is also stored at
                                                out_offset + i of every device's gathered vector (peer memory) */
    int n_peers, out_offset;
    float* depth_out;       /* render: [rows*cols] */
    int* err_flag;
    /* shared-memory carve-up (bytes from the dynamic smem base) */
    uint32_t off_tri, off_v4, off_XY, off_list, off_wt;
    uint32_t wcap;          /* inverse-depth tile capacity in pixels */
};

#include "host_workers.cuh"
#include "ptx_wrappers.cuh"
#include "state_kernels.cuh"
#include "particle_kernel.cuh"

/* ========================================================================================== */
/* host side                                                                                  */
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                         \
    do                                                                                         \
    {                                                                                          \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail((int)e__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),     \
                        __FILE__, __LINE__);                                                   \
    } while (0)

struct dbot_b200_sensor
{
    dbot_b200_config cfg;
    int device;
    cudaStream_t stream;
    dbot_kparams kp; /* template: per-call fields are filled at launch */
    int grid_max, smem_bytes, num_sms;
    int plan_smem_max; /* dynamic shared memory the update-planning kernel may use for its tables */
    int cull_signs[DBOT_B200_MAX_OBJECTS];
    /* model constants */
    double lambda, wD, c, log_c;
    /* device memory */
    unsigned char* d_mesh;
    dbot_chunk* d_chunks;
    float* d_obs;
    float4* d_tab;
    float* d_pool;          /* 2 * max_particles physical occlusion slots */
    int32_t* d_loc[2];      /* logical -> physical slot, current / next generation */
    int32_t* d_free;        /* stack of unused physical slots + its height */
    int32_t* d_free_count;
    int32_t *d_cnt, *d_last, *d_pending, *d_order;
    int4* d_slot_rect;      /* touched rectangle of every physical slot */
    double background;      /* stored value of untouched pixels */
    int4* d_plan;
    float* d_poses;         /* poses of a host call and, behind them, its parent indices */
    double* d_out;
    double* d_result;       /* { error flag (in the first word), log-weights }: one copy back per call */
    float* d_depth;
    int* d_err;
    /* pinned staging */
    float* h_poses;
    double* h_out;
    double* h_result;
    float* h_obs;
    float* h_depth;
    int* h_err;
    /* state */
    int cur, n_slots;       /* cur: which d_loc is current */
    double t_obs, t_update;
    /* lazy propagation: a stored value r of any slot stands for the occlusion probability
     * A * r + B as of t_update.  dbot's OcclusionModel propagates by an affine map that is the same
     * for every pixel of every slot, so the maps of successive updates compose into these two
     * numbers (kept in double) instead of being applied to 4*P bytes per particle per frame; an
     * update copies stored values verbatim and encodes the posterior of the covered pixels as
     * (o - B) / A.  Equivalent to the eager propagation of dbot's GPU evaluator. */
    double A, B;
    /* measurement */
    cudaEvent_t ev_begin, ev_end, ev_k0, ev_k1, ev_obs;
    int obs_in_flight; /* the staged depth frame may still be on its way to the device */
    int profile;
    float prof_ms[DBOT_B200_KERNEL_COUNT];
    int64_t prof_launches[DBOT_B200_KERNEL_COUNT];
    int64_t launches;
    int64_t device_bytes;
    int update_since_sync;   /* an update was enqueued through the device-pointer entry points */
    int poisoned;            /* an update failed on the device: the slot tables are half-built, reset() required */
    struct dbot_multi* multi; /* != nullptr: this handle shards its particles over several GPUs (multi_device.cuh) */
    struct dbot_peer_gather* peer; /* != nullptr: one rank of a process-per-GPU group (peer_gather.cuh) */
    dbot_shard_worker* stager[DBOT_STAGERS]; /* helpers that copy pieces of a pageable depth frame into h_obs */
    dbot_stage_job* stage;
    int stagers_wanted; /* DBOT_B200_OPT_STAGING_THREADS, 0 .. DBOT_STAGERS */
    /* DBOT_B200_HOST_TIMING=1: host microseconds spent in the calls of a frame (printed by destroy) */
    int host_timing;
    int64_t ht_obs_calls, ht_ll_calls;
    double ht_us[4]; /* set_observation (pageable) | loglikes: stage + enqueue | loglikes: wait + copy out | spare */
};
static double host_now_us()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#define DBOT_CHECK_USABLE(s, what)                                                                          \
    do                                                                                                      \
    {                                                                                                       \
        if ((s)->poisoned)                                                                                  \
            return fail(-13, what ": an earlier update failed on the device and left the occlusion-slot "   \
                             "tables half-built; call reset() first");                                      \
    } while (0)

static uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

/* multi_device.cuh */
static int multi_create(const dbot_b200_config* cfg, dbot_b200_sensor* s);
static int multi_destroy(dbot_b200_sensor* s);
static int multi_set_observation(dbot_b200_sensor* s, const float* depth_m);
static dbot_b200_sensor* multi_owner(dbot_b200_sensor* s, int32_t slot, int32_t* local);
static int multi_timer_begin(dbot_b200_sensor* s);
static int multi_timer_end(dbot_b200_sensor* s, float* out_ms);
static int multi_n_slots(dbot_b200_sensor* s);
static dbot_b200_sensor* multi_child(dbot_b200_sensor* s, int g);
static int multi_count(dbot_b200_sensor* s);
/* peer_gather.cuh */
static void peer_gather_close(dbot_b200_sensor* s);
static void stop_stagers(dbot_b200_sensor* s);
#define DBOT_SINGLE_ONLY(s, what)                                                                      \
    do                                                                                                 \
    {                                                                                                  \
        if ((s)->multi)                                                                                \
            return fail(-12, what ": device-pointer entry points need a single-device handle "         \
                             "(this one shards over %d GPUs)", multi_count(s));                        \
    } while (0)

/* Back-face culling is applied to an object only when its mesh is PROVEN closed and consistently
 * oriented: after welding vertices with bit-identical positions, every directed edge u->v must occur
 * exactly once and its opposite v->u exactly once; every connected component must enclose volume
 * of the same sign.  Then along any ray from outside the nearest surface is camera-facing, so
 * dropping the triangles that face away leaves the depth image unchanged (up to the rounding that
 * decides between the two triangles meeting at a silhouette edge; quantified in tests/).  dbot's
 * renderer draws every triangle; open, inside-out-in-parts or non-manifold meshes are therefore
 * rendered exactly as before (return 0).  Returns the sign of the snapped screen-space area
 * (x right, y down) of the triangles to keep. */
static int mesh_cull_sign(const float* V, int nv, const int32_t* T, int nt, double fxfy)
{
    if (nv < 4 || nt < 4) return 0;
    struct key3
    {
        uint32_t a, b, c;
        bool operator==(const key3& o) const { return a == o.a && b == o.b && c == o.c; }
    };
    struct hash3
    {
        size_t operator()(const key3& k) const
        {
            uint64_t h = k.a * 0x9E3779B97F4A7C15ull;
            h ^= (h >> 29) + k.b * 0xBF58476D1CE4E5B9ull;
            h ^= (h >> 31) + k.c * 0x94D049BB133111EBull;
            return (size_t)(h ^ (h >> 32));
        }
    };
    std::unordered_map<key3, int32_t, hash3> weld;
    std::vector<int32_t> canon((size_t)nv);
    for (int i = 0; i < nv; ++i)
    {
        key3 k;
        memcpy(&k, V + 3 * (size_t)i, sizeof(k));
        canon[i] = weld.emplace(k, i).first->second;
    }
    std::unordered_map<uint64_t, int32_t> edge_tri; /* directed edge -> triangle */
    edge_tri.reserve((size_t)nt * 3);
    auto ekey = [](int32_t u, int32_t v) { return ((uint64_t)(uint32_t)u << 32) | (uint32_t)v; };
    for (int t = 0; t < nt; ++t)
    {
        const int32_t a = canon[T[3 * t]], b = canon[T[3 * t + 1]], c = canon[T[3 * t + 2]];
        if (a == b || b == c || c == a) return 0;
        const int32_t e[3][2] = {{a, b}, {b, c}, {c, a}};
        for (auto& uv : e)
            if (!edge_tri.emplace(ekey(uv[0], uv[1]), t).second) return 0; /* used twice in one direction */
    }
    std::vector<int32_t> parent((size_t)nt);
    for (int t = 0; t < nt; ++t) parent[t] = t;
    auto find = [&](int32_t i) {
        while (parent[i] != i) i = parent[i] = parent[parent[i]];
        return i;
    };
    for (const auto& kv : edge_tri)
    {
        const int32_t u = (int32_t)(kv.first >> 32), v = (int32_t)(kv.first & 0xffffffffu);
        auto opp = edge_tri.find(ekey(v, u));
        if (opp == edge_tri.end()) return 0; /* boundary edge */
        const int32_t ra = find(kv.second), rb = find(opp->second);
        if (ra != rb) parent[ra] = rb;
    }
    std::vector<double> vol((size_t)nt, 0.0);
    for (int t = 0; t < nt; ++t)
    {
        const float *a = V + 3 * (size_t)T[3 * t], *b = V + 3 * (size_t)T[3 * t + 1], *c = V + 3 * (size_t)T[3 * t + 2];
        vol[find(t)] += (double)a[0] * ((double)b[1] * c[2] - (double)b[2] * c[1]) +
                        (double)a[1] * ((double)b[2] * c[0] - (double)b[0] * c[2]) +
                        (double)a[2] * ((double)b[0] * c[1] - (double)b[1] * c[0]);
    }
    int sgn = 0;
    for (int t = 0; t < nt; ++t)
    {
        if (find(t) != t) continue;
        const int sv = vol[t] > 0.0 ? 1 : (vol[t] < 0.0 ? -1 : 0);
        if (sv == 0 || (sgn != 0 && sv != sgn)) return 0;
        sgn = sv;
    }
    /* outward-oriented triangles facing the camera have negative area where x grows to the right and y
     * downwards (fx fy > 0); a mirrored image flips that */
    return fxfy > 0.0 ? -sgn : sgn;
}

/* split every object mesh into chunks of at most DBOT_CHUNK_MAX_VERTS vertices and
 * DBOT_CHUNK_MAX_TRIS triangles with chunk-local uint16 indices (shared vertices are
 * duplicated; a duplicated vertex transforms to bit-identical coordinates).  Vertices are
 * stored as float4 (read once per particle straight from L2), triangle lists as uint16
 * triples (TMA-staged into shared memory). */
static int build_chunks(const dbot_b200_config* cfg, std::vector<dbot_chunk>& chunks,
                        std::vector<unsigned char>& blob, int* cull_signs)
{
    for (int o = 0; o < cfg->n_objects; ++o)
    {
        const int nv = cfg->n_vertices[o], nt = cfg->n_triangles[o];
        cull_signs[o] = 0;
        if (nv < 0 || nt < 0) return fail(-2, "object %d: negative vertex/triangle count", o);
        const float* V = cfg->vertices[o];
        const int32_t* T = cfg->triangles[o];
        if ((nv > 0 && !V) || (nt > 0 && !T)) return fail(-1, "object %d: null mesh pointer", o);
        for (int i = 0; i < 3 * nt; ++i)
            if (T[i] < 0 || T[i] >= nv) return fail(-6, "object %d: triangle index out of range", o);
        if (cfg->cull_backfaces == 0) cull_signs[o] = mesh_cull_sign(V, nv, T, nt, cfg->K[0] * cfg->K[4]);
        int t = 0;
        while (t < nt)
        {
            std::unordered_map<int32_t, uint16_t> local;
            std::vector<float> cv;
            std::vector<uint16_t> ct;
            while (t < nt && (int)ct.size() / 3 < DBOT_CHUNK_MAX_TRIS)
            {
                int fresh = 0;
                for (int k = 0; k < 3; ++k)
                    if (!local.count(T[3 * t + k])) ++fresh;
                /* worst case all three are new (repeated indices inside one triangle count once) */
                if ((int)local.size() + fresh > DBOT_CHUNK_MAX_VERTS) break;
                for (int k = 0; k < 3; ++k)
                {
                    const int32_t g = T[3 * t + k];
                    auto it = local.find(g);
                    if (it == local.end())
                    {
                        it = local.emplace(g, (uint16_t)local.size()).first;
                        cv.push_back(V[3 * g]);
                        cv.push_back(V[3 * g + 1]);
                        cv.push_back(V[3 * g + 2]);
                        cv.push_back(0.0f); /* float4 per vertex: one 128-bit load */
                    }
                    ct.push_back(it->second);
                }
                ++t;
            }
            dbot_chunk ch;
            memset(&ch, 0, sizeof(ch));
            ch.obj = o;
            ch.cull_sign = cull_signs[o];
            ch.n_verts = (int)cv.size() / 4;
            ch.n_tris = (int)ct.size() / 3;
            ch.vert_bytes = round_up((uint32_t)(cv.size() * sizeof(float)), 16);
            ch.tri_bytes = round_up((uint32_t)(ct.size() * sizeof(uint16_t)), 16);
            ch.vert_off = (uint32_t)blob.size();
            blob.resize(blob.size() + ch.vert_bytes, 0);
            memcpy(blob.data() + ch.vert_off, cv.data(), cv.size() * sizeof(float));
            ch.tri_off = (uint32_t)blob.size();
            blob.resize(blob.size() + ch.tri_bytes, 0);
            memcpy(blob.data() + ch.tri_off, ct.data(), ct.size() * sizeof(uint16_t));
            chunks.push_back(ch);
        }
    }
    if (chunks.empty())
    {
        /* no triangles at all: keep one empty chunk so the kernel has something to walk */
        dbot_chunk ch;
        memset(&ch, 0, sizeof(ch));
        ch.vert_bytes = ch.tri_bytes = 16;
        ch.vert_off = 0;
        ch.tri_off = 16;
        blob.resize(32, 0);
        chunks.push_back(ch);
    }
    return 0;
}

extern "C" int dbot_b200_abi_version(void) { return DBOT_B200_ABI_VERSION; }
extern "C" const char* dbot_b200_last_error(void) { return g_last_error.c_str(); }

extern "C" int dbot_b200_destroy(dbot_b200_sensor* s)
{
    if (!s) return fail(-1, "destroy: null handle");
    if (s->multi)
    {
        multi_destroy(s);
        delete s;
        return 0;
    }
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->host_timing && s->ht_ll_calls > 0)
        fprintf(stderr, "dbot_b200 host timing, device %d: set_observation (pageable) %.1f us x %lld, loglikes stage + "
                        "enqueue %.1f us, wait + copy out %.1f us x %lld\n", s->device,
                s->ht_us[0] / (double)std::max<int64_t>(s->ht_obs_calls, 1), (long long)s->ht_obs_calls,
                s->ht_us[1] / (double)s->ht_ll_calls, s->ht_us[2] / (double)s->ht_ll_calls, (long long)s->ht_ll_calls);
    peer_gather_close(s);
    stop_stagers(s);
    cudaFree(s->d_mesh); cudaFree(s->d_chunks); cudaFree(s->d_obs); cudaFree(s->d_tab);
    cudaFree(s->d_pool); cudaFree(s->d_loc[0]); cudaFree(s->d_loc[1]); cudaFree(s->d_free);
    cudaFree(s->d_free_count); cudaFree(s->d_cnt); cudaFree(s->d_last); cudaFree(s->d_pending);
    cudaFree(s->d_order); cudaFree(s->d_plan); cudaFree(s->d_slot_rect); cudaFree(s->d_poses);
    cudaFree(s->d_result); cudaFree(s->d_depth);
    cudaFreeHost(s->h_poses); cudaFreeHost(s->h_result);
    cudaFreeHost(s->h_obs); cudaFreeHost(s->h_depth);
    if (s->ev_begin) cudaEventDestroy(s->ev_begin);
    if (s->ev_end) cudaEventDestroy(s->ev_end);
    if (s->ev_k0) cudaEventDestroy(s->ev_k0);
    if (s->ev_k1) cudaEventDestroy(s->ev_k1);
    if (s->ev_obs) cudaEventDestroy(s->ev_obs);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return 0;
}

static int create_impl(const dbot_b200_config* cfg, dbot_b200_sensor* s)
{
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(e != cudaSuccess ? (int)e : (int)cudaErrorNoDevice,
                    "no CUDA device: libdbot_b200 has no CPU fallback (%s)", cudaGetErrorString(e));
    const int device_id = (cfg->n_devices == 1 && cfg->device_ids) ? cfg->device_ids[0] : cfg->device_id;
    if (device_id < 0 || device_id >= n_dev)
        return fail(-2, "device_id %d out of range (have %d)", device_id, n_dev);
    s->device = device_id;
    s->stagers_wanted = DBOT_STAGERS;
    {
        const char* t = getenv("DBOT_B200_HOST_TIMING");
        s->host_timing = t && *t == '1';
    }
    CUDA_TRY(cudaSetDevice(s->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, s->device));
    if (prop.major < 10)
        return fail(-8, "device %d is sm_%d%d; this library is built for sm_100a only", s->device,
                    prop.major, prop.minor);
    s->num_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&s->ev_begin));
    CUDA_TRY(cudaEventCreate(&s->ev_end));
    CUDA_TRY(cudaEventCreate(&s->ev_k0));
    CUDA_TRY(cudaEventCreate(&s->ev_k1));
    CUDA_TRY(cudaEventCreateWithFlags(&s->ev_obs, cudaEventDisableTiming));

    /* mesh */
    std::vector<dbot_chunk> chunks;
    std::vector<unsigned char> blob;
    int rc = build_chunks(cfg, chunks, blob, s->cull_signs);
    if (rc) return rc;
    CUDA_TRY(cudaMalloc(&s->d_mesh, blob.size()));
    CUDA_TRY(cudaMemcpy(s->d_mesh, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&s->d_chunks, chunks.size() * sizeof(dbot_chunk)));
    CUDA_TRY(cudaMemcpy(s->d_chunks, chunks.data(), chunks.size() * sizeof(dbot_chunk),
                        cudaMemcpyHostToDevice));
    s->device_bytes += (int64_t)blob.size() + (int64_t)(chunks.size() * sizeof(dbot_chunk));

    dbot_kparams& kp = s->kp;
    memset(&kp, 0, sizeof(kp));
    kp.rows = cfg->rows;
    kp.cols = cfg->cols;
    kp.pitch = (int)round_up((uint32_t)cfg->cols, 8); /* 256-bit rows */
    kp.n_obj = cfg->n_objects;
    kp.n_chunks = (int)chunks.size();
    kp.slot_stride = (size_t)kp.rows * kp.pitch;
    kp.chunks = s->d_chunks;
    kp.mesh = s->d_mesh;
    const double fx = cfg->K[0], sk = cfg->K[1], cx = cfg->K[2], fy = cfg->K[4], cy = cfg->K[5];
    kp.K.fx = (float)fx; kp.K.sk = (float)sk; kp.K.cx = (float)cx; kp.K.fy = (float)fy; kp.K.cy = (float)cy;
    kp.K.k00 = (float)(1.0 / fx);
    kp.K.k01 = (float)(-sk / (fx * fy));
    kp.K.k02 = (float)((sk * cy - cx * fy) / (fx * fy));
    kp.K.k11 = (float)(1.0 / fy);
    kp.K.k12 = (float)(-cy / fy);
    s->lambda = -log(0.5) / cfg->half_life_depth; /* dbot KinectPixelModel: rate from half life */
    s->wD = cfg->tail_weight / cfg->max_depth;
    s->c = cfg->p_occluded_occluded - cfg->p_occluded_visible;
    s->log_c = log(s->c);
    kp.M.lam_rsqrt2 = (float)(s->lambda / 1.4142135623730951);
    kp.M.nlam_log2e = (float)(-s->lambda * 1.4426950408889634);
    kp.M.model_sigma = (float)cfg->model_sigma;
    kp.M.sigma_factor = (float)cfg->sigma_factor;

    /* shared-memory carve-up: triangle list | transformed vertices (float4 + Y) | 1/z tile */
    int max_nv = 1, max_nt = 1;
    for (const dbot_chunk& ch : chunks)
    {
        max_nv = std::max(max_nv, ch.n_verts);
        max_nt = std::max(max_nt, ch.n_tris);
    }
    uint32_t off = 0;
    kp.off_tri = off; off += round_up((uint32_t)max_nt * 6u, 128);
    kp.off_v4 = off;  off += round_up((uint32_t)max_nv * 16u, 128);
    kp.off_XY = off;  off += round_up((uint32_t)max_nv * 8u, 128);
    kp.off_list = off; off += round_up((uint32_t)max_nt * 2u, 128);
    kp.off_wt = off;
    const void* kernels[] = {(const void*)dbot_particle_kernel<DBOT_MODE_EVAL, false>,
                             (const void*)dbot_particle_kernel<DBOT_MODE_EVAL, true>,
                             (const void*)dbot_particle_kernel<DBOT_MODE_UPDATE, false>,
                             (const void*)dbot_particle_kernel<DBOT_MODE_UPDATE, true>,
                             (const void*)dbot_particle_kernel<DBOT_MODE_RENDER, false>,
                             (const void*)dbot_particle_kernel<DBOT_MODE_RENDER, true>};
    /* what the kernels declare as static __shared__ (first-row masks, counters, staged poses, hand-off) comes
     * off the per-block limit; the rest is the dynamic carve-up */
    size_t static_max = 0;
    for (const void* k : kernels)
    {
        cudaFuncAttributes fa;
        CUDA_TRY(cudaFuncGetAttributes(&fa, k));
        static_max = std::max(static_max, fa.sharedSizeBytes);
    }
    const int smem_max = ((int)prop.sharedMemPerBlockOptin - (int)static_max - 64) / 128 * 128;
    if ((int)off + 4 * (int)round_up((uint32_t)kp.cols, 4) > smem_max)
        return fail(-9, "image too wide for the shared-memory depth tile (cols=%d)", kp.cols);
    s->smem_bytes = smem_max;
    kp.wcap = ((uint32_t)(smem_max - (int)off) / 16u) * 4u; /* pixels, multiple of 4 */
    for (const void* k : kernels)
        CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, s->smem_bytes));
    s->plan_smem_max = (int)prop.sharedMemPerBlockOptin - 4096;
    CUDA_TRY(cudaFuncSetAttribute((const void*)dbot_plan_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  s->plan_smem_max));
    s->grid_max = s->num_sms; /* one persistent CTA per SM (launch bounds: 1 CTA/SM) */

    /* buffers */
    const size_t P = (size_t)kp.rows * kp.cols;
    const size_t np = (size_t)cfg->max_particles;
    const size_t state_bytes = np * kp.slot_stride * sizeof(float);
    CUDA_TRY(cudaMalloc(&s->d_obs, P * sizeof(float)));
    CUDA_TRY(cudaMalloc(&s->d_tab, kp.slot_stride * sizeof(float4)));
    CUDA_TRY(cudaMalloc(&s->d_pool, 2 * state_bytes));
    CUDA_TRY(cudaMalloc(&s->d_loc[0], 2 * np * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&s->d_loc[1], 2 * np * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&s->d_free, 2 * np * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&s->d_free_count, sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&s->d_cnt, 2 * np * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&s->d_last, 2 * np * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&s->d_pending, 2 * np * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&s->d_order, np * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&s->d_slot_rect, 2 * np * sizeof(int4)));
    CUDA_TRY(cudaMalloc(&s->d_plan, np * sizeof(int4)));
    /* per-call traffic of the host entry points, packed so that each direction is ONE copy: poses and parent
     * indices of a call go up back to back; the error flag sits in front of the log-weights and comes back with
     * them */
    CUDA_TRY(cudaMalloc(&s->d_poses, np * 12 * kp.n_obj * sizeof(float) + np * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&s->d_result, (np + 1) * sizeof(double)));
    CUDA_TRY(cudaMemset(s->d_result, 0, sizeof(double)));
    s->d_err = (int*)s->d_result;
    s->d_out = s->d_result + 1;
    CUDA_TRY(cudaMalloc(&s->d_depth, P * sizeof(float)));
    s->device_bytes += (int64_t)(2 * P * sizeof(float) + kp.slot_stride * sizeof(float4) +
                                 2 * state_bytes + np * (12 * kp.n_obj * 4 + 4 + 8 + 8 * 4 + 16) + 8);
    CUDA_TRY(cudaHostAlloc(&s->h_poses, np * 12 * kp.n_obj * sizeof(float) + np * sizeof(int32_t), cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc(&s->h_result, (np + 1) * sizeof(double), cudaHostAllocDefault));
    s->h_err = (int*)s->h_result;
    s->h_out = s->h_result + 1;
    *s->h_err = 0;
    CUDA_TRY(cudaHostAlloc(&s->h_obs, P * sizeof(float), cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc(&s->h_depth, P * sizeof(float), cudaHostAllocDefault));
    /* until the first set_observation every pixel is "no data" */
    CUDA_TRY(cudaMemsetAsync(s->d_obs, 0xff, P * sizeof(float), s->stream));
    CUDA_TRY(cudaMemsetAsync(s->d_tab, 0xff, kp.slot_stride * sizeof(float4), s->stream));
    return dbot_b200_reset(s);
}

extern "C" int dbot_b200_create(const dbot_b200_config* cfg, dbot_b200_sensor** out)
{
    if (!cfg || !out) return fail(-1, "create: null argument");
    *out = nullptr;
    if (cfg->rows <= 0 || cfg->cols <= 0) return fail(-2, "create: bad resolution %dx%d", cfg->cols, cfg->rows);
    if (cfg->n_objects <= 0 || cfg->n_objects > DBOT_B200_MAX_OBJECTS)
        return fail(-2, "create: n_objects=%d not in [1,%d]", cfg->n_objects, DBOT_B200_MAX_OBJECTS);
    if (cfg->max_particles <= 0) return fail(-2, "create: max_particles must be positive");
    if (!cfg->vertices || !cfg->n_vertices || !cfg->triangles || !cfg->n_triangles)
        return fail(-1, "create: null mesh arrays");
    if (cfg->K[3] != 0.0 || cfg->K[6] != 0.0 || cfg->K[7] != 0.0 || cfg->K[8] != 1.0 ||
        cfg->K[0] == 0.0 || cfg->K[4] == 0.0)
        return fail(-3, "create: camera matrix must be [[fx s cx][0 fy cy][0 0 1]]");
    const double c = cfg->p_occluded_occluded - cfg->p_occluded_visible;
    if (!(c > 0.0) || c == 1.0) return fail(-4, "create: need 0 < p_occluded_occluded - p_occluded_visible < 1");
    if (!(cfg->half_life_depth > 0.0) || !(cfg->max_depth > 0.0) || !(cfg->model_sigma > 0.0))
        return fail(-4, "create: half_life_depth, max_depth and model_sigma must be positive");
    if (cfg->occlusion_mode != 0) return fail(-2, "create: occlusion_mode %d is reserved", cfg->occlusion_mode);
    if (cfg->cull_backfaces != 0 && cfg->cull_backfaces != 1)
        return fail(-2, "create: cull_backfaces must be 0 (where proven safe) or 1 (never)");
    /* the span records pack pixel coordinates into 16 bits */
    if (cfg->rows > 32768 || cfg->cols > 32768) return fail(-2, "create: resolution %dx%d above 32768", cfg->cols, cfg->rows);
    dbot_b200_sensor* s = new dbot_b200_sensor();
    memset(s, 0, sizeof(*s));
    s->cfg = *cfg;
    s->cfg.vertices = nullptr; s->cfg.triangles = nullptr;
    s->cfg.n_vertices = nullptr; s->cfg.n_triangles = nullptr;
    if (cfg->n_devices < 0) return fail(-2, "create: n_devices must not be negative");
    int rc = cfg->n_devices > 1 ? multi_create(cfg, s) : create_impl(cfg, s);
    if (rc)
    {
        std::string keep = g_last_error;
        dbot_b200_destroy(s);
        g_last_error = keep;
        return rc;
    }
    *out = s;
    return 0;
}

static void prof_begin(dbot_b200_sensor* s)
{
    if (s->profile) cudaEventRecord(s->ev_k0, s->stream);
}
static void prof_end(dbot_b200_sensor* s, int which)
{
    s->launches++;
    if (!s->profile) return;
    cudaEventRecord(s->ev_k1, s->stream);
    cudaEventSynchronize(s->ev_k1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s->ev_k0, s->ev_k1);
    s->prof_ms[which] += ms;
    s->prof_launches[which]++;
}

/* RbSensor::reset: one logical slot (physical slot 0) at initial_occlusion_prob,
 * every other physical slot free */
__global__ void dbot_reset_slots_kernel(int32_t* loc, int32_t* free_stack, int32_t* free_count, int4* slot_rect,
                                        int n_phys, int rows, int row8)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0)
    {
        loc[0] = 0;
        *free_count = n_phys - 1;
    }
    if (i < n_phys - 1) free_stack[i] = n_phys - 1 - i; /* top of the stack = slot 1 */
    if (i < n_phys) slot_rect[i] = make_int4(rows, -1, row8, -1); /* every slot holds the background everywhere */
}

__global__ void dbot_free_all_slots_kernel(int32_t* free_stack, int32_t* free_count, int n_phys)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *free_count = n_phys;
    if (i < n_phys) free_stack[i] = n_phys - 1 - i;
}

static int multi_reset(dbot_b200_sensor* s);

extern "C" int dbot_b200_reset(dbot_b200_sensor* s)
{
    if (!s) return fail(-1, "reset: null handle");
    s->poisoned = 0;
    if (s->multi) return multi_reset(s);
    CUDA_TRY(cudaSetDevice(s->device));
    s->cur = 0;
    s->n_slots = 1;
    s->t_obs = 0.0;
    s->t_update = 0.0;
    s->A = 1.0;
    s->B = 0.0;
    const int n_phys = 2 * s->cfg.max_particles;
    /* EVERY physical slot starts as the background image: the copies of an update rely on a recycled slot
     * holding the background outside its recorded rows */
    s->background = s->cfg.initial_occlusion_prob;
    dbot_fill_kernel<<<s->num_sms * 8, 256, 0, s->stream>>>(s->d_pool, (size_t)n_phys * s->kp.slot_stride,
                                                            (float)s->background);
    dbot_reset_slots_kernel<<<(n_phys + 255) / 256, 256, 0, s->stream>>>(s->d_loc[0], s->d_free, s->d_free_count,
                                                                         s->d_slot_rect, n_phys, s->kp.rows,
                                                                         s->kp.pitch >> 3);
    s->launches += 2;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

static int launch_observation(dbot_b200_sensor* s, const float* d_depth)
{
    const int total = (int)s->kp.slot_stride;
    prof_begin(s);
    dbot_observation_kernel<<<(total + 255) / 256, 256, 0, s->stream>>>(
        d_depth, s->d_tab, s->kp.rows, s->kp.cols, s->kp.pitch, s->cfg.tail_weight,
        s->cfg.model_sigma, s->cfg.sigma_factor, s->lambda, s->wD);
    CUDA_TRY(cudaGetLastError());
    prof_end(s, DBOT_B200_KERNEL_OBSERVATION);
    s->t_obs += s->cfg.delta_time;
    return 0;
}

/* page-locked host memory can be handed to the copy engine as it is */
/* the helpers that stage pageable depth frames are started with the first such frame */
static void ensure_stagers(dbot_b200_sensor* s)
{
    if (s->stage) return;
    s->stage = new dbot_stage_job();
    for (int k = 0; k < s->stagers_wanted; ++k)
    {
        dbot_shard_worker* w = new dbot_shard_worker();
        w->spin = 4000; /* frames are a third of a millisecond apart at best: sleep between them */
        w->th = std::thread([w] { w->loop(-1); });
        s->stager[k] = w;
    }
}
static void stop_stagers(dbot_b200_sensor* s)
{
    for (int k = 0; k < DBOT_STAGERS; ++k)
        if (s->stager[k])
        {
            s->stager[k]->stop();
            delete s->stager[k];
            s->stager[k] = nullptr;
        }
    delete s->stage;
    s->stage = nullptr;
}
/* arm the staging job for one frame and wake the helpers; the caller then claims pieces itself */
static dbot_stage_job* begin_staging(dbot_b200_sensor* s, float* dst, const float* src, size_t count)
{
    ensure_stagers(s);
    dbot_stage_job* job = s->stage;
    job->arm(dst, src, count * sizeof(float), count >= 65536 ? DBOT_STAGE_PIECES : 1);
    if (job->np > 1)
        for (int k = 0; k < DBOT_STAGERS; ++k)
            if (s->stager[k]) s->stager[k]->post([job] { job->help(); });
    return job;
}

extern "C" int dbot_b200_set_option(dbot_b200_sensor* s, int32_t option, int64_t value)
{
    if (!s) return fail(-1, "set_option: null handle");
    if (option == DBOT_B200_OPT_STAGING_THREADS)
    {
        if (value < 0 || value > DBOT_STAGERS) return fail(-2, "set_option: staging threads must be 0 .. %d", DBOT_STAGERS);
        stop_stagers(s); /* restarted with the next pageable frame */
        s->stagers_wanted = (int)value;
        return 0;
    }
    return fail(-2, "set_option: unknown option %d", option);
}

static bool is_pinned_host(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

/* RbSensor::set_observation */
extern "C" int dbot_b200_set_observation(dbot_b200_sensor* s, const float* depth_m)
{
    if (!s || !depth_m) return fail(-1, "set_observation: null argument");
    DBOT_CHECK_USABLE(s, "set_observation");
    if (s->multi) return multi_set_observation(s, depth_m);
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t bytes = (size_t)s->kp.rows * s->kp.cols * sizeof(float);
    if (is_pinned_host(depth_m))
    {
        /* the caller's buffer is page-locked: copy straight from it, and be done with it on return -- with the
         * copy, that is; the table kernel is left running like everything else on the stream */
        CUDA_TRY(cudaMemcpyAsync(s->d_obs, depth_m, bytes, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaEventRecord(s->ev_obs, s->stream));
        int rc = launch_observation(s, s->d_obs);
        if (rc) return rc;
        CUDA_TRY(cudaEventSynchronize(s->ev_obs));
        s->obs_in_flight = 0;
        return 0;
    }
    /* pageable: stage in the handle's own pinned buffer; the copy and the table kernel are left
     * running -- every later call on this handle is ordered behind them on the same stream, and the
     * staging buffer is not touched again before they are done */
    const double ht0 = s->host_timing ? host_now_us() : 0.0;
    if (s->obs_in_flight) CUDA_TRY(cudaEventSynchronize(s->ev_obs));
    /* in pieces that this thread and the staging helpers claim one by one; every piece goes to the device as soon
     * as it and all pieces before it are staged, so the copy engine works while the rest is still being staged */
    dbot_stage_job* job = begin_staging(s, s->h_obs, depth_m, (size_t)s->kp.rows * s->kp.cols);
    cudaError_t copy_err = cudaSuccess;
    for (int issued = 0; issued < job->np;)
    {
        const bool mine = job->claim_one();
        while (issued < job->np && job->is_done(issued))
        {
            const size_t a = job->cut(issued), b = job->cut(issued + 1);
            if (copy_err == cudaSuccess)
                copy_err = cudaMemcpyAsync((char*)s->d_obs + a, (const char*)s->h_obs + a, b - a, cudaMemcpyHostToDevice,
                                           s->stream);
            ++issued;
        }
        if (!mine && issued < job->np) __builtin_ia32_pause(); /* a helper is still copying a piece it claimed */
    }
    CUDA_TRY(copy_err);
    CUDA_TRY(cudaEventRecord(s->ev_obs, s->stream));
    s->obs_in_flight = 1;
    const int orc = launch_observation(s, s->d_obs);
    if (s->host_timing)
    {
        s->ht_us[0] += host_now_us() - ht0;
        s->ht_obs_calls++;
    }
    return orc;
}

extern "C" int dbot_b200_set_observation_device(dbot_b200_sensor* s, const float* d_depth_m)
{
    if (!s || !d_depth_m) return fail(-1, "set_observation_device: null argument");
    DBOT_CHECK_USABLE(s, "set_observation_device");
    DBOT_SINGLE_ONLY(s, "set_observation_device");
    CUDA_TRY(cudaSetDevice(s->device));
    return launch_observation(s, d_depth_m);
}

/* occlusion propagation over dt as an affine map o -> a o + b (dbot OcclusionModel) */
static void propagation_affine(const dbot_b200_sensor* s, double dt, double* a, double* b)
{
    const double ct = exp(dt * s->log_c);
    const double k = (1.0 - s->cfg.p_occluded_occluded) * (ct - 1.0) / (s->c - 1.0);
    *a = ct;
    *b = 1.0 - ct - k;
}
/* stored value -> occlusion probability as of the current observation time */
static void state_affine_now(const dbot_b200_sensor* s, double* A, double* B)
{
    double a, b;
    propagation_affine(s, s->t_obs - s->t_update, &a, &b);
    *A = a * s->A;
    *B = a * s->B + b;
}

#define DBOT_REBASE_BELOW 8.673617379884035e-19 /* 2^-60 */

static int launch_particles(dbot_b200_sensor* s, const float* d_poses, int32_t* d_indices, int n,
                            int update, double* d_out)
{
    double A, B;
    state_affine_now(s, &A, &B);
    if (update && A < DBOT_REBASE_BELOW)
    {
        /* fold the accumulated map into the stored values before it leaves float range */
        /* every physical slot, free ones included: they must keep holding the (new) background value */
        dbot_rebase_kernel<<<s->num_sms * 8, 256, 0, s->stream>>>(
            s->d_pool, (size_t)2 * s->cfg.max_particles * s->kp.slot_stride, (float)s->A, (float)s->B);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        s->background = (double)fmaf((float)s->A, (float)s->background, (float)s->B);
        s->A = 1.0;
        s->B = 0.0;
        state_affine_now(s, &A, &B);
    }
    dbot_kparams kp = s->kp;
    kp.n = n;
    kp.n_slots = s->n_slots;
    kp.tab = s->d_tab;
    kp.pool = s->d_pool;
    kp.loc = s->d_loc[s->cur];
    kp.plan = s->d_plan;
    kp.order = s->d_order;
    kp.pending = s->d_pending;
    kp.slot_rect = s->d_slot_rect;
    kp.background = (float)s->background;
    kp.poses = d_poses;
    kp.indices = d_indices;
    kp.out = d_out;
    kp.err_flag = s->d_err;
    kp.M.prop_a = (float)A;
    kp.M.prop_b = (float)B;
    kp.post_b = (float)B;
    kp.post_inv_a = (float)(1.0 / A);
    if (update)
    {
        prof_begin(s);
        const size_t table_bytes = (size_t)s->n_slots * 2 * sizeof(int32_t);
        const int in_smem = table_bytes <= (size_t)s->plan_smem_max;
        dbot_plan_update_kernel<<<1, DBOT_PLAN_THREADS, in_smem ? table_bytes : 0, s->stream>>>(
            d_indices, n, s->n_slots, s->d_loc[s->cur], s->d_loc[s->cur ^ 1], s->d_free, s->d_free_count,
            s->d_cnt, s->d_last, s->d_pending, s->d_order, s->d_plan, s->d_err, in_smem);
        CUDA_TRY(cudaGetLastError());
        prof_end(s, DBOT_B200_KERNEL_PLAN);
    }
    const int grid = std::min(n, s->grid_max);
    prof_begin(s);
    if (update)
    {
        /* adopting children spin on their siblings' copies: every CTA must be resident */
        void* args[] = {(void*)&kp};
        const void* kern = kp.n_chunks > 1 ? (const void*)dbot_particle_kernel<DBOT_MODE_UPDATE, true>
                                           : (const void*)dbot_particle_kernel<DBOT_MODE_UPDATE, false>;
        CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(grid),
                                             dim3(DBOT_CT + DBOT_ST), args, (size_t)s->smem_bytes, s->stream));
    }
    else
    {
        if (kp.n_chunks > 1)
            dbot_particle_kernel<DBOT_MODE_EVAL, true><<<grid, DBOT_CT, s->smem_bytes, s->stream>>>(kp);
        else
            dbot_particle_kernel<DBOT_MODE_EVAL, false><<<grid, DBOT_CT, s->smem_bytes, s->stream>>>(kp);
    }
    CUDA_TRY(cudaGetLastError());
    prof_end(s, update ? DBOT_B200_KERNEL_UPDATE : DBOT_B200_KERNEL_LOGLIKES);
    if (update)
    {
        s->cur ^= 1;
        s->n_slots = n;
        s->t_update = s->t_obs;
        s->A = A;
        s->B = B;
    }
    return 0;
}

/* device error flag -> error code; a failed update leaves the handle unusable until reset() */
static int device_error(dbot_b200_sensor* s, const char* what, int was_update)
{
    const int code = *s->h_err;
    if (!code) return 0;
    cudaMemsetAsync(s->d_err, 0, sizeof(int), s->stream);
    if (was_update) s->poisoned = 1;
    if (code == 4)
        return fail(-16, "%s: a peer rank did not deliver its log-weights within the time limit (did every rank make "
                         "the same gathering call?)", what);
    if (code == 3) return fail(-11, "%s: an adopting particle timed out waiting for its siblings' copies", what);
    if (code == 2) return fail(-10, "%s: the occlusion-slot pool is exhausted (too many imported slots)", what);
    return fail(-7, "%s: a parent index was not a valid occlusion slot", what);
}

/* an update with no particles: every slot is given back, the handle's clock and lazy-propagation state
 * advance exactly as with particles (a shard that receives no particle of a call must stay in step with
 * the other shards) */
static int empty_update(dbot_b200_sensor* s)
{
    double A, B;
    state_affine_now(s, &A, &B);
    const int n_phys = 2 * s->cfg.max_particles;
    dbot_free_all_slots_kernel<<<(n_phys + 255) / 256, 256, 0, s->stream>>>(s->d_free, s->d_free_count, n_phys);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    s->n_slots = 0;
    s->t_update = s->t_obs;
    s->A = A;
    s->B = B;
    return 0;
}

/* poses and parent indices of one call: staged back to back in page-locked memory, one copy up */
static int stage_call_inputs(dbot_b200_sensor* s, const float* poses, const int32_t* indices, int32_t n,
                             int32_t** d_indices)
{
    const size_t pose_floats = (size_t)n * 12 * s->kp.n_obj;
    memcpy(s->h_poses, poses, pose_floats * sizeof(float));
    memcpy(s->h_poses + pose_floats, indices, (size_t)n * sizeof(int32_t));
    CUDA_TRY(cudaMemcpyAsync(s->d_poses, s->h_poses, pose_floats * sizeof(float) + (size_t)n * sizeof(int32_t),
                             cudaMemcpyHostToDevice, s->stream));
    *d_indices = (int32_t*)(s->d_poses + pose_floats);
    return 0;
}

/* first half of RbSensor::loglikes: stage the host arrays, enqueue copies and kernels, no wait */
static int loglikes_enqueue(dbot_b200_sensor* s, const float* poses, const int32_t* indices, int32_t n,
                            int32_t update)
{
    CUDA_TRY(cudaSetDevice(s->device));
    int32_t* d_indices = nullptr;
    int rc = stage_call_inputs(s, poses, indices, n, &d_indices);
    if (rc) return rc;
    return launch_particles(s, s->d_poses, d_indices, n, update, s->d_out);
}
/* second half: log-weights and the error flag back, wait, report */
static int loglikes_finish(dbot_b200_sensor* s, int32_t n, int32_t update, double* out_loglikes)
{
    CUDA_TRY(cudaSetDevice(s->device));
    /* the error flag and, behind it, the log-weights: one copy */
    const size_t back = (n > 0 && out_loglikes) ? (size_t)(n + 1) * sizeof(double) : sizeof(int);
    CUDA_TRY(cudaMemcpyAsync(s->h_result, s->d_result, back, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    int rc = device_error(s, "loglikes", update);
    if (rc) return rc;
    if (n > 0 && out_loglikes) memcpy(out_loglikes, s->h_out, (size_t)n * sizeof(double));
    return 0;
}

/* particle-sharded evaluation: the update of a call in which this shard received no particle */
extern "C" int dbot_b200_update_empty(dbot_b200_sensor* s)
{
    if (!s) return fail(-1, "update_empty: null handle");
    DBOT_CHECK_USABLE(s, "update_empty");
    DBOT_SINGLE_ONLY(s, "update_empty");
    CUDA_TRY(cudaSetDevice(s->device));
    return empty_update(s);
}

static int multi_loglikes(dbot_b200_sensor* s, const float* poses, int32_t* indices, int32_t n, int32_t update,
                          double* out_loglikes);

/* RbSensor::loglikes */
extern "C" int dbot_b200_loglikes(dbot_b200_sensor* s, const float* poses, int32_t* indices,
                                  int32_t n, int32_t update, double* out_loglikes)
{
    if (!s || n < 0) return fail(-1, "loglikes: null handle or negative n");
    if (n > 0 && (!poses || !indices || !out_loglikes)) return fail(-1, "loglikes: null argument");
    DBOT_CHECK_USABLE(s, "loglikes");
    if (n > s->cfg.max_particles)
        return fail(-2, "loglikes: n=%d exceeds max_particles=%d", n, s->cfg.max_particles);
    if (s->multi) return multi_loglikes(s, poses, indices, n, update, out_loglikes);
    for (int i = 0; i < n; ++i)
        if (indices[i] < 0 || indices[i] >= s->n_slots)
            return fail(-7, "loglikes: indices[%d]=%d is not an occlusion slot (have %d)", i,
                        indices[i], s->n_slots);
    if (n == 0) return 0; /* as dbot's model: a call without particles changes nothing, update or not */
    const double ht0 = s->host_timing ? host_now_us() : 0.0;
    int rc = loglikes_enqueue(s, poses, indices, n, update);
    if (rc) return rc;
    const double ht1 = s->host_timing ? host_now_us() : 0.0;
    rc = loglikes_finish(s, n, update, out_loglikes);
    if (rc) return rc;
    if (s->host_timing)
    {
        s->ht_us[1] += ht1 - ht0;
        s->ht_us[2] += host_now_us() - ht1;
        s->ht_ll_calls++;
    }
    if (update)
        for (int i = 0; i < n; ++i) indices[i] = i;
    return 0;
}

extern "C" int dbot_b200_loglikes_device(dbot_b200_sensor* s, const float* d_poses,
                                         int32_t* d_indices, int32_t n, int32_t update,
                                         double* d_out_loglikes)
{
    if (!s || !d_poses || !d_indices || !d_out_loglikes) return fail(-1, "loglikes_device: null argument");
    if (n < 0 || n > s->cfg.max_particles)
        return fail(-2, "loglikes_device: n=%d exceeds max_particles=%d", n, s->cfg.max_particles);
    DBOT_CHECK_USABLE(s, "loglikes_device");
    DBOT_SINGLE_ONLY(s, "loglikes_device");
    if (n == 0) return 0;
    CUDA_TRY(cudaSetDevice(s->device));
    if (update) s->update_since_sync = 1;
    return launch_particles(s, d_poses, d_indices, n, update, d_out_loglikes);
}

extern "C" void* dbot_b200_stream(dbot_b200_sensor* s) { return (s && !s->multi) ? (void*)s->stream : nullptr; }

extern "C" int dbot_b200_synchronize(dbot_b200_sensor* s)
{
    if (!s) return fail(-1, "synchronize: null handle");
    DBOT_SINGLE_ONLY(s, "synchronize");
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaMemcpyAsync(s->h_err, s->d_err, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    const int was_update = s->update_since_sync;
    s->update_since_sync = 0;
    return device_error(s, "loglikes_device", was_update);
}

/* RigidBodyRenderer::Render for one pose set */
extern "C" int dbot_b200_render(dbot_b200_sensor* s, const float* pose, float* out_depth)
{
    if (!s || !pose || !out_depth) return fail(-1, "render: null argument");
    if (s->multi) s = multi_child(s, 0);
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t P = (size_t)s->kp.rows * s->kp.cols;
    const size_t pose_bytes = (size_t)12 * s->kp.n_obj * sizeof(float);
    float* d_pose = nullptr;
    CUDA_TRY(cudaMalloc(&d_pose, pose_bytes));
    cudaError_t e = cudaMemcpyAsync(d_pose, pose, pose_bytes, cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess)
    {
        dbot_fill_kernel<<<64, 256, 0, s->stream>>>(s->d_depth, P, INFINITY);
        s->launches++;
        dbot_kparams kp = s->kp;
        kp.n = 1;
        kp.n_slots = 1;
        kp.poses = d_pose;
        kp.depth_out = s->d_depth;
        kp.err_flag = s->d_err;
        prof_begin(s);
        if (kp.n_chunks > 1)
            dbot_particle_kernel<DBOT_MODE_RENDER, true><<<1, DBOT_CT, s->smem_bytes, s->stream>>>(kp);
        else
            dbot_particle_kernel<DBOT_MODE_RENDER, false><<<1, DBOT_CT, s->smem_bytes, s->stream>>>(kp);
        e = cudaGetLastError();
        prof_end(s, DBOT_B200_KERNEL_RENDER);
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(s->h_depth, s->d_depth, P * sizeof(float), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d_pose);
    if (e != cudaSuccess) return fail((int)e, "render failed: %s", cudaGetErrorString(e));
    memcpy(out_depth, s->h_depth, P * sizeof(float));
    return 0;
}

/* ---- state read-back ---- */
extern "C" int dbot_b200_n_slots(dbot_b200_sensor* s) { return !s ? -1 : (s->multi ? multi_n_slots(s) : s->n_slots); }
extern "C" int dbot_b200_cull_sign(dbot_b200_sensor* s, int32_t object)
{
    return (s && object >= 0 && object < s->cfg.n_objects) ? s->cull_signs[object] : 0;
}
extern "C" double dbot_b200_observation_time(dbot_b200_sensor* s)
{
    return !s ? NAN : (s->multi ? multi_child(s, 0)->t_obs : s->t_obs);
}

/* ---- occlusion-slot migration between handles (cross-GPU resampling) ---- */
__global__ void dbot_export_slot_kernel(const float* __restrict__ pool, const int32_t* __restrict__ loc,
                                        int slot, size_t slot_stride, float* __restrict__ dst)
{
    const float4* src = reinterpret_cast<const float4*>(pool + (size_t)loc[slot] * slot_stride);
    float4* out = reinterpret_cast<float4*>(dst);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < slot_stride / 4;
         i += (size_t)gridDim.x * blockDim.x)
        out[i] = src[i];
}
__global__ void dbot_import_pop_kernel(int32_t* loc, int32_t* free_stack, int32_t* free_count, int slot,
                                       int* err_flag, int4* slot_rect, int rows, int row8)
{
    const int f = *free_count;
    if (f <= 0)
    {
        atomicExch(err_flag, 2);
        loc[slot] = loc[0];
        return;
    }
    loc[slot] = free_stack[f - 1];
    slot_rect[free_stack[f - 1]] = make_int4(0, rows - 1, 0, row8 - 1); /* an imported image may differ from the background anywhere */
    *free_count = f - 1;
}
__global__ void dbot_import_copy_kernel(float* __restrict__ pool, const int32_t* __restrict__ loc, int slot,
                                        size_t slot_stride, const float* __restrict__ src)
{
    float4* out = reinterpret_cast<float4*>(pool + (size_t)loc[slot] * slot_stride);
    const float4* in = reinterpret_cast<const float4*>(src);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < slot_stride / 4;
         i += (size_t)gridDim.x * blockDim.x)
        out[i] = in[i];
}

extern "C" int64_t dbot_b200_slot_floats(dbot_b200_sensor* s) { return s ? (int64_t)s->kp.slot_stride : -1; }

extern "C" int dbot_b200_export_slot_device(dbot_b200_sensor* s, int32_t slot, float* d_dst)
{
    if (!s || !d_dst) return fail(-1, "export_slot: null argument");
    DBOT_SINGLE_ONLY(s, "export_slot");
    if (slot < 0 || slot >= s->n_slots) return fail(-7, "export_slot: slot %d of %d", slot, s->n_slots);
    CUDA_TRY(cudaSetDevice(s->device));
    dbot_export_slot_kernel<<<64, 256, 0, s->stream>>>(s->d_pool, s->d_loc[s->cur], slot, s->kp.slot_stride, d_dst);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int dbot_b200_import_slot_device(dbot_b200_sensor* s, const float* d_src, int32_t* out_slot)
{
    if (!s || !d_src || !out_slot) return fail(-1, "import_slot: null argument");
    DBOT_SINGLE_ONLY(s, "import_slot");
    /* live + imported slots stay within half of the pool, so that an update can always find a fresh
     * slot for every copying child (every imported slot is some child's parent and is adopted) */
    if (s->n_slots >= 2 * s->cfg.max_particles - 1)
        return fail(-2, "import_slot: %d slots are live, the pool holds %d", s->n_slots, 2 * s->cfg.max_particles);
    CUDA_TRY(cudaSetDevice(s->device));
    const int slot = s->n_slots;
    dbot_import_pop_kernel<<<1, 1, 0, s->stream>>>(s->d_loc[s->cur], s->d_free, s->d_free_count, slot, s->d_err,
                                                   s->d_slot_rect, s->kp.rows, s->kp.pitch >> 3);
    dbot_import_copy_kernel<<<64, 256, 0, s->stream>>>(s->d_pool, s->d_loc[s->cur], slot, s->kp.slot_stride, d_src);
    s->launches += 2;
    CUDA_TRY(cudaGetLastError());
    s->n_slots = slot + 1;
    *out_slot = slot;
    return 0;
}

/* rectangle of a slot that may differ from the background: { first row, last row, first column, last column },
 * first > last: none */
extern "C" int dbot_b200_slot_rect(dbot_b200_sensor* s, int32_t slot, int32_t* out_rect)
{
    if (!s || !out_rect) return fail(-1, "slot_rect: null argument");
    if (s->multi)
    {
        if (slot < 0 || slot >= multi_n_slots(s)) return fail(-7, "slot_rect: slot %d of %d", slot, multi_n_slots(s));
        int32_t local = 0;
        dbot_b200_sensor* owner = multi_owner(s, slot, &local);
        return dbot_b200_slot_rect(owner, local, out_rect);
    }
    if (slot < 0 || slot >= s->n_slots) return fail(-7, "slot_rect: slot %d of %d", slot, s->n_slots);
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    int32_t phys = 0;
    int4 r;
    CUDA_TRY(cudaMemcpy(&phys, s->d_loc[s->cur] + slot, sizeof(int32_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&r, s->d_slot_rect + phys, sizeof(int4), cudaMemcpyDeviceToHost));
    out_rect[0] = r.x;
    out_rect[1] = r.y;
    /* columns are kept in 256-bit words (8 floats): the copy's unit */
    out_rect[2] = r.z <= r.w ? std::min(r.z * 8, s->kp.cols) : s->kp.cols;
    out_rect[3] = r.z <= r.w ? std::min(r.w * 8 + 7, s->kp.cols - 1) : -1;
    return 0;
}
/* first / last image row of that rectangle */
extern "C" int dbot_b200_slot_rows(dbot_b200_sensor* s, int32_t slot, int32_t* out_first, int32_t* out_last)
{
    if (!out_first || !out_last) return fail(-1, "slot_rows: null argument");
    int32_t r[4];
    const int rc = dbot_b200_slot_rect(s, slot, r);
    if (rc) return rc;
    *out_first = r[0];
    *out_last = r[1];
    return 0;
}

extern "C" int dbot_b200_free_slot_count(dbot_b200_sensor* s)
{
    if (!s) return -1;
    if (s->multi)
    {
        int total = 0;
        for (int g = 0; g < multi_count(s); ++g)
        {
            const int f = dbot_b200_free_slot_count(multi_child(s, g));
            if (f < 0) return -1;
            total += f;
        }
        return total;
    }
    if (cudaSetDevice(s->device) != cudaSuccess || cudaStreamSynchronize(s->stream) != cudaSuccess) return -1;
    int32_t f = -1;
    if (cudaMemcpy(&f, s->d_free_count, sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return f;
}

extern "C" int dbot_b200_get_occlusion(dbot_b200_sensor* s, int32_t slot, float* out_occ)
{
    if (!s || !out_occ) return fail(-1, "get_occlusion: null argument");
    if (s->multi)
    {
        if (slot < 0 || slot >= multi_n_slots(s)) return fail(-7, "get_occlusion: slot %d of %d", slot, multi_n_slots(s));
        int32_t local = 0;
        dbot_b200_sensor* owner = multi_owner(s, slot, &local);
        return dbot_b200_get_occlusion(owner, local, out_occ);
    }
    if (slot < 0 || slot >= s->n_slots) return fail(-7, "get_occlusion: slot %d of %d", slot, s->n_slots);
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    const int rows = s->kp.rows, cols = s->kp.cols, pitch = s->kp.pitch;
    std::vector<float> tmp((size_t)rows * cols);
    int32_t phys = 0;
    CUDA_TRY(cudaMemcpy(&phys, s->d_loc[s->cur] + slot, sizeof(int32_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy2D(tmp.data(), (size_t)cols * sizeof(float),
                          s->d_pool + (size_t)phys * s->kp.slot_stride,
                          (size_t)pitch * sizeof(float), (size_t)cols * sizeof(float), rows,
                          cudaMemcpyDeviceToHost));
    double A, B;
    state_affine_now(s, &A, &B);
    for (size_t i = 0; i < tmp.size(); ++i) out_occ[i] = fmaf((float)A, tmp[i], (float)B);
    return 0;
}

/* ---- measurement hooks ---- */
extern "C" int dbot_b200_timer_begin(dbot_b200_sensor* s)
{
    if (!s) return fail(-1, "timer_begin: null handle");
    if (s->multi) return multi_timer_begin(s);
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaEventRecord(s->ev_begin, s->stream));
    return 0;
}
extern "C" int dbot_b200_timer_end(dbot_b200_sensor* s, float* out_ms)
{
    if (!s || !out_ms) return fail(-1, "timer_end: null argument");
    if (s->multi) return multi_timer_end(s, out_ms);
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaEventRecord(s->ev_end, s->stream));
    CUDA_TRY(cudaEventSynchronize(s->ev_end));
    CUDA_TRY(cudaEventElapsedTime(out_ms, s->ev_begin, s->ev_end));
    return 0;
}
extern "C" int dbot_b200_profile_enable(dbot_b200_sensor* s, int32_t on)
{
    if (!s) return fail(-1, "profile_enable: null handle");
    if (s->multi)
    {
        for (int g = 0; g < multi_count(s); ++g) dbot_b200_profile_enable(multi_child(s, g), on);
        return 0;
    }
    s->profile = on ? 1 : 0;
    memset(s->prof_ms, 0, sizeof(s->prof_ms));
    memset(s->prof_launches, 0, sizeof(s->prof_launches));
    return 0;
}
extern "C" int dbot_b200_profile_read(dbot_b200_sensor* s, float* out_ms, int64_t* out_launches)
{
    if (!s || !out_ms || !out_launches) return fail(-1, "profile_read: null argument");
    if (s->multi) /* the slowest shard's time per kernel kind, launches summed over the shards */
    {
        float ms[DBOT_B200_KERNEL_COUNT];
        int64_t cnt[DBOT_B200_KERNEL_COUNT];
        for (int k = 0; k < DBOT_B200_KERNEL_COUNT; ++k) { out_ms[k] = 0.f; out_launches[k] = 0; }
        for (int g = 0; g < multi_count(s); ++g)
        {
            dbot_b200_profile_read(multi_child(s, g), ms, cnt);
            for (int k = 0; k < DBOT_B200_KERNEL_COUNT; ++k)
            {
                out_ms[k] = std::max(out_ms[k], ms[k]);
                out_launches[k] += cnt[k];
            }
        }
        return 0;
    }
    memcpy(out_ms, s->prof_ms, sizeof(s->prof_ms));
    memcpy(out_launches, s->prof_launches, sizeof(s->prof_launches));
    memset(s->prof_ms, 0, sizeof(s->prof_ms));
    memset(s->prof_launches, 0, sizeof(s->prof_launches));
    return 0;
}
extern "C" int64_t dbot_b200_launch_count(dbot_b200_sensor* s)
{
    if (!s) return -1;
    if (!s->multi) return s->launches;
    int64_t t = 0;
    for (int g = 0; g < multi_count(s); ++g) t += multi_child(s, g)->launches;
    return t;
}
extern "C" int64_t dbot_b200_device_bytes(dbot_b200_sensor* s)
{
    if (!s) return -1;
    if (!s->multi) return s->device_bytes;
    int64_t t = 0;
    for (int g = 0; g < multi_count(s); ++g) t += multi_child(s, g)->device_bytes;
    return t;
}

#include "dbot_filter.cuh"
#include "multi_device.cuh"
#include "peer_gather.cuh"
static int multi_n_slots(dbot_b200_sensor* s) { return s->multi->n_slots; }
static dbot_b200_sensor* multi_child(dbot_b200_sensor* s, int g) { return s->multi->child[g]; }
static int multi_count(dbot_b200_sensor* s) { return s->multi->G; }
