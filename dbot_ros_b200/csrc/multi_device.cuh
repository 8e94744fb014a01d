/*
 * multi_device.cuh -- one handle, several GPUs (BASELINE config 4).
 * Included at the end of dbot_b200.cu (same translation unit).
 *
 * dbot_ros builds sensor and tracker in ONE process (ref:source/dbot_ros/tracker/particle_tracker_node.cpp:
 * 201-218), so the multi-GPU evaluator must live behind the same handle RbSensorBuilder::build() returns:
 * create() with n_devices > 1 makes one shard (a complete single-device sensor: own stream, own occlusion
 * pool) per GPU.  Particles are independent, so a loglikes call is cut into contiguous slices -- shard g
 * evaluates particles [g c, (g+1) c), c = ceil(n / G), and owns the occlusion slots of exactly those
 * particles after an update; frame, mesh and camera are replicated.  The one exchange of the path is the
 * all-gather of the log-weights, done in one of two ways (cfg.gather_mode):
 *   0  ncclAllGather (ncclCommInitAll, one ncclGroupStart/End per call) over NVLink / NVSwitch into every
 *      device's gathered vector.  NCCL is bound at run time (dlopen of the libnccl.so.2 the process
 *      already holds, else the system one), so the library has no link-time dependency on it.
 *   1  fused: the particle kernel's epilogue stores each log-weight straight into every device's
 *      gathered vector through peer memory (st.global on NVLink-mapped pointers) -- no collective launch
 *      at all; the transfer overlaps the next particle's rasterisation.
 * Either way one D2H copy of the gathered vector from device 0 ends the call.
 *
 * A child whose parent slot lives on another shard (after a resampling step)
 * gets the parent's stored image first: the needing shard pops a free slot and copies the image out of
 * the owner's pool over NVLink (peer loads in dbot_import_peer_kernel).  Stored values are lazily
 * propagated (value = A r + B); every shard sees the same sequence of observations and updates -- shards
 * that receive no particle of an update advance through empty_update() -- so A and B are identical
 * everywhere and images can move verbatim.
 */
#pragma once
#include <dlfcn.h>

#include <chrono>
#include <thread>

/* the handful of NCCL declarations this file needs (bound with dlsym; nccl.h is not required to build) */
typedef struct ncclComm* dbot_nccl_comm;
typedef int (*pfn_ncclCommInitAll)(dbot_nccl_comm*, int, const int*);
typedef int (*pfn_ncclCommDestroy)(dbot_nccl_comm);
typedef int (*pfn_ncclGroupStart)(void);
typedef int (*pfn_ncclGroupEnd)(void);
typedef int (*pfn_ncclAllGather)(const void*, void*, size_t, int, dbot_nccl_comm, cudaStream_t);
typedef const char* (*pfn_ncclGetErrorString)(int);
#define DBOT_NCCL_FLOAT64 8 /* ncclFloat64 / ncclDouble */

struct dbot_mig_desc
{
    const float* src_pool;  /* the owner's slot pool (peer memory) */
    const int32_t* src_loc; /* the owner's logical -> physical table */
    int32_t src_slot;       /* logical slot on the owner */
    int32_t pad;
};

struct dbot_multi
{
    int G;
    int dev[DBOT_B200_MAX_DEVICES];
    dbot_b200_sensor* child[DBOT_B200_MAX_DEVICES];
    int gather_mode;
    int cnt_max;      /* particles per shard at max_particles */
    int n_slots;      /* logical slots over all shards */
    int fresh;        /* no update since reset(): every shard holds its own copy of slot 0 */
    int prev_cnt;     /* shard width of the last update: logical slot p lives on shard p / prev_cnt */
    /* gathered log-weights: [G * cnt_max] doubles on every device, shard g's slice at g * cnt */
    double* d_all[DBOT_B200_MAX_DEVICES];
    double* h_all;
    float* h_obs;
    int obs_in_flight;
    cudaEvent_t ev_done[DBOT_B200_MAX_DEVICES], ev_begin[DBOT_B200_MAX_DEVICES], ev_end[DBOT_B200_MAX_DEVICES];
    cudaEvent_t ev_frame; /* the depth frame has landed in the first GPU's memory */
    /* NCCL */
    void* nccl_lib;
    pfn_ncclCommInitAll CommInitAll;
    pfn_ncclCommDestroy CommDestroy;
    pfn_ncclGroupStart GroupStart;
    pfn_ncclGroupEnd GroupEnd;
    pfn_ncclAllGather AllGather;
    pfn_ncclGetErrorString GetErrorString;
    dbot_nccl_comm comms[DBOT_B200_MAX_DEVICES];
    int have_comms;
    /* parents imported since the last update, per shard: global slot -> local slot */
    std::unordered_map<int32_t, int32_t>* imported[DBOT_B200_MAX_DEVICES];
    dbot_mig_desc* d_mig[DBOT_B200_MAX_DEVICES];
    dbot_mig_desc* h_mig[DBOT_B200_MAX_DEVICES];
    int64_t migrated;
    std::vector<int32_t>* local_idx;
    dbot_shard_worker* worker[DBOT_B200_MAX_DEVICES]; /* [0] unused: shard 0 runs on the caller's thread */
    /* DBOT_B200_HOST_TIMING=1: where the host time of a frame goes (printed by destroy) */
    int timing;
    int64_t timed_calls;
    double phase_us[8];
    cudaEvent_t tev[DBOT_B200_MAX_DEVICES][4]; /* frame in / table done / particle launch / particle kernel done */
    double gpu_us[DBOT_B200_MAX_DEVICES][3];
};
static const char* const dbot_multi_phase_names[8] = {"obs: stage frame", "obs: enqueue on all shards", "loglikes: check + localise parents",
                                                      "loglikes: enqueue on all shards", "loglikes: gather set-up",
                                                      "loglikes: wait for all shards", "loglikes: copy out", ""};
struct dbot_phase_clock
{
    dbot_multi* m;
    std::chrono::steady_clock::time_point t;
    explicit dbot_phase_clock(dbot_multi* mm) : m(mm), t(std::chrono::steady_clock::now()) {}
    void mark(int phase)
    {
        if (!m->timing) return;
        const auto now = std::chrono::steady_clock::now();
        m->phase_us[phase] += std::chrono::duration<double, std::micro>(now - t).count();
        t = now;
    }
};

/* run fn(g) for every shard: shard 0 on this thread, the others on their workers; returns when all are done */
template <class F>
static void multi_for_each_shard(dbot_multi* m, F fn)
{
    for (int g = 1; g < m->G; ++g) m->worker[g]->post([fn, g] { fn(g); });
    fn(0);
    for (int g = 1; g < m->G; ++g) m->worker[g]->wait();
}

/* pop n_mig free physical slots for the logical slots first .. first + n_mig - 1 */
__global__ void dbot_import_pop_many_kernel(int32_t* loc, const int32_t* free_stack, int32_t* free_count, int first,
                                            int n_mig, int* err_flag, int4* slot_rect, int rows, int row8)
{
    const int f = *free_count;
    if (f < n_mig)
    {
        if (threadIdx.x == 0) atomicExch(err_flag, 2);
        for (int k = threadIdx.x; k < n_mig; k += blockDim.x) loc[first + k] = loc[0];
        return;
    }
    for (int k = threadIdx.x; k < n_mig; k += blockDim.x)
    {
        loc[first + k] = free_stack[f - 1 - k];
        slot_rect[free_stack[f - 1 - k]] = make_int4(0, rows - 1, 0, row8 - 1); /* imported: may differ from the background anywhere */
    }
    __syncthreads();
    if (threadIdx.x == 0) *free_count = f - n_mig;
}
/* blockIdx.y = migration k: the owner's stored image -> this shard's new slot, read over NVLink */
__global__ void dbot_import_peer_kernel(float* __restrict__ pool, const int32_t* __restrict__ loc, int first,
                                        size_t slot_stride, const dbot_mig_desc* __restrict__ desc)
{
    const dbot_mig_desc d = desc[blockIdx.y];
    const float4* src = reinterpret_cast<const float4*>(d.src_pool + (size_t)d.src_loc[d.src_slot] * slot_stride);
    float4* dst = reinterpret_cast<float4*>(pool + (size_t)loc[first + blockIdx.y] * slot_stride);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < slot_stride / 4;
         i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

static void multi_bounds(int n, int G, int g, int* lo, int* hi, int* cnt)
{
    const int c = (n + G - 1) / G;
    *cnt = c;
    *lo = std::min(n, g * c);
    *hi = std::min(n, (g + 1) * c);
}

static int multi_sync_all(dbot_multi* m)
{
    for (int g = 0; g < m->G; ++g)
    {
        CUDA_TRY(cudaSetDevice(m->dev[g]));
        CUDA_TRY(cudaStreamSynchronize(m->child[g]->stream));
    }
    return 0;
}

static int multi_destroy(dbot_b200_sensor* s)
{
    dbot_multi* m = s->multi;
    if (!m) return 0;
    if (m->timing && m->timed_calls > 0)
    {
        fprintf(stderr, "dbot_b200 host timing, %d GPUs, %lld loglikes calls, microseconds per call:\n", m->G,
                (long long)m->timed_calls);
        for (int k = 0; k < 7; ++k)
            fprintf(stderr, "  %-40s %9.1f\n", dbot_multi_phase_names[k], m->phase_us[k] / (double)m->timed_calls);
        for (int g = 0; g < m->G; ++g)
            fprintf(stderr, "  GPU %d stream: frame in + table %8.1f   idle until the particle launch %8.1f   poses + plan + "
                            "particle kernel %8.1f\n", m->dev[g], m->gpu_us[g][0] / (double)m->timed_calls,
                    m->gpu_us[g][1] / (double)m->timed_calls, m->gpu_us[g][2] / (double)m->timed_calls);
    }
    stop_stagers(s);
    for (int g = 1; g < m->G; ++g)
        if (m->worker[g])
        {
            m->worker[g]->stop();
            delete m->worker[g];
        }
    for (int g = 0; g < m->G; ++g)
    {
        if (!m->child[g]) continue;
        cudaSetDevice(m->dev[g]);
        cudaStreamSynchronize(m->child[g]->stream);
        if (m->have_comms && m->comms[g]) m->CommDestroy(m->comms[g]);
        cudaFree(m->d_all[g]);
        cudaFree(m->d_mig[g]);
        cudaFreeHost(m->h_mig[g]);
        if (m->ev_done[g]) cudaEventDestroy(m->ev_done[g]);
        if (m->ev_begin[g]) cudaEventDestroy(m->ev_begin[g]);
        if (m->ev_end[g]) cudaEventDestroy(m->ev_end[g]);
        delete m->imported[g];
        dbot_b200_destroy(m->child[g]);
    }
    if (m->ev_frame) cudaEventDestroy(m->ev_frame);
    cudaFreeHost(m->h_all);
    cudaFreeHost(m->h_obs);
    delete m->local_idx;
    /* the NCCL library stays loaded: other users in the process (e.g. torch) share it */
    delete m;
    s->multi = nullptr;
    return 0;
}

static int multi_load_nccl(dbot_multi* m)
{
    const char* env = getenv("DBOT_B200_NCCL_LIB");
    void* h = nullptr;
    if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL); /* the copy the process already holds */
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h)
        return fail(-14, "n_devices > 1 with gather_mode 0 needs NCCL: libnccl.so.2 not found (%s); set "
                         "DBOT_B200_NCCL_LIB or use gather_mode 1 (fused peer stores)", dlerror());
    m->nccl_lib = h;
    m->CommInitAll = (pfn_ncclCommInitAll)dlsym(h, "ncclCommInitAll");
    m->CommDestroy = (pfn_ncclCommDestroy)dlsym(h, "ncclCommDestroy");
    m->GroupStart = (pfn_ncclGroupStart)dlsym(h, "ncclGroupStart");
    m->GroupEnd = (pfn_ncclGroupEnd)dlsym(h, "ncclGroupEnd");
    m->AllGather = (pfn_ncclAllGather)dlsym(h, "ncclAllGather");
    m->GetErrorString = (pfn_ncclGetErrorString)dlsym(h, "ncclGetErrorString");
    if (!m->CommInitAll || !m->CommDestroy || !m->GroupStart || !m->GroupEnd || !m->AllGather || !m->GetErrorString)
        return fail(-14, "libnccl.so.2 lacks an entry point this library needs");
    return 0;
}

static int multi_create(const dbot_b200_config* cfg, dbot_b200_sensor* s)
{
    const int G = cfg->n_devices;
    if (G > DBOT_B200_MAX_DEVICES) return fail(-2, "create: n_devices=%d above %d", G, DBOT_B200_MAX_DEVICES);
    if (cfg->gather_mode != 0 && cfg->gather_mode != 1)
        return fail(-2, "create: gather_mode must be 0 (ncclAllGather) or 1 (fused peer stores)");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(e != cudaSuccess ? (int)e : (int)cudaErrorNoDevice,
                    "no CUDA device: libdbot_b200 has no CPU fallback (%s)", cudaGetErrorString(e));
    dbot_multi* m = new dbot_multi();
    memset((void*)m, 0, sizeof(*m));
    s->multi = m;
    m->G = G;
    m->gather_mode = cfg->gather_mode;
    m->cnt_max = (cfg->max_particles + G - 1) / G;
    m->local_idx = new std::vector<int32_t>();
    s->stagers_wanted = DBOT_STAGERS;
    {
        const char* t = getenv("DBOT_B200_HOST_TIMING");
        m->timing = t && *t == '1';
    }
    for (int g = 0; g < G; ++g)
    {
        m->dev[g] = cfg->device_ids ? cfg->device_ids[g] : g;
        if (m->dev[g] < 0 || m->dev[g] >= n_dev)
            return fail(-2, "create: device_ids[%d]=%d out of range (have %d)", g, m->dev[g], n_dev);
        /* shards may share a GPU only without NCCL (which wants one rank per device): useful to exercise the
         * sharding, migration and fused-gather logic on a one-GPU machine */
        for (int h = 0; h < g; ++h)
            if (m->dev[h] == m->dev[g] && m->gather_mode == 0)
                return fail(-2, "create: device %d listed twice (allowed only with gather_mode 1)", m->dev[g]);
    }
    for (int g = 0; g < G; ++g)
    {
        dbot_b200_config c = *cfg;
        c.n_devices = 1;
        c.device_ids = nullptr;
        c.device_id = m->dev[g];
        c.max_particles = m->cnt_max;
        int rc = dbot_b200_create(&c, &m->child[g]);
        if (rc) return rc;
        m->imported[g] = new std::unordered_map<int32_t, int32_t>();
        CUDA_TRY(cudaSetDevice(m->dev[g]));
        CUDA_TRY(cudaMalloc(&m->d_all[g], (size_t)G * m->cnt_max * sizeof(double)));
        CUDA_TRY(cudaMalloc(&m->d_mig[g], (size_t)m->cnt_max * sizeof(dbot_mig_desc)));
        CUDA_TRY(cudaHostAlloc(&m->h_mig[g], (size_t)m->cnt_max * sizeof(dbot_mig_desc), cudaHostAllocDefault));
        CUDA_TRY(cudaEventCreateWithFlags(&m->ev_done[g], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreate(&m->ev_begin[g]));
        CUDA_TRY(cudaEventCreate(&m->ev_end[g]));
        if (m->timing)
            for (int e = 0; e < 4; ++e) CUDA_TRY(cudaEventCreate(&m->tev[g][e]));
    }
    /* every shard reads the others' pools (slot migration) and, in fused mode, writes their gathered vectors */
    for (int g = 0; g < G; ++g)
        for (int h = 0; h < G; ++h)
        {
            if (m->dev[g] == m->dev[h]) continue;
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, m->dev[g], m->dev[h]));
            if (!can) return fail(-15, "create: device %d cannot access device %d (no NVLink / P2P)", m->dev[g], m->dev[h]);
            CUDA_TRY(cudaSetDevice(m->dev[g]));
            e = cudaDeviceEnablePeerAccess(m->dev[h], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled)
                cudaGetLastError();
            else if (e != cudaSuccess)
                return fail((int)e, "cudaDeviceEnablePeerAccess(%d -> %d): %s", m->dev[g], m->dev[h], cudaGetErrorString(e));
        }
    CUDA_TRY(cudaSetDevice(m->dev[0]));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_frame, cudaEventDisableTiming));
    CUDA_TRY(cudaHostAlloc(&m->h_all, (size_t)G * m->cnt_max * sizeof(double), cudaHostAllocPortable));
    CUDA_TRY(cudaHostAlloc(&m->h_obs, (size_t)cfg->rows * cfg->cols * sizeof(float), cudaHostAllocPortable));
    if (m->gather_mode == 0)
    {
        int rc = multi_load_nccl(m);
        if (rc) return rc;
        const int nrc = m->CommInitAll(m->comms, G, m->dev);
        if (nrc != 0) return fail(-14, "ncclCommInitAll failed: %s", m->GetErrorString(nrc));
        m->have_comms = 1;
    }
    for (int g = 1; g < G; ++g)
    {
        m->worker[g] = new dbot_shard_worker();
        dbot_shard_worker* w = m->worker[g];
        const int device = m->dev[g];
        w->th = std::thread([w, device] { w->loop(device); });
    }
    m->n_slots = 1;
    m->prev_cnt = 1;
    m->fresh = 1;
    s->device = m->dev[0];
    s->kp = m->child[0]->kp;
    for (int o = 0; o < cfg->n_objects; ++o) s->cull_signs[o] = m->child[0]->cull_signs[o];
    return 0;
}

static int multi_reset(dbot_b200_sensor* s)
{
    dbot_multi* m = s->multi;
    for (int g = 0; g < m->G; ++g)
    {
        int rc = dbot_b200_reset(m->child[g]);
        if (rc) return rc;
        m->imported[g]->clear();
    }
    m->n_slots = 1; /* every shard holds its own copy of slot 0 */
    m->prev_cnt = 1;
    m->fresh = 1;
    return 0;
}

static int multi_set_observation(dbot_b200_sensor* s, const float* depth_m)
{
    dbot_multi* m = s->multi;
    if (m->obs_in_flight)
    {
        /* the staging buffer and the first GPU's frame buffer may still be read */
        int rc = multi_sync_all(m);
        if (rc) return rc;
    }
    dbot_phase_clock clk(m);
    /* The frame crosses PCIe ONCE, to the first GPU -- staged in pieces that this thread and the staging helpers
     * claim, every piece on its way as soon as it and the pieces before it are staged.  The other GPUs do not
     * upload their own copy (measured on an 8-GPU box: eight uploads of the same page-locked buffer take 0.5 ms,
     * they run one after the other): their table kernels read the raw frame out of the first GPU's memory over
     * NVLink (peer loads, 1.2 MB each) once its copy has landed. */
    dbot_b200_sensor* c0 = m->child[0];
    CUDA_TRY(cudaSetDevice(m->dev[0]));
    if (m->timing)
        for (int g = 0; g < m->G; ++g)
        {
            CUDA_TRY(cudaSetDevice(m->dev[g]));
            CUDA_TRY(cudaEventRecord(m->tev[g][0], m->child[g]->stream));
        }
    CUDA_TRY(cudaSetDevice(m->dev[0]));
    {
        dbot_stage_job* job = begin_staging(s, m->h_obs, depth_m, (size_t)s->kp.rows * s->kp.cols);
        cudaError_t copy_err = cudaSuccess;
        for (int issued = 0; issued < job->np;)
        {
            const bool mine = job->claim_one();
            while (issued < job->np && job->is_done(issued))
            {
                const size_t a = job->cut(issued), b = job->cut(issued + 1);
                if (copy_err == cudaSuccess)
                    copy_err = cudaMemcpyAsync((char*)c0->d_obs + a, (const char*)m->h_obs + a, b - a,
                                               cudaMemcpyHostToDevice, c0->stream);
                ++issued;
            }
            if (!mine && issued < job->np) __builtin_ia32_pause(); /* a helper is still copying a piece it claimed */
        }
        CUDA_TRY(copy_err);
    }
    CUDA_TRY(cudaEventRecord(m->ev_frame, c0->stream));
    clk.mark(0);
    int shard_rc[DBOT_B200_MAX_DEVICES] = {0};
    std::string shard_err[DBOT_B200_MAX_DEVICES];
    multi_for_each_shard(m, [&](int g) {
        dbot_b200_sensor* c = m->child[g];
        auto body = [&]() -> int {
            CUDA_TRY(cudaSetDevice(m->dev[g]));
            if (g > 0) CUDA_TRY(cudaStreamWaitEvent(c->stream, m->ev_frame, 0));
            const int orc = launch_observation(c, c0->d_obs);
            if (m->timing) CUDA_TRY(cudaEventRecord(m->tev[g][1], c->stream));
            return orc;
        };
        shard_rc[g] = body();
        if (shard_rc[g]) shard_err[g] = g_last_error; /* thread-local: carried back to the caller's thread */
    });
    clk.mark(1);
    m->obs_in_flight = 1;
    for (int g = 0; g < m->G; ++g)
        if (shard_rc[g])
        {
            g_last_error = shard_err[g];
            return shard_rc[g];
        }
    return 0;
}

/* bring every parent that lives on another shard than its child to the child's shard; fills
 * m->local_idx with each particle's parent slot in its own shard's numbering */
static int multi_localise_parents(dbot_b200_sensor* s, const int32_t* indices, int n)
{
    dbot_multi* m = s->multi;
    const int G = m->G;
    m->local_idx->resize((size_t)n);
    int32_t* local = m->local_idx->data();
    int any_migration = 0;
    for (int g = 0; g < G; ++g)
    {
        int lo, hi, cnt;
        multi_bounds(n, G, g, &lo, &hi, &cnt);
        dbot_b200_sensor* c = m->child[g];
        int n_mig = 0;
        for (int i = lo; i < hi; ++i)
        {
            const int p = indices[i];
            if (m->fresh) /* after reset every shard holds slot 0 itself */
            {
                local[i] = 0;
                continue;
            }
            const int owner = p / m->prev_cnt;
            if (owner == g)
            {
                local[i] = p - owner * m->prev_cnt;
                continue;
            }
            auto it = m->imported[g]->find(p);
            if (it == m->imported[g]->end())
            {
                if (c->n_slots + n_mig >= 2 * c->cfg.max_particles - 1 || n_mig >= m->cnt_max)
                    return fail(-10, "loglikes: shard %d would have to import more parent slots than its pool holds", g);
                dbot_b200_sensor* o = m->child[owner];
                dbot_mig_desc d;
                d.src_pool = o->d_pool;
                d.src_loc = o->d_loc[o->cur];
                d.src_slot = p - owner * m->prev_cnt;
                d.pad = 0;
                m->h_mig[g][n_mig] = d;
                it = m->imported[g]->emplace(p, c->n_slots + n_mig).first;
                ++n_mig;
            }
            local[i] = it->second;
        }
        if (n_mig > 0)
        {
            CUDA_TRY(cudaSetDevice(m->dev[g]));
            CUDA_TRY(cudaMemcpyAsync(m->d_mig[g], m->h_mig[g], (size_t)n_mig * sizeof(dbot_mig_desc),
                                     cudaMemcpyHostToDevice, c->stream));
            dbot_import_pop_many_kernel<<<1, 256, 0, c->stream>>>(c->d_loc[c->cur], c->d_free, c->d_free_count,
                                                                  c->n_slots, n_mig, c->d_err, c->d_slot_rect, c->kp.rows,
                                                                  c->kp.pitch >> 3);
            dbot_import_peer_kernel<<<dim3(16, n_mig), 256, 0, c->stream>>>(c->d_pool, c->d_loc[c->cur], c->n_slots,
                                                                           c->kp.slot_stride, m->d_mig[g]);
            c->launches += 2;
            CUDA_TRY(cudaGetLastError());
            c->n_slots += n_mig;
            m->migrated += n_mig;
            any_migration = 1;
        }
    }
    /* an update may write into parent slots in place (slot adoption): every image must have been read
     * out of its owner's pool before any shard starts its particle kernel */
    if (any_migration) return multi_sync_all(m);
    return 0;
}

static int multi_loglikes(dbot_b200_sensor* s, const float* poses, int32_t* indices, int32_t n, int32_t update,
                          double* out_loglikes)
{
    dbot_multi* m = s->multi;
    const int G = m->G;
    dbot_phase_clock clk(m);
    for (int i = 0; i < n; ++i)
        if (indices[i] < 0 || indices[i] >= m->n_slots)
            return fail(-7, "loglikes: indices[%d]=%d is not an occlusion slot (have %d)", i, indices[i], m->n_slots);
    if (n == 0) return 0;
    int rc = multi_localise_parents(s, indices, n);
    if (rc) return rc;
    clk.mark(2);
    const int32_t* local = m->local_idx->data();
    const size_t pose_stride = (size_t)12 * s->kp.n_obj;
    int cnt = (n + G - 1) / G;
    /* every shard's thread stages that shard's slice and enqueues its copies and kernels */
    int shard_rc[DBOT_B200_MAX_DEVICES] = {0};
    std::string shard_err[DBOT_B200_MAX_DEVICES];
    auto enqueue_shard = [&](int g) {
        int lo, hi, c_unused;
        multi_bounds(n, G, g, &lo, &hi, &c_unused);
        dbot_b200_sensor* c = m->child[g];
        auto body = [&]() -> int {
            CUDA_TRY(cudaSetDevice(m->dev[g]));
            if (hi == lo) return update ? empty_update(c) : 0;
            /* the shard's kernel writes its log-weights to its own d_out and, fused mode, into every device's
             * gathered vector at g * cnt + i */
            c->kp.n_peers = 0;
            if (m->gather_mode == 1)
            {
                c->kp.n_peers = G;
                c->kp.out_offset = g * cnt;
                for (int h = 0; h < G; ++h) c->kp.peer_out[h] = m->d_all[h];
            }
            if (m->timing) CUDA_TRY(cudaEventRecord(m->tev[g][2], c->stream));
            const int erc = loglikes_enqueue(c, poses + (size_t)lo * pose_stride, local + lo, hi - lo, update);
            if (m->timing) CUDA_TRY(cudaEventRecord(m->tev[g][3], c->stream));
            return erc;
        };
        shard_rc[g] = body();
        if (shard_rc[g]) shard_err[g] = g_last_error; /* thread-local: carried back to the caller's thread */
    };
    multi_for_each_shard(m, enqueue_shard);
    clk.mark(3);
    for (int g = 0; g < G; ++g)
        if (shard_rc[g])
        {
            g_last_error = shard_err[g];
            if (update) s->poisoned = 1; /* other shards may already have committed this update */
            return shard_rc[g];
        }
    if (m->gather_mode == 0)
    {
        int nrc = m->GroupStart();
        for (int g = 0; g < G && nrc == 0; ++g)
            nrc = m->AllGather(m->child[g]->d_out, m->d_all[g], (size_t)cnt, DBOT_NCCL_FLOAT64, m->comms[g],
                               m->child[g]->stream);
        const int erc = m->GroupEnd();
        if (nrc == 0) nrc = erc;
        if (nrc != 0) return fail(-14, "ncclAllGather failed: %s", m->GetErrorString(nrc));
    }
    else
    {
        /* device 0's copy is complete once every shard's kernel has finished */
        for (int g = 1; g < G; ++g)
        {
            CUDA_TRY(cudaSetDevice(m->dev[g]));
            CUDA_TRY(cudaEventRecord(m->ev_done[g], m->child[g]->stream));
        }
        CUDA_TRY(cudaSetDevice(m->dev[0]));
        for (int g = 1; g < G; ++g) CUDA_TRY(cudaStreamWaitEvent(m->child[0]->stream, m->ev_done[g], 0));
    }
    CUDA_TRY(cudaSetDevice(m->dev[0]));
    CUDA_TRY(cudaMemcpyAsync(m->h_all, m->d_all[0], (size_t)G * cnt * sizeof(double), cudaMemcpyDeviceToHost,
                             m->child[0]->stream));
    clk.mark(4);
    int first_error = 0;
    std::string first_text;
    multi_for_each_shard(m, [&](int g) { /* error flags back, wait: every shard on its own thread */
        shard_rc[g] = loglikes_finish(m->child[g], 0, update, nullptr);
        if (shard_rc[g]) shard_err[g] = g_last_error;
    });
    clk.mark(5);
    if (m->timing && m->obs_in_flight)
        for (int g = 0; g < G; ++g)
            for (int e = 0; e < 3; ++e)
            {
                float ms = 0.f;
                cudaSetDevice(m->dev[g]);
                if (cudaEventElapsedTime(&ms, m->tev[g][e], m->tev[g][e + 1]) == cudaSuccess) m->gpu_us[g][e] += 1e3 * ms;
                else cudaGetLastError();
            }
    for (int g = 0; g < G; ++g)
        if (shard_rc[g] && !first_error)
        {
            first_error = shard_rc[g];
            first_text = shard_err[g];
        }
    m->obs_in_flight = 0;
    if (first_error)
    {
        if (update) s->poisoned = 1;
        g_last_error = first_text;
        return first_error;
    }
    for (int g = 0; g < G; ++g)
    {
        int lo, hi;
        multi_bounds(n, G, g, &lo, &hi, &cnt);
        if (hi > lo) memcpy(out_loglikes + lo, m->h_all + (size_t)g * cnt, (size_t)(hi - lo) * sizeof(double));
    }
    if (update)
    {
        for (int i = 0; i < n; ++i) indices[i] = i;
        m->n_slots = n;
        m->prev_cnt = cnt;
        m->fresh = 0;
        for (int g = 0; g < G; ++g) m->imported[g]->clear();
    }
    clk.mark(6);
    m->timed_calls += m->timing;
    return 0;
}

static dbot_b200_sensor* multi_owner(dbot_b200_sensor* s, int32_t slot, int32_t* local)
{
    dbot_multi* m = s->multi;
    if (m->fresh)
    {
        *local = 0;
        return m->child[0];
    }
    const int owner = slot / m->prev_cnt;
    *local = slot - owner * m->prev_cnt;
    return m->child[owner];
}

static int multi_timer_begin(dbot_b200_sensor* s)
{
    dbot_multi* m = s->multi;
    int rc = multi_sync_all(m);
    if (rc) return rc;
    for (int g = 0; g < m->G; ++g)
    {
        CUDA_TRY(cudaSetDevice(m->dev[g]));
        CUDA_TRY(cudaEventRecord(m->ev_begin[g], m->child[g]->stream));
    }
    return 0;
}
/* the slowest device's time between the two marks */
static int multi_timer_end(dbot_b200_sensor* s, float* out_ms)
{
    dbot_multi* m = s->multi;
    float worst = 0.f;
    for (int g = 0; g < m->G; ++g)
    {
        CUDA_TRY(cudaSetDevice(m->dev[g]));
        CUDA_TRY(cudaEventRecord(m->ev_end[g], m->child[g]->stream));
    }
    for (int g = 0; g < m->G; ++g)
    {
        float ms = 0.f;
        CUDA_TRY(cudaSetDevice(m->dev[g]));
        CUDA_TRY(cudaEventSynchronize(m->ev_end[g]));
        CUDA_TRY(cudaEventElapsedTime(&ms, m->ev_begin[g], m->ev_end[g]));
        worst = std::max(worst, ms);
    }
    *out_ms = worst;
    return 0;
}

extern "C" int dbot_b200_device_count(dbot_b200_sensor* s) { return !s ? -1 : (s->multi ? s->multi->G : 1); }
extern "C" int64_t dbot_b200_migrated_slots(dbot_b200_sensor* s) { return !s ? -1 : (s->multi ? s->multi->migrated : 0); }
