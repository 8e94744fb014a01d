/*
 * peer_gather.cuh -- the all-gather of log-weights for ONE PROCESS PER GPU (torchrun, dbot_ros_b200/sharded.py),
 * fused into the particle kernel over NVLink peer memory.  Included at the end of dbot_b200.cu.
 *
 * multi_device.cuh does the same inside one process (gather_mode 1), where every GPU's memory is reachable
 * through cudaDeviceEnablePeerAccess.  Across processes the gathered vectors are shared as CUDA IPC memory:
 *   - every rank allocates one block: two gathered vectors [world x width_max] doubles (double-buffered by the
 *     parity of the call number) and one arrival counter per peer; exports it (cudaIpcGetMemHandle), the ranks
 *     exchange the 64-byte handles through whatever they have (torch.distributed.all_gather_object in
 *     sharded.py) and map each other's blocks (cudaIpcOpenMemHandle);
 *   - a gathering loglikes call launches the particle kernel with peer_out[g] = rank g's gathered vector: the
 *     kernel's epilogue stores each log-weight at rank * width + i of EVERY rank's vector (st.global over
 *     NVLink / NVSwitch, overlapped with the next particle's rasterisation) -- no collective launch;
 *   - one warp then tells every peer "my slice of call #seq is in your memory" (st.release.sys of seq into the
 *     peer's arrival counter for this rank, ordered behind the particle kernel by the stream) and waits until
 *     every peer has said the same (ld.acquire.sys on its own counters).  After that kernel the local gathered
 *     vector is complete; it is read back with one D2H copy.
 * Two buffers are enough: a rank can start call k+1 (which writes buffer (k+1)&1 everywhere) only after every
 * peer has signalled call k, and a peer signals call k+1 only after -- in stream order -- it has read call k's
 * buffer, so nobody writes buffer k&1 again (call k+2) before everybody has read it.
 * The wait is bounded (about 10 s): a peer that never arrives becomes error code -16, not a hung GPU.
 */
#pragma once

struct dbot_peer_gather
{
    int world, rank, width_max;
    unsigned char* d_block;                           /* this rank's exported allocation */
    unsigned char* peer_block[DBOT_B200_MAX_DEVICES]; /* every rank's block as mapped here ([rank] = d_block) */
    size_t vec_bytes;                                 /* one gathered vector */
    size_t flags_off;                                 /* arrival counters: uint32[DBOT_B200_MAX_DEVICES] */
    size_t block_bytes;
    uint32_t seq;  /* gathering calls so far */
    double* h_all; /* pinned [world * width_max] */
    int connected;
};

struct dbot_peer_signal
{
    uint32_t* peer_flags[DBOT_B200_MAX_DEVICES]; /* rank g's arrival counters */
    uint32_t* own_flags;
    int world, rank;
    uint32_t seq;
    int* err_flag;
};

#define DBOT_PEER_WAIT_CYCLES 20000000000LL /* about 10 s */

__global__ void dbot_peer_signal_wait_kernel(const dbot_peer_signal f)
{
    const int g = threadIdx.x;
    if (g >= f.world) return;
    /* the particle kernel before this one on the stream has stored this rank's slice into rank g's vector */
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.peer_flags[g] + f.rank), "r"(f.seq) : "memory");
    const long long t0 = clock64();
    for (;;)
    {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f.own_flags + g) : "memory");
        if ((int32_t)(v - f.seq) >= 0) break;
        __nanosleep(200);
        if (clock64() - t0 > DBOT_PEER_WAIT_CYCLES)
        {
            atomicExch(f.err_flag, 4);
            break;
        }
    }
}

static void peer_gather_close(dbot_b200_sensor* s)
{
    dbot_peer_gather* pg = s->peer;
    if (!pg) return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    for (int g = 0; g < pg->world; ++g)
        if (g != pg->rank && pg->peer_block[g]) cudaIpcCloseMemHandle(pg->peer_block[g]);
    cudaFree(pg->d_block);
    cudaFreeHost(pg->h_all);
    delete pg;
    s->peer = nullptr;
}

extern "C" int dbot_b200_peer_gather_export(dbot_b200_sensor* s, int32_t world, int32_t rank, int32_t width_max,
                                            void* out_handle)
{
    if (!s || !out_handle) return fail(-1, "peer_gather_export: null argument");
    DBOT_SINGLE_ONLY(s, "peer_gather_export");
    if (world < 1 || world > DBOT_B200_MAX_DEVICES || rank < 0 || rank >= world || width_max < 1)
        return fail(-2, "peer_gather_export: world=%d rank=%d width_max=%d (at most %d ranks)", world, rank, width_max,
                    DBOT_B200_MAX_DEVICES);
    static_assert(sizeof(cudaIpcMemHandle_t) == DBOT_B200_IPC_HANDLE_BYTES, "IPC handle size");
    peer_gather_close(s);
    CUDA_TRY(cudaSetDevice(s->device));
    dbot_peer_gather* pg = new dbot_peer_gather();
    memset((void*)pg, 0, sizeof(*pg));
    s->peer = pg;
    pg->world = world;
    pg->rank = rank;
    pg->width_max = width_max;
    pg->vec_bytes = ((size_t)world * width_max * sizeof(double) + 255) / 256 * 256;
    pg->flags_off = 2 * pg->vec_bytes;
    pg->block_bytes = pg->flags_off + 256;
    CUDA_TRY(cudaMalloc(&pg->d_block, pg->block_bytes));
    CUDA_TRY(cudaMemset(pg->d_block, 0, pg->block_bytes));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaHostAlloc(&pg->h_all, (size_t)world * width_max * sizeof(double), cudaHostAllocDefault));
    pg->peer_block[rank] = pg->d_block;
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, pg->d_block));
    memcpy(out_handle, &h, sizeof(h));
    s->device_bytes += (int64_t)pg->block_bytes;
    return 0;
}

extern "C" int dbot_b200_peer_gather_connect(dbot_b200_sensor* s, const void* handles)
{
    if (!s || !handles) return fail(-1, "peer_gather_connect: null argument");
    dbot_peer_gather* pg = s->peer;
    if (!pg) return fail(-2, "peer_gather_connect: call peer_gather_export first");
    CUDA_TRY(cudaSetDevice(s->device));
    const unsigned char* hs = (const unsigned char*)handles;
    for (int g = 0; g < pg->world; ++g)
    {
        if (g == pg->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs + (size_t)g * sizeof(h), sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
        {
            cudaGetLastError();
            return fail(-15, "peer_gather_connect: cannot map rank %d's gathered vector (%s); is there NVLink / P2P "
                             "between the two GPUs and do the processes share an IPC namespace?",
                        g, cudaGetErrorString(e));
        }
        pg->peer_block[g] = (unsigned char*)p;
    }
    pg->connected = 1;
    return 0;
}

/* device pointer of the gathered vector the last gathering call filled ([world * width] doubles) */
extern "C" void* dbot_b200_peer_gather_result(dbot_b200_sensor* s)
{
    if (!s || !s->peer) return nullptr;
    return s->peer->d_block + (size_t)(s->peer->seq & 1u) * s->peer->vec_bytes;
}

static int peer_gather_enqueue(dbot_b200_sensor* s, const float* d_poses, int32_t* d_indices, int32_t n_local,
                               int32_t update, double* d_out, int32_t width)
{
    dbot_peer_gather* pg = s->peer;
    if (!pg || !pg->connected) return fail(-2, "loglikes_gather: peer_gather_export / _connect have not been called");
    if (width < 1 || width > pg->width_max || n_local > width)
        return fail(-2, "loglikes_gather: width=%d (at most %d), n_local=%d", width, pg->width_max, n_local);
    pg->seq += 1u;
    const size_t buf = (size_t)(pg->seq & 1u) * pg->vec_bytes;
    int rc = 0;
    if (n_local > 0)
    {
        s->kp.n_peers = pg->world;
        s->kp.out_offset = pg->rank * width;
        for (int g = 0; g < pg->world; ++g) s->kp.peer_out[g] = (double*)(pg->peer_block[g] + buf);
        rc = launch_particles(s, d_poses, d_indices, n_local, update, d_out);
        s->kp.n_peers = 0;
        if (rc) return rc;
    }
    else if (update && (rc = empty_update(s)))
        return rc;
    dbot_peer_signal f;
    for (int g = 0; g < pg->world; ++g) f.peer_flags[g] = (uint32_t*)(pg->peer_block[g] + pg->flags_off);
    f.own_flags = (uint32_t*)(pg->d_block + pg->flags_off);
    f.world = pg->world;
    f.rank = pg->rank;
    f.seq = pg->seq;
    f.err_flag = s->d_err;
    dbot_peer_signal_wait_kernel<<<1, 32, 0, s->stream>>>(f);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

/* device-resident variant: pointers on this rank's GPU, nothing is waited for; the gathered vector is at
 * dbot_b200_peer_gather_result() once the stream has run */
extern "C" int dbot_b200_loglikes_gather_device(dbot_b200_sensor* s, const float* d_poses, int32_t* d_indices,
                                                int32_t n_local, int32_t update, double* d_out, int32_t width)
{
    if (!s) return fail(-1, "loglikes_gather_device: null handle");
    if (n_local < 0 || n_local > s->cfg.max_particles)
        return fail(-2, "loglikes_gather_device: n_local=%d exceeds max_particles=%d", n_local, s->cfg.max_particles);
    if (n_local > 0 && (!d_poses || !d_indices || !d_out)) return fail(-1, "loglikes_gather_device: null argument");
    DBOT_CHECK_USABLE(s, "loglikes_gather_device");
    DBOT_SINGLE_ONLY(s, "loglikes_gather_device");
    CUDA_TRY(cudaSetDevice(s->device));
    if (update) s->update_since_sync = 1;
    return peer_gather_enqueue(s, d_poses, d_indices, n_local, update, d_out, width);
}

/* RbSensor::loglikes of a particle-sharded sensor, one process per GPU: this rank's slice in (host arrays,
 * parents in this rank's own slot numbering), every rank's log-weights out: out_all[g * width + i] is particle
 * i of rank g's slice.  Every rank of the group must make the same sequence of gathering calls. */
extern "C" int dbot_b200_loglikes_gather(dbot_b200_sensor* s, const float* poses_local, const int32_t* indices_local,
                                         int32_t n_local, int32_t update, int32_t width, double* out_all)
{
    if (!s || !out_all) return fail(-1, "loglikes_gather: null argument");
    if (n_local < 0 || n_local > s->cfg.max_particles)
        return fail(-2, "loglikes_gather: n_local=%d exceeds max_particles=%d", n_local, s->cfg.max_particles);
    if (n_local > 0 && (!poses_local || !indices_local)) return fail(-1, "loglikes_gather: null argument");
    DBOT_CHECK_USABLE(s, "loglikes_gather");
    DBOT_SINGLE_ONLY(s, "loglikes_gather");
    dbot_peer_gather* pg = s->peer;
    if (!pg || !pg->connected) return fail(-2, "loglikes_gather: peer_gather_export / _connect have not been called");
    for (int i = 0; i < n_local; ++i)
        if (indices_local[i] < 0 || indices_local[i] >= s->n_slots)
            return fail(-7, "loglikes_gather: indices[%d]=%d is not an occlusion slot of this rank (have %d)", i,
                        indices_local[i], s->n_slots);
    CUDA_TRY(cudaSetDevice(s->device));
    const double ht0 = s->host_timing ? host_now_us() : 0.0;
    int32_t* d_indices = nullptr;
    int rc = n_local > 0 ? stage_call_inputs(s, poses_local, indices_local, n_local, &d_indices) : 0;
    if (rc) return rc;
    rc = peer_gather_enqueue(s, s->d_poses, d_indices, n_local, update, s->d_out, width);
    if (rc) return rc;
    const size_t bytes = (size_t)pg->world * width * sizeof(double);
    CUDA_TRY(cudaMemcpyAsync(pg->h_all, dbot_b200_peer_gather_result(s), bytes, cudaMemcpyDeviceToHost, s->stream));
    const double ht1 = s->host_timing ? host_now_us() : 0.0;
    rc = loglikes_finish(s, 0, update, nullptr);
    s->obs_in_flight = 0;
    if (rc) return rc;
    memcpy(out_all, pg->h_all, bytes);
    if (s->host_timing)
    {
        s->ht_us[1] += ht1 - ht0;
        s->ht_us[2] += host_now_us() - ht1;
        s->ht_ll_calls++;
    }
    return 0;
}

/* Host-side check of a sharded call's parent vector (what sharded.py needs before every call): are all n
 * parents valid slots, and does every child's parent live on the child's own rank?  Ranks own contiguous,
 * balanced slices (sharded.py shard_bounds: the first n % world ranks hold one item more).
 * Returns 0 = all local, 1 = some parent lives on another rank (out_first = the first such child), -7 = a parent
 * is not a slot. */
static inline int shard_owner_of(int i, int base, int extra)
{
    const int cut = extra * (base + 1);
    return i < cut ? i / (base + 1) : extra + (i - cut) / (base > 0 ? base : 1);
}
extern "C" int dbot_b200_shard_scan(const int32_t* indices, int32_t n, int32_t n_slots, int32_t world,
                                    int32_t* out_first)
{
    if ((n > 0 && !indices) || n < 0 || n_slots < 1 || world < 1) return fail(-1, "shard_scan: bad argument");
    const int cb = n / world, ce = n % world, pb = n_slots / world, pe = n_slots % world;
    int first = -1;
    for (int i = 0; i < n; ++i)
    {
        const int p = indices[i];
        if (p < 0 || p >= n_slots)
            return fail(-7, "loglikes: indices[%d]=%d is not an occlusion slot (have %d)", i, p, n_slots);
        if (first < 0 && n_slots > 1 && shard_owner_of(i, cb, ce) != shard_owner_of(p, pb, pe)) first = i;
    }
    if (out_first) *out_first = first;
    return first >= 0 ? 1 : 0;
}
