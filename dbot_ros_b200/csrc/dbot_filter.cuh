/*
 * dbot_filter.cuh -- device-resident particle-filter step around the sensor (SURVEY.md section 8f,
 * row N1).  Included at the end of dbot_b200.cu (same translation unit: it drives the sensor through
 * launch_observation / launch_particles and never leaves the device between them).
 *
 * What dbot does per frame on the host (Rao-Blackwellised coordinate particle filter + ParticleTracker;
 * un-vendored dbot, call pattern restated in SURVEY.md A.7; PARITY UNPINNED like the rest of dbot):
 *     set_observation(frame)
 *     per sampling block b (one per object):
 *         draw N(0,1) for the block's 6 dimensions; next = transition(states, noise)
 *         ll = sensor.loglikes(next, indices, update = last block)
 *         log_w += ll - previous ll; normalise; if KL(w || uniform) > max_kl_divergence: resample
 *     states = next; fold the weighted mean delta into the integrated pose, re-base the particles
 * Here every one of these O(N) steps is a kernel on the sensor's stream, so that states, poses,
 * parent indices and weights never cross PCIe: per frame the host sends the depth image and gets back
 * n_objects * 6 doubles.  The motion model is the one ref:config/particle_tracker.yaml:51-63
 * parameterises (ref:source/dbot_ros/tracker/particle_tracker_node.cpp:138-159):
 *     v' = velocity_factor * v + sigma * w / dt,   pose' = pose + v' * dt,   w ~ N(0, I)
 * dbot_ros_b200/tracker.py holds the same loop in numpy (the reference for the tests, including the
 * counter-based noise generator below).
 */
#pragma once

struct dbot_b200_filter
{
    dbot_b200_sensor* s;
    dbot_b200_filter_config cfg;
    int n, n_obj;
    uint64_t frame;
    /* device state; [2] = ping-pong across a (possibly identity) resampling gather */
    double* d_states[2];   /* [n][n_obj][12] position, rotation vector, linear and angular velocity (deltas) */
    double* d_next[2];     /* [n][n_obj][12] */
    double* d_noise[2];    /* [n][n_obj][6]  */
    double* d_prev_ll[2];  /* [n] */
    int32_t* d_indices[2]; /* [n] */
    double* d_log_w;       /* [n] */
    double* d_ll;          /* [n] sensor output */
    float* d_poses;        /* [n][n_obj][12] */
    int32_t* d_parents;    /* [n] */
    double* d_cdf;         /* [n] */
    double* d_integrated;  /* [n_obj][6] poses + [n_obj][9] rotation matrices + [n_obj][6] mean velocities */
    int32_t* d_info;       /* [0] resamplings since initialize, [1] resample flag of the current block */
    double* h_integrated;  /* pinned, same layout */
    int32_t* h_info;       /* pinned [2] */
    int cur;
    /* Tracker::track's moving average over the published state (centre frame), host side */
    double smoothed[DBOT_B200_MAX_OBJECTS][12];
    int have_smoothed;
};

/* ---- host: rotation vector <-> quaternion, for the moving average and the centre <-> model frame ---- */
static void dbot_h_quat(const double* v, double* q)
{
    const double a = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const double k = a > 1e-12 ? sin(0.5 * a) / a : 0.5;
    q[0] = cos(0.5 * a); q[1] = k * v[0]; q[2] = k * v[1]; q[3] = k * v[2];
}
static void dbot_h_rotvec(const double* q, double* v)
{
    double w = q[0], x = q[1], y = q[2], z = q[3];
    if (w < 0.0) { w = -w; x = -x; y = -y; z = -z; }
    const double sn = sqrt(x * x + y * y + z * z);
    const double a = 2.0 * atan2(sn, w);
    const double k = sn > 1e-12 ? a / sn : 2.0;
    v[0] = k * x; v[1] = k * y; v[2] = k * z;
}
/* R(v) c, Rodrigues */
static void dbot_h_rotate(const double* v, const double* c, double* out)
{
    const double a = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (!(a > 0.0)) { out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; return; }
    const double k[3] = {v[0] / a, v[1] / a, v[2] / a};
    const double cs = cos(a), sn = sin(a);
    const double kc = k[0] * c[0] + k[1] * c[1] + k[2] * c[2];
    const double kxc[3] = {k[1] * c[2] - k[2] * c[1], k[2] * c[0] - k[0] * c[2], k[0] * c[1] - k[1] * c[0]};
    for (int i = 0; i < 3; ++i) out[i] = c[i] * cs + kxc[i] * sn + k[i] * kc * (1.0 - cs);
}

/* ---- counter-based noise: splitmix64 of (seed, frame, block, particle, pair) -> Box-Muller ---- */
__host__ __device__ inline uint64_t dbot_splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ inline void dbot_normal_pair(uint64_t key, double* a, double* b)
{
    const uint64_t h1 = dbot_splitmix64(key), h2 = dbot_splitmix64(h1);
    const double u1 = ((double)(h1 >> 11) + 1.0) * (1.0 / 9007199254740992.0); /* (0, 1] */
    const double u2 = (double)(h2 >> 11) * (1.0 / 9007199254740992.0);         /* [0, 1) */
    const double r = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    *a = r * cs;
    *b = r * sn;
}
__host__ __device__ inline uint64_t dbot_noise_key(uint64_t seed, uint64_t frame, int block, int particle, int pair)
{
    return dbot_splitmix64(seed ^ (frame * 0x100000001B3ull)) ^ ((uint64_t)block << 56) ^
           ((uint64_t)particle << 8) ^ (uint64_t)pair;
}

/* ---- rotation vector <-> matrix (double) ---- */
__device__ inline void dbot_rotmat(const double* v, double* R)
{
    const double angle = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (!(angle > 0.0))
    {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        return;
    }
    const double x = v[0] / angle, y = v[1] / angle, z = v[2] / angle;
    const double c = cos(angle), sn = sin(angle), C1 = 1.0 - c;
    R[0] = c + x * x * C1;     R[1] = x * y * C1 - z * sn; R[2] = x * z * C1 + y * sn;
    R[3] = y * x * C1 + z * sn; R[4] = c + y * y * C1;     R[5] = y * z * C1 - x * sn;
    R[6] = z * x * C1 - y * sn; R[7] = z * y * C1 + x * sn; R[8] = c + z * z * C1;
}
__device__ inline void dbot_matmul3(const double* A, const double* B, double* C)
{
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}
__device__ inline void dbot_rotvec(const double* R, double* v)
{
    double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
    c = fmin(fmax(c, -1.0), 1.0);
    const double angle = acos(c), sn = sin(angle);
    const double scale = sn > 1e-9 ? angle / (2.0 * sn) : 0.5;
    v[0] = (R[7] - R[5]) * scale;
    v[1] = (R[2] - R[6]) * scale;
    v[2] = (R[3] - R[1]) * scale;
}

/* ---- kernels ---- */
__global__ void dbot_filter_init_kernel(double* states, double* log_w, double* prev_ll, int32_t* indices,
                                        int32_t* info, int n, int n_obj)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) info[0] = info[1] = 0;
    if (i >= n) return;
    for (int k = 0; k < 12 * n_obj; ++k) states[(size_t)i * 12 * n_obj + k] = 0.0;
    log_w[i] = -log((double)n);
    prev_ll[i] = 0.0;
    indices[i] = 0;
}

/* noise for sampling block `block`, transition of every object, pose composition for the sensor */
__global__ void dbot_filter_propagate_kernel(const double* __restrict__ states, double* __restrict__ noise,
                                             double* __restrict__ next, float* __restrict__ poses,
                                             const double* __restrict__ integrated, dbot_b200_filter_config cfg,
                                             double dt, uint64_t frame, int block, int n, int n_obj)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* nz = noise + (size_t)i * 6 * n_obj;
    if (block == 0)
        for (int k = 0; k < 6 * n_obj; ++k) nz[k] = 0.0;
    for (int pr = 0; pr < 3; ++pr)
        dbot_normal_pair(dbot_noise_key(cfg.seed, frame, block, i, pr), &nz[6 * block + 2 * pr],
                         &nz[6 * block + 2 * pr + 1]);
    for (int o = 0; o < n_obj; ++o)
    {
        const double* st = states + ((size_t)i * n_obj + o) * 12;
        double* nx = next + ((size_t)i * n_obj + o) * 12;
        for (int d = 0; d < 6; ++d)
        {
            const double sig = d < 3 ? cfg.linear_sigma[d] : cfg.angular_sigma[d - 3];
            const double v = cfg.velocity_factor * st[6 + d] + nz[6 * o + d] * (sig / dt);
            nx[6 + d] = v;
            nx[d] = st[d] + v * dt;
        }
        double Rd[9], R[9];
        dbot_rotmat(nx + 3, Rd);
        dbot_matmul3(Rd, integrated + 6 * n_obj + 9 * o, R);
        float* p = poses + ((size_t)i * n_obj + o) * 12;
        for (int k = 0; k < 9; ++k) p[k] = (float)R[k];
        for (int k = 0; k < 3; ++k) p[9 + k] = (float)(nx[k] + integrated[6 * o + k]);
    }
}

#define DBOT_FILTER_THREADS 1024
__device__ inline double dbot_block_sum(double v, double* s_red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(DBOT_FULL, v, off);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < DBOT_FILTER_THREADS / 32; ++w) t += s_red[w]; /* fixed order: deterministic */
    return t;
}
__device__ inline double dbot_block_max(double v, double* s_red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_down_sync(DBOT_FULL, v, off));
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double t = s_red[0];
    for (int w = 1; w < DBOT_FILTER_THREADS / 32; ++w) t = fmax(t, s_red[w]);
    return t;
}

/* weights of one sampling block (one CTA): log_w += ll - prev; normalise; KL against uniform; when it
 * exceeds max_kl_divergence, systematic resampling -> parents (else identity) */
__global__ void __launch_bounds__(DBOT_FILTER_THREADS, 1)
    dbot_filter_weights_kernel(const double* __restrict__ ll, double* __restrict__ prev_ll,
                               double* __restrict__ log_w, double* __restrict__ cdf,
                               int32_t* __restrict__ parents, int32_t* __restrict__ info, double max_kl,
                               uint64_t seed, uint64_t frame, int block, int n)
{
    __shared__ double s_red[DBOT_FILTER_THREADS / 32];
    __shared__ double s_carry;
    const int t = threadIdx.x;
    double mx = -1.0e300;
    for (int i = t; i < n; i += DBOT_FILTER_THREADS)
    {
        const double v = log_w[i] + (ll[i] - prev_ll[i]);
        log_w[i] = v;
        prev_ll[i] = ll[i];
        mx = fmax(mx, v);
    }
    mx = dbot_block_max(mx, s_red);
    double se = 0.0;
    for (int i = t; i < n; i += DBOT_FILTER_THREADS) se += exp(log_w[i] - mx);
    se = dbot_block_sum(se, s_red);
    const double lse = mx + log(se);
    double ent = 0.0;
    for (int i = t; i < n; i += DBOT_FILTER_THREADS)
    {
        const double lw = log_w[i] - lse;
        log_w[i] = lw;
        ent += exp(lw) * lw;
    }
    ent = dbot_block_sum(ent, s_red);
    const double kl = log((double)n) + ent;
    const bool resample = kl > max_kl;
    if (t == 0)
    {
        info[1] = resample ? 1 : 0;
        if (resample) info[0] += 1;
        s_carry = 0.0;
    }
    __syncthreads();
    if (!resample)
    {
        for (int i = t; i < n; i += DBOT_FILTER_THREADS) parents[i] = i;
        return;
    }
    /* inclusive prefix sum of the weights in index order (tiles of DBOT_FILTER_THREADS) */
    for (int base = 0; base < n; base += DBOT_FILTER_THREADS)
    {
        const int i = base + t;
        double v = i < n ? exp(log_w[i]) : 0.0;
        const int lane = t & 31, warp = t >> 5;
        for (int d = 1; d < 32; d <<= 1)
        {
            const double o = __shfl_up_sync(DBOT_FULL, v, d);
            if (lane >= d) v += o;
        }
        __syncthreads();
        if (lane == 31) s_red[warp] = v;
        __syncthreads();
        double off = s_carry;
        for (int w = 0; w < warp; ++w) off += s_red[w];
        if (i < n) cdf[i] = off + v;
        __syncthreads();
        if (t == DBOT_FILTER_THREADS - 1) s_carry = off + v;
        __syncthreads();
    }
    const double total = s_carry;
    const uint64_t h = dbot_splitmix64(dbot_noise_key(seed, frame, block, 0x7fffff, 7));
    const double u0 = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    for (int j = t; j < n; j += DBOT_FILTER_THREADS)
    {
        const double u = (u0 + (double)j) / (double)n * total;
        int lo = 0, hi = n - 1; /* first index with cdf > u */
        while (lo < hi)
        {
            const int mid = (lo + hi) >> 1;
            if (cdf[mid] > u) hi = mid; else lo = mid + 1;
        }
        parents[j] = lo;
    }
    for (int i = t; i < n; i += DBOT_FILTER_THREADS) log_w[i] = -log((double)n);
}

/* gather by parents (identity when the block did not resample) into the other buffer set */
__global__ void dbot_filter_gather_kernel(const int32_t* __restrict__ parents, const double* __restrict__ st_in,
                                          double* __restrict__ st_out, const double* __restrict__ nx_in,
                                          double* __restrict__ nx_out, const double* __restrict__ nz_in,
                                          double* __restrict__ nz_out, const double* __restrict__ pl_in,
                                          double* __restrict__ pl_out, const int32_t* __restrict__ ix_in,
                                          int32_t* __restrict__ ix_out, int n, int n_obj)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int par = parents[i];
    const size_t a = (size_t)i * 12 * n_obj, b = (size_t)par * 12 * n_obj;
    for (int k = 0; k < 12 * n_obj; ++k)
    {
        st_out[a + k] = st_in[b + k];
        nx_out[a + k] = nx_in[b + k];
    }
    for (int k = 0; k < 6 * n_obj; ++k) nz_out[(size_t)i * 6 * n_obj + k] = nz_in[(size_t)par * 6 * n_obj + k];
    pl_out[i] = pl_in[par];
    ix_out[i] = ix_in[par];
}

/* ParticleTracker::on_track: weighted mean delta -> integrated pose, particles re-based on it (one CTA) */
__global__ void __launch_bounds__(DBOT_FILTER_THREADS, 1)
    dbot_filter_mean_kernel(double* __restrict__ states, const double* __restrict__ next,
                            const double* __restrict__ log_w, double* __restrict__ integrated, int n, int n_obj)
{
    __shared__ double s_red[DBOT_FILTER_THREADS / 32];
    __shared__ double s_mean[12];
    __shared__ double s_Rm[9];
    const int t = threadIdx.x;
    for (int o = 0; o < n_obj; ++o)
    {
        for (int d = 0; d < 12; ++d) /* pose delta and velocity (the State the reference publishes carries both) */
        {
            double acc = 0.0;
            for (int i = t; i < n; i += DBOT_FILTER_THREADS) acc += exp(log_w[i]) * next[((size_t)i * n_obj + o) * 12 + d];
            acc = dbot_block_sum(acc, s_red);
            if (t == 0) s_mean[d] = acc;
        }
        __syncthreads();
        if (t == 0)
        {
            double R0n[9], v[3];
            dbot_rotmat(s_mean + 3, s_Rm);
            dbot_matmul3(s_Rm, integrated + 6 * n_obj + 9 * o, R0n);
            for (int k = 0; k < 9; ++k) integrated[6 * n_obj + 9 * o + k] = R0n[k];
            dbot_rotvec(R0n, v);
            for (int k = 0; k < 3; ++k)
            {
                integrated[6 * o + k] += s_mean[k];
                integrated[6 * o + 3 + k] = v[k];
            }
            for (int k = 0; k < 6; ++k) integrated[15 * n_obj + 6 * o + k] = s_mean[6 + k];
        }
        __syncthreads();
        for (int i = t; i < n; i += DBOT_FILTER_THREADS)
        {
            const double* nx = next + ((size_t)i * n_obj + o) * 12;
            double* st = states + ((size_t)i * n_obj + o) * 12;
            double Rd[9], RmT[9], R[9], v[3];
            dbot_rotmat(nx + 3, Rd);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) RmT[3 * r + c] = s_Rm[3 * c + r];
            dbot_matmul3(Rd, RmT, R);
            dbot_rotvec(R, v);
            for (int k = 0; k < 3; ++k)
            {
                st[k] = nx[k] - s_mean[k];
                st[3 + k] = v[k];
            }
            for (int k = 6; k < 12; ++k) st[k] = nx[k];
        }
        __syncthreads();
    }
}

__global__ void dbot_filter_set_pose_kernel(double* integrated, const double* poses6, int n_obj)
{
    const int o = threadIdx.x;
    if (o >= n_obj) return;
    for (int k = 0; k < 6; ++k) integrated[6 * o + k] = poses6[6 * o + k];
    dbot_rotmat(poses6 + 6 * o + 3, integrated + 6 * n_obj + 9 * o);
}

/* ---- host side ---- */
extern "C" int dbot_b200_filter_destroy(dbot_b200_filter* f)
{
    if (!f) return fail(-1, "filter_destroy: null handle");
    cudaSetDevice(f->s->device);
    cudaStreamSynchronize(f->s->stream);
    for (int k = 0; k < 2; ++k)
    {
        cudaFree(f->d_states[k]); cudaFree(f->d_next[k]); cudaFree(f->d_noise[k]);
        cudaFree(f->d_prev_ll[k]); cudaFree(f->d_indices[k]);
    }
    cudaFree(f->d_log_w); cudaFree(f->d_ll); cudaFree(f->d_poses); cudaFree(f->d_parents);
    cudaFree(f->d_cdf); cudaFree(f->d_integrated); cudaFree(f->d_info);
    cudaFreeHost(f->h_integrated); cudaFreeHost(f->h_info);
    delete f;
    return 0;
}

extern "C" int dbot_b200_filter_create(dbot_b200_sensor* s, const dbot_b200_filter_config* cfg,
                                       dbot_b200_filter** out)
{
    if (!s || !cfg || !out) return fail(-1, "filter_create: null argument");
    *out = nullptr;
    DBOT_SINGLE_ONLY(s, "filter_create");
    const int n_obj = s->kp.n_obj;
    const int n = cfg->n_particles;
    if (n <= 0 || n > s->cfg.max_particles)
        return fail(-2, "filter_create: n_particles=%d not in [1,%d]", n, s->cfg.max_particles);
    if (!(cfg->velocity_factor >= 0.0) || !(cfg->max_kl_divergence >= 0.0))
        return fail(-4, "filter_create: velocity_factor and max_kl_divergence must be non-negative");
    if (!(cfg->moving_average_update_rate > 0.0) || cfg->moving_average_update_rate > 1.0)
        return fail(-4, "filter_create: moving_average_update_rate must be in (0, 1]");
    CUDA_TRY(cudaSetDevice(s->device));
    dbot_b200_filter* f = new dbot_b200_filter();
    memset(f, 0, sizeof(*f));
    f->s = s;
    f->cfg = *cfg;
    f->n = n;
    f->n_obj = n_obj;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    const size_t st = (size_t)n * n_obj * 12 * sizeof(double);
    for (int k = 0; k < 2; ++k)
    {
        alloc((void**)&f->d_states[k], st);
        alloc((void**)&f->d_next[k], st);
        alloc((void**)&f->d_noise[k], st / 2);
        alloc((void**)&f->d_prev_ll[k], n * sizeof(double));
        alloc((void**)&f->d_indices[k], n * sizeof(int32_t));
    }
    alloc((void**)&f->d_log_w, n * sizeof(double));
    alloc((void**)&f->d_ll, n * sizeof(double));
    alloc((void**)&f->d_poses, (size_t)n * n_obj * 12 * sizeof(float));
    alloc((void**)&f->d_parents, n * sizeof(int32_t));
    alloc((void**)&f->d_cdf, n * sizeof(double));
    alloc((void**)&f->d_integrated, (size_t)n_obj * 21 * sizeof(double));
    alloc((void**)&f->d_info, 2 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&f->h_integrated, (size_t)n_obj * 21 * sizeof(double), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&f->h_info, 2 * sizeof(int32_t), cudaHostAllocDefault);
    if (e != cudaSuccess)
    {
        dbot_b200_filter_destroy(f);
        return fail((int)e, "filter_create: %s", cudaGetErrorString(e));
    }
    *out = f;
    return 0;
}

/* Tracker::initialize (ref:source/dbot_ros/tracker/particle_tracker_node.cpp:252) */
extern "C" int dbot_b200_filter_initialize(dbot_b200_filter* f, const double* poses)
{
    if (!f || !poses) return fail(-1, "filter_initialize: null argument");
    dbot_b200_sensor* s = f->s;
    int rc = dbot_b200_reset(s);
    if (rc) return rc;
    memcpy(f->h_integrated, poses, (size_t)f->n_obj * 6 * sizeof(double));
    /* dbot Tracker::initialize -> to_center_coordinate_system: the filter tracks the centred mesh, whose
     * origin sits at R c in the model frame the caller speaks (center_object_frame,
     * ref:source/dbot_ros/tracker/particle_tracker_node.cpp:210-214) */
    for (int o = 0; o < f->n_obj; ++o)
    {
        double rc[3];
        dbot_h_rotate(poses + 6 * o + 3, f->cfg.centers[o], rc);
        for (int k = 0; k < 3; ++k) f->h_integrated[6 * o + k] += rc[k];
    }
    f->have_smoothed = 0;
    /* staged through d_cdf-independent scratch: the tail of d_integrated is rewritten by the kernel */
    double* d_tmp = nullptr;
    CUDA_TRY(cudaMalloc(&d_tmp, (size_t)f->n_obj * 6 * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d_tmp, f->h_integrated, (size_t)f->n_obj * 6 * sizeof(double),
                                    cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess)
    {
        dbot_filter_set_pose_kernel<<<1, 32, 0, s->stream>>>(f->d_integrated, d_tmp, f->n_obj);
        dbot_filter_init_kernel<<<(f->n + 255) / 256, 256, 0, s->stream>>>(f->d_states[0], f->d_log_w, f->d_prev_ll[0],
                                                                          f->d_indices[0], f->d_info, f->n, f->n_obj);
        s->launches += 2;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d_tmp);
    if (e != cudaSuccess) return fail((int)e, "filter_initialize: %s", cudaGetErrorString(e));
    f->cur = 0;
    f->frame = 0;
    return 0;
}

/* Tracker::track (ref:source/dbot_ros/object_tracker_ros.hpp:49-61): one depth frame in, the published state
 * [n_obj][12] out (position, rotation vector, linear and angular velocity; model frame, moving average
 * applied); out_resamplings = resampling steps since initialize */
extern "C" int dbot_b200_filter_track(dbot_b200_filter* f, const float* depth_m, double* out_states,
                                      int32_t* out_resamplings)
{
    if (!f || !depth_m || !out_states) return fail(-1, "filter_track: null argument");
    dbot_b200_sensor* s = f->s;
    CUDA_TRY(cudaSetDevice(s->device));
    int rc = dbot_b200_set_observation(s, depth_m);
    if (rc) return rc;
    const int n = f->n, n_obj = f->n_obj, blocks = (n + 255) / 256;
    const double dt = s->cfg.delta_time;
    CUDA_TRY(cudaMemsetAsync(f->d_prev_ll[f->cur], 0, n * sizeof(double), s->stream));
    for (int b = 0; b < n_obj; ++b)
    {
        const int c = f->cur;
        dbot_filter_propagate_kernel<<<blocks, 256, 0, s->stream>>>(f->d_states[c], f->d_noise[c], f->d_next[c],
                                                                    f->d_poses, f->d_integrated, f->cfg, dt,
                                                                    f->frame, b, n, n_obj);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        rc = launch_particles(s, f->d_poses, f->d_indices[c], n, b == n_obj - 1, f->d_ll);
        if (rc) return rc;
        dbot_filter_weights_kernel<<<1, DBOT_FILTER_THREADS, 0, s->stream>>>(
            f->d_ll, f->d_prev_ll[c], f->d_log_w, f->d_cdf, f->d_parents, f->d_info, f->cfg.max_kl_divergence,
            f->cfg.seed, f->frame, b, n);
        dbot_filter_gather_kernel<<<blocks, 256, 0, s->stream>>>(
            f->d_parents, f->d_states[c], f->d_states[c ^ 1], f->d_next[c], f->d_next[c ^ 1], f->d_noise[c],
            f->d_noise[c ^ 1], f->d_prev_ll[c], f->d_prev_ll[c ^ 1], f->d_indices[c], f->d_indices[c ^ 1], n, n_obj);
        s->launches += 2;
        CUDA_TRY(cudaGetLastError());
        f->cur ^= 1;
    }
    const int c = f->cur;
    dbot_filter_mean_kernel<<<1, DBOT_FILTER_THREADS, 0, s->stream>>>(f->d_states[c], f->d_next[c], f->d_log_w,
                                                                      f->d_integrated, n, n_obj);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(f->h_integrated, f->d_integrated, (size_t)n_obj * 21 * sizeof(double),
                             cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaMemcpyAsync(f->h_info, f->d_info, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->h_err, s->d_err, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    rc = device_error(s, "filter_track", 1);
    if (rc) return rc;
    /* dbot Tracker::track: moving average over the state (positions and velocities linearly, orientations as
     * quaternions), then back to the model frame (to_model_coordinate_system) */
    const double a = f->cfg.moving_average_update_rate;
    for (int o = 0; o < n_obj; ++o)
    {
        double cur[12];
        for (int k = 0; k < 6; ++k) cur[k] = f->h_integrated[6 * o + k];
        for (int k = 0; k < 6; ++k) cur[6 + k] = f->h_integrated[15 * n_obj + 6 * o + k];
        double* sm = f->smoothed[o];
        if (!f->have_smoothed || a >= 1.0)
            memcpy(sm, cur, sizeof(cur));
        else
        {
            double qa[4], qb[4];
            dbot_h_quat(sm + 3, qa);
            dbot_h_quat(cur + 3, qb);
            const double sgn = (qa[0] * qb[0] + qa[1] * qb[1] + qa[2] * qb[2] + qa[3] * qb[3]) < 0.0 ? -1.0 : 1.0;
            double nrm = 0.0;
            for (int k = 0; k < 4; ++k)
            {
                qa[k] = (1.0 - a) * qa[k] + a * sgn * qb[k];
                nrm += qa[k] * qa[k];
            }
            nrm = sqrt(nrm);
            for (int k = 0; k < 4; ++k) qa[k] /= nrm;
            for (int k = 0; k < 3; ++k) sm[k] = (1.0 - a) * sm[k] + a * cur[k];
            dbot_h_rotvec(qa, sm + 3);
            for (int k = 6; k < 12; ++k) sm[k] = (1.0 - a) * sm[k] + a * cur[k];
        }
        double rc3[3];
        dbot_h_rotate(sm + 3, f->cfg.centers[o], rc3);
        for (int k = 0; k < 12; ++k) out_states[12 * o + k] = sm[k];
        for (int k = 0; k < 3; ++k) out_states[12 * o + k] -= rc3[k];
    }
    f->have_smoothed = 1;
    if (out_resamplings) *out_resamplings = f->h_info[0];
    f->frame += 1;
    return 0;
}
