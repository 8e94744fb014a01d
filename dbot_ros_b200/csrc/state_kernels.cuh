/*
 * state_kernels.cuh -- per-frame observation table, occlusion-slot pool maintenance and the update plan
 * Part of libdbot_b200.so; included by dbot_b200.cu (one translation unit).
 */
#pragma once

/* ------------------------------------------------------------------------------------------ */
/* K1: per-frame observation table.  Everything in the pixel densities that depends only on
 * the observed depth y (dbot KinectPixelModel) is evaluated once per frame in
 * double and stored as float4 { y, 1/(sqrt2 sigma), C_v/p_inf, (w/D)/p_inf }.                */
__global__ void dbot_observation_kernel(const float* __restrict__ depth, float4* __restrict__ tab,
                                        int rows, int cols, int pitch, double tail_weight,
                                        double model_sigma, double sigma_factor, double lambda,
                                        double wD)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * pitch) return;
    const int r = idx / pitch, c = idx - r * pitch;
    float4 t = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
    if (c < cols)
    {
        const float yf = depth[r * cols + c];
        if (yf == yf)
        {
            const double y = (double)yf;
            const double sigma = model_sigma + sigma_factor * y * y;
            const double g = 1.0 / (1.4142135623730951 * sigma);
            const double cv = (1.0 - tail_weight) / (2.5066282746310002 * sigma);
            const double co =
                0.5 * (1.0 - tail_weight) * lambda * exp(0.5 * lambda * (-2.0 * y + lambda * sigma * sigma));
            const double pinf = wD + 2.0 * co;
            t = make_float4(yf, (float)g, (float)(cv / pinf), (float)(wD / pinf));
        }
    }
    tab[idx] = t;
}

__global__ void dbot_fill_kernel(float* __restrict__ dst, size_t n, float v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = v;
}

/* rebase of the lazily propagated state: stored <- A * stored + B in every physical slot, the unused ones
 * included (they keep holding the background value, see dbot_kparams::slot_rows); rare: once the
 * accumulated factor A has shrunk to 2^-60, i.e. after roughly 80 s of tracking */
__global__ void dbot_rebase_kernel(float* __restrict__ pool, size_t n, float A, float B)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        pool[i] = __fmaf_rn(A, pool[i], B);
}

/* ------------------------------------------------------------------------------------------ */
/* Update planning (one CTA).  dbot's update gives every particle a new occlusion image: a copy of
 * its parent's, with the covered pixels replaced.  Copying 4*P bytes per particle is the HBM cost
 * of the path, and half of it is avoidable: the LAST child of every parent simply adopts the
 * parent's physical slot (nothing else will ever read that parent again), only the other children
 * get a fresh slot and a copy.  Without resampling (parents = identity) nothing is copied at all.
 * Physical slots live in a pool of 2 * max_particles; `loc` maps logical slot -> physical slot;
 * `free_stack` holds the unused ones.  Outputs: plan[i] = { source, destination, parent, 0 },
 * order[] = copying children first (so that an adopting child never waits for a copy that has not
 * been scheduled yet), pending[parent] = number of copies that must read the parent before the
 * adopting child may write into it, loc_next[i] = destination. */
#define DBOT_PLAN_THREADS 1024
__device__ __forceinline__ int dbot_block_exclusive_scan(int v, int* s_warp, int* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const int o = __shfl_up_sync(DBOT_FULL, inc, d);
        if (lane >= d) inc += o;
    }
    __syncthreads(); /* s_warp may still be read from a previous scan */
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0)
    {
        const int w = s_warp[lane];
        int winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const int o = __shfl_up_sync(DBOT_FULL, winc, d);
            if (lane >= d) winc += o;
        }
        s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[warp] + inc - v;
}

__global__ void __launch_bounds__(DBOT_PLAN_THREADS, 1)
    dbot_plan_update_kernel(const int32_t* __restrict__ parents, int n, int n_slots,
                            const int32_t* __restrict__ loc, int32_t* __restrict__ loc_next,
                            int32_t* __restrict__ free_stack, int32_t* __restrict__ free_count,
                            int32_t* cnt, int32_t* last,
                            int32_t* __restrict__ pending, int32_t* __restrict__ order,
                            int4* __restrict__ plan, int* __restrict__ err_flag, int tables_in_smem)
{
    __shared__ int s_warp[33];
    extern __shared__ int32_t s_tables[];
    const int t = threadIdx.x;
    /* children per parent and each parent's last child: in shared memory when they fit (the kernel is a chain of
     * dependent look-ups by one CTA; with the two tables in global memory it takes 11 us for 1000 particles) */
    if (tables_in_smem)
    {
        cnt = s_tables;
        last = s_tables + n_slots;
    }
    for (int sidx = t; sidx < n_slots; sidx += DBOT_PLAN_THREADS)
    {
        cnt[sidx] = 0;
        last[sidx] = -1;
    }
    __syncthreads();
    for (int i = t; i < n; i += DBOT_PLAN_THREADS)
    {
        const int par = parents[i];
        if (par >= 0 && par < n_slots)
        {
            atomicAdd(cnt + par, 1);
            atomicMax(last + par, i);
        }
        else
            atomicExch(err_flag, 1);
    }
    __syncthreads();
    /* contiguous ranges per thread keep the order of copies and adoptions stable */
    const int per = (n + DBOT_PLAN_THREADS - 1) / DBOT_PLAN_THREADS;
    const int i0 = min(t * per, n), i1 = min(i0 + per, n);
    int copies = 0, adopts = 0;
    for (int i = i0; i < i1; ++i)
    {
        const int par = parents[i];
        const bool ok = par >= 0 && par < n_slots;
        if (ok && last[par] == i) ++adopts; else ++copies; /* an orphan takes a fresh slot, unfilled */
    }
    int n_copies, n_adopts;
    int crank = dbot_block_exclusive_scan(copies, s_warp, &n_copies);
    int arank = dbot_block_exclusive_scan(adopts, s_warp, &n_adopts);
    const int f0 = *free_count;
    if (f0 < n_copies)
    {
        /* cannot happen while live + imported slots <= pool / 2 (see dbot_b200_import_slot) */
        if (t == 0) atomicExch(err_flag, 2);
        for (int i = i0; i < i1; ++i)
        {
            plan[i] = make_int4(-1, 0, 0, 0);
            order[i] = i;
            loc_next[i] = loc[0];
        }
        return;
    }
    for (int i = i0; i < i1; ++i)
    {
        const int par = parents[i];
        const bool ok = par >= 0 && par < n_slots;
        int4 pl;
        if (ok && last[par] == i)
        {
            pl = make_int4(loc[par], loc[par], par, 0);
#ifdef DBOT_ORDER_COPIES_FIRST
            order[n_copies + arank] = i;
#endif
            ++arank;
        }
        else
        {
            pl = make_int4(ok ? loc[par] : -1, free_stack[f0 - 1 - crank], ok ? par : 0, 0);
#ifdef DBOT_ORDER_COPIES_FIRST
            order[crank] = i;
#endif
            ++crank;
        }
#ifndef DBOT_ORDER_COPIES_FIRST
        /* the caller's order: copying and adopting children alternate as resampling left them, so that the
         * copy stream's HBM traffic is spread over the whole launch; a parent's copies sit just before its
         * adopter and are taken by the neighbouring CTAs' free-running streams */
        order[i] = i;
#endif
        plan[i] = pl;
        loc_next[i] = pl.y;
    }
    /* parents without children give their slots back */
    const int pers = (n_slots + DBOT_PLAN_THREADS - 1) / DBOT_PLAN_THREADS;
    const int s0 = min(t * pers, n_slots), s1 = min(s0 + pers, n_slots);
    int freed = 0;
    for (int sidx = s0; sidx < s1; ++sidx) freed += (cnt[sidx] == 0);
    int n_freed;
    int frank = dbot_block_exclusive_scan(freed, s_warp, &n_freed);
    __syncthreads(); /* every thread has taken its fresh slots from the top of the stack */
    for (int sidx = s0; sidx < s1; ++sidx)
    {
        if (cnt[sidx] == 0)
            free_stack[f0 - n_copies + frank++] = loc[sidx];
        pending[sidx] = max(cnt[sidx] - 1, 0);
    }
    __syncthreads();
    if (t == 0) *free_count = f0 - n_copies + n_freed;
}

