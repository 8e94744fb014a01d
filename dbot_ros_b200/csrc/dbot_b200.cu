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
#include <string>
#include <unordered_map>
#include <vector>

#define DBOT_CW 16                 /* compute warps per CTA */
#define DBOT_CT (DBOT_CW * 32)     /* compute threads */
#define DBOT_SW 4                  /* streaming warps per CTA (update launches only) */
#define DBOT_ST (DBOT_SW * 32)
#define DBOT_STREAM_UNROLL 4       /* 256-bit loads in flight per streaming thread */
/* update launches start every thread with 96 registers (640 threads); the streaming warp group
 * hands most of its share to the four compute warp groups (setmaxnreg) */
#ifndef DBOT_PREFETCH_AHEAD
#define DBOT_PREFETCH_AHEAD 8      /* chunks of DBOT_ST x DBOT_STREAM_UNROLL x 32 B (16 KB) */
#endif
#define DBOT_STREAM_REGS 56
#define DBOT_COMPUTE_REGS 104
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
    int32_t obj, n_verts, n_tris, pad;
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
    size_t slot_stride;     /* rows*pitch */
    float post_b, post_inv_a; /* update: the posterior o is stored as (o - post_b) * post_inv_a */
    const float* poses;     /* [n][n_obj][12] */
    int32_t* indices;       /* [n] */
    double* out;            /* [n] */
    float* depth_out;       /* render: [rows*cols] */
    int* err_flag;
    /* shared-memory carve-up (bytes from the dynamic smem base) */
    uint32_t off_tri, off_v4, off_Y, off_wt;
    uint32_t wcap;          /* inverse-depth tile capacity in pixels */
};

/* ------------------------------------------------------------------------------------------ */
/* PTX wrappers: mbarrier, TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP), named barriers,     */
/* 256-bit streaming loads / stores (SASS LDG.256 / STG.256, new with sm_100)                   */
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
/* barrier among the compute warps only (the streaming warps never take part) */
__device__ __forceinline__ void compute_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(DBOT_CT) : "memory");
}
/* producer/consumer hand-off between the streaming and the compute warps on hardware named
 * barriers (no polling): the producer side arrives, the consumer side syncs */
#define DBOT_BAR_COPIED 2
#define DBOT_BAR_CONSUMED 3
__device__ __forceinline__ void role_arrive(int id)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(DBOT_CT + DBOT_ST) : "memory");
}
__device__ __forceinline__ void role_sync(int id)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(DBOT_CT + DBOT_ST) : "memory");
}
__device__ __forceinline__ void ldg256(const float* p, float (&v)[8])
{
    asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]),
                   "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256_stream(float* p, const float (&v)[8])
{
    asm volatile(
        "st.global.L1::no_allocate.L2::evict_first.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
        : "memory");
}

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

/* rebase of the lazily propagated state: stored <- A * stored + B in every live slot (rare: once the
 * accumulated factor A has shrunk to 2^-60, i.e. after roughly 80 s of tracking) */
__global__ void dbot_rebase_kernel(float* __restrict__ pool, const int32_t* __restrict__ loc, int n_slots,
                                   size_t slot_stride, float A, float B)
{
    for (int sidx = blockIdx.y; sidx < n_slots; sidx += gridDim.y)
    {
        float* dst = pool + (size_t)loc[sidx] * slot_stride;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < slot_stride;
             i += (size_t)gridDim.x * blockDim.x)
            dst[i] = __fmaf_rn(A, dst[i], B);
    }
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
                            int32_t* __restrict__ cnt, int32_t* __restrict__ last,
                            int32_t* __restrict__ pending, int32_t* __restrict__ order,
                            int4* __restrict__ plan, int* __restrict__ err_flag)
{
    __shared__ int s_warp[33];
    const int t = threadIdx.x;
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
            order[n_copies + arank++] = i;
        }
        else
        {
            pl = make_int4(ok ? loc[par] : -1, free_stack[f0 - 1 - crank], ok ? par : 0, 0);
            order[crank++] = i;
        }
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

/* ------------------------------------------------------------------------------------------ */
/* raster: the triangles of one staged chunk into the inverse-depth tile covering rows
 * [row0,row1] x columns [col0,col1].
 *
 * A warp takes 32 triangles at a time (batches handed out by a shared-memory counter); every
 * lane sets its own triangle up (convergent).  The pixel boxes of the 32 triangles are cut into
 * blocks of 8 columns x 4 rows and the blocks are dealt out evenly over the lanes (prefix sum
 * over the block counts, owner table in shared memory, owner's parameters by shuffle), so every
 * lane always has a block to work on however unequal the triangles are.  A block is processed
 * row by row without any branch: 8 x (three int32 edge functions, sign test, one fma for the
 * inverse depth, address select, shared-memory max on the bits of the positive float).
 * Pixels of a block outside the triangle's own box fail the edge test by themselves, so no
 * box mask is needed -- unless the box was cut by the raster window; those triangles, the ones
 * that need int64 edge functions and the ones whose box dwarfs the batch are instead walked by
 * the whole warp, 8 x 4 pixels per step with explicit bounds.                                  */
#define DBOT_OWNER_CAP 512 /* blocks per batch: 32 lanes x at most 16 blocks each */

__device__ __forceinline__ void dbot_zmax(uint32_t* p, uint32_t v)
{
    if (v > *p) atomicMax(p, v); /* plain read first: fragments that lose the test cost no atomic */
}
__device__ __forceinline__ void dbot_raster_chunk(const uint16_t* __restrict__ s_tri, int ntri,
                                                  const float4* __restrict__ s_v4,
                                                  const int* __restrict__ s_Y, const dbot_camera& K,
                                                  uint32_t* s_wt, unsigned char* s_owner, int* s_batch,
                                                  int row0, int row1, int col0, int col1, int lane)
{
    const int ww = col1 - col0 + 1;
    const uint32_t tile_s = smem_u32(s_wt);
    /* dummy words: 8 consecutive ones per lane (the column offset stays an immediate of the atomic),
     * staggered so that the 32 lanes of a warp hit 32 different banks at every column; one set serves
     * all warps -- a warp instruction owns the shared-memory pipe while it executes */
    __shared__ uint32_t s_dummy[8 * 31 + 7 + 8];
    const uint32_t dummy_s = smem_u32(&s_dummy[8 * lane + (lane >> 2)]);
    for (;;)
    {
        int base = 0;
        if (lane == 0) base = atomicAdd(s_batch, 32);
        base = __shfl_sync(DBOT_FULL, base, 0);
        if (base >= ntri) break;
        const int t = base + lane;
        dbot_tri s;
        bool ok = false;
        if (t < ntri)
        {
            const int ia = s_tri[3 * t], ib = s_tri[3 * t + 1], ic = s_tri[3 * t + 2];
            const float4 pa = s_v4[ia], pb = s_v4[ib], pc = s_v4[ic];
            dbot_vertex a, b, c;
            a.x = pa.x; a.y = pa.y; a.z = pa.z; a.X = __float_as_int(pa.w); a.Y = s_Y[ia];
            b.x = pb.x; b.y = pb.y; b.z = pb.z; b.X = __float_as_int(pb.w); b.Y = s_Y[ib];
            c.x = pc.x; c.y = pc.y; c.z = pc.z; c.X = __float_as_int(pc.w); c.Y = s_Y[ic];
            ok = dbot_setup_triangle(a, b, c, K, row0, row1, col0, col1, s);
        }
        __syncwarp();
        const int bw = ok ? (s.cmax - s.cmin + 8) >> 3 : 0; /* blocks of 8 columns */
        const int bh = ok ? (s.rmax - s.rmin + 4) >> 2 : 0; /* blocks of 4 rows */
        int tot = bw * bh;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(DBOT_FULL, tot, off);
        /* whole-warp walk: int64 edge functions, window-cut boxes, > 4 x the batch mean */
        const bool big = ok && (!s.small || s.clipped || bw * bh > min(16, max(8, tot >> 3)));
        const bool mine = ok && !big;
        const int nb = mine ? bw * bh : 0;

        /* edge functions at (cmin, rmin), oriented so that inside means all three >= 0 */
        int e0c = 0, e1c = 0, e2c = 0, e0r = 0, e1r = 0, e2r = 0, A0 = -1, A1 = -1, A2 = -1;
        if (mine)
        {
            e0c = -(s.Yb - s.Ya) * DBOT_SUBPIX; e0r = (s.Xb - s.Xa) * DBOT_SUBPIX;
            e1c = -(s.Yc - s.Yb) * DBOT_SUBPIX; e1r = (s.Xc - s.Xb) * DBOT_SUBPIX;
            e2c = -(s.Ya - s.Yc) * DBOT_SUBPIX; e2r = (s.Xa - s.Xc) * DBOT_SUBPIX;
            const int px0 = s.cmin * DBOT_SUBPIX, py0 = s.rmin * DBOT_SUBPIX;
            A0 = (s.Xb - s.Xa) * (py0 - s.Ya) - (s.Yb - s.Ya) * (px0 - s.Xa);
            A1 = (s.Xc - s.Xb) * (py0 - s.Yb) - (s.Yc - s.Yb) * (px0 - s.Xb);
            A2 = (s.Xa - s.Xc) * (py0 - s.Yc) - (s.Ya - s.Yc) * (px0 - s.Xc);
            if (s.flip)
            {
                A0 = -A0; A1 = -A1; A2 = -A2;
                e0c = -e0c; e1c = -e1c; e2c = -e2c;
                e0r = -e0r; e1r = -e1r; e2r = -e2r;
            }
        }
        /* origin of the box and its width in blocks, packed: cmin | rmin << 12 | bw << 24 */
        const int origin = mine ? (s.cmin | (s.rmin << 12) | (bw << 24)) : 0;
        /* inclusive prefix sum of the block counts, owner table */
        int P = nb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const int v = __shfl_up_sync(DBOT_FULL, P, d);
            if (lane >= d) P += v;
        }
#ifdef DBOT_DBG_NO_SWEEP
        const int H = (A0 == 0x7654321) ? 1 : 0;
#else
        const int H = __shfl_sync(DBOT_FULL, P, 31);
#endif
        const int Pex = P - nb;
        for (int k = 0; k < nb; ++k) s_owner[Pex + k] = (unsigned char)lane;
        __syncwarp();

        for (int qb = 0; qb < H; qb += 32)
        {
            const int q = qb + lane;
            const bool active = q < H;
            const int j = active ? s_owner[q] : 0;
            const int idx = q - __shfl_sync(DBOT_FULL, Pex, j);
            const int o_origin = __shfl_sync(DBOT_FULL, origin, j);
            int o_e0c = __shfl_sync(DBOT_FULL, e0c, j), o_e1c = __shfl_sync(DBOT_FULL, e1c, j),
                o_e2c = __shfl_sync(DBOT_FULL, e2c, j);
            int o_e0r = __shfl_sync(DBOT_FULL, e0r, j);
            const int o_e1r = __shfl_sync(DBOT_FULL, e1r, j), o_e2r = __shfl_sync(DBOT_FULL, e2r, j);
            const int o_A0 = __shfl_sync(DBOT_FULL, A0, j), o_A1 = __shfl_sync(DBOT_FULL, A1, j),
                      o_A2 = __shfl_sync(DBOT_FULL, A2, j);
            const float o_m0 = __shfl_sync(DBOT_FULL, s.m0, j), o_m1 = __shfl_sync(DBOT_FULL, s.m1, j),
                        o_m2 = __shfl_sync(DBOT_FULL, s.m2, j);
            const float o_wlo = __shfl_sync(DBOT_FULL, s.wlo, j), o_whi = __shfl_sync(DBOT_FULL, s.whi, j);
            const int o_bw = o_origin >> 24;
            /* idx -> (row block, column block); exact for these small integers */
            const int pr = (o_bw == 1) ? idx : __float2int_rz(__fdividef((float)idx + 0.5f, (float)o_bw));
            const int bx = idx - pr * o_bw;
            const int cx0 = (o_origin & 0xfff) + 8 * bx, ry0 = ((o_origin >> 12) & 0xfff) + 4 * pr;
            int x0 = o_A0 + (8 * bx) * o_e0c + (4 * pr) * o_e0r;
            int x1 = o_A1 + (8 * bx) * o_e1c + (4 * pr) * o_e1r;
            int x2 = o_A2 + (8 * bx) * o_e2c + (4 * pr) * o_e2r;
            if (!active) /* nothing to do in the last, partial round: every test fails */
            {
                x0 = -1;
                o_e0c = 0;
                o_e0r = 0;
            }
            const float fc0 = (float)cx0;
            float fr = (float)ry0;
            uint32_t trow = tile_s + 4u * (uint32_t)((ry0 - row0) * ww + (cx0 - col0));
            float fc[8];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) fc[cc] = dbot_add(fc0, (float)cc);
#pragma unroll
            for (int rr = 0; rr < 4; ++rr)
            {
                const float wr = dbot_fmaf(o_m1, fr, o_m2);
                /* Branch-free resolve.  ptxas wraps every predicated shared-memory atomic in a
                 * BSSY/BRA/BSYNC triple, and a read-compare-branch in front of it costs more issue
                 * slots than the atomics it saves; so every covered pixel sends its inverse depth
                 * to the tile's max, and every pixel outside the triangle sends it to a per-lane
                 * dummy word instead (own bank: no conflict, no branch). */
#define DBOT_SWEEP_PIXEL(cc)                                                                          \
    {                                                                                                  \
        const int eor = (x0 + (cc) * o_e0c) | (x1 + (cc) * o_e1c) | (x2 + (cc) * o_e2c);              \
        const uint32_t bits = dbot_f2u(fminf(fmaxf(dbot_fmaf(o_m0, fc[cc], wr), o_wlo), o_whi));       \
        const uint32_t a = eor >= 0 ? trow : dummy_s;                                                   \
        asm volatile("red.shared.max.u32 [%0+%2], %1;" ::"r"(a), "r"(bits), "n"(4 * (cc)) : "memory"); \
    }
                DBOT_SWEEP_PIXEL(0) DBOT_SWEEP_PIXEL(1) DBOT_SWEEP_PIXEL(2) DBOT_SWEEP_PIXEL(3)
                DBOT_SWEEP_PIXEL(4) DBOT_SWEEP_PIXEL(5) DBOT_SWEEP_PIXEL(6) DBOT_SWEEP_PIXEL(7)
#undef DBOT_SWEEP_PIXEL
                x0 += o_e0r; x1 += o_e1r; x2 += o_e2r;
                fr = dbot_add(fr, 1.0f);
                trow += 4u * (uint32_t)ww;
            }
        }
        __syncwarp();

        unsigned m = __ballot_sync(DBOT_FULL, big);
        while (m)
        {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            dbot_tri q;
            q.Xa = __shfl_sync(DBOT_FULL, s.Xa, j); q.Ya = __shfl_sync(DBOT_FULL, s.Ya, j);
            q.Xb = __shfl_sync(DBOT_FULL, s.Xb, j); q.Yb = __shfl_sync(DBOT_FULL, s.Yb, j);
            q.Xc = __shfl_sync(DBOT_FULL, s.Xc, j); q.Yc = __shfl_sync(DBOT_FULL, s.Yc, j);
            q.cmin = __shfl_sync(DBOT_FULL, s.cmin, j); q.cmax = __shfl_sync(DBOT_FULL, s.cmax, j);
            q.rmin = __shfl_sync(DBOT_FULL, s.rmin, j); q.rmax = __shfl_sync(DBOT_FULL, s.rmax, j);
            q.flip = __shfl_sync(DBOT_FULL, s.flip, j);
            q.small = 0;
            q.clipped = 0;
            q.m0 = __shfl_sync(DBOT_FULL, s.m0, j); q.m1 = __shfl_sync(DBOT_FULL, s.m1, j);
            q.m2 = __shfl_sync(DBOT_FULL, s.m2, j);
            q.wlo = __shfl_sync(DBOT_FULL, s.wlo, j); q.whi = __shfl_sync(DBOT_FULL, s.whi, j);
            const int dc = lane & 7, dr = lane >> 3;
            for (int r0 = q.rmin; r0 <= q.rmax; r0 += 4)
            {
                const int r = r0 + dr;
                for (int c0 = q.cmin; c0 <= q.cmax; c0 += 8)
                {
                    const int c = c0 + dc;
                    if (r <= q.rmax && c <= q.cmax && dbot_covers(q, c, r))
                        dbot_zmax(s_wt + (r - row0) * ww + (c - col0),
                                  dbot_fragment_w(q, (float)c, (float)r));
                }
            }
        }
        __syncwarp();
    }
}

/* ------------------------------------------------------------------------------------------ */
/* K2: fused transform + raster + likelihood (+ occlusion stream) kernel.
 * Warp roles: warps [0, DBOT_CW) compute; in update launches warps [DBOT_CW, DBOT_CW+DBOT_SW)
 * stream the parent's propagated occlusion image into the child's slot while the compute warps
 * transform and rasterise the same particle; the compute warps then overwrite the covered
 * pixels with the posterior.  Hand-off through two named barriers (copied / consumed). */
template <int MODE>
__global__ void __launch_bounds__(DBOT_CT + (MODE == DBOT_MODE_UPDATE ? DBOT_ST : 0), 1)
    dbot_particle_kernel(const __grid_constant__ dbot_kparams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar;      /* TMA mesh staging */
    __shared__ int s_bb[4];
    __shared__ int s_batch;                      /* next triangle batch of the current raster pass */
    __shared__ unsigned char s_owner_all[DBOT_CW][DBOT_OWNER_CAP]; /* block -> owning lane, per warp */
    __shared__ double s_red[DBOT_CW];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
    {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    /* ================= streaming warps (update launches only) =================
     * A child that does not adopt its parent's slot (see dbot_plan_update_kernel) starts as a bit
     * copy of the parent's stored image; propagation is applied lazily (dbot_b200_sensor::A). */
    if (MODE == DBOT_MODE_UPDATE && warp >= DBOT_CW)
    {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DBOT_STREAM_REGS));
        const int sl = tid - DBOT_CT;
        const int n8 = (int)(p.slot_stride >> 3);
        /* The loads below keep only DBOT_STREAM_UNROLL x 32 B per thread in flight -- far too
         * little for HBM latency.  One lane therefore runs a TMA prefetch (cp.async.bulk.prefetch.L2)
         * DBOT_PREFETCH_AHEAD chunks ahead of the loop, across particle boundaries, so that the
         * loads hit L2. */
        constexpr int CHUNK8 = DBOT_ST * DBOT_STREAM_UNROLL;       /* 256-bit words per chunk */
        const int n_chunks = (n8 + CHUNK8 - 1) / CHUNK8;
        auto next_copy = [&](int k) { /* source of the next copying child of this CTA */
            for (k += gridDim.x; k < p.n; k += gridDim.x)
            {
                const int4 pl = p.plan[p.order[k]];
                if (pl.x >= 0 && pl.x != pl.y) return pl.x;
            }
            return -1;
        };
        auto prefetch_chunk = [&](const float* src, int c) {
            const int q0 = c * CHUNK8;
            const int nq = min(CHUNK8, n8 - q0);
            if (nq > 0)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + ((size_t)q0 << 3)),
                             "r"(nq * 32)
                             : "memory");
        };
        bool first = true;
        for (int k = blockIdx.x; k < p.n; k += gridDim.x)
        {
            const int4 pl = p.plan[p.order[k]];
            if (pl.x < 0) continue;
            if (pl.x != pl.y)
            {
                const float* in = p.pool + (size_t)pl.x * p.slot_stride;
                float* out = p.pool + (size_t)pl.y * p.slot_stride;
                const float* in_next = nullptr;
                if (sl == 0)
                {
                    const int snext = next_copy(k);
                    if (snext >= 0) in_next = p.pool + (size_t)snext * p.slot_stride;
                    if (first)
                        for (int c = 0; c < DBOT_PREFETCH_AHEAD; ++c) prefetch_chunk(in, c);
                }
                first = false;
#ifdef DBOT_DBG_NO_STREAM
                if (p.n == 0x7654321)
#endif
                for (int c = 0; c < n_chunks; ++c)
                {
                    if (sl == 0)
                    {
                        const int cp = c + DBOT_PREFETCH_AHEAD;
                        if (cp < n_chunks)
                            prefetch_chunk(in, cp);
                        else if (in_next)
                            prefetch_chunk(in_next, cp - n_chunks);
                    }
                    const int base = c * CHUNK8 + sl;
                    float v[DBOT_STREAM_UNROLL][8];
#pragma unroll
                    for (int u = 0; u < DBOT_STREAM_UNROLL; ++u)
                    {
                        const int q = base + u * DBOT_ST;
                        if (q < n8) ldg256(in + ((size_t)q << 3), v[u]);
                    }
#pragma unroll
                    for (int u = 0; u < DBOT_STREAM_UNROLL; ++u)
                    {
                        const int q = base + u * DBOT_ST;
                        if (q < n8) stg256_stream(out + ((size_t)q << 3), v[u]);
                    }
                }
            }
            role_arrive(DBOT_BAR_COPIED);
            /* at most one particle ahead: wait until the compute warps have taken this copy */
            role_sync(DBOT_BAR_CONSUMED);
        }
        return;
    }

    /* ================= compute warps ================= */
    if (MODE == DBOT_MODE_UPDATE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(DBOT_COMPUTE_REGS));
    uint16_t* s_tri = (uint16_t*)(smem + p.off_tri);
    float4* s_v4 = (float4*)(smem + p.off_v4);
    int* s_Y = (int*)(smem + p.off_Y);
    uint32_t* s_wt = (uint32_t*)(smem + p.off_wt);
    uint32_t parity = 0;
    int staged = -1;

    /* TMA-stage the triangle list of one mesh chunk: one thread arms the mbarrier with the
     * byte count and issues the bulk copy; every compute thread waits on the barrier phase. */
    auto stage_chunk = [&](int c) {
        if (staged == c) return;
        compute_sync(); /* all readers of the previous chunk are done */
        if (tid == 0)
        {
            const dbot_chunk ch = p.chunks[c];
            mbar_expect_tx(&s_bar, ch.tri_bytes);
            tma_load_1d(s_tri, p.mesh + ch.tri_off, ch.tri_bytes, &s_bar);
        }
        mbar_wait(&s_bar, parity);
        parity ^= 1u;
        staged = c;
    };
    /* v' = R v + t, projection and snapping of one chunk's vertices into shared memory */
    auto transform_chunk = [&](const dbot_chunk& ch, const float* pose_g, int& mnx, int& mxx,
                               int& mny, int& mxy) {
        float pose[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) pose[e] = __ldg(pose_g + e);
        const float4* vg = reinterpret_cast<const float4*>(p.mesh + ch.vert_off);
        for (int v = tid; v < ch.n_verts; v += DBOT_CT)
        {
            const float4 o = __ldg(vg + v);
            const dbot_vertex t = dbot_transform_vertex(o.x, o.y, o.z, pose, p.K);
            s_v4[v] = make_float4(t.x, t.y, t.z, __int_as_float(t.X));
            s_Y[v] = t.Y;
            if (t.X != DBOT_VERT_INVALID)
            {
                mnx = min(mnx, t.X); mxx = max(mxx, t.X);
                mny = min(mny, t.Y); mxy = max(mxy, t.Y);
            }
        }
    };

    for (int k = blockIdx.x; k < p.n; k += gridDim.x)
    {
        /* update launches walk the particles in the planned order (copying children first) */
        const int i = (MODE == DBOT_MODE_UPDATE) ? p.order[k] : k;
        const float* pose = p.poses + (size_t)i * 12 * p.n_obj;
        int4 pl = make_int4(0, 0, 0, 0);
        if (MODE == DBOT_MODE_UPDATE)
            pl = p.plan[i];
        else if (MODE == DBOT_MODE_EVAL)
        {
            const int slot = p.indices[i];
            pl.x = (slot < 0 || slot >= p.n_slots) ? -1 : p.loc[slot];
        }
        if (pl.x < 0)
        {
            if (tid == 0)
            {
                p.out[i] = __longlong_as_double(0x7ff8000000000000LL);
                atomicExch(p.err_flag, 1);
            }
            continue;
        }
        /* stored occlusion values are read from the parent's slot; an adopting sibling leaves it
         * alone until this particle has signed off (pending count, end of the particle) */
        const float* occ_in = p.pool + (size_t)pl.x * p.slot_stride;
        float* occ_out = (MODE == DBOT_MODE_UPDATE) ? p.pool + (size_t)pl.y * p.slot_stride : nullptr;

#ifdef DBOT_DBG_TIMELINE
        long long tl[8];
        int tli = 0;
#define DBOT_TL() do { if (tli < 8) tl[tli++] = clock64(); } while (0)
#else
#define DBOT_TL() do { } while (0)
#endif
        DBOT_TL();
        /* ---- pass 1: transform every chunk, collect the screen window of the particle ---- */
        if (tid == 0)
        {
            s_bb[0] = INT_MAX; s_bb[1] = INT_MIN; s_bb[2] = INT_MAX; s_bb[3] = INT_MIN;
        }
        compute_sync(); /* also fences the previous particle's use of shared memory */
        int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
        for (int c = 0; c < p.n_chunks; ++c)
        {
            if (c > 0) compute_sync(); /* multi-chunk: only the bounding box survives this pass */
            const dbot_chunk ch = p.chunks[c];
            transform_chunk(ch, pose + 12 * ch.obj, mnx, mxx, mny, mxy);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
        {
            mnx = min(mnx, __shfl_xor_sync(DBOT_FULL, mnx, off));
            mxx = max(mxx, __shfl_xor_sync(DBOT_FULL, mxx, off));
            mny = min(mny, __shfl_xor_sync(DBOT_FULL, mny, off));
            mxy = max(mxy, __shfl_xor_sync(DBOT_FULL, mxy, off));
        }
        if (lane == 0)
        {
            atomicMin(&s_bb[0], mnx); atomicMax(&s_bb[1], mxx);
            atomicMin(&s_bb[2], mny); atomicMax(&s_bb[3], mxy);
        }
        compute_sync();
        int c0 = 0, c1 = -1, r0 = 0, r1 = -1;
        if (s_bb[0] <= s_bb[1])
        {
            c0 = max(dbot_ceil_subpix(s_bb[0]), 0);
            c1 = min(dbot_floor_subpix(s_bb[1]), p.cols - 1);
            r0 = max(dbot_ceil_subpix(s_bb[2]), 0);
            r1 = min(dbot_floor_subpix(s_bb[3]), p.rows - 1);
        }
        DBOT_TL(); /* transform + window done */
        const bool have_window = (c0 <= c1) && (r0 <= r1);
        const int ww = c1 - c0 + 1;
        float acc = 0.0f;
        bool copied = false; /* update: this particle's slot is ready for the posterior */
        auto await_slot = [&]() {
            /* a copying child: its slot holds the parent's image once the streaming warps arrive.
             * An adopting child owns the parent's slot itself and waits until all its siblings are
             * done reading it (they sign off at the end of their particle). */
            role_sync(DBOT_BAR_COPIED);
            role_arrive(DBOT_BAR_CONSUMED);
            if (pl.x == pl.y)
            {
                if (tid == 0)
                {
                    while (atomicAdd(p.pending + pl.z, 0) > 0) __nanosleep(200);
                    __threadfence();
                }
                compute_sync();
            }
            copied = true;
        };

        if (have_window)
        {
            const int band_rows = max(1, (int)(p.wcap / (uint32_t)ww));
            for (int rb0 = r0; rb0 <= r1; rb0 += band_rows)
            {
                const int rb1 = min(rb0 + band_rows - 1, r1);
                const int npx = (rb1 - rb0 + 1) * ww;
                compute_sync(); /* previous band's readers of the tile are done */
                {
                    uint4* w4 = reinterpret_cast<uint4*>(s_wt);
                    const int n4 = (npx + 3) >> 2; /* wcap is a multiple of 4 */
                    for (int q = tid; q < n4; q += DBOT_CT) w4[q] = make_uint4(0u, 0u, 0u, 0u);
                }
                if (p.n_chunks == 1)
                {
                    stage_chunk(0); /* resident after the first particle */
                    if (tid == 0) s_batch = 0;
                    compute_sync();
                    DBOT_TL(); /* tile cleared */
                    dbot_raster_chunk(s_tri, p.chunks[0].n_tris, s_v4, s_Y, p.K, s_wt, s_owner_all[warp],
                                      &s_batch, rb0, rb1, c0, c1, lane);
                    DBOT_TL(); /* this warp's raster done */
                }
                else
                {
                    for (int c = 0; c < p.n_chunks; ++c)
                    {
                        stage_chunk(c); /* leading barrier orders tile init / previous raster */
                        const dbot_chunk ch = p.chunks[c];
                        int d0 = INT_MAX, d1 = INT_MIN, d2 = INT_MAX, d3 = INT_MIN;
                        transform_chunk(ch, pose + 12 * ch.obj, d0, d1, d2, d3);
                        if (tid == 0) s_batch = 0;
                        compute_sync();
                        dbot_raster_chunk(s_tri, ch.n_tris, s_v4, s_Y, p.K, s_wt, s_owner_all[warp],
                                          &s_batch, rb0, rb1, c0, c1, lane);
                    }
                }
                compute_sync();
                if (MODE == DBOT_MODE_UPDATE && !copied) await_slot();
                /* ---- per-pixel mixture over the band: a warp per row, DBOT_UNIT_PX x 32 columns
                 * (and their loads) in flight per step ---- */
                if (MODE == DBOT_MODE_RENDER)
                {
                    for (int r = rb0 + warp; r <= rb1; r += DBOT_CW)
                        for (int cc = lane; cc < ww; cc += 32)
                            p.depth_out[(size_t)r * p.cols + c0 + cc] =
                                dbot_depth_from_w(s_wt[(r - rb0) * ww + cc]);
                }
                else
                {
                    struct unit
                    {
                        uint32_t wb[DBOT_UNIT_PX];
                        float4 tab[DBOT_UNIT_PX];
                        float o_in[DBOT_UNIT_PX];
                        size_t g; /* element offset of the unit's first column in the state / table */
                    };
                    auto fetch = [&](int r, int cb, unit& u) {
                        const uint32_t* wrow = s_wt + (r - rb0) * ww + cb + lane;
                        u.g = (size_t)r * p.pitch + c0 + cb + lane;
#pragma unroll
                        for (int j = 0; j < DBOT_UNIT_PX; ++j)
                            u.wb[j] = (cb + 32 * j + lane < ww) ? wrow[32 * j] : DBOT_W_EMPTY;
#pragma unroll
                        for (int j = 0; j < DBOT_UNIT_PX; ++j)
                            if (u.wb[j] != DBOT_W_EMPTY)
                            {
                                u.tab[j] = __ldg(p.tab + u.g + 32 * j);
                                /* update launches may read a slot whose covered pixels were written
                                 * in an earlier launch phase by its adopter-to-be: L2-coherent, not .nc */
                                u.o_in[j] = (MODE == DBOT_MODE_UPDATE) ? __ldcg(occ_in + u.g + 32 * j)
                                                                       : __ldg(occ_in + u.g + 32 * j);
                            }
                    };
                    auto evaluate = [&](const unit& u) {
#pragma unroll
                        for (int j = 0; j < DBOT_UNIT_PX; ++j)
                            if (u.wb[j] != DBOT_W_EMPTY && u.tab[j].x == u.tab[j].x)
                            {
                                float o_new;
                                /* the correctly rounded reciprocal, not MUFU alone: the oracle's
                                 * (and dbot's) predicted depth is the float the renderer stores, and
                                 * d - y cancels -- one ulp of d is 1e-5 in the Gaussian's argument */
                                acc = dbot_add(acc, dbot_pixel_term(dbot_rcp(dbot_u2f(u.wb[j])), u.tab[j],
                                                                    u.o_in[j], p.M, o_new));
                                if (MODE == DBOT_MODE_UPDATE)
                                    __stcs(occ_out + u.g + 32 * j, dbot_mul(dbot_sub(o_new, p.post_b), p.post_inv_a));
                            }
                    };
#ifdef DBOT_DBG_NO_PIXELS
                    if (p.n == 0x7654321)
#endif
                    for (int r = rb0 + warp; r <= rb1; r += DBOT_CW)
                        for (int cb = 0; cb < ww; cb += 32 * DBOT_UNIT_PX)
                        {
                            unit ua;
                            fetch(r, cb, ua);
                            evaluate(ua);
                        }
                }
            }
        }
        if (MODE == DBOT_MODE_UPDATE && !copied) await_slot();
        DBOT_TL(); /* this warp's mixture done */

        /* ---- one log-weight per particle: fixed shuffle tree, double across warps ---- */
        if (MODE != DBOT_MODE_RENDER)
        {
            double v = (double)acc;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(DBOT_FULL, v, off);
            if (lane == 0) s_red[warp] = v;
            compute_sync();
            if (tid == 0)
            {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < DBOT_CW; ++w) s += s_red[w];
                p.out[i] = s;
#ifdef DBOT_DBG_TIMELINE
                DBOT_TL();
                if (blockIdx.x == 3 && k >= 3 * (int)gridDim.x && k < 5 * (int)gridDim.x)
                    printf("mode %d particle %d cycles: transform+window %lld clear %lld raster(w0) %lld "
                           "mixture(w0, incl. wait for other warps' raster) %lld reduce %lld total %lld\n",
                           MODE, i, tl[1] - tl[0], tl[2] - tl[1], tl[3] - tl[2], tl[4] - tl[3], tl[5] - tl[4],
                           tl[5] - tl[0]);
#endif
                if (MODE == DBOT_MODE_UPDATE)
                {
                    p.indices[i] = i;
                    if (pl.x != pl.y) /* done with the parent's slot: its adopter may write into it */
                    {
                        __threadfence();
                        atomicSub(p.pending + pl.z, 1);
                    }
                }
            }
        }
    }
}

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
    int4* d_plan;
    float* d_poses;
    int32_t* d_indices;
    double* d_out;
    float* d_depth;
    int* d_err;
    /* pinned staging */
    float* h_poses;
    int32_t* h_indices;
    double* h_out;
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
};

static uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

/* split every object mesh into chunks of at most DBOT_CHUNK_MAX_VERTS vertices and
 * DBOT_CHUNK_MAX_TRIS triangles with chunk-local uint16 indices (shared vertices are
 * duplicated; a duplicated vertex transforms to bit-identical coordinates).  Vertices are
 * stored as float4 (read once per particle straight from L2), triangle lists as uint16
 * triples (TMA-staged into shared memory). */
static int build_chunks(const dbot_b200_config* cfg, std::vector<dbot_chunk>& chunks,
                        std::vector<unsigned char>& blob)
{
    for (int o = 0; o < cfg->n_objects; ++o)
    {
        const int nv = cfg->n_vertices[o], nt = cfg->n_triangles[o];
        if (nv < 0 || nt < 0) return fail(-2, "object %d: negative vertex/triangle count", o);
        const float* V = cfg->vertices[o];
        const int32_t* T = cfg->triangles[o];
        if ((nv > 0 && !V) || (nt > 0 && !T)) return fail(-1, "object %d: null mesh pointer", o);
        for (int i = 0; i < 3 * nt; ++i)
            if (T[i] < 0 || T[i] >= nv) return fail(-6, "object %d: triangle index out of range", o);
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
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    cudaFree(s->d_mesh); cudaFree(s->d_chunks); cudaFree(s->d_obs); cudaFree(s->d_tab);
    cudaFree(s->d_pool); cudaFree(s->d_loc[0]); cudaFree(s->d_loc[1]); cudaFree(s->d_free);
    cudaFree(s->d_free_count); cudaFree(s->d_cnt); cudaFree(s->d_last); cudaFree(s->d_pending);
    cudaFree(s->d_order); cudaFree(s->d_plan); cudaFree(s->d_poses); cudaFree(s->d_indices);
    cudaFree(s->d_out); cudaFree(s->d_depth); cudaFree(s->d_err);
    cudaFreeHost(s->h_poses); cudaFreeHost(s->h_indices); cudaFreeHost(s->h_out);
    cudaFreeHost(s->h_obs); cudaFreeHost(s->h_depth); cudaFreeHost(s->h_err);
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
    if (cfg->device_id < 0 || cfg->device_id >= n_dev)
        return fail(-2, "device_id %d out of range (have %d)", cfg->device_id, n_dev);
    s->device = cfg->device_id;
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
    int rc = build_chunks(cfg, chunks, blob);
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
    kp.off_Y = off;   off += round_up((uint32_t)max_nv * 4u, 128);
    kp.off_wt = off;
    /* static __shared__ (owner tables, dummy words, hand-off) + slack */
    const int smem_max = (int)prop.sharedMemPerBlockOptin - 1024 - DBOT_CW * DBOT_OWNER_CAP - DBOT_CW * 128;
    if ((int)off + 4 * (int)round_up((uint32_t)kp.cols, 4) > smem_max)
        return fail(-9, "image too wide for the shared-memory depth tile (cols=%d)", kp.cols);
    s->smem_bytes = smem_max;
    kp.wcap = ((uint32_t)(smem_max - (int)off) / 16u) * 4u; /* pixels, multiple of 4 */
    CUDA_TRY(cudaFuncSetAttribute(dbot_particle_kernel<DBOT_MODE_EVAL>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, s->smem_bytes));
    CUDA_TRY(cudaFuncSetAttribute(dbot_particle_kernel<DBOT_MODE_UPDATE>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, s->smem_bytes));
    CUDA_TRY(cudaFuncSetAttribute(dbot_particle_kernel<DBOT_MODE_RENDER>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, s->smem_bytes));
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
    CUDA_TRY(cudaMalloc(&s->d_plan, np * sizeof(int4)));
    CUDA_TRY(cudaMalloc(&s->d_poses, np * 12 * kp.n_obj * sizeof(float)));
    CUDA_TRY(cudaMalloc(&s->d_indices, np * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&s->d_out, np * sizeof(double)));
    CUDA_TRY(cudaMalloc(&s->d_depth, P * sizeof(float)));
    CUDA_TRY(cudaMalloc(&s->d_err, sizeof(int)));
    CUDA_TRY(cudaMemset(s->d_err, 0, sizeof(int)));
    s->device_bytes += (int64_t)(2 * P * sizeof(float) + kp.slot_stride * sizeof(float4) +
                                 2 * state_bytes + np * (12 * kp.n_obj * 4 + 4 + 8 + 8 * 4 + 16) + 8);
    CUDA_TRY(cudaHostAlloc(&s->h_poses, np * 12 * kp.n_obj * sizeof(float), cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc(&s->h_indices, np * sizeof(int32_t), cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc(&s->h_out, np * sizeof(double), cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc(&s->h_obs, P * sizeof(float), cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc(&s->h_depth, P * sizeof(float), cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc(&s->h_err, sizeof(int), cudaHostAllocDefault));
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
    dbot_b200_sensor* s = new dbot_b200_sensor();
    memset(s, 0, sizeof(*s));
    s->cfg = *cfg;
    s->cfg.vertices = nullptr; s->cfg.triangles = nullptr;
    s->cfg.n_vertices = nullptr; s->cfg.n_triangles = nullptr;
    int rc = create_impl(cfg, s);
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
__global__ void dbot_reset_slots_kernel(int32_t* loc, int32_t* free_stack, int32_t* free_count, int n_phys)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0)
    {
        loc[0] = 0;
        *free_count = n_phys - 1;
    }
    if (i < n_phys - 1) free_stack[i] = n_phys - 1 - i; /* top of the stack = slot 1 */
}

extern "C" int dbot_b200_reset(dbot_b200_sensor* s)
{
    if (!s) return fail(-1, "reset: null handle");
    CUDA_TRY(cudaSetDevice(s->device));
    s->cur = 0;
    s->n_slots = 1;
    s->t_obs = 0.0;
    s->t_update = 0.0;
    s->A = 1.0;
    s->B = 0.0;
    const int n_phys = 2 * s->cfg.max_particles;
    dbot_fill_kernel<<<64, 256, 0, s->stream>>>(s->d_pool, s->kp.slot_stride,
                                                (float)s->cfg.initial_occlusion_prob);
    dbot_reset_slots_kernel<<<(n_phys + 255) / 256, 256, 0, s->stream>>>(s->d_loc[0], s->d_free,
                                                                         s->d_free_count, n_phys);
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
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t bytes = (size_t)s->kp.rows * s->kp.cols * sizeof(float);
    if (is_pinned_host(depth_m))
    {
        /* the caller's buffer is page-locked: copy straight from it, and be done with it on return */
        CUDA_TRY(cudaMemcpyAsync(s->d_obs, depth_m, bytes, cudaMemcpyHostToDevice, s->stream));
        int rc = launch_observation(s, s->d_obs);
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        return 0;
    }
    /* pageable: stage in the handle's own pinned buffer; the copy and the table kernel are left
     * running -- every later call on this handle is ordered behind them on the same stream, and the
     * staging buffer is not touched again before they are done */
    if (s->obs_in_flight) CUDA_TRY(cudaEventSynchronize(s->ev_obs));
    memcpy(s->h_obs, depth_m, bytes);
    CUDA_TRY(cudaMemcpyAsync(s->d_obs, s->h_obs, bytes, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaEventRecord(s->ev_obs, s->stream));
    s->obs_in_flight = 1;
    return launch_observation(s, s->d_obs);
}

extern "C" int dbot_b200_set_observation_device(dbot_b200_sensor* s, const float* d_depth_m)
{
    if (!s || !d_depth_m) return fail(-1, "set_observation_device: null argument");
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
        dbot_rebase_kernel<<<dim3(32, std::min(s->n_slots, 4096)), 256, 0, s->stream>>>(
            s->d_pool, s->d_loc[s->cur], s->n_slots, s->kp.slot_stride, (float)s->A, (float)s->B);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
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
        dbot_plan_update_kernel<<<1, DBOT_PLAN_THREADS, 0, s->stream>>>(
            d_indices, n, s->n_slots, s->d_loc[s->cur], s->d_loc[s->cur ^ 1], s->d_free, s->d_free_count,
            s->d_cnt, s->d_last, s->d_pending, s->d_order, s->d_plan, s->d_err);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    const int grid = std::min(n, s->grid_max);
    prof_begin(s);
    if (update)
    {
        /* adopting children spin on their siblings' copies: every CTA must be resident */
        void* args[] = {(void*)&kp};
        CUDA_TRY(cudaLaunchCooperativeKernel((const void*)dbot_particle_kernel<DBOT_MODE_UPDATE>, dim3(grid),
                                             dim3(DBOT_CT + DBOT_ST), args, (size_t)s->smem_bytes, s->stream));
    }
    else
        dbot_particle_kernel<DBOT_MODE_EVAL><<<grid, DBOT_CT, s->smem_bytes, s->stream>>>(kp);
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

/* RbSensor::loglikes */
extern "C" int dbot_b200_loglikes(dbot_b200_sensor* s, const float* poses, int32_t* indices,
                                  int32_t n, int32_t update, double* out_loglikes)
{
    if (!s || !poses || !indices || !out_loglikes) return fail(-1, "loglikes: null argument");
    if (n < 0 || n > s->cfg.max_particles)
        return fail(-2, "loglikes: n=%d exceeds max_particles=%d", n, s->cfg.max_particles);
    for (int i = 0; i < n; ++i)
        if (indices[i] < 0 || indices[i] >= s->n_slots)
            return fail(-7, "loglikes: indices[%d]=%d is not an occlusion slot (have %d)", i,
                        indices[i], s->n_slots);
    if (n == 0) return 0;
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t pose_bytes = (size_t)n * 12 * s->kp.n_obj * sizeof(float);
    memcpy(s->h_poses, poses, pose_bytes);
    memcpy(s->h_indices, indices, (size_t)n * sizeof(int32_t));
    CUDA_TRY(cudaMemcpyAsync(s->d_poses, s->h_poses, pose_bytes, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->d_indices, s->h_indices, (size_t)n * sizeof(int32_t),
                             cudaMemcpyHostToDevice, s->stream));
    int rc = launch_particles(s, s->d_poses, s->d_indices, n, update, s->d_out);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(s->h_out, s->d_out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost,
                             s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    memcpy(out_loglikes, s->h_out, (size_t)n * sizeof(double));
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
    if (n == 0) return 0;
    CUDA_TRY(cudaSetDevice(s->device));
    return launch_particles(s, d_poses, d_indices, n, update, d_out_loglikes);
}

extern "C" void* dbot_b200_stream(dbot_b200_sensor* s) { return s ? (void*)s->stream : nullptr; }

extern "C" int dbot_b200_synchronize(dbot_b200_sensor* s)
{
    if (!s) return fail(-1, "synchronize: null handle");
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaMemcpyAsync(s->h_err, s->d_err, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (*s->h_err)
    {
        const int code = *s->h_err;
        cudaMemsetAsync(s->d_err, 0, sizeof(int), s->stream);
        if (code == 2) return fail(-10, "the occlusion-slot pool is exhausted (too many imported slots)");
        return fail(-7, "loglikes_device: an index was not a valid occlusion slot");
    }
    return 0;
}

/* RigidBodyRenderer::Render for one pose set */
extern "C" int dbot_b200_render(dbot_b200_sensor* s, const float* pose, float* out_depth)
{
    if (!s || !pose || !out_depth) return fail(-1, "render: null argument");
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
        dbot_particle_kernel<DBOT_MODE_RENDER><<<1, DBOT_CT, s->smem_bytes, s->stream>>>(kp);
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
extern "C" int dbot_b200_n_slots(dbot_b200_sensor* s) { return s ? s->n_slots : -1; }
extern "C" double dbot_b200_observation_time(dbot_b200_sensor* s) { return s ? s->t_obs : NAN; }

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
                                       int* err_flag)
{
    const int f = *free_count;
    if (f <= 0)
    {
        atomicExch(err_flag, 2);
        loc[slot] = loc[0];
        return;
    }
    loc[slot] = free_stack[f - 1];
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
    /* live + imported slots stay within half of the pool, so that an update can always find a fresh
     * slot for every copying child (every imported slot is some child's parent and is adopted) */
    if (s->n_slots >= 2 * s->cfg.max_particles - 1)
        return fail(-2, "import_slot: %d slots are live, the pool holds %d", s->n_slots, 2 * s->cfg.max_particles);
    CUDA_TRY(cudaSetDevice(s->device));
    const int slot = s->n_slots;
    dbot_import_pop_kernel<<<1, 1, 0, s->stream>>>(s->d_loc[s->cur], s->d_free, s->d_free_count, slot, s->d_err);
    dbot_import_copy_kernel<<<64, 256, 0, s->stream>>>(s->d_pool, s->d_loc[s->cur], slot, s->kp.slot_stride, d_src);
    s->launches += 2;
    CUDA_TRY(cudaGetLastError());
    s->n_slots = slot + 1;
    *out_slot = slot;
    return 0;
}

extern "C" int dbot_b200_free_slot_count(dbot_b200_sensor* s)
{
    if (!s) return -1;
    if (cudaSetDevice(s->device) != cudaSuccess || cudaStreamSynchronize(s->stream) != cudaSuccess) return -1;
    int32_t f = -1;
    if (cudaMemcpy(&f, s->d_free_count, sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return f;
}

extern "C" int dbot_b200_get_occlusion(dbot_b200_sensor* s, int32_t slot, float* out_occ)
{
    if (!s || !out_occ) return fail(-1, "get_occlusion: null argument");
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
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaEventRecord(s->ev_begin, s->stream));
    return 0;
}
extern "C" int dbot_b200_timer_end(dbot_b200_sensor* s, float* out_ms)
{
    if (!s || !out_ms) return fail(-1, "timer_end: null argument");
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaEventRecord(s->ev_end, s->stream));
    CUDA_TRY(cudaEventSynchronize(s->ev_end));
    CUDA_TRY(cudaEventElapsedTime(out_ms, s->ev_begin, s->ev_end));
    return 0;
}
extern "C" int dbot_b200_profile_enable(dbot_b200_sensor* s, int32_t on)
{
    if (!s) return fail(-1, "profile_enable: null handle");
    s->profile = on ? 1 : 0;
    memset(s->prof_ms, 0, sizeof(s->prof_ms));
    memset(s->prof_launches, 0, sizeof(s->prof_launches));
    return 0;
}
extern "C" int dbot_b200_profile_read(dbot_b200_sensor* s, float* out_ms, int64_t* out_launches)
{
    if (!s || !out_ms || !out_launches) return fail(-1, "profile_read: null argument");
    memcpy(out_ms, s->prof_ms, sizeof(s->prof_ms));
    memcpy(out_launches, s->prof_launches, sizeof(s->prof_launches));
    memset(s->prof_ms, 0, sizeof(s->prof_ms));
    memset(s->prof_launches, 0, sizeof(s->prof_launches));
    return 0;
}
extern "C" int64_t dbot_b200_launch_count(dbot_b200_sensor* s) { return s ? s->launches : -1; }
extern "C" int64_t dbot_b200_device_bytes(dbot_b200_sensor* s) { return s ? s->device_bytes : -1; }

#include "dbot_filter.cuh"
