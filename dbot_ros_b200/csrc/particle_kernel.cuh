/*
 * particle_kernel.cuh -- the fused transform / raster / mixture (/ occlusion copy) kernel, one persistent CTA per SM
 * Part of libdbot_b200.so; included by dbot_b200.cu (one translation unit).
 */
#pragma once

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
 * that need int64 edge functions and the ones with more than DBOT_BIG_BLOCKS blocks are instead
 * walked by the whole warp, 8 x 4 pixels per step with explicit bounds.                        */
#define DBOT_BOX_CHUNKS 32 /* chunks beyond this many are never culled against a tile */
#define DBOT_OWNER_CAP 512 /* blocks per batch: 32 lanes x at most 16 blocks each */
#define DBOT_BIG_BLOCKS 16 /* a triangle with more blocks than a lane's share of the table: whole-warp walk */

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
        /* whole-warp walk: int64 edge functions, window-cut boxes, more blocks than a lane's share of the owner table */
        const bool big = ok && (!s.small || s.clipped || bw * bh > DBOT_BIG_BLOCKS);
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
/* K2: fused transform + raster + likelihood (+ occlusion copy) kernel.
 * Warp roles: warps [0, DBOT_CW) compute; in update launches warps [DBOT_CW, DBOT_CW+DBOT_SW)
 * copy the parent's stored occlusion image into the child's slot -- for the children that do not
 * adopt their parent's slot (dbot_plan_update_kernel) -- while the compute warps transform and
 * rasterise the same particle; the compute warps then overwrite the covered pixels with the
 * (lazily encoded) posterior.  Hand-off through two named barriers (copied / consumed); an
 * adopting child additionally waits on its parent's pending counter. */
template <int MODE, bool MULTI> /* MULTI: the mesh is split into several chunks (big or multi-object) */
__global__ void __launch_bounds__(DBOT_CT + (MODE == DBOT_MODE_UPDATE ? DBOT_ST : 0), 1)
    dbot_particle_kernel(const __grid_constant__ dbot_kparams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar;      /* TMA mesh staging */
    __shared__ int s_bb[4];
    __shared__ int s_cbb[DBOT_BOX_CHUNKS][4]; /* snapped box of each mesh chunk (multi-chunk meshes) */
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
        if (MULTI && tid < 4 * DBOT_BOX_CHUNKS) s_cbb[tid >> 2][tid & 3] = (tid & 1) ? INT_MIN : INT_MAX;
        compute_sync(); /* also fences the previous particle's use of shared memory */
        int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
        for (int c = 0; c < p.n_chunks; ++c)
        {
            if (c > 0) compute_sync(); /* multi-chunk: only the bounding boxes survive this pass */
            const dbot_chunk ch = p.chunks[c];
            int a0 = INT_MAX, a1 = INT_MIN, a2 = INT_MAX, a3 = INT_MIN;
            transform_chunk(ch, pose + 12 * ch.obj, a0, a1, a2, a3);
            mnx = min(mnx, a0); mxx = max(mxx, a1); mny = min(mny, a2); mxy = max(mxy, a3);
            if (MULTI && c < DBOT_BOX_CHUNKS)
            {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1)
                {
                    a0 = min(a0, __shfl_xor_sync(DBOT_FULL, a0, off));
                    a1 = max(a1, __shfl_xor_sync(DBOT_FULL, a1, off));
                    a2 = min(a2, __shfl_xor_sync(DBOT_FULL, a2, off));
                    a3 = max(a3, __shfl_xor_sync(DBOT_FULL, a3, off));
                }
                if (lane == 0)
                {
                    atomicMin(&s_cbb[c][0], a0); atomicMax(&s_cbb[c][1], a1);
                    atomicMin(&s_cbb[c][2], a2); atomicMax(&s_cbb[c][3], a3);
                }
            }
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
                    /* siblings never wait for anything, so this ends; the bound (about a second) only
                     * turns a scheduling accident into an error code instead of a hung GPU */
                    int spins = 0;
                    while (atomicAdd(p.pending + pl.z, 0) > 0)
                    {
                        __nanosleep(200);
                        if (++spins > (1 << 22))
                        {
                            atomicExch(p.err_flag, 3);
                            break;
                        }
                    }
                    __threadfence();
                }
                compute_sync();
            }
            copied = true;
        };

        if (have_window)
        {
            /* The window is cut into tiles that fit the shared-memory z tile: row bands of the full
             * width, or -- when the window is wider than tall, e.g. several objects side by side --
             * column bands of the full height, so that a tile meets as few mesh chunks as possible. */
            const int wc0 = c0, wc1 = c1, wh = r1 - r0 + 1;
            int tile_rows = max(1, (int)(p.wcap / (uint32_t)ww)), tile_cols = ww;
            if (MULTI && tile_rows < wh && ww > wh && (uint32_t)wh * 32u <= p.wcap)
            {
                tile_rows = wh;
                tile_cols = (int)(p.wcap / (uint32_t)wh);
            }
            for (int rb0 = r0; rb0 <= r1; rb0 += tile_rows)
            for (int cb0 = wc0; cb0 <= wc1; cb0 += tile_cols)
            {
                const int rb1 = min(rb0 + tile_rows - 1, r1);
                const int c0 = cb0, c1 = min(cb0 + tile_cols - 1, wc1), ww = c1 - c0 + 1; /* this tile */
                const int npx = (rb1 - rb0 + 1) * ww;
                compute_sync(); /* previous band's readers of the tile are done */
                {
                    uint4* w4 = reinterpret_cast<uint4*>(s_wt);
                    const int n4 = (npx + 3) >> 2; /* wcap is a multiple of 4 */
                    for (int q = tid; q < n4; q += DBOT_CT) w4[q] = make_uint4(0u, 0u, 0u, 0u);
                }
                if (!MULTI)
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
                        if (c < DBOT_BOX_CHUNKS) /* a chunk whose box misses the tile is not even staged */
                        {
                            const int b0 = s_cbb[c][0], b1 = s_cbb[c][1], b2 = s_cbb[c][2], b3 = s_cbb[c][3];
                            if (b0 > b1 || dbot_ceil_subpix(b0) > c1 || dbot_floor_subpix(b1) < c0 ||
                                dbot_ceil_subpix(b2) > rb1 || dbot_floor_subpix(b3) < rb0)
                                continue;
                        }
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

