/*
 * particle_kernel.cuh -- the fused transform / raster / mixture (/ occlusion copy) kernel, one persistent CTA per SM
 * Part of libdbot_b200.so; included by dbot_b200.cu (one translation unit).
 */
#pragma once

/* ------------------------------------------------------------------------------------------ */
/* raster: the triangles of one staged chunk into the inverse-depth tile covering rows
 * [row0,row1] x columns [col0,col1].  Three steps, all in shared memory:
 *
 * 1. prepass (whole CTA, one triangle per thread and step): snapped vertices -> validity,
 *    orientation (back faces of a mesh proven closed are dropped here and never set up), pixel box
 *    clipped to the tile.  Survivors are appended to a compact list with one shared-memory atomic
 *    per warp (ballot + popc); the few triangles that do not fit a batch lane (more than
 *    DBOT_SPAN_BATCH_ROWS rows, or int64 edge functions) go to the back of the same list.
 * 2. set-up (a warp takes 32 survivors at a time, one per lane): plane equation bit for bit as the
 *    oracle, the three edges sorted into long / short as dbot_setup_spans describes; the span record
 *    stays in the lane's registers.
 * 3. units: the pixel rows of the 32 triangles are dealt out evenly over the lanes -- unit q of the
 *    batch is row q - first_row(owner) of its owner, the owner being found from a 1024-bit mask of
 *    first rows (one word per lane) by two shuffles and a popc.  A unit fetches its owner's record
 *    by shuffle, computes the exact first and last covered column of its row
 *    (dbot_row_span) and sends only those pixels' inverse depths to the tile's max
 *    (red.shared.max on the bits of the positive float).  No pixel outside a triangle is ever
 *    tested: 35 k fragments and 15 k row units per particle for the 5120-triangle mesh, where the
 *    block sweep of round 1 tested 385 k pixels.                                              */
#define DBOT_BOX_CHUNKS 32 /* chunks beyond this many are never culled against a tile */
#define DBOT_LIST_WIDE 0x8000u /* flag in the survivor list: int64 block walk instead of span units */

__device__ __forceinline__ void dbot_zmax(uint32_t* p, uint32_t v)
{
    if (v > *p) atomicMax(p, v); /* plain read first: fragments that lose the test cost no atomic */
}

struct dbot_raster_smem
{
    const uint16_t* tri;  /* [ntri][3] chunk-local vertex indices (TMA-staged) */
    const float4* v4;     /* [nv] camera-frame x, y, z and 1/z */
    const int2* XY;       /* [nv] snapped image position */
    uint16_t* list;       /* [ntri] survivors from the front, tall / wide triangles from the back */
    uint32_t* mask;       /* this warp's first-row mask, 32 words */
    uint32_t* wt;         /* inverse-depth tile */
    int* counters;        /* [0] survivors | tall + wide << 16, [2] next batch */
};

__device__ __forceinline__ dbot_vertex dbot_load_vertex(const dbot_raster_smem& S, int i)
{
    const float4 a = S.v4[i];
    const int2 b = S.XY[i];
    dbot_vertex v;
    v.x = a.x; v.y = a.y; v.z = a.z; v.iz = a.w; v.X = b.x; v.Y = b.y;
    return v;
}

/* step 1; every compute thread calls it, then a compute_sync() publishes the list */
__device__ __forceinline__ void dbot_raster_prepass(const dbot_raster_smem& S, int ntri, int cull_sign,
                                                    int row0, int row1, int col0, int col1, int tid)
{
    const int lane = tid & 31;
    const unsigned lt = (1u << lane) - 1u;
    /* two triangles per thread and step (a warp takes 64 consecutive ones): two independent chains of
     * index -> vertex -> box loads in flight, one shared atomic for both */
    for (int base = 2 * (tid - lane); base < ntri; base += 2 * DBOT_CT)
    {
        int kind[2];
#pragma unroll
        for (int u = 0; u < 2; ++u)
        {
            const int t = base + 32 * u + lane;
            kind[u] = DBOT_TRI_REJECT;
            if (t < ntri)
            {
                const int2 a = S.XY[S.tri[3 * t]], b = S.XY[S.tri[3 * t + 1]], c = S.XY[S.tri[3 * t + 2]];
                int cmin, cmax, rmin, rmax;
                kind[u] = dbot_prepass_triangle(a.x, a.y, b.x, b.y, c.x, c.y, cull_sign, row0, row1, col0, col1,
                                                cmin, cmax, rmin, rmax);
            }
        }
        const unsigned f0 = __ballot_sync(DBOT_FULL, kind[0] == DBOT_TRI_SPANS);
        const unsigned f1 = __ballot_sync(DBOT_FULL, kind[1] == DBOT_TRI_SPANS);
        const unsigned b0 = __ballot_sync(DBOT_FULL, kind[0] >= DBOT_TRI_TALL);
        const unsigned b1 = __ballot_sync(DBOT_FULL, kind[1] >= DBOT_TRI_TALL);
        int base_fb = 0;
        if (lane == 0 && (f0 | f1 | b0 | b1))
            base_fb = atomicAdd(&S.counters[0], (__popc(f0) + __popc(f1)) | ((__popc(b0) + __popc(b1)) << 16));
        __syncwarp();
        base_fb = __shfl_sync(DBOT_FULL, base_fb, 0);
        const int front = base_fb & 0xffff, back = base_fb >> 16;
#pragma unroll
        for (int u = 0; u < 2; ++u)
        {
            const int t = base + 32 * u + lane;
            const unsigned mf = u ? f1 : f0, mb = u ? b1 : b0;
            const int pf = (u ? __popc(f0) : 0) + __popc(mf & lt), pb = (u ? __popc(b0) : 0) + __popc(mb & lt);
            if (kind[u] == DBOT_TRI_SPANS)
                S.list[front + pf] = (uint16_t)t;
            else if (kind[u] >= DBOT_TRI_TALL)
                S.list[ntri - 1 - (back + pb)] = (uint16_t)(t | (kind[u] == DBOT_TRI_WIDE ? DBOT_LIST_WIDE : 0u));
        }
    }
}

/* one (triangle, row) unit: row j of the box of the triangle whose span record the lane `owner` holds in its
 * registers (s); every field the unit needs comes over by shuffle -- the records never touch shared memory,
 * which leaves it to the z tile (a 5120-triangle particle's whole window fits one tile) */
__device__ __forceinline__ void dbot_span_unit(const dbot_span_rec& s, int owner, int j, int aux, bool active,
                                               uint32_t tile_s, uint32_t dummy_s, int row0, int col0, int ww)
{
    const int bwm1 = aux & 0xff, split = (aux >> 8) & 0xff;
    const bool long_left = (aux >> 16) & 1, lower = j >= split;
    /* long edge, and the short edge that bounds this row on the other side */
    const int a0 = __shfl_sync(DBOT_FULL, s.A.g0, owner), a1 = __shfl_sync(DBOT_FULL, s.A.gr, owner),
              a2 = __shfl_sync(DBOT_FULL, s.A.gc, owner);
    const int u0 = __shfl_sync(DBOT_FULL, s.B1.g0, owner), u1 = __shfl_sync(DBOT_FULL, s.B1.gr, owner),
              u2 = __shfl_sync(DBOT_FULL, s.B1.gc, owner);
    const int v0 = __shfl_sync(DBOT_FULL, s.B2.g0, owner), v1 = __shfl_sync(DBOT_FULL, s.B2.gr, owner),
              v2 = __shfl_sync(DBOT_FULL, s.B2.gc, owner);
    const float m0 = __shfl_sync(DBOT_FULL, s.m0, owner), m1 = __shfl_sync(DBOT_FULL, s.m1, owner),
                m2 = __shfl_sync(DBOT_FULL, s.m2, owner);
    const float wlo = __shfl_sync(DBOT_FULL, s.wlo, owner), whi = __shfl_sync(DBOT_FULL, s.whi, owner);
    const int org = __shfl_sync(DBOT_FULL, s.org, owner);
    const int b0 = lower ? v0 : u0, b1 = lower ? v1 : u1, b2 = lower ? v2 : u2;
    dbot_edge L, R;
    L.g0 = long_left ? a0 : b0; L.gr = long_left ? a1 : b1; L.gc = long_left ? a2 : b2;
    R.g0 = long_left ? b0 : a0; R.gr = long_left ? b1 : a1; R.gc = long_left ? b2 : a2;
    L.nrc = -dbot_rcp_fast((float)L.gc); /* estimates only: dbot_row_span corrects them exactly */
    R.nrc = -dbot_rcp_fast((float)R.gc);
    int first;
    int n = dbot_row_span(L, R, j, bwm1, first);
    if (!active || n < 0) n = 0;
    const int r = (org >> 16) + j, c = (org & 0xffff) + first;
    const float wr = dbot_fmaf(m1, (float)r, m2);
    float fc = (float)c;
    uint32_t addr = tile_s + 4u * (uint32_t)((r - row0) * ww + (c - col0));
    /* The pixel loop runs the same number of times in every lane (the longest span of the round): a lane
     * whose span is done sends its inverse depths to its own dummy word instead.  A data-dependent trip count
     * would make this a divergent loop around an atomic, which the compiler must treat as a potential spin
     * lock (YIELD in the loop: lanes that finish early are released from the reconvergence point and run the
     * rest of the round as separate groups -- measured: half the lanes idle through the whole unit loop). */
    const int nmax = __reduce_max_sync(DBOT_FULL, n);
    for (int k = 0; k < nmax; ++k)
    {
        const uint32_t bits = dbot_f2u(fminf(fmaxf(dbot_fmaf(m0, fc, wr), wlo), whi));
        const uint32_t a = k < n ? addr : dummy_s;
        asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(a), "r"(bits) : "memory");
        fc = dbot_add(fc, 1.0f);
        addr += 4u;
    }
}

/* steps 2 and 3; called by every compute warp after the prepass has been published */
__device__ __forceinline__ void dbot_raster_sweep(const dbot_raster_smem& S, int ntri, const dbot_camera& K,
                                                  int row0, int row1, int col0, int col1, int lane)
{
    const int ww = col1 - col0 + 1;
    const uint32_t tile_s = smem_u32(S.wt);
    /* per-lane dummy words for finished spans: the warp's first-row mask, which is dead while the units run
     * and rewritten at the start of the next batch (one word per lane: no bank conflict) */
    const uint32_t dummy_s = smem_u32(S.mask + lane);
    const int n_front = S.counters[0] & 0xffff, n_back = S.counters[0] >> 16;
    const int front_batches = (n_front + 31) >> 5;
    const unsigned le = 0xffffffffu >> (31 - lane);
    for (;;)
    {
        int batch = 0;
        if (lane == 0) batch = atomicAdd(&S.counters[2], 1);
        __syncwarp();
        batch = __shfl_sync(DBOT_FULL, batch, 0);
        if (batch >= front_batches + n_back) break;

        /* ---- which triangles: 32 survivors, or one tall / wide triangle from the back ---- */
        int t = -1;
        bool wide = false;
        if (batch < front_batches)
        {
            const int k = 32 * batch + lane;
            if (k < n_front) t = S.list[k];
        }
        else
        {
            const unsigned e = S.list[ntri - 1 - (batch - front_batches)];
            wide = (e & DBOT_LIST_WIDE) != 0;
            if (lane == 0 || wide) t = (int)(e & (DBOT_LIST_WIDE - 1u));
        }
        if (wide)
        {
            /* int64 edge functions: the whole warp walks the box, 8 x 4 pixels per step */
            const dbot_vertex a = dbot_load_vertex(S, S.tri[3 * t]), b = dbot_load_vertex(S, S.tri[3 * t + 1]),
                              c = dbot_load_vertex(S, S.tri[3 * t + 2]);
            dbot_tri q;
            if (dbot_setup_triangle(a, b, c, K, row0, row1, col0, col1, q))
            {
                const int dc = lane & 7, dr = lane >> 3;
                for (int r0 = q.rmin; r0 <= q.rmax; r0 += 4)
                {
                    const int r = r0 + dr;
                    for (int c0 = q.cmin; c0 <= q.cmax; c0 += 8)
                    {
                        const int cc = c0 + dc;
                        if (r <= q.rmax && cc <= q.cmax && dbot_covers(q, cc, r))
                            dbot_zmax(S.wt + (r - row0) * ww + (cc - col0), dbot_fragment_w(q, (float)cc, (float)r));
                    }
                }
            }
            __syncwarp();
            continue;
        }

        /* ---- set-up: one triangle per lane -> span record in the lane's registers ---- */
        int nrows = 0, aux = 0;
        dbot_span_rec s;
        memset(&s, 0, sizeof(s));
        if (t >= 0)
        {
            const dbot_vertex a = dbot_load_vertex(S, S.tri[3 * t]), b = dbot_load_vertex(S, S.tri[3 * t + 1]),
                              c = dbot_load_vertex(S, S.tri[3 * t + 2]);
            int cmin = dbot_ceil_subpix(dbot_min3(a.X, b.X, c.X)), cmax = dbot_floor_subpix(dbot_max3(a.X, b.X, c.X));
            int rmin = dbot_ceil_subpix(dbot_min3(a.Y, b.Y, c.Y)), rmax = dbot_floor_subpix(dbot_max3(a.Y, b.Y, c.Y));
            cmin = max(cmin, col0); cmax = min(cmax, col1);
            rmin = max(rmin, row0); rmax = min(rmax, row1);
            dbot_setup_spans(a, b, c, K, cmin, cmax, rmin, rmax, s);
            nrows = rmax - rmin + 1;
            aux = s.aux;
        }
        /* first unit of every lane's triangle (exclusive prefix sum of the row counts) */
        int P = nrows;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const int v = __shfl_up_sync(DBOT_FULL, P, d);
            if (lane >= d) P += v;
        }
        const int H = __shfl_sync(DBOT_FULL, P, 31);
        const int Pex = P - nrows;
        /* bit q of the mask: a triangle starts at unit q */
        S.mask[lane] = 0u;
        __syncwarp();
        if (nrows > 0) atomicOr(&S.mask[Pex >> 5], 1u << (Pex & 31));
        __syncwarp();
        const uint32_t word = S.mask[lane];
        int wpre = __popc(word);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const int v = __shfl_up_sync(DBOT_FULL, wpre, d);
            if (lane >= d) wpre += v;
        }
        wpre -= __popc(word);
        const int pa = Pex | (aux << 11);

        /* ---- units ---- */
        for (int qb = 0; qb < H; qb += 32)
        {
            const int q = qb + lane;
            const uint32_t w = __shfl_sync(DBOT_FULL, word, qb >> 5);
            const int pre = __shfl_sync(DBOT_FULL, wpre, qb >> 5);
            /* triangles hold their lanes in order and every one of them has at least one row, so the
             * owner of unit q is the number of first-row bits at or below q, minus one */
            const int owner = min(pre + __popc(w & le) - 1, 31);
            const int opa = __shfl_sync(DBOT_FULL, pa, owner);
            dbot_span_unit(s, owner, q - (opa & 0x7ff), opa >> 11, q < H, tile_s, dummy_s, row0, col0, ww);
        }
        __syncwarp(); /* the dummy words (first-row mask) are rewritten by the next batch */
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
    __shared__ int s_cnt[4];                     /* raster pass: survivors, tall + wide, next batch */
    __shared__ uint32_t s_mask[DBOT_CW][32];     /* first-row masks of the span batches, per warp */
    __shared__ double s_red[DBOT_CW];
    __shared__ int s_copied;                     /* update: copies this CTA's streaming warps have finished */
    __shared__ int s_gave_up;                    /* update: a wait for a slot ran into its bound (error 3) */

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
    {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        s_copied = 0;
        s_gave_up = 0;
        s_cnt[0] = 0; s_cnt[1] = 0; s_cnt[2] = 0;
    }
    __syncthreads();

    /* ================= streaming warps (update launches only) =================
     * A child that does not adopt its parent's slot (see dbot_plan_update_kernel) starts as a bit
     * copy of the parent's stored image; propagation is applied lazily (dbot_b200_sensor::A).
     * The stream runs free: every destination is a fresh slot nobody else touches, so the copies of
     * this CTA's later particles start as soon as the earlier ones are done, however far the compute
     * warps are behind -- the HBM traffic of the launch is spread over its whole duration instead of
     * being demanded particle by particle.  After each copy the stream (a) counts it in s_copied, which
     * is what the compute warps of that particle wait for before they touch the child's slot, and
     * (b) signs the parent off (pending): the compute warps read the child's own copy, never the
     * parent, so a sibling that adopts the parent's slot is held up by copies only. */
    if (MODE == DBOT_MODE_UPDATE && warp >= DBOT_CW)
    {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DBOT_STREAM_REGS));
        const int sl = tid - DBOT_CT;
        const int row8 = p.pitch >> 3; /* 256-bit words per image row */
        /* Only the rectangle in which the parent's image can differ from the background is copied
         * (dbot_kparams::slot_rect: rows, and columns in 256-bit words): every image is the background value
         * outside its recorded rectangle, the recycled destination slot included -- whose own old rectangle is
         * first put back to the background where the parent's does not cover it.  A tracked object touches a
         * patch of the image, not all of it; when it has wandered everywhere this is the dense copy again.
         * The copy walks the rectangle as one flat sequence of 256-bit words (row and word back from the flat
         * index by a multiply-high).  The loads keep only DBOT_STREAM_UNROLL x 32 B per thread in flight -- far
         * too little for HBM latency -- so the group's first warp runs a TMA prefetch
         * (cp.async.bulk.prefetch.L2, one row segment per lane) DBOT_PREFETCH_AHEAD chunks ahead of the loop,
         * across particle boundaries, and the loads hit L2. */
        constexpr int CHUNK8 = DBOT_ST * DBOT_STREAM_UNROLL; /* 256-bit words per chunk */
        struct patch /* a rectangle of one slot as a flat word sequence */
        {
            const float* base; /* the slot */
            int r0, w0, nrows, wcols, n8;
            uint32_t inv; /* floor(2^32 / wcols) + 1 */
        };
        auto make_patch = [&](int phys, const int4& r) {
            patch c;
            c.base = p.pool + (size_t)phys * p.slot_stride;
            c.r0 = r.x; c.w0 = r.z;
            c.nrows = r.x <= r.y ? r.y - r.x + 1 : 0;
            c.wcols = r.z <= r.w ? r.w - r.z + 1 : 0;
            if (c.nrows == 0 || c.wcols == 0) c.nrows = c.wcols = 0;
            c.n8 = c.nrows * c.wcols;
            c.inv = c.wcols > 1 ? 0xffffffffu / (uint32_t)c.wcols + 1u : 0u;
            return c;
        };
        auto word_of = [&](const patch& c, int q) -> uint32_t { /* flat word q of the patch -> 256-bit word of the slot */
            const uint32_t row = c.wcols > 1 ? __umulhi((uint32_t)q, c.inv) : (uint32_t)q;
            const uint32_t w = (uint32_t)q - row * (uint32_t)c.wcols;
            return (uint32_t)(c.r0 + (int)row) * (uint32_t)row8 + (uint32_t)(c.w0 + (int)w);
        };
        auto next_copy = [&](int k) { /* source patch of this CTA's next copy */
            for (k += gridDim.x; k < p.n; k += gridDim.x)
            {
                const int4 pl = p.plan[p.order[k]];
                if (pl.x >= 0 && pl.x != pl.y) return make_patch(pl.x, p.slot_rect[pl.x]);
            }
            return make_patch(0, make_int4(0, -1, 0, -1));
        };
        auto prefetch_chunk = [&](const patch& c, int chunk) { /* called by the group's first warp */
            const int q0 = chunk * CHUNK8;
            if (q0 >= c.n8) return;
            const int q1 = min(q0 + CHUNK8, c.n8) - 1;
            const int ra = c.wcols > 1 ? (int)__umulhi((uint32_t)q0, c.inv) : q0;
            const int rb = c.wcols > 1 ? (int)__umulhi((uint32_t)q1, c.inv) : q1;
            for (int r = ra + sl; r <= rb; r += 32)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(
                                 c.base + (((size_t)(c.r0 + r) * (size_t)row8 + (size_t)c.w0) << 3)),
                             "r"(c.wcols * 32)
                             : "memory");
        };
        auto fill_rect = [&](float* slot, int ra, int rb, int wa, int wb) { /* rows x words of a slot <- background */
            if (ra > rb || wa > wb) return;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = p.background;
            const patch c = make_patch(0, make_int4(ra, rb, wa, wb));
            for (int q = sl; q < c.n8; q += DBOT_ST) stg256_stream(slot + ((size_t)word_of(c, q) << 3), v);
        };
        bool first = true;
        int done = 0;
        for (int k = blockIdx.x; k < p.n; k += gridDim.x)
        {
            const int4 pl = p.plan[p.order[k]];
            if (pl.x < 0) continue;
            if (pl.x != pl.y)
            {
                const int4 pr = p.slot_rect[pl.x];
                const int4 dr = p.slot_rect[pl.y];
                const patch src = make_patch(pl.x, pr);
                float* slot = p.pool + (size_t)pl.y * p.slot_stride;
                const int n_chunks = (src.n8 + CHUNK8 - 1) / CHUNK8;
                patch nxt = make_patch(0, make_int4(0, -1, 0, -1));
                if (sl < 32)
                {
                    nxt = next_copy(k);
                    if (first)
                        for (int c = 0; c < DBOT_PREFETCH_AHEAD; ++c) prefetch_chunk(src, c);
                }
                first = false;
                /* what the recycled slot held outside the parent's rectangle */
                if (src.n8 > 0)
                {
                    fill_rect(slot, dr.x, min(dr.y, pr.x - 1), dr.z, dr.w);                      /* rows above */
                    fill_rect(slot, max(dr.x, pr.y + 1), dr.y, dr.z, dr.w);                      /* rows below */
                    fill_rect(slot, max(dr.x, pr.x), min(dr.y, pr.y), dr.z, min(dr.w, pr.z - 1)); /* left of it */
                    fill_rect(slot, max(dr.x, pr.x), min(dr.y, pr.y), max(dr.z, pr.w + 1), dr.w); /* right of it */
                }
                else
                    fill_rect(slot, dr.x, dr.y, dr.z, dr.w);
                for (int c = 0; c < n_chunks; ++c)
                {
                    if (sl < 32)
                    {
                        const int cp = c + DBOT_PREFETCH_AHEAD;
                        if (cp < n_chunks)
                            prefetch_chunk(src, cp);
                        else
                            prefetch_chunk(nxt, cp - n_chunks);
                    }
                    const int base = c * CHUNK8 + sl;
                    float v[DBOT_STREAM_UNROLL][8];
                    uint32_t off[DBOT_STREAM_UNROLL];
#pragma unroll
                    for (int u = 0; u < DBOT_STREAM_UNROLL; ++u)
                    {
                        const int q = base + u * DBOT_ST;
                        off[u] = word_of(src, min(q, src.n8 - 1));
                        if (q < src.n8) ldg256(src.base + ((size_t)off[u] << 3), v[u]);
                    }
#pragma unroll
                    for (int u = 0; u < DBOT_STREAM_UNROLL; ++u)
                    {
                        const int q = base + u * DBOT_ST;
                        if (q < src.n8) stg256_stream(slot + ((size_t)off[u] << 3), v[u]);
                    }
                }
                __threadfence(); /* this thread's part of the copy is visible device-wide */
                stream_sync();   /* ... and so is everybody else's */
                if (sl == 0)
                {
                    /* the child's image differs from the background where the parent's does */
                    p.slot_rect[pl.y] = src.n8 > 0 ? pr : make_int4(p.rows, -1, row8, -1);
                    __threadfence(); /* in L2 before this particle's compute warps reduce their window into it */
                    atomicSub(p.pending + pl.z, 1); /* one reader of the parent's slot less */
                    ++done;
                    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(&s_copied)), "r"(done)
                                 : "memory");
                }
            }
        }
        return;
    }

    /* ================= compute warps ================= */
    if (MODE == DBOT_MODE_UPDATE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(DBOT_COMPUTE_REGS));
    uint16_t* s_tri = (uint16_t*)(smem + p.off_tri);
    float4* s_v4 = (float4*)(smem + p.off_v4);
    int2* s_XY = (int2*)(smem + p.off_XY);
    uint32_t* s_wt = (uint32_t*)(smem + p.off_wt);
    dbot_raster_smem S;
    S.tri = s_tri;
    S.v4 = s_v4;
    S.XY = s_XY;
    S.list = (uint16_t*)(smem + p.off_list);
    S.mask = s_mask[warp];
    S.wt = s_wt;
    S.counters = s_cnt;
    uint32_t parity = 0;
    int staged = -1;
    /* the tile starts empty and every mixture pass hands it back empty (it clears what it reads) */
    {
        uint4* w4 = reinterpret_cast<uint4*>(s_wt);
        for (int q = tid; q < (int)(p.wcap >> 2); q += DBOT_CT) w4[q] = make_uint4(0u, 0u, 0u, 0u);
    }

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
        /* four vertices per thread and step, all four loads in flight before the first is used */
        for (int v0 = tid; v0 < ch.n_verts; v0 += 4 * DBOT_CT)
        {
            float4 o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
            {
                const int v = v0 + u * DBOT_CT;
                o[u] = __ldg(vg + (v < ch.n_verts ? v : v0));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
            {
                const int v = v0 + u * DBOT_CT;
                if (v < ch.n_verts)
                {
                    const dbot_vertex t = dbot_transform_vertex(o[u].x, o[u].y, o[u].z, pose, p.K);
                    s_v4[v] = make_float4(t.x, t.y, t.z, t.iz);
                    s_XY[v] = make_int2(t.X, t.Y);
                    if (t.X != DBOT_VERT_INVALID)
                    {
                        mnx = min(mnx, t.X); mxx = max(mxx, t.X);
                        mny = min(mny, t.Y); mxy = max(mxy, t.Y);
                    }
                }
            }
        }
    };

    int my_copies = 0; /* update: copying children this CTA's compute warps have met so far */
    for (int k = blockIdx.x; k < p.n; k += gridDim.x)
    {
        /* update launches walk the particles in the planned order */
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
                for (int g = 0; g < p.n_peers; ++g)
                    p.peer_out[g][p.out_offset + i] = __longlong_as_double(0x7ff8000000000000LL);
                atomicExch(p.err_flag, 1);
            }
            continue;
        }
        /* stored occlusion values: the parent's slot, or -- a copying child of an update -- the child's own
         * copy of it (ready once the streaming warps have counted it in s_copied) */
        float* occ_out = (MODE == DBOT_MODE_UPDATE) ? p.pool + (size_t)pl.y * p.slot_stride : nullptr;
        const float* occ_in = (MODE == DBOT_MODE_UPDATE) ? occ_out : p.pool + (size_t)pl.x * p.slot_stride;
        const bool copier = MODE == DBOT_MODE_UPDATE && pl.x != pl.y;
        const int copy_idx = copier ? my_copies++ : 0;

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
        auto await_slot = [&]() -> bool {
            /* a copying child: its slot holds the parent's image once the streaming warps have counted the
             * copy.  An adopting child owns the parent's slot itself and waits until every copy of it has
             * been taken (pending).  Neither wait depends on anything that waits itself, so both end; the
             * bound (about a second) only turns a scheduling accident into an error code instead of a hung
             * GPU -- and the particle then leaves the slot alone (a sibling may still be reading it): the CTA's
             * compute warps stop, the call fails with -11 and the handle wants a reset(). */
            if (tid == 0)
            {
                int spins = 0;
                for (;;)
                {
                    int ready;
                    if (copier)
                    {
                        int c;
                        asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(c) : "r"(smem_u32(&s_copied)) : "memory");
                        ready = c > copy_idx;
                    }
                    else
                        ready = atomicAdd(p.pending + pl.z, 0) <= 0;
                    if (ready) break;
                    __nanosleep(copier ? 64 : 200);
                    if (++spins > (1 << 22))
                    {
                        atomicExch(p.err_flag, 3);
                        s_gave_up = 1;
                        break;
                    }
                }
                __threadfence();
            }
            compute_sync();
            copied = true;
            return s_gave_up == 0;
        };

        if (have_window)
        {
            /* The window is cut into tiles that fit the shared-memory z tile: row bands of the full
             * width, or -- when the window is wider than tall, e.g. several objects side by side --
             * column bands of the full height, so that a tile meets as few mesh chunks as possible. */
            const int wc0 = c0, wc1 = c1, wh = r1 - r0 + 1;
            int tile_rows = wh, tile_cols = ww; /* the usual case: the whole window is one tile */
            if ((uint32_t)ww * (uint32_t)wh > p.wcap)
            {
                tile_rows = max(1, (int)(p.wcap / (uint32_t)ww));
                if (MULTI && tile_rows < wh && ww > wh && (uint32_t)wh * 32u <= p.wcap)
                {
                    tile_rows = wh;
                    tile_cols = (int)(p.wcap / (uint32_t)wh);
                }
                /* as few tiles as the capacity allows, of equal size: every tile costs a prepass over the mesh */
                tile_rows = (wh + (wh + tile_rows - 1) / tile_rows - 1) / ((wh + tile_rows - 1) / tile_rows);
                tile_cols = (ww + (ww + tile_cols - 1) / tile_cols - 1) / ((ww + tile_cols - 1) / tile_cols);
            }
            for (int rb0 = r0; rb0 <= r1; rb0 += tile_rows)
            for (int cb0 = wc0; cb0 <= wc1; cb0 += tile_cols)
            {
                const int rb1 = min(rb0 + tile_rows - 1, r1);
                const int c0 = cb0, c1 = min(cb0 + tile_cols - 1, wc1), ww = c1 - c0 + 1; /* this tile */
                if (!MULTI)
                {
                    stage_chunk(0); /* resident after the first particle */
                    /* the raster counters were put back to zero when the previous sweep ended (below); a
                     * particle's first tile comes after the window barrier, later ones need their own */
                    if (rb0 != r0 || cb0 != wc0) compute_sync();
                    DBOT_TL();
                    dbot_raster_prepass(S, p.chunks[0].n_tris, p.chunks[0].cull_sign, rb0, rb1, c0, c1, tid);
                    compute_sync();
                    DBOT_TL(); /* prepass done */
                    dbot_raster_sweep(S, p.chunks[0].n_tris, p.K, rb0, rb1, c0, c1, lane);
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
                        stage_chunk(c); /* leading barrier orders the previous raster / mixture pass */
                        const dbot_chunk ch = p.chunks[c];
                        int d0 = INT_MAX, d1 = INT_MIN, d2 = INT_MAX, d3 = INT_MIN;
                        transform_chunk(ch, pose + 12 * ch.obj, d0, d1, d2, d3);
                        if (tid == 0) { s_cnt[0] = 0; s_cnt[1] = 0; s_cnt[2] = 0; }
                        compute_sync();
                        dbot_raster_prepass(S, ch.n_tris, ch.cull_sign, rb0, rb1, c0, c1, tid);
                        compute_sync();
                        dbot_raster_sweep(S, ch.n_tris, p.K, rb0, rb1, c0, c1, lane);
                    }
                }
                compute_sync();
                if (tid == 0) { s_cnt[0] = 0; s_cnt[1] = 0; s_cnt[2] = 0; } /* every warp has left the sweep */
                if (MODE == DBOT_MODE_UPDATE && !copied && !await_slot())
                {
                    /* the slot never became this particle's (error 3): these warps write no more state in this
                     * launch.  The streaming warps go on -- other CTAs' adopters wait for their copies only. */
                    if (tid == 0) p.out[i] = __longlong_as_double(0x7ff8000000000000LL);
                    return;
                }
                /* ---- per-pixel mixture over the band: a warp per row, DBOT_UNIT_PX x 32 columns
                 * (and their loads) in flight per step ---- */
                if (MODE == DBOT_MODE_RENDER)
                {
                    for (int r = rb0 + warp; r <= rb1; r += DBOT_CW)
                        for (int cc = lane; cc < ww; cc += 32)
                        {
                            p.depth_out[(size_t)r * p.cols + c0 + cc] =
                                dbot_depth_from_w(s_wt[(r - rb0) * ww + cc]);
                            s_wt[(r - rb0) * ww + cc] = DBOT_W_EMPTY;
                        }
                }
                else
                {
                    /* The tile is one flat array (row-major over the tile's own width): a warp takes 32 x
                     * DBOT_UNIT_PX consecutive entries at a time, whatever rows they fall in -- no lane idles
                     * because a row ended.  Row and column come back from the flat index by a multiply-high
                     * with floor(2^32 / ww) + 1 (exact for these sizes: index * ww < 2^32; a tile one column wide is
                     * its own case, 2^32 does not fit the multiplier). */
                    struct unit
                    {
                        uint32_t wb[DBOT_UNIT_PX];
                        float4 tab[DBOT_UNIT_PX];
                        float o_in[DBOT_UNIT_PX];
                        uint32_t g[DBOT_UNIT_PX]; /* element offset of the pixel in the state / table */
                    };
                    const int npx = (rb1 - rb0 + 1) * ww;
                    const uint32_t inv_ww = ww > 1 ? 0xffffffffu / (uint32_t)ww + 1u : 0u;
                    const uint32_t g0 = (uint32_t)rb0 * (uint32_t)p.pitch + (uint32_t)c0;
                    auto fetch = [&](int base, unit& u) -> bool {
                        bool any = false;
#pragma unroll
                        for (int j = 0; j < DBOT_UNIT_PX; ++j)
                        {
                            const int q = base + 32 * j + lane;
                            u.wb[j] = q < npx ? s_wt[q] : DBOT_W_EMPTY;
                            any |= u.wb[j] != DBOT_W_EMPTY;
                        }
                        /* nothing covered in these 32 x DBOT_UNIT_PX entries (window corners, the gap between two
                         * objects): next */
                        if (!__any_sync(DBOT_FULL, any)) return false;
                        /* every lane loads its pixel's table entry and stored occlusion value, covered or not
                         * (a lane past the tile's end loads the tile's last pixel): the loads are issued for the
                         * warp anyway, and without a branch around them there is nothing to initialise and
                         * nothing to reconverge */
#pragma unroll
                        for (int j = 0; j < DBOT_UNIT_PX; ++j)
                        {
                            const int q = base + 32 * j + lane;
                            if (MODE != DBOT_MODE_UPDATE)
                            {
                                /* launches that write no state: only covered pixels are loaded (measured: 3 %
                                 * less time per eval launch than loading for every lane; update launches, whose
                                 * compute warps are short of registers, gain 6 % from the branch-free form below) */
                                const uint32_t rr = ww > 1 ? __umulhi((uint32_t)q, inv_ww) : (uint32_t)q;
                                u.g[j] = g0 + rr * (uint32_t)p.pitch + ((uint32_t)q - rr * (uint32_t)ww);
                                u.tab[j] = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
                                u.o_in[j] = 0.f;
                                if (u.wb[j] != DBOT_W_EMPTY)
                                {
                                    s_wt[q] = DBOT_W_EMPTY;
                                    u.tab[j] = __ldg(p.tab + u.g[j]);
                                    u.o_in[j] = __ldg(occ_in + u.g[j]);
                                }
                                continue;
                            }
                            const uint32_t qc = (uint32_t)min(q, npx - 1);
                            const uint32_t rr = ww > 1 ? __umulhi(qc, inv_ww) : qc;
                            u.g[j] = g0 + rr * (uint32_t)p.pitch + (qc - rr * (uint32_t)ww);
                            if (u.wb[j] != DBOT_W_EMPTY) s_wt[q] = DBOT_W_EMPTY; /* hand the tile back empty */
                            u.tab[j] = __ldg(p.tab + u.g[j]);
                            /* update launches may read a slot whose covered pixels were written
                             * in an earlier launch phase by its adopter-to-be: L2-coherent, not .nc */
                            u.o_in[j] = (MODE == DBOT_MODE_UPDATE) ? __ldcg(occ_in + u.g[j]) : __ldg(occ_in + u.g[j]);
                        }
                        return true;
                    };
                    auto evaluate = [&](const unit& u) {
                        /* straight-line code for all DBOT_UNIT_PX pixels of the unit: a lane without a covered,
                         * observed pixel computes on whatever it holds and the result is dropped by a select, so
                         * that the four pixels' dependency chains (two MUFU.EX2, two MUFU.RCP, MUFU.LG2 each)
                         * interleave instead of running one basic block after the other */
#pragma unroll
                        for (int j = 0; j < DBOT_UNIT_PX; ++j)
                        {
                            const bool valid = u.wb[j] != DBOT_W_EMPTY && u.tab[j].x == u.tab[j].x;
                            float o_new;
                            /* the correctly rounded reciprocal, not MUFU alone: the oracle's (and dbot's)
                             * predicted depth is the float the renderer stores, and d - y cancels -- one ulp
                             * of d is 1e-5 in the Gaussian's argument */
                            const float d = dbot_rcp_normal(dbot_u2f(valid ? u.wb[j] : 0x3f800000u));
                            const float term = dbot_pixel_term_log2(d, u.tab[j], u.o_in[j], p.M, o_new);
                            acc = dbot_add(acc, valid ? term : 0.0f);
                            if (MODE == DBOT_MODE_UPDATE && valid)
                                __stcs(occ_out + u.g[j], dbot_mul(dbot_sub(o_new, p.post_b), p.post_inv_a));
                        }
                    };
                    for (int base = warp * (32 * DBOT_UNIT_PX); base < npx; base += DBOT_CW * 32 * DBOT_UNIT_PX)
                    {
                        unit ua;
                        if (fetch(base, ua)) evaluate(ua);
                    }
                }
            }
        }
        DBOT_TL(); /* this warp's mixture done (a particle without a window writes nothing: no wait for its slot) */

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
                s *= 0.6931471805599453; /* the pixel terms are summed in bits */
                p.out[i] = s;
                /* fused gather (multi_device.cuh, gather_mode 1): the log-weight goes straight into every
                 * GPU's gathered vector over NVLink peer memory */
                for (int g = 0; g < p.n_peers; ++g) p.peer_out[g][p.out_offset + i] = s;
#ifdef DBOT_DBG_TIMELINE
                DBOT_TL();
                if (blockIdx.x == 3 && k >= 3 * (int)gridDim.x && k < 5 * (int)gridDim.x)
                    printf("mode %d particle %d cycles: transform+window %lld sync %lld prepass %lld raster(w0) %lld "
                           "mixture(w0, incl. wait for other warps' raster) %lld reduce %lld total %lld\n",
                           MODE, i, tl[1] - tl[0], tl[2] - tl[1], tl[3] - tl[2], tl[4] - tl[3], tl[5] - tl[4],
                           tl[6] - tl[5], tl[6] - tl[0]);
#endif
                if (MODE == DBOT_MODE_UPDATE)
                {
                    p.indices[i] = i;
                    if (have_window) /* posteriors were written inside the window: rows [r0, r1], columns [c0, c1] */
                    {
                        /* four reductions that nobody waits for (a load + store would hold this warp, and with
                         * it the next particle's first barrier, for one trip to L2) */
                        int* rect = reinterpret_cast<int*>(p.slot_rect + pl.y);
                        atomicMin(rect + 0, r0); atomicMax(rect + 1, r1);
                        atomicMin(rect + 2, c0 >> 3); atomicMax(rect + 3, c1 >> 3);
                    }
                }
            }
        }
    }
}

