/*
 * render_streams.cuh -- engine 2 of the render kernels: orbits sorted by length into three streams, one kernel each,
 * every kernel a single instruction stream at (nearly) full lanes.  Included by render_generic.cuh.
 *
 * Why.  The escape loop (reference: src/main/cuda/fractals/mandelbrot.cu:15-20) runs 1 ... maxIterations trips per
 * orbit.  Engine 1 (render_refill.cuh) keeps short and long orbits in the same warps: a warp that holds 22 long orbits
 * and churns through short ones in its other lanes runs tested blocks (6 FP64 instructions per trip plus predicates)
 * for all 32 lanes most of the time, and every orbit's end costs a share of a scheduling pass of some 500 instructions
 * (lane counters on c2: a quarter of pass A's lane-trips sit in tested blocks that are 68 % useful; 45 % of the issue
 * slots are not FP64).  Here the three kinds of work never meet:
 *
 *   probe    one warp per work item (a vote tile, or a (tile, sample) pair), lane = pixel, lock step: every orbit's
 *            first trips with a test per trip (quadratic.cuh step()).  Orbits that end -- most of a frame's orbits --
 *            deliver their result.  The survivors are appended to the LONG list (32 bytes each: where the result
 *            goes, the trips run and the state after them -- restarting an orbit from its plane point instead costs its
 *            first 64 trips a second time, a seventh of c2's executed trips).  A tile in which nobody has ended after CHAOS_PROBE_BAIL trips stops probing: deep inside or
 *            deep zoom, every pixel is long.
 *   long     persistent warps, lane refill from the long list (one atomicAdd per warp and refill, the n-th idle lane
 *            takes the n-th claimed entry).  ONLY the untested stream: groups of 32 trips of 5 FP64 instructions, one
 *            escape test and one recurrence compare per group.  An orbit leaves when (a) its state recurred bit for
 *            bit: proven never to escape, result delivered; (b) the iteration limit is reached; (c) a group's test
 *            failed or the tail is shorter than a group: the state before that group goes to the FINISH list.
 *            No lane ever replays anything here, and no warp ever runs a tested block.
 *   finish   one lane per entry, lock step, at most a group (+ tail) of tested trips from the stored state: finds the
 *            exact trip and delivers.
 *
 * Two things follow from the loop's latency (one orbit advances a trip per 30 cycles with 2 warps per scheduler, per 97
 * with 8; tools/loopbench.cu): the long list is handed out LONGEST FIRST where lengths can be guessed (its hot region,
 * see stream_probe), and the long kernel runs on FEWER WARPS when the list is short (stream_long, occ_orbits_per_lane).
 *
 * Work items and where results go depend on the pass (chaos_render_args::phase), as in engine 1:
 *   phase 0  one sample per pixel: item = tile; result = the pixel's record;
 *   phase 1  pass A, samples 0 and 1 of every pixel: item = (tile, sample); escape times parked in the record;
 *   phase 3  pass C, the rounds pass B / the classifier exported: item = (exported tile, round); escape times to the
 *            export arrays.
 * Trip counts are those of quadratic.cuh whatever kernel runs which part of an orbit: nothing here touches the
 * arithmetic, and the exact counters (pixel_iterations, samples, skipped) are added exactly once per orbit, where it
 * delivers.
 */
#ifndef CHAOS_RENDER_STREAMS_CUH
#define CHAOS_RENDER_STREAMS_CUH

#define CHAOS_PROBE_BAIL 16u     /* probe trips after which a tile in which no orbit has ended goes to the long list as it is */

/* an orbit on its way from one kernel to the next (long list: after the probe's trips; finish list: before the group that
 * failed).  Entries of both lists are CHAOS_LIST_STRIDE bytes apart whatever Real is. */
template <class Real> struct alignas(16) finish_item {
    uint32_t a, b;      /* destination, see stream_dest */
    uint32_t it, pad;   /* trips run so far */
    Real x, y;          /* orbit state (Orbit::save) after them */
};
#define CHAOS_LIST_STRIDE 32u
static_assert(sizeof(finish_item<double>) == CHAOS_LIST_STRIDE && sizeof(finish_item<float>) == CHAOS_LIST_STRIDE, "list entries are 32 bytes");
template <class Real> static __device__ __forceinline__ finish_item<Real> *list_entry(void *list, uint32_t idx)
{
    return reinterpret_cast<finish_item<Real> *>(static_cast<char *>(list) + (size_t)idx * CHAOS_LIST_STRIDE);
}

/* the three ways a result is addressed, packed into two words */
struct stream_dest {
    uint32_t a, b;
};

template <class Real, class FractalT> struct stream_frame {
    typedef typename FractalT::template Orbit<Real> Orbit;
    frame_map<Real> fm;
    Real dx0, dy0, dx1, dy1;
    const Real *s_dx, *s_dy;    /* phase 3: offsets of rounds 0 .. CHAOS_EXPORT_ROUNDS-1 (shared memory) */
    uint32_t phase;

    __device__ __forceinline__ void init(const chaos_render_args &a, Real *sdx, Real *sdy)
    {
        fm.init(a);
        phase = a.phase;
        sample_delta<Real>(0u, 0.f, dx0, dy0);
        sample_delta<Real>(1u, 0.f, dx1, dy1);
        s_dx = sdx; s_dy = sdy;
        if (phase == 3u) {
            if (threadIdx.x < CHAOS_EXPORT_ROUNDS)
                sample_delta<Real>(threadIdx.x, sqrtf(__fadd_rn(a.max_ss, -2.0f)), sdx[threadIdx.x], sdy[threadIdx.x]);
            __syncthreads();
        }
    }
    /* destination words -> pixel and sample round */
    __device__ __forceinline__ void decode(const chaos_render_args &a, stream_dest d, uint32_t &px, uint32_t &py, uint32_t &rnd) const
    {
        if (phase == 3u) {
            uint32_t x0, y0;
            tile_origin(a, a.exp.tile[d.a], x0, y0);
            const uint32_t pix = d.b & 0xffu;
            rnd = d.b >> 8;
            px = x0 + (pix & 7u); py = y0 + (pix >> 3);
        } else {
            rnd = d.a >> 31;
            px = d.a & 0x7fffffffu; py = d.b;
        }
    }
    __device__ __forceinline__ void start(Orbit &o, uint32_t px, uint32_t py, uint32_t rnd, const orbit_ctx &ctx) const
    {
        Real dx, dy, cx, cy;
        if (phase == 3u) { dx = s_dx[rnd]; dy = s_dy[rnd]; }
        else if (rnd) { dx = dx1; dy = dy1; }
        else { dx = dx0; dy = dy0; }
        fm.template plane_point<fused_plane_y<FractalT>::value>(px, py, dx, dy, cx, cy);
        o.start(cx, cy, ctx);
    }
    /* an orbit is over: et = Orbit::finish(), it = the reference loop's trip count, sk = trips proven, not executed */
    __device__ __forceinline__ void deliver(const chaos_render_args &a, stream_dest d, uint32_t et, uint32_t it, uint32_t sk) const
    {
        if (phase == 3u) {
            const uint32_t rnd = d.b >> 8;
            export_et(a, d.a, rnd)[d.b & 0xffu] = et;
            atomicAdd(&a.exp.iters[(size_t)d.a * CHAOS_EXPORT_ROUNDS + rnd], (unsigned long long)it);
            if (sk) atomicAdd(&a.exp.skipped[(size_t)d.a * CHAOS_EXPORT_ROUNDS + rnd], (unsigned long long)sk);
        } else if (phase == 1u) {    /* pass A: sample 0 in `value` with the orbit's trip count and what it cost; sample 1 in `isReused` */
            chaos_pixel_info *rec = record_at(a.out, a.out_pitch, d.a & 0x7fffffffu, d.b);
            if ((d.a >> 31) == 0u) {
                *reinterpret_cast<float2 *>(&rec->value) = make_float2(__uint_as_float(et), __uint_as_float(it));
                rec->weight_of_new_samples = __uint_as_float(it - sk);
            } else {
                rec->is_reused = et;
            }
        } else {                     /* S == 1: value = (float)(sum / 1), weight = 1 (:152-153) */
            store_record(record_at(a.out, a.out_pitch, d.a, d.b), __uint2float_rn(et), 1.0f, 0u, 0.f);
        }
    }
};

/* exact work of the orbits a lane delivered.  Pass C (phase 3) does not count here: whether a round it ran exists at all
 * is known only to pass D, which adds the per-round sums deliver() left in the export arrays. */
struct stream_totals {
    unsigned long long iters, nsamples, skipped;
    __device__ __forceinline__ void add(uint32_t it, uint32_t sk) { iters += it; skipped += sk; nsamples += 1; }
    __device__ __forceinline__ void flush(const chaos_render_args &a) const { if (a.phase != 3u) flush_counters(a, iters, nsamples, skipped); }
};

/* ---- probe ------------------------------------------------------------------------------------------------ */
template <class Real, class Orbit>
static __device__ __forceinline__ void put_survivor(void *list, uint32_t idx, stream_dest d, uint32_t it, const Orbit &o)
{
    finish_item<Real> f;
    f.a = d.a; f.b = d.b; f.it = it; f.pad = 0u;
    o.save(f.x, f.y);
    *list_entry<Real>(list, idx) = f;
}
template <class Real, class FractalT>
static __device__ void stream_probe(const chaos_render_args &a)
{
    typedef typename FractalT::template Orbit<Real> Orbit;
    __shared__ Real s_dx[CHAOS_EXPORT_ROUNDS], s_dy[CHAOS_EXPORT_ROUNDS];
    stream_frame<Real, FractalT> sf;
    sf.init(a, s_dx, s_dy);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t max_iter = a.max_iter;
    const orbit_ctx ctx = {a.max_iter, a.shortcuts};
    const bool pass_c = a.phase == 3u, pass_a = a.phase == 1u;
    const uint32_t S0 = min(64u, __float2uint_rz(roundf(a.max_ss)));
    const uint32_t rounds_per_tile = pass_c ? S0 - 2u : pass_a ? 2u : 1u;
    const uint32_t n_items = pass_c ? min(a.counters->n_exported, a.exp.capacity) * rounds_per_tile : a.n_tiles * rounds_per_tile;
    chaos_stream_ctl *ctl = &a.counters->stream[pass_c ? 1 : 0];
    const uint32_t T0 = min(max_iter, a.probe_trips);
    stream_totals tot = {0ull, 0ull, 0ull};
    /* Items are dealt statically, warp w takes items w, w + warps, ...: an item's work is bounded (T0 trips), neighbouring
     * items go to neighbouring warps, and a common cursor would be the bottleneck -- half a million atomics on one
     * address take longer than the arithmetic of this kernel. */
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_items; t += warps) {
        uint32_t x0, y0, rnd = 0u, e = 0u;
        if (pass_c) {
            e = t / rounds_per_tile;
            rnd = 2u + (t - e * rounds_per_tile);
            if (rnd < a.exp.first[e]) continue;                 /* pass B had taken this round already */
            tile_origin(a, a.exp.tile[e], x0, y0);
        } else if (pass_a) {
            rnd = t & 1u;
            tile_origin(a, t >> 1, x0, y0);
        } else {
            tile_origin(a, t, x0, y0);
        }
        const uint32_t px = x0 + (lane & 7u), py = y0 + (lane >> 3);
        const bool inb = px < a.width && py < a.height;
        stream_dest d;
        if (pass_c) { d.a = e; d.b = (rnd << 8) | lane; }
        else { d.a = px | (rnd << 31); d.b = py; }
        Orbit o;
        uint32_t it = 0;
        bool ended = false, bailed = false;
        if (inb) sf.start(o, px, py, rnd, ctx);
        if (!Orbit::kResumable) {       /* one opaque call (fractal.cuh ClassicOrbit): the whole orbit, here */
            if (inb) { o.run(it, max_iter, true); ended = true; }
        } else {
            uint32_t lim = 0;
            while (lim < T0) {
                lim = min(lim + 8u, T0);
                if (inb && !ended) ended = o.run(it, lim, true) || it >= max_iter;
                const uint32_t live = __ballot_sync(CHAOS_FULL_MASK, inb && !ended);
                if (!live) break;
                if (lim >= CHAOS_PROBE_BAIL && !__any_sync(CHAOS_FULL_MASK, inb && ended)) { bailed = true; break; }
            }
        }
        if (inb && ended) {
            sf.deliver(a, d, o.finish(it, max_iter), it, o.skipped());
            tot.add(it, o.skipped());
        }
        /* survivors go to the long list.  Longest first: the launch ends when its last orbit does, and an orbit of maxIterations
         * trips that starts when the list runs dry IS the tail.  What is known about an orbit's length:
         *   pass C  the executed trips of the pixel's sample 0 are still in its record;
         *   other passes  a survivor of a tile in which some orbit ended sits next to the set's boundary -- where escapes are
         *           slow and convergence is slower -- while a tile in which nobody ended within the first trips (`bailed`) lies
         *           deep inside or deep in a zoom, where every orbit is alike.
         * Orbits expected to be long go to the list's hot region (its END, growing down), which the long kernel hands out
         * first.  Pass A and the one-sample pass cannot overflow the list (it holds every orbit they have), so there the two
         * ends cannot meet; pass C reserves the hot region (hot_capacity entries).
         * (Measured and removed: guessing the length from the orbit's rate of convergence -- the ratio of |z(64) - z(48)|^2 to
         * |z(48) - z(32)|^2 is |multiplier|^32 -- with a third region for the longest.  It does what it should: of c2's 46 000
         * orbits per pass that run all 10 000 trips, 92 % then start in the first 0.4 ms of the long kernel instead of evenly
         * over its 0.85 ms; the kernel is not shorter for it, because its last 0.4 ms belong to the few hundred long orbits any
         * guess misses) */
        bool hot = false;
        if (a.hot_capacity && inb && !ended)
            hot = pass_c ? __float_as_uint(record_at(a.out, a.out_pitch, px, py)->weight_of_new_samples) >= a.hot_trips : !bailed;
        const uint32_t surv_hot = __ballot_sync(CHAOS_FULL_MASK, hot);
        if (surv_hot) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&ctl->n_hot, (unsigned int)__popc(surv_hot));
            base = __shfl_sync(CHAOS_FULL_MASK, base, 0);
            if (hot) {
                const uint32_t idx = base + __popc(surv_hot & lanemask_lt());
                if (idx < a.hot_capacity) put_survivor<Real>(a.long_list, a.list_capacity - 1u - idx, d, it, o);
                else hot = false;                                   /* region full: an ordinary entry */
            }
        }
        const uint32_t surv = __ballot_sync(CHAOS_FULL_MASK, inb && !ended && !hot);
        if (surv) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&ctl->n_long, (unsigned int)__popc(surv));
            base = __shfl_sync(CHAOS_FULL_MASK, base, 0);
            if (inb && !ended && !hot) {
                const uint32_t idx = base + __popc(surv & lanemask_lt());
                if (idx < a.list_capacity - (pass_c ? a.hot_capacity : 0u)) {
                    put_survivor<Real>(a.long_list, idx, d, it, o);
                } else {                 /* list full (never with the host's sizing for passes 0 and A): the orbit is finished here */
                    run_whole(o, it, max_iter);
                    sf.deliver(a, d, o.finish(it, max_iter), it, o.skipped());
                    tot.add(it, o.skipped());
                }
            }
        }
    }
    tot.flush(a);
}

/* ---- orbit pool (protocol of render_refill.cuh, as an object) ------------------------------------------------ */
/* Once the long list is dry the warps drain on their own: lanes empty one by one while every instruction still takes
 * a whole issue slot, and an orbit of 10 000 trips takes 8 warps per scheduler 0.5 ms.  A warp left with fewer than
 * pool_min_lanes orbits parks them (state and all) and claims a warpful back: thin warps merge into full ones, warps that
 * come away empty end and leave their scheduler to the others, which then run faster.  One bounded MPMC ring per shard of
 * warps; hand-over through each entry's state word (free for lap g -> full -> free for lap g + 1), tagged with the
 * launch's epoch -- see render_refill.cuh for why each of these is the way it is.  Every wait is bounded: a hand-over
 * that does not arrive within CHAOS_SPIN_LIMIT polls raises chaos_counters::abort and the host fails the frame. */
#define CHAOS_SPIN_LIMIT (1u << 24)
template <class Rec> struct stream_pool {
    static_assert(sizeof(Rec) <= CHAOS_POOL_TAG_OFFSET, "parked orbit does not fit a pool entry");
    static constexpr uint32_t kWords = (sizeof(Rec) + 15u) / 16u;
    unsigned char *ring;
    uint32_t ring_size, epoch, lane;
    pool_ctl_ref pc;
    unsigned int *abort_flag;
    bool on;

    __device__ __forceinline__ void init(const chaos_render_args &a, uint32_t which)
    {
        lane = threadIdx.x & 31u;
        on = a.pool != nullptr && a.pool_min_lanes > 0u;
        const uint32_t shard = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) % CHAOS_POOL_SHARDS;
        ring_size = a.pool_capacity / CHAOS_POOL_SHARDS;
        ring = a.pool + (size_t)shard * ring_size * CHAOS_POOL_STRIDE;
        pc.live = &a.counters->pool[which][shard].live;
        pc.reserved = &a.counters->pool[which][shard].reserved;
        pc.head = &a.counters->pool[which][shard].head;
        epoch = a.pool_epoch + which;
        abort_flag = &a.counters->abort;
        if (on && lane == 0) atomicAdd(pc.live, 1u);
    }
    __device__ __forceinline__ bool wait_tag(const unsigned char *entry, uint32_t want) const
    {
        const unsigned int *tag = reinterpret_cast<const unsigned int *>(entry + CHAOS_POOL_TAG_OFFSET);
        for (uint32_t spins = 0; ld_volatile(tag) != want; ++spins)
            if (spins >= CHAOS_SPIN_LIMIT || ((spins & 1023u) == 1023u && ld_volatile(abort_flag))) { *abort_flag = 1u; return false; }
        return true;
    }
    /* empty lanes take parked orbits over; true for a lane that got one */
    __device__ __forceinline__ bool claim(bool empty, Rec &rec) const
    {
        const uint32_t idle = __ballot_sync(CHAOS_FULL_MASK, empty);
        if (!idle) return false;
        uint32_t base = 0, take = 0;
        if (lane == 0) take = pool_claim(pc, (uint32_t)__popc(idle), base);
        take = __shfl_sync(CHAOS_FULL_MASK, take, 0);
        base = __shfl_sync(CHAOS_FULL_MASK, base, 0);
        const uint32_t rank = __popc(idle & lanemask_lt());
        if (!empty || rank >= take) return false;
        unsigned char *entry = ring + (size_t)((base + rank) % ring_size) * CHAOS_POOL_STRIDE;
        if (!wait_tag(entry, pool_state(epoch, base + rank, ring_size, true))) return false;
        __threadfence();
        union { Rec r; uint4 w[kWords]; } u;
        const uint4 *src = reinterpret_cast<const uint4 *>(entry);
#pragma unroll
        for (uint32_t k = 0; k < kWords; ++k) u.w[k] = __ldcg(src + k);
        rec = u.r;
        __threadfence();            /* read before the entry is handed back */
        *reinterpret_cast<volatile unsigned int *>(entry + CHAOS_POOL_TAG_OFFSET) = pool_state(epoch, base + rank + ring_size, ring_size, false);
        return true;
    }
    /* all lanes that hold an orbit park it; false (nothing parked) if the ring has no room */
    __device__ __forceinline__ bool park(bool holds, const Rec &rec) const
    {
        const uint32_t running = __ballot_sync(CHAOS_FULL_MASK, holds);
        const uint32_t n_run = (uint32_t)__popc(running);
        uint32_t start = 0xffffffffu;
        if (lane == 0) {
            for (uint32_t spins = 0; spins < CHAOS_SPIN_LIMIT; ++spins) {
                const unsigned int r = ld_volatile(pc.reserved);
                if (r + n_run - ld_volatile(pc.head) > ring_size) break;
                if (atomicCAS(pc.reserved, r, r + n_run) == r) { start = r; break; }
            }
        }
        start = __shfl_sync(CHAOS_FULL_MASK, start, 0);
        if (start == 0xffffffffu) return false;
        if (holds) {
            const uint32_t index = start + __popc(running & lanemask_lt());
            unsigned char *entry = ring + (size_t)(index % ring_size) * CHAOS_POOL_STRIDE;
            /* the entry is free once the consumer of the previous lap has read it; anything from an older launch is free */
            if (index >= ring_size) { wait_tag(entry, pool_state(epoch, index, ring_size, false)); __threadfence(); }
            *reinterpret_cast<Rec *>(entry) = rec;
            __threadfence();
            *reinterpret_cast<volatile unsigned int *>(entry + CHAOS_POOL_TAG_OFFSET) = pool_state(epoch, index, ring_size, true);
        }
        __syncwarp();
        return true;
    }
    /* the warp holds nothing and would end: false = it may; true = it is the last one alive of its shard and something
     * is still parked, so it stays (and keeps whatever it claims from now on) */
    __device__ __forceinline__ bool must_stay() const
    {
        uint32_t stay = 0u;
        if (lane == 0 && atomicSub(pc.live, 1u) == 1u) {
            __threadfence();
            if (ld_volatile(pc.head) != ld_volatile(pc.reserved) && !ld_volatile(abort_flag)) { atomicAdd(pc.live, 1u); stay = 1u; }
        }
        return __shfl_sync(CHAOS_FULL_MASK, stay, 0) != 0u;
    }
};

/* ---- long ------------------------------------------------------------------------------------------------- */
/* When to refill.  A pass costs the warp about as much as 20 trips of all its lanes; a lane that is out (ended, stalled or
 * empty) costs 1/32 of the warp per trip it waits.  So lanes that are out run up a debt of lane-trips, and the pass is
 * taken when the debt reaches `sched_idle_lanes_indep` x 64 (default 10 x 64 = 20 trips x 32 lanes). */
template <class Orbit> struct stream_parked {
    Orbit o;
    uint32_t it, a, b;
    CHAOS_LS(unsigned long long tt;)      /* (diagnostics) when the orbit was taken from the list */
};
template <class Real, class FractalT>
static __device__ void stream_long(const chaos_render_args &a)
{
    typedef typename FractalT::template Orbit<Real> Orbit;
    typedef finish_item<Real> fin_t;
    typedef stream_parked<Orbit> parked_t;
    __shared__ Real s_dx[CHAOS_EXPORT_ROUNDS], s_dy[CHAOS_EXPORT_ROUNDS];
    stream_frame<Real, FractalT> sf;
    sf.init(a, s_dx, s_dy);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t max_iter = a.max_iter;
    const uint32_t nb = max(a.block_iters & ~31u, 32u);   /* whole groups: a block that ends inside a group would read as a tail */
    const orbit_ctx ctx = {a.max_iter, a.shortcuts};
    const uint32_t which = a.phase == 3u ? 1u : 0u;
    chaos_stream_ctl *ctl = &a.counters->stream[which];
    const uint32_t n_hot = min(ctl->n_hot, a.hot_capacity);      /* the hot region: the list's last n_hot entries, handed out first */
    const uint32_t n = n_hot + min(ctl->n_long, a.list_capacity - (a.phase == 3u ? a.hot_capacity : 0u));
    /* few orbits per lane: the upper CTA layers of every SM stay out (see chaos_render_args::occ_orbits_per_lane) */
    if (a.sm_count) {
        const uint32_t layer = blockIdx.x / a.sm_count;
        const uint32_t per_lane = n / (gridDim.x * blockDim.x);
        if (layer >= 1u && per_lane < a.occ_orbits_per_lane[min(layer, 3u) - 1u]) return;
    }
    const uint32_t max_debt = max(a.sched_idle_lanes_indep, 1u) * 64u;
    stream_totals tot = {0ull, 0ull, 0ull};
    stream_pool<parked_t> pool;
    pool.init(a, which);
    const bool pooling = pool.on && Orbit::kResumable;
    const uint32_t drain_interval = min(max(max_iter / (nb * 64u), 2u), 16u);   /* blocks between two looks at the pool */

    Orbit o;
    stream_dest d = {0u, 0u};
    uint32_t it = 0;
    bool busy = false, fin = false, stall = false;   /* fin: over, result known; stall: needs tested trips (finish list) */
    bool dry = n == 0u, keep_all = false;
    uint32_t debt = 0, drain_wait = 0, park_cooldown = 0;
    CHAOS_LS(lane_stats ls; ls.init();)
    CHAOS_LS(unsigned long long tt_orbit = 0ull;)
    CHAOS_LS(unsigned long long tt_start, tt_dry = 0ull; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt_start));)
    for (;;) {
        CHAOS_LS(if (dry && !tt_dry) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt_dry));)
        CHAOS_LS(ls.before(busy, fin || stall, false, false, dry, it);)
        CHAOS_LS(const long long ck0 = clock64(); const uint32_t it_before = it;)
        if (busy && !fin && !stall) {
            const uint32_t lim = min(it + nb, max_iter);
            const bool e = o.run(it, lim, false);
            fin = e || it >= max_iter;
            stall = !fin && o.wants_tested();
        }
        CHAOS_LS({
            const long long ck1 = clock64();
            const uint32_t adv = __reduce_max_sync(CHAOS_FULL_MASK, (fin && o.skipped()) ? 0u : it - it_before);
            if (lane == 0u && a.phase != 0u && adv) {
                atomicAdd(&a.counters->long_hist[which][3][dry ? 4 : 0], (unsigned long long)(ck1 - ck0));
                atomicAdd(&a.counters->long_hist[which][3][dry ? 5 : 1], (unsigned long long)adv);
                atomicAdd(&a.counters->long_hist[which][3][dry ? 6 : 2], 1ull);
            }
        })
        CHAOS_LS(ls.after((fin ? it - o.skipped() : it) - ls.it0);)
        const uint32_t running = __ballot_sync(CHAOS_FULL_MASK, busy && !fin && !stall);
        if (running) {
            if (!dry) {
                debt += (32u - (uint32_t)__popc(running)) * nb;
                if (debt < max_debt) continue;
            } else {
                /* nothing to refill from but the pool: a look every few blocks while lanes are out -- the orbits still running
                 * are the launch's critical path, a pass between any two of their blocks would stretch it */
                if (!pooling || running == CHAOS_FULL_MASK || ++drain_wait < drain_interval) continue;
            }
        }
        debt = 0u; drain_wait = 0u;
        CHAOS_LS(ls.pass();)
        /* retire */
        CHAOS_LS(if (fin && it >= max_iter && !o.skipped() && a.phase != 0u) {      /* (diagnostics) the orbits that ran all the way */
            unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            unsigned long long *w = &a.counters->lane_stats[3][0][which * 4u];
            atomicAdd(w + 0, 1ull); atomicAdd(w + 1, now - tt_orbit); atomicMax(w + 2, now - tt_orbit); atomicMax(w + 3, tt_orbit - tt_start);
            atomicAdd(&a.counters->lane_stats[1][which][min((tt_orbit - tt_start) >> 17, 7ull)], 1ull);       /* when they started, 0.131 ms bins */
        })
        if (fin) {
            sf.deliver(a, d, o.finish(it, max_iter), it, o.skipped());
            tot.add(it, o.skipped());
            fin = false; busy = false;
        }
        CHAOS_LS(if (stall && max_iter - it < 32u && a.phase != 0u) atomicAdd(&a.counters->lane_stats[1][which][min((tt_orbit - tt_start) >> 17, 7ull)], 1ull);)
        const uint32_t stalled = __ballot_sync(CHAOS_FULL_MASK, stall);
        if (stalled) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&ctl->n_finish, (unsigned int)__popc(stalled));
            base = __shfl_sync(CHAOS_FULL_MASK, base, 0);
            if (stall) {
                const uint32_t idx = base + __popc(stalled & lanemask_lt());
                if (idx < a.list_capacity) {
                    fin_t f;
                    f.a = d.a; f.b = d.b; f.it = it; f.pad = 0u;
                    o.save(f.x, f.y);
                    *list_entry<Real>(a.finish_list, idx) = f;
                } else {                 /* list full: the tested trips are run here */
                    o.run(it, max_iter, true);
                    sf.deliver(a, d, o.finish(it, max_iter), it, o.skipped());
                    tot.add(it, o.skipped());
                }
                stall = false; busy = false;
            }
        }
        /* refill */
        if (!dry) {
            const uint32_t idle = __ballot_sync(CHAOS_FULL_MASK, !busy);
            const uint32_t cnt = (uint32_t)__popc(idle);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&ctl->long_cursor, cnt);
            base = __shfl_sync(CHAOS_FULL_MASK, base, 0);
            const uint32_t idx = base + __popc(idle & lanemask_lt());
            if (!busy && idx < n) {
                const fin_t ent = *list_entry<Real>(a.long_list, idx < n_hot ? a.list_capacity - 1u - idx : idx - n_hot);
                d.a = ent.a; d.b = ent.b;
                uint32_t px, py, rnd;
                sf.decode(a, d, px, py, rnd);
                sf.start(o, px, py, rnd, ctx);
                o.resume(ent.x, ent.y);          /* goes on where the probe stopped */
                it = ent.it;
                busy = true;
                CHAOS_LS(asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt_orbit));)
            }
            if (base + cnt >= n) dry = true;
        }
        if (dry && pooling) {
            if (park_cooldown) --park_cooldown;
            for (int again = 0; again < 2; ++again) {
                parked_t rec;
                if (pool.claim(!busy, rec)) { o = rec.o; it = rec.it; d.a = rec.a; d.b = rec.b; busy = true; CHAOS_LS(tt_orbit = rec.tt;) }
                /* too few orbits left for a whole warp's issue slots: park them all, then claim a warpful */
                const uint32_t n_run = (uint32_t)__popc(__ballot_sync(CHAOS_FULL_MASK, busy));
                if (again || !n_run || n_run >= a.pool_min_lanes || keep_all || park_cooldown) break;
                park_cooldown = CHAOS_PARK_COOLDOWN;
                rec.o = o; rec.it = it; rec.a = d.a; rec.b = d.b;
                CHAOS_LS(rec.tt = tt_orbit;)
                if (!pool.park(busy, rec)) break;
                busy = false;
            }
        }
        if (!__any_sync(CHAOS_FULL_MASK, busy)) {
            if (!pooling || !pool.must_stay()) break;
            keep_all = true;
        }
    }
    tot.flush(a);
    CHAOS_LS(ls.flush(a, a.phase == 1u ? 0 : a.phase == 3u ? 2 : 3);)
    /* (diagnostics) when the launch's warps started, saw the list dry, and ended: min / max over the warps, as ~t for the minima */
    CHAOS_LS(if (lane == 0u && a.phase != 0u) {
        unsigned long long tt_end; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt_end));
        for (unsigned long long bin = 0; bin <= min((tt_end - tt_start) >> 16, 31ull); ++bin) atomicAdd(&a.counters->long_hist[which][0][bin], 1ull);
        unsigned long long *w = &a.counters->lane_stats[3][1][which * 4u];
        atomicMax(w + 0, ~tt_start); if (tt_dry) { atomicMax(w + 1, ~tt_dry); atomicMax(w + 2, tt_dry); } atomicMax(w + 3, tt_end);
    })
}

/* ---- finish ----------------------------------------------------------------------------------------------- */
template <class Real, class FractalT>
static __device__ void stream_finish(const chaos_render_args &a)
{
    typedef typename FractalT::template Orbit<Real> Orbit;
    typedef finish_item<Real> fin_t;
    __shared__ Real s_dx[CHAOS_EXPORT_ROUNDS], s_dy[CHAOS_EXPORT_ROUNDS];
    stream_frame<Real, FractalT> sf;
    sf.init(a, s_dx, s_dy);
    const uint32_t max_iter = a.max_iter;
    const orbit_ctx ctx = {a.max_iter, a.shortcuts};
    chaos_stream_ctl *ctl = &a.counters->stream[a.phase == 3u ? 1 : 0];
    const uint32_t n = min(ctl->n_finish, a.list_capacity);
    stream_totals tot = {0ull, 0ull, 0ull};
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const fin_t f = *list_entry<Real>(a.finish_list, idx);
        stream_dest d = {f.a, f.b};
        uint32_t px, py, rnd;
        sf.decode(a, d, px, py, rnd);
        Orbit o;
        sf.start(o, px, py, rnd, ctx);
        o.resume(f.x, f.y);
        uint32_t it = f.it;
        o.run(it, max_iter, true);
        sf.deliver(a, d, o.finish(it, max_iter), it, o.skipped());
        tot.add(it, o.skipped());
    }
    tot.flush(a);
}

#endif
