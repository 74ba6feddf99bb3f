/*
 * render_refill.cuh -- engine 1 of the main render kernels: persistent warps, work stealing over vote tiles,
 * finished lanes refilled with pending orbits.  Included by render_generic.cuh.
 *
 * Why: the escape loop runs 1 ... maxIterations trips per sample and neighbouring pixels differ by orders of
 * magnitude near the set's boundary, so "one warp = one 8x4 tile, all lanes wait for the slowest" (the reference's
 * mapping, fractalRendererGeneric.cu:36-53) leaves most FP64 lanes idle there.  Here an ORBIT (one sample of one
 * pixel) is the unit of work: tiles come from a global cursor, lanes iterate in blocks of trips, and at a
 * scheduling pass the lanes whose orbit ended retire their result and take the next pending orbit (rank among idle
 * lanes -> n-th pending pixel).
 *
 * One sample per pixel (round(maxSuperSampling) == 1: configs c1, c4): one launch of independent orbits
 * (render_main_independent, kMode 0).
 *
 * Several samples: the reference's early-termination decision (:128-150) is an ALL-vote over the tile after each
 * sample, so sample i+1 of a tile needs sample i of all its pixels.  The frame runs in passes:
 *   pass A   samples 0 and 1 of every pixel need no vote (the first decision comes after sample 1): independent orbits
 *            (kMode 1, work item = (tile, sample)); escape times, trip count and executed trips wait in the pixel's record;
 *   classify/order   one warp per tile takes the decision after sample 1.  Tiles it ends get their final record; tiles
 *            that go on and are set to use their whole budget are exported (see pass C); the rest get a cost class and are
 *            counting-sorted: tile_order, most expensive first;
 *   pass B   rounds 2.. with the votes (render_main_rounds).  Every warp owns CHAOS_REFILL_SLOTS tile slots in shared
 *            memory (per pixel: sum of the decided rounds' escape times + one escape time per round for the first
 *            ten rounds, the state sampleTheFractal keeps in registers, :96-127); lanes take pending orbits of any
 *            slot and any round in flight; when the round that is next in order is complete the warp evaluates the
 *            decision cooperatively, lane p speaking for pixel p, with the reference's predicates and votes -- so
 *            sample counts, sums and therefore every stored record are the reference's;
 *   pass C   the remaining rounds of the EXPORTED tiles (tiles set to use their whole sample budget; every tile when
 *            few are left), as independent orbits of one GPU-wide pool (kMode 2);
 *   pass D   the decisions of those tiles, replayed over the stored rounds (replay_exported).
 * Once the tile queue of pass A or C is dry, warps left with few orbits park them in the orbit pool and full warps are
 * repacked from it (below); a frame is cut into strands whose pass chains run next to each other (chaos_abi.cpp).
 * No result depends on the order in which anything runs; only the launches' tails do.
 */
#ifndef CHAOS_RENDER_REFILL_CUH
#define CHAOS_RENDER_REFILL_CUH

#define CHAOS_TILE_DONE 0xffffffffu   /* tile_key of a tile chaosClassifyTiles finished (its decision after sample 1 ended it) */
#ifndef CHAOS_REFILL_SLOTS
#define CHAOS_REFILL_SLOTS 4
#endif
#define CHAOS_REFILL_WARPS (CHAOS_RENDER_THREADS / 32)

/* Rounds 0 .. CHAOS_OVERLAP_ROUNDS-1 of a tile keep one escape time per pixel and round (the reference keeps the same
 * ten values in samples[], :96), so several of them may be IN FLIGHT at once; their effects become visible to the
 * decision logic strictly in round order.  Later rounds (maxSuperSampling > 10) only add to the running sum and run
 * one at a time. */
#define CHAOS_OVERLAP_ROUNDS CHAOS_ADAPTIVE_THRESHOLD

struct refill_slot_hdr {
    uint32_t x0, y0;   /* tile origin */
    uint32_t S;        /* current sample bound (sampleCount) */
    uint32_t dec;      /* the round whose decision is outstanding; every earlier round is complete and decided */
    uint32_t issued;   /* rounds dec .. issued-1 are in flight */
    uint32_t inb;      /* in-bounds pixels = the voters */
    uint32_t active;
    uint32_t rmask;    /* bit j: round slot j (= round % CHAOS_OVERLAP_ROUNDS) still has orbits no lane has taken */
    uint32_t tile;     /* the tile's index */
    uint32_t pend[CHAOS_OVERLAP_ROUNDS];   /* per round slot: pixels whose orbit has not been handed to a lane yet */
    uint32_t left[CHAOS_OVERLAP_ROUNDS];   /* per round slot: orbits not yet retired */
    /* work of the retired orbits of a round; counted into the frame's totals when the round is DECIDED, dropped if the
     * tile stops before it (a round started ahead of its turn that turns out not to exist) */
    unsigned long long iters[CHAOS_OVERLAP_ROUNDS], skipped[CHAOS_OVERLAP_ROUNDS];
};

struct refill_warp_store {
    uint32_t sum[CHAOS_REFILL_SLOTS][32];                            /* escape times of the decided rounds, summed */
    uint32_t et[CHAOS_REFILL_SLOTS][CHAOS_OVERLAP_ROUNDS][32];       /* escape time per round and pixel */
    refill_slot_hdr hdr[CHAOS_REFILL_SLOTS];
};

#define CHAOS_REFILL_SMEM_BYTES (sizeof(refill_warp_store) * CHAOS_REFILL_WARPS)

static __device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

/* position of the n-th (0-based) set bit of m; n < popc(m).  (__fns is a loop of up to 32 steps.) */
static __device__ __forceinline__ uint32_t nth_set_bit(uint32_t m, uint32_t n)
{
    uint32_t pos = 0;
#pragma unroll
    for (uint32_t w = 16u; w; w >>= 1) {
        const uint32_t c = __popc(m & ((1u << w) - 1u));
        if (n >= c) { n -= c; m >>= w; pos += w; }
    }
    return pos;
}

/* 64-bit add in shared memory from 32-bit pieces (a 64-bit shared atomicAdd is a compare-and-swap loop) */
static __device__ __forceinline__ void shared_add64(unsigned long long *p, uint32_t v)
{
    unsigned int *w = reinterpret_cast<unsigned int *>(p);
    const unsigned int old = atomicAdd(w, v);
    if (old + v < old) atomicAdd(w + 1, 1u);
}

/* ---- phases ------------------------------------------------------------------------------------------
 * All lanes of a warp run the escape loop together for a BLOCK of trips, then meet at a scheduling point.
 * A block is either TESTED (every trip carries its escape test; short: CHAOS_TESTED_BLOCK trips) or UNTESTED
 * (args.block_iters trips of the orbits' cheaper instruction stream, see Orbit::run in fractal.cuh).  The kind is
 * uniform over the warp, so the warp executes one instruction stream either way:
 *   - a block is tested if any lane holds a new orbit (most orbits end within a few trips; they then give the
 *     lane back after a short block instead of a long one) or any lane asks for tests (its untested group failed:
 *     the tested block finds the exact trip while every other lane keeps advancing);
 *   - otherwise it is untested.
 * (Two orbits per lane with interleaved chains was measured in round 1 and removed: no gain once the loop is
 * pipe-bound, a loss with sample rounds.) */
#ifndef CHAOS_GROUP
#define CHAOS_GROUP 32u          /* (quadratic.cuh; modules without that loop still need the block length) */
#endif
#define CHAOS_TESTED_BLOCK (CHAOS_GROUP + 8u)   /* >= quadratic_orbit::kGroup + 1: the replay of a failed group ends inside one tested block */

/* A scheduling pass costs the warp some hundred instructions, a lane that waits for one costs nothing but its share
 * of the next blocks.  So a pass is not taken for every finished orbit: finished lanes wait until `idle_lanes` lanes
 * are finished or empty, or `CHAOS_SCHED_MAX_WAIT` blocks have gone by since the first of them finished, or no lane is
 * running any more.  (Results do not depend on when a pass is taken.) */
#define CHAOS_SCHED_MAX_WAIT 8u
static __device__ __forceinline__ bool take_scheduling_pass(bool fin, bool busy, uint32_t idle_lanes, uint32_t &waited)
{
    const uint32_t n_fin = __popc(__ballot_sync(CHAOS_FULL_MASK, fin));
    const uint32_t n_run = __popc(__ballot_sync(CHAOS_FULL_MASK, busy && !fin));
    if (n_run == 0u) { waited = 0u; return true; }
    if (n_fin == 0u) { waited = 0u; return false; }
    if (32u - n_run >= idle_lanes || ++waited >= CHAOS_SCHED_MAX_WAIT) { waited = 0u; return true; }
    return false;
}

#ifdef CHAOS_LANE_STATS
struct lane_stats {
    unsigned long long v[2][8];
    uint32_t it0;
    bool run0, tested0;
    __device__ void init() { for (int i = 0; i < 2; ++i) for (int k = 0; k < 8; ++k) v[i][k] = 0ull; }
    /* before a block: what every lane is about to do */
    __device__ void before(bool busy, bool fin, bool wants_tested, bool tested, bool queue_empty, uint32_t it)
    {
        it0 = it; tested0 = tested;
        run0 = busy && !fin && (tested || !wants_tested);
        cls = run0 ? 0 : (busy && !fin) ? CHAOS_LS_REPLAY_WAIT : (busy && fin) ? CHAOS_LS_FIN_WAIT : queue_empty ? CHAOS_LS_IDLE_DRY : CHAOS_LS_IDLE_QUEUE;
    }
    int cls;
    /* after it: `adv` = trips this lane advanced (a proven orbit: up to the proof) */
    __device__ void after(uint32_t adv)
    {
        const uint32_t a = run0 ? adv : 0u;
        const uint32_t L = __reduce_max_sync(CHAOS_FULL_MASK, a);
        const int t = tested0 ? 0 : 1;
        v[t][CHAOS_LS_CAPACITY] += L; v[t][CHAOS_LS_USEFUL] += a;
        if (cls) v[t][cls] += L;
        if ((threadIdx.x & 31u) == 0) v[t][CHAOS_LS_BLOCKS] += 1;
    }
    __device__ void pass() { if ((threadIdx.x & 31u) == 0) v[0][CHAOS_LS_PASSES] += 1; }
    __device__ void flush(const chaos_render_args &a, int which)
    {
        for (int i = 0; i < 2; ++i) for (int k = 0; k < 8; ++k) {
            unsigned long long x = v[i][k];
            for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(CHAOS_FULL_MASK, x, o);
            if ((threadIdx.x & 31u) == 0 && x) atomicAdd(&a.counters->lane_stats[which][i][k], x);
        }
    }
};
#define CHAOS_LS(...) __VA_ARGS__
#else
#define CHAOS_LS(...)
#endif

template <class Orbit>
static __device__ __forceinline__ bool run_block(Orbit &o, uint32_t &it, bool busy, bool tested, uint32_t nb, uint32_t max_iter)
{
    if (!busy) return false;
    const uint32_t lim = Orbit::kResumable ? min(it + (tested ? CHAOS_TESTED_BLOCK : nb), max_iter) : max_iter;
    const bool ended = o.run(it, lim, tested);
    return ended || it >= max_iter;
}

/* ---- orbit pool (chaos_render_args::pool) ------------------------------------------------------ */
/* A bounded multi-producer multi-consumer ring per shard of warps.  Producers reserve a range of indices (never more than
 * a ring's length ahead of the consumers' cursor) and consumers claim reserved indices, both with one compare-and-swap
 * on a cursor; the hand-over of an entry goes through the entry's own state word (its last word):
 *     free for lap g  ->  [producer writes the orbit]  ->  full, lap g  ->  [consumer reads it]  ->  free for lap g + 1
 * so a producer never writes over an entry of the previous lap that its consumer has claimed but not read yet, and a
 * consumer never reads an entry its producer has reserved but not written yet.  (Both waits are short and on a word of
 * their own.  Publishing through one common counter in reservation order made thousands of warps spin on one word, and
 * a single ring made them retry their compare-and-swap against each other -- O(warps^2) atomics, frames 10x slower.
 * Tags without the "free" state let a producer lap a slow consumer when the SM's store path was backed up behind the
 * compose kernel's PCIe writes: frames that never ended.)  The word carries the launch's epoch, so a ring is never
 * reset: whatever an older launch left there reads as "free for lap 0".  A ring holds 32 x its warps entries -- every
 * orbit its warps can hold at once.
 * `live` counts the warps of the shard that have not ended.  A warp ends by decrementing it; the one that brings it to
 * zero looks at the ring once more and stays if something is parked, so a parked orbit is always picked up. */
#define CHAOS_PARK_COOLDOWN 4u   /* looks at the pool a warp that parked lets go by before it parks again */
template <class Orbit> struct parked_orbit {
    Orbit o;
    uint32_t it, px, py, tile, rnd;
};
#define CHAOS_POOL_TAG_OFFSET (CHAOS_POOL_STRIDE - 4u)   /* the tag is the entry's last word */
struct pool_ctl_ref {
    unsigned int *live, *reserved, *head;
};
/* state word of entry `index`: full = false: free for this index' lap; true: holds this index' orbit */
static __device__ __forceinline__ uint32_t pool_state(uint32_t epoch, uint32_t index, uint32_t ring_size, bool full)
{
    return (epoch << 8) | ((2u * (index / ring_size) + (full ? 1u : 0u)) & 0xffu);
}
static __device__ __forceinline__ unsigned int ld_volatile(const unsigned int *p) { return *reinterpret_cast<const volatile unsigned int *>(p); }
/* every wait of the hand-over is bounded: a loop that does not end within 2^24 polls (seconds; a hand-over takes
 * microseconds) raises chaos_counters::abort and leaves; the host then fails the frame with CHAOS_ERR_CUDA */
#ifdef CHAOS_POOL_DEBUG   /* diagnostics: name the loop that does not end */
#define CHAOS_SPIN_GUARD(n, what, ...) if (++(n) >= (1u << 24)) { printf("pool spin: " what "\n", __VA_ARGS__); *chaos_abort_flag = 1u; break; }
#else
#define CHAOS_SPIN_GUARD(n, what, ...) if (++(n) >= (1u << 24)) { *chaos_abort_flag = 1u; break; }
#endif
/* lane 0: claim up to `want` reserved entries; returns how many, first index in `base` */
static __device__ __forceinline__ uint32_t pool_claim(const pool_ctl_ref &c, uint32_t want, uint32_t &base)
{
    for (;;) {      /* lock-free: a failed compare-and-swap means another warp made progress */
        const unsigned int h = ld_volatile(c.head), p = ld_volatile(c.reserved);
        if (p <= h) return 0u;
        const uint32_t take = min(want, p - h);
        if (atomicCAS(c.head, h, h + take) == h) { base = h; return take; }
    }
}

/* ---- independent orbits ---------------------------------------------------------------------- */
/* kMode 0: one sample per pixel, the frame's only launch.  1: pass A, sample 0 of every pixel.  2: pass C, the rounds
 * pass B exported -- a work item is (exported tile, round), handed out exactly like a tile. */
static __device__ __forceinline__ uint32_t *export_et(const chaos_render_args &a, uint32_t e, uint32_t round)
{
    return a.exp.et + ((size_t)e * CHAOS_EXPORT_ROUNDS + round) * 32u;
}
template <class Real, class FractalT, int kMode>
static __device__ void render_main_independent(const chaos_render_args &a)
{
    constexpr bool kProbe = kMode == 1, kExport = kMode == 2;
    typedef typename FractalT::template Orbit<Real> Orbit;
    frame_map<Real> fm;
    fm.init(a);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t max_iter = a.max_iter;
    const uint32_t nb = a.block_iters;
    const orbit_ctx ctx = {a.max_iter, a.shortcuts};
    Real dx0, dy0;
    sample_delta<Real>(0u, 0.f, dx0, dy0);                  /* sample 0 sits at offset 0/3, */
    Real dx1, dy1;
    sample_delta<Real>(1u, 0.f, dx1, dy1);                  /* sample 1 at 1/3 (a division: once per launch, not per orbit) */
    /* pass C: the offsets of rounds 2 .. 9 (integer and float divisions each) once per CTA */
    __shared__ Real s_dx[kExport ? CHAOS_EXPORT_ROUNDS : 1], s_dy[kExport ? CHAOS_EXPORT_ROUNDS : 1];
    if (kExport) {
        if (threadIdx.x < CHAOS_EXPORT_ROUNDS) sample_delta<Real>(threadIdx.x, sqrtf(__fadd_rn(a.max_ss, -2.0f)), s_dx[threadIdx.x], s_dy[threadIdx.x]);
        __syncthreads();
    }
    const uint32_t S0 = min(64u, __float2uint_rz(roundf(a.max_ss)));
    const float spr = sqrtf(__fadd_rn(a.max_ss, -2.0f));
    /* work items per tile: pass C rounds 2 .. S0-1; pass A rounds 0 and 1 (both exist whenever S0 >= 2: the first
     * decision comes after sample 1, :128); one sample per pixel otherwise */
    const uint32_t rounds_per_tile = kExport ? S0 - 2u : kProbe ? 2u : 1u;
    uint32_t n_items = kExport ? min(a.counters->n_exported, a.exp.capacity) * rounds_per_tile : a.n_tiles * rounds_per_tile;
    unsigned int *cursor = kExport ? &a.counters->next_export_item : &a.counters->next_tile;
    /* cross-GPU stealing (one-sample frames, chaos_render_args::steal_world): the rank whose tiles are being handed out now,
     * and per lane the rank its orbit belongs to */
    const bool stealing = kMode == 0 && a.steal_world > 1u;
    uint32_t victim = a.steal_rank, owner = a.steal_rank;
    unsigned long long own_done = 0;
    if (stealing && blockIdx.x == 0 && threadIdx.x == 0) {      /* the counters were cleared for this frame: the cursor is open */
        __threadfence_system();
        *reinterpret_cast<volatile unsigned int *>(&a.counters->frame_seq) = a.frame_seq;
    }

    Orbit o;
    uint32_t it = 0, px = 0, py = 0, tile = 0, rnd = 0;      /* pass C: tile = export index, px = pixel within the tile */
    bool busy = false, fin = false;                          /* fin: the orbit is over and waits to be retired */
    bool first = true, queue_empty = false, tested = true;
    uint32_t pend = 0, x0 = 0, y0 = 0, cur_tile = 0, cur_round = 0;   /* warp-uniform: the tile being handed out */
    uint32_t waited = 0;
    unsigned long long iters = 0, nsamples = 0, skipped = 0;
    /* orbit pool: only orbits that can be suspended can change warps */
    typedef parked_orbit<Orbit> parked_t;
    static_assert(sizeof(parked_t) <= CHAOS_POOL_TAG_OFFSET, "parked orbit does not fit a pool entry");
    /* (not in the single launch of a one-sample frame: its drain is a small part of it -- c4: 0.6 % -- and the bookkeeping
     * in the loop cost c1 20 % and c4 3 %) */
    const bool pooling = kMode != 0 && Orbit::kResumable && a.pool != nullptr && a.pool_min_lanes > 0u;
    /* One ring per shard of warps: with a single ring thousands of warps that park and claim at the same moment (the
     * queue runs dry for all of them at once) retry their compare-and-swap on one word -- O(warps^2) atomics, frames 10x
     * slower.  The last warp alive of a shard keeps what is left of it. */
    const uint32_t shard = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) % CHAOS_POOL_SHARDS;
    const uint32_t ring_size = a.pool_capacity / CHAOS_POOL_SHARDS;
    unsigned char *const ring = a.pool + (size_t)shard * ring_size * CHAOS_POOL_STRIDE;
    const pool_ctl_ref pc = {&a.counters->pool[kExport ? 1 : 0][shard].live, &a.counters->pool[kExport ? 1 : 0][shard].reserved,
                             &a.counters->pool[kExport ? 1 : 0][shard].head};
    const uint32_t pool_epoch = a.pool_epoch + (kExport ? 1u : 0u);
    unsigned int *const chaos_abort_flag = &a.counters->abort;
    bool keep_all = false;                                   /* this warp is the launch's last one: it parks nothing (any more) */
    uint32_t park_cooldown = 0, drain_wait = 0;
    const uint32_t drain_interval = min(max(max_iter / (nb * 64u), 2u), 16u);   /* blocks */
    uint32_t stay_spins = 0;
    if (pooling && lane == 0) atomicAdd(pc.live, 1u);

    CHAOS_LS(lane_stats ls; ls.init();)
    uint32_t loop_spins = 0; (void)loop_spins;
    for (;;) {
#ifdef CHAOS_POOL_DEBUG
        if (++loop_spins >= 3000000u) {
            if (lane == 0) printf("main loop: warp %u shard %u busy %x fin %x queue_empty %d keep_all %d head %u reserved %u live %u it %u tested %d wants %x\n",
                                  (blockIdx.x * blockDim.x + threadIdx.x) >> 5, shard, __ballot_sync(1u, 1), 0u, (int)queue_empty, (int)keep_all,
                                  ld_volatile(pc.head), ld_volatile(pc.reserved), ld_volatile(pc.live), it, (int)tested, 0u);
            break;
        }
#endif
        CHAOS_LS(ls.before(busy, fin, o.wants_tested(), tested, queue_empty, it);)
        fin |= run_block(o, it, busy && !fin, tested, nb, max_iter);
        CHAOS_LS(ls.after((fin ? it - o.skipped() : it) - ls.it0);)
        tested = __any_sync(CHAOS_FULL_MASK, busy && !fin && o.wants_tested());
        /* Queue dry: a look at the pool every few blocks while lanes are empty.  Not after every block: the orbits still
         * running are the launch's critical path, and a pass (a few hundred instructions and a round trip to L2 for the
         * ring's cursors) between any two blocks of a warp that runs alone nearly doubles the time its orbits take
         * (c4 on 2 GPUs: 19 -> 23 ms).  The interval grows with the iteration limit: long orbits, long drain. */
        bool drain_pass = false;
        if (pooling && queue_empty && __any_sync(CHAOS_FULL_MASK, !busy || fin) && ++drain_wait >= drain_interval) { drain_pass = true; drain_wait = 0u; }
        if (!take_scheduling_pass(fin, busy, a.sched_idle_lanes_indep, waited) && !first && !drain_pass) continue;
        first = false;
        CHAOS_LS(ls.pass();)
        const bool retired_now = fin;
        const uint32_t retired_owner = owner;
        if (fin) {
            fin = false;
            uint32_t et = o.finish(it, max_iter);
            iters += it;
            skipped += o.skipped();
            nsamples += 1;
            if (kExport) {   /* pass C: the escape time waits for pass D; the work is counted when pass D knows the round existed */
                export_et(a, tile, rnd)[px] = et;
                atomicAdd(&a.exp.iters[(size_t)tile * CHAOS_EXPORT_ROUNDS + rnd], (unsigned long long)it);
                if (o.skipped()) atomicAdd(&a.exp.skipped[(size_t)tile * CHAOS_EXPORT_ROUNDS + rnd], (unsigned long long)o.skipped());
            } else if (kProbe) { /* pass A: park the escape times in the record for chaosClassifyTiles and pass B -- sample 0 in
                                  * `value`, with the orbit's trip count and what it cost (a proven never-ending orbit is
                                  * cheap); sample 1 in the `isReused` word */
                chaos_pixel_info *rec = record_at(a.out, a.out_pitch, px, py);
                if (rnd == 0u) {
                    *reinterpret_cast<float2 *>(&rec->value) = make_float2(__uint_as_float(et), __uint_as_float(it));
                    rec->weight_of_new_samples = __uint_as_float(it - o.skipped());
                } else {
                    rec->is_reused = et;
                }
            }
            else if (!stealing || owner == a.steal_rank) {   /* S == 1: value = (float)(sum / 1), weight = 1 (:152-153) */
                store_record(record_at(a.out, a.out_pitch, px, py), __uint2float_rn(et), 1.0f, 0u, 0.f);
                own_done += 1;
            } else {      /* a stolen orbit: the record goes straight into its owner's buffer (a peer store over NVLink) */
                store_record(record_at(a.steal_out[owner], a.out_pitch, px, py), __uint2float_rn(et), 1.0f, 0u, 0.f);
                __threadfence_system();
            }
            busy = false;
        }
        if (stealing) {   /* tell the owners how many of their orbits ended here: one system-scope add per owner and warp */
            const bool foreign = retired_now && retired_owner != a.steal_rank;
            const uint32_t fm = __ballot_sync(CHAOS_FULL_MASK, foreign);
            if (fm) {
                if (foreign) {
                    const uint32_t same = __match_any_sync(fm, retired_owner);
                    if (lane == (uint32_t)__ffs(same) - 1u)
                        atomicAdd_system(&a.steal_counters[retired_owner]->foreign_done, (unsigned long long)__popc(same));
                }
                __syncwarp();
            }
        }
        for (;;) {
            uint32_t idle = __ballot_sync(CHAOS_FULL_MASK, !busy);
            if (!idle) break;
            if (!pend) {
                if (queue_empty) break;
                uint32_t t = 0;
                if (lane == 0) t = (stealing && victim != a.steal_rank) ? atomicAdd_system(cursor, 1u) : atomicAdd(cursor, 1u);
                t = __shfl_sync(CHAOS_FULL_MASK, t, 0);
                if (t >= n_items) {
                    if (!stealing) { queue_empty = true; break; }
                    /* this rank's tiles are handed out: go on with the next rank whose counters belong to this frame (a rank
                     * that has not started the frame yet is passed over, not waited for) */
                    uint32_t next = a.steal_world;
                    if (lane == 0) {
                        /* ranks are visited once each, in ring order after the own rank */
                        for (uint32_t q = (victim + 1u) % a.steal_world; q != a.steal_rank; q = (q + 1u) % a.steal_world) {
                            const unsigned int seq = *reinterpret_cast<const volatile unsigned int *>(&a.steal_counters[q]->frame_seq);
                            if (seq == a.frame_seq && a.steal_n_tiles[q]) { next = q; break; }
                        }
                    }
                    next = __shfl_sync(CHAOS_FULL_MASK, next, 0);
                    if (next >= a.steal_world) { queue_empty = true; break; }
                    victim = next;
                    cursor = &a.steal_counters[victim]->next_tile;
                    n_items = a.steal_n_tiles[victim];
                    continue;
                }
                if (kExport) {
                    cur_tile = t / rounds_per_tile;                       /* export index */
                    cur_round = 2u + (t - cur_tile * rounds_per_tile);
                    if (cur_round < a.exp.first[cur_tile]) continue;      /* pass B had taken this round already */
                    tile_origin(a, a.exp.tile[cur_tile], x0, y0);
                } else if (kProbe) {
                    cur_tile = t >> 1;
                    cur_round = t & 1u;
                    tile_origin(a, cur_tile, x0, y0);
                } else {
                    cur_tile = t;
                    if (stealing) tile_origin_of(a.tiles_x, victim, a.part_count, a.band_tile_rows, t, x0, y0);   /* the owner's band deal */
                    else tile_origin(a, t, x0, y0);
                }
                pend = __ballot_sync(CHAOS_FULL_MASK, (x0 + (lane & 7u)) < a.width && (y0 + (lane >> 3)) < a.height);
            }
            uint32_t rank = __popc(idle & lanemask_lt());
            bool take = !busy && rank < (uint32_t)__popc(pend);
            uint32_t mypix = 0;
            if (take) {
                mypix = (pend == CHAOS_FULL_MASK && idle == CHAOS_FULL_MASK) ? lane : nth_set_bit(pend, rank);
                tile = cur_tile;
                Real cx, cy;
                if (kExport) {
                    const Real dx = s_dx[cur_round], dy = s_dy[cur_round];
                    fm.template plane_point<fused_plane_y<FractalT>::value>(x0 + (mypix & 7u), y0 + (mypix >> 3), dx, dy, cx, cy);
                    px = mypix;
                    rnd = cur_round;
                } else {
                    px = x0 + (mypix & 7u);
                    py = y0 + (mypix >> 3);
                    Real dx = dx0, dy = dy0;
                    if (kProbe) { if (cur_round) { dx = dx1; dy = dy1; } rnd = cur_round; }
                    fm.template plane_point<fused_plane_y<FractalT>::value>(px, py, dx, dy, cx, cy);
                }
                o.start(cx, cy, ctx);
                it = 0;
                busy = true;
                owner = victim;
                tested = true;                               /* new orbits start with a tested block */
            }
            pend &= ~__reduce_or_sync(CHAOS_FULL_MASK, take ? (1u << mypix) : 0u);
        }
        if (pooling && queue_empty) {
            if (park_cooldown) --park_cooldown;
            for (int again = 0; again < 2; ++again) {
                /* empty lanes take parked orbits over */
                const uint32_t idle = __ballot_sync(CHAOS_FULL_MASK, !busy);
                if (idle) {
                    uint32_t base = 0, take = 0;
                    if (lane == 0) take = pool_claim(pc, (uint32_t)__popc(idle), base);
                    take = __shfl_sync(CHAOS_FULL_MASK, take, 0);
                    base = __shfl_sync(CHAOS_FULL_MASK, base, 0);
                    const uint32_t rank = __popc(idle & lanemask_lt());
                    if (!busy && rank < take) {
                        union { parked_t rec; uint4 w[(sizeof(parked_t) + 15u) / 16u]; } u;
                        const unsigned char *entry = ring + (size_t)((base + rank) % ring_size) * CHAOS_POOL_STRIDE;
                        const uint32_t want = pool_state(pool_epoch, base + rank, ring_size, true);
                        uint32_t spins = 0;
                        while (ld_volatile(reinterpret_cast<const unsigned int *>(entry + CHAOS_POOL_TAG_OFFSET)) != want) {
                            CHAOS_SPIN_GUARD(spins, "tag shard %u index %u want %x have %x head %u reserved %u", shard, base + rank, want,
                                             ld_volatile(reinterpret_cast<const unsigned int *>(entry + CHAOS_POOL_TAG_OFFSET)), ld_volatile(pc.head), ld_volatile(pc.reserved));
                        }
                        __threadfence();
                        const uint4 *src = reinterpret_cast<const uint4 *>(entry);
#pragma unroll
                        for (uint32_t k = 0; k < (sizeof(parked_t) + 15u) / 16u; ++k) u.w[k] = __ldcg(src + k);
                        o = u.rec.o; it = u.rec.it; px = u.rec.px; py = u.rec.py; tile = u.rec.tile; rnd = u.rec.rnd;
                        __threadfence();            /* read before the entry is handed back */
                        *reinterpret_cast<volatile unsigned int *>(const_cast<unsigned char *>(entry) + CHAOS_POOL_TAG_OFFSET) =
                            pool_state(pool_epoch, base + rank + ring_size, ring_size, false);
                        busy = true;
                        tested = tested || o.wants_tested();
                    }
                }
                /* Too few orbits left for a whole warp's pipe slots: park them all, then claim a warpful -- warps that park
                 * at about the same time repack their orbits into full warps, and the ones that come away empty end.  (No
                 * orbit waits in the pool for long: whoever parks claims right afterwards.)  A warp that stays thin all the
                 * same leaves it at that for a while. */
                const uint32_t running = __ballot_sync(CHAOS_FULL_MASK, busy);
                const uint32_t n_run = (uint32_t)__popc(running);
                if (again || !n_run || n_run >= a.pool_min_lanes || keep_all || park_cooldown) break;
                park_cooldown = CHAOS_PARK_COOLDOWN;
                uint32_t start = 0xffffffffu;
                if (lane == 0) {             /* reserve a range of the ring (a full ring just stops taking orbits) */
                    for (;;) {
                        const unsigned int r = ld_volatile(pc.reserved);
                        if (r + n_run - ld_volatile(pc.head) > ring_size) break;
                        if (atomicCAS(pc.reserved, r, r + n_run) == r) { start = r; break; }
                    }
                }
                start = __shfl_sync(CHAOS_FULL_MASK, start, 0);
                if (start == 0xffffffffu) break;
                if (busy) {
                    const uint32_t index = start + __popc(running & lanemask_lt());
                    unsigned char *entry = ring + (size_t)(index % ring_size) * CHAOS_POOL_STRIDE;
                    /* the entry is free once the consumer of the previous lap has read it; anything from an older launch is free */
                    if (index >= ring_size) {
                        const uint32_t free_now = pool_state(pool_epoch, index, ring_size, false);
                        uint32_t spins = 0;
                        while (ld_volatile(reinterpret_cast<const unsigned int *>(entry + CHAOS_POOL_TAG_OFFSET)) != free_now) {
                            CHAOS_SPIN_GUARD(spins, "free shard %u index %u want %x have %x", shard, index, free_now,
                                             ld_volatile(reinterpret_cast<const unsigned int *>(entry + CHAOS_POOL_TAG_OFFSET)));
                        }
                        __threadfence();
                    }
                    parked_t rec;
                    rec.o = o; rec.it = it; rec.px = px; rec.py = py; rec.tile = tile; rec.rnd = rnd;
                    *reinterpret_cast<parked_t *>(entry) = rec;
                    __threadfence();
                    *reinterpret_cast<volatile unsigned int *>(entry + CHAOS_POOL_TAG_OFFSET) = pool_state(pool_epoch, index, ring_size, true);
                    busy = false;
                }
                __syncwarp();
            }
        }
        tested = __any_sync(CHAOS_FULL_MASK, tested);
        if (!__any_sync(CHAOS_FULL_MASK, busy)) {
            if (!pooling) break;
            /* end of this warp -- unless it is the last one alive and something is still parked */
            uint32_t stay = 0u;
            if (lane == 0) {
                if (atomicSub(pc.live, 1u) == 1u) {
                    __threadfence();
                    if (ld_volatile(pc.head) != ld_volatile(pc.reserved)) { atomicAdd(pc.live, 1u); stay = 1u; }
                }
            }
            if (!__shfl_sync(CHAOS_FULL_MASK, stay, 0)) break;
            keep_all = true;
            CHAOS_SPIN_GUARD(stay_spins, "stay shard %u head %u reserved %u live %u", shard, ld_volatile(pc.head), ld_volatile(pc.reserved), ld_volatile(pc.live));
        }
    }
    if (!kExport) flush_counters(a, iters, nsamples, skipped);
    if (stealing) {
        for (int o2 = 16; o2; o2 >>= 1) own_done += __shfl_xor_sync(CHAOS_FULL_MASK, own_done, o2);
        if (lane == 0 && own_done) atomicAdd(&a.counters->own_done, own_done);
    }
    CHAOS_LS(ls.flush(a, kMode == 1 ? 0 : kMode == 2 ? 2 : 3);)
}

/* ---- general case: sample rounds with tile-wide votes ----------------------------------------- */
/* the first round after i at which the decision block is entered (:128); S if there is none */
static __device__ __forceinline__ uint32_t next_decision_round(bool adaptive, uint32_t i, uint32_t S)
{
    if (adaptive && i + 1u < CHAOS_ADAPTIVE_THRESHOLD) return i + 1u;
    return (S >> 1) > i ? (S >> 1) : S;
}

/*
 * A tile's rounds are decided in order, but they need not RUN in order: the decision after round i only needs rounds
 * 0..i complete.  After each decision the slot issues
 *   - every round up to the next decision point (certain to run: nothing can change S before it), and
 *   - if the tile looks set to use its whole sample budget -- some pixel's mean is 0 (its dispersion is undefined, which
 *     vetoes both stop rules: tiles the set's boundary runs through, in modules that report 0 inside), or, from
 *     sample 2 on, some pixel's dispersion is above 1 (not even the loosest stop rule fired) -- all remaining rounds
 *     below CHAOS_OVERLAP_ROUNDS, ahead of their turn.
 * A round started ahead of its turn is dropped without trace if the tile stops before it: its orbits are abandoned
 * where they are and their trips are not counted.  So the records and the exact counters do not depend on the policy;
 * only the tile's critical path does (one orbit instead of up to S sequential ones), and more orbits are pending per
 * warp to keep the lanes filled.
 */
template <class Real, class FractalT>
static __device__ void render_main_rounds(const chaos_render_args &a, refill_warp_store &ws)
{
    constexpr bool kResume = true;      /* sample 0 of every pixel was taken by pass A */
    typedef typename FractalT::template Orbit<Real> Orbit;
    constexpr int K = CHAOS_REFILL_SLOTS;
    constexpr uint32_t R = CHAOS_OVERLAP_ROUNDS;
    frame_map<Real> fm;
    fm.init(a);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t max_iter = a.max_iter;
    const uint32_t nb = a.block_iters;
    const orbit_ctx ctx = {a.max_iter, a.shortcuts};
    const bool adaptive = (a.flags & CHAOS_FLAG_ADAPTIVE_SS) != 0u;
    const float scf = a.max_ss;                               /* host guarantees >= 1 (:174) */
    const uint32_t S0 = min(64u, __float2uint_rz(roundf(scf)));
    const float spr = sqrtf(__fadd_rn(scf, -2.0f));
    /* the sample offsets of all rounds (integer and float divisions each) once per CTA */
    __shared__ Real s_dx[64], s_dy[64];
    if (threadIdx.x < 64u) sample_delta<Real>(threadIdx.x, spr, s_dx[threadIdx.x], s_dy[threadIdx.x]);
    __syncthreads();
    const uint32_t n_cont = kResume ? a.counters->n_continuing : a.n_tiles;   /* tiles chaosClassifyTiles left for this pass */
    /* Few tiles left (a frame whose tiles mostly ended after sample 1): this pass would be all latency -- a tile's rounds
     * run one after the other here, each as long as its longest orbit, with most lanes of the GPU empty (c2: 0.4 ms at
     * 6 % of the lanes).  Then every tile is exported as it arrives: its rounds run side by side in pass C.  Rounds that
     * turn out not to exist are wasted work, which is why a frame with many such tiles (c2ex2) does not do it. */
    const bool export_everything = n_cont <= a.export_all_below;
    if (export_everything && a.exp.capacity && a.export_all_done) return;     /* chaosExportAll has taken them all */

    for (uint32_t w = lane; w < sizeof(ws.hdr) / 4u; w += 32u) reinterpret_cast<uint32_t *>(ws.hdr)[w] = 0u;
    __syncwarp();

    Orbit o;
    uint32_t it = 0, slot = 0, pix = 0, rnd = 0;
    bool busy = false, fin = false;                          /* fin: the orbit is over and waits to be retired */
    bool first = true, queue_empty = false, tested = true;
    uint32_t waited = 0;
    unsigned long long iters = 0, nsamples = 0, skipped = 0;   /* lane 0 carries the warp's totals */
    /* diagnostics (args.warp_trace): when the warp started, saw the queue run dry, and ended; what it did */
    unsigned long long tr_start = 0, tr_dry = 0, tr_blocks = 0, tr_passes = 0, tr_tiles = 0, tr_lane_blocks = 0;
    if (a.warp_trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_start));
    CHAOS_LS(lane_stats ls; ls.init();)

    for (;;) {
        /* (1) iterate: one block of trips for the orbit this lane holds */
        if (a.warp_trace) { tr_blocks += 1; tr_lane_blocks += __popc(__ballot_sync(CHAOS_FULL_MASK, busy && !fin)); }
        CHAOS_LS(ls.before(busy, fin, o.wants_tested(), tested, queue_empty, it);)
        fin |= run_block(o, it, busy && !fin, tested, nb, max_iter);
        CHAOS_LS(ls.after((fin ? it - o.skipped() : it) - ls.it0);)
        tested = __any_sync(CHAOS_FULL_MASK, busy && !fin && o.wants_tested());
        if (!take_scheduling_pass(fin, busy, a.sched_idle_lanes_rounds, waited) && !first) continue;
        first = false;
        CHAOS_LS(ls.pass();)

        if (a.warp_trace) tr_passes += 1;
        /* (2) retire finished orbits into their slot (:125-127) */
        uint32_t touched = __reduce_or_sync(CHAOS_FULL_MASK, fin ? (1u << slot) : 0u);   /* slots that got a result (or a new tile) */
        if (fin) {
            fin = false;
            const uint32_t et = o.finish(it, max_iter);
            const uint32_t j = rnd % R;
            if (rnd < R) ws.et[slot][rnd][pix] = et;
            else ws.sum[slot][pix] += et;                   /* rounds >= R run one at a time, straight into the sum */
            shared_add64(&ws.hdr[slot].iters[j], it);
            const uint32_t sk = o.skipped();
            if (sk) shared_add64(&ws.hdr[slot].skipped[j], sk);
            atomicSub(&ws.hdr[slot].left[j], 1u);
            busy = false;
        }
        __syncwarp();

        /* (3) per slot: decide every round that is complete and next in order (:128-150); issue further rounds, or
         *     finish the tile and take a new one */
        for (int k = 0; k < K; ++k) {
          for (;;) {
            while ((touched >> k) & 1u) {                   /* only a slot that got a result can have completed a round */
                const refill_slot_hdr &hk = ws.hdr[k];
                const uint32_t act = hk.active, i = hk.dec, issued = hk.issued, inb = hk.inb;
                uint32_t S = hk.S;
                const uint32_t j = i % R;
                const bool complete = act && i < issued && hk.left[j] == 0u && hk.pend[j] == 0u;
                const unsigned long long r_iters = hk.iters[j], r_skipped = hk.skipped[j];
                __syncwarp();                               /* every lane has read the header before lane 0 rewrites it */
                if (!complete) break;
                const bool part = (inb >> lane) & 1u;
                if (i < R && part) ws.sum[k][lane] += ws.et[k][i][lane];
                const uint32_t sum = ws.sum[k][lane];
                /* (round 1 came with the tile: pass A ran and counted it) */
                if (lane == 0 && !(kResume && i == 1u)) { iters += r_iters; skipped += r_skipped; nsamples += (unsigned long long)__popc(inb); }
                bool blocked = false;                       /* some pixel rules out even the loosest stop rule */
                if (decision_entered(adaptive, i, S)) {
                    vote_preds p = {true, true, true, false};
                    if (part) {
                        float sm[CHAOS_ADAPTIVE_THRESHOLD];
#pragma unroll
                        for (uint32_t q = 0; q < CHAOS_ADAPTIVE_THRESHOLD; ++q) sm[q] = (q <= i) ? __uint2float_rn(ws.et[k][q][lane]) : 0.f;
                        p = decision_preds(sm, i, sum);
                    }
                    const bool all_eq = __all_sync(CHAOS_FULL_MASK, p.eq);
                    const bool all_lt = __all_sync(CHAOS_FULL_MASK, p.lt);
                    const bool all_le = __all_sync(CHAOS_FULL_MASK, p.le);
                    S = decision_update(i, S, all_eq, all_lt, all_le);
                    /* after sample 1 every dispersion is a division by zero, so the vote says nothing then; a pixel whose
                     * mean is 0 (all its samples 0: inside the set for modules that report 0 there) vetoes at every i */
                    blocked = __any_sync(CHAOS_FULL_MASK, p.zero_mean && part) || (i >= 2u && !all_le);
                }
                /* A tile that looks set to use its whole budget leaves here: pass C runs its remaining rounds as independent
                 * orbits of one GPU-wide pool (a tile of 32 never-ending pixels is 7 x 32 full-length orbits -- seven waves
                 * of this warp's lanes, the whole launch's critical path when it stays here), pass D replays the decisions. */
                uint32_t e = 0xffffffffu;
                if (kResume && i + 1u < S && (blocked || export_everything) && S <= CHAOS_EXPORT_ROUNDS && issued == i + 1u && i >= 1u && a.exp.capacity) {
                    if (lane == 0) e = atomicAdd(&a.counters->n_exported, 1u);
                    e = __shfl_sync(CHAOS_FULL_MASK, e, 0);
                    if (e >= a.exp.capacity) e = 0xffffffffu;
                }
                if (e != 0xffffffffu) {
                    for (uint32_t q = 0; q <= i; ++q) export_et(a, e, q)[lane] = part ? ws.et[k][q][lane] : 0u;
                    if (lane == 0) {
                        a.exp.tile[e] = hk.tile; a.exp.first[e] = i + 1u;
                        if (a.late_tiles) {
                            const uint32_t gt = (hk.y0 >> 2) * a.tiles_x + (hk.x0 >> 3);
                            atomicOr(&a.late_tiles[gt >> 5], 1u << (gt & 31u));
                        }
                    }
                    if (lane >= i + 1u && lane < CHAOS_EXPORT_ROUNDS) {
                        a.exp.iters[(size_t)e * CHAOS_EXPORT_ROUNDS + lane] = 0ull;
                        a.exp.skipped[(size_t)e * CHAOS_EXPORT_ROUNDS + lane] = 0ull;
                    }
                    __syncwarp();
                    if (lane < sizeof(refill_slot_hdr) / 4u) reinterpret_cast<uint32_t *>(&ws.hdr[k])[lane] = 0u;
                    for (uint32_t w = 32u + lane; w < sizeof(refill_slot_hdr) / 4u; w += 32u) reinterpret_cast<uint32_t *>(&ws.hdr[k])[w] = 0u;
                    __syncwarp();
                    break;
                }
                if (i + 1u < S) {
                    /* rounds certain to run, then rounds ahead of their turn; a round >= R only when it is the next one */
                    uint32_t upto = min(next_decision_round(adaptive, i, S) + 1u, S);
                    if (blocked) upto = S;
                    upto = max(min(upto, R), i + 2u);
                    upto = max(upto, issued);
                    if (lane == 0) {
                        refill_slot_hdr &h = ws.hdr[k];
                        h.iters[j] = 0ull; h.skipped[j] = 0ull;
                        h.dec = i + 1u;
                        h.S = S;
                        uint32_t rm = h.rmask;
                        for (uint32_t r = issued; r < upto; ++r) {
                            h.pend[r % R] = inb;
                            h.left[r % R] = (uint32_t)__popc(inb);
                            rm |= 1u << (r % R);
                        }
                        h.rmask = rm;
                        h.issued = upto;
                    }
                    __syncwarp();
                } else {
                    if (part)
                        store_record(record_at(a.out, a.out_pitch, hk.x0 + (lane & 7u), hk.y0 + (lane >> 3)),
                                     __uint2float_rn(sum / S), __uint2float_rn(S), 0u, 0.f);
                    if (busy && slot == (uint32_t)k) busy = false;      /* rounds started ahead of their turn: abandoned */
                    __syncwarp();
                    if (lane < sizeof(refill_slot_hdr) / 4u) reinterpret_cast<uint32_t *>(&ws.hdr[k])[lane] = 0u;
                    for (uint32_t w = 32u + lane; w < sizeof(refill_slot_hdr) / 4u; w += 32u) reinterpret_cast<uint32_t *>(&ws.hdr[k])[w] = 0u;
                    __syncwarp();
                    break;
                }
            }
            bool loaded = false;
            if (!ws.hdr[k].active && !queue_empty) {
                uint32_t t = 0;
                if (lane == 0) {
                    if (kResume) {
                        /* tile_order runs from the most to the least expensive tile.  Slot 0 draws from the expensive end,
                         * the other slots from the cheap end: a warp then works on ONE heavy tile at a time (up to 7 x 32
                         * full-length orbits: seven waves of its 32 lanes) next to light ones, instead of four heavy tiles
                         * at once while other warps run dry -- measured on c2: SM busy time 3.9 .. 8.1 Mcycles before. */
                        if (atomicAdd(&a.counters->claimed_b, 1u) >= n_cont) t = 0xffffffffu;
                        else if (k == 0) t = a.tile_order[atomicAdd(&a.counters->next_tile_b, 1u)];
                        else t = a.tile_order[n_cont - 1u - atomicAdd(&a.counters->tail_tile_b, 1u)];
                    } else {
                        t = atomicAdd(&a.counters->next_tile, 1u);
                    }
                }
                t = __shfl_sync(CHAOS_FULL_MASK, t, 0);
                if (t >= a.n_tiles) {
                    queue_empty = true;
                    if (a.warp_trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_dry));
                } else {
                    tr_tiles += 1;
                    uint32_t x0, y0;
                    tile_origin(a, t, x0, y0);
                    const bool in = (x0 + (lane & 7u)) < a.width && (y0 + (lane >> 3)) < a.height;
                    const uint32_t inb = __ballot_sync(CHAOS_FULL_MASK, in);
                    uint32_t first_round = 0u, upto = 1u;
                    if (kResume) {           /* samples 0 and 1 were taken by pass A; their escape times sit in the record */
                        uint32_t et0 = 0u, et1 = 0u;
                        if (in) {
                            const float4 rec = *reinterpret_cast<const float4 *>(record_at(a.out, a.out_pitch, x0 + (lane & 7u), y0 + (lane >> 3)));
                            et0 = __float_as_uint(rec.x); et1 = __float_as_uint(rec.z);
                        }
                        ws.sum[k][lane] = et0;             /* round 1 is added when it is decided, below */
                        ws.et[k][0][lane] = et0;
                        ws.et[k][1][lane] = et1;
                        first_round = 1u;
                        upto = 2u;
                        loaded = true;
                    } else {
                        ws.sum[k][lane] = 0u;
                    }
                    if (lane == 0) {
                        refill_slot_hdr &h = ws.hdr[k];
                        h.tile = t; h.x0 = x0; h.y0 = y0; h.S = S0; h.dec = first_round; h.issued = min(upto, max(S0, first_round + 1u)); h.inb = inb;
                        uint32_t rm = 0u;
                        for (uint32_t r = first_round; r < h.issued && !kResume; ++r) {   /* (kResume: round 1 is complete already) */
                            h.pend[r % R] = inb;
                            h.left[r % R] = (uint32_t)__popc(inb);
                            rm |= 1u << (r % R);
                        }
                        h.rmask = rm;
                        h.active = 1u;
                    }
                }
            }
            __syncwarp();
            if (!loaded) break;
            touched |= 1u << k;                            /* round 1 came with the tile: decide it right away */
          }
        }

        /* (4) refill: idle lanes take pending orbits, from any slot, earliest round first */
        for (int k = 0; k < K; ++k) {
            uint32_t idle = __ballot_sync(CHAOS_FULL_MASK, !busy);
            if (!idle) break;
            const uint32_t dec = ws.hdr[k].dec, issued = ws.hdr[k].issued;
            if (!ws.hdr[k].rmask) continue;
            for (uint32_t r = dec; r < issued && idle; ++r) {
                const uint32_t j = r % R;
                const uint32_t pend = ws.hdr[k].pend[j];
                if (!pend) continue;
                const uint32_t rank = __popc(idle & lanemask_lt());
                const bool take = !busy && rank < (uint32_t)__popc(pend);
                uint32_t mypix = 0;
                if (take) {
                    mypix = (pend == CHAOS_FULL_MASK && idle == CHAOS_FULL_MASK) ? lane : nth_set_bit(pend, rank);
                    slot = k;
                    pix = mypix;
                    rnd = r;
                    Real cx, cy;
                    const Real dx = s_dx[r], dy = s_dy[r];
                    fm.template plane_point<fused_plane_y<FractalT>::value>(ws.hdr[k].x0 + (mypix & 7u), ws.hdr[k].y0 + (mypix >> 3), dx, dy, cx, cy);
                    o.start(cx, cy, ctx);
                    it = 0;
                    busy = true;
                    tested = true;                           /* new orbits start with a tested block */
                }
                const uint32_t taken = __reduce_or_sync(CHAOS_FULL_MASK, take ? (1u << mypix) : 0u);
                __syncwarp();
                if (lane == 0) {
                    ws.hdr[k].pend[j] = pend & ~taken;
                    if ((pend & ~taken) == 0u) ws.hdr[k].rmask &= ~(1u << j);
                }
                __syncwarp();
                idle = __ballot_sync(CHAOS_FULL_MASK, !busy);
            }
        }
        tested = __any_sync(CHAOS_FULL_MASK, tested);

        /* (5) nothing running after a full scheduling pass = no tile left anywhere for this warp */
        if (!__any_sync(CHAOS_FULL_MASK, busy)) break;
    }
    if (a.warp_trace && lane == 0) {
        unsigned long long tr_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_end));
        unsigned long long *w = a.warp_trace + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8u;
        w[0] = tr_start; w[1] = tr_dry; w[2] = tr_end; w[3] = tr_blocks; w[4] = tr_passes; w[5] = tr_tiles; w[6] = tr_lane_blocks; w[7] = nsamples;
    }
    flush_counters(a, iters, nsamples, skipped);
    CHAOS_LS(ls.flush(a, 1);)
}

/* pass B as a kernel body: one slot store per warp in dynamic shared memory */
template <class Real, class FractalT>
static __device__ __forceinline__ void render_pass_b(const chaos_render_args &a)
{
    extern __shared__ __align__(16) unsigned char chaos_dyn_smem[];
    refill_warp_store *stores = reinterpret_cast<refill_warp_store *>(chaos_dyn_smem);
    render_main_rounds<Real, FractalT>(a, stores[threadIdx.x >> 5]);
}

/* ---- pass D: the decisions of an exported tile, replayed over its stored rounds (:128-150) ----------------- */
static __device__ void replay_exported(const chaos_render_args &a)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = min(a.counters->n_exported, a.exp.capacity);
    const bool adaptive = (a.flags & CHAOS_FLAG_ADAPTIVE_SS) != 0u;
    const uint32_t S0 = min(64u, __float2uint_rz(roundf(a.max_ss)));
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long iters = 0, nsamples = 0, skipped = 0;   /* lane 0 carries the warp's totals */
    for (uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < n; e += warps) {
        uint32_t x0, y0;
        tile_origin(a, a.exp.tile[e], x0, y0);
        const uint32_t px = x0 + (lane & 7u), py = y0 + (lane >> 3);
        const bool part = px < a.width && py < a.height;
        const uint32_t n_part = __popc(__ballot_sync(CHAOS_FULL_MASK, part));
        const uint32_t first = a.exp.first[e];
        float sm[CHAOS_ADAPTIVE_THRESHOLD];
        uint32_t sum = 0;
#pragma unroll
        for (uint32_t q = 0; q < CHAOS_ADAPTIVE_THRESHOLD; ++q) {
            const uint32_t v = (part && q < S0) ? export_et(a, e, q)[lane] : 0u;
            sm[q] = __uint2float_rn(v);
            if (q < first) sum += v;
        }
        uint32_t S = S0;
        for (uint32_t i = first; ; ++i) {
            sum += part ? export_et(a, e, i)[lane] : 0u;
            if (lane == 0) {
                iters += a.exp.iters[(size_t)e * CHAOS_EXPORT_ROUNDS + i];
                skipped += a.exp.skipped[(size_t)e * CHAOS_EXPORT_ROUNDS + i];
                nsamples += n_part;
            }
            if (decision_entered(adaptive, i, S)) {
                vote_preds p = {true, true, true, false};
                if (part) {
                    float sq[CHAOS_ADAPTIVE_THRESHOLD];
#pragma unroll
                    for (uint32_t q = 0; q < CHAOS_ADAPTIVE_THRESHOLD; ++q) sq[q] = (q <= i) ? sm[q] : 0.f;
                    p = decision_preds(sq, i, sum);
                }
                const bool all_eq = __all_sync(CHAOS_FULL_MASK, p.eq);
                const bool all_lt = __all_sync(CHAOS_FULL_MASK, p.lt);
                const bool all_le = __all_sync(CHAOS_FULL_MASK, p.le);
                S = decision_update(i, S, all_eq, all_lt, all_le);
            }
            if (i + 1u >= S) break;
        }
        if (part) store_record(record_at(a.out, a.out_pitch, px, py), __uint2float_rn(sum / S), __uint2float_rn(S), 0u, 0.f);
    }
    flush_counters(a, iters, nsamples, skipped);
}

/* ---- cross-GPU stealing: the owner's side ---------------------------------------------------------------------- */
/* After its own launch a rank waits until every pixel of its tiles has been finished by somebody: own_done counts the
 * orbits that ended here, foreign_done those the other ranks finished (they add to it with system-scope atomics after
 * their record stores).  One thread; bounded: a rank that died leaves the frame void, not hanging. */
static __device__ void wait_foreign(const chaos_render_args &a)
{
    if (blockIdx.x != 0 || threadIdx.x != 0 || a.steal_world <= 1u) return;
    const volatile unsigned long long *own = &a.counters->own_done, *foreign = &a.counters->foreign_done;
    for (uint32_t spins = 0; *own + *foreign < a.own_pixels; ++spins) {
        if (spins >= (1u << 26)) { a.counters->abort = 1u; break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

/* ---- cost classes between the two passes ------------------------------------------------------ */
/* Expected cost of a tile's remaining rounds from what pass A left in its records: do the pixels' trip counts agree
 * (then the i == 1 vote will most likely end the tile after one more round), and how many trips did pass A EXECUTE in
 * the tile (with exact recurrence a never-ending orbit may have cost a few hundred trips or all of maxIterations, and
 * only the latter kind makes a tile expensive):
 *   longest orbit x (S0 - 1) rounds if the pixels disagree (the tile will probably use its whole sample budget),
 *   x 1 if all agree (the i == 1 vote will most likely end it after one more round).
 * (Also ranking a tile by its 8 neighbours' longest orbit -- to catch boundary tiles whose own sample-0 orbits all
 * escaped -- was measured: it made c2 slower, 19.0 vs 17.7 ms, and is not done.)
 * class = 36 - floor(log2(est)), class 0 is taken first by pass B.  One thread per tile; the histogram is
 * accumulated per block in shared memory (one global atomic per class and block). */
static __device__ void classify_tiles(const chaos_render_args &a)
{
    __shared__ uint32_t hist[CHAOS_COST_BUCKETS];
    for (uint32_t k = threadIdx.x; k < CHAOS_COST_BUCKETS; k += blockDim.x) hist[k] = 0u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t S0 = min(64u, __float2uint_rz(roundf(a.max_ss)));
    const bool adaptive = (a.flags & CHAOS_FLAG_ADAPTIVE_SS) != 0u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < a.n_tiles; t += warps) {   /* one warp per tile, lane = pixel */
        uint32_t x0, y0;
        tile_origin(a, t, x0, y0);
        const uint32_t px = x0 + (lane & 7u), py = y0 + (lane >> 3);
        const bool part = px < a.width && py < a.height;
        uint32_t trips = 0, cost = 0, et0 = 0, et1 = 0;
        if (part) {
            const float4 rec = *reinterpret_cast<const float4 *>(record_at(a.out, a.out_pitch, px, py));   /* pass A: (et0, trips0, et1, cost0) */
            et0 = __float_as_uint(rec.x);
            trips = __float_as_uint(rec.y);
            et1 = __float_as_uint(rec.z);
            cost = min(__float_as_uint(rec.w), 1u << 26);
        }
        /* the decision after sample 1 (:128-150), exactly as pass B takes it: most tiles of a frame end here, and those get
         * their final record now; only the others go to pass B */
        uint32_t S = S0;
        bool blocked = false;        /* some pixel's mean is 0: the tile looks set to use its whole budget (see render_main_rounds) */
        if (decision_entered(adaptive, 1u, S)) {
            /* decision_preds at i == 1, written out: the dispersion over ONE sample divides by (uint) i - 1 = 0, so it is +inf or
             * NaN for every pixel and neither `disp < 0.01` nor `disp <= 1` can hold -- only the first rule (:140-142, both
             * samples equal) can end the tile here.  (chaosPassB / pass D evaluate the general form for i >= 2.) */
            const float s0 = __uint2float_rn(et0), s1 = __uint2float_rn(et1);
            const bool eq = !part || fabsf(__fsub_rn(s0, s1)) < FLT_EPSILON;
            const bool zero_mean = part && __uint2float_rn((et0 + et1) / 2u) == 0.f;
            if (__all_sync(CHAOS_FULL_MASK, eq)) S = 2u;
            blocked = __any_sync(CHAOS_FULL_MASK, zero_mean);
            /* Also set to go on: a pixel whose first two samples lie so far apart that its dispersion after sample 2 is above 1
             * whatever sample 2 is (variance >= d^2 / 2 for d = et0 - et1, mean <= (et0 + et1 + maxIterations) / 3).  Pass B
             * would export such a tile after one more round; it may as well go now.  (A scheduling hint: pass D replays
             * the reference's decisions whatever was guessed here.) */
            const float d = __uint2float_rn(et0 > et1 ? et0 - et1 : et1 - et0);
            blocked = blocked || __any_sync(CHAOS_FULL_MASK, part && 3.0f * d * d > 2.0f * (__uint2float_rn(et0 + et1) + __uint2float_rn(a.max_iter)));
        }
        if (2u >= S) {
            if (part) store_record(record_at(a.out, a.out_pitch, px, py), __uint2float_rn((et0 + et1) / S), __uint2float_rn(S), 0u, 0.f);
            if (lane == 0) a.tile_key[t] = CHAOS_TILE_DONE;
            continue;
        }
        /* a tile that goes on (S is still S0 then) and is set to use its whole budget skips pass B: its remaining rounds go
         * to pass C as independent orbits, its decisions to pass D */
        if (blocked && S <= CHAOS_EXPORT_ROUNDS && a.exp.capacity) {
            uint32_t e = 0u;
            if (lane == 0) e = atomicAdd(&a.counters->n_exported, 1u);
            e = __shfl_sync(CHAOS_FULL_MASK, e, 0);
            if (e < a.exp.capacity) {
                export_et(a, e, 0u)[lane] = part ? et0 : 0u;
                export_et(a, e, 1u)[lane] = part ? et1 : 0u;
                if (lane == 0) {
                    a.exp.tile[e] = t; a.exp.first[e] = 2u;
                    a.tile_key[t] = CHAOS_TILE_DONE;
                    if (a.late_tiles) {
                        const uint32_t gt = (y0 >> 2) * a.tiles_x + (x0 >> 3);
                        atomicOr(&a.late_tiles[gt >> 5], 1u << (gt & 31u));
                    }
                }
                if (lane >= 2u && lane < CHAOS_EXPORT_ROUNDS) {
                    a.exp.iters[(size_t)e * CHAOS_EXPORT_ROUNDS + lane] = 0ull;
                    a.exp.skipped[(size_t)e * CHAOS_EXPORT_ROUNDS + lane] = 0ull;
                }
                continue;
            }
        }
        const uint32_t first = __shfl_sync(CHAOS_FULL_MASK, trips, __ffs(__ballot_sync(CHAOS_FULL_MASK, part)) - 1);
        const bool uniform = __all_sync(CHAOS_FULL_MASK, !part || trips == first);
        const uint32_t total = __reduce_add_sync(CHAOS_FULL_MASK, cost);          /* <= 32 x 2^26 */
        if (lane == 0) {
            const unsigned long long est = (unsigned long long)(total | 1u) * (uniform ? 1u : S - 1u);
            const uint32_t key = (uint32_t)__clzll((long long)est) - 27u;      /* est < 2^37: clzll in [27,63] -> key in [0,36] */
            a.tile_key[t] = key;
            atomicAdd(&hist[key], 1u);
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < CHAOS_COST_BUCKETS; k += blockDim.x)
        if (hist[k]) atomicAdd(&a.counters->bucket_count[k], hist[k]);
}
/* Few tiles left after the decision on sample 1 (a frame whose tiles mostly ended there): every one of them is exported
 * (see render_main_rounds: pass B would be all latency).  Exporting needs nothing of pass B's machinery -- the tile's two
 * escape times go to the export arrays exactly as chaosClassifyTiles does it for the tiles it exports itself -- so a light
 * kernel does it, one warp per tile, and pass B, launched right after, finds nothing to do and ends at once. */
static __device__ void export_all_tiles(const chaos_render_args &a)
{
    const uint32_t n_cont = a.counters->n_continuing;
    if (n_cont > a.export_all_below || !a.exp.capacity) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n_cont; k += warps) {
        const uint32_t t = a.tile_order[k];
        uint32_t x0, y0;
        tile_origin(a, t, x0, y0);
        const uint32_t px = x0 + (lane & 7u), py = y0 + (lane >> 3);
        const bool part = px < a.width && py < a.height;
        uint32_t et0 = 0u, et1 = 0u;
        if (part) {
            const float4 rec = *reinterpret_cast<const float4 *>(record_at(a.out, a.out_pitch, px, py));
            et0 = __float_as_uint(rec.x); et1 = __float_as_uint(rec.z);
        }
        uint32_t e = 0u;
        if (lane == 0) e = atomicAdd(&a.counters->n_exported, 1u);
        e = __shfl_sync(CHAOS_FULL_MASK, e, 0);
        if (e >= a.exp.capacity) continue;           /* (cannot happen: the arrays hold every tile of the launch) */
        export_et(a, e, 0u)[lane] = et0;
        export_et(a, e, 1u)[lane] = et1;
        if (lane == 0) {
            a.exp.tile[e] = t; a.exp.first[e] = 2u;
            if (a.late_tiles) {
                const uint32_t gt = (y0 >> 2) * a.tiles_x + (x0 >> 3);
                atomicOr(&a.late_tiles[gt >> 5], 1u << (gt & 31u));
            }
        }
        if (lane >= 2u && lane < CHAOS_EXPORT_ROUNDS) {
            a.exp.iters[(size_t)e * CHAOS_EXPORT_ROUNDS + lane] = 0ull;
            a.exp.skipped[(size_t)e * CHAOS_EXPORT_ROUNDS + lane] = 0ull;
        }
    }
}

/* counting-sort scatter: one thread per tile; every block reserves one range per class */
static __device__ void order_tiles(const chaos_render_args &a)
{
    __shared__ uint32_t base[CHAOS_COST_BUCKETS], cnt[CHAOS_COST_BUCKETS], off[CHAOS_COST_BUCKETS];
    for (uint32_t k = threadIdx.x; k < CHAOS_COST_BUCKETS; k += blockDim.x) cnt[k] = 0u;
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (uint32_t j = 0; j < CHAOS_COST_BUCKETS; ++j) { base[j] = acc; acc += a.counters->bucket_count[j]; }
        if (blockIdx.x == 0) a.counters->n_continuing = acc;
    }
    __syncthreads();
    /* a block owns a contiguous chunk of tiles */
    const uint32_t per_block = (a.n_tiles + gridDim.x - 1u) / gridDim.x;
    const uint32_t t0 = blockIdx.x * per_block, t1 = min(a.n_tiles, t0 + per_block);
    for (uint32_t t = t0 + threadIdx.x; t < t1; t += blockDim.x)
        if (a.tile_key[t] != CHAOS_TILE_DONE) atomicAdd(&cnt[a.tile_key[t]], 1u);
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < CHAOS_COST_BUCKETS; k += blockDim.x) {
        off[k] = cnt[k] ? base[k] + atomicAdd(&a.counters->bucket_cursor[k], cnt[k]) : 0u;
        cnt[k] = 0u;
    }
    __syncthreads();
    for (uint32_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        const uint32_t key = a.tile_key[t];
        if (key != CHAOS_TILE_DONE) a.tile_order[off[key] + atomicAdd(&cnt[key], 1u)] = t;
    }
}

#endif
