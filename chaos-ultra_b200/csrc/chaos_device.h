/*
 * chaos_device.h -- the launch contract between the host library (chaos_abi.cpp) and a fractal
 * module (one cubin per fractal, built from fractals/<name>.cu + render_generic.cuh).
 *
 * It replaces the positional kernel parameter lists the reference marshals from Java
 * (cudarenderer/RenderingKernel.java:21-25, KernelMain.java:19-20, KernelAdvanced.java:26-29,
 * KernelCompose.java:24-33): every kernel takes ONE by-value struct.  Entry names stay the
 * reference's (FractalRenderingModule.java:91-97) so a module is still found by name.
 */
#ifndef CHAOS_DEVICE_H
#define CHAOS_DEVICE_H

#include <stdint.h>
#ifndef __CUDACC__
struct uint2 { unsigned int x, y; };   /* host side of the launch contract (vector_types.h is CUDA-only) */
#endif

#define CHAOS_MODULE_ABI 37u

/* helpers.cuh:106-130 -- the 16-byte record both frame buffers hold */
struct chaos_pixel_info {
    float value;
    float weight;
    uint32_t is_reused; /* byte 0 = isReused, bytes 1..3 = padding (written as 0 here) */
    float weight_of_new_samples;
};

/* fractalRendererGeneric.cu:157-161 */
#define CHAOS_FLAG_ADAPTIVE_SS (1u << 0)
#define CHAOS_FLAG_FOVEATION (1u << 2)
#define CHAOS_FLAG_SAMPLE_REUSE (1u << 3)
#define CHAOS_FLAG_IS_ZOOMING (1u << 4)
#define CHAOS_FLAG_ZOOMING_IN (1u << 5)

/* count-preserving shortcuts of the escape loop (quadratic.cuh items 3 and 4) */
#define CHAOS_SHORTCUT_DEFER_TEST (1u << 0)  /* test the escape condition once per group of trips, replay on failure */
#define CHAOS_SHORTCUT_RECURRENCE (1u << 1)  /* an orbit whose state recurs bit for bit is reported as never escaping */
#define CHAOS_SHORTCUT_DENSE_COMPARE (1u << 2)  /* ... and is compared with its kept state every 8 trips, not only at group ends (set by the
                                                  * host unless the renderer's previous frame proved next to nothing: quadratic.cuh) */

/* device counters, one block per renderer (zeroed by the host before each render call) */
#define CHAOS_MAX_PEERS 8          /* ranks of one NVSwitch domain */
#define CHAOS_COST_BUCKETS 37
#define CHAOS_POOL_SHARDS 128   /* the orbit pool is this many independent rings (warp w uses ring w % shards) */
/* engine 2 (render_streams.cuh): the long and finish lists of one probe -> long -> finish chain */
struct chaos_stream_ctl {
    unsigned int n_long;          /* orbits the probe appended to the long list (may exceed the capacity: clamp) */
    unsigned int long_cursor;     /* entries the long kernel has claimed */
    unsigned int n_finish;        /* orbits the long kernel appended to the finish list */
    unsigned int n_hot;           /* orbits the probe put into the list's hot region (expected to be long: the long kernel starts them first) */
};
struct chaos_counters {
    unsigned int next_tile;             /* work-stealing cursor over vote tiles */
    unsigned int next_tile_b;           /* pass B: cursor from the expensive end of tile_order (fast frames: the sampling pass' cursor) */
    unsigned int tail_tile_b;           /* pass B: cursor from the cheap end */
    unsigned int claimed_b;             /* pass B: tiles claimed from either end */
    unsigned int n_exported;            /* pass B: tiles whose remaining rounds were handed to pass C (may exceed export.capacity: clamp) */
    unsigned int next_export_item;      /* pass C: work-stealing cursor over (exported tile, round) items */
    unsigned int n_continuing;          /* chaosOrderTiles: tiles whose decision after sample 1 did not end them = pass B's tiles */
    unsigned int abort;                 /* set by a kernel that gave up waiting for a hand-over in the orbit pool: the frame is void */
    unsigned long long pixel_iterations;
    unsigned long long samples;
    unsigned long long skipped_iterations; /* part of pixel_iterations that was proven, not executed (exact recurrence) */
    unsigned int bucket_count[CHAOS_COST_BUCKETS + 3];  /* tiles per cost class (chaosClassifyTiles) */
    unsigned int bucket_cursor[CHAOS_COST_BUCKETS + 3]; /* fill position per class (chaosOrderTiles) */
    /* orbit pool of the independent-orbit passes ([0] pass A or the single launch, [1] pass C), see chaos_render_args::pool */
    struct { unsigned int live, reserved, head, pad; } pool[2][CHAOS_POOL_SHARDS];
    /* diagnostics, modules built with -DCHAOS_LANE_STATS only (host: CHAOS_LANE_STATS=1 prints them): lane-trips of the
     * escape loop by what the lane was doing, [pass A/B/C/main][tested, untested][CHAOS_LS_*] */
    unsigned long long lane_stats[4][2][8];
    chaos_stream_ctl stream[2];         /* engine 2: [0] one-sample frame or pass A, [1] pass C */
    /* diagnostics (CHAOS_LANE_STATS modules), [pass A, pass C]: [0] warps of the long kernel alive per 65.5 us bin; [3] cycles
     * inside the loop / trips of its longest lane / calls, before ([0..2]) and after ([4..6]) the list ran dry */
    unsigned long long long_hist[2][4][32];
    /* cross-GPU tile stealing (one-sample frames, render_refill.cuh): this rank's tile cursor (next_tile) is open to the
     * other ranks once frame_seq says the counters belong to the current frame; orbits finished here / by other ranks */
    unsigned int frame_seq;
    unsigned int pad1;
    unsigned long long own_done, foreign_done;
};
#define CHAOS_LS_CAPACITY 0   /* 32 x trips the warp spent in the block */
#define CHAOS_LS_USEFUL 1     /* trips the lanes advanced */
#define CHAOS_LS_REPLAY_WAIT 2 /* lanes waiting for a tested block */
#define CHAOS_LS_FIN_WAIT 3   /* lanes with a finished orbit waiting for a scheduling pass */
#define CHAOS_LS_IDLE_QUEUE 4 /* empty lanes while the queue still had work */
#define CHAOS_LS_IDLE_DRY 5   /* empty lanes after the queue ran dry */
#define CHAOS_LS_BLOCKS 6     /* blocks */
#define CHAOS_LS_PASSES 7     /* scheduling passes */

/* Tiles that will (almost surely) use their whole sample budget leave pass B after a decision: their remaining rounds
 * are run by pass C as independent orbits of one GPU-wide pool, and pass D replays the decisions over the stored
 * escape times (render_refill.cuh).  Rounds are stored for S <= CHAOS_EXPORT_ROUNDS only. */
#define CHAOS_EXPORT_ROUNDS 10
struct chaos_export {
    uint32_t capacity;                  /* tiles the arrays below can hold (0 = no export) */
    uint32_t *tile;                     /* [capacity] which tile */
    uint32_t *first;                    /* [capacity] first round pass C computes for it (the earlier ones came with it) */
    uint32_t *et;                       /* [capacity][CHAOS_EXPORT_ROUNDS][32] escape time per round and pixel */
    unsigned long long *iters;          /* [capacity][CHAOS_EXPORT_ROUNDS] trips of the round's orbits (reference count) */
    unsigned long long *skipped;        /* [capacity][CHAOS_EXPORT_ROUNDS] of which proven, not executed */
};

struct chaos_render_args {
    chaos_pixel_info *out;  /* output records, pitch-linear */
    uint64_t out_pitch;     /* bytes */
    const chaos_pixel_info *in; /* previous frame (advanced kernels) */
    uint64_t in_pitch;
    chaos_counters *counters;
    double image[4];        /* lb.x lb.y rt.x rt.y; float kernels read imagef (host-cast, KernelMainFloat.java:19-21) */
    double image_reused[4];
    float imagef[4];
    float image_reusedf[4];
    uint32_t width, height;
    uint32_t max_iter;
    float max_ss;
    uint32_t flags;
    uint32_t focus_x, focus_y;
    float focus_d2_thr;     /* squared pixel radius of the focus area (host: (tan(thr) * 60 / 0.02652)^2), see pass R */
    uint32_t tiles_x;       /* ceil(width / 8) */
    uint32_t tile_rows;     /* ceil(height / 4) */
    /* multi-GPU row-band partition: this launch covers bands b with b % part_count == part_index */
    uint32_t part_index, part_count, band_tile_rows;
    uint32_t n_tiles;       /* vote tiles owned by this launch */
    uint32_t *tile_key;     /* [n_tiles] cost class of each tile after pass A */
    uint32_t *tile_order;   /* [n_tiles] tiles sorted by descending expected cost: the order pass B takes them in */
    uint32_t phase;         /* 0 = whole render in one launch; 1 = pass A (sample 0 of every pixel); 2 = pass B (the rest);
                             * 3 = pass C (the rounds pass B exported, as independent orbits) */
    chaos_export exp;
    uint32_t *late_tiles;   /* NULL, or one bit per vote tile of the FRAME (row-major): set by pass B for the tiles it exports -- the
                             * tiles that are not final when the frame-wide compose starts */
    uint32_t engine;        /* 0 = tile-synchronous, 1 = lane-refill scheduler */
    uint32_t force_exact;   /* 1 = always the reference's 7-operation trip (differential check) */
    uint32_t block_iters;   /* engine 1: trips between two scheduling points (multiple of 4) */
    uint32_t shortcuts;     /* CHAOS_SHORTCUT_* bits an Orbit may use; 0 with force_exact */
    unsigned long long *warp_trace;   /* NULL, or [warps][8]: per-warp timeline of the rounds engine (CHAOS_WARP_TRACE, diagnostics) */
    uint32_t sched_idle_lanes_indep;   /* engine 1: finished or empty lanes a warp lets accumulate before a scheduling pass, */
    uint32_t sched_idle_lanes_rounds;  /* independent orbits / sample rounds (1 = a pass after every block that ended an orbit) */
    /* Orbit pool (independent-orbit passes).  Once the tile queue is dry a warp's lanes empty one by one while its
     * instructions still take whole FP-pipe slots.  A warp left with fewer than pool_min_lanes running orbits parks them
     * here (state and all) and ends; warps with empty lanes take parked orbits over.  So the drain of a pass runs in few,
     * full warps, and SM share goes back to the next pass early.  NULL / 0 = off. */
    unsigned char *pool;               /* [pool_capacity][CHAOS_POOL_STRIDE] */
    uint32_t pool_capacity;            /* entries of all CHAOS_POOL_SHARDS rings together; a ring holds >= 32 x its warps */
    uint32_t pool_min_lanes;
    uint32_t export_all_below;         /* pass B exports every tile it gets when chaosClassifyTiles left it at most this many */
    uint32_t export_all_done;          /* 1 = chaosExportAll ran before pass B and took that case over */
    /* fast frame, pass R: the pixels it finishes are coloured right away (their record is in registers) instead of being
     * read back by compose: 16 R + 16 W + 4 W bytes per pixel instead of 16 R + 16 W and 16 R + 4 W.  NULL = compose does
     * it all.  The tiles pass R hands to pass S are flagged in late_tiles; a filtered compose colours them afterwards. */
    uint32_t *fuse_rgba;               /* the frame (device, mapped host or a peer's), width*height */
    const uint32_t *fuse_palette;
    uint32_t fuse_palette_len;
    uint32_t pool_epoch;               /* distinguishes this launch's entries from older ones (host: += 2 per frame; pass C uses epoch + 1) */
    /* engine 2 (render_streams.cuh) */
    void *long_list;                   /* [list_capacity] finish_item<Real>, 32 bytes apart: the orbits that outlived the probe, with their state */
    void *finish_list;                 /* [list_capacity] finish_item<Real>, 32 bytes apart: orbits that need a last group of tested trips */
    uint32_t list_capacity;
    uint32_t probe_trips;              /* tested trips an orbit gets in the probe kernel before it goes to the long list */
    /* Multi-GPU fast frames: the frame is cut into one slab of slab_rows pixel rows per rank (slab q = rows q * slab_rows ...),
     * every rank keeps the records of its own slab, and the reprojection reads the previous frame's row j from the rank that
     * owns it: in_peer[j / slab_rows], that rank's primary buffer mapped into this process (CUDA IPC; the own slab is plain
     * local memory).  Under a zoom about a point a pixel's origin lies in its own slab except for a thin ring at the slab's
     * edges, so almost all taps are local and the rest are peer loads over NVLink.  0 = one buffer (`in`). */
    const chaos_pixel_info *in_peer[CHAOS_MAX_PEERS];
    uint32_t slab_rows;
    /* Long kernel: how many of its resident CTAs per SM take part.  An orbit is a chain of dependent FP64 instructions: with 8
     * warps per scheduler a trip takes ~97 cycles, with 2 it takes ~30 at 80 % of the throughput.  When the list holds few
     * orbits per lane the launch is as long as its longest orbit, so it runs on fewer warps: CTA layer k (blockIdx / sm_count)
     * takes part only if the list holds at least occ_orbits_per_lane[k-1] orbits per lane of the full grid. */
    uint32_t sm_count;
    uint32_t occ_orbits_per_lane[3];
    /* Cross-GPU work stealing, one-sample frames (the deep-zoom configuration).  Every rank's tile cursor and record buffer
     * are mapped into every other rank's process (CUDA IPC).  A rank whose own tiles are handed out goes on with the next
     * rank's: it claims tiles from that rank's cursor with system-scope atomics over NVLink, iterates them like its own and
     * stores the records straight into the owner's buffer; the owner waits (chaosWaitForeign) until as many orbits have
     * been finished as its tiles hold pixels.  steal_world = 0: off. */
    uint32_t steal_world, steal_rank, frame_seq;
    uint32_t steal_n_tiles[CHAOS_MAX_PEERS];            /* vote tiles rank q owns */
    chaos_counters *steal_counters[CHAOS_MAX_PEERS];    /* rank q's counter block (strand 0) */
    chaos_pixel_info *steal_out[CHAOS_MAX_PEERS];       /* rank q's output records */
    unsigned long long own_pixels;                      /* in-bounds pixels of this rank's tiles: what chaosWaitForeign waits for */
    uint32_t hot_capacity;             /* the last hot_capacity entries of long_list are the hot region, filled from the end (0 = none) */
    uint32_t hot_trips;                /* pass C: a pixel whose sample 0 executed at least this many trips is expected to be long */
};
#define CHAOS_POOL_STRIDE 128u

struct chaos_compose_args {
    const chaos_pixel_info *in;
    uint64_t in_pitch;
    uint32_t *out_rgba;     /* device or mapped-host pointer, width*height, row 0 = top */
    const uint32_t *palette;
    uint32_t palette_len;
    uint32_t width, height;
    float max_ss;
    uint32_t part_index, part_count, band_rows;
    /* NULL, or one bit per vote tile of the frame (row-major, tiles_x per row): compose only the tiles whose bit is set
     * (the ones pass D finished after the frame-wide compose had started) */
    const uint32_t *only_tiles;
    uint32_t tiles_x;
};

#endif
