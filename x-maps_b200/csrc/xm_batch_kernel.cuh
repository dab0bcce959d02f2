// xm_batch_kernel.cuh — ONE persistent kernel for a batch of independent frames.
//
// Frames do not depend on each other (python/depth_reprojection_pipe.py:121-167 keeps no state across
// frames), so a batch of them can be rendered by one grid that never drains between frames:
//
//   * Every CTA has two kinds of warps that never wait for each other inside a frame:
//       - the lean warp-specialised event pipeline of events_lean_kernel (one producer lane, eight
//         consumer warps, mbarrier ring fed by TMA bulk copies of the chunks and of their X-map windows).
//         All chunks of the batch form ONE ordered list handed out by a global counter; the producer takes
//         its chunks TWO ahead and reads a chunk's first / last timestamp ONE ahead;
//       - the epilogue warps (two for large frames, four for small ones).  Projector view, default: the STRIP
//         epilogue -- two barrier-free passes per frame, one warp per item (pass 1: decode + 7x7 dilation of the
//         remap targets' window into a small u16 map; pass 2: one gather per output pixel), the items of all
//         frames in one ordered list handed out by a global counter (batch_strip_warps).  Otherwise (camera view,
//         `batch_strips = 0`): groups of kTileGroupThreads threads take 32x32 output tiles (camera view: 4096
//         pixels) of frame f from that frame's ticket counter, then move on to frame f+1 (batch_tile_groups).
//         Either way an item of frame f starts as soon as every chunk of frame f has been scattered.
//     So the epilogue of a frame runs on the same SMs, at the same time, as the event stream of the next
//     frames: HBM streaming, L2 gathers and the dilation overlap instead of alternating, and there is no
//     launch, drain or pipeline fill per frame.  (A first version let the eight consumer warps execute the
//     tiles in line: a tile is a chain of four barrier-separated phases, ~8 us during which the CTA's event
//     stream stood still -- slower than separate kernels, see EXPERIMENTS_r01.md.)
//   * A tile of frame f may only read the scatter map once every chunk of frame f has been scattered.
//     Each CTA counts the chunks it finished per frame and publishes the count (after a fence) when its
//     consumers leave the frame; the tile groups spin on the frame's count.  A chunk only ever waits for
//     tiles of an earlier frame and a tile only for chunks of its own frame, and chunks are handed out in
//     frame order, so this cannot deadlock.
//   * Frames rotate through n_maps scatter maps (3 ... 6, chosen per launch; epoch-tagged keys as everywhere
//     else), so the events of the next frames never disturb the cells frame f's epilogue still has to read; the
//     first chunk of frame f + n_maps a warp sees waits for frame f's finished-tile (strips: pass-1) count.
//   * Time bounds: batch_bounds_kernel (one tiny launch per batch) looks up first / last valid event
//     of every frame.  Events outside the assumed bounds flag their frame; batch_redo_kernel (one tiny
//     launch per batch) re-renders flagged frames exactly (device-side tail launches of the two-pass
//     kernels), so unsorted input stays bit-exact without any host round trip.
#pragma once
#include "xm_fused_kernel.cuh"

namespace xm {

constexpr int kBatchMax = 32;    // frames per launch (the frame table travels in the kernel parameters)
#ifndef XM_BATCH_MAPS
#define XM_BATCH_MAPS 3
#endif
constexpr int kBatchMaps = XM_BATCH_MAPS;  // scatter maps in rotation (default; BatchParams::n_maps is what a launch uses)
constexpr int kBatchMapsMax = 8;
constexpr int kBatchDilMaps = 4;           // strip epilogue: dilated maps in rotation
constexpr int kBatchHeader = 1152;  // mbarriers, ring descriptors, CTA accumulators, per-warp frame constants (2 slots)
constexpr int kCamTilePx = 4096;  // camera-view epilogue item
#ifndef XM_TILE_WARPS
#define XM_TILE_WARPS 4
#endif
#ifndef XM_POLL_NS0  // back-off of a tile group leader waiting for a frame's events: first / longest sleep
#define XM_POLL_NS0 200
#define XM_POLL_NS1 1600
#endif
#ifndef XM_AGG_ABOVE
#define XM_AGG_ABOVE 100
#endif
#ifndef XM_BATCH_CTAS
#define XM_BATCH_CTAS 2
#endif
constexpr int kTileWarps = XM_TILE_WARPS;  // epilogue warps per CTA
constexpr int kBatchCtasPerSm = XM_BATCH_CTAS;  // resident CTAs per SM the kernel is compiled for (register budget)
constexpr int kTileGroupThreads = 64;     // threads that share one tile
constexpr int kTileGroups = kTileWarps * 32 / kTileGroupThreads;
constexpr int kBatchThreads = kWsThreads + kTileWarps * 32;

struct BatchFrame {
    const int4* events;
    void* dst;
    long long n;
};

struct BatchParams {
    // per-event tables / geometry (as EventParams)
    int polarity;
    const int* lut_xy;
    int cam_w, cam_h;
    const short* xmap_t;
    int xmap_w, xmap_h, col_stride;
    int t_px_scale, x_offset;
    int rect_w, rect_h;
    int cap_cols, stages, win_stages;
    // "alive" table: per block of 2^s x 2^s camera pixels the range of (quantised) time columns in which an event of a
    // pixel of the block CAN be an inlier (derived exactly from the LUT and the X-map at table upload: the union over
    // the block's pixels of [first, last] column with disparity >= 0; a pixel's inlier columns are one interval for any
    // monotonic X-map row, and a conservative hull otherwise).  Events outside their block's range -- 69 % of a uniform
    // stream, of which 70 % are no inliers -- are only counted and bounds-checked: no LUT gather, no X-map lookup, no
    // scatter.  Entry (u16) = lo | (hi - lo) << 8 in units of 2^alive_qs columns, dead block = 255 | 0 << 8 (no column
    // reaches 255 units); entry (bx, by) at by * alive_pitch + bx, the index clamped so that any 16-bit coordinate
    // pair reads inside the table.
    const unsigned* alive;   // the table, as 32-bit words for the copy into shared memory
    int alive_words;         // words to copy
    int alive_shift, alive_pitch, alive_qs;
    unsigned alive_last;     // last entry (index clamp)
    unsigned long long* maps[kBatchMapsMax];
    int n_maps;             // maps in rotation: frame f scatters into maps[f % n_maps]
    unsigned epoch0;        // frame f scatters with epoch0 + f
    FrameState* states;     // [n_frames + 1]; block n_frames is the control block (next_chunk = item counter)
    // epilogue (ep.map / ep.dst / ep.state are set per tile)
    EpilogueParams ep;
    int tiles_x, tile_items;  // items per frame epilogue (strip epilogue: pass-1 items)
    // strip epilogue (projector view; see xm_frame_kernels.cuh): every epilogue WARP works on its own, no shared memory
    int strips;               // 1: on
    int p2_items;             // pass-2 items per frame
    StripWindow win;          // window of the rectified image pass 1 produces (bounding box of the remap targets)
    int strip_lag;            // blocks the pass-2 items of a frame trail its pass-1 items by in the item list (1 ... kBatchDilMaps - 1)
    const unsigned* pix_cell;  // per output pixel: its remap target's cell in the rectified image, 0xffffffff = none
    unsigned short* dil[kBatchDilMaps];  // dilated disparity maps in rotation (frame f uses f % kBatchDilMaps)
    int n_frames;
    unsigned long long* dbg;  // XM_DEBUG_HOOKS + debug & 8: per frame [0] first consumer enters, [1] last chunk count published,
                              // [2] first tile group sees the frame complete, [3] last tile finished (global timer, ns)
    int debug;  // timing experiments only (results WRONG): 16 = skip the epilogue work of tile items
    int hard_frames;  // 1: publish a frame's chunk count before touching the next frame (few chunks per CTA and frame:
                      //    the tiles would otherwise wait for every CTA's NEXT chunk)
    unsigned total_items;
    unsigned first_item[kBatchMax + 2];  // first chunk of frame f in the batch-wide chunk list; [n_frames] = total
    BatchFrame frames[kBatchMax];
};

__host__ __device__ __forceinline__ unsigned batch_chunks(long long n) { return static_cast<unsigned>((n + kEvChunk - 1) / kEvChunk); }

// Live lists: the front half of a chunk compacts the events that can become inliers (polarity, inside the image,
// "alive" pixel block) warp by warp into dense lists -- gathered LUT word, (time column | index inside the chunk),
// and for the camera view the pixel index -- so that the back half only runs over those (~40 % of a uniform stream).
// Two buffers (front half of chunk c+1 / back half of chunk c), kEvChunk 32-bit entries per list and buffer.
constexpr int kListBytes = kEvChunk * 4;  // one list of one buffer
__host__ __device__ __forceinline__ int batch_list_bytes(bool cam) { return (cam ? 3 : 2) * 2 * kListBytes; }

inline int batch_smem_bytes(int stages, int win_stages, int win_bytes, int region_cells, int alive_words = 0, bool cam = false) {
    // header | live lists | event ring | X-map window ring | two u16 regions per tile group | alive bitmap
    return kBatchHeader + batch_list_bytes(cam) + stages * (kEvChunk * 16) + win_stages * win_bytes + kTileGroups * region_cells * 4 + alive_words * 4;
}

struct BatchBoundsParams {
    const int4* events[kBatchMax];
    long long n[kBatchMax];
    long long lo[kBatchMax], hi[kBatchMax];
    unsigned given_mask;  // bit f: bounds of frame f are lo[f] / hi[f] (else first / last valid event)
    int polarity;
    int t_px_scale;
    int n_frames;
    FrameState* states;
};

// grid = n_frames + 1 CTAs of 64 threads: clears the state blocks and sets the bounds of every frame
__global__ void __launch_bounds__(64) batch_bounds_kernel(const __grid_constant__ BatchBoundsParams p) {
    const int f = blockIdx.x;
    FrameState* st = p.states + f;
    if (threadIdx.x == 0) {
        st->red_lo = 0;
        st->red_hi = 0;
        st->n_valid = 0;
        st->n_inliers = 0;
        st->flags = 0;
        st->redo = 0;
        st->epoch_used = 0;
        st->blocks_done = 0;
        st->any_valid = 0;
        st->next_chunk = 0;
        st->next_tile = 0;
        st->fix_chunk = 0;
        st->p2_ticket = 0;
        st->p2_done = 0;
    }
    if (f >= p.n_frames) {
        if (threadIdx.x == 0) st->t_lo_bits = st->t_hi_bits = 0;
        return;
    }
    if (p.given_mask & (1u << f)) {
        if (threadIdx.x == 0) {
            st->t_lo_bits = p.lo[f];
            st->t_hi_bits = p.hi[f];
        }
    } else {
        scan_sorted_bounds(p.events[f], p.n[f], p.polarity, threadIdx.x >> 5, threadIdx.x & 31, &st->t_lo_bits);
    }
    __syncthreads();  // (both branches are uniform per block)
    if (threadIdx.x == 0) {
        IntCol ic;
        ic.init(st->t_lo_bits, st->t_hi_bits, p.t_px_scale);
        st->ic0 = make_int4(static_cast<int>(ic.lo), static_cast<int>(ic.lo >> 32), static_cast<int>(ic.range), static_cast<int>(ic.scale2));
        st->ic1 = make_int4(static_cast<int>(ic.d), static_cast<int>(ic.M), ic.sh, ic.ok ? 1 : 0);
    }
}

// release fence for the "data, then counter" hand-offs below (lighter than __threadfence(), which is fence.sc)
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// one lane of a converged warp, without reading %laneid / %tid (the compiler re-reads the special register inside
// the hot loop to save a register; the S2R round trip showed up as ~8 % of the consumers' stall samples)
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ long long ld_time_nc(const int4* ev) {
    long long t;
    asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(t) : "l"(reinterpret_cast<const char*>(ev) + 8));
    return t;
}
__device__ __forceinline__ void sts128_a(unsigned addr, const int4& v) {
    asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ int2 lds64_a(unsigned addr) {
    int2 r;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
    return r;
}

// One epilogue item of frame f, executed by one group of kTileGroupThreads threads (`tid` inside the
// group, named barrier `bar_id`).
template <bool CAM>
static __device__ __noinline__ void batch_tile(const BatchParams& bp, int f, int t, unsigned short* bufA, unsigned short* bufB, int tid,
                                               int bar_id) {
    const unsigned epoch = bp.epoch0 + static_cast<unsigned>(f);
#ifdef XM_PROBE_NO_TILE
    if (true) {
#else
    if (bp.debug & 16) {
#endif
        // timing experiment: no epilogue work (results WRONG)
    } else if (CAM) {
        const unsigned long long* mp = bp.maps[f % bp.n_maps];
        const int n_px = bp.ep.out_w * bp.ep.out_h;
        const int end = min(n_px, (t + 1) * kCamTilePx);
        for (int i = t * kCamTilePx + tid; i < end; i += kTileGroupThreads)
            emit_pixel_int(bp.ep.out, bp.frames[f].dst, i, key_disparity(__ldcg(mp + i), epoch));
    } else {
        EpilogueParams q = bp.ep;
        q.map = bp.maps[f % bp.n_maps];
        q.dst = bp.frames[f].dst;
        const int by = t / bp.tiles_x;
        proj7_tile_late<kTileGroupThreads, 6, 8>(q, t - by * bp.tiles_x, by, bp.tiles_x, epoch, bufA, bufB, tid, bar_id);
    }
}

// The loop of one epilogue group (kTileGroupThreads threads, named barrier 1 + grp): for every frame of the
// batch, wait until all of its chunks (CHUNK events each) have been scattered, then take tiles from the frame's
// ticket counter until they run out.
template <bool CAM, int CHUNK>
__device__ __forceinline__ void batch_tile_groups(const BatchParams& bp, int grp, int gtid, unsigned short* bufA, unsigned short* bufB,
                                                  volatile int* s_ticket) {
    const int bar_id = 1 + grp;
    const int n_tiles = bp.tile_items;
    for (int f = 0; f < bp.n_frames; ++f) {
        FrameState* st = bp.states + f;
        int next_ticket = 0;
        if (gtid == 0) {
            const unsigned need = static_cast<unsigned>((bp.frames[f].n + CHUNK - 1) / CHUNK);
            // (back-off: a group leader polls every 0.2 ... 1.6 us; the 3-map ring gives the tiles two frames of slack)
            spin_until_ge(&st->blocks_done, need, XM_POLL_NS0, XM_POLL_NS1);
            next_ticket = static_cast<int>(atomicAdd(&st->fix_chunk, 1u));
#ifdef XM_DEBUG_HOOKS
            if (bp.dbg) atomicMin(bp.dbg + f * 4 + 2, global_timer_ns());
#endif
        }
        for (;;) {
            group_sync<kTileGroupThreads>(bar_id);  // the previous tile is done with the buffers / the frame is complete
            if (gtid == 0) {
                *s_ticket = next_ticket;
                // the ticket after this one is requested now and read after the tile: its round trip is hidden
                if (next_ticket < n_tiles) next_ticket = static_cast<int>(atomicAdd(&st->fix_chunk, 1u));
            }
            group_sync<kTileGroupThreads>(bar_id);
            const int t = *s_ticket;
            if (t >= n_tiles) break;
            batch_tile<CAM>(bp, f, t, bufA, bufB, gtid, bar_id);
            group_sync<kTileGroupThreads>(bar_id);
            if (gtid == 0) {
                fence_acq_rel_gpu();
                atomicAdd(&st->next_tile, 1u);  // this tile no longer needs the frame's scatter map
#ifdef XM_DEBUG_HOOKS
                if (bp.dbg) atomicMax(bp.dbg + f * 4 + 3, global_timer_ns());
#endif
            }
        }
    }
}

// The loop of one epilogue WARP with the strip epilogue (projector view).  The pass-1 items (strip x row segment: decode
// + 7x7 dilation into the frame's dilated map) and pass-2 items (a run of output pixels) of ALL frames form one
// ordered list handed out by a global counter, so warps do not march through the frames in step: with small frames
// the items of several frames are in flight at the same time.  A pass-1 item waits for its frame's chunks (and for
// pass 2 of the frame that used the same dilated map), a pass-2 item for the frame's finished pass-1 count.  The
// scatter map is free as soon as pass 1 is complete (`next_tile` counts finished pass-1 items: what the event warps of
// frame f + n_maps wait for).  Every wait is for items that come earlier in the list (or for event work that only
// depends on such items) and an item is only ever held by a warp that is running it or waiting to, so this cannot
// deadlock (tests/test_batch_protocol_model.py models it, with a negative control).
template <int CHUNK>
__device__ __forceinline__ void batch_strip_warps(const BatchParams& bp, int lane) {
    constexpr unsigned kFull = 0xffffffffu;
    const int n_p1 = bp.tile_items, n_p2 = bp.p2_items;
    const unsigned per_frame = static_cast<unsigned>(n_p1 + n_p2);
    const unsigned total = per_frame * static_cast<unsigned>(bp.n_frames);
    const int n_strips = bp.win.strips;
    const int n_px = bp.ep.out_w * bp.ep.out_h;
    unsigned* const counter = &bp.states[bp.n_frames].next_tile;
    auto wait_for = [&](const unsigned* c, unsigned need) {
        spin_until_ge(c, need, XM_POLL_NS0, XM_POLL_NS1);
    };
#ifdef XM_DEBUG_HOOKS  // lane 0's cycles per phase and pass: [pass * 4 + {ticket, wait, work, publish}], [8 + pass] items
    long long acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#define XM_STRIP_CLK(v) v = clock64()
#else
#define XM_STRIP_CLK(v)
#endif
    // List order: blocks of (pass 1 of frame b, pass 2 of frame b - lag): by the time a warp reaches the pass-2 items
    // of a frame, that frame's pass 1 has had `lag` blocks' worth of items to complete, so warps spend their time on
    // items that can run instead of holding items that cannot (lag < kBatchDilMaps: pass 1 of frame f waits for pass
    // 2 of frame f - kBatchDilMaps, which must come earlier in the list).
    // (No look-ahead ticket: a warp that held its next item while working on the current one kept that item from every
    // other warp -- with small frames a serial chain pass 2 (f) -> pass 1 (f + 1) through the warps, 31 us per frame
    // instead of 14.)
    unsigned ticket = 0;
    for (;;) {
        XM_STRIP_CLK(c0);
        if (lane == 0) ticket = atomicAdd(counter, 1u);
        const unsigned t = __shfl_sync(kFull, ticket, 0);
        if (t >= total) break;
        XM_STRIP_CLK(c1);
        int f, j;  // frame, item (j < n_p1: pass 1, else pass 2 item j - n_p1)
        {
            // block b of the list = pass 1 of frame b (b < n) followed by pass 2 of frame b - lag (b >= lag)
            const int n = bp.n_frames, lag = bp.strip_lag;
            const unsigned head = static_cast<unsigned>(min(lag, n)) * static_cast<unsigned>(n_p1);  // pass-1-only blocks
            const int mixed = max(n - lag, 0);
            if (t < head) {
                f = static_cast<int>(t / static_cast<unsigned>(n_p1));
                j = static_cast<int>(t - static_cast<unsigned>(f) * static_cast<unsigned>(n_p1));
            } else if (t - head < static_cast<unsigned>(mixed) * per_frame) {
                const unsigned k = t - head;
                const int b = static_cast<int>(k / per_frame);
                const int r = static_cast<int>(k - static_cast<unsigned>(b) * per_frame);
                f = r < n_p1 ? b + lag : b;
                j = r;
            } else {  // pass-2-only blocks of the last frames
                const unsigned k = t - head - static_cast<unsigned>(mixed) * per_frame;
                const int b = static_cast<int>(k / static_cast<unsigned>(n_p2));
                f = mixed + b;
                j = n_p1 + static_cast<int>(k - static_cast<unsigned>(b) * static_cast<unsigned>(n_p2));
            }
        }
        FrameState* st = bp.states + f;
        unsigned short* const dil = bp.dil[f % kBatchDilMaps];
        if (j < n_p1) {
            if (lane == 0) {
                wait_for(&st->blocks_done, static_cast<unsigned>((bp.frames[f].n + CHUNK - 1) / CHUNK));
                if (f >= kBatchDilMaps) wait_for(&bp.states[f - kBatchDilMaps].p2_done, static_cast<unsigned>(n_p2));
#ifdef XM_DEBUG_HOOKS
                if (bp.dbg) atomicMin(bp.dbg + f * 4 + 2, global_timer_ns());
#endif
            }
            __syncwarp();
            XM_STRIP_CLK(c2);
            const int seg = j / n_strips;
            if (!(bp.debug & 16))
                strip_dilate_item(bp.maps[f % bp.n_maps], dil, bp.rect_w, bp.rect_h, bp.epoch0 + static_cast<unsigned>(f),
                                  bp.win.x0 + (j - seg * n_strips) * kStripCols, bp.win.y0 + seg * bp.win.rows,
                                  min(bp.win.y0 + (seg + 1) * bp.win.rows, bp.win.y1), lane);
            __syncwarp();
            XM_STRIP_CLK(c3);
            if (lane == 0) {
                fence_acq_rel_gpu();
                atomicAdd(&st->next_tile, 1u);
            }
        } else {
            RemapPipe rp;
            strip_remap_begin(rp, bp.pix_cell, n_px, (j - n_p1) * bp.win.blocks * kRemapBlockPx, lane);  // (static table: requested before the wait)
            if (lane == 0) wait_for(&st->next_tile, static_cast<unsigned>(n_p1));
            __syncwarp();
            XM_STRIP_CLK(c2);
            if (!(bp.debug & 16)) strip_remap_run(rp, bp.ep.out, bp.frames[f].dst, bp.pix_cell, dil, n_px, (j - n_p1) * bp.win.blocks * kRemapBlockPx, bp.win.blocks, lane);
            __syncwarp();
            XM_STRIP_CLK(c3);
            if (lane == 0) {
                fence_acq_rel_gpu();
                atomicAdd(&st->p2_done, 1u);
#ifdef XM_DEBUG_HOOKS
                if (bp.dbg) atomicMax(bp.dbg + f * 4 + 3, global_timer_ns());
#endif
            }
        }
#ifdef XM_DEBUG_HOOKS
        {
            const long long c4 = clock64();
            const int ps = j < n_p1 ? 0 : 1;
            acc[ps * 4 + 0] += c1 - c0;
            acc[ps * 4 + 1] += c2 - c1;
            acc[ps * 4 + 2] += c3 - c2;
            acc[ps * 4 + 3] += c4 - c3;
            acc[8 + ps] += 1;
        }
#endif
    }
#ifdef XM_DEBUG_HOOKS
    if (bp.dbg && lane == 0)
        for (int i = 0; i < 10; ++i) atomicAdd(bp.dbg + 256 + i, static_cast<unsigned long long>(acc[i]));
#endif
}

// The back half's work on two list entries per lane, as one block of PTX (see `back` in batch_kernel): HEAD = lookup,
// disparity, key; then the scatter -- PLAIN: one predicated RED per inlier; AGG: one RED per distinct cell of the warp
// round (match.any on the cell, the highest lane of a group holds the highest event index), which pays when many
// consecutive events share a cell (a scanning projector lights few pixels at a time) and costs 25 % on a uniform
// stream -- and TAIL = counts.
#define XM_BACK_PTX_HEAD \
"{\n\t" \
                    ".reg .pred pa0, pa1, ph0, ph1, pi0, pi1, pm0, pm1;\n\t" \
                    ".reg .b32 m0, m1, l0, l1, y0, y1, c0, c1, r0, r1, a0, a1, x0, x1, xc0, xc1, d0, d1, cl0, cl1, k0, k1, j1;\n\t" \
                    ".reg .s16 xs0, xs1;\n\t" \
                    ".reg .b64 ad0, ad1, kk0, kk1;\n\t" \
                    "ld.shared.u32 m0, [%2 + 8192];\n\t" \
                    "ld.shared.u32 m1, [%2 + 8320];\n\t" \
                    "ld.shared.u32 l0, [%2];\n\t" \
                    "ld.shared.u32 l1, [%2 + 128];\n\t" \
                    "add.u32 j1, %3, 32;\n\t" \
                    "shr.s32 y0, l0, 16;\n\t" \
                    "shr.s32 y1, l1, 16;\n\t" \
                    "and.b32 c0, m0, 0xffff;\n\t" \
                    "and.b32 c1, m1, 0xffff;\n\t" \
                    "sub.u32 r0, c0, %5;\n\t" \
                    "sub.u32 r1, c1, %5;\n\t" \
                    "setp.lt.u32 pa0, y0, %7;\n\t" \
                    "setp.lt.u32 pa1, y1, %7;\n\t" \
                    "setp.lt.and.u32 pa0, %3, %4, pa0;\n\t" \
                    "setp.lt.and.u32 pa1, j1, %4, pa1;\n\t" \
                    "setp.lt.and.u32 ph0, r0, %6, pa0;\n\t" \
                    "setp.lt.and.u32 ph1, r1, %6, pa1;\n\t" \
                    "setp.ge.and.u32 pm0, r0, %6, pa0;\n\t" \
                    "setp.ge.and.u32 pm1, r1, %6, pa1;\n\t" \
                    "mad.lo.u32 a0, r0, %8, y0;\n\t" \
                    "mad.lo.u32 a1, r1, %8, y1;\n\t" \
                    "shl.b32 a0, a0, 1;\n\t" \
                    "shl.b32 a1, a1, 1;\n\t" \
                    "add.u32 a0, a0, %9;\n\t" \
                    "add.u32 a1, a1, %9;\n\t" \
                    "mov.b16 xs0, 0;\n\t" \
                    "mov.b16 xs1, 0;\n\t" \
                    "@ph0 ld.shared.s16 xs0, [a0];\n\t" \
                    "@ph1 ld.shared.s16 xs1, [a1];\n\t" \
                    "cvt.s32.s16 x0, xs0;\n\t" \
                    "cvt.s32.s16 x1, xs1;\n\t" \
                    "sub.s32 x0, x0, %10;\n\t" \
                    "sub.s32 x1, x1, %10;\n\t" \
                    "cvt.s32.s16 xc0, l0;\n\t" \
                    "cvt.s32.s16 xc1, l1;\n\t" \
                    "sub.s32 d0, x0, xc0;\n\t" \
                    "sub.s32 d1, x1, xc1;\n\t" \
                    "and.b32 k0, d0, 0x8000;\n\t" \
                    "and.b32 k1, d1, 0x8000;\n\t" \
                    "setp.eq.and.u32 pi0, k0, 0, ph0;\n\t" \
                    "setp.eq.and.u32 pi1, k1, 0, ph1;\n\t" \
                    "mad.lo.s32 cl0, y0, %11, x0;\n\t" \
                    "mad.lo.s32 cl1, y1, %11, x1;\n\t" \
                    "mad.wide.s32 ad0, cl0, 8, %12;\n\t" \
                    "mad.wide.s32 ad1, cl1, 8, %12;\n\t" \
                    "prmt.b32 k0, d0, m0, 0x7610;\n\t" \
                    "prmt.b32 k1, d1, m1, 0x7610;\n\t" \
                    "add.u32 k0, k0, %13;\n\t" \
                    "add.u32 k1, k1, %13;\n\t" \
                    "mov.b64 kk0, {k0, %14};\n\t" \
                    "mov.b64 kk1, {k1, %14};\n\t"
#define XM_BACK_PTX_RED_PLAIN \
                    "@pi0 red.global.max.u64 [ad0], kk0;\n\t" \
                    "@pi1 red.global.max.u64 [ad1], kk1;\n\t"
#define XM_BACK_PTX_RED_AGG \
                    "{\n\t" \
                    ".reg .b32 g0, g1, q0, q1;\n\t" \
                    ".reg .pred pl0, pl1;\n\t" \
                    "sub.s32 q0, -1, %15;\n\t" \
                    "selp.b32 g0, cl0, q0, pi0;\n\t" \
                    "selp.b32 g1, cl1, q0, pi1;\n\t" \
                    "match.any.sync.b32 g0, g0, 0xffffffff;\n\t" \
                    "match.any.sync.b32 g1, g1, 0xffffffff;\n\t" \
                    "shr.u32 g0, g0, %15;\n\t" \
                    "shr.u32 g1, g1, %15;\n\t" \
                    "setp.eq.and.u32 pl0, g0, 1, pi0;\n\t" \
                    "setp.eq.and.u32 pl1, g1, 1, pi1;\n\t" \
                    "@pl0 red.global.max.u64 [ad0], kk0;\n\t" \
                    "@pl1 red.global.max.u64 [ad1], kk1;\n\t" \
                    "}\n\t"
#define XM_BACK_PTX_TAIL \
                    "selp.u32 %0, 1, 0, pi0;\n\t" \
                    "@pi1 add.u32 %0, %0, 1;\n\t" \
                    "selp.u32 %1, 1, 0, pm0;\n\t" \
                    "@pm1 or.b32 %1, %1, 2;\n\t" \
                    "}"
#define XM_BACK_PTX_OPERANDS \
                    : "=r"(n_in), "=r"(n_miss) \
                    : "r"(a_e), "r"(j0), "r"(count), "r"(win_lo), "r"(win_n), "r"(static_cast<unsigned>(bp.xmap_h) - 1u), "r"(static_cast<unsigned>(bp.col_stride)), \
                      "r"(a_win_c), "r"(bp.x_offset), "r"(bp.rect_w), "l"(map), "r"(idx_lo), "r"(key_hi), "r"(lane) \
                    : "memory"

// What a chunk carries from its front half to its back half (one register):
//   bit 0 constant slot | bits 1-3 events of this thread that pass the polarity mask | bit 4 a pixel outside the
//   camera image | bit 5 a timestamp outside the assumed bounds | bits 8.. entries of the warp's live list
constexpr unsigned kCarryOob = 1u << 4, kCarryTb = 1u << 5;
constexpr unsigned kAggregateAbove = XM_AGG_ABOVE;  // live-list entries (of 128) above which a chunk's scatter is warp-aggregated
constexpr unsigned kMetaSkip = 0x8000u;  // list entry of an event whose column is not usable (bounds violation)
constexpr unsigned kMetaOff = 2 * kListBytes, kPixOff = 4 * kListBytes;  // (column | index << 16) / camera pixel lists behind the LUT words

// Can an event of this record's pixel block (first word x | y << 16) be an inlier at time column q?  (alive table,
// see BatchParams; clamped index: any coordinates are safe)
__device__ __forceinline__ bool batch_alive(const BatchParams& bp, unsigned a_alive, unsigned xy, unsigned q) {
    const unsigned idx = min((xy >> 16 >> bp.alive_shift) * static_cast<unsigned>(bp.alive_pitch) + ((xy & 0xffffu) >> bp.alive_shift), bp.alive_last);
    const unsigned e = lds_u16_a(a_alive + idx * 2u);
    return (q >> bp.alive_qs) - (e & 0xffu) <= (e >> 8);
}

// GENERAL front half of a chunk (whole warp, out of line): partial chunks, frames the integer time column does not
// cover, and chunks in which the fast pass found an exact rounding tie or a timestamp outside the assumed bounds.
// Evaluates the reference's own float64 expression for every event and (re)writes the warp's live list.  The live
// flags depend on the time column, so the list a fast pass left behind may be laid out differently: the caller waits
// for the gathers that pass started (cp.async.wait_all) before this one reuses the slots.  Returns the chunk's carry
// (without the constant slot).
template <bool CAM>
static __device__ __noinline__ unsigned batch_front_general(const BatchParams& bp, int f, unsigned a_stage, unsigned a_list_c, int limit,
                                                            unsigned a_alive, int tid) {
    TimeCol<false> tc;
    tc.init(__ldcg(&bp.states[f].t_lo_bits), __ldcg(&bp.states[f].t_hi_bits), bp.t_px_scale);
    const unsigned pol_mask = bp.polarity ? 0xffffu : 0u;
    const unsigned lt_mask = (1u << (tid & 31)) - 1u;
    bool tb = false, oob = false;
    unsigned count = 0, kept = 0;
#pragma unroll 1
    for (int k = 0; k < kEvPerThread; ++k) {
        const int4 rec = lds128_a(a_stage + k * (kEvThreads * 16));  // the stage is ours until release()
        const unsigned ex = static_cast<unsigned>(rec.x) & 0xffffu, ey = static_cast<unsigned>(rec.x) >> 16;
        const bool valid = ((static_cast<unsigned>(rec.y) ^ 1u) & pol_mask) == 0u && (k * kEvThreads + tid < limit);
        const bool ok = valid && ex < static_cast<unsigned>(bp.cam_w) && ey < static_cast<unsigned>(bp.cam_h);
        bool viol = false;
        int cc = 0;
        if (ok) {  // (not kept, or not inside the image: no time column)
            const long long t_bits = (static_cast<long long>(rec.w) << 32) | static_cast<unsigned>(rec.z);
            cc = tc.column(t_bits, viol);    // exact for every event (== the integer form wherever that is valid)
            if (cc < 0) cc += bp.xmap_w;     // NumPy negative index (only reachable with wrong bounds)
            viol = viol || cc < 0 || cc >= bp.xmap_w;
        }
        // (an event whose column is not usable flags the frame below and needs no list entry)
        const bool live = ok && !viol && batch_alive(bp, a_alive, static_cast<unsigned>(rec.x), static_cast<unsigned>(cc));
        const unsigned mask = __ballot_sync(0xffffffffu, live);
        const unsigned pos = count + __popc(mask & lt_mask);
        count += __popc(mask);
        kept += valid ? 1u : 0u;
        oob = oob || (valid && !ok);  // the reference raises IndexError here
        tb = tb || viol;  // dead events too: a timestamp outside the assumed bounds invalidates the frame's normalisation
        if (live) {
            const int px = static_cast<int>(ey * static_cast<unsigned>(bp.cam_w) + ex);
            cp_async_4_a(a_list_c + pos * 4u, bp.lut_xy + px);
            sts32_a(a_list_c + kMetaOff + pos * 4u, static_cast<unsigned>(cc) | (static_cast<unsigned>(k * kEvThreads + tid) << 16));
            if (CAM) sts32_a(a_list_c + kPixOff + pos * 4u, static_cast<unsigned>(px));
        }
    }
    return (kept << 1) | (oob ? kCarryOob : 0u) | (tb ? kCarryTb : 0u) | (count << 8);
}

// AGG: the instantiation whose dense chunks aggregate their scatter per warp round (see XM_BACK_PTX_*).  A second
// instantiation rather than a run-time switch: the mere presence of the second PTX block costs the plain path 1.6 %.
// TW: epilogue warps per CTA.  The strip epilogue of large frames needs only two (it waits most of the time), which
// leaves the per-event loop 80 registers instead of 72; small frames, whose pace the epilogue sets, take four.
#ifndef XM_TILE_WARPS_LARGE
#define XM_TILE_WARPS_LARGE 2
#endif
constexpr int kTileWarpsLarge = XM_TILE_WARPS_LARGE;
template <bool CAM, bool AGG = false, int TW = kTileWarps>
__global__ void __launch_bounds__(kWsThreads + TW * 32, kBatchCtasPerSm) batch_kernel(const __grid_constant__ BatchParams bp) {
    constexpr int kGroups = TW * 32 / kTileGroupThreads;
    constexpr int kThreads = kWsThreads + TW * 32;
    extern __shared__ __align__(128) unsigned char ev_smem[];
    uint64_t* full_ev = reinterpret_cast<uint64_t*>(ev_smem);
    uint64_t* empty_ev = reinterpret_cast<uint64_t*>(ev_smem + 32);
    uint64_t* full_win = reinterpret_cast<uint64_t*>(ev_smem + 64);
    uint64_t* empty_win = reinterpret_cast<uint64_t*>(ev_smem + 96);
    int2* win_meta = reinterpret_cast<int2*>(ev_smem + 128);  // [4] (first column, count) of a window stage
    int2* ev_meta = reinterpret_cast<int2*>(ev_smem + 160);   // [4] (kind << 16 | frame, chunk or tile index); x < 0: end
    unsigned* s_acc = reinterpret_cast<unsigned*>(ev_smem + 192);  // [8][4] per-frame CTA accumulators: valid, inliers, flags, warps
    const int lut_tile = batch_list_bytes(CAM);  // the live lists
    unsigned char* ring = ev_smem + kBatchHeader + lut_tile;
    const int win_bytes = bp.cap_cols * bp.col_stride * 2;
    unsigned char* win_ring = ring + bp.stages * (kEvChunk * 16);

    int tid;  // (opaque: the compiler otherwise re-reads the special register inside the hot loop to save a register)
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const bool producer = warp == kEvThreads / 32;
    const int B = bp.n_frames;
    unsigned* const item_counter = &bp.states[B].next_chunk;

    if (tid == 0) {
        for (int s = 0; s < kWsMaxStages; ++s) {
            mbar_init(full_ev + s, 1);
            mbar_init(empty_ev + s, kEvThreads / 32);
            mbar_init(full_win + s, 1);
            mbar_init(empty_win + s, kEvThreads / 32);
        }
    }
    if (tid < 32) s_acc[tid] = 0u;
    unsigned* const s_alive = reinterpret_cast<unsigned*>(win_ring + bp.win_stages * win_bytes + kGroups * bp.ep.region_cap * 4);
    for (int i = tid; i < bp.alive_words; i += kThreads) s_alive[i] = __ldg(bp.alive + i);
    __syncthreads();

    if (warp > kEvThreads / 32) {
        // ---- epilogue groups ----------------------------------------------------------------------
        if (!CAM && bp.strips) {
            batch_strip_warps<kEvChunk>(bp, lane);
            return;
        }
        const int grp = (tid - kWsThreads) / kTileGroupThreads, gtid = (tid - kWsThreads) % kTileGroupThreads;
        unsigned short* bufA = reinterpret_cast<unsigned short*>(win_ring + bp.win_stages * win_bytes) + grp * (2 * bp.ep.region_cap);
        unsigned short* bufB = bufA + bp.ep.region_cap;
        volatile int* s_ticket = reinterpret_cast<volatile int*>(ev_smem + 1088) + grp;
        batch_tile_groups<CAM, kEvChunk>(bp, grp, gtid, bufA, bufB, s_ticket);
        return;
    }

    if (producer) {
        if (lane != 0) return;
        // ---- producer lane ------------------------------------------------------------------------
        int se = 0, sw = 0;
        unsigned pe = 0, pw = 0;
        int issued = 0;
        const uint64_t pol = make_evict_first_policy();
        // chunk -> (frame, chunk index inside the frame); x = -1: past the end (chunks arrive in increasing order)
        int slot = 0;
        unsigned s_first = bp.first_item[0], s_next = bp.first_item[1];
        auto decode = [&](unsigned it) -> int2 {
            if (it >= bp.total_items) return make_int2(-1, 0);
            while (it >= s_next) {
                ++slot;
                s_first = s_next;
                s_next = bp.first_item[slot + 1];
            }
            return make_int2(slot, static_cast<int>(it - s_first));
        };
        // first / last timestamp of a chunk (loads only; consumed one iteration later)
        auto fetch_ts = [&](const int2& d, long long& ta, long long& tb) {
            ta = tb = 0;
            if (d.x < 0) return;
            const BatchFrame& fr = bp.frames[d.x];
            const long long first = static_cast<long long>(d.y) * kEvChunk;
            const long long last = min(fr.n, first + kEvChunk) - 1;
            ta = ld_time_nc(fr.events + first);
            tb = ld_time_nc(fr.events + last);
        };

        int tc_frame = -1;
        TimeCol<false> tc;
        tc.init(0, 0, bp.t_px_scale);

        int2 dA = decode(atomicAdd(item_counter, 1u));
        long long tAa, tAb;
        fetch_ts(dA, tAa, tAb);
        unsigned itB = dA.x >= 0 ? atomicAdd(item_counter, 1u) : 0xffffffffu;
        for (;;) {
            // look-ahead: decode the next item (its atomic was issued one iteration ago), start its
            // timestamp loads, and issue the atomic of the one after
            const int2 dB = dA.x >= 0 ? decode(itB) : make_int2(-1, 0);
            long long tBa, tBb;
            fetch_ts(dB, tBa, tBb);
            const unsigned itC = dB.x >= 0 ? atomicAdd(item_counter, 1u) : 0xffffffffu;

            if (issued >= bp.stages) mbar_wait_relaxed(empty_ev + se, pe);
            ev_meta[se] = dA;
            if (dA.x < 0) {
                mbar_arrive(full_ev + se);  // descriptor only (end of the batch)
            } else {
                const BatchFrame& fr = bp.frames[dA.x];
                const long long first = static_cast<long long>(dA.y) * kEvChunk;
                const unsigned count = static_cast<unsigned>(min(static_cast<long long>(kEvChunk), fr.n - first));
                mbar_expect_tx(full_ev + se, count * 16u);
                tma_load_1d_hint(ring + se * (kEvChunk * 16), fr.events + first, count * 16u, full_ev + se, pol);
            }
            ++issued;
            if (++se == bp.stages) {
                se = 0;
                if (issued > bp.stages) pe ^= 1u;
            }
            if (dA.x < 0) break;
            if (bp.cap_cols > 0) {
                // X-map window: the columns between the chunk's first and last record
                if (dA.x != tc_frame) {
                    tc_frame = dA.x;
                    const FrameState* st = bp.states + tc_frame;
                    tc.init(__ldcg(&st->t_lo_bits), __ldcg(&st->t_hi_bits), bp.t_px_scale);
                }
                bool va, vb;
                int ca = tc.column(tAa, va), cb = tc.column(tAb, vb);
                ca = min(max(ca, 0), bp.xmap_w - 1);
                cb = min(max(cb, 0), bp.xmap_w - 1);
                const int lo = min(ca, cb);
                const int n = min(min(max(ca, cb) - lo + 1, bp.cap_cols), bp.xmap_w - lo);
                mbar_wait_relaxed(empty_win + sw, pw ^ 1u);  // first round passes immediately
                win_meta[sw] = make_int2(lo, n);
                const unsigned bytes = static_cast<unsigned>(n) * bp.col_stride * 2u;
                mbar_expect_tx(full_win + sw, bytes);
                tma_load_1d(win_ring + sw * win_bytes, bp.xmap_t + static_cast<long long>(lo) * bp.col_stride, bytes, full_win + sw);
                if (++sw == bp.win_stages) {
                    sw = 0;
                    pw ^= 1u;
                }
            }
            dA = dB;
            tAa = tBa;
            tAb = tBb;
            itB = itC;
        }
        return;
    }

    // ---- consumers (8 warps) ------------------------------------------------------------------------
    const unsigned sbase = smem_u32(ev_smem);
    const unsigned a_full_ev = sbase, a_empty_ev = sbase + 32, a_full_win = sbase + 64, a_empty_win = sbase + 96;
    const unsigned a_wmeta = sbase + 128, a_emeta = sbase + 160;
    // live lists of this warp: kEvPerThread * 32 entries per list and buffer
    //   LUT words at a_list + buffer * kListBytes, (column | index << 16) at + 2 * kListBytes, camera pixel at + 4 * kListBytes
    const unsigned a_list = sbase + kBatchHeader + warp * (kEvPerThread * 32 * 4);
    const unsigned a_ring = sbase + kBatchHeader + lut_tile + tid * 16;      // + slot * 16384 + k * 4096
    const unsigned a_win = sbase + kBatchHeader + lut_tile + bp.stages * (kEvChunk * 16);
    const unsigned a_alive = smem_u32(s_alive);
    const unsigned lt_mask = (1u << lane) - 1u;
    // geometry is read from the parameter bank where it is used (constant operands, no registers)
#define XM_B_YLIM (static_cast<unsigned>(bp.xmap_h) - 1u)
#define XM_B_CAMW (static_cast<unsigned>(bp.cam_w))
#define XM_B_CAMH (static_cast<unsigned>(bp.cam_h))
    const unsigned pol_mask = bp.polarity ? 0xffffu : 0u;

    int fe = 0, bw = 0;
    unsigned fpe = 0, bpw = 0;
    unsigned fpar = 0, bpar = 0;  // list buffers of the next front / back half

    // per-frame state: all statistics are taken in the BACK half, so they belong to exactly one frame
    int cur_f = -1;    // frame of the back half (the one whose statistics are being accumulated)
    int front_f = -1;  // frame whose constants the front half prepared last
    unsigned fslot = 0;  // constant slot of front_f (toggles with every new frame; a chunk carries its slot to the back half)
    unsigned my_chunks = 0, n_valid = 0, n_inl = 0, flags = 0;
    // The constants of the frames a warp is working on live in shared memory (two slots of 48 B per warp,
    // alternating per new frame, written by lane 0 when the warp's FRONT half enters the frame, read back with
    // broadcast loads where they are used; the BACK half is at most one chunk -- hence one frame -- behind): unlike
    // the single-frame kernels, where they are kernel parameters, they would otherwise occupy ~14
    // registers across the whole per-event loop and push it into local-memory spills.
    //   [0] t_min lo, hi, range, 2*scale   [1] d, M, shift, ok   [2] map lo, hi, epoch << 16, n_events
    const unsigned a_fc = sbase + 320 + warp * 96;  // + slot * 48

    // leaves frame cur_f: statistics and the count of finished chunks go to the frame's state block
    // (once per CTA: the last of the eight warps to leave forwards the CTA's totals)
    auto leave_frame = [&]() {
        if (cur_f < 0) return;
        n_valid = __reduce_add_sync(0xffffffffu, n_valid);
        n_inl = __reduce_add_sync(0xffffffffu, n_inl);
        flags = __reduce_or_sync(0xffffffffu, flags);
        fence_acq_rel_gpu();  // every lane: its scatter atomics are ordered before the counts below
        __syncwarp();
        if (lane == 0) {
            unsigned* acc = s_acc + (cur_f & 7) * 4;
            if (n_valid) atomicAdd(acc + 0, n_valid);
            if (n_inl) atomicAdd(acc + 1, n_inl);
            if (flags) atomicOr(acc + 2, flags);
            __threadfence_block();
            if (atomicAdd(acc + 3, 1u) == kEvThreads / 32 - 1) {
                __threadfence_block();
                const unsigned v = atomicExch(acc + 0, 0u), i = atomicExch(acc + 1, 0u), fl = atomicExch(acc + 2, 0u);
                atomicExch(acc + 3, 0u);
                FrameState* st = bp.states + cur_f;
                if (v) atomicAdd(&st->n_valid, static_cast<unsigned long long>(v));
                if (i) atomicAdd(&st->n_inliers, static_cast<unsigned long long>(i));
                if (fl) atomicOr(&st->flags, fl);
                fence_acq_rel_gpu();
                atomicAdd(&st->blocks_done, my_chunks);
#ifdef XM_DEBUG_HOOKS
                if (bp.dbg) atomicMax(bp.dbg + cur_f * 4 + 1, global_timer_ns());
#endif
            }
        }
        __syncwarp();
        n_valid = n_inl = flags = 0;
        my_chunks = 0;
        cur_f = -1;
    };
    // FRONT side: constants of frame f into the warp's other slot
    auto prepare_frame = [&](int f) {
        front_f = f;
        fslot ^= 1u;
        const FrameState* st = bp.states + f;
        if (lane == 0) {
            // the frame's time-column constants as batch_bounds_kernel left them (one division per frame, not per warp);
            // requested BEFORE the wait below, which they do not depend on: one L2 round trip per frame instead of two
            const int4 k0 = __ldcg(&st->ic0), k1 = __ldcg(&st->ic1);
            if (f >= bp.n_maps) {
                // this frame scatters into the map frame f - n_maps used: all of that frame's tiles must have read it
                spin_until_ge(&bp.states[f - bp.n_maps].next_tile, static_cast<unsigned>(bp.tile_items), 64u, 64u);
            }
#ifdef XM_DEBUG_HOOKS
            if (bp.dbg) atomicMin(bp.dbg + f * 4 + 0, global_timer_ns());
#endif
            const unsigned a = a_fc + fslot * 48;
            const unsigned long long mp = reinterpret_cast<unsigned long long>(bp.maps[f % bp.n_maps]);
            sts128_a(a, k0);
            sts128_a(a + 16, k1);
            sts128_a(a + 32, make_int4(static_cast<int>(mp), static_cast<int>(mp >> 32),
                                       static_cast<int>((bp.epoch0 + static_cast<unsigned>(f)) << 16), static_cast<int>(bp.frames[f].n)));
        }
        __syncwarp();  // (also orders every lane's scatter after lane 0's acquire)
    };

    auto peek = [&]() -> int2 {
        mbar_wait_a(a_full_ev + fe * 8, fpe);
        return lds64_a(a_emeta + fe * 8);
    };
    auto release = [&]() {
        __syncwarp();
        if (elect_one()) mbar_arrive_a(a_empty_ev + fe * 8);
        if (++fe == bp.stages) {
            fe = 0;
            fpe ^= 1u;
        }
    };

    auto alive = [&](unsigned xy, unsigned q) -> bool { return batch_alive(bp, a_alive, xy, q); };

    // FRONT half of chunk g of frame f (stage fe): polarity / image / alive tests, bounds check and time column of
    // every event; live events are compacted into the warp's lists and their LUT gathers started; releases the stage.
    auto front = [&](int f, int g) -> unsigned {
        if (f != front_f) prepare_frame(f);
        const unsigned a_fcf = a_fc + fslot * 48;
        const unsigned a_stage = a_ring + fe * (kEvChunk * 16);
        const unsigned a_list_c = a_list + fpar * kListBytes;
        fpar ^= 1u;
        const int4 c1 = lds128_a(a_fcf + 16);
        const unsigned left = static_cast<unsigned>(lds32_a(a_fcf + 44)) - static_cast<unsigned>(g) * kEvChunk;
        unsigned carry;
        if (left >= kEvChunk && c1.w != 0) {
            // FAST pass: a full chunk of a frame whose time columns have the integer form
            IntCol ic;
            {
                const int4 c0 = lds128_a(a_fcf);
                ic.lo = (static_cast<unsigned long long>(static_cast<unsigned>(c0.y)) << 32) | static_cast<unsigned>(c0.x);
                ic.range = static_cast<unsigned>(c0.z);
                ic.scale2 = static_cast<unsigned>(c0.w);
                ic.d = static_cast<unsigned>(c1.x);
                ic.M = static_cast<unsigned>(c1.y);
                ic.sh = c1.z;
                ic.ok = true;
            }
            bool any_bad = false, oob = false;
            unsigned kept = 0;   // events of this thread that pass the polarity mask (n_valid)
            unsigned count = 0;  // entries of the warp's live list so far (warp-uniform)
            // XM_FRONT_GROUP events at a time (registers for the raw records vs. independent work in flight)
#ifndef XM_FRONT_GROUP
#define XM_FRONT_GROUP 2
#endif
#pragma unroll
            for (int h = 0; h < kEvPerThread; h += XM_FRONT_GROUP) {
                int4 raw[XM_FRONT_GROUP];
#pragma unroll
                for (int j = 0; j < XM_FRONT_GROUP; ++j) raw[j] = lds128_a(a_stage + (h + j) * (kEvThreads * 16));
#pragma unroll
                for (int j = 0; j < XM_FRONT_GROUP; ++j) {
                    const int k = h + j;
                    const unsigned xy = static_cast<unsigned>(raw[j].x);
                    const unsigned ex = xy & 0xffffu, ey = xy >> 16;
                    const bool valid = ((static_cast<unsigned>(raw[j].y) ^ 1u) & pol_mask) == 0u;  // polarity: p == 1, or everything
                    const bool ok = valid && ex < XM_B_CAMW && ey < XM_B_CAMH;
                    // can a pixel of this block be an inlier at this time column?  (shared-memory table; looked up
                    // unconditionally: a branch per event costs more than the load, and the clamped index is safe for
                    // any coordinates; a wrong q of a `bad` event only matters until the general pass redoes the chunk)
                    const long long t_bits = (static_cast<long long>(raw[j].w) << 32) | static_cast<unsigned>(raw[j].z);
                    bool bad;
                    const unsigned q = ic.column(t_bits, bad);
                    const bool live = ok & alive(xy, q);
                    // dead events too: a timestamp outside the assumed bounds invalidates the frame's normalisation
                    any_bad = any_bad || (ok && bad);
                    oob = oob || (valid && !ok);  // the reference raises IndexError here
                    kept += valid ? 1u : 0u;
                    const unsigned mask = __ballot_sync(0xffffffffu, live);
                    const unsigned pos = count + __popc(mask & lt_mask);
                    count += __popc(mask);
                    if (live) {
                        const int px = static_cast<int>(ey * XM_B_CAMW + ex);
                        // (gather first: ptxas pads a shared-memory store that is followed by an LDGSTS with three dummy loads)
#ifdef XM_DEBUG_HOOKS  // timing experiments only (results WRONG): 2 = no LUT gathers, 64 = gathers hit one L1-resident 16 KB slice
                        if (!(bp.debug & 2)) cp_async_4_a(a_list_c + pos * 4u, bp.lut_xy + ((bp.debug & 64) ? (px & 0xfff) : px));
#else
                        cp_async_4_a(a_list_c + pos * 4u, bp.lut_xy + px);
#endif
                        sts32_a(a_list_c + kMetaOff + pos * 4u, q | (static_cast<unsigned>(k * kEvThreads + tid) << 16));
                        if (CAM) sts32_a(a_list_c + kPixOff + pos * 4u, static_cast<unsigned>(px));
                    }
                }
            }
            carry = (kept << 1) | (oob ? kCarryOob : 0u) | (count << 8);
            if (__any_sync(0xffffffffu, any_bad)) {
                cp_async_wait_all();  // the list is laid out anew: no gather of this pass may land in it afterwards
                __syncwarp();
                carry = batch_front_general<CAM>(bp, f, a_stage, a_list_c, kEvChunk, a_alive, tid);
            }
        } else {
            carry = batch_front_general<CAM>(bp, f, a_stage, a_list_c, left < kEvChunk ? static_cast<int>(left) : kEvChunk, a_alive, tid);
        }
        cp_async_commit();
        release();
        return carry | fslot;
    };

    // BACK half: the warp's live list -- X-map lookups (window in shared memory, else through L2), disparity, scatter
    auto back = [&](int f, unsigned carry, int g) {
        const unsigned slot = carry & 1u;
        const unsigned count = carry >> 8;
        if (f != cur_f) {
            // first chunk of a new frame: the previous frame's scatter is complete for this warp.  Its fence
            // and counters run while the gathers of the next chunk (issued by the front half) are in flight.
            leave_frame();
            cur_f = f;
        }
        n_valid += (carry >> 1) & 7u;
        if (carry & (kCarryOob | kCarryTb)) {
            if (carry & kCarryOob) flags |= kStatusPixelOob;
            if (carry & kCarryTb) flags |= kStatusTBounds;
        }
        const unsigned a_fcb = a_fc + slot * 48;
        unsigned win_lo = 0, win_n = 0;
        unsigned a_win_c = a_list;  // any valid address: without a window every lookup misses
        if (bp.cap_cols > 0) {
            mbar_wait_a(a_full_win + bw * 8, bpw);
            const int2 wm = lds64_a(a_wmeta + bw * 8);
            win_lo = static_cast<unsigned>(wm.x);
            win_n = static_cast<unsigned>(wm.y);
            a_win_c = a_win + bw * win_bytes;
        }
        const unsigned a_list_c = a_list + bpar * kListBytes + lane * 4u;
        bpar ^= 1u;
        const int4 c2 = lds128_a(a_fcb + 32);
        unsigned long long* const map = reinterpret_cast<unsigned long long*>(
            (static_cast<unsigned long long>(static_cast<unsigned>(c2.y)) << 32) | static_cast<unsigned>(c2.x));
        // event index = g * kEvChunk + index inside the chunk: its bits 16.. are the same for the whole chunk
        static_assert(kEvChunk == 1024, "key halves below assume 1024-event chunks");
        const unsigned key_hi = static_cast<unsigned>(c2.z) | (static_cast<unsigned>(g) >> 6);
        const unsigned idx_lo = (static_cast<unsigned>(g) & 63u) << 26;  // bits 16.. of the key's low word
        unsigned missed = 0;  // entries whose column is not in the staged window (or that carry the skip flag)
        // Dense chunk (nearly every event is live: what a scanning projector produces, consecutive events from the few
        // pixels it lights) -> aggregate the scatter per warp round; a uniform stream (~40 % live) keeps the plain form.
        const bool aggregate = AGG && count > kAggregateAbove;
#pragma unroll 1
        for (unsigned jb = 0; jb < count; jb += 64u) {  // (warp-uniform trip count)
            const unsigned j0 = jb + lane;
            // Two list entries per lane and round, everything for entries whose column is in the window, in one block of
            // PTX: predicates stay predicates (no masks in registers), the scatter is a predicated RED.
            const unsigned a_e = a_list_c + jb * 4u;
            unsigned n_in, n_miss;
            if (!CAM) {
                if (AGG && aggregate)
                    asm volatile(XM_BACK_PTX_HEAD XM_BACK_PTX_RED_AGG XM_BACK_PTX_TAIL XM_BACK_PTX_OPERANDS);
                else
                    asm volatile(XM_BACK_PTX_HEAD XM_BACK_PTX_RED_PLAIN XM_BACK_PTX_TAIL XM_BACK_PTX_OPERANDS);
                n_inl += n_in;
                missed |= n_miss;
            } else {
                missed |= 3u;  // camera view: the general loop below does everything
            }
        }
        if (__any_sync(0xffffffffu, missed != 0u)) {
            // entries the block above left out: column outside the staged window (read the transposed table through
            // L2), skip flag, camera view
#pragma unroll 1
            for (unsigned j = lane; j < count; j += 32u) {
                const unsigned meta = static_cast<unsigned>(lds32_a(a_list_c + kMetaOff + (j - lane) * 4u));
                const int lut = lds32_a(a_list_c + (j - lane) * 4u);
                const int ycr = lut >> 16, xcr = static_cast<short>(lut & 0xffff);
                const unsigned col = meta & 0xffffu;
                if (static_cast<unsigned>(ycr) >= XM_B_YLIM || (meta & kMetaSkip)) continue;
                const unsigned rel = col - win_lo;
                int xp;
                if (rel < win_n) {
                    if (!CAM) continue;  // done above
                    xp = lds_s16_a(a_win_c + (rel * static_cast<unsigned>(bp.col_stride) + static_cast<unsigned>(ycr)) * 2u);
                } else {
                    xp = __ldg(bp.xmap_t + static_cast<long long>(col) * bp.col_stride + ycr);
                }
                const int disp = static_cast<short>(xp - xcr - bp.x_offset);  // int16 arithmetic wraps
                if (disp < 0) continue;
                // projector view: x_rect + disp = x_map - x_offset, in [0, rect_w) for verified tables
                const int cell = CAM ? lds32_a(a_list_c + kPixOff + (j - lane) * 4u) : ycr * bp.rect_w + (xp - bp.x_offset);
                const unsigned key_lo = (idx_lo + (meta & 0xffff0000u)) | (static_cast<unsigned>(disp) & 0xffffu);
                const unsigned long long key = (static_cast<unsigned long long>(key_hi) << 32) | key_lo;
#ifdef XM_DEBUG_HOOKS  // 1 = no scatter atomics
                red_max_u64_if(map + cell, key, !(bp.debug & 1));
#else
                red_max_u64_if(map + cell, key, true);
#endif
                ++n_inl;
            }
        }
        __syncwarp();  // the list buffer is rewritten by the front half after next
        if (bp.cap_cols > 0) {
            if (elect_one()) mbar_arrive_a(a_empty_win + bw * 8);
            if (++bw == bp.win_stages) {
                bw = 0;
                bpw ^= 1u;
            }
        }
        ++my_chunks;
    };

#ifdef XM_DEBUG_HOOKS  // 32 = consumers only take and release the stages (pure TMA stream through this pipeline)
    if (bp.debug & 32) {
        for (;;) {
            const int2 mm = peek();
            if (mm.x < 0) break;
            if (mm.x != cur_f) {  // the tile groups still wait for every frame's chunk count
                leave_frame();
                cur_f = mm.x;
            }
            ++my_chunks;
            const int4 r = lds128_a(a_ring + fe * (kEvChunk * 16));
            if (r.x == 0x7fffffff && r.y == 0x7fffffff) flags |= 1u << 30;  // keep the load
            release();
            if (bp.cap_cols > 0) {
                mbar_wait_a(a_full_win + bw * 8, bpw);
                __syncwarp();
                if (elect_one()) mbar_arrive_a(a_empty_win + bw * 8);
                if (++bw == bp.win_stages) {
                    bw = 0;
                    bpw ^= 1u;
                }
            }
        }
        leave_frame();
        return;
    }
#endif
    // The software pipeline runs across frame boundaries: FRONT half of chunk c+1 (possibly the first chunk
    // of the next frame), then BACK half of chunk c.  (One call site each: the two halves are long, and the loop
    // body has to stay inside the instruction cache.)
    bool have_cur = false;
    int f_cur = -1, g_cur = 0;
    unsigned s_cur = 0;
    for (;;) {
        const int2 m = peek();  // (does not consume the stage: front() releases it)
        const bool more = m.x >= 0;
        // Entering frame F waits for the tiles of frame F - n_maps, which wait for every warp's count
        // of that frame's chunks.  This warp publishes a frame's count only in the BACK half of a later
        // frame's chunk, so if one of its two unpublished frames (cur_f: back half, f_cur: front half done)
        // is that old, the pipeline is drained first (tiny frames / many more CTAs than chunks per frame).
        bool drain = false;
        if (more && have_cur && m.x != f_cur) {
            drain = bp.hard_frames != 0;
            if (!drain && m.x != front_f && m.x >= bp.n_maps) {
                const int must_be_out = m.x - bp.n_maps;
                drain = f_cur <= must_be_out || (cur_f >= 0 && cur_f <= must_be_out);
            }
        }
        const bool do_front = more && !drain;
        unsigned s_nxt = 0;
        if (do_front) s_nxt = front(m.x, m.y);
        if (have_cur) {
            if (do_front)
                cp_async_wait<1>();  // the gathers of the current chunk have landed; the next chunk's stay in flight
            else
                cp_async_wait<0>();
            __syncwarp();  // ... those of every lane: a list entry is read by another lane than the one that gathered it
            back(f_cur, s_cur, g_cur);
            have_cur = false;
        }
        if (do_front) {
            f_cur = m.x;
            g_cur = m.y;
            s_cur = s_nxt;
            have_cur = true;
        } else if (more) {
            leave_frame();  // drained: publish, then take the same chunk again
        } else {
            break;
        }
    }
    leave_frame();
}

// ---------------------------------------------------------------------------------------------------
// Exact re-render of the frames whose optimistic bounds were violated (unsorted input): one thread
// walks the batch and tail-launches, per flagged frame, the bounds reduction, the lean per-event
// kernel on the reduced bounds and the epilogue -- the same kernels the single-frame path uses.
// ---------------------------------------------------------------------------------------------------
struct BatchRedoParams {
    EventParams ev;       // shared fields; events / n / epoch / state / map are set per frame
    EpilogueParams ep;    // shared fields; dst / epoch / state are set per frame
    unsigned epoch_redo0;  // frame f re-renders with epoch_redo0 + f
    int n_frames;
    int sm_count;
    int k1_grid_max, k1_smem;
    int tiles_x, tiles_y, k2_smem;
    int view;
    FrameState* states;
    BatchFrame frames[kBatchMax];
};

__global__ void __launch_bounds__(32) batch_redo_kernel(const __grid_constant__ BatchRedoParams rp) {
    // one lane per frame looks at its flags; normally nothing is flagged and the kernel ends here
    static_assert(kBatchMax <= 32, "one lane per frame");
    const int lf = threadIdx.x;
    const bool mine = lf < rp.n_frames && (*reinterpret_cast<volatile unsigned*>(&rp.states[lf].flags) & kStatusTBounds);
    if (!__any_sync(0xffffffffu, mine)) return;
    if (threadIdx.x != 0) return;
    for (int f = 0; f < rp.n_frames; ++f) {
        FrameState* st = rp.states + f;
        const unsigned fl = *reinterpret_cast<volatile unsigned*>(&st->flags);
        if (!(fl & kStatusTBounds)) continue;
        st->flags = fl & ~(kStatusPixelOob | kStatusScatterOob);
        st->redo = 1;
        st->blocks_done = 0;
        st->next_chunk = 0;
        st->red_lo = 0;
        st->red_hi = 0;
        const BatchFrame& fr = rp.frames[f];
        long long b = (fr.n + 2047) / 2048;
        b = b < 1 ? 1 : (b > rp.sm_count * 8 ? rp.sm_count * 8 : b);
        bounds_reduce_kernel<false><<<static_cast<int>(b), 256, 0, cudaStreamTailLaunch>>>(fr.events, fr.n, rp.ev.polarity, st);
        EventParams q = rp.ev;
        q.events = fr.events;
        q.n = fr.n;
        q.epoch = rp.epoch_redo0 + f;
        q.state = st;
        q.bounds_mode = 2;
        q.arm_fixup = 0;
        q.use_pdl = 0;
        long long g = (fr.n + kEvChunk - 1) / kEvChunk;
        g = g < 1 ? 1 : (g > rp.k1_grid_max ? rp.k1_grid_max : g);
        if (rp.view == 1)
            events_lean_kernel<true><<<static_cast<int>(g), kWsThreads, rp.k1_smem, cudaStreamTailLaunch>>>(q);
        else
            events_lean_kernel<false><<<static_cast<int>(g), kWsThreads, rp.k1_smem, cudaStreamTailLaunch>>>(q);
        EpilogueParams e = rp.ep;
        e.dst = fr.dst;
        e.state = st;
        e.epoch = q.epoch - 1;  // the epilogue adds state->redo (= 1)
        e.recycle = nullptr;
        e.use_pdl = 0;
        if (rp.view == 1) {
            long long eb = (static_cast<long long>(e.out_w) * e.out_h + 511) / 512;
            eb = eb > rp.sm_count * 8 ? rp.sm_count * 8 : eb;
            epilogue_camera_kernel<<<static_cast<int>(eb), 256, 0, cudaStreamTailLaunch>>>(e);
        } else {
            epilogue_projector7_kernel<<<dim3(rp.tiles_x, rp.tiles_y), 256, rp.k2_smem, cudaStreamTailLaunch>>>(e);
        }
    }
}

#undef XM_B_YLIM
#undef XM_B_CAMW
#undef XM_B_CAMH

}  // namespace xm
