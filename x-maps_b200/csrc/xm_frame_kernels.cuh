// xm_frame_kernels.cuh — the per-frame hot path as sm_100a kernels.
//
//   K0  bounds_*        t.min() / t.max() of the polarity-masked events          (x_maps_disparity.py:12-13)
//   K1  events_*_kernel polarity mask -> rectify LUT -> time column -> X-map lookup -> disparity ->
//                       last-write-wins scatter as a 64-bit atomicMax            (depth_reprojection_pipe.py:114,128,142,150-160)
//   K2  epilogue_*      [7x7 dilate + nearest remap] -> depth / disparity / BGR  (disp_to_depth.py:46-115)
//
// Data layout in HBM (all owned by the context):
//   lut_xy    int32 [cam_h * cam_w]        low 16 bits = x_rect, high 16 bits = y_rect (one gather per event)
//   xmap_t    int16 [xmap_w][col_stride]   the X-map TRANSPOSED: one time column is contiguous, so the
//                                          columns a time-sorted chunk of events needs are ONE contiguous
//                                          range that a single 1-D bulk (TMA) copy stages into shared memory
//   remap_xy  int16 [proj_h * proj_w][2]   (x_rect, y_rect) of every projector pixel
//   map       u64   [max(rect, cam) cells] scatter keys  epoch:16 | event index:32 | disparity:16
#pragma once
#include "xm_device.cuh"

namespace xm {

// Device-resident per-context state block (mirrored to XmFrameStatus by xm_frame_status).
struct FrameState {
    long long t_lo_bits;  // bounds in use: int64, or float64 bit pattern
    long long t_hi_bits;
    unsigned long long red_lo;  // reduction scratch, encoded so that 0 is the identity of atomicMax:
    unsigned long long red_hi;  //   red_lo = max(~u), red_hi = max(u + 1), u = order-preserving unsigned image of t
    unsigned long long n_valid;
    unsigned long long n_inliers;
    unsigned int flags;       // XM_STATUS_*
    unsigned int redo;        // 1: the optimistic bounds were wrong, the fix-up pass is running / ran
    unsigned int epoch_used;  // (staged path only) epoch of the last scatter
    unsigned int blocks_done; // last-block detection
    unsigned int any_valid;   // reduction saw at least one valid event
    unsigned int next_chunk;  // dynamic chunk scheduler of the lean K1
    unsigned int next_tile;   // fused kernel: epilogue tile scheduler
    unsigned int fix_chunk;   // fused kernel: chunk scheduler of the fix-up pass
    unsigned int p2_ticket;   // batch kernel, strip epilogue: pass-2 item scheduler
    unsigned int p2_done;     //   ... and finished pass-2 items
    unsigned int pad_[2];
    // batch kernel: the frame's integer time-column constants (IntCol), worked out once by batch_bounds_kernel instead
    // of by every consumer warp (a 64-bit division): lo (2 words), range, 2 * scale | d, M, shift, ok
    int4 ic0, ic1;
};
static_assert(sizeof(FrameState) == 128 && offsetof(FrameState, ic0) == 96, "FrameState layout (16-byte aligned constants)");

constexpr unsigned kStatusTBounds = 0x1u;
constexpr unsigned kStatusPixelOob = 0x2u;
constexpr unsigned kStatusScatterOob = 0x4u;

constexpr int kEvThreads = 256;  // threads per CTA of the event kernels
constexpr int kEvPerThread = 4;  // events per thread per chunk (independent 128-bit loads in flight)
constexpr int kEvChunk = kEvThreads * kEvPerThread;

struct EventParams {
    const int4* events;
    long long n;
    int polarity;  // keep only p == 1
    const int* lut_xy;
    int cam_w, cam_h;
    const short* xmap_t;
    int xmap_w, xmap_h, col_stride;
    int t_px_scale, x_offset;
    int rect_w, rect_h;
    int view;  // 0 projector, 1 camera
    unsigned long long* map;
    unsigned epoch;
    FrameState* state;
    int cap_cols;     // X-map columns that fit the shared-memory window (0 = never stage)
    int lookahead;    // extra columns fetched ahead of a time-sorted stream
    int bounds_mode;  // 0: first / last valid event (found by every CTA), 1: given_lo / given_hi, 2: state->t_lo/hi_bits
    long long given_lo, given_hi;
    int use_pdl;      // 1: execute the griddepcontrol instructions (0 for device-launched fix-up grids)
    int arm_fixup;    // 1: the last CTA launches the exact fix-up when an event violated the assumed bounds
    int fix_reduce_grid;  // grid of the fix-up's bounds reduction
    int smem_bytes;       // dynamic shared memory of this launch (re-used by the fix-up launch)
    int stages;       // depth of the shared-memory event ring (1..kMaxStages)
    int win_stages;   // (warp-specialised K1) depth of the X-map window ring
    unsigned long long* dbg;  // optional per-CTA phase timestamps (8 words per CTA), or NULL
    int debug;        // timing experiments only (results WRONG): 1 = coalesced LUT gather, 2 = coalesced scatter cells
};

// order-preserving map of a timestamp (int64, or float64 bit pattern) to uint64
template <bool F64>
__device__ __forceinline__ unsigned long long time_to_ordered(long long t_bits) {
    const long long k = F64 ? f64_bits_to_sortable(t_bits) : t_bits;
    return static_cast<unsigned long long>(k) ^ 0x8000000000000000ULL;
}
template <bool F64>
__device__ __forceinline__ long long ordered_to_time(unsigned long long u) {
    const long long k = static_cast<long long>(u ^ 0x8000000000000000ULL);
    return F64 ? sortable_to_f64_bits(k) : k;
}

// Programmatic dependent launch (PDL): let the next kernel of the stream start its prologue while
// this one runs / wait until everything the previous kernel wrote is visible.  Both are no-ops for
// launches without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ bool event_valid(const EventFields& e, int polarity) {
    return !polarity || e.p == 1;
}

// ---------------------------------------------------------------------------------------------
// Bounds of a time-sorted frame: t of the first / last valid event.  Called by the first two warps
// of a CTA (warp 0 scans forward, warp 1 backward; normally one step each); the results land in
// out[0] / out[1] (0 if the frame has no valid event).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void scan_sorted_bounds(const int4* __restrict__ events, long long n, int polarity, int warp, int lane,
                                                   long long* out) {
    long long found = -1;
    if (warp == 0) {
        for (long long base = 0; base < n; base += 32) {
            long long i = base + lane;
            bool ok = false;
            if (i < n) ok = event_valid(unpack_event(__ldg(events + i)), polarity);
            unsigned m = __ballot_sync(0xffffffffu, ok);
            if (m) {
                found = base + (__ffs(m) - 1);
                break;
            }
        }
        if (lane == 0) out[0] = found >= 0 ? unpack_event(__ldg(events + found)).t_bits : 0;
    } else if (warp == 1) {
        for (long long top = n; top > 0; top -= 32) {
            long long i = top - 1 - lane;
            bool ok = false;
            if (i >= 0) ok = event_valid(unpack_event(__ldg(events + i)), polarity);
            unsigned m = __ballot_sync(0xffffffffu, ok);
            if (m) {
                found = top - 1 - (__ffs(m) - 1);
                break;
            }
        }
        if (lane == 0) out[1] = found >= 0 ? unpack_event(__ldg(events + found)).t_bits : 0;
    }
}

// K0a (stage-by-stage path only; the fused path finds its bounds inside K1): reset the state block
// and set the bounds.   mode 0: first / last valid event      mode 1: bounds given by the caller
__global__ void bounds_init_kernel(const int4* __restrict__ events, long long n, int polarity, int mode,
                                   long long given_lo, long long given_hi, unsigned epoch, FrameState* st) {
    if (threadIdx.x == 0) {
        st->n_valid = 0;
        st->n_inliers = 0;
        st->flags = 0;
        st->redo = 0;
        st->epoch_used = epoch;
        st->blocks_done = 0;
        st->any_valid = 0;
        st->red_lo = 0;
        st->red_hi = 0;
    }
    if (mode == 1) {
        if (threadIdx.x == 0) {
            st->t_lo_bits = given_lo;
            st->t_hi_bits = given_hi;
        }
        return;
    }
    scan_sorted_bounds(events, n, polarity, threadIdx.x >> 5, threadIdx.x & 31, &st->t_lo_bits);
}

// ---------------------------------------------------------------------------------------------
// K0b: exact min / max over all valid events (one extra pass over the stream).  The last CTA to
// finish publishes the result into state->t_lo_bits / t_hi_bits and clears the frame counters.
// Used by XM_TBOUNDS_REDUCE and, tail-launched from K1, by the fix-up of wrong optimistic bounds.
// ---------------------------------------------------------------------------------------------
template <bool F64>
__global__ void __launch_bounds__(256) bounds_reduce_kernel(const int4* __restrict__ events, long long n, int polarity,
                                                            FrameState* st) {
    unsigned long long lo = 0xffffffffffffffffULL, hi = 0;  // ordered domain
    bool any = false;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        // default L2 policy: the second pass (K1) re-reads these lines and may still find them in L2
        EventFields e = unpack_event(ld_event_plain(events + i));
        if (event_valid(e, polarity)) {
            const unsigned long long u = time_to_ordered<F64>(e.t_bits);
            lo = u < lo ? u : lo;
            hi = u > hi ? u : hi;
            any = true;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o);
        const unsigned long long h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    any = __any_sync(0xffffffffu, any);
    __shared__ unsigned s_last;
    if ((threadIdx.x & 31) == 0 && any) {
        atomicMax(&st->red_lo, ~lo);
        atomicMax(&st->red_hi, hi + 1ULL);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(&st->blocks_done, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        const unsigned long long rl = *reinterpret_cast<volatile unsigned long long*>(&st->red_lo);
        const unsigned long long rh = *reinterpret_cast<volatile unsigned long long*>(&st->red_hi);
        const bool found = rh != 0ULL;
        st->t_lo_bits = found ? ordered_to_time<F64>(~rl) : 0;
        st->t_hi_bits = found ? ordered_to_time<F64>(rh - 1ULL) : 0;
        st->any_valid = found ? 1u : 0u;
        st->blocks_done = 0;
        st->n_valid = 0;
        st->n_inliers = 0;
        st->next_chunk = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// K1: the fused per-event kernel.
//
// Each CTA owns one contiguous span of the event buffer and walks it in chunks of kEvChunk events.
//
//  * Event stream: the chunks are streamed into a ring of shared-memory stages with 1-D bulk async
//    copies (TMA, L2 evict-first hint) signalled on per-stage mbarriers; thread 0 keeps `stages`
//    chunks in flight, so HBM latency is covered by the ring and not by registers.
//  * Software pipeline: iteration c runs the FRONT half of chunk c+1 (take kEvPerThread events per
//    thread from the stage with conflict-free 128-bit shared loads, polarity mask, issue the
//    dependent packed-LUT gathers, compute the X-map time column, reduce the chunk's column range)
//    and then the BACK half of chunk c (X-map lookup, disparity, scatter), so the L2 latency of the
//    LUT gathers is hidden behind a whole chunk of work.
//  * X-map: for a time-sorted stream a chunk needs 1-3 adjacent time columns, which in the
//    transposed layout is ONE contiguous byte range; when it is not resident thread 0 stages it
//    with a single bulk copy one chunk ahead of its use, and lookups are shared-memory reads.
//    Chunks whose range does not fit (unsorted input) read the transposed table through L2.
//  * Scatter: 64-bit atomicMax whose key orders by event index (last write wins, deterministic).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxStages = 6;
constexpr int kEvSmemHeader = 256;             // mbarriers + reduction scratch
constexpr int kEvLutBytes = 2 * kEvChunk * 4;  // double-buffered gathered LUT words

__host__ __device__ inline int events_smem_bytes(int stages, int win_bytes) {
    return kEvSmemHeader + kEvLutBytes + stages * kEvChunk * 16 + win_bytes;
}

struct ChunkRegs {
    int col[kEvPerThread];  // X-map time column, -1 = event dropped
    int pix[kEvPerThread];  // camera pixel index
};

// Column range of a chunk: `lo` = min column as unsigned (dropped events contribute UINT_MAX),
// `hi` = max column (dropped events contribute -1).
struct ColRange {
    unsigned lo;
    int hi;
    __device__ __forceinline__ void add(int col) {
        lo = min(lo, static_cast<unsigned>(col));
        hi = max(hi, col);
    }
};

// FRONT half.  FULL: the chunk has kEvChunk events (no per-event bound check).  FAST: TimeCol::fast.
// The packed-LUT words are gathered asynchronously into `s_lut` (one cp.async group per chunk).
// The FAST variant is written branch-free so that the kEvPerThread independent dependency chains
// (shared load -> unpack -> float64 column) interleave; the rare exact-tie case is patched up after.
template <bool F64, bool FULL, bool FAST>
__device__ __forceinline__ void front_half(const EventParams& p, const TimeCol<F64>& tc, const int4* stage, int* s_lut, int tid,
                                           int limit, unsigned pol_mask, ChunkRegs& r, ColRange& range, unsigned& n_valid,
                                           unsigned& flags) {
    if (FAST) {
        int4 raw[kEvPerThread];
#pragma unroll
        for (int k = 0; k < kEvPerThread; ++k) raw[k] = stage[k * kEvThreads + tid];
        unsigned tie_mask = 0;
#pragma unroll
        for (int k = 0; k < kEvPerThread; ++k) {
            const bool in = FULL || k * kEvThreads + tid < limit;
            const unsigned ex = static_cast<unsigned>(raw[k].x) & 0xffffu, ey = static_cast<unsigned>(raw[k].x) >> 16;
            // polarity: keep p == 1 (pol_mask = 0xffff) or everything (pol_mask = 0)
            const bool valid = in && (((static_cast<unsigned>(raw[k].y) ^ 1u) & pol_mask) == 0u);
            const bool ok = valid && ex < static_cast<unsigned>(p.cam_w) && ey < static_cast<unsigned>(p.cam_h);
            const int pix = ok ? static_cast<int>(ey) * p.cam_w + static_cast<int>(ex) : 0;
            if (ok) cp_async_4(s_lut + k * kEvThreads + tid, p.lut_xy + ((p.debug & 1) ? ((k * kEvThreads + tid) & 0xffff) : pix));
            const long long t_bits = (static_cast<long long>(raw[k].w) << 32) | static_cast<unsigned>(raw[k].z);
            bool viol, tie;
            int cc = tc.column_fast_nb(t_bits, viol, tie);
            viol = viol && ok;
            tie_mask |= (ok && !viol && tie) ? (1u << k) : 0u;
            n_valid += valid ? 1u : 0u;
            flags |= (valid && !ok) ? kStatusPixelOob : 0u;  // the reference raises IndexError here
            flags |= viol ? kStatusTBounds : 0u;
            cc = ok ? (viol ? 0 : cc) : -1;
            r.col[k] = cc;
            r.pix[k] = pix;
        }
        if (tie_mask) {  // exact ties of the rounding: evaluate the reference's own expression
#pragma unroll
            for (int k = 0; k < kEvPerThread; ++k)
                if (tie_mask & (1u << k))
                    r.col[k] = tc.exact_from_bits((static_cast<long long>(raw[k].w) << 32) | static_cast<unsigned>(raw[k].z));
        }
#pragma unroll
        for (int k = 0; k < kEvPerThread; ++k) range.add(r.col[k]);
    } else {
#pragma unroll
        for (int k = 0; k < kEvPerThread; ++k) {
            int cc = -1;
            int pix = 0;
            if (FULL || k * kEvThreads + tid < limit) {
                const int4 raw = stage[k * kEvThreads + tid];
                if ((((static_cast<unsigned>(raw.y) & 0xffffu) ^ 1u) & pol_mask) == 0u) {
                    ++n_valid;
                    const unsigned ex = static_cast<unsigned>(raw.x) & 0xffffu, ey = static_cast<unsigned>(raw.x) >> 16;
                    if (ex < static_cast<unsigned>(p.cam_w) && ey < static_cast<unsigned>(p.cam_h)) {
                        pix = static_cast<int>(ey) * p.cam_w + static_cast<int>(ex);
                        cp_async_4(s_lut + k * kEvThreads + tid, p.lut_xy + pix);
                        const long long t_bits = (static_cast<long long>(raw.w) << 32) | static_cast<unsigned>(raw.z);
                        bool viol;
                        cc = tc.column(t_bits, viol);
                        if (cc < 0) cc += p.xmap_w;  // NumPy negative index (only reachable with wrong bounds)
                        viol = viol || cc < 0 || cc >= p.xmap_w;
                        if (viol) cc = 0;
                        flags |= viol ? kStatusTBounds : 0u;
                    } else {
                        flags |= kStatusPixelOob;  // the reference raises IndexError here
                    }
                }
            }
            r.col[k] = cc;
            r.pix[k] = pix;
            range.add(cc);
        }
    }
    cp_async_commit();
}


// ---------------------------------------------------------------------------------------------
// K1, warp-specialised general variant (k1_variant = 1; the lean integer-time kernel below is the default).
//
// The general per-event kernel (float64 or integer timestamps, verified or unverified tables), without any CTA-wide
// barrier in the steady state:
//   * warp 8 is the PRODUCER: one lane streams the CTA's span through a ring of event stages
//     (TMA bulk copies, `full_ev` / `empty_ev` mbarriers) and, for every chunk, stages the X-map time
//     columns between the columns of the chunk's first and last event into a ring of window buffers
//     (`full_win` / `empty_win`); the window's (first column, count) travels in shared memory.
//   * warps 0-7 are CONSUMERS: each runs the front half of chunk c+1 (events out of the stage,
//     stage released right away, LUT gathers started with cp.async) and then the back half of chunk c
//     (window lookup, disparity, scatter).  An event whose column is not in its chunk's window
//     (unsorted input, or more columns than the window holds) reads the X-map through L2 instead,
//     so no block-wide agreement on the column range is needed any more.
// Warps drift apart by up to the ring depth; the only synchronisation is mbarrier arrive / wait.
// ---------------------------------------------------------------------------------------------
constexpr int kWsThreads = kEvThreads + 32;
constexpr int kWsMaxStages = 4;

__host__ __device__ inline int events_ws_smem_bytes(int ev_stages, int win_stages, int win_bytes) {
    return kEvSmemHeader + kEvLutBytes + ev_stages * kEvChunk * 16 + win_stages * win_bytes;
}

template <bool SAFE>
__device__ __forceinline__ void back_ws(const EventParams& p, const ChunkRegs& r, const int* s_lut, const short* s_win, int win_lo,
                                        int win_n, int tid, unsigned idx_base, unsigned& n_inl, unsigned& flags) {
    const unsigned y_lim = static_cast<unsigned>(p.xmap_h - 1);
    int lut[kEvPerThread];
#pragma unroll
    for (int k = 0; k < kEvPerThread; ++k) lut[k] = s_lut[k * kEvThreads + tid];
    int xp[kEvPerThread];
    bool y_ok[kEvPerThread];
    unsigned miss = 0;
#pragma unroll
    for (int k = 0; k < kEvPerThread; ++k) {
        const int ycr = lut[k] >> 16;
        // x_maps_disparity.py:23: 0 <= y_rect < H - 1 (last row excluded)
        y_ok[k] = r.col[k] >= 0 && static_cast<unsigned>(ycr) < y_lim;
        const int rel = r.col[k] - win_lo;
        const bool in_win = static_cast<unsigned>(rel) < static_cast<unsigned>(win_n);
        xp[k] = s_win[(y_ok[k] && in_win) ? rel * p.col_stride + ycr : 0];
        miss |= (y_ok[k] && !in_win) ? (1u << k) : 0u;
    }
    if (miss) {  // column outside the staged window: read the transposed table through L2
#pragma unroll
        for (int k = 0; k < kEvPerThread; ++k)
            if (miss & (1u << k)) xp[k] = __ldg(p.xmap_t + static_cast<long long>(r.col[k]) * p.col_stride + (lut[k] >> 16));
    }
#pragma unroll
    for (int k = 0; k < kEvPerThread; ++k) {
        const int xcr = static_cast<short>(lut[k] & 0xffff);
        const int ycr = lut[k] >> 16;
        const int disp = static_cast<short>(xp[k] - xcr - p.x_offset);  // int16 arithmetic wraps
        bool inl = y_ok[k] && disp >= 0;
        n_inl += inl ? 1u : 0u;
        int cell;
        if (p.view == 1) {
            cell = r.pix[k];
        } else if (SAFE) {
            cell = ycr * p.rect_w + (xp[k] - p.x_offset);  // = x_rect + disp, in [0, rect_w) for verified tables
        } else {
            int xpr = static_cast<short>(xcr + disp);
            xpr += xpr < 0 ? p.rect_w : 0;  // NumPy negative index wraps once
            const bool in_map = xpr >= 0 && xpr < p.rect_w && ycr < p.rect_h;
            flags |= (inl && !in_map) ? kStatusScatterOob : 0u;  // the reference raises IndexError here
            inl = inl && in_map;
            cell = ycr * p.rect_w + xpr;
        }
        if (p.debug & 2) cell = static_cast<int>((idx_base + static_cast<unsigned>(k * kEvThreads)) & 0xfffffu);
        if (inl) atomicMax(p.map + cell, make_key32(p.epoch, idx_base + static_cast<unsigned>(k * kEvThreads), disp));
    }
}

template <bool F64, bool SAFE>
__global__ void __launch_bounds__(kWsThreads, 3) events_ws_kernel(const EventParams p) {
    extern __shared__ __align__(128) unsigned char ev_smem[];
    uint64_t* full_ev = reinterpret_cast<uint64_t*>(ev_smem);         // [kWsMaxStages]
    uint64_t* empty_ev = reinterpret_cast<uint64_t*>(ev_smem + 32);   // [kWsMaxStages]
    uint64_t* full_win = reinterpret_cast<uint64_t*>(ev_smem + 64);   // [kWsMaxStages]
    uint64_t* empty_win = reinterpret_cast<uint64_t*>(ev_smem + 96);  // [kWsMaxStages]
    int2* win_meta = reinterpret_cast<int2*>(ev_smem + 128);          // [kWsMaxStages] (first column, count)
    long long* s_bounds = reinterpret_cast<long long*>(ev_smem + 160);  // [2]
    int* s_lut = reinterpret_cast<int*>(ev_smem + kEvSmemHeader);     // [2][kEvChunk]
    unsigned char* ring = ev_smem + kEvSmemHeader + kEvLutBytes;
    const int win_bytes = p.cap_cols * p.col_stride * 2;
    unsigned char* win_ring = ring + p.stages * (kEvChunk * 16);

    FrameState* st = p.state;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const bool producer = warp == kEvThreads / 32;

    // span of this CTA: equal shares, boundaries on multiples of 32 events (512 B)
    const long long per = ((p.n + gridDim.x - 1) / gridDim.x + 31) & ~31LL;
    const long long span_lo = per * blockIdx.x < p.n ? per * blockIdx.x : p.n;
    const long long span_hi = span_lo + per < p.n ? span_lo + per : p.n;
    const int span_len = static_cast<int>(span_hi - span_lo);
    const int n_chunks = (span_len + kEvChunk - 1) / kEvChunk;
    const int4* span_ptr = p.events + span_lo;
    const unsigned pol_mask = p.polarity ? 0xffffu : 0u;

    if (p.use_pdl) pdl_launch_dependents();
    if (tid == 0) {
        for (int s = 0; s < kWsMaxStages; ++s) {
            mbar_init(full_ev + s, 1);
            mbar_init(empty_ev + s, kEvThreads / 32);
            mbar_init(full_win + s, 1);
            mbar_init(empty_win + s, kEvThreads / 32);
        }
    }
    // t.min() / t.max() of a time-sorted frame are its first / last valid event; every CTA looks them
    // up itself (two cached 512-byte reads) instead of waiting for a separate kernel
    if (p.bounds_mode == 0) scan_sorted_bounds(p.events, p.n, p.polarity, warp, lane, s_bounds);
    __syncthreads();
    // everything above only reads inputs; below this line the kernel touches the state block and the
    // scatter map, which the previous frame's epilogue may still be using
    if (p.use_pdl) pdl_wait();

    long long t_lo, t_hi;
    if (p.bounds_mode == 0) {
        t_lo = s_bounds[0];
        t_hi = s_bounds[1];
    } else if (p.bounds_mode == 1) {
        t_lo = p.given_lo;
        t_hi = p.given_hi;
    } else {
        t_lo = st->t_lo_bits;
        t_hi = st->t_hi_bits;
    }
    if (p.bounds_mode != 2 && blockIdx.x == 0 && tid == 0) {  // for xm_frame_status
        st->t_lo_bits = t_lo;
        st->t_hi_bits = t_hi;
    }
    TimeCol<F64> tc;
    tc.init(t_lo, t_hi, p.t_px_scale);

    unsigned n_valid = 0, n_inl = 0, flags = 0;

    if (producer) {
        if (lane == 0) {
            const uint64_t pol = make_evict_first_policy();
            int se = 0, sw = 0;
            unsigned pe = 0, pw = 0;  // parity of the empty barriers' phase to wait for (from the second round on)
            for (int c = 0; c < n_chunks; ++c) {
                const int first = c * kEvChunk;
                const int count = min(kEvChunk, span_len - first);
                // time stamps of the chunk's first and last record (any polarity) bound its columns
                long long ta = 0, tb = 0;
                if (p.cap_cols > 0) {
                    const int4 a = __ldg(span_ptr + first), b = __ldg(span_ptr + first + count - 1);
                    ta = (static_cast<long long>(a.w) << 32) | static_cast<unsigned>(a.z);
                    tb = (static_cast<long long>(b.w) << 32) | static_cast<unsigned>(b.z);
                }
                if (c >= p.stages) mbar_wait(empty_ev + se, pe);
                mbar_expect_tx(full_ev + se, static_cast<unsigned>(count) * 16u);
                tma_load_1d_hint(ring + se * (kEvChunk * 16), span_ptr + first, static_cast<unsigned>(count) * 16u, full_ev + se, pol);
                if (++se == p.stages) {
                    se = 0;
                    if (c >= p.stages) pe ^= 1u;
                }
                if (p.cap_cols > 0) {
                    bool va, vb;
                    int ca = tc.column(ta, va), cb = tc.column(tb, vb);
                    ca = min(max(ca, 0), p.xmap_w - 1);
                    cb = min(max(cb, 0), p.xmap_w - 1);
                    const int lo = min(ca, cb);
                    const int n = min(min(max(ca, cb) - lo + 1, p.cap_cols), p.xmap_w - lo);
                    if (c >= p.win_stages) mbar_wait(empty_win + sw, pw);
                    win_meta[sw] = make_int2(lo, n);
                    const unsigned bytes = static_cast<unsigned>(n) * p.col_stride * 2u;
                    mbar_expect_tx(full_win + sw, bytes);
                    tma_load_1d(win_ring + sw * win_bytes, p.xmap_t + static_cast<long long>(lo) * p.col_stride, bytes, full_win + sw);
                    if (++sw == p.win_stages) {
                        sw = 0;
                        if (c >= p.win_stages) pw ^= 1u;
                    }
                }
            }
        }
    } else {
        int fe = 0, bw = 0;        // ring positions of the next front (events) / back (window)
        unsigned fpe = 0, bpw = 0;  // parities of the full barriers
        auto front = [&](int c, ChunkRegs& r) {
            mbar_wait(full_ev + fe, fpe);
            const int4* stage = reinterpret_cast<const int4*>(ring + fe * (kEvChunk * 16));
            int* lut_dst = s_lut + (c & 1) * kEvChunk;
            const int limit = span_len - c * kEvChunk;
            ColRange range{0xffffffffu, -1};
            if (limit >= kEvChunk) {
                if (tc.fast)
                    front_half<F64, true, true>(p, tc, stage, lut_dst, tid, limit, pol_mask, r, range, n_valid, flags);
                else
                    front_half<F64, true, false>(p, tc, stage, lut_dst, tid, limit, pol_mask, r, range, n_valid, flags);
            } else {
                if (tc.fast)
                    front_half<F64, false, true>(p, tc, stage, lut_dst, tid, limit, pol_mask, r, range, n_valid, flags);
                else
                    front_half<F64, false, false>(p, tc, stage, lut_dst, tid, limit, pol_mask, r, range, n_valid, flags);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_ev + fe);  // this warp has its events in registers
            if (++fe == p.stages) {
                fe = 0;
                fpe ^= 1u;
            }
        };
        ChunkRegs cur;
        if (n_chunks > 0) front(0, cur);
        for (int c = 0; c < n_chunks; ++c) {
            ChunkRegs nxt;
            const bool has_next = c + 1 < n_chunks;
            if (has_next) {
                front(c + 1, nxt);
                cp_async_wait<1>();  // the gathers of chunk c have landed; those of chunk c+1 stay in flight
            } else {
                cp_async_wait<0>();
            }
            int win_lo = 0, win_n = 0;
            const short* s_win = reinterpret_cast<const short*>(win_ring);
            if (p.cap_cols > 0) {
                mbar_wait(full_win + bw, bpw);
                const int2 meta = win_meta[bw];
                win_lo = meta.x;
                win_n = meta.y;
                s_win = reinterpret_cast<const short*>(win_ring + bw * win_bytes);
            }
            const unsigned idx_base = static_cast<unsigned>(span_lo) + static_cast<unsigned>(c * kEvChunk + tid);
            back_ws<SAFE>(p, cur, s_lut + (c & 1) * kEvChunk, s_win, win_lo, win_n, tid, idx_base, n_inl, flags);
            if (p.cap_cols > 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_win + bw);
                if (++bw == p.win_stages) {
                    bw = 0;
                    bpw ^= 1u;
                }
            }
#pragma unroll
            for (int k = 0; k < kEvPerThread; ++k) {
                cur.col[k] = nxt.col[k];
                cur.pix[k] = nxt.pix[k];
            }
        }
    }

    // per-CTA statistics -> one atomic per warp
    n_valid = __reduce_add_sync(0xffffffffu, n_valid);
    n_inl = __reduce_add_sync(0xffffffffu, n_inl);
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0) {
        if (n_valid) atomicAdd(&st->n_valid, static_cast<unsigned long long>(n_valid));
        if (n_inl) atomicAdd(&st->n_inliers, static_cast<unsigned long long>(n_inl));
        if (flags) atomicOr(&st->flags, flags);
    }
    if (p.arm_fixup) {
        // last CTA: if any event violated the assumed bounds, launch the exact fix-up (see events_kernel)
        __shared__ unsigned s_last;
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            s_last = (atomicAdd(&st->blocks_done, 1u) == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last && tid == 0) {
            __threadfence();
            unsigned f = *reinterpret_cast<volatile unsigned*>(&st->flags);
            st->blocks_done = 0;
            if (f & kStatusTBounds) {
                st->redo = 1;
                st->flags = f & ~(kStatusPixelOob | kStatusScatterOob);
                bounds_reduce_kernel<F64><<<p.fix_reduce_grid, 256, 0, cudaStreamTailLaunch>>>(p.events, p.n, p.polarity, st);
                EventParams q = p;
                q.epoch = p.epoch + 1;
                q.arm_fixup = 0;
                q.use_pdl = 0;
                q.bounds_mode = 2;
                events_ws_kernel<F64, SAFE><<<gridDim.x, kWsThreads, p.smem_bytes, cudaStreamTailLaunch>>>(q);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1, lean variant: the warp-specialised pipeline of events_ws_kernel with the per-event
// instruction stream cut down for the common case -- integer timestamps (IntCol: no float64 in the
// hot loop), verified tables (no scatter bound checks), raw shared-memory addresses hoisted out of
// the loop, validity / inlier bookkeeping as bit masks.  Every exceptional event (outside the assumed
// time bounds, exact rounding tie, pixel outside the image, column outside the staged window) is
// flagged in a mask and handled by a slow path that evaluates the reference's own expressions.
// CAM: camera-view scatter (cell = event pixel) instead of projector view.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <bool CAM>
__device__ __forceinline__ void lean_events_phase(const EventParams& p, unsigned char* ev_smem, unsigned& n_valid, unsigned& n_inl,
                                                  unsigned& flags) {
    if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 0] = global_timer_ns();
    uint64_t* full_ev = reinterpret_cast<uint64_t*>(ev_smem);
    uint64_t* empty_ev = reinterpret_cast<uint64_t*>(ev_smem + 32);
    uint64_t* full_win = reinterpret_cast<uint64_t*>(ev_smem + 64);
    uint64_t* empty_win = reinterpret_cast<uint64_t*>(ev_smem + 96);
    int2* win_meta = reinterpret_cast<int2*>(ev_smem + 128);   // [4] (first column, count) of a window stage
    int* ev_meta = reinterpret_cast<int*>(ev_smem + 160);      // [4] global chunk index of an event stage, -1 = no more work
    long long* s_bounds = reinterpret_cast<long long*>(ev_smem + 176);
    uint64_t* bounds_bar = reinterpret_cast<uint64_t*>(ev_smem + 192);
    unsigned char* ring = ev_smem + kEvSmemHeader + kEvLutBytes;
    const int win_bytes = p.cap_cols * p.col_stride * 2;
    unsigned char* win_ring = ring + p.stages * (kEvChunk * 16);

    FrameState* st = p.state;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const bool producer = warp == kEvThreads / 32;
    const int total_chunks = static_cast<int>((p.n + kEvChunk - 1) / kEvChunk);  // chunks are handed out dynamically
    const unsigned pol_mask = p.polarity ? 0xffffu : 0u;

    if (p.use_pdl) pdl_launch_dependents();
    if (tid == 0) {
        for (int s = 0; s < kWsMaxStages; ++s) {
            mbar_init(full_ev + s, 1);
            mbar_init(empty_ev + s, kEvThreads / 32);
            mbar_init(full_win + s, 1);
            mbar_init(empty_win + s, kEvThreads / 32);
        }
        mbar_init(bounds_bar, p.bounds_mode == 0 ? 2 : 1);
    }
    __syncthreads();

    // ---- producer state (lane 0 of the last warp) ---------------------------------------------------
    int se = 0, sw = 0;
    unsigned pe = 0, pw = 0;
    int issued = 0;       // event stages filled so far
    bool drained = false;  // the global chunk counter ran out
    const uint64_t pol = make_evict_first_policy();
    // grabs the next chunk and streams it into the ring; returns its index (-1: none left, sentinel posted)
    auto produce_events = [&]() -> int {
        const int g = drained ? total_chunks : static_cast<int>(atomicAdd(&st->next_chunk, 1u));
        if (issued >= p.stages) mbar_wait(empty_ev + se, pe);
        int ret = g;
        if (g >= total_chunks) {
            drained = true;
            ev_meta[se] = -1;
            mbar_arrive(full_ev + se);  // completes the phase: consumers wake up and see the sentinel
            ret = -1;
        } else {
            const long long first = static_cast<long long>(g) * kEvChunk;
            const unsigned count = static_cast<unsigned>(min(static_cast<long long>(kEvChunk), p.n - first));
            ev_meta[se] = g;
            mbar_expect_tx(full_ev + se, count * 16u);
            tma_load_1d_hint(ring + se * (kEvChunk * 16), p.events + first, count * 16u, full_ev + se, pol);
        }
        ++issued;
        if (++se == p.stages) {
            se = 0;
            if (issued > p.stages) pe ^= 1u;
        }
        return ret;
    };

    // Without programmatic launch nothing else is running: start streaming the first chunks before the
    // time bounds are known (the X-map windows, which need them, follow below).
    int early[kWsMaxStages];
    int n_early = 0;
    if (producer && lane == 0 && !p.use_pdl) {
        for (; n_early < p.stages; ++n_early) {
            early[n_early] = produce_events();
            if (early[n_early] < 0) {
                ++n_early;
                break;
            }
        }
    }

    // ---- time bounds ----------------------------------------------------------------------------
    // t.min() / t.max() of a time-sorted frame are its first / last valid event: warps 0 and 1 look them up
    if (p.bounds_mode == 0) {
        if (warp < 2) {
            scan_sorted_bounds(p.events, p.n, p.polarity, warp, lane, s_bounds);
            __syncwarp();
            if (lane == 0) mbar_arrive(bounds_bar);
        }
    } else if (tid == 0) {
        mbar_arrive(bounds_bar);
    }
    mbar_wait(bounds_bar, 0);
    // everything above only reads inputs (and, without PDL, this frame's own chunk counter); below the
    // kernel touches the state block and the scatter map, which the previous frame's epilogue may still use
    if (p.use_pdl) pdl_wait();

    long long t_lo, t_hi;
    if (p.bounds_mode == 0) {
        t_lo = s_bounds[0];
        t_hi = s_bounds[1];
    } else if (p.bounds_mode == 1) {
        t_lo = p.given_lo;
        t_hi = p.given_hi;
    } else {
        t_lo = st->t_lo_bits;
        t_hi = st->t_hi_bits;
    }
    if (p.bounds_mode != 2 && blockIdx.x == 0 && tid == 0) {
        st->t_lo_bits = t_lo;
        st->t_hi_bits = t_hi;
    }
    TimeCol<false> tc;  // float64 path: slow cases and the producer's window columns
    tc.init(t_lo, t_hi, p.t_px_scale);
    IntCol ic;
    ic.init(t_lo, t_hi, p.t_px_scale);
    if (p.dbg && tid == 0) p.dbg[blockIdx.x * 8 + 1] = global_timer_ns();

    if (producer) {
        if (lane == 0) {
            // stages the X-map columns between the columns of the chunk's first and last record
            auto produce_window = [&](int g) {
                const long long first = static_cast<long long>(g) * kEvChunk;
                const long long last = min(p.n, first + kEvChunk) - 1;
                const int4 a = __ldg(p.events + first), b = __ldg(p.events + last);
                const long long ta = (static_cast<long long>(a.w) << 32) | static_cast<unsigned>(a.z);
                const long long tb = (static_cast<long long>(b.w) << 32) | static_cast<unsigned>(b.z);
                bool va, vb;
                int ca = tc.column(ta, va), cb = tc.column(tb, vb);
                ca = min(max(ca, 0), p.xmap_w - 1);
                cb = min(max(cb, 0), p.xmap_w - 1);
                const int lo = min(ca, cb);
                const int n = min(min(max(ca, cb) - lo + 1, p.cap_cols), p.xmap_w - lo);
                if (g >= 0) {
                }
                mbar_wait(empty_win + sw, pw ^ 1u);  // first round: passes immediately (phase -1 counts as done)
                win_meta[sw] = make_int2(lo, n);
                const unsigned bytes = static_cast<unsigned>(n) * p.col_stride * 2u;
                mbar_expect_tx(full_win + sw, bytes);
                tma_load_1d(win_ring + sw * win_bytes, p.xmap_t + static_cast<long long>(lo) * p.col_stride, bytes, full_win + sw);
                if (++sw == p.win_stages) {
                    sw = 0;
                    pw ^= 1u;
                }
            };
            bool more = true;
            for (int i = 0; i < n_early; ++i) {
                if (early[i] < 0) {
                    more = false;
                    break;
                }
                if (p.cap_cols > 0) produce_window(early[i]);
            }
            while (more) {
                const int g = produce_events();
                if (g < 0) break;
                if (p.cap_cols > 0) produce_window(g);
            }
        }
    } else {
        // raw shared addresses, computed once
        const unsigned sbase = smem_u32(ev_smem);
        const unsigned a_full_ev = sbase, a_empty_ev = sbase + 32, a_full_win = sbase + 64, a_empty_win = sbase + 96;
        const unsigned a_wmeta = sbase + 128, a_emeta = sbase + 160;
        const unsigned a_lut = sbase + kEvSmemHeader + tid * 4;                  // + parity * 4096 + k * 1024
        const unsigned a_ring = sbase + kEvSmemHeader + kEvLutBytes + tid * 16;  // + slot * 16384 + k * 4096
        const unsigned a_win = sbase + kEvSmemHeader + kEvLutBytes + p.stages * (kEvChunk * 16);
        const unsigned y_lim = static_cast<unsigned>(p.xmap_h - 1);
        const unsigned cam_w = static_cast<unsigned>(p.cam_w), cam_h = static_cast<unsigned>(p.cam_h);
        const int col_stride = p.col_stride, x_offset = p.x_offset, rect_w = p.rect_w;
        const int* const lut_xy = p.lut_xy;
        unsigned long long* const map = p.map;
        const unsigned epoch16 = p.epoch << 16;
        const bool ic_ok = ic.ok;

        int fe = 0, bw = 0;
        unsigned fpe = 0, bpw = 0;
        unsigned fpar = 0;  // LUT double-buffer half the next front half writes

        // FRONT half of the next chunk: events out of the stage, LUT gathers started, columns computed.
        // Returns the chunk's global index, -1 when the producer has run out of work.
        auto front = [&](int (&col)[kEvPerThread], int (&pix)[kEvPerThread]) -> int {
            mbar_wait_a(a_full_ev + fe * 8, fpe);
            const int g = lds32_a(a_emeta + fe * 4);
            if (g < 0) return -1;
            const unsigned a_stage = a_ring + fe * (kEvChunk * 16);
            const unsigned a_lut_c = a_lut + fpar * (kEvChunk * 4);
            fpar ^= 1u;
            const long long left = p.n - static_cast<long long>(g) * kEvChunk;
            const int limit = left < kEvChunk ? static_cast<int>(left) : kEvChunk;
            int4 raw[kEvPerThread];
#pragma unroll
            for (int k = 0; k < kEvPerThread; ++k) raw[k] = lds128_a(a_stage + k * (kEvThreads * 16));
            unsigned vmask = 0, bad_mask = 0;
#pragma unroll
            for (int k = 0; k < kEvPerThread; ++k) {
                const unsigned ex = static_cast<unsigned>(raw[k].x) & 0xffffu, ey = static_cast<unsigned>(raw[k].x) >> 16;
                bool valid = ((static_cast<unsigned>(raw[k].y) ^ 1u) & pol_mask) == 0u;  // polarity: p == 1, or everything
                if (limit < kEvChunk) valid = valid && (k * kEvThreads + tid < limit);
                const bool ok = valid && ex < cam_w && ey < cam_h;
                const int px = static_cast<int>(ey * cam_w + ex);
                if (ok) cp_async_4_a(a_lut_c + k * (kEvThreads * 4), lut_xy + ((p.debug & 4) ? (px & 0x3fff) : ((p.debug & 1) ? ((k * kEvThreads + tid) & 0xffff) : px)));
                const long long t_bits = (static_cast<long long>(raw[k].w) << 32) | static_cast<unsigned>(raw[k].z);
                bool bad;
                const unsigned q = ic.column(t_bits, bad);
                col[k] = ok ? static_cast<int>(q) : -1;
                if (CAM) pix[k] = px;
                vmask |= valid ? (1u << k) : 0u;
                bad_mask |= (valid && (!ok || bad || !ic_ok)) ? (1u << k) : 0u;
            }
            cp_async_commit();
            n_valid += __popc(vmask);
            if (bad_mask) {  // slow path: the reference's own float64 expression / error flags
#pragma unroll
                for (int k = 0; k < kEvPerThread; ++k) {
                    if (!(bad_mask & (1u << k))) continue;
                    if (col[k] < 0) {
                        flags |= kStatusPixelOob;  // the reference raises IndexError here
                        continue;
                    }
                    const long long t_bits = (static_cast<long long>(raw[k].w) << 32) | static_cast<unsigned>(raw[k].z);
                    bool viol;
                    int cc = tc.column(t_bits, viol);
                    if (cc < 0) cc += p.xmap_w;  // NumPy negative index (only reachable with wrong bounds)
                    viol = viol || cc < 0 || cc >= p.xmap_w;
                    if (viol) {
                        flags |= kStatusTBounds;
                        cc = 0;
                    }
                    col[k] = cc;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_a(a_empty_ev + fe * 8);  // this warp has its events in registers
            if (++fe == p.stages) {
                fe = 0;
                fpe ^= 1u;
            }
            return g;
        };

        int col_cur[kEvPerThread], pix_cur[kEvPerThread];
        int g_cur = front(col_cur, pix_cur);
        unsigned bpar = 0;  // LUT double-buffer half the next back half reads
        if (p.dbg && tid == 0) p.dbg[blockIdx.x * 8 + 2] = global_timer_ns();
        int iter = 0;
        while (g_cur >= 0) {
            if (p.dbg && tid == 0 && iter == 1) p.dbg[blockIdx.x * 8 + 3] = global_timer_ns();
            ++iter;
            int col_nxt[kEvPerThread], pix_nxt[kEvPerThread];
            const int g_nxt = front(col_nxt, pix_nxt);
            if (g_nxt >= 0)
                cp_async_wait<1>();  // the gathers of the current chunk have landed; the next chunk's stay in flight
            else
                cp_async_wait<0>();
            // ---- BACK half of the current chunk ----------------------------------------------------------
            int win_lo = 0;
            unsigned win_n = 0;
            unsigned a_win_c = a_lut;  // any valid address: without a window every lookup misses
            if (p.cap_cols > 0) {
                mbar_wait_a(a_full_win + bw * 8, bpw);
                win_lo = lds32_a(a_wmeta + bw * 8);
                win_n = static_cast<unsigned>(lds32_a(a_wmeta + bw * 8 + 4));
                a_win_c = a_win + bw * win_bytes;
            }
            const unsigned a_lut_c = a_lut + bpar * (kEvChunk * 4);
            bpar ^= 1u;
            int lut[kEvPerThread], xp[kEvPerThread];
#pragma unroll
            for (int k = 0; k < kEvPerThread; ++k) lut[k] = lds32_a(a_lut_c + k * (kEvThreads * 4));
            unsigned hit_mask = 0, miss_mask = 0;
#pragma unroll
            for (int k = 0; k < kEvPerThread; ++k) {
                const unsigned ycr = static_cast<unsigned>(lut[k] >> 16);
                const unsigned rel = static_cast<unsigned>(col_cur[k] - win_lo);  // dropped events (col = -1) wrap to huge
                const bool y_ok = ycr < y_lim;  // x_maps_disparity.py:23: 0 <= y_rect < H - 1 (last row excluded)
                const bool hit = y_ok && rel < win_n;
                xp[k] = lds_s16_a(a_win_c + (hit ? (rel * col_stride + ycr) * 2u : 0u));
                hit_mask |= hit ? (1u << k) : 0u;
                miss_mask |= (y_ok && !hit && col_cur[k] >= 0) ? (1u << k) : 0u;
            }
            if (miss_mask) {  // column outside the staged window: read the transposed table through L2
#pragma unroll
                for (int k = 0; k < kEvPerThread; ++k)
                    if (miss_mask & (1u << k))
                        xp[k] = __ldg(p.xmap_t + static_cast<long long>(col_cur[k]) * col_stride + (lut[k] >> 16));
                hit_mask |= miss_mask;
            }
            const unsigned idx0 = static_cast<unsigned>(g_cur) * kEvChunk + static_cast<unsigned>(tid);
            unsigned imask = 0;
#pragma unroll
            for (int k = 0; k < kEvPerThread; ++k) {
                const int xcr = static_cast<short>(lut[k] & 0xffff);
                const int ycr = lut[k] >> 16;
                const int disp = static_cast<short>(xp[k] - xcr - x_offset);  // int16 arithmetic wraps
                const bool inl = ((hit_mask >> k) & 1u) && disp >= 0;
                // projector view: x_rect + disp = x_map - x_offset, in [0, rect_w) for verified tables
                const int cell = CAM ? pix_cur[k] : ycr * rect_w + (xp[k] - x_offset);
                const unsigned idx = idx0 + static_cast<unsigned>(k * kEvThreads);
                const unsigned long long key =
                    (static_cast<unsigned long long>(epoch16 | (idx >> 16)) << 32) | ((idx << 16) | static_cast<unsigned>(disp));
                red_max_u64_if(map + ((p.debug & 2) ? static_cast<int>(idx & 0xfffffu) : cell), key, inl);
                imask |= inl ? (1u << k) : 0u;
            }
            n_inl += __popc(imask);
            if (p.cap_cols > 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive_a(a_empty_win + bw * 8);
                if (++bw == p.win_stages) {
                    bw = 0;
                    bpw ^= 1u;
                }
            }
#pragma unroll
            for (int k = 0; k < kEvPerThread; ++k) {
                col_cur[k] = col_nxt[k];
                if (CAM) pix_cur[k] = pix_nxt[k];
            }
            g_cur = g_nxt;
        }
    }

}

template <bool CAM>
__global__ void __launch_bounds__(kWsThreads, 3) events_lean_kernel(const EventParams p) {
    extern __shared__ __align__(128) unsigned char ev_smem[];
    FrameState* st = p.state;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    unsigned n_valid = 0, n_inl = 0, flags = 0;
    lean_events_phase<CAM>(p, ev_smem, n_valid, n_inl, flags);
    if (p.dbg && tid == 0) p.dbg[blockIdx.x * 8 + 5] = global_timer_ns();
    n_valid = __reduce_add_sync(0xffffffffu, n_valid);
    n_inl = __reduce_add_sync(0xffffffffu, n_inl);
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0) {
        if (n_valid) atomicAdd(&st->n_valid, static_cast<unsigned long long>(n_valid));
        if (n_inl) atomicAdd(&st->n_inliers, static_cast<unsigned long long>(n_inl));
        if (flags) atomicOr(&st->flags, flags);
    }
    if (p.arm_fixup) {
        __shared__ unsigned s_last;
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            s_last = (atomicAdd(&st->blocks_done, 1u) == gridDim.x - 1);
        }
        __syncthreads();
        if (p.dbg && tid == 0) {
            p.dbg[blockIdx.x * 8 + 6] = global_timer_ns();
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.dbg[blockIdx.x * 8 + 7] = smid;
        }
        if (s_last && tid == 0) {
            __threadfence();
            unsigned f = *reinterpret_cast<volatile unsigned*>(&st->flags);
            st->blocks_done = 0;
            if (f & kStatusTBounds) {
                st->redo = 1;
                st->flags = f & ~(kStatusPixelOob | kStatusScatterOob);
                bounds_reduce_kernel<false><<<p.fix_reduce_grid, 256, 0, cudaStreamTailLaunch>>>(p.events, p.n, p.polarity, st);
                EventParams q = p;
                q.epoch = p.epoch + 1;
                q.arm_fixup = 0;
                q.use_pdl = 0;
                q.bounds_mode = 2;
                events_lean_kernel<CAM><<<gridDim.x, kWsThreads, p.smem_bytes, cudaStreamTailLaunch>>>(q);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2 (camera view): decode the key of every camera pixel and emit.
// ---------------------------------------------------------------------------------------------
struct EpilogueParams {
    const unsigned long long* map;
    const FrameState* state;
    unsigned epoch;          // epoch of the frame's first pass; the fix-up pass (state->redo) used epoch + 1
    FrameState* recycle;     // state block of the frame after next: cleared here (ring of 4), or NULL
    int use_pdl;             // 1: execute the griddepcontrol instructions
    const short2* remap_xy;
    const short4* tile_box;  // per 32x32 output tile: bounding box (x0, y0, x1, y1) of its remap targets; x1 < 0: none
    // per output pixel: cell of the pixel's remap target inside its tile's shared-memory region (row stride and padding
    // as tile_region() lays it out), 0xffff = outside the rectified image; NULL: derive it from remap_xy per pixel
    const unsigned short* tile_off;
    int rect_w, rect_h;
    int out_w, out_h;  // projector (view 0) or camera (view 1) size
    int radius;        // dilate / 2
    int region_cap;    // cells of the shared-memory region (per buffer)
    OutputSpec out;
    void* dst;
};

__device__ __forceinline__ void recycle_state(FrameState* st) {
    unsigned long long* w = reinterpret_cast<unsigned long long*>(st);
#pragma unroll
    for (int i = 0; i < static_cast<int>(sizeof(FrameState) / 8); ++i) w[i] = 0ULL;
}

__global__ void __launch_bounds__(256) epilogue_camera_kernel(const EpilogueParams p) {
    if (p.use_pdl) {
        pdl_launch_dependents();
        pdl_wait();
    }
    const unsigned epoch = p.epoch + p.state->redo;
    if (p.recycle && blockIdx.x == 0 && threadIdx.x == 0) recycle_state(p.recycle);
    const long long n = static_cast<long long>(p.out_w) * p.out_h;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        emit_pixel_int(p.out, p.dst, i, key_disparity(p.map[i], epoch));
    }
}

// ---------------------------------------------------------------------------------------------
// K2 (projector view): out[v,u] = max over the (2r+1)^2 window around remap(v,u) of the rectified
// disparity map, 0 if remap(v,u) is outside it  ==  cv2.dilate + cv2.remap(NEAREST, BORDER_CONSTANT)
// (disp_to_depth.py:76-97).  One CTA per 32x32 output tile: the tile's source bounding box (+ halo)
// is decoded from the key map into shared memory once, dilated horizontally in shared memory, and
// each output pixel then takes a (2r+1)-tap vertical max.  Tiles whose box does not fit the
// shared-memory region fall back to direct window reads.
// ---------------------------------------------------------------------------------------------
constexpr int kTile = 32;

// R > 0: compile-time dilate radius (unrolled taps); R == 0: runtime radius p.radius.
template <int R>
__global__ void __launch_bounds__(256, 7) epilogue_projector_kernel(const EpilogueParams p) {
    extern __shared__ __align__(128) unsigned char ev_smem[];
    unsigned short* s_raw = reinterpret_cast<unsigned short*>(ev_smem);
    unsigned short* s_h = s_raw + p.region_cap;

    if (p.use_pdl) pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int u0 = blockIdx.x * kTile, v0 = blockIdx.y * kTile;
    const int r = R > 0 ? R : p.radius;

    // the tile's source bounding box was computed when the remap table was uploaded, so the region
    // loads below do not have to wait for the remap loads (one dependent memory round trip less)
    const short4 box = __ldg(p.tile_box + blockIdx.y * gridDim.x + blockIdx.x);
    const int x0 = box.x, y0 = box.y, x1 = box.z, y1 = box.w;
    if (p.out.kind == 0 && p.out.depth_lut && tid < 128) prefetch_l1(p.out.depth_lut + tid * 32);  // first 16 KB of the depth table

    short2 m[4];
    unsigned inside = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int u = u0 + lane, v = v0 + warp + k * 8;
        m[k] = make_short2(-1, -1);
        if (u < p.out_w && v < p.out_h) m[k] = __ldg(p.remap_xy + v * p.out_w + u);
    }

    // up to here only calibration tables were read; the key map and the state block belong to K1
    if (p.use_pdl) pdl_wait();
    const unsigned epoch = p.epoch + p.state->redo;
    if (p.recycle && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) recycle_state(p.recycle);

    int val[4] = {0, 0, 0, 0};
    if (x1 >= 0) {
        // region = bounding box + halo; cells outside the image count as 0 (disparities are >= 0,
        // so ignoring a tap, as cv2.dilate does at the border, equals reading 0)
        const int rx0 = x0 - r, ry0 = y0 - r;
        const int rw = x1 - x0 + 1 + 2 * r, rh = y1 - y0 + 1 + 2 * r;
        if (rw * rh <= p.region_cap) {
            // flat index -> (row, col) with a multiply-high instead of a division (exact: c * rw < 2^32)
            const unsigned magic_w = 0xffffffffu / static_cast<unsigned>(rw) + 1u;
            for (int c = tid; c < rw * rh; c += 256) {
                const int ry = static_cast<int>(__umulhi(static_cast<unsigned>(c), magic_w));
                const int rx = c - ry * rw;
                const int gx = rx0 + rx, gy = ry0 + ry;
                unsigned short d = 0;
                if (gx >= 0 && gx < p.rect_w && gy >= 0 && gy < p.rect_h)
                    d = static_cast<unsigned short>(key_disparity(p.map[gy * p.rect_w + gx], epoch));
                s_raw[c] = d;
            }
            __syncthreads();
            // horizontal max, only for the columns the vertical pass can read: rx in [r, rw - r)
            const int iw = rw - 2 * r;
            const unsigned magic_i = 0xffffffffu / static_cast<unsigned>(iw) + 1u;
            for (int c = tid; c < iw * rh; c += 256) {
                const int ry = static_cast<int>(__umulhi(static_cast<unsigned>(c), magic_i));
                const int rx = c - ry * iw + r;
                const unsigned short* row = s_raw + ry * rw + rx;
                unsigned best = 0;
                if (R > 0) {
#pragma unroll
                    for (int d = -R; d <= R; ++d) best = max(best, static_cast<unsigned>(row[d]));
                } else {
                    for (int d = -r; d <= r; ++d) best = max(best, static_cast<unsigned>(row[d]));
                }
                s_h[ry * rw + rx] = static_cast<unsigned short>(best);
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (m[k].x >= 0 && m[k].x < p.rect_w && m[k].y >= 0 && m[k].y < p.rect_h) inside |= 1u << k;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!(inside & (1u << k))) continue;
                const unsigned short* colp = s_h + (m[k].y - ry0) * rw + (m[k].x - rx0);
                unsigned best = 0;
                if (R > 0) {
#pragma unroll
                    for (int d = -R; d <= R; ++d) best = max(best, static_cast<unsigned>(colp[d * rw]));
                } else {
                    for (int d = -r; d <= r; ++d) best = max(best, static_cast<unsigned>(colp[d * rw]));
                }
                val[k] = static_cast<int>(best);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (m[k].x >= 0 && m[k].x < p.rect_w && m[k].y >= 0 && m[k].y < p.rect_h) inside |= 1u << k;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!(inside & (1u << k))) continue;
                int best = 0;
                for (int dy = -r; dy <= r; ++dy) {
                    const int gy = m[k].y + dy;
                    if (gy < 0 || gy >= p.rect_h) continue;
                    for (int dx = -r; dx <= r; ++dx) {
                        const int gx = m[k].x + dx;
                        if (gx < 0 || gx >= p.rect_w) continue;
                        best = max(best, key_disparity(p.map[gy * p.rect_w + gx], epoch));
                    }
                }
                val[k] = best;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int u = u0 + lane, v = v0 + warp + k * 8;
        if (u < p.out_w && v < p.out_h) emit_pixel_int(p.out, p.dst, v * p.out_w + u, val[k]);
    }
}

// ---------------------------------------------------------------------------------------------
// K2 (projector view), 7x7 dilate, even rect_w: the hot variant.
//
// Same tiling, but the dilation is done as a full separable pass over the tile's source region with
// register sliding windows instead of 7 taps per cell:
//   A  region (bounding box + halo, left edge aligned to an even x) decoded from the key map with
//      128-bit loads (two cells) into shared memory as packed u16 pairs, rows padded with zeros
//   B  horizontal 7-max: a thread takes 16 cells of a row as 8 words and produces 8 outputs with
//      the 2-4-7 max ladder (32 max ops instead of 48 loads + 48 max)
//   C  vertical 7-max on column PAIRS with per-halfword SIMD max (__vmaxu2): 14 words in, 8 x 2
//      outputs, same ladder
//   D  every output pixel reads its dilated value at its own (x_rect, y_rect) and converts.
// ---------------------------------------------------------------------------------------------
#ifndef XM_ROW_PHASE_A
#define XM_ROW_PHASE_A 1
#endif
constexpr int kRowPad = 4;    // zero cells left of the data in every shared-memory row
constexpr int kRowExtra = 16;  // total padding per row (4 left + 12 right)

__device__ __forceinline__ unsigned magic_div(int d) { return 0xffffffffu / static_cast<unsigned>(d) + 1u; }

// One 32x32 output tile, processed by a group of NT threads (NT = 64, 128 or 256; `gtid` = index inside
// the group) that synchronises on the named barrier `bar_id`: several groups of one CTA can work on
// different tiles at the same time, which is what hides the L2 latency of the region loads.
template <int NT>
__device__ __forceinline__ void group_sync(int bar_id) {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NT) : "memory");
}

// Geometry of a tile's source region in shared memory.
struct TileRegion {
    int rx0, ry0, rw, rh, stride;
    bool any;   // the tile maps into the rectified image at all
    bool fits;  // the region fits the shared-memory buffers (else: direct 49-tap reads)
};

__device__ __forceinline__ TileRegion tile_region(const EpilogueParams& p, const short4& box) {
    constexpr int R = 3;
    TileRegion g;
    const int x0 = box.x, y0 = box.y, x1 = box.z, y1 = box.w;
    g.any = x1 >= 0;
    g.rx0 = (x0 - R) & ~1;               // even
    g.rw = (x1 + R - g.rx0 + 2) & ~1;    // even number of data cells
    g.ry0 = y0 - R;
    g.rh = y1 - y0 + 1 + 2 * R;
    g.stride = g.rw + kRowExtra;         // cells per shared-memory row
    g.fits = g.any && g.stride * g.rh <= p.region_cap;
    return g;
}

// Phases A-C for a region that fits: afterwards bufA holds the 7x7-dilated disparities of the region
// (row stride g.stride, kRowPad zero cells left of the data); ends with a group barrier.
template <int NT, int UA>
__device__ __forceinline__ void proj7_dilate_region(const EpilogueParams& p, const TileRegion& g, unsigned epoch, unsigned short* bufA,
                                                    unsigned short* bufB, int tid, int bar_id) {
    constexpr int R = 3;
    const int rx0 = g.rx0, ry0 = g.ry0, rw = g.rw, rh = g.rh, stride = g.stride;
    {
        {
            // ---- A: decode the region (pads included, written as zeros) -------------------------
            {
                const int pairs = stride >> 1;
                unsigned* dst = reinterpret_cast<unsigned*>(bufA);
                const unsigned ep16 = epoch & 0xffffu;
                if (XM_ROW_PHASE_A && pairs <= 32) {
                    // one warp per region row, one lane per cell pair: everything that depends on x is a per-lane
                    // constant and a row costs one 128-bit load, two epoch tests and one shared store per lane;
                    // UA rows are in flight per warp
                    constexpr int NW = NT / 32;
                    const int lane = tid & 31, wrp = tid >> 5;
                    const int cx = 2 * lane - kRowPad;  // data coordinate of the pair's first cell (even)
                    const int gx = rx0 + cx;
                    const bool slot = lane < pairs;
                    const bool live_x = slot && cx >= 0 && cx < rw && gx >= 0 && gx < p.rect_w;
                    for (int r0 = wrp; r0 < rh; r0 += UA * NW) {
                        uint4 kk[UA];
#pragma unroll
                        for (int j = 0; j < UA; ++j) {
                            const int gy = ry0 + r0 + j * NW;
                            kk[j] = make_uint4(0u, 0u, 0u, 0u);
                            if (live_x && r0 + j * NW < rh && gy >= 0 && gy < p.rect_h)
                                kk[j] = __ldcg(reinterpret_cast<const uint4*>(p.map + gy * p.rect_w + gx));
                        }
#pragma unroll
                        for (int j = 0; j < UA; ++j) {
                            const int row = r0 + j * NW;
                            // key = epoch:16 | index:32 | disparity:16 -> the epoch is the top half of the high word
                            const unsigned d0 = (kk[j].y >> 16) == ep16 ? (kk[j].x & 0xffffu) : 0u;
                            const unsigned d1 = (kk[j].w >> 16) == ep16 ? (kk[j].z << 16) : 0u;
                            if (slot && row < rh) dst[row * pairs + lane] = d0 | d1;
                        }
                    }
                } else {
                    // UA cell pairs per thread per round, all loads of a round issued before the first decode
                    const unsigned mg = magic_div(pairs);
                    const int total = pairs * rh;
                    for (int c0 = tid; c0 < total; c0 += UA * NT) {
                        ulonglong2 kk[UA];
                        bool live[UA];
#pragma unroll
                        for (int j = 0; j < UA; ++j) {
                            const int c = c0 + j * NT;
                            const int ry = static_cast<int>(__umulhi(static_cast<unsigned>(c), mg));
                            const int cx = 2 * (c - ry * pairs) - kRowPad;  // data coordinate of the pair's first cell (even)
                            const int gx = rx0 + cx, gy = ry0 + ry;
                            live[j] = c < total && cx >= 0 && cx < rw && gx >= 0 && gx < p.rect_w && gy >= 0 && gy < p.rect_h;
                            kk[j] = make_ulonglong2(0ULL, 0ULL);
                            if (live[j]) kk[j] = __ldcg(reinterpret_cast<const ulonglong2*>(p.map + gy * p.rect_w + gx));
                        }
#pragma unroll
                        for (int j = 0; j < UA; ++j) {
                            const int c = c0 + j * NT;
                            if (c < total)
                                dst[c] = live[j] ? (static_cast<unsigned>(key_disparity(kk[j].x, epoch)) |
                                                    (static_cast<unsigned>(key_disparity(kk[j].y, epoch)) << 16))
                                                 : 0u;
                        }
                    }
                }
            }
            group_sync<NT>(bar_id);
            // ---- B: horizontal 7-max, 8 outputs per task ------------------------------------------
            {
                const int segs = (rw + 7) >> 3;
                const unsigned mg = magic_div(segs);
                for (int t = tid; t < segs * rh; t += NT) {
                    const int ry = static_cast<int>(__umulhi(static_cast<unsigned>(t), mg));
                    const int seg = t - ry * segs;
                    const unsigned* src = reinterpret_cast<const unsigned*>(bufA + ry * stride) + 4 * seg;
                    unsigned v[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const unsigned w = src[i];
                        v[2 * i] = w & 0xffffu;
                        v[2 * i + 1] = w >> 16;
                    }
                    // v[i] = padded cell 8*seg + i = data cell 8*seg + i - 4; out[j] = max(v[j+1 .. j+7])
                    unsigned m2[15], m4[13];
#pragma unroll
                    for (int i = 1; i < 15; ++i) m2[i] = max(v[i], v[i + 1]);
#pragma unroll
                    for (int i = 1; i < 13; ++i) m4[i] = max(m2[i], m2[i + 2]);
                    unsigned* dst = reinterpret_cast<unsigned*>(bufB + ry * stride) + 4 * seg + 2;
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        const unsigned a = max(m4[j + 1], m4[j + 4]);
                        const unsigned b = max(m4[j + 2], m4[j + 5]);
                        dst[j >> 1] = a | (b << 16);
                    }
                }
            }
            group_sync<NT>(bar_id);
            // ---- C: vertical 7-max on column pairs, 8 output rows per task -----------------------------
            {
                const int orows = rh - 2 * R;  // rows a pixel of this tile can map to: [R, rh - R)
                const int vsegs = (orows + 7) >> 3;
                const int cpairs = rw >> 1;
                const unsigned mg = magic_div(cpairs);
                for (int t = tid; t < vsegs * cpairs; t += NT) {
                    const int vs = static_cast<int>(__umulhi(static_cast<unsigned>(t), mg));
                    const int q = t - vs * cpairs;
                    const int word = (kRowPad >> 1) + q;  // word index of the pair inside a row
                    const int rbase = 8 * vs;             // first input row; output rows R + 8*vs + j
                    unsigned x[14];
#pragma unroll
                    for (int i = 0; i < 14; ++i) {
                        const int ry = rbase + i;
                        x[i] = ry < rh ? reinterpret_cast<const unsigned*>(bufB + ry * stride)[word] : 0u;
                    }
                    unsigned m2[13], m4[11];
#pragma unroll
                    for (int i = 0; i < 13; ++i) m2[i] = __vmaxu2(x[i], x[i + 1]);
#pragma unroll
                    for (int i = 0; i < 11; ++i) m4[i] = __vmaxu2(m2[i], m2[i + 2]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int ry = R + rbase + j;
                        if (ry < rh - R) reinterpret_cast<unsigned*>(bufA + ry * stride)[word] = __vmaxu2(m4[j], m4[j + 3]);
                    }
                }
            }
            group_sync<NT>(bar_id);
        }
    }
}

// 49-tap fallback for one pixel (regions that do not fit the buffers)
__device__ __forceinline__ int proj7_direct(const EpilogueParams& p, const short2 m, unsigned epoch) {
    constexpr int R = 3;
    int best = 0;
    for (int dy = -R; dy <= R; ++dy) {
        const int gy = m.y + dy;
        if (gy < 0 || gy >= p.rect_h) continue;
        for (int dx = -R; dx <= R; ++dx) {
            const int gx = m.x + dx;
            if (gx < 0 || gx >= p.rect_w) continue;
            best = max(best, key_disparity(p.map[gy * p.rect_w + gx], epoch));
        }
    }
    return best;
}

template <int NT, int UA = 1>
__device__ __forceinline__ void proj7_tile(const EpilogueParams& p, int bx, int by, int tiles_x, unsigned epoch, unsigned short* bufA,
                                           unsigned short* bufB, int gtid, int bar_id) {
    constexpr int PX = kTile * kTile / NT;  // output pixels per thread
    constexpr int ROWS = NT / 32;           // tile rows covered by one pass of the group
    const int lane = gtid & 31, warp = gtid >> 5;
    const int u0 = bx * kTile, v0 = by * kTile;
    const TileRegion g = tile_region(p, __ldg(p.tile_box + by * tiles_x + bx));

    // the remap coordinates are requested before the region work and used after it
    short2 m[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        const int u = u0 + lane, v = v0 + warp + k * ROWS;
        m[k] = make_short2(-1, -1);
        if (u < p.out_w && v < p.out_h) m[k] = __ldg(p.remap_xy + v * p.out_w + u);
    }
    if (g.fits) proj7_dilate_region<NT, UA>(p, g, epoch, bufA, bufB, gtid, bar_id);
    int val[PX];
    int idx[PX];
    bool live[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        val[k] = 0;
        const bool inside = g.any && static_cast<unsigned>(m[k].x) < static_cast<unsigned>(p.rect_w) &&
                            static_cast<unsigned>(m[k].y) < static_cast<unsigned>(p.rect_h);  // negative coordinates wrap to large values
        if (inside && g.fits) val[k] = bufA[(m[k].y - g.ry0) * g.stride + (m[k].x - g.rx0) + kRowPad];
        if (inside && !g.fits) val[k] = proj7_direct(p, m[k], epoch);
        const int u = u0 + lane, v = v0 + warp + k * ROWS;
        live[k] = u < p.out_w && v < p.out_h;
        idx[k] = v * p.out_w + u;
    }
    emit_pixels_int<PX>(p.out, p.dst, idx, val, live);
}

// Same tile for a small group inside a register-tight kernel (batch_kernel's epilogue warps): the remap
// coordinates are loaded AFTER the region work, PXB pixels at a time, so nothing but the region geometry
// is live across phases A-C and the region loads can be issued UA at a time without spilling.
template <int NT, int UA, int PXB>
__device__ __forceinline__ void proj7_tile_late(const EpilogueParams& p, int bx, int by, int tiles_x, unsigned epoch, unsigned short* bufA,
                                                unsigned short* bufB, int gtid, int bar_id) {
    constexpr int PX = kTile * kTile / NT;
    constexpr int ROWS = NT / 32;
    static_assert(PX % PXB == 0, "pixel batch must divide the pixels per thread");
    const int lane = gtid & 31, warp = gtid >> 5;
    const int u0 = bx * kTile, v0 = by * kTile;
    const TileRegion g = tile_region(p, __ldg(p.tile_box + by * tiles_x + bx));
    if (g.fits) proj7_dilate_region<NT, UA>(p, g, epoch, bufA, bufB, gtid, bar_id);
    const int u = u0 + lane;
    if (u >= p.out_w) return;  // (after the last group barrier)
    const bool full = v0 + kTile <= p.out_h;  // every row of the tile exists: no per-pixel row test
    const short2* rp = p.remap_xy + static_cast<long long>(v0 + warp) * p.out_w + u;
    const int pix0 = (v0 + warp) * p.out_w + u;  // (frames are limited to 2^30 pixels at context creation)
    const int step = ROWS * p.out_w;
    if (p.tile_off != nullptr && g.fits) {
        // the pixel's cell inside the region was worked out when the remap table was uploaded: one 16-bit load, one
        // shared-memory read, one table look-up and one store per pixel
        const unsigned short* op = p.tile_off + pix0;
        for (int k0 = 0; k0 < PX; k0 += PXB) {
            unsigned off[PXB];
            bool live[PXB];
            int val[PXB], idx[PXB];
#pragma unroll
            for (int j = 0; j < PXB; ++j) {
                live[j] = full || v0 + warp + (k0 + j) * ROWS < p.out_h;
                off[j] = 0xffffu;
                if (live[j]) off[j] = __ldg(op + (k0 + j) * step);
            }
#pragma unroll
            for (int j = 0; j < PXB; ++j) {
                val[j] = off[j] != 0xffffu ? bufA[off[j]] : 0;
                idx[j] = pix0 + (k0 + j) * step;
            }
            emit_pixels_int<PXB>(p.out, p.dst, idx, val, live);
        }
        return;
    }
    for (int k0 = 0; k0 < PX; k0 += PXB) {
        short2 m[PXB];
        bool live[PXB];
#pragma unroll
        for (int j = 0; j < PXB; ++j) {
            live[j] = full || v0 + warp + (k0 + j) * ROWS < p.out_h;
            m[j] = make_short2(-1, -1);
            if (live[j]) m[j] = __ldg(rp + (k0 + j) * step);
        }
        int val[PXB];
        int idx[PXB];
#pragma unroll
        for (int j = 0; j < PXB; ++j) {
            val[j] = 0;
            // unsigned compares: negative coordinates wrap to large values
            const bool inside = g.any && static_cast<unsigned>(m[j].x) < static_cast<unsigned>(p.rect_w) &&
                                static_cast<unsigned>(m[j].y) < static_cast<unsigned>(p.rect_h);
            if (inside && g.fits) val[j] = bufA[(m[j].y - g.ry0) * g.stride + (m[j].x - g.rx0) + kRowPad];
            if (inside && !g.fits) val[j] = proj7_direct(p, m[j], epoch);
            idx[j] = pix0 + (k0 + j) * step;
        }
        emit_pixels_int<PXB>(p.out, p.dst, idx, val, live);
    }
}

// ---------------------------------------------------------------------------------------------
// Strip epilogue (projector view, 7x7 dilation, even rect_w): the same result in two passes WITHOUT shared memory or
// barriers, each item done by ONE warp.
//   pass 1 (rectified domain, every cell decoded once): a warp walks down kStripRows + 6 rows of a strip of 64 cells
//     (lane l holds cells 2l, 2l+1 as a u16 pair; lanes 0, 1, 30, 31 are the halo), decodes the keys, takes the
//     horizontal 7-max with four shuffles and the vertical 7-max with a register sliding window (three SIMD max per
//     row), and stores the dilated u16 disparities of its 56 x kStripRows cells into a small map (rect_w x rect_h x 2 B);
//   pass 2 (output domain): per output pixel one streamed 32-bit load of the pixel's cell index (table built with the
//     remap table), one 16-bit gather from the dilated map (L2), one depth-table look-up, one store.
// cv2.dilate ignores taps outside the image and disparities are >= 0, so "outside" and "undefined" are both 0.
// ---------------------------------------------------------------------------------------------
constexpr int kStripCols = 56;   // cells a strip produces per row
#ifndef XM_STRIP_ROWS
#define XM_STRIP_ROWS 42
#endif
#ifndef XM_STRIP_BATCH
#define XM_STRIP_BATCH 8
#endif
#ifndef XM_REMAP_PX
#define XM_REMAP_PX 8
#endif
#ifndef XM_REMAP_BLOCKS
#define XM_REMAP_BLOCKS 8
#endif
// Item sizes are chosen per launch (StripWindow::rows / ::blocks): small items for small frames, whose epilogue is a
// latency chain (more items in flight), large ones for large frames (fewer tickets, waits and fences).
constexpr int kStripRows = XM_STRIP_ROWS;    // default rows a pass-1 item produces (it reads rows + 6)
constexpr int kStripBatch = XM_STRIP_BATCH;  // rows whose key loads are in flight together: an item costs (rows + 6) / kStripBatch L2 round trips
constexpr int kRemapPx = XM_REMAP_PX;        // output pixels per lane and block of a pass-2 item
constexpr int kRemapBlocks = XM_REMAP_BLOCKS;  // default blocks per pass-2 item (software-pipelined: cell indices two blocks ahead, gathers one)
constexpr int kRemapBlockPx = 32 * kRemapPx;   // output pixels of a block

// The window of the rectified image pass 1 has to produce: the bounding box of the remap targets (x0 rounded down to even).
struct StripWindow {
    int x0, y0, x1, y1;  // cells [x0, x1) x [y0, y1); x0 even
    int rows, blocks;    // rows per pass-1 item, blocks per pass-2 item
    int strips, items;   // strips across, pass-1 items (strips x row segments)
};
inline StripWindow strip_window(int bx0, int by0, int bx1, int by1, int rows, int blocks) {  // inclusive bounding box
    StripWindow w;
    w.x0 = bx0 & ~1;
    w.y0 = by0;
    w.x1 = bx1 + 1;
    w.y1 = by1 + 1;
    w.rows = rows;
    w.blocks = blocks;
    w.strips = (w.x1 - w.x0 + kStripCols - 1) / kStripCols;
    w.items = w.strips * ((w.y1 - w.y0 + rows - 1) / rows);
    return w;
}

__device__ __forceinline__ unsigned ldg_stream_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// pass 1, one item: output cells [x_out, x_out + 56) x [y_out, y_end); x_out even
__device__ __forceinline__ void strip_dilate_item(const unsigned long long* __restrict__ map, unsigned short* __restrict__ dil, int rect_w,
                                                  int rect_h, unsigned epoch, int x_out, int y_out, int y_end, int lane) {
    constexpr unsigned kFull = 0xffffffffu;
    const unsigned ep16 = epoch & 0xffffu;
    const int gx = x_out - 4 + 2 * lane;  // first cell of this lane's pair (even; rect_w is even)
    const bool col_in = gx >= 0 && gx < rect_w;
    const bool store_lane = lane >= 2 && lane < 30 && col_in;
    const int y_first = y_out - 3;  // first row read
    unsigned h1 = 0u, m2a = 0u, m2b = 0u, m4a = 0u, m4b = 0u, m4c = 0u;  // sliding window: rows i-1 (h, m2, m4), i-2 (m2, m4), i-3 (m4)
#pragma unroll 1
    for (int i0 = 0; y_first + i0 - 3 < y_end; i0 += kStripBatch) {  // (the batch's last row is the window end of output row y_first + i0 + kStripBatch - 4)
        uint4 cur[kStripBatch];
#pragma unroll
        for (int j = 0; j < kStripBatch; ++j) {
            const int gy = y_first + i0 + j;
            cur[j] = make_uint4(0u, 0u, 0u, 0u);
            if (col_in && gy >= 0 && gy < rect_h) cur[j] = __ldcg(reinterpret_cast<const uint4*>(map + static_cast<long long>(gy) * rect_w + gx));
        }
#pragma unroll
        for (int j = 0; j < kStripBatch; ++j) {
            // key = epoch:16 | index:32 | disparity:16 -> the epoch is the top half of the high word
            const unsigned d0 = (cur[j].y >> 16) == ep16 ? (cur[j].x & 0xffffu) : 0u;
            const unsigned d1 = (cur[j].w >> 16) == ep16 ? (cur[j].z & 0xffffu) : 0u;
            const unsigned pr = d0 | (d1 << 16);
            // horizontal: cells 2l-3 .. 2l+3 = hi(l-2), pair(l-1), pair(l), pair(l+1); cells 2l-2 .. 2l+4 = pair(l-1 .. l+1), lo(l+2)
            const unsigned m = max(d0, d1);
            const unsigned common = max(max(__shfl_up_sync(kFull, m, 1), m), __shfl_down_sync(kFull, m, 1));
            const unsigned ha = max(common, __shfl_up_sync(kFull, pr, 2) >> 16);
            const unsigned hb = max(common, __shfl_down_sync(kFull, pr, 2) & 0xffffu);
            const unsigned h = ha | (hb << 16);
            // vertical: max of rows i-6 .. i  (m2_i = rows i-1..i, m4_i = rows i-3..i)
            const unsigned m2 = __vmaxu2(h, h1);
            const unsigned m4 = __vmaxu2(m2, m2b);
            const unsigned out = __vmaxu2(m4, m4c);
            h1 = h;
            m2b = m2a;
            m2a = m2;
            m4c = m4b;
            m4b = m4a;
            m4a = m4;
            const int gy = y_first + i0 + j - 3;  // the row whose window ends with row i
            if (store_lane && i0 + j >= 6 && gy < y_end)
                *reinterpret_cast<unsigned*>(dil + static_cast<long long>(gy) * rect_w + gx) = out;
        }
    }
}

// pass 2, one item = output pixels [first, first + blocks * kRemapBlockPx) in blocks of 32 x kRemapPx,
// as a three-stage software pipeline: cell indices of block b + 2 (static table, streamed), gathers from the dilated
// map for block b + 1, depth-table look-ups and stores of block b.  strip_remap_begin() only touches the static table,
// so it can run before the frame's dilated map is complete.
struct RemapPipe {
    unsigned c0[kRemapPx], c1[kRemapPx];  // cell indices of the next two blocks
};
__device__ __forceinline__ void strip_remap_cells(unsigned (&cell)[kRemapPx], const unsigned* __restrict__ pix_cell, int n_px, int first) {
#pragma unroll
    for (int j = 0; j < kRemapPx; ++j) {
        cell[j] = 0xffffffffu;
        if (first + j * 32 < n_px) cell[j] = ldg_stream_u32(pix_cell + first + j * 32);
    }
}
__device__ __forceinline__ void strip_remap_begin(RemapPipe& rp, const unsigned* __restrict__ pix_cell, int n_px, int first, int lane) {
    const int base = first + lane;
    strip_remap_cells(rp.c0, pix_cell, n_px, base);
    strip_remap_cells(rp.c1, pix_cell, n_px, base + kRemapBlockPx);
}
__device__ __forceinline__ void strip_remap_run(RemapPipe& rp, const OutputSpec& o, void* dst, const unsigned* __restrict__ pix_cell,
                                                const unsigned short* __restrict__ dil, int n_px, int first, int blocks, int lane) {
    const int base = first + lane;
    int val[kRemapPx];
#pragma unroll
    for (int j = 0; j < kRemapPx; ++j) val[j] = rp.c0[j] != 0xffffffffu ? static_cast<int>(__ldcg(dil + rp.c0[j])) : 0;
#pragma unroll 1
    for (int b = 0; b < blocks; ++b) {
        // gathers of block b + 1, cell indices of block b + 3 (c0 is free once its gathers are issued)
        int nxt[kRemapPx];
#pragma unroll
        for (int j = 0; j < kRemapPx; ++j) nxt[j] = (b + 1 < blocks && rp.c1[j] != 0xffffffffu) ? static_cast<int>(__ldcg(dil + rp.c1[j])) : 0;
#pragma unroll
        for (int j = 0; j < kRemapPx; ++j) rp.c0[j] = rp.c1[j];
        if (b + 2 < blocks)
            strip_remap_cells(rp.c1, pix_cell, n_px, base + (b + 2) * kRemapBlockPx);
        else {
#pragma unroll
            for (int j = 0; j < kRemapPx; ++j) rp.c1[j] = 0xffffffffu;
        }
        int idx[kRemapPx];
        bool live[kRemapPx];
#pragma unroll
        for (int j = 0; j < kRemapPx; ++j) {
            idx[j] = base + b * kRemapBlockPx + j * 32;
            live[j] = idx[j] < n_px;
        }
        emit_pixels_int<kRemapPx>(o, dst, idx, val, live);
#pragma unroll
        for (int j = 0; j < kRemapPx; ++j) val[j] = nxt[j];
    }
}

__global__ void __launch_bounds__(256, 7) epilogue_projector7_kernel(const EpilogueParams p) {
    extern __shared__ __align__(128) unsigned char ev_smem[];
    unsigned short* bufA = reinterpret_cast<unsigned short*>(ev_smem);
    unsigned short* bufB = bufA + p.region_cap;
    if (p.use_pdl) pdl_launch_dependents();
    const int tid = threadIdx.x;
    if (p.out.kind == 0 && p.out.depth_lut && tid < 128) prefetch_l1(p.out.depth_lut + tid * 32);
    // the key map and the state block belong to K1
    if (p.use_pdl) pdl_wait();
    const unsigned epoch = p.epoch + p.state->redo;
    if (p.recycle && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) recycle_state(p.recycle);
    proj7_tile<256>(p, blockIdx.x, blockIdx.y, gridDim.x, epoch, bufA, bufB, tid, 0);
}

}  // namespace xm
