// xm_batch2_kernel.cuh — the batch kernel with occupancy-driven event warps.
//
// Same batch structure as batch_kernel (ordered chunk list, per-frame completion counts, epilogue warp groups,
// three scatter maps in rotation, bounds / redo kernels), but the event side is deliberately simple: no shared
// memory ring, no TMA, no software pipeline.  Every event warp takes its OWN chunks of 128 events (4 per lane,
// four independent 128-bit streaming loads) from the global counter, gathers the packed LUT words and the X-map
// cells with plain read-only loads and scatters; the latency is hidden by the number of warps (12 event warps x
// 2 CTAs per SM at <= 64 registers) instead of by staging.  tools/gather_probe.cu measured that form at 34-35 us
// per 5 M events against 39-41 us for the staged pipeline: the per-event loop is bound by the L1 -> L2 sector
// rate (LUT gathers 18.6 us, stream 15 us, atomics 10 us on their own), which staging cannot improve, while it
// costs issue slots, registers (fewer warps) and 150 KB of shared memory that is worth more as L1 (the X-map
// columns of the moment and ~15 % of the LUT gathers then hit L1).
#pragma once
#include "xm_batch_kernel.cuh"

namespace xm {

constexpr int kB2EventWarps = 12;
constexpr int kB2TileWarps = kTileWarps;
constexpr int kB2Threads = (kB2EventWarps + kB2TileWarps) * 32;
constexpr int kB2Sub = 32 * kEvPerThread;    // events per inner step of a warp (4 per lane)
constexpr int kB2Steps = 8;                  // inner steps per chunk: one atomic hands out 1024 events (39 k atomics per
                                             // 5 M-event frame on ONE counter cost more than the events themselves)
constexpr int kB2Chunk = kB2Sub * kB2Steps;  // events per chunk
constexpr int kB2Header = 1024;              // [kBatchMax][5] CTA accumulators (640 B; the event warps of a CTA are
                                             // not coupled, so no two frames may share a slot), tile tickets at 960

inline int batch2_smem_bytes(int region_cells) { return kB2Header + kTileGroups * region_cells * 4; }
__host__ __device__ __forceinline__ unsigned batch2_chunks(long long n) { return static_cast<unsigned>((n + kB2Chunk - 1) / kB2Chunk); }

// exact column of a timestamp the integer fast path could not decide (outside the assumed bounds, exact tie):
// the reference's own float64 expression.  Returns -3 for a bounds violation.
static __device__ __noinline__ int batch2_slow_column(long long t_bits, const FrameState* st, int t_px_scale, int xmap_w) {
    TimeCol<false> tc;
    tc.init(__ldcg(&st->t_lo_bits), __ldcg(&st->t_hi_bits), t_px_scale);
    bool viol;
    int cc = tc.column(t_bits, viol);
    if (cc < 0) cc += xmap_w;  // NumPy negative index (only reachable with wrong bounds)
    viol = viol || cc < 0 || cc >= xmap_w;
    return viol ? -3 : cc;
}

template <bool CAM>
__global__ void __launch_bounds__(kB2Threads, 2) batch2_kernel(const __grid_constant__ BatchParams bp) {
    extern __shared__ __align__(128) unsigned char ev_smem[];
    unsigned* s_acc = reinterpret_cast<unsigned*>(ev_smem);  // [kBatchMax][5]: valid, inliers, flags, warps arrived, chunks
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int B = bp.n_frames;
    if (tid < kBatchMax * 5) s_acc[tid] = 0u;
    __syncthreads();

    if (warp >= kB2EventWarps) {
        const int t = tid - kB2EventWarps * 32;
        const int grp = t / kTileGroupThreads, gtid = t % kTileGroupThreads;
        unsigned short* bufA = reinterpret_cast<unsigned short*>(ev_smem + kB2Header) + grp * (2 * bp.ep.region_cap);
        batch_tile_groups<CAM, kB2Chunk>(bp, grp, gtid, bufA, bufA + bp.ep.region_cap, reinterpret_cast<volatile int*>(ev_smem + 960) + grp);
        return;
    }

    // ---- event warps ------------------------------------------------------------------------------------
    unsigned* const counter = &bp.states[B].next_chunk;
    int slot = 0;
    unsigned s_first = bp.first_item[0], s_next = bp.first_item[1];

    int cur_f = -1;
    unsigned my_chunks = 0, n_valid = 0, n_inl = 0, flags = 0;
    IntCol ic;
    ic.init(0, 0, bp.t_px_scale);
    unsigned epoch16 = 0, n_f = 0;
    unsigned long long* map = bp.maps[0];
    const int4* ev = nullptr;
    const unsigned pol_mask = bp.polarity ? 0xffffu : 0u;

    // Every warp passes through every frame in order (frames it got no chunk of included), so each frame
    // collects exactly kB2EventWarps arrivals per CTA and the last warp forwards the CTA's totals.
    auto leave = [&](int f) {
        n_valid = __reduce_add_sync(0xffffffffu, n_valid);
        n_inl = __reduce_add_sync(0xffffffffu, n_inl);
        flags = __reduce_or_sync(0xffffffffu, flags);
        if (my_chunks) fence_acq_rel_gpu();  // every lane: its scatter atomics are ordered before the counts below
        __syncwarp();
        if (lane == 0) {
            unsigned* acc = s_acc + f * 5;
            if (n_valid) atomicAdd(acc + 0, n_valid);
            if (n_inl) atomicAdd(acc + 1, n_inl);
            if (flags) atomicOr(acc + 2, flags);
            if (my_chunks) atomicAdd(acc + 4, my_chunks);
            __threadfence_block();
            if (atomicAdd(acc + 3, 1u) == kB2EventWarps - 1) {
                __threadfence_block();
                const unsigned v = atomicExch(acc + 0, 0u), i = atomicExch(acc + 1, 0u), fl = atomicExch(acc + 2, 0u);
                const unsigned ch = atomicExch(acc + 4, 0u);
                atomicExch(acc + 3, 0u);
                FrameState* st = bp.states + f;
                if (v) atomicAdd(&st->n_valid, static_cast<unsigned long long>(v));
                if (i) atomicAdd(&st->n_inliers, static_cast<unsigned long long>(i));
                if (fl) atomicOr(&st->flags, fl);
                if (ch) {
                    fence_acq_rel_gpu();
                    atomicAdd(&st->blocks_done, ch);
                }
            }
        }
        __syncwarp();
        n_valid = n_inl = flags = 0;
        my_chunks = 0;
    };
    auto advance_to = [&](int f) {  // f == B: end of the batch
        while (cur_f < f) {
            if (cur_f >= 0) leave(cur_f);
            ++cur_f;
        }
        if (f >= B) return;
        if (f >= kBatchMaps) {
            // this frame scatters into the map frame f - kBatchMaps used: all of that frame's tiles must have read it
            if (lane == 0) {
                const unsigned* done = &bp.states[f - kBatchMaps].next_tile;
                const unsigned need = static_cast<unsigned>(bp.tile_items);
                while (ld_acquire_u32(done) < need) __nanosleep(64);
            }
            __syncwarp();
        }
        const FrameState* st = bp.states + f;
        ic.init(__ldcg(&st->t_lo_bits), __ldcg(&st->t_hi_bits), bp.t_px_scale);
        epoch16 = (bp.epoch0 + static_cast<unsigned>(f)) << 16;
        map = bp.maps[f % kBatchMaps];
        n_f = static_cast<unsigned>(bp.frames[f].n);
        ev = bp.frames[f].events;
    };

    // the chunk after the current one is requested (lane 0) before the current one is processed
    unsigned pending = 0;
    if (lane == 0) pending = atomicAdd(counter, 1u);
    for (;;) {
        const unsigned it = __shfl_sync(0xffffffffu, pending, 0);
        if (it >= bp.total_items) break;
        if (lane == 0) pending = atomicAdd(counter, 1u);
        while (it >= s_next) {
            ++slot;
            s_first = s_next;
            s_next = bp.first_item[slot + 1];
        }
        if (slot != cur_f) advance_to(slot);
        const unsigned g = it - s_first;
        for (int step = 0; step < kB2Steps; ++step) {
        const unsigned first = g * kB2Chunk + static_cast<unsigned>(step) * kB2Sub;
        if (first >= n_f) break;
        const unsigned base = first + static_cast<unsigned>(lane);
        const unsigned left = n_f - first;  // events from the start of this step to the end of the frame

        int4 raw[kEvPerThread];
#pragma unroll
        for (int k = 0; k < kEvPerThread; ++k) {
            raw[k] = make_int4(0, 0, 0, 0);
            if (left >= kB2Sub || k * 32 + lane < static_cast<int>(left)) raw[k] = __ldcs(ev + base + k * 32);
        }
        // packed LUT word of every kept event; col: time column, or -1 not kept, -2 pixel outside the image
        int lut[kEvPerThread], col[kEvPerThread];
        unsigned bad_mask = 0, kept = 0;
#pragma unroll
        for (int k = 0; k < kEvPerThread; ++k) {
            const unsigned ex = static_cast<unsigned>(raw[k].x) & 0xffffu, ey = static_cast<unsigned>(raw[k].x) >> 16;
            bool valid = ((static_cast<unsigned>(raw[k].y) ^ 1u) & pol_mask) == 0u;  // polarity: p == 1, or everything
            if (left < kB2Sub) valid = valid && (k * 32 + lane < static_cast<int>(left));
            const bool ok = valid && ex < static_cast<unsigned>(bp.cam_w) && ey < static_cast<unsigned>(bp.cam_h);
            const int px = static_cast<int>(ey * static_cast<unsigned>(bp.cam_w) + ex);
            lut[k] = 0x7fff0000;  // y_rect = 32767: fails the row test below
            if (ok) lut[k] = __ldg(bp.lut_xy + px);
            const long long t_bits = (static_cast<long long>(raw[k].w) << 32) | static_cast<unsigned>(raw[k].z);
            bool bad;
            const unsigned q = ic.column(t_bits, bad);
            col[k] = ok ? static_cast<int>(q) : (valid ? -2 : -1);
            kept += valid;
            bad_mask |= (ok && (bad || !ic.ok)) ? (1u << k) : 0u;
            if (CAM) raw[k].x = px;  // the camera-view scatter cell
        }
        n_valid += kept;
        if (bad_mask) {
#pragma unroll
            for (int k = 0; k < kEvPerThread; ++k)
                if (bad_mask & (1u << k)) {
                    const long long t_bits = (static_cast<long long>(raw[k].w) << 32) | static_cast<unsigned>(raw[k].z);
                    col[k] = batch2_slow_column(t_bits, bp.states + cur_f, bp.t_px_scale, bp.xmap_w);
                }
        }
        // X-map cells (the columns of the moment are L1-resident: all warps of the SM work on neighbouring times)
        int xp[kEvPerThread];
        unsigned live_mask = 0;
#pragma unroll
        for (int k = 0; k < kEvPerThread; ++k) {
            const unsigned ycr = static_cast<unsigned>(lut[k] >> 16);
            const bool live = col[k] >= 0 && ycr < static_cast<unsigned>(bp.xmap_h) - 1u;  // x_maps_disparity.py:23 (last row excluded)
            xp[k] = 0;
            if (live) xp[k] = __ldg(bp.xmap_t + static_cast<long long>(col[k]) * bp.col_stride + ycr);
            live_mask |= live ? (1u << k) : 0u;
            if (col[k] == -2) flags |= kStatusPixelOob;  // the reference raises IndexError here
            if (col[k] == -3) flags |= kStatusTBounds;
        }
        unsigned inl_count = 0;
#pragma unroll
        for (int k = 0; k < kEvPerThread; ++k) {
            const int xcr = static_cast<short>(lut[k] & 0xffff);
            const int ycr = lut[k] >> 16;
            const int disp = static_cast<short>(xp[k] - xcr - bp.x_offset);  // int16 arithmetic wraps
            const bool inl = ((live_mask >> k) & 1u) && disp >= 0;
            // projector view: x_rect + disp = x_map - x_offset, in [0, rect_w) for verified tables
            const int cell = CAM ? raw[k].x : ycr * bp.rect_w + (xp[k] - bp.x_offset);
            const unsigned idx = base + static_cast<unsigned>(k * 32);
            const unsigned long long key =
                (static_cast<unsigned long long>(epoch16 | (idx >> 16)) << 32) | ((idx << 16) | static_cast<unsigned>(disp));
            red_max_u64_if(map + cell, key, inl);
            inl_count += inl;
        }
        n_inl += inl_count;
        }  // step
        ++my_chunks;
    }
    advance_to(B);
}

}  // namespace xm
