// xm_fused_kernel.cuh — ONE kernel per frame.
//
// frame_kernel<CAM> runs, in one persistent grid of co-resident CTAs (148 x occupancy):
//
//   phase 1  the lean per-event pipeline (lean_events_phase: TMA event ring, X-map window ring,
//            cp.async LUT gathers, integer time columns, 64-bit RED scatter), chunks handed out by a
//            global counter;
//   barrier  a grid-wide barrier (arrival counter in the frame's state block).  Every CTA has pushed
//            its statistics and status flags before arriving, so afterwards all CTAs agree on whether
//            an event violated the assumed time bounds;
//   fix-up   (only then) exact bounds by a grid-wide reduction, a second barrier, and a plain second
//            pass over the events with a fresh epoch -- slow but rare, and exact for any input;
//   phase 2  the per-pixel epilogue (dilate + remap + depth / disparity / BGR), 32x32 output tiles
//            handed out by a second counter (camera view: a grid-stride loop).
//
// Compared with separate kernels this removes the K1 -> K2 kernel boundary (~8 us of drain + launch on a
// B200), needs no device-side launch for the fix-up, and is therefore compatible with programmatic
// dependent launch: the next frame's prologue (barrier init, bounds look-up) overlaps this frame's tail.
#pragma once
#include "xm_frame_kernels.cuh"

namespace xm {

constexpr int kTileGroup = 128;  // threads that share one epilogue tile in the fused kernel

struct FrameParams {
    EventParams ev;
    EpilogueParams ep;
    int tiles_x, tiles_y;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Spin until *p >= need (a counter that only grows), then acquire.  The polls are RELAXED loads: an acquire load is
// an LDG followed by CCTL.IVALL (invalidate the SM's whole L1), and the batch kernel's epilogue warps polled 350 000
// times per frame -- an L1 flush every 12 ns per SM (ncu r3r).  One acquire load after the last poll gives the same
// ordering (a fence instead would also wait for the warp's own REDs and gathers in flight: measured slower).
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void spin_until_ge(const unsigned* p, unsigned need, unsigned ns0, unsigned ns1) {
    if (ld_acquire_u32(p) >= need) return;  // (already there: one round trip)
    for (unsigned ns = ns0; ld_relaxed_u32(p) < need; ns = min(ns * 2u, ns1)) __nanosleep(ns);
    (void)ld_acquire_u32(p);
}

// All CTAs of the grid must be resident (the host sizes the grid with the occupancy API).
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        spin_until_ge(counter, target, 40u, 40u);
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ void flush_stats(FrameState* st, unsigned n_valid, unsigned n_inl, unsigned flags) {
    n_valid = __reduce_add_sync(0xffffffffu, n_valid);
    n_inl = __reduce_add_sync(0xffffffffu, n_inl);
    flags = __reduce_or_sync(0xffffffffu, flags);
    if ((threadIdx.x & 31) == 0) {
        if (n_valid) atomicAdd(&st->n_valid, static_cast<unsigned long long>(n_valid));
        if (n_inl) atomicAdd(&st->n_inliers, static_cast<unsigned long long>(n_inl));
        if (flags) atomicOr(&st->flags, flags);
    }
}

template <bool CAM>
__global__ void __launch_bounds__(kWsThreads, 3) frame_kernel(const FrameParams fp) {
    extern __shared__ __align__(128) unsigned char ev_smem[];
    __shared__ int s_work;
    const EventParams& p = fp.ev;
    const EpilogueParams& q = fp.ep;
    FrameState* st = p.state;
    const int tid = threadIdx.x;

    // ---- phase 1 ----------------------------------------------------------------------------------
    unsigned n_valid = 0, n_inl = 0, flags = 0;
    lean_events_phase<CAM>(p, ev_smem, n_valid, n_inl, flags);
    if (p.dbg && tid == 0) p.dbg[blockIdx.x * 8 + 5] = global_timer_ns();
    flush_stats(st, n_valid, n_inl, flags);
    if (q.recycle && blockIdx.x == 0 && tid == 0) recycle_state(q.recycle);  // state block of the frame after next
    if (q.out.kind == 0 && q.out.depth_lut && tid < 128) prefetch_l1(q.out.depth_lut + tid * 32);

    grid_barrier(&st->blocks_done, gridDim.x);
    if (p.dbg && tid == 0) p.dbg[blockIdx.x * 8 + 6] = global_timer_ns();
    unsigned epoch = p.epoch;

    // ---- fix-up: an event lay outside the assumed [t_min, t_max] ------------------------------------
    if (p.arm_fixup && (ld_acquire_u32(&st->flags) & kStatusTBounds)) {
        // (a) exact bounds; CTA 0 also clears the counters of the first pass (all CTAs flushed before the barrier)
        if (blockIdx.x == 0 && tid == 0) {
            st->n_valid = 0;
            st->n_inliers = 0;
            st->redo = 1;
            atomicAnd(&st->flags, ~(kStatusPixelOob | kStatusScatterOob));
        }
        unsigned long long lo = 0xffffffffffffffffULL, hi = 0;
        bool any = false;
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + tid; i < p.n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
            const EventFields e = unpack_event(ld_event_plain(p.events + i));
            if (event_valid(e, p.polarity)) {
                const unsigned long long u = time_to_ordered<false>(e.t_bits);
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
        if ((tid & 31) == 0 && any) {
            atomicMax(&st->red_lo, ~lo);
            atomicMax(&st->red_hi, hi + 1ULL);
        }
        grid_barrier(&st->blocks_done, 2u * gridDim.x);
        const unsigned long long rl = *reinterpret_cast<volatile unsigned long long*>(&st->red_lo);
        const unsigned long long rh = *reinterpret_cast<volatile unsigned long long*>(&st->red_hi);
        const bool found = rh != 0ULL;
        const long long t_lo = found ? ordered_to_time<false>(~rl) : 0;
        const long long t_hi = found ? ordered_to_time<false>(rh - 1ULL) : 0;
        if (blockIdx.x == 0 && tid == 0) {
            st->t_lo_bits = t_lo;
            st->t_hi_bits = t_hi;
        }
        // (b) plain second pass with a fresh epoch (chunks from a second counter)
        epoch = p.epoch + 1;
        TimeCol<false> tc;
        tc.init(t_lo, t_hi, p.t_px_scale);
        const int total_chunks = static_cast<int>((p.n + kEvChunk - 1) / kEvChunk);
        n_valid = n_inl = flags = 0;
        for (;;) {
            __syncthreads();
            if (tid == 0) s_work = static_cast<int>(atomicAdd(&st->fix_chunk, 1u));
            __syncthreads();
            const int g = s_work;
            if (g >= total_chunks) break;
            if (tid >= kEvThreads) continue;
#pragma unroll
            for (int k = 0; k < kEvPerThread; ++k) {
                const long long i = static_cast<long long>(g) * kEvChunk + k * kEvThreads + tid;
                if (i >= p.n) continue;
                const EventFields e = unpack_event(ld_event_plain(p.events + i));
                if (!event_valid(e, p.polarity)) continue;
                ++n_valid;
                if (e.x >= static_cast<unsigned>(p.cam_w) || e.y >= static_cast<unsigned>(p.cam_h)) {
                    flags |= kStatusPixelOob;
                    continue;
                }
                const int pix = static_cast<int>(e.y) * p.cam_w + static_cast<int>(e.x);
                const int lut = __ldg(p.lut_xy + pix);
                bool viol;
                int cc = tc.column(e.t_bits, viol);
                if (cc < 0) cc += p.xmap_w;
                if (viol || cc < 0 || cc >= p.xmap_w) {  // cannot happen with exact bounds
                    flags |= kStatusTBounds;
                    continue;
                }
                const int xcr = static_cast<short>(lut & 0xffff);
                const int ycr = lut >> 16;
                if (static_cast<unsigned>(ycr) >= static_cast<unsigned>(p.xmap_h - 1)) continue;
                const int xp = __ldg(p.xmap_t + static_cast<long long>(cc) * p.col_stride + ycr);
                const int disp = static_cast<short>(xp - xcr - p.x_offset);
                if (disp < 0) continue;
                ++n_inl;
                const int cell = CAM ? pix : ycr * p.rect_w + (xp - p.x_offset);  // verified tables: inside the map
                atomicMax(p.map + cell, make_key32(epoch, static_cast<unsigned>(i), disp));
            }
        }
        flush_stats(st, n_valid, n_inl, flags);
        grid_barrier(&st->blocks_done, 3u * gridDim.x);
    }

    // ---- phase 2: epilogue --------------------------------------------------------------------------
    if (CAM) {
        const int n_px = q.out_w * q.out_h;
        for (int i = blockIdx.x * blockDim.x + tid; i < n_px; i += gridDim.x * blockDim.x)
            emit_pixel_int(q.out, q.dst, i, key_disparity(q.map[i], epoch));
    } else if (tid < kEvThreads) {
        // the consumer threads split into groups of kTileGroup threads; every group takes whole tiles from
        // the global counter and works through them independently (named barriers), so that several
        // tiles per CTA are in flight and the L2 latency of one group's region loads hides behind the others
        constexpr int NT = kTileGroup;
        const int grp = tid / NT, gtid = tid % NT;
        unsigned short* bufA = reinterpret_cast<unsigned short*>(ev_smem + kEvSmemHeader) + grp * (2 * q.region_cap);
        unsigned short* bufB = bufA + q.region_cap;
        int* s_tile = reinterpret_cast<int*>(ev_smem + 208);  // [kEvThreads / NT] (unused header words)
        const int n_tiles = fp.tiles_x * fp.tiles_y;
        for (;;) {
            group_sync<NT>(1 + grp);  // the previous tile's gather is done with the buffers
            if (gtid == 0) s_tile[grp] = static_cast<int>(atomicAdd(&st->next_tile, 1u));
            group_sync<NT>(1 + grp);
            const int t = s_tile[grp];
            if (t >= n_tiles) break;
            const int by = t / fp.tiles_x;
            proj7_tile<NT>(q, t - by * fp.tiles_x, by, fp.tiles_x, epoch, bufA, bufB, gtid, 1 + grp);
        }
    }
    if (p.dbg && (tid == 0 || tid == kEvThreads - 1)) {
        atomicMax(p.dbg + blockIdx.x * 8 + 4, global_timer_ns());
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.dbg[blockIdx.x * 8 + 7] = smid;
    }
}

}  // namespace xm
