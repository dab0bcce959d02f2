// xm_stream_kernels.cuh — the rows either side of the depth path:
//
//   N4  per-frame de-duplication filters       python/frame_event_filter.py:19-128
//   N2  frame segmentation (trigger finder)    python/trigger_finder.py:146-189
//
// Both are built from one order-preserving "flagged compaction" (count per 1024 elements, one-block
// scan, ordered write) instantiated with small functors, so the filters' survivors come out in the
// row-major key order of the reference's boolean-mask read-back and the pause list in stream order.
#pragma once
#include <climits>

#include "xm_stage_kernels.cuh"

namespace xm {

constexpr unsigned kStatusFilterPolarity = 0x8u;  // YT filter: an event with p != 1 (the reference's xp array would not line up)
constexpr unsigned kStatusFilterIndex = 0x10u;    // YT filter: column outside the (y, x_rect) image (reference: IndexError)

// ---------------------------------------------------------------------------------------------
// flagged compaction: pred(i) selects, emit(i, position) writes; blocks of kCompactBlock elements
// ---------------------------------------------------------------------------------------------
template <typename Pred>
__global__ void __launch_bounds__(256) flag_count_kernel(const Pred pred, long long n, unsigned* __restrict__ counts) {
    const long long base = static_cast<long long>(blockIdx.x) * kCompactBlock;
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long i = base + threadIdx.x * 4 + k;
        if (i < n && pred(i)) ++c;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ unsigned s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < 8; ++w) t += s[w];
        counts[blockIdx.x] = t;
    }
}

template <typename Pred, typename Emit>
__global__ void __launch_bounds__(256) flag_write_kernel(const Pred pred, const Emit emit, long long n, const unsigned* __restrict__ offsets) {
    const long long base = static_cast<long long>(blockIdx.x) * kCompactBlock;
    bool keep[4];
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long i = base + threadIdx.x * 4 + k;
        keep[k] = i < n && pred(i);
        c += keep[k];
    }
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    __shared__ unsigned s[8];
    if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned off = offsets[blockIdx.x] + incl - c;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) off += s[w];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (keep[k]) emit(base + threadIdx.x * 4 + k, off++);
}

// ---------------------------------------------------------------------------------------------
// A0 materialised: PolarityFilterAlgorithm(1).process_events = events[events["p"] == 1], order kept
// (python/depth_reprojection_pipe.py:43,114; restated in frame_event_filter.py:21)
// ---------------------------------------------------------------------------------------------
struct PolarityPred {
    const int4* events;
    __device__ __forceinline__ bool operator()(long long i) const { return unpack_event(ld_event_plain(events + i)).p == 1; }
};
struct CopyEventEmit {
    const int4* events;
    int4* out;
    __device__ __forceinline__ void operator()(long long i, unsigned pos) const { out[pos] = ld_event_plain(events + i); }
};

// ---------------------------------------------------------------------------------------------
// N2: ActivityNoiseFilterAlgorithm(width, height, threshold).process_events (Metavision, closed binary; call
// site python/depth_reprojection_pipe.py:65-67,116-117).  Restated from its published semantics -- an event
// passes iff one of the 8 neighbours of its pixel saw an event less than `threshold` us earlier (t_n > t -
// threshold, per-pixel timestamps start at 0 and are carried across packets; the event's own pixel is not a
// witness) -- the test oracle's ActivityNoiseFilterOracle restates the same ("parity unpinned").
//
// The definition is sequential (every event updates its pixel before the next event is looked at).  On a
// time-sorted packet that spans less than `threshold` it is equivalent to an order-free test: neighbour n is a
// witness for event i iff an event of THIS packet with a smaller index sits on n (it is younger than the
// threshold by construction) or the timestamp carried in from earlier packets is.  So: first[n] = smallest
// event index on pixel n (atomicMin), the test reads first[] and the carried image, survivors are written by
// the ordered compaction, and last[n] takes the packet's timestamps afterwards (atomicMax).  Longer packets are
// cut into sub-packets that each span less than the threshold (xm_activity_filter).
// ---------------------------------------------------------------------------------------------
struct ActivityParams {
    const int4* events;  // the sub-packet
    long long n;
    long long threshold;
    int cols, rows;
    unsigned* first;       // [rows * cols] first event index of the sub-packet per pixel, 0xffffffff = none
    long long* last;       // [rows * cols] latest timestamp per pixel carried across packets
    unsigned* unsorted;    // device flag: a timestamp smaller than its predecessor's
};

__global__ void __launch_bounds__(256) activity_clear_kernel(unsigned* first, long long cells) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < cells; i += static_cast<long long>(gridDim.x) * blockDim.x)
        first[i] = 0xffffffffu;
}

__global__ void __launch_bounds__(256) activity_mark_kernel(const ActivityParams p) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const EventFields e = unpack_event(ld_event_plain(p.events + i));
        if (i > 0 && unpack_event(ld_event_plain(p.events + i - 1)).t_bits > e.t_bits) *p.unsorted = 1u;
        if (e.x < static_cast<unsigned>(p.cols) && e.y < static_cast<unsigned>(p.rows))
            atomicMin(p.first + static_cast<long long>(e.y) * p.cols + e.x, static_cast<unsigned>(i));
    }
}

struct ActivityPred {
    ActivityParams p;
    __device__ __forceinline__ bool operator()(long long i) const {
        const EventFields e = unpack_event(ld_event_plain(p.events + i));
        if (e.x >= static_cast<unsigned>(p.cols) || e.y >= static_cast<unsigned>(p.rows)) return false;
        const long long th = e.t_bits - p.threshold;
        const int x = static_cast<int>(e.x), y = static_cast<int>(e.y);
        bool keep = false;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= p.rows) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx;
                if ((dx == 0 && dy == 0) || xx < 0 || xx >= p.cols) continue;
                const long long c = static_cast<long long>(yy) * p.cols + xx;
                keep = keep || __ldg(p.first + c) < static_cast<unsigned>(i) || __ldg(p.last + c) > th;
            }
        }
        return keep;
    }
};

// output position = *base + position inside the sub-packet
struct ActivityEmit {
    const int4* events;
    int4* out;
    const long long* base;
    __device__ __forceinline__ void operator()(long long i, unsigned pos) const { out[*base + pos] = ld_event_plain(events + i); }
};

__global__ void __launch_bounds__(256) activity_update_kernel(const ActivityParams p) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const EventFields e = unpack_event(ld_event_plain(p.events + i));
        if (e.x < static_cast<unsigned>(p.cols) && e.y < static_cast<unsigned>(p.rows))
            atomicMax(p.last + static_cast<long long>(e.y) * p.cols + e.x, e.t_bits);
    }
}

// *total += *part (one sub-packet's survivors)
__global__ void activity_advance_kernel(long long* total, const long long* part) { *total += *part; }

// Cuts a time-sorted packet into sub-packets that each span less than `threshold`: bounds[0] = 0, bounds[k + 1] =
// first index whose timestamp is >= t[bounds[k]] + threshold, ... until n; *count = number of sub-packets
// (at most `cap`, the last one takes the rest).  One thread, a binary search per cut.
__global__ void activity_plan_kernel(const int4* events, long long n, long long threshold, long long* bounds, int cap, int* count) {
    int k = 0;
    long long lo = 0;
    bounds[0] = 0;
    while (lo < n && k < cap - 1) {
        const long long limit = unpack_event(ld_event_plain(events + lo)).t_bits + threshold;
        long long a = lo + 1, b = n;  // first index in (lo, n] with t >= limit
        while (a < b) {
            const long long m = a + (b - a) / 2;
            if (unpack_event(ld_event_plain(events + m)).t_bits >= limit) b = m; else a = m + 1;
        }
        bounds[++k] = a;
        lo = a;
    }
    if (lo < n) bounds[++k] = n;
    *count = k;
}

// ---------------------------------------------------------------------------------------------
// N4: de-duplication filters.  Key image of `rows` x `stride` cells; per cell the index of the first
// and of the last event that hit it (atomicMin / atomicMax).  Modes as XM_FILTER_*.
// ---------------------------------------------------------------------------------------------
constexpr int kFilterFirstYT = 1, kFilterFirstXY = 2, kFilterLastXY = 3, kFilterMeanXY = 4;

struct FilterParams {
    const int4* events;
    long long n;
    const short* xp;  // YT: rectified x of every event (rectify_cam_coords_i16), else NULL
    int mode;
    int rows, cols;   // camera image
    int stride;       // cells per key-image row: cols (XY) or the largest rectified x + 1 the table can produce (YT)
    int as_reference; // 1: "first" filters keep the LAST duplicate, like the reference as it runs under NumPy
    unsigned* first;  // [rows * stride]
    unsigned* last;
    int* xp_max;      // device scalar
    FrameState* state;
};

__global__ void __launch_bounds__(256) filter_prepare_kernel(const FilterParams p) {
    const long long cells = static_cast<long long>(p.rows) * p.stride;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < cells; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        p.first[i] = 0xffffffffu;
        p.last[i] = 0u;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.xp_max = INT_MIN;
}

// xp_i16.max() (frame_event_filter.py:75)
__global__ void __launch_bounds__(256) filter_xpmax_kernel(const FilterParams p) {
    int m = INT_MIN;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n; i += static_cast<long long>(gridDim.x) * blockDim.x)
        m = max(m, static_cast<int>(p.xp[i]));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m != INT_MIN) atomicMax(p.xp_max, m);
}

__global__ void __launch_bounds__(256) filter_mark_kernel(const FilterParams p) {
    unsigned flags = 0;
    const int width = p.mode == kFilterFirstYT ? *p.xp_max + 1 : p.cols;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const EventFields e = unpack_event(ld_event_plain(p.events + i));
        if (e.p != 1) {
            if (p.mode == kFilterFirstYT) flags |= kStatusFilterPolarity;
            continue;  // events[events["p"] == 1]  (:21,47,72,104)
        }
        if (e.x >= static_cast<unsigned>(p.cols) || e.y >= static_cast<unsigned>(p.rows)) {
            flags |= kStatusPixelOob;
            continue;
        }
        int col = static_cast<int>(e.x);
        if (p.mode == kFilterFirstYT) {
            col = p.xp[i];
            if (col < 0) col += width;  // a negative NumPy index counts from the end of the axis
            if (col < 0 || col >= p.stride) {
                flags |= kStatusFilterIndex;
                continue;
            }
        }
        const long long cell = static_cast<long long>(e.y) * p.stride + col;
        atomicMin(p.first + cell, static_cast<unsigned>(i));
        atomicMax(p.last + cell, static_cast<unsigned>(i) + 1u);
    }
    flags = __reduce_or_sync(0xffffffffu, flags);
    if ((threadIdx.x & 31) == 0 && flags) atomicOr(&p.state->flags, flags);
}

struct FilterPred {
    const unsigned* last;
    __device__ __forceinline__ bool operator()(long long i) const { return last[i] != 0u; }
};

struct FilterEmit {
    FilterParams p;
    int4* out;
    __device__ __forceinline__ int t32(unsigned idx) const {  // int32_image[...] = events["t"]: the timestamp wraps
        return static_cast<int>(static_cast<unsigned>(unpack_event(ld_event_plain(p.events + idx)).t_bits));
    }
    __device__ __forceinline__ void operator()(long long cell, unsigned pos) const {
        const unsigned y = static_cast<unsigned>(cell / p.stride);
        unsigned x = static_cast<unsigned>(cell - static_cast<long long>(y) * p.stride);
        const unsigned li = p.last[cell] - 1u;
        const unsigned fi = p.as_reference ? li : p.first[cell];
        int t;
        if (p.mode == kFilterLastXY) {
            t = t32(li);
        } else if (p.mode == kFilterMeanXY) {
            t = static_cast<int>(static_cast<unsigned>(t32(li)) + static_cast<unsigned>(t32(fi))) >> 1;  // int32 add wraps, // 2 floors
        } else {
            t = t32(fi);
            if (p.mode == kFilterFirstYT) x = unpack_event(ld_event_plain(p.events + fi)).x;
        }
        out[pos] = make_int4(static_cast<int>((x & 0xffffu) | (y << 16)), 1, t, t >> 31);
    }
};

// ---------------------------------------------------------------------------------------------
// N2: frame segmentation.  pauses = nonzero(diff(t) >= thresh) (:155); the first pair of consecutive
// pauses further apart than half a frame decides (:161-187).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long event_time(const int4* ev, long long i) { return unpack_event(ld_event_plain(ev + i)).t_bits; }

struct PausePred {
    const int4* events;
    long long n;
    long long thresh;
    __device__ __forceinline__ bool operator()(long long i) const {
        return i + 1 < n && event_time(events, i + 1) - event_time(events, i) >= thresh;
    }
};
struct PauseEmit {
    unsigned* idx;
    __device__ __forceinline__ void operator()(long long i, unsigned pos) const { idx[pos] = static_cast<unsigned>(i); }
};

struct TriggerScratch {
    long long n_pauses;
    unsigned first_pair;  // smallest k with t[pause[k+1]] - t[pause[k]] > frame / 2, 0xffffffff = none
    unsigned pad;
};

__global__ void __launch_bounds__(256) trigger_pairs_kernel(const int4* __restrict__ events, const unsigned* __restrict__ idx,
                                                            TriggerScratch* sc, double half_frame_us) {
    const long long pairs = sc->n_pauses - 1;
    unsigned best = 0xffffffffu;
    for (long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < pairs; k += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long gap = event_time(events, idx[k + 1]) - event_time(events, idx[k]);
        if (static_cast<double>(gap) > half_frame_us) {
            best = static_cast<unsigned>(k);
            break;  // k only grows in this thread
        }
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if ((threadIdx.x & 31) == 0 && best != 0xffffffffu) atomicMin(&sc->first_pair, best);
}

// result[0..5] = status, prev_idx, next_idx, n_pauses, start_time, end_time
__global__ void trigger_decide_kernel(const int4* __restrict__ events, const unsigned* __restrict__ idx, const TriggerScratch* sc,
                                      double frame_us, long long min_events, long long* __restrict__ result) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long status = -1, prev = -1, next = -1, t0 = -1, t1 = -1;
    const unsigned k = sc->first_pair;
    if (k != 0xffffffffu) {
        prev = idx[k];
        next = idx[k + 1];
        const long long gap = event_time(events, next) - event_time(events, prev);
        if (static_cast<double>(gap) <= frame_us && next - prev > min_events) {
            status = 1;
            t0 = event_time(events, prev + 2);
            t1 = event_time(events, next - 2);
        } else {
            status = 0;
        }
    }
    result[0] = status;
    result[1] = prev;
    result[2] = next;
    result[3] = sc->n_pauses;
    result[4] = t0;
    result[5] = t1;
    result[6] = 0;
    result[7] = 0;
}

}  // namespace xm
