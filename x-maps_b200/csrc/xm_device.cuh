// xm_device.cuh — device-side building blocks shared by the X-maps kernels (sm_100a).
//
// Everything here is integer / IEEE-exact arithmetic: the path must reproduce the reference's
// NumPy results bit for bit, so no fast-math, explicit _rn intrinsics where a
// fused multiply-add could otherwise be contracted.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace xm {

// ---------------------------------------------------------------------------------------------
// Event record: Metavision EventCD, 16-byte AoS, read as one int4.
//   .x = x | y << 16     .y = p (low 16 bits) | pad     .z/.w = t (int64 or float64, little endian)
// ---------------------------------------------------------------------------------------------
struct EventFields {
    unsigned x, y;
    int p;
    long long t_bits;
};

__device__ __forceinline__ EventFields unpack_event(const int4& r) {
    EventFields e;
    e.x = static_cast<unsigned>(r.x) & 0xffffu;
    e.y = static_cast<unsigned>(r.x) >> 16;
    e.p = static_cast<int>(static_cast<short>(r.y & 0xffff));
    e.t_bits = (static_cast<long long>(r.w) << 32) | static_cast<unsigned>(r.z);
    return e;
}

// Streaming 128-bit load: read-only path, do not allocate in L1, evict-first in L2 so the
// event stream (80 MB / frame) does not push the tables and the scatter map out of the 126 MB L2.
__device__ __forceinline__ uint64_t make_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ int4 ld_event_stream(const int4* p, uint64_t policy) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(policy));
    return r;
}

__device__ __forceinline__ int4 ld_event_plain(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// ---------------------------------------------------------------------------------------------
// Time bounds.  For float64 timestamps the bounds are kept as "sortable" int64 so that the same
// integer atomics implement min / max (IEEE order == integer order after this mapping).
// ---------------------------------------------------------------------------------------------
__device__ __host__ __forceinline__ long long f64_bits_to_sortable(long long b) {
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __host__ __forceinline__ long long sortable_to_f64_bits(long long s) {
    return s ^ ((s >> 63) & 0x7fffffffffffffffLL);
}

// Normalisation constants of one frame (python/x_maps_disparity.py:12-19).
template <bool F64>
struct TimeNorm {
    long long lo_i;  // int64 mode
    long long hi_i;
    double lo_f;     // float64 mode
    double hi_f;
    double den;      // (max_t - min_t) as float64
    double scale;    // T_PX_SCALE

    __device__ __forceinline__ void init(long long lo_bits, long long hi_bits, int t_px_scale) {
        scale = static_cast<double>(t_px_scale);
        if (F64) {
            lo_f = __longlong_as_double(lo_bits);
            hi_f = __longlong_as_double(hi_bits);
            den = __dsub_rn(hi_f, lo_f);
            lo_i = hi_i = 0;
        } else {
            lo_i = lo_bits;
            hi_i = hi_bits;
            den = static_cast<double>(hi_bits - lo_bits);
            lo_f = hi_f = 0.0;
        }
    }

    // true if t lies outside [lo, hi] (only possible when the bounds were assumed, not reduced)
    __device__ __forceinline__ bool outside(long long t_bits) const {
        if (F64) {
            double t = __longlong_as_double(t_bits);
            return !(t >= lo_f && t <= hi_f);
        }
        return t_bits < lo_i || t_bits > hi_i;
    }

    // int16(rint(((t - min) / (max - min)) * T_PX_SCALE)) in float64, two roundings, half-even.
    // 0/0 (all timestamps equal) is NaN in NumPy and casts to 0.
    __device__ __forceinline__ int column(long long t_bits) const {
        double num = F64 ? __dsub_rn(__longlong_as_double(t_bits), lo_f) : static_cast<double>(t_bits - lo_i);
        double r = rint(__dmul_rn(__ddiv_rn(num, den), scale));
        if (!(r == r)) return 0;
        // clamp only guards memory safety for flagged (out-of-bounds) events
        r = fmin(fmax(r, -32768.0), 32767.0);
        return static_cast<int>(r);
    }
};

// Same arithmetic with a cheap common case.  a = fl(num * fl(scale / den)) differs from the
// reference's s = fl(fl(num / den) * scale) by < 4e-13 (both are within 2 ulp-ish of the exact
// value, which is < 2^15), so rint(a) == rint(s) unless a lies within 1e-4 of a half-integer; only
// then (exact ties in practice) is the division evaluated.  Integer timestamps whose range fits 31
// bits also avoid the 64-bit int -> double conversion.
template <bool F64>
struct TimeCol {
    unsigned long long lo_u, range;
    double lo_f, hi_f, den, scale, inv;
    bool fast;

    __device__ __forceinline__ void init(long long lo_bits, long long hi_bits, int t_px_scale) {
        scale = static_cast<double>(t_px_scale);
        if (F64) {
            lo_f = __longlong_as_double(lo_bits);
            hi_f = __longlong_as_double(hi_bits);
            den = __dsub_rn(hi_f, lo_f);
            lo_u = range = 0;
            fast = den > 0.0 && den < 1.7e308;
        } else {
            lo_u = static_cast<unsigned long long>(lo_bits);
            range = hi_bits >= lo_bits ? static_cast<unsigned long long>(hi_bits) - lo_u : 0ULL;
            den = static_cast<double>(hi_bits - lo_bits);
            lo_f = hi_f = 0.0;
            fast = range > 0 && range < 0x80000000ULL && hi_bits >= lo_bits;
        }
        inv = fast ? __ddiv_rn(scale, den) : 0.0;
    }

    __device__ __forceinline__ int exact(double num) const {
        double r = rint(__dmul_rn(__ddiv_rn(num, den), scale));
        if (!(r == r)) return 0;  // 0/0 (all timestamps equal) is NaN in NumPy and casts to 0
        r = fmin(fmax(r, -32768.0), 32767.0);  // memory safety only (flagged events)
        return static_cast<int>(r);
    }

    // int16(rint(((t - min) / (max - min)) * T_PX_SCALE)); `viol` = t outside the assumed [min, max]
    __device__ __forceinline__ int column(long long t_bits, bool& viol) const {
        double num;
        if (F64) {
            const double t = __longlong_as_double(t_bits);
            viol = !(t >= lo_f && t <= hi_f);
            num = __dsub_rn(t, lo_f);
        } else {
            const unsigned long long dt = static_cast<unsigned long long>(t_bits) - lo_u;
            viol = dt > range;
            num = fast ? static_cast<double>(static_cast<unsigned>(dt)) : static_cast<double>(static_cast<long long>(dt));
        }
        if (fast && !viol) {
            const double a = __dmul_rn(num, inv);
            const int k = __double2int_rn(a);
            if (fabs(__dsub_rn(a, static_cast<double>(k))) < 0.4999) return k;
        }
        return exact(num);
    }

    // Branch-free hot-loop variant (valid only when `fast`): the caller patches `tie` cases with
    // exact_from_bits().  The returned column is meaningful only if !viol && !tie.
    __device__ __forceinline__ int column_fast_nb(long long t_bits, bool& viol, bool& tie) const {
        double num;
        if (F64) {
            const double t = __longlong_as_double(t_bits);
            viol = !(t >= lo_f && t <= hi_f);
            num = __dsub_rn(t, lo_f);
        } else {
            const unsigned long long dt = static_cast<unsigned long long>(t_bits) - lo_u;
            viol = dt > range;
            num = static_cast<double>(static_cast<unsigned>(dt));
        }
        const double a = __dmul_rn(num, inv);
        const int k = __double2int_rn(a);
        tie = !(fabs(__dsub_rn(a, static_cast<double>(k))) < 0.4999);
        return k;
    }
    __device__ __forceinline__ int exact_from_bits(long long t_bits) const {
        if (F64) return exact(__dsub_rn(__longlong_as_double(t_bits), lo_f));
        return exact(static_cast<double>(static_cast<long long>(static_cast<unsigned long long>(t_bits) - lo_u)));
    }

    // Hot-loop variant, valid only when `fast`: returns a column in [0, T_PX_SCALE] (0 if `viol`).
    __device__ __forceinline__ int column_fast(long long t_bits, bool& viol) const {
        double num;
        if (F64) {
            const double t = __longlong_as_double(t_bits);
            viol = !(t >= lo_f && t <= hi_f);
            num = __dsub_rn(t, lo_f);
        } else {
            const unsigned long long dt = static_cast<unsigned long long>(t_bits) - lo_u;
            viol = dt > range;
            num = static_cast<double>(static_cast<unsigned>(dt));
        }
        const double a = __dmul_rn(num, inv);
        const int k = __double2int_rn(a);
        if (viol) return 0;
        if (fabs(__dsub_rn(a, static_cast<double>(k))) < 0.4999) return k;
        return exact(num);
    }
};

// Integer timestamps, all-integer evaluation.  With dt = t - min, R = max - min, s = T_PX_SCALE the
// reference's rint(fl(fl(dt / R) * s)) equals floor((2 dt s + R) / (2 R)) unless 2 dt s + R is a multiple
// of 2R (an exact .5 tie, where half-even and the double rounding decide): the float64 result is within
// 4e-13 of dt s / R while a non-tie is at least 1 / (2R) > 2e-10 away from any half-integer.  The
// division by the frame constant 2R is a multiply-high with a magic number (exact for n < 2^31).
struct IntCol {
    unsigned long long lo;
    unsigned range, scale2, d, M;
    int sh;
    bool ok;  // R > 0 and (2 s + 1) R < 2^31 (frames up to ~1.5 s at T_PX_SCALE = 719)

    __device__ __forceinline__ void init(long long lo_bits, long long hi_bits, int t_px_scale) {
        lo = static_cast<unsigned long long>(lo_bits);
        const unsigned long long r = hi_bits > lo_bits ? static_cast<unsigned long long>(hi_bits) - lo : 0ULL;
        const unsigned long long nmax = r * static_cast<unsigned long long>(2 * t_px_scale + 1);
        ok = r > 0 && t_px_scale > 0 && r < 0x40000000ULL && nmax < 0x80000000ULL;
        range = static_cast<unsigned>(r);
        scale2 = 2u * static_cast<unsigned>(t_px_scale);
        d = 2u * range;
        M = 0;
        sh = 0;
        if (ok) {
            const int e = 31 - __clz(d);  // floor(log2 d), d >= 2
            if ((d & (d - 1u)) == 0u) {   // power of two: n / d = (n >> 1) >> (e - 1)
                M = 0x80000000u;
                sh = e - 1;
            } else {                      // 2^e < d < 2^(e+1): M = ceil(2^(32+e) / d) fits 32 bits
                M = static_cast<unsigned>(((1ULL << (32 + e)) + d - 1u) / d);
                sh = e;
            }
        }
    }
    // column of t; `bad`: t outside [min, max] or an exact rounding tie (caller takes the float64 path)
    __device__ __forceinline__ unsigned column(long long t_bits, bool& bad) const {
        const unsigned long long dt = static_cast<unsigned long long>(t_bits) - lo;
        const unsigned n = static_cast<unsigned>(dt) * scale2 + range;
        const unsigned q = __umulhi(n, M) >> sh;
        bad = dt > range || n == q * d;
        return q;
    }
};

// ---------------------------------------------------------------------------------------------
// Scatter keys: the reference's `map[rows, cols] = vals` keeps the LAST duplicate.  Each event
// contributes key = epoch:16 | event_index:32 | disparity:16 and the map keeps the maximum, i.e.
// the highest event index of the current frame; older frames (lower epoch) lose automatically,
// so the map never has to be cleared between frames.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long make_key(unsigned epoch, unsigned long long index, int disp) {
    return (static_cast<unsigned long long>(epoch) << 48) | ((index & 0xffffffffULL) << 16) |
           static_cast<unsigned long long>(static_cast<unsigned>(disp) & 0xffffu);
}

// same key from a 32-bit event index, built from two 32-bit halves
__device__ __forceinline__ unsigned long long make_key32(unsigned epoch, unsigned index, int disp) {
    const unsigned hi = (epoch << 16) | (index >> 16);
    const unsigned lo = (index << 16) | (static_cast<unsigned>(disp) & 0xffffu);
    return (static_cast<unsigned long long>(hi) << 32) | lo;
}

__device__ __forceinline__ int key_disparity(unsigned long long key, unsigned epoch) {
    return (static_cast<unsigned>(key >> 48) == epoch) ? static_cast<int>(key & 0xffffULL) : 0;
}

// ---------------------------------------------------------------------------------------------
// Output conversions (python/disp_to_depth.py)
// ---------------------------------------------------------------------------------------------
// disparity_to_depth_rectified :46-63 — float64 divide, float64 max, one rounding to float32.
__device__ __forceinline__ float disparity_to_depth(float d, double depth_scale) {
    if (d == 0.0f) return 0.0f;
    return __double2float_rn(fmax(__ddiv_rn(depth_scale, static_cast<double>(d)), 1e-9));
}

// clip_normalize_uint8_depth_frame :7-21 — float32 clip / subtract / divide, float64 * 255, truncate.
__device__ __forceinline__ unsigned depth_to_u8(float depth, float z_near, float z_far) {
    if (depth == 0.0f) return 0u;
    float v = fmaxf(fminf(depth, z_far), z_near);
    float frac = __fdiv_rn(__fsub_rn(v, z_near), __fsub_rn(z_far, z_near));
    double scaled = __dmul_rn(static_cast<double>(frac), 255.0);
    return static_cast<unsigned>(static_cast<int>(scaled)) & 0xffu;
}

struct OutputSpec {
    int kind;            // XM_OUT_*
    double depth_scale;  // P2[0,3]
    float z_near, z_far;
    const unsigned char* turbo_bgr;  // [256][3]
    const float* depth_lut;          // [32768] depth of every integer disparity (exact f64 divide), or NULL
};

// Writes one pixel of the frame from an integer-valued disparity.
__device__ __forceinline__ void emit_pixel(const OutputSpec& o, void* out, long long idx, float disp) {
    if (o.kind == 1) {  // XM_OUT_DISPARITY
        static_cast<float*>(out)[idx] = disp;
    } else if (o.kind == 0) {  // XM_OUT_DEPTH
        static_cast<float*>(out)[idx] = disparity_to_depth(disp, o.depth_scale);
    } else {  // XM_OUT_BGR: TURBO colour map, white where there is no depth (:24-43)
        unsigned u = depth_to_u8(disparity_to_depth(disp, o.depth_scale), o.z_near, o.z_far);
        unsigned char* px = static_cast<unsigned char*>(out) + idx * 3;
        if (u == 0u) {
            px[0] = 255; px[1] = 255; px[2] = 255;
        } else {
            px[0] = o.turbo_bgr[u * 3 + 0];
            px[1] = o.turbo_bgr[u * 3 + 1];
            px[2] = o.turbo_bgr[u * 3 + 2];
        }
    }
}

// Same for the fused epilogues, whose disparities are integers in [0, 32767]: the depth comes from a
// per-context table filled once with disparity_to_depth() (bit-identical, no division per pixel).
__device__ __forceinline__ void emit_pixel_int(const OutputSpec& o, void* out, long long idx, int disp) {
    if (o.kind == 1) {
        static_cast<float*>(out)[idx] = static_cast<float>(disp);
    } else if (o.kind == 0) {
        float z = 0.0f;
        if (disp != 0) z = o.depth_lut ? __ldg(o.depth_lut + disp) : disparity_to_depth(static_cast<float>(disp), o.depth_scale);
        static_cast<float*>(out)[idx] = z;
    } else {
        emit_pixel(o, out, idx, static_cast<float>(disp));
    }
}

// N pixels at once: all depth-table look-ups are issued before the first store, so their latencies overlap
// (`live` masks pixels that do not exist; idx[j] is only used when live[j]).
template <int N>
__device__ __forceinline__ void emit_pixels_int(const OutputSpec& o, void* out, const int (&idx)[N], const int (&disp)[N], const bool (&live)[N]) {
    if (o.kind == 0 && o.depth_lut) {
        float z[N];
#pragma unroll
        for (int j = 0; j < N; ++j) z[j] = (live[j] && disp[j] != 0) ? __ldg(o.depth_lut + disp[j]) : 0.0f;
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (live[j]) static_cast<float*>(out)[idx[j]] = z[j];
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (live[j]) emit_pixel_int(o, out, idx[j], disp[j]);
    }
}

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA) helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the waiting thread is parked until the phase completes (or the hint
// expires) instead of re-issuing try_wait + branch.  Without the hint the wait loops of the producer lanes and
// the consumer warps were 14 % of all warp instructions the batch kernel issued (ncu source view, r02f).
#ifndef XM_MBAR_HINT_NS
#define XM_MBAR_HINT_NS 0x989680
#endif
constexpr unsigned kMbarSuspendHintNs = XM_MBAR_HINT_NS;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "XM_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra XM_DONE_%=;\n\t"
        "bra XM_WAIT_%=;\n\t"
        "XM_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(kMbarSuspendHintNs)
        : "memory");
}
// Wait of a thread that is NOT on the critical path (a producer lane waiting for a free ring slot): polls with a
// sleep in between, so that its wait loop does not take issue slots from the warps that do the work.
__device__ __forceinline__ bool mbar_test(uint64_t* bar, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, unsigned parity) {
#ifdef XM_PRODUCER_SLEEP_NS
    while (!mbar_test(bar, parity)) __nanosleep(XM_PRODUCER_SLEEP_NS);
#else
    mbar_wait(bar, parity);
#endif
}
// global -> shared bulk copy, completion signalled on the mbarrier (bytes % 16 == 0, 16 B aligned)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// 4-byte asynchronous gather global -> shared (LDGSTS): completion is tracked per thread by
// cp.async groups, not by a register scoreboard, so a later chunk's gathers never stall this one
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// raw 32-bit shared-address variants (the address arithmetic is hoisted out of the hot loops)
__device__ __forceinline__ void mbar_wait_a(unsigned bar_addr, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "XM_WAITA_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra XM_DONEA_%=;\n\t"
        "bra XM_WAITA_%=;\n\t"
        "XM_DONEA_%=:\n\t"
        "}" ::"r"(bar_addr),
        "r"(parity), "r"(kMbarSuspendHintNs)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(unsigned bar_addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ int4 lds128_a(unsigned addr) {
    int4 r;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ int lds32_a(unsigned addr) {
    int r;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ void sts32_a(unsigned addr, unsigned v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int lds_s16_a(unsigned addr) {
    int r;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ unsigned lds_u16_a(unsigned addr) {
    unsigned r;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_4_a(unsigned smem_addr, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(gmem_src) : "memory");
}
// predicated 64-bit max reduction (no return value -> RED)
__device__ __forceinline__ void red_max_u64_if(unsigned long long* addr, unsigned long long val, bool pred) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.b32 q, %2, 0;\n\t"
        "@q red.global.max.u64 [%0], %1;\n\t"
        "}" ::"l"(addr),
        "l"(val), "r"(static_cast<int>(pred))
        : "memory");
}

// same with an L2 cache-policy hint (evict-first for the one-pass event stream)
__device__ __forceinline__ void tma_load_1d_hint(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

}  // namespace xm
