// xm_stage_kernels.cuh — the same path stage by stage, with materialised intermediates.
//
// These kernels back the reference's per-stage call surface (rectify_cam_coords_*,
// compute_event_disparity, compute_disp_map_*, remap_rectified_disp_map_to_proj,
// disparity_to_depth_rectified, colorize_depth_from_disp, construct_point_cloud) for callers that
// look at the intermediates (python/eval/compute_depth_x_maps.py:97-122, dump_frame_data in
// python/depth_reprojection_pipe.py:19-34).  The fused kernels of xm_frame_kernels.cuh never
// materialise them; these are the compatibility path, written for clarity first.
#pragma once
#include "xm_frame_kernels.cuh"

namespace xm {

// CamProjMaps.rectify_cam_coords_i16 (cam_proj_calibration.py:277-281)
__global__ void rectify_i16_kernel(const int4* __restrict__ events, long long n, const int* __restrict__ lut_xy, int cam_w,
                                   int cam_h, short* __restrict__ xr, short* __restrict__ yr, FrameState* st) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        EventFields e = unpack_event(ld_event_plain(events + i));
        int v = 0;
        if (e.x < static_cast<unsigned>(cam_w) && e.y < static_cast<unsigned>(cam_h))
            v = __ldg(lut_xy + static_cast<int>(e.y) * cam_w + static_cast<int>(e.x));
        else
            atomicOr(&st->flags, kStatusPixelOob);
        xr[i] = static_cast<short>(v & 0xffff);
        yr[i] = static_cast<short>(v >> 16);
    }
}

// CamProjMaps.rectify_cam_coords_f32 (cam_proj_calibration.py:272-275)
__global__ void rectify_f32_kernel(const int4* __restrict__ events, long long n, const float* __restrict__ lx,
                                   const float* __restrict__ ly, int cam_w, int cam_h, float* __restrict__ xr,
                                   float* __restrict__ yr, FrameState* st) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        EventFields e = unpack_event(ld_event_plain(events + i));
        float a = 0.f, b = 0.f;
        if (e.x < static_cast<unsigned>(cam_w) && e.y < static_cast<unsigned>(cam_h)) {
            const int pix = static_cast<int>(e.y) * cam_w + static_cast<int>(e.x);
            a = __ldg(lx + pix);
            b = __ldg(ly + pix);
        } else {
            atomicOr(&st->flags, kStatusPixelOob);
        }
        xr[i] = a;
        yr[i] = b;
    }
}

// compute_disparity (x_maps_disparity.py:9-32), one thread per event, un-compacted outputs.
struct DisparityParams {
    const int4* events;
    long long n;
    int polarity;
    const int* lut_xy;
    int cam_w, cam_h;
    const short* x_rect;  // optional
    const short* y_rect;
    const short* xmap_t;
    int xmap_w, xmap_h, col_stride;
    int t_px_scale, x_offset;
    short* disp_full;
    unsigned char* mask;
    FrameState* state;
    int verify;
};

template <bool F64>
__global__ void __launch_bounds__(256) event_disparity_kernel(const DisparityParams p) {
    TimeNorm<F64> tn;
    tn.init(p.state->t_lo_bits, p.state->t_hi_bits, p.t_px_scale);
    unsigned n_valid = 0, n_inl = 0, flags = 0;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        EventFields e = unpack_event(ld_event_plain(p.events + i));
        int disp = -1;
        bool inl = false;
        if (event_valid(e, p.polarity)) {
            ++n_valid;
            int xcr = 0, ycr = -1;
            if (p.x_rect) {
                xcr = p.x_rect[i];
                ycr = p.y_rect[i];
            } else if (e.x < static_cast<unsigned>(p.cam_w) && e.y < static_cast<unsigned>(p.cam_h)) {
                int v = __ldg(p.lut_xy + static_cast<int>(e.y) * p.cam_w + static_cast<int>(e.x));
                xcr = static_cast<short>(v & 0xffff);
                ycr = v >> 16;
            } else {
                flags |= kStatusPixelOob;
            }
            if (p.verify && tn.outside(e.t_bits)) flags |= kStatusTBounds;
            int c = tn.column(e.t_bits);
            if (c < 0) c += p.xmap_w;
            if (c < 0 || c >= p.xmap_w) {
                flags |= kStatusTBounds;
                c = 0;
            }
            if (ycr >= 0 && ycr < p.xmap_h - 1) {
                int xp = __ldg(p.xmap_t + static_cast<long long>(c) * p.col_stride + ycr);
                int d = static_cast<short>(xp - xcr - p.x_offset);
                if (d >= 0) {
                    disp = d;
                    inl = true;
                    ++n_inl;
                }
            }
        }
        p.disp_full[i] = static_cast<short>(disp);
        p.mask[i] = inl ? 1 : 0;
    }
    n_valid = __reduce_add_sync(0xffffffffu, n_valid);
    n_inl = __reduce_add_sync(0xffffffffu, n_inl);
    flags = __reduce_or_sync(0xffffffffu, flags);
    if ((threadIdx.x & 31) == 0) {
        if (n_valid) atomicAdd(&p.state->n_valid, static_cast<unsigned long long>(n_valid));
        if (n_inl) atomicAdd(&p.state->n_inliers, static_cast<unsigned long long>(n_inl));
        if (flags) atomicOr(&p.state->flags, flags);
    }
}

// ---------------------------------------------------------------------------------------------
// Order-preserving compaction (the `disp[disp_inlier_mask]` of x_maps_disparity.py:32) in three
// steps: per-block counts, one-block exclusive scan of the counts, per-block ordered write.
// ---------------------------------------------------------------------------------------------
constexpr int kCompactBlock = 1024;  // elements per block (256 threads x 4)

__global__ void __launch_bounds__(256) compact_count_kernel(const unsigned char* __restrict__ mask, long long n,
                                                            unsigned* __restrict__ counts) {
    const long long base = static_cast<long long>(blockIdx.x) * kCompactBlock;
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        long long i = base + threadIdx.x * 4 + k;
        if (i < n && mask[i]) ++c;
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

// single block: counts[i] <- exclusive prefix; *total <- sum
__global__ void __launch_bounds__(1024) compact_scan_kernel(unsigned* counts, long long n_blocks, long long* total) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (long long base = 0; base < n_blocks; base += 1024) {
        long long i = base + threadIdx.x;
        unsigned v = i < n_blocks ? counts[i] : 0u;
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned long long w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += t;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        unsigned long long warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0ULL;
        unsigned long long excl = s_carry + warp_off + incl - v;
        if (i < n_blocks) counts[i] = static_cast<unsigned>(excl);  // totals < 2^32 (n_events fits the key's 32 bits)
        __syncthreads();
        if (threadIdx.x == 1023) s_carry += warp_off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = static_cast<long long>(s_carry);
}

template <typename T>
__global__ void __launch_bounds__(256) compact_write_kernel(const T* __restrict__ vals, const unsigned char* __restrict__ mask,
                                                            long long n, const unsigned* __restrict__ offsets,
                                                            T* __restrict__ out) {
    const long long base = static_cast<long long>(blockIdx.x) * kCompactBlock;
    bool keep[4];
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        long long i = base + threadIdx.x * 4 + k;
        keep[k] = i < n && mask[i];
        c += keep[k];
    }
    // exclusive scan of c over the block (thread order == element order)
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    __shared__ unsigned s[8];
    if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned off = offsets[blockIdx.x] + incl - c;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) off += s[w];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (keep[k]) out[off++] = vals[base + threadIdx.x * 4 + k];
    }
}

// ---------------------------------------------------------------------------------------------
// `m = zeros(h, w); m[rows, cols] = vals` with last-write-wins (cam_proj_calibration.py:301-302,
// 315-316): keys into the context's scatter map, then a decode pass.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scatter_keys_kernel(const short* __restrict__ rows, const short* __restrict__ cols,
                                                           const short* __restrict__ vals, long long n, int h, int w,
                                                           unsigned long long* map, unsigned epoch, FrameState* st) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        int r = rows[i], c = cols[i];
        if (r < 0) r += h;  // NumPy negative indices
        if (c < 0) c += w;
        if (r < 0 || r >= h || c < 0 || c >= w) {
            atomicOr(&st->flags, kStatusScatterOob);
            continue;
        }
        // value keeps all 16 bits (the staged path may carry any int16, incl. negative values)
        atomicMax(map + static_cast<long long>(r) * w + c, make_key(epoch, static_cast<unsigned long long>(i), vals[i]));
    }
}

__global__ void __launch_bounds__(256) decode_map_kernel(const unsigned long long* __restrict__ map, long long n, unsigned epoch,
                                                         float* __restrict__ out) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        unsigned long long k = map[i];
        float v = 0.f;
        if (static_cast<unsigned>(k >> 48) == epoch) v = static_cast<float>(static_cast<short>(k & 0xffffULL));
        out[i] = v;
    }
}

// DisparityToDepth.remap_rectified_disp_map_to_proj (disp_to_depth.py:76-97) on a float32 map.
// Generic values (negative allowed): border taps are skipped, exactly as cv2.dilate does.
__global__ void __launch_bounds__(256) dilate_remap_kernel(const float* __restrict__ src, int rect_w, int rect_h,
                                                           const short2* __restrict__ remap_xy, int out_w, int out_h, int radius,
                                                           float* __restrict__ dst) {
    const long long n = static_cast<long long>(out_w) * out_h;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        short2 m = __ldg(remap_xy + i);
        float best = 0.f;
        if (m.x >= 0 && m.x < rect_w && m.y >= 0 && m.y < rect_h) {
            best = -INFINITY;
            for (int dy = -radius; dy <= radius; ++dy) {
                int gy = m.y + dy;
                if (gy < 0 || gy >= rect_h) continue;
                for (int dx = -radius; dx <= radius; ++dx) {
                    int gx = m.x + dx;
                    if (gx < 0 || gx >= rect_w) continue;
                    best = fmaxf(best, __ldg(src + static_cast<long long>(gy) * rect_w + gx));
                }
            }
        }
        dst[i] = best;
    }
}

// disparity_to_depth_rectified / colorize_depth_from_disp on flat float32 arrays
__global__ void __launch_bounds__(256) convert_kernel(const float* __restrict__ disp, long long n, OutputSpec out, void* dst) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        emit_pixel(out, dst, i, disp[i]);
}

// depth of every integer disparity 0..n-1 (context set-up; feeds emit_pixel_int)
__global__ void __launch_bounds__(256) depth_lut_kernel(float* __restrict__ lut, int n, double depth_scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lut[i] = disparity_to_depth(static_cast<float>(i), depth_scale);
}

// CamProjMaps.construct_point_cloud (cam_proj_calibration.py:319-331), float32 arithmetic.
struct Mat4f {
    float m[16];
};
__global__ void __launch_bounds__(256) point_cloud_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                          const float* __restrict__ d, long long n, Mat4f q,
                                                          float* __restrict__ xyz) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float v0 = __fadd_rn(x[i], d[i]), v1 = y[i], v2 = -d[i];
        float o[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float acc = __fmul_rn(q.m[r * 4 + 0], v0);
            acc = __fadd_rn(acc, __fmul_rn(q.m[r * 4 + 1], v1));
            acc = __fadd_rn(acc, __fmul_rn(q.m[r * 4 + 2], v2));
            acc = __fadd_rn(acc, q.m[r * 4 + 3]);
            o[r] = acc;
        }
        xyz[i * 3 + 0] = __fdiv_rn(o[0], o[3]);
        xyz[i * 3 + 1] = -__fdiv_rn(o[1], o[3]);
        xyz[i * 3 + 2] = -__fdiv_rn(o[2], o[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// compute_x_map_from_time_map (python/x_map.py:5-55).  One CTA per rectified row: the row of the
// time map sits in shared memory, each thread owns time coordinates and scans all x (broadcast
// reads).  float64 arithmetic, first minimum wins, cells equal to 0 are undefined.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_xmap_kernel(const float* __restrict__ time_map, int h, int w, int x_map_width,
                                                         int t_px_scale, int x_offset, int num_scanlines,
                                                         short* __restrict__ x_map, float* __restrict__ t_diffs) {
    extern __shared__ float s_row[];
    const int y = blockIdx.x;
    for (int x = threadIdx.x; x < w; x += blockDim.x) s_row[x] = time_map[static_cast<long long>(y) * w + x];
    __syncthreads();
    const double max_t_diff = __ddiv_rn(2.0, static_cast<double>(num_scanlines));
    for (int tc = threadIdx.x; tc < x_map_width; tc += blockDim.x) {
        short xv = 0;
        float dv = 0.f;
        const double t = __ddiv_rn(static_cast<double>(tc), static_cast<double>(t_px_scale));
        if (t != 0.0) {
            double best = INFINITY;
            int best_x = -1;
            for (int x = 0; x < w; ++x) {
                const float tm = s_row[x];
                if (tm == 0.f) continue;
                const double diff = fabs(__dsub_rn(t, static_cast<double>(tm)));
                if (diff < best) {
                    best = diff;
                    best_x = x;
                }
            }
            if (best_x != -1 && best <= max_t_diff) {
                xv = static_cast<short>(best_x + x_offset);
                dv = static_cast<float>(best);
            }
        }
        x_map[static_cast<long long>(y) * x_map_width + tc] = xv;
        if (t_diffs) t_diffs[static_cast<long long>(y) * x_map_width + tc] = dv;
    }
}

// ---------------------------------------------------------------------------------------------
// XM_FLAG_BILINEAR: bilinear X-map lookup (BASELINE config 3; NOT in the reference, whose lookup is nearest:
// x_maps_disparity.py:19,25 round both coordinates).  Defined here and restated by the oracle's
// frame_disparity_map_bilinear, float64 operation by operation:
//   (x_r, y_r) = float32 rectification LUT (rectify_cam_coords_f32, cam_proj_calibration.py:272-275)
//   c   = (t - t_min) / (t_max - t_min) * T_PX_SCALE            un-rounded (0 / 0 -> 0)
//   c0 = floor(c), fc = c - c0, y0 = floor(y_r), fy = y_r - y0;  needs 0 <= y0, y0 + 1 <= H - 1, 0 <= c0 <= T_PX_SCALE
//   taps X[y0, c0], X[y0, c1], X[y0 + 1, c0], X[y0 + 1, c1], c1 = min(c0 + 1, W - 1), weights (1 - fy)(1 - fc) ...;
//   taps equal to 0 are undefined cells (x_map.py:5-55) and are left out, the remaining weights renormalised
//   x_p = sum(w v) / sum(w);  disparity = float32((x_p - x_r) - X_OFFSET), inlier iff >= 0
//   scatter cell: projector view (rint(y_r), rint(x_p - X_OFFSET)), camera view (y, x); last event wins; the map holds
//   the float32 disparity.  Key = (event index + 1) << 32 | float bits, kept with atomicMax in a cleared map.
// ---------------------------------------------------------------------------------------------
struct BilinearParams {
    const int4* events;
    long long n;
    int polarity;
    const float* lut_x;
    const float* lut_y;
    int cam_w, cam_h;
    const short* xmap_t;
    int xmap_w, xmap_h, col_stride;
    int t_px_scale, x_offset;
    int rect_w, rect_h;
    int view;  // 0 projector, 1 camera
    unsigned long long* map;
    FrameState* state;
    int verify;
};

template <bool F64>
__global__ void __launch_bounds__(256) bilinear_scatter_kernel(const BilinearParams p) {
    TimeNorm<F64> tn;
    tn.init(p.state->t_lo_bits, p.state->t_hi_bits, p.t_px_scale);
    unsigned n_valid = 0, n_inl = 0, flags = 0;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const EventFields e = unpack_event(ld_event_plain(p.events + i));
        if (!event_valid(e, p.polarity)) continue;
        ++n_valid;
        if (e.x >= static_cast<unsigned>(p.cam_w) || e.y >= static_cast<unsigned>(p.cam_h)) {
            flags |= kStatusPixelOob;
            continue;
        }
        if (p.verify && tn.outside(e.t_bits)) flags |= kStatusTBounds;
        const int pix = static_cast<int>(e.y) * p.cam_w + static_cast<int>(e.x);
        const double xr = static_cast<double>(__ldg(p.lut_x + pix)), yr = static_cast<double>(__ldg(p.lut_y + pix));
        const double num_t = F64 ? __dsub_rn(__longlong_as_double(e.t_bits), tn.lo_f) : static_cast<double>(e.t_bits - tn.lo_i);
        double c = __dmul_rn(__ddiv_rn(num_t, tn.den), tn.scale);
        if (!(c == c)) c = 0.0;
        const double c0 = floor(c), y0 = floor(yr);
        if (!(y0 >= 0.0 && y0 + 1.0 <= static_cast<double>(p.xmap_h - 1) && c0 >= 0.0 && c0 <= static_cast<double>(p.t_px_scale))) continue;
        const double fc = __dsub_rn(c, c0), fy = __dsub_rn(yr, y0);
        const int ic0 = static_cast<int>(c0), iy0 = static_cast<int>(y0);
        const int ic1 = min(ic0 + 1, p.xmap_w - 1);
        const int v00 = __ldg(p.xmap_t + static_cast<long long>(ic0) * p.col_stride + iy0);
        const int v01 = __ldg(p.xmap_t + static_cast<long long>(ic1) * p.col_stride + iy0);
        const int v10 = __ldg(p.xmap_t + static_cast<long long>(ic0) * p.col_stride + iy0 + 1);
        const int v11 = __ldg(p.xmap_t + static_cast<long long>(ic1) * p.col_stride + iy0 + 1);
        const double gy = __dsub_rn(1.0, fy), gc = __dsub_rn(1.0, fc);
        const double w00 = __dmul_rn(gy, gc), w01 = __dmul_rn(gy, fc), w10 = __dmul_rn(fy, gc), w11 = __dmul_rn(fy, fc);
        double num = 0.0, den = 0.0;
        if (v00) { num = __dadd_rn(num, __dmul_rn(w00, static_cast<double>(v00))); den = __dadd_rn(den, w00); }
        if (v01) { num = __dadd_rn(num, __dmul_rn(w01, static_cast<double>(v01))); den = __dadd_rn(den, w01); }
        if (v10) { num = __dadd_rn(num, __dmul_rn(w10, static_cast<double>(v10))); den = __dadd_rn(den, w10); }
        if (v11) { num = __dadd_rn(num, __dmul_rn(w11, static_cast<double>(v11))); den = __dadd_rn(den, w11); }
        if (!(den > 0.0)) continue;
        const double xp = __ddiv_rn(num, den);
        const float disp = __double2float_rn(__dsub_rn(__dsub_rn(xp, xr), static_cast<double>(p.x_offset)));
        if (!(disp >= 0.0f)) continue;
        long long cell;
        if (p.view == 1) {
            cell = pix;
        } else {
            const double row = rint(yr), col = rint(__dsub_rn(xp, static_cast<double>(p.x_offset)));
            if (!(row >= 0.0 && row < static_cast<double>(p.rect_h) && col >= 0.0 && col < static_cast<double>(p.rect_w))) {
                flags |= kStatusScatterOob;
                continue;
            }
            cell = static_cast<long long>(row) * p.rect_w + static_cast<long long>(col);
        }
        ++n_inl;
        atomicMax(p.map + cell, (static_cast<unsigned long long>(i + 1) << 32) | static_cast<unsigned>(__float_as_int(disp)));
    }
    n_valid = __reduce_add_sync(0xffffffffu, n_valid);
    n_inl = __reduce_add_sync(0xffffffffu, n_inl);
    flags = __reduce_or_sync(0xffffffffu, flags);
    if ((threadIdx.x & 31) == 0) {
        if (n_valid) atomicAdd(&p.state->n_valid, static_cast<unsigned long long>(n_valid));
        if (n_inl) atomicAdd(&p.state->n_inliers, static_cast<unsigned long long>(n_inl));
        if (flags) atomicOr(&p.state->flags, flags);
    }
}

// float32 disparity map out of the bilinear keys; clears the map for the next frame on the way
__global__ void __launch_bounds__(256) bilinear_decode_kernel(unsigned long long* __restrict__ map, long long n, float* __restrict__ out) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const unsigned long long k = map[i];
        out[i] = k ? __int_as_float(static_cast<int>(k & 0xffffffffULL)) : 0.f;
        if (k) map[i] = 0ULL;
    }
}

// ---------------------------------------------------------------------------------------------
// N3 (second half): initUndistortRectifyMapInverse (python/cam_proj_calibration.py:31-41, used at :246-270)
// = cv2.undistortPoints over the whole pixel grid.  OpenCV's point loop (cvUndistortPointsInternal, default
// criteria = 5 iterations, no tilt model) restated operation by operation in float64 without contraction, so the
// float32 maps -- and the int16 tables rounded from them -- are bit-identical to the host's:
//   x = (u - cx) / fx (as * (1 / fx)), y likewise; 5 x { r2; icdist = (1 + ((k7 r2 + k6) r2 + k5) r2) / (1 + ((k4 r2 +
//   k1) r2 + k0) r2); dX = 2 k2 x y + k3 (r2 + 2 x x) + k8 r2 + k9 r2 r2; dY likewise; x = (x0 - dX) icdist; ... };
//   [xx yy ww] = RR [x y 1], RR = P[:, :3] R (passed in, computed by the caller the way OpenCV does); out = xx / ww.
// ---------------------------------------------------------------------------------------------
struct InverseLutParams {
    double ifx, ify, cx, cy;
    double k[12];   // k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 (OpenCV order)
    double rr[9];   // row-major
    int w, h;
    int has_dist;
};

__global__ void __launch_bounds__(256) build_inverse_lut_kernel(const InverseLutParams p, float* __restrict__ mapx, float* __restrict__ mapy,
                                                                short* __restrict__ xy_i16, unsigned* __restrict__ overflow) {
    const long long n = static_cast<long long>(p.w) * p.h;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(i / p.w), u = static_cast<int>(i - static_cast<long long>(v) * p.w);
        double x = __dmul_rn(__dsub_rn(static_cast<double>(u), p.cx), p.ifx);
        double y = __dmul_rn(__dsub_rn(static_cast<double>(v), p.cy), p.ify);
        const double x00 = x, y00 = y;
        if (p.has_dist) {
            const double x0 = x, y0 = y;
            const double* k = p.k;
#pragma unroll 1
            for (int j = 0; j < 5; ++j) {
                const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
                const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[7], r2), k[6]), r2), k[5]), r2));
                const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[4], r2), k[1]), r2), k[0]), r2));
                const double icdist = __ddiv_rn(num, den);
                if (icdist < 0) {  // (OpenCV: give up and return the undistorted start value)
                    x = x00;
                    y = y00;
                    break;
                }
                const double two_k2 = __dmul_rn(2.0, k[2]), two_k3 = __dmul_rn(2.0, k[3]);
                const double dx = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(two_k2, x), y),
                                                                  __dmul_rn(k[3], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x)))),
                                                        __dmul_rn(k[8], r2)),
                                              __dmul_rn(__dmul_rn(k[9], r2), r2));
                const double dy = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(k[2], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
                                                                  __dmul_rn(__dmul_rn(two_k3, x), y)),
                                                        __dmul_rn(k[10], r2)),
                                              __dmul_rn(__dmul_rn(k[11], r2), r2));
                x = __dmul_rn(__dsub_rn(x0, dx), icdist);
                y = __dmul_rn(__dsub_rn(y0, dy), icdist);
            }
        }
        const double xx = __dadd_rn(__dadd_rn(__dmul_rn(p.rr[0], x), __dmul_rn(p.rr[1], y)), p.rr[2]);
        const double yy = __dadd_rn(__dadd_rn(__dmul_rn(p.rr[3], x), __dmul_rn(p.rr[4], y)), p.rr[5]);
        const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(p.rr[6], x), __dmul_rn(p.rr[7], y)), p.rr[8]));
        const float fx = __double2float_rn(__dmul_rn(xx, ww)), fy = __double2float_rn(__dmul_rn(yy, ww));
        if (mapx) mapx[i] = fx;
        if (mapy) mapy[i] = fy;
        if (xy_i16) {  // mapf_to_i16 (:44-48): np.rint (half-even), range-checked
            const float rx = rintf(fx), ry = rintf(fy);
            if (!(rx >= -32768.0f && rx <= 32767.0f && ry >= -32768.0f && ry <= 32767.0f)) *overflow = 1u;
            xy_i16[2 * i] = static_cast<short>(static_cast<int>(rx));
            xy_i16[2 * i + 1] = static_cast<short>(static_cast<int>(ry));
        }
    }
}

}  // namespace xm
