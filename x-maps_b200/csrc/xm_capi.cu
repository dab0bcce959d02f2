// xm_capi.cu — C ABI of the B200-native X-maps depth path (include/xmaps_b200.h).
//
// Plain pointers and sizes only; no torch types.  The context owns device copies of the
// calibration tables (re-packed for the kernels), the 64-bit scatter map and a small state block;
// event buffers and outputs belong to the caller.  Every entry point returns a status code and
// records a thread-local message for xm_last_error().
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/xmaps_b200.h"
#include "xm_stage_kernels.cuh"
#include "xm_fused_kernel.cuh"
#include "xm_batch_kernel.cuh"

#ifndef XM_BATCH_SMEM_CAP  // shared memory the resident batch-kernel CTAs of one SM may take together (above 196 KB the L1 shrinks to 28 KB)
#define XM_BATCH_SMEM_CAP (196 * 1024)
#endif
#include "xm_stream_kernels.cuh"

namespace {

thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define XM_CUDA(expr)                                                                                       \
    do {                                                                                                    \
        cudaError_t e__ = (expr);                                                                           \
        if (e__ != cudaSuccess) return fail(XM_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

#define XM_LAUNCHED()                                                                                   \
    do {                                                                                                \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                             \
        cudaError_t e__ = cudaGetLastError();                                                           \
        if (e__ != cudaSuccess) return fail(XM_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e__)); \
    } while (0)

// 256-entry BGR table of OpenCV's COLORMAP_TURBO is supplied by the host side at context creation
// (xm_ctx_set_colormap); until then XM_OUT_BGR is refused.

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace

struct XmCtx {
    int device = 0;
    int sm_count = 148;
    int cam_w = 0, cam_h = 0, rect_w = 0, rect_h = 0, proj_w = 0, proj_h = 0;
    int xmap_w = 0, xmap_h = 0, col_stride = 0, t_px_scale = 0, x_offset = 0, dilate = 7;
    double depth_scale = 0.0;
    int* d_lut_xy = nullptr;
    std::vector<int> h_lut_xy;          // host copy of the packed LUT (the alive bitmap is rebuilt when the X-map changes)
    unsigned* d_alive = nullptr;        // the batch kernel's alive table: per camera-pixel block the time columns that can yield inliers (build_alive)
    unsigned* d_alive_ones = nullptr;   // the same table with every column alive everywhere (option alive = 0)
    int alive_shift = 3, alive_pitch = 0, alive_qs = 0, alive_entries = 0, alive_words = 0;  // block = 2^shift pixels, entries per row, column quantum 2^qs, u16 entries, 32-bit words
    long long alive_px = 0;             // camera pixels inside alive blocks (diagnostic, option "alive_px")
    int opt_alive = 1;
    int lut_x_min = 0;                  // smallest rectified x of the LUT (lut_safe = lut_x_min > -x_offset)
    float* d_lut_x_f32 = nullptr;
    float* d_lut_y_f32 = nullptr;
    unsigned long long* d_bil_map = nullptr;  // XM_FLAG_BILINEAR: key map (kept cleared between frames), float rect / out maps
    float* d_bil_rect = nullptr;
    float* d_bil_out = nullptr;
    short* d_xmap_t = nullptr;
    short2* d_remap_xy = nullptr;
    short4* d_tile_box = nullptr;  // bounding box of the remap targets of every 32x32 output tile
    unsigned short* d_tile_off = nullptr;  // per output pixel: its cell inside the tile's shared-memory region (EpilogueParams::tile_off)
    int opt_tile_off = 1;
    // strip epilogue of the batch kernel (projector view): per-pixel cell index table and the dilated-map ring
    unsigned* d_pix_cell = nullptr;
    unsigned short* d_dil = nullptr;  // kBatchDilMaps x rect_w x rect_h
    int strip_box[4] = {0, 0, 0, 0};  // bounding box of the remap targets (inclusive): the window pass 1 produces
    int opt_tile_warps = 0;  // epilogue warps per CTA of the strip epilogue: 0 = auto, 2 or 4
    int opt_strip_lag = 1;  // (2, 3 measured: no difference, EXPERIMENTS_r02.md #36)
    int opt_strip_rows = 0, opt_strip_blocks = 0;  // item sizes of the strip epilogue (0 = auto: by events per frame)
    int opt_batch_strips = 1;
    int opt_scatter_aggregate = 2;  // batch kernel: 0 plain scatter, 1 aggregate dense chunks per warp round, 2 auto (from the last batch's inlier fraction)
    bool agg_dense = false, agg_pending = false;
    cudaEvent_t agg_event = nullptr;
    unsigned long long* h_agg_stats = nullptr;  // pinned: n_valid, n_inliers of the last sampled batch frame
    unsigned char* d_turbo = nullptr;
    unsigned long long* d_dbg = nullptr;  // per-CTA phase timestamps (option debug & 8)
    float* d_depth_lut = nullptr;  // [32768], exact depth of every integer disparity
    bool have_turbo = false;
    unsigned long long* d_map = nullptr;
    long long map_cells = 0;
    xm::FrameState* d_state = nullptr;  // ring of kStateSlots blocks; frame f uses slot f % kStateSlots
    unsigned long long frame_no = 0;    // operations that used a state slot so far
    int last_slot = 0;                  // slot of the most recent frame (xm_frame_status)
    bool slot_dirty[4] = {false, false, false, false};
    bool prev_was_frame = false;        // the previous launch on the stream was a fused frame's epilogue
    unsigned epoch = 0;
    // compaction scratch
    unsigned* d_counts = nullptr;
    long long counts_cap = 0;
    // filters / trigger finder scratch
    int lut_x_max = 0;                 // largest rectified x of the LUT: bounds the YT filter's key image
    unsigned* d_filter_first = nullptr;  // [cam_h * max(cam_w, lut_x_max + 1)]
    unsigned* d_filter_last = nullptr;
    unsigned* d_pause_idx = nullptr;
    long long pause_cap = 0;
    void* d_stream_scratch = nullptr;  // TriggerScratch + xp_max
    unsigned* d_act_first = nullptr;    // [cam_h * cam_w] activity filter: first event index of the sub-packet per pixel
    long long* d_act_last = nullptr;    // [cam_h * cam_w] ... latest timestamp per pixel, carried across packets (starts at 0)
    long long* d_act_plan = nullptr;    // [kActPlanCap + 1] sub-packet bounds, then: count (int), unsorted flag (unsigned), running total, part total
    // staging for xm_frame_host
    void* d_stage_ev = nullptr;
    long long stage_ev_cap = 0;
    void* d_stage_out = nullptr;
    long long stage_out_cap = 0;
    // options
    int opt_stage_xmap = 1;
    int opt_auto_fixup = 1;
    int opt_lookahead = 1;
    int opt_ctas_per_sm = 0;  // 0 = occupancy query
    int opt_smem_cols_bytes = 12 * 1024;
    int opt_k2_variant = 1;   // 1: sliding-window epilogue (7x7, even rect_w), 0: per-tap epilogue
    int opt_reserve_sms = 0;  // SMs the persistent batch kernel leaves free (room for NCCL's copy kernels next to it)
    int opt_coop = 1;         // 1: frame_kernel (grid barrier inside) is launched cooperatively
    int opt_fused = 1;        // 1: one fused kernel per frame where the lean path applies
    int fused_occ = 0;        // resident CTAs per SM of frame_kernel
    int opt_batch = 1;        // 1: xm_frame_batch renders uniform batches with one persistent kernel per <= 32 frames
                              //    (event warps + dedicated epilogue warps; 47 vs 57 us per 5 M-event frame, EXPERIMENTS_r01.md)
    int batch_occ = 0, batch_smem[2] = {0, 0}, batch_cols[2] = {0, 0};  // launch configuration of batch_kernel ([view])
    unsigned long long* d_map_ring[xm::kBatchMapsMax] = {};  // [0] = d_map; the others are allocated when a batch first uses them
    int opt_batch_maps = 0;  // scatter maps in rotation (0 = auto: kBatchMaps, more for batches of small frames)
    xm::FrameState* d_bstate = nullptr;  // [kBatchMax + 1] state blocks of the current batch
    const xm::FrameState* status_src = nullptr;  // state block xm_frame_status reports (NULL: d_state + last_slot)
    int opt_pdl = 1;          // programmatic dependent launch between K1 / K2 / next K1
    int opt_safe_tables = 1;  // use the check-free scatter when the tables were verified
    int opt_stages = 2;  // depth of the shared-memory event ring of K1
    int opt_debug = 0;        // timing experiments only
    int opt_win_stages = 2;   // depth of the X-map window ring (warp-specialised K1)
    int opt_batch_win_stages = 3;  // ... of the batch kernel (its back half runs a whole front half behind: 2 -> 3 stages = -0.7 us per 5 M-event frame)
    int opt_k1_variant = 2;   // 2: lean warp-specialised K1 (integer time, verified tables; else falls back to 1),
                              // 1: warp-specialised K1 (mbarrier pipelines), 0: block-barrier K1
    int opt_region_cells = 48 * 64;  // per buffer; replaced at creation by the largest tile region of the remap table
    long long region_need = 0;
    // per-kernel CUDA-event timing (option "profile"): pairs around K1 and K2 of every frame
    int opt_profile = 0;
    std::vector<cudaEvent_t> prof_events;  // 3 per frame: before K1, after K1 (+ fix-up), after K2
    size_t prof_used = 0;
    double prof_k1_ms = 0.0, prof_k2_ms = 0.0;
    long long prof_frames = 0;
    long long prof_batch_extra = 0;  // frames rendered by batch launches beyond one per event triple
    // derived
    int ev_occ_i64 = 0, ev_occ_f64 = 0;
    int ev_smem = 0, cap_cols = 0;
    bool lut_safe = false, xmap_safe = false;  // table ranges verified: scatter targets always inside the map
};

namespace {

constexpr int kStateSlots = 4;

// Picks the state block of the next operation.  Fused frames rely on the block having been cleared
// by the epilogue of the frame before the previous one; anything else (first frames, stage-by-stage
// calls in between) is cleared here.
int acquire_state_slot(XmCtx* c, cudaStream_t s, bool need_clean, cudaError_t* err) {
    *err = cudaSuccess;
    const int slot = static_cast<int>(c->frame_no % kStateSlots);
    c->frame_no += 1;
    if (need_clean && c->slot_dirty[slot]) {
        *err = cudaMemsetAsync(c->d_state + slot, 0, sizeof(xm::FrameState), s);
        c->slot_dirty[slot] = false;
        c->prev_was_frame = false;  // a memset sits between the kernels: no programmatic dependency
    }
    c->last_slot = slot;
    return slot;
}

// state block for a stage-by-stage call (cleared here unless the callee resets it itself)
int staged_state(XmCtx* c, cudaStream_t s, bool need_clean, xm::FrameState** out) {
    cudaError_t err;
    c->status_src = nullptr;
    const int slot = acquire_state_slot(c, s, need_clean, &err);
    XM_CUDA(err);
    c->slot_dirty[slot] = true;
    c->prev_was_frame = false;
    *out = c->d_state + slot;
    return XM_OK;
}

unsigned next_epoch(XmCtx* c, unsigned count, cudaStream_t s, cudaError_t* err) {
    // epochs live in the top 16 bits of a key; 0 means "never written".  On wrap the map is cleared.
    *err = cudaSuccess;
    if (c->epoch + count > 0xffffu) {
        *err = cudaMemsetAsync(c->d_map, 0, static_cast<size_t>(c->map_cells) * 8, s);
        for (int i = 1; i < xm::kBatchMapsMax && *err == cudaSuccess; ++i)
            if (c->d_map_ring[i]) *err = cudaMemsetAsync(c->d_map_ring[i], 0, static_cast<size_t>(c->map_cells) * 8, s);
        c->epoch = 0;
    }
    unsigned e = c->epoch + 1;
    c->epoch += count;
    return e;
}

size_t fa_static_bytes(void (*k)(xm::BatchParams)) {
    cudaFuncAttributes fa;
    return cudaFuncGetAttributes(&fa, k) == cudaSuccess ? fa.sharedSizeBytes : 0;
}

using EvKernel = void (*)(xm::EventParams);
EvKernel ev_kernel(bool f64, bool safe, int variant = 1, bool cam = false) {
    // 2: the lean integer-time kernel (verified tables); float64 timestamps and unverified tables take the general one
    if (variant == 2 && !f64 && safe) return cam ? xm::events_lean_kernel<true> : xm::events_lean_kernel<false>;
    if (f64) return safe ? xm::events_ws_kernel<true, true> : xm::events_ws_kernel<true, false>;
    return safe ? xm::events_ws_kernel<false, true> : xm::events_ws_kernel<false, false>;
}

// Launch with (optionally) the programmatic-stream-serialization attribute: the kernel may start
// while its predecessor in the stream is still running and synchronises with griddepcontrol.wait.
// `coop`: cooperative launch -- the runtime guarantees that every CTA of the grid is resident at the same time (or
// rejects the launch), which a kernel with a hand-rolled grid barrier needs when other streams, NCCL or a second
// context share the GPU.
template <typename P>
cudaError_t launch_pdl(void (*kernel)(P), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, const P& params, bool coop = false) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (coop) {
        attr[n].id = cudaLaunchAttributeCooperative;
        attr[n].val.cooperative = 1;
        ++n;
    }
    cfg.attrs = n ? attr : nullptr;
    cfg.numAttrs = n;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, params);
    if (e != cudaSuccess && pdl && coop) {
        // the two attributes together are refused by some drivers: keep the co-residency guarantee, drop the overlap
        cudaGetLastError();
        cfg.attrs = attr + 1;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kernel, params);
    }
    return e;
}

// shared-memory tile regions of the batch kernel's epilogue groups: none for the camera view and the strip epilogue
int batch_region_cells(const XmCtx* c, bool cam) { return (cam || (c->opt_batch_strips && c->d_pix_cell)) ? 0 : c->opt_region_cells; }

int configure_event_kernels(XmCtx* c) {
    int cols = 0;
    if (c->opt_stage_xmap && c->col_stride > 0) cols = c->opt_smem_cols_bytes / (c->col_stride * 2);
    if (cols > c->xmap_w) cols = c->xmap_w;
    c->cap_cols = cols;
    const int win_bytes = cols * c->col_stride * 2;
    const int variant = c->opt_k1_variant;
    const int threads = xm::kWsThreads;
    c->ev_smem = xm::events_ws_smem_bytes(c->opt_stages, c->opt_win_stages, win_bytes);
    // the attribute is per function, not per context: always allow the device maximum so that
    // contexts with different X-map geometries can coexist in one process
    int optin = 0;
    XM_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    c->ev_occ_i64 = c->ev_occ_f64 = 1 << 30;
    for (int f64 = 0; f64 < 2; ++f64)
        for (int safe = 0; safe < 3; ++safe) {  // safe == 2: the second (camera-view) lean instantiation
            EvKernel k = ev_kernel(f64 != 0, safe != 0, variant, safe == 2);
            cudaFuncAttributes fa;
            XM_CUDA(cudaFuncGetAttributes(&fa, k));
            const int dyn = optin - static_cast<int>(fa.sharedSizeBytes);
            if (c->ev_smem > dyn) return fail(XM_ERR_UNSUPPORTED, "event kernel needs %d B of shared memory, device allows %d", c->ev_smem, dyn);
            XM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
            int occ = 0;
            XM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, c->ev_smem));
            int& dst = f64 ? c->ev_occ_f64 : c->ev_occ_i64;
            dst = occ < dst ? occ : dst;
        }
    if (c->ev_occ_i64 < 1 || c->ev_occ_f64 < 1) return fail(XM_ERR_UNSUPPORTED, "event kernel does not fit an SM (smem %d B)", c->ev_smem);
    c->fused_occ = 0;
    if (variant == 2) {
        int occ_min = 1 << 30;
        for (int cam = 0; cam < 2; ++cam) {
            void (*k)(xm::FrameParams) = cam ? xm::frame_kernel<true> : xm::frame_kernel<false>;
            cudaFuncAttributes fa;
            XM_CUDA(cudaFuncGetAttributes(&fa, k));
            const int dyn = optin - static_cast<int>(fa.sharedSizeBytes);
            if (c->ev_smem > dyn) {
                occ_min = 0;
                break;
            }
            XM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
            int occ = 0;
            XM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, xm::kWsThreads, c->ev_smem));
            occ_min = occ < occ_min ? occ : occ_min;
        }
        c->fused_occ = occ_min;
    }
    // batch kernel: the same pipeline plus epilogue warp groups with their own tile regions, 2 CTAs per SM
    // (416 threads x 72 registers); keep them below the 196 KB shared-memory carve-out (above it the L1
    // shrinks to 28 KB and the LUT gathers slow down by 1.5x, EXPERIMENTS_r01.md)
    c->batch_occ = 0;
    if (variant == 2) {
        int occ_min = 1 << 30;
        for (int cam = 0; cam < 2; ++cam) {
            int bcols = cols;
            auto smem_for = [&](int k) {
                return xm::batch_smem_bytes(c->opt_stages, c->opt_batch_win_stages, k * c->col_stride * 2, batch_region_cells(c, cam != 0), c->alive_words, cam != 0);
            };
            while (bcols > 0 && xm::kBatchCtasPerSm * (smem_for(bcols) + 1024) > XM_BATCH_SMEM_CAP) --bcols;
            c->batch_cols[cam] = bcols;
            c->batch_smem[cam] = smem_for(bcols);
            void (*k)(xm::BatchParams) = cam ? xm::batch_kernel<true> : xm::batch_kernel<false>;
            XM_CUDA(cudaFuncSetAttribute(xm::batch_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         optin - static_cast<int>(fa_static_bytes(xm::batch_kernel<false, true>))));
            XM_CUDA(cudaFuncSetAttribute(xm::batch_kernel<false, false, xm::kTileWarpsLarge>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         optin - static_cast<int>(fa_static_bytes(xm::batch_kernel<false, false, xm::kTileWarpsLarge>))));
            XM_CUDA(cudaFuncSetAttribute(xm::batch_kernel<false, true, xm::kTileWarpsLarge>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         optin - static_cast<int>(fa_static_bytes(xm::batch_kernel<false, true, xm::kTileWarpsLarge>))));
            cudaFuncAttributes fa;
            XM_CUDA(cudaFuncGetAttributes(&fa, k));
            const int dyn = optin - static_cast<int>(fa.sharedSizeBytes);
            if (c->batch_smem[cam] > dyn) {
                occ_min = 0;
                break;
            }
            XM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
            int occ = 0;
            XM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, xm::kBatchThreads, c->batch_smem[cam]));
            occ_min = occ < occ_min ? occ : occ_min;
        }
        c->batch_occ = occ_min;
    }
    return XM_OK;
}

// The batch kernel's "alive" table (BatchParams::alive): per block of 2^s x 2^s camera pixels the hull of the time
// columns at which an event of a pixel of the block can be an inlier.  Exactly the reference's conditions
// (x_maps_disparity.py:23-30) evaluated per pixel against its X-map row: 0 <= y_rect < rows - 1 and
// int16(x_map[y_rect, col] - x_rect - X_OFFSET) >= 0; first and last such column per pixel (for a monotonic row every
// column between them is an inlier too; for any other table the hull is merely conservative), union over the block,
// quantised outwards to 2^qs columns.
int build_alive(XmCtx* c, const int16_t* h_x_map, int rows, int cols, int x_offset) {
    int s = 3;  // smallest block (>= 8 x 8 pixels) whose table stays below 12 KB of shared memory
    auto blocks = [&](int dim) { return (dim + (1 << s) - 1) >> s; };
    while (static_cast<long long>(blocks(c->cam_w)) * blocks(c->cam_h) > 6144) ++s;
    const int bw = blocks(c->cam_w), bh = blocks(c->cam_h), n = bw * bh;
    int qs = 0;  // (no column may reach 255 units: that value marks dead blocks)
    while (((cols - 1) >> qs) > 254) ++qs;
    std::vector<int> lo(n, 1 << 30), hi(n, -1);
    for (int y = 0; y < c->cam_h; ++y)
        for (int x = 0; x < c->cam_w; ++x) {
            const int packed = c->h_lut_xy[static_cast<size_t>(y) * c->cam_w + x];
            const int xcr = static_cast<short>(packed & 0xffff), ycr = packed >> 16;
            if (ycr < 0 || ycr >= rows - 1) continue;
            const int16_t* row = h_x_map + static_cast<size_t>(ycr) * cols;
            const int k = xcr + x_offset;
            int first = 0;
            while (first < cols && static_cast<short>(row[first] - k) < 0) ++first;
            if (first == cols) continue;  // no column makes this pixel an inlier
            int last = cols - 1;
            while (static_cast<short>(row[last] - k) < 0) --last;
            const int b = (y >> s) * bw + (x >> s);
            lo[b] = first < lo[b] ? first : lo[b];
            hi[b] = last > hi[b] ? last : hi[b];
        }
    const int words = (n * 2 + 3) / 4;
    std::vector<unsigned short> tab(static_cast<size_t>(words) * 2, 255u), ones(static_cast<size_t>(words) * 2, 0xff00u);
    long long alive_px = 0;
    for (int by = 0; by < bh; ++by)
        for (int bx = 0; bx < bw; ++bx) {
            const int b = by * bw + bx;
            if (hi[b] < 0) continue;
            const int lq = lo[b] >> qs, hq = hi[b] >> qs;
            tab[b] = static_cast<unsigned short>(lq | ((hq - lq) << 8));
            const int w = std::min(1 << s, c->cam_w - (bx << s)), h = std::min(1 << s, c->cam_h - (by << s));
            alive_px += static_cast<long long>(w) * h;
        }
    if (c->alive_words != words) {
        cudaFree(c->d_alive);
        cudaFree(c->d_alive_ones);
        c->d_alive = c->d_alive_ones = nullptr;
        XM_CUDA(cudaMalloc(&c->d_alive, words * sizeof(unsigned)));
        XM_CUDA(cudaMalloc(&c->d_alive_ones, words * sizeof(unsigned)));
    }
    XM_CUDA(cudaMemcpy(c->d_alive, tab.data(), words * sizeof(unsigned), cudaMemcpyHostToDevice));
    XM_CUDA(cudaMemcpy(c->d_alive_ones, ones.data(), words * sizeof(unsigned), cudaMemcpyHostToDevice));
    c->alive_shift = s;
    c->alive_pitch = bw;
    c->alive_qs = qs;
    c->alive_entries = n;
    c->alive_words = words;
    c->alive_px = alive_px;
    return XM_OK;
}

int upload_xmap(XmCtx* c, const int16_t* h_x_map, int rows, int cols, int t_px_scale, int x_offset) {
    if (!h_x_map || rows <= 0 || cols <= 0) return fail(XM_ERR_INVALID_ARG, "x_map: bad shape %d x %d", rows, cols);
    // asserts of the reference (x_maps_disparity.py:52-53)
    if (rows > 32767) return fail(XM_ERR_TABLE_RANGE, "x_map has %d rows, int16 indices allow 32767", rows);
    if (cols > 32767) return fail(XM_ERR_TABLE_RANGE, "x_map has %d time columns, int16 indices allow 32767", cols);
    // The kernels index the table with the time column rint(t_norm * t_px_scale) in [0, t_px_scale] without a
    // bounds test; the reference would raise IndexError for a column >= cols (x_maps_disparity.py:25).
    if (t_px_scale <= 0 || t_px_scale >= cols)
        return fail(XM_ERR_INVALID_ARG, "t_px_scale = %d must lie in [1, %d) (time columns of the X-map)", t_px_scale, cols);
    if (x_offset <= 0 || x_offset > 32767) return fail(XM_ERR_INVALID_ARG, "x_offset = %d must lie in [1, 32767]", x_offset);
    // With every defined X-map cell in [x_offset, x_offset + rect_w) (and LUT x > -x_offset, checked at
    // creation, so that undefined cells can never produce disp >= 0) an inlier's scatter target
    // x_rect + disp = x_map - x_offset is always inside the map: K1 may skip the checks.
    bool safe = rows <= c->rect_h;
    for (size_t i = 0, e = static_cast<size_t>(rows) * cols; i < e && safe; ++i) {
        const int v = h_x_map[i];
        safe = v == 0 || (v >= x_offset && v < x_offset + c->rect_w);
    }
    c->xmap_safe = safe;
    // (re)derive the LUT's half of the check-free scatter condition for THIS x_offset: undefined cells (0) must
    // never produce disp >= 0, i.e. every rectified x > -x_offset
    c->lut_safe = c->lut_x_min > -x_offset;
    {
        int rc = build_alive(c, h_x_map, rows, cols, x_offset);
        if (rc) return rc;
    }
    const int stride = (rows + 7) & ~7;  // 16-byte multiple: one column range = one bulk copy
    std::vector<short> t(static_cast<size_t>(cols) * stride, 0);
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) t[static_cast<size_t>(x) * stride + y] = h_x_map[static_cast<size_t>(y) * cols + x];
    if (c->d_xmap_t) cudaFree(c->d_xmap_t);
    c->d_xmap_t = nullptr;
    XM_CUDA(cudaMalloc(&c->d_xmap_t, t.size() * sizeof(short)));
    XM_CUDA(cudaMemcpy(c->d_xmap_t, t.data(), t.size() * sizeof(short), cudaMemcpyHostToDevice));
    c->xmap_w = cols;
    c->xmap_h = rows;
    c->col_stride = stride;
    c->t_px_scale = t_px_scale;
    c->x_offset = x_offset;
    return configure_event_kernels(c);
}

int grid_for(long long n, int threads, int per_thread, int max_blocks) {
    long long b = (n + static_cast<long long>(threads) * per_thread - 1) / (static_cast<long long>(threads) * per_thread);
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return static_cast<int>(b);
}

xm::OutputSpec make_output(const XmCtx* c, int kind, double depth_scale, float z_near, float z_far) {
    xm::OutputSpec o;
    o.kind = kind;
    o.depth_scale = depth_scale;
    o.z_near = z_near;
    o.z_far = z_far;
    o.turbo_bgr = c->d_turbo;
    o.depth_lut = (depth_scale == c->depth_scale) ? c->d_depth_lut : nullptr;
    return o;
}

// Launches K0 for one frame.  `epoch` is the epoch K1 will use.
int launch_bounds(XmCtx* c, xm::FrameState* st, const void* d_events, long long n, uint32_t flags, int time_bounds, long long t_min,
                  long long t_max, unsigned epoch, cudaStream_t s) {
    const int polarity = (flags & XM_FLAG_POLARITY) ? 1 : 0;
    const bool f64 = (flags & XM_FLAG_TIME_F64) != 0;
    const int4* ev = static_cast<const int4*>(d_events);
    const int mode = time_bounds == XM_TBOUNDS_SORTED ? 0 : 1;
    xm::bounds_init_kernel<<<1, 64, 0, s>>>(ev, n, polarity, mode, t_min, t_max, epoch, st);
    XM_LAUNCHED();
    if (time_bounds == XM_TBOUNDS_REDUCE) {
        int grid = grid_for(n, 256, 8, c->sm_count * 8);
        if (f64)
            xm::bounds_reduce_kernel<true><<<grid, 256, 0, s>>>(ev, n, polarity, st);
        else
            xm::bounds_reduce_kernel<false><<<grid, 256, 0, s>>>(ev, n, polarity, st);
        XM_LAUNCHED();
    }
    return XM_OK;
}

int check_frame_args(const XmCtx* c, const XmFrameArgs* a) {
    if (!c || !a) return fail(XM_ERR_INVALID_ARG, "null context or arguments");
    if (a->n_events < 0) return fail(XM_ERR_INVALID_ARG, "n_events = %lld", static_cast<long long>(a->n_events));
    if (a->n_events > 0xffffffffLL) return fail(XM_ERR_UNSUPPORTED, "n_events = %lld exceeds 2^32 - 1 per frame", static_cast<long long>(a->n_events));
    if (a->n_events > 0 && !a->d_events) return fail(XM_ERR_INVALID_ARG, "d_events is NULL");
    if (reinterpret_cast<uintptr_t>(a->d_events) & 15) return fail(XM_ERR_INVALID_ARG, "d_events must be 16-byte aligned");
    if (a->view != XM_VIEW_PROJECTOR && a->view != XM_VIEW_CAMERA) return fail(XM_ERR_INVALID_ARG, "unknown view %d", a->view);
    if (a->time_bounds < XM_TBOUNDS_REDUCE || a->time_bounds > XM_TBOUNDS_GIVEN) return fail(XM_ERR_INVALID_ARG, "unknown time_bounds %d", a->time_bounds);
    if (a->output < XM_OUT_DEPTH || a->output > XM_OUT_BGR) return fail(XM_ERR_INVALID_ARG, "unknown output %d", a->output);
    if (a->flags & ~(XM_FLAG_POLARITY | XM_FLAG_TIME_F64 | XM_FLAG_BILINEAR)) return fail(XM_ERR_INVALID_ARG, "unknown flags 0x%x", a->flags);
    if ((a->flags & XM_FLAG_BILINEAR) && !c->d_lut_x_f32) return fail(XM_ERR_INVALID_ARG, "XM_FLAG_BILINEAR needs XmTables.lut_x_f32 / lut_y_f32");
    if (!c->d_xmap_t) return fail(XM_ERR_NO_XMAP, "no X-map: supply XmTables.x_map or call xm_ctx_set_xmap");
    if (a->view == XM_VIEW_PROJECTOR && !c->d_remap_xy) return fail(XM_ERR_INVALID_ARG, "projector view needs XmTables.remap_xy");
    if (a->output == XM_OUT_BGR && !c->have_turbo) return fail(XM_ERR_INVALID_ARG, "XM_OUT_BGR needs a colour map (xm_ctx_set_colormap)");
    return XM_OK;
}

// profile support: drain recorded event triples into the accumulators (synchronises)
int profile_drain(XmCtx* c) {
    for (size_t i = 0; i + 2 < c->prof_used; i += 3) {
        float a = 0.f, b = 0.f;
        XM_CUDA(cudaEventSynchronize(c->prof_events[i + 2]));
        XM_CUDA(cudaEventElapsedTime(&a, c->prof_events[i], c->prof_events[i + 1]));
        XM_CUDA(cudaEventElapsedTime(&b, c->prof_events[i + 1], c->prof_events[i + 2]));
        c->prof_k1_ms += a;
        c->prof_k2_ms += b;
        c->prof_frames += 1;
    }
    c->prof_used = 0;
    return XM_OK;
}

int profile_mark(XmCtx* c, cudaStream_t s) {
    if (c->prof_used == c->prof_events.size()) {
        if (c->prof_events.size() >= 3 * 4096) {
            int rc = profile_drain(c);
            if (rc) return rc;
        } else {
            for (int i = 0; i < 3 * 64; ++i) {
                cudaEvent_t e;
                XM_CUDA(cudaEventCreate(&e));
                c->prof_events.push_back(e);
            }
        }
    }
    XM_CUDA(cudaEventRecord(c->prof_events[c->prof_used++], s));
    c->prev_was_frame = false;  // an event record sits between the kernels
    return XM_OK;
}

// XM_FLAG_BILINEAR: bounds, bilinear scatter of float disparities, decode, [dilate + remap], convert
int bilinear_impl(XmCtx* c, const XmFrameArgs* a, cudaStream_t s) {
    const bool cam = a->view == XM_VIEW_CAMERA;
    const long long cells = c->map_cells;
    const long long out_px = cam ? static_cast<long long>(c->cam_w) * c->cam_h : static_cast<long long>(c->proj_w) * c->proj_h;
    if (!c->d_bil_map) {
        XM_CUDA(cudaMalloc(&c->d_bil_map, static_cast<size_t>(cells) * 8));
        XM_CUDA(cudaMemset(c->d_bil_map, 0, static_cast<size_t>(cells) * 8));
        XM_CUDA(cudaMalloc(&c->d_bil_rect, static_cast<size_t>(cells) * 4));
        const long long mx = std::max(static_cast<long long>(c->cam_w) * c->cam_h, static_cast<long long>(c->proj_w) * c->proj_h);
        XM_CUDA(cudaMalloc(&c->d_bil_out, static_cast<size_t>(mx > 0 ? mx : 1) * 4));
    }
    xm::FrameState* st;
    int rc = staged_state(c, s, true, &st);
    if (rc) return rc;
    rc = launch_bounds(c, st, a->d_events, a->n_events, a->flags, a->time_bounds, a->t_min, a->t_max, c->epoch, s);
    if (rc) return rc;
    if (a->n_events > 0) {
        xm::BilinearParams p;
        p.events = static_cast<const int4*>(a->d_events);
        p.n = a->n_events;
        p.polarity = (a->flags & XM_FLAG_POLARITY) ? 1 : 0;
        p.lut_x = c->d_lut_x_f32;
        p.lut_y = c->d_lut_y_f32;
        p.cam_w = c->cam_w;
        p.cam_h = c->cam_h;
        p.xmap_t = c->d_xmap_t;
        p.xmap_w = c->xmap_w;
        p.xmap_h = c->xmap_h;
        p.col_stride = c->col_stride;
        p.t_px_scale = c->t_px_scale;
        p.x_offset = c->x_offset;
        p.rect_w = c->rect_w;
        p.rect_h = c->rect_h;
        p.view = cam ? 1 : 0;
        p.map = c->d_bil_map;
        p.state = st;
        p.verify = a->time_bounds != XM_TBOUNDS_REDUCE;
        const int grid = grid_for(a->n_events, 256, 4, c->sm_count * 8);
        if (a->flags & XM_FLAG_TIME_F64)
            xm::bilinear_scatter_kernel<true><<<grid, 256, 0, s>>>(p);
        else
            xm::bilinear_scatter_kernel<false><<<grid, 256, 0, s>>>(p);
        XM_LAUNCHED();
    }
    const long long src_cells = cam ? static_cast<long long>(c->cam_w) * c->cam_h : static_cast<long long>(c->rect_w) * c->rect_h;
    float* disp_map = cam ? c->d_bil_out : c->d_bil_rect;
    xm::bilinear_decode_kernel<<<grid_for(src_cells, 256, 4, c->sm_count * 8), 256, 0, s>>>(c->d_bil_map, src_cells, disp_map);
    XM_LAUNCHED();
    if (!cam) {
        xm::dilate_remap_kernel<<<grid_for(out_px, 256, 1, c->sm_count * 16), 256, 0, s>>>(c->d_bil_rect, c->rect_w, c->rect_h, c->d_remap_xy, c->proj_w,
                                                                                           c->proj_h, c->dilate / 2, c->d_bil_out);
        XM_LAUNCHED();
    }
    const xm::OutputSpec o = make_output(c, a->output, c->depth_scale, a->z_near, a->z_far);
    xm::convert_kernel<<<grid_for(out_px, 256, 4, c->sm_count * 8), 256, 0, s>>>(c->d_bil_out, out_px, o, a->d_out);
    XM_LAUNCHED();
    return XM_OK;
}

int frame_impl(XmCtx* c, const XmFrameArgs* a, cudaStream_t s) {
    if (!a->d_out) return fail(XM_ERR_INVALID_ARG, "d_out is NULL");
    if (a->flags & XM_FLAG_BILINEAR) return bilinear_impl(c, a, s);
    const bool f64 = (a->flags & XM_FLAG_TIME_F64) != 0;
    const bool assumed = a->time_bounds != XM_TBOUNDS_REDUCE;
    const bool fixup = assumed && c->opt_auto_fixup;
    cudaError_t err;
    const unsigned epoch = next_epoch(c, fixup ? 2u : 1u, s, &err);
    XM_CUDA(err);
    const int slot = acquire_state_slot(c, s, true, &err);
    XM_CUDA(err);
    c->status_src = nullptr;
    xm::FrameState* st = c->d_state + slot;
    int rc = XM_OK;
    const int polarity = (a->flags & XM_FLAG_POLARITY) ? 1 : 0;

    if (a->time_bounds == XM_TBOUNDS_REDUCE && a->n_events > 0) {
        // exact two-pass form: one extra pass over the stream publishes t.min() / t.max() into the state
        const int rgrid = grid_for(a->n_events, 256, 8, c->sm_count * 8);
        if (f64)
            xm::bounds_reduce_kernel<true><<<rgrid, 256, 0, s>>>(static_cast<const int4*>(a->d_events), a->n_events, polarity, st);
        else
            xm::bounds_reduce_kernel<false><<<rgrid, 256, 0, s>>>(static_cast<const int4*>(a->d_events), a->n_events, polarity, st);
        XM_LAUNCHED();
        c->prev_was_frame = false;
    }

    xm::EventParams p;
    p.events = static_cast<const int4*>(a->d_events);
    p.n = a->n_events;
    p.polarity = polarity;
    p.lut_xy = c->d_lut_xy;
    p.cam_w = c->cam_w;
    p.cam_h = c->cam_h;
    p.xmap_t = c->d_xmap_t;
    p.xmap_w = c->xmap_w;
    p.xmap_h = c->xmap_h;
    p.col_stride = c->col_stride;
    p.t_px_scale = c->t_px_scale;
    p.x_offset = c->x_offset;
    p.rect_w = c->rect_w;
    p.rect_h = c->rect_h;
    p.view = a->view;
    p.map = c->d_map;
    p.epoch = epoch;
    p.state = st;
    p.bounds_mode = a->time_bounds == XM_TBOUNDS_SORTED ? 0 : (a->time_bounds == XM_TBOUNDS_GIVEN ? 1 : 2);
    p.given_lo = a->t_min;
    p.given_hi = a->t_max;
    // griddepcontrol and a device-side tail launch do not mix in one kernel (measured: the grid hangs),
    // so programmatic dependent launch is only used when K1 cannot launch its fix-up
    const bool use_pdl = c->opt_pdl && !fixup;
    p.use_pdl = use_pdl ? 1 : 0;
    p.cap_cols = c->cap_cols;
    p.lookahead = c->opt_lookahead;
    p.stages = c->opt_stages;
    p.win_stages = c->opt_win_stages;
    p.debug = c->opt_debug;
    p.dbg = (c->opt_debug & 8) ? c->d_dbg : nullptr;
    p.arm_fixup = fixup ? 1 : 0;
    p.fix_reduce_grid = grid_for(a->n_events, 256, 8, c->sm_count * 8);
    p.smem_bytes = c->ev_smem;

    // ---- one fused kernel per frame (lean path) -----------------------------------------------------
    const bool tables_safe = c->lut_safe && c->xmap_safe && c->opt_safe_tables;
    const bool fused = c->opt_fused && c->opt_k1_variant == 2 && c->fused_occ > 0 && !f64 && tables_safe &&
                       a->time_bounds != XM_TBOUNDS_REDUCE &&
                       (a->view == XM_VIEW_CAMERA || (c->dilate == 7 && !(c->rect_w & 1))) &&
                       static_cast<size_t>(c->opt_region_cells) * 4 * (xm::kEvThreads / xm::kTileGroup) + xm::kEvSmemHeader <=
                           static_cast<size_t>(c->ev_smem);
    if (fused) {
        xm::FrameParams fp;
        fp.ev = p;
        fp.ev.use_pdl = c->opt_pdl ? 1 : 0;  // no device-side launch in this kernel: PDL is always possible
        fp.ev.arm_fixup = fixup ? 1 : 0;
        xm::EpilogueParams& q = fp.ep;
        q.map = c->d_map;
        q.state = st;
        q.epoch = epoch;
        q.use_pdl = 0;
        const int rslot = (slot + 2) % kStateSlots;
        q.recycle = c->d_state + rslot;
        c->slot_dirty[rslot] = false;
        c->slot_dirty[slot] = true;
        q.remap_xy = c->d_remap_xy;
        q.tile_box = c->d_tile_box;
    q.tile_off = c->opt_tile_off ? c->d_tile_off : nullptr;
        q.rect_w = c->rect_w;
        q.rect_h = c->rect_h;
        q.radius = c->dilate / 2;
        q.region_cap = c->opt_region_cells;
        q.out = make_output(c, a->output, c->depth_scale, a->z_near, a->z_far);
        q.dst = a->d_out;
        q.out_w = a->view == XM_VIEW_CAMERA ? c->cam_w : c->proj_w;
        q.out_h = a->view == XM_VIEW_CAMERA ? c->cam_h : c->proj_h;
        fp.tiles_x = (c->proj_w + xm::kTile - 1) / xm::kTile;
        fp.tiles_y = (c->proj_h + xm::kTile - 1) / xm::kTile;
        // all CTAs must be co-resident (grid-wide barrier inside): never more than occupancy x SMs
        const int occ = c->opt_ctas_per_sm > 0 && c->opt_ctas_per_sm < c->fused_occ ? c->opt_ctas_per_sm : c->fused_occ;
        const int grid = c->sm_count * occ;
        if (c->opt_profile) {
            rc = profile_mark(c, s);
            if (rc) return rc;
        }
        void (*k)(xm::FrameParams) = a->view == XM_VIEW_CAMERA ? xm::frame_kernel<true> : xm::frame_kernel<false>;
        // frame_kernel has a grid-wide barrier: launched cooperatively so that co-residency is guaranteed
        XM_CUDA(launch_pdl(k, dim3(grid), dim3(xm::kWsThreads), c->ev_smem, s, c->opt_pdl && c->prev_was_frame, fp, c->opt_coop != 0));
        XM_LAUNCHED();
        c->prev_was_frame = true;
        if (c->opt_profile) {
            rc = profile_mark(c, s);
            if (rc) return rc;
            rc = profile_mark(c, s);
            if (rc) return rc;
        }
        return XM_OK;
    }

    if (c->opt_profile) {
        rc = profile_mark(c, s);
        if (rc) return rc;
    }
    if (a->n_events > 0) {
        const int occ = c->opt_ctas_per_sm > 0 ? c->opt_ctas_per_sm : (f64 ? c->ev_occ_f64 : c->ev_occ_i64);
        const int grid = grid_for(a->n_events, xm::kEvThreads, xm::kEvPerThread, c->sm_count * occ);
        // when an event violates the assumed bounds the last CTA of K1 tail-launches the exact
        // two-pass fix-up from the device (no extra host launches in the common case)
        const EvKernel k1 = ev_kernel(f64, c->lut_safe && c->xmap_safe && c->opt_safe_tables, c->opt_k1_variant, a->view == XM_VIEW_CAMERA);
        const int threads = xm::kWsThreads;
        // programmatic dependent launch: K1's input-only prologue may overlap the previous frame's epilogue
        XM_CUDA(launch_pdl(k1, dim3(grid), dim3(threads), c->ev_smem, s, use_pdl && c->prev_was_frame, p));
        XM_LAUNCHED();
        c->prev_was_frame = true;
    } else {
        c->prev_was_frame = false;
    }

    if (c->opt_profile) {
        rc = profile_mark(c, s);
        if (rc) return rc;
    }
    xm::EpilogueParams q;
    q.map = c->d_map;
    q.state = st;
    q.epoch = epoch;
    q.use_pdl = use_pdl ? 1 : 0;
    {   // this epilogue clears the state block of the frame after next (ring of kStateSlots)
        const int rslot = (slot + 2) % kStateSlots;
        q.recycle = c->d_state + rslot;
        c->slot_dirty[rslot] = false;
        c->slot_dirty[slot] = true;
    }
    q.remap_xy = c->d_remap_xy;
    q.tile_box = c->d_tile_box;
    q.tile_off = c->opt_tile_off ? c->d_tile_off : nullptr;
    q.rect_w = c->rect_w;
    q.rect_h = c->rect_h;
    q.radius = c->dilate / 2;
    q.region_cap = c->opt_region_cells;
    q.out = make_output(c, a->output, c->depth_scale, a->z_near, a->z_far);
    q.dst = a->d_out;
    const bool pdl2 = use_pdl && c->prev_was_frame;  // K2 directly follows this frame's K1
    if (a->view == XM_VIEW_CAMERA) {
        q.out_w = c->cam_w;
        q.out_h = c->cam_h;
        const int grid = grid_for(static_cast<long long>(c->cam_w) * c->cam_h, 256, 2, c->sm_count * 8);
        XM_CUDA(launch_pdl(xm::epilogue_camera_kernel, dim3(grid), dim3(256), 0, s, pdl2, q));
    } else {
        q.out_w = c->proj_w;
        q.out_h = c->proj_h;
        dim3 grid((c->proj_w + xm::kTile - 1) / xm::kTile, (c->proj_h + xm::kTile - 1) / xm::kTile);
        const size_t smem = static_cast<size_t>(c->opt_region_cells) * 4;
        if (c->dilate == 7 && !(c->rect_w & 1) && c->opt_k2_variant == 1)
            XM_CUDA(launch_pdl(xm::epilogue_projector7_kernel, grid, dim3(256), smem, s, pdl2, q));
        else if (c->dilate == 7)
            XM_CUDA(launch_pdl(xm::epilogue_projector_kernel<3>, grid, dim3(256), smem, s, pdl2, q));
        else
            XM_CUDA(launch_pdl(xm::epilogue_projector_kernel<0>, grid, dim3(256), smem, s, pdl2, q));
    }
    c->prev_was_frame = true;
    XM_LAUNCHED();
    if (c->opt_profile) {
        rc = profile_mark(c, s);
        if (rc) return rc;
    }
    return XM_OK;
}

// true if frames a[0..n) can be rendered by one batch_kernel launch sequence
bool batch_applies(const XmCtx* c, const XmFrameArgs* a, int n) {
    if (!c->opt_batch || n < 2 || c->opt_k1_variant != 2 || c->batch_occ < 1) return false;
    if (!(c->lut_safe && c->xmap_safe && c->opt_safe_tables)) return false;
    if (static_cast<long long>(c->proj_w) * c->proj_h <= 0 && a[0].view == XM_VIEW_PROJECTOR) return false;
    if (a[0].view == XM_VIEW_PROJECTOR && !(c->dilate == 7 && !(c->rect_w & 1))) return false;
    for (int i = 0; i < n; ++i) {
        if (a[i].view != a[0].view || a[i].output != a[0].output || a[i].flags != a[0].flags) return false;
        if (a[i].z_near != a[0].z_near || a[i].z_far != a[0].z_far) return false;
        if ((a[i].flags & (XM_FLAG_TIME_F64 | XM_FLAG_BILINEAR)) || a[i].time_bounds == XM_TBOUNDS_REDUCE || !a[i].d_out) return false;
    }
    return true;
}

int batch_impl(XmCtx* c, const XmFrameArgs* a, int n, cudaStream_t s) {
    // lazily: the extra scatter maps and the batch's state blocks
    if (!c->d_bstate) {
        XM_CUDA(cudaMalloc(&c->d_bstate, (xm::kBatchMax + 1) * sizeof(xm::FrameState)));
        XM_CUDA(cudaMemset(c->d_bstate, 0, (xm::kBatchMax + 1) * sizeof(xm::FrameState)));
    }
    c->d_map_ring[0] = c->d_map;
    const bool cam = a[0].view == XM_VIEW_CAMERA;
    const int polarity = (a[0].flags & XM_FLAG_POLARITY) ? 1 : 0;
    const bool fixup = c->opt_auto_fixup != 0;
    cudaError_t err;
    const unsigned epoch0 = next_epoch(c, static_cast<unsigned>(fixup ? 2 * n : n), s, &err);
    XM_CUDA(err);
    c->prev_was_frame = false;

    xm::BatchBoundsParams bb;
    memset(&bb, 0, sizeof(bb));
    bb.polarity = polarity;
    bb.t_px_scale = c->t_px_scale;
    bb.n_frames = n;
    bb.states = c->d_bstate;
    for (int f = 0; f < n; ++f) {
        bb.events[f] = static_cast<const int4*>(a[f].d_events);
        bb.n[f] = a[f].n_events;
        if (a[f].time_bounds == XM_TBOUNDS_GIVEN) {
            bb.given_mask |= 1u << f;
            bb.lo[f] = a[f].t_min;
            bb.hi[f] = a[f].t_max;
        }
    }
    xm::batch_bounds_kernel<<<n + 1, 64, 0, s>>>(bb);
    XM_LAUNCHED();

    xm::BatchParams bp;
    memset(&bp, 0, sizeof(bp));
    bp.polarity = polarity;
    bp.lut_xy = c->d_lut_xy;
    bp.cam_w = c->cam_w;
    bp.cam_h = c->cam_h;
    bp.xmap_t = c->d_xmap_t;
    bp.xmap_w = c->xmap_w;
    bp.xmap_h = c->xmap_h;
    bp.col_stride = c->col_stride;
    bp.t_px_scale = c->t_px_scale;
    bp.x_offset = c->x_offset;
    bp.rect_w = c->rect_w;
    bp.rect_h = c->rect_h;
    bp.cap_cols = c->batch_cols[a[0].view == XM_VIEW_CAMERA ? 1 : 0];
    bp.stages = c->opt_stages;
    bp.win_stages = c->opt_batch_win_stages;
    bp.alive = c->opt_alive ? c->d_alive : c->d_alive_ones;
    bp.alive_words = c->alive_words;
    bp.alive_shift = c->alive_shift;
    bp.alive_pitch = c->alive_pitch;
    bp.alive_qs = c->alive_qs;
    bp.alive_last = static_cast<unsigned>(c->alive_entries - 1);
    bp.epoch0 = epoch0;
    bp.states = c->d_bstate;
    xm::EpilogueParams& q = bp.ep;
    q.map = nullptr;
    q.state = nullptr;
    q.epoch = 0;
    q.use_pdl = 0;
    q.recycle = nullptr;
    q.remap_xy = c->d_remap_xy;
    q.tile_box = c->d_tile_box;
    q.tile_off = c->opt_tile_off ? c->d_tile_off : nullptr;
    q.rect_w = c->rect_w;
    q.rect_h = c->rect_h;
    q.radius = c->dilate / 2;
    q.region_cap = batch_region_cells(c, cam);
    q.out = make_output(c, a[0].output, c->depth_scale, a[0].z_near, a[0].z_far);
    q.dst = nullptr;
    q.out_w = cam ? c->cam_w : c->proj_w;
    q.out_h = cam ? c->cam_h : c->proj_h;
    const int tiles_x = (c->proj_w + xm::kTile - 1) / xm::kTile, tiles_y = (c->proj_h + xm::kTile - 1) / xm::kTile;
    bp.tiles_x = tiles_x;
    bp.tile_items = cam ? (c->cam_w * c->cam_h + xm::kCamTilePx - 1) / xm::kCamTilePx : tiles_x * tiles_y;
    long long epi_items = (static_cast<long long>(bp.tile_items) + xm::kTileGroups - 1) / xm::kTileGroups;  // per CTA and round
    int tile_warps = xm::kTileWarps;  // epilogue warps per CTA (strip epilogue of large frames: kTileWarpsLarge)
    if (!cam && c->opt_batch_strips && c->d_pix_cell) {
        if (!c->d_dil) XM_CUDA(cudaMalloc(&c->d_dil, static_cast<size_t>(xm::kBatchDilMaps) * c->rect_w * c->rect_h * sizeof(unsigned short)));
        bp.strips = 1;
        bp.pix_cell = c->d_pix_cell;
        for (int i = 0; i < xm::kBatchDilMaps; ++i) bp.dil[i] = c->d_dil + static_cast<size_t>(i) * c->rect_w * c->rect_h;
        // item sizes: large frames leave the epilogue warps plenty of slack -> few large items; small frames are a latency chain
        long long ev_total = 0;
        for (int f = 0; f < n; ++f) ev_total += a[f].n_events;
        // "large": the event stream of a frame outweighs its epilogue (window cells + output pixels) -- measured cross-over
        // on the default geometry (1.7 M cells + pixels) at ~3 M events per frame
        const long long epi_units = static_cast<long long>(c->strip_box[2] - c->strip_box[0] + 1) * (c->strip_box[3] - c->strip_box[1] + 1) +
                                    static_cast<long long>(c->proj_w) * c->proj_h;
        const bool large = 4 * ev_total > 7 * epi_units * n;
        const int rows = c->opt_strip_rows > 0 ? c->opt_strip_rows : (large ? 2 * xm::kStripRows + 6 : xm::kStripRows);
        const int blocks = c->opt_strip_blocks > 0 ? c->opt_strip_blocks : (large ? 2 * xm::kRemapBlocks : xm::kRemapBlocks);
        bp.win = xm::strip_window(c->strip_box[0], c->strip_box[1], c->strip_box[2], c->strip_box[3], rows, blocks);
        bp.tile_items = bp.win.items;
        bp.strip_lag = c->opt_strip_lag;
        const long long item_px = static_cast<long long>(blocks) * xm::kRemapBlockPx;
        bp.p2_items = static_cast<int>((static_cast<long long>(c->proj_w) * c->proj_h + item_px - 1) / item_px);
        epi_items = (bp.tile_items + bp.p2_items + xm::kTileWarps - 1) / xm::kTileWarps;
        tile_warps = c->opt_tile_warps > 0 ? c->opt_tile_warps : (large ? xm::kTileWarpsLarge : xm::kTileWarps);
    }
    bp.n_frames = n;
    bp.debug = c->opt_debug;
    bp.dbg = nullptr;
    if ((c->opt_debug & 8) && c->d_dbg) {  // per-frame hand-off timestamps (XM_DEBUG_HOOKS builds): min slots start at ~0, max slots at 0
        bp.dbg = c->d_dbg;
        std::vector<unsigned long long> init(xm::kBatchMax * 4 + 256, 0ULL);  // [256 ..): phase cycle counters of the strip epilogue
        for (int f = 0; f < xm::kBatchMax; ++f) {
            init[f * 4 + 0] = init[f * 4 + 2] = ~0ULL;
            init[f * 4 + 1] = init[f * 4 + 3] = 0ULL;
        }
        XM_CUDA(cudaMemcpyAsync(c->d_dbg, init.data(), init.size() * 8, cudaMemcpyHostToDevice, s));
    }
    unsigned long long items64 = 0;
    for (int f = 0; f < n; ++f) {
        bp.first_item[f] = static_cast<unsigned>(items64);
        items64 += xm::batch_chunks(a[f].n_events);
    }
    if (items64 > 0xfffffff0ULL) return fail(XM_ERR_UNSUPPORTED, "batch too large");
    unsigned items = static_cast<unsigned>(items64);
    bp.first_item[n] = bp.first_item[n + 1] = items;
    bp.total_items = items;
    for (int f = 0; f < n; ++f) {
        bp.frames[f].events = static_cast<const int4*>(a[f].d_events);
        bp.frames[f].dst = a[f].d_out;
        bp.frames[f].n = a[f].n_events;
    }
    const int kocc = c->batch_occ;
    const int occ = c->opt_ctas_per_sm > 0 && c->opt_ctas_per_sm < kocc ? c->opt_ctas_per_sm : kocc;
    // no more CTAs than there is work: a CTA takes chunks in pairs and runs kTileGroups tiles at a time
    long long want = (static_cast<long long>(items) + 1) / 2;
    const long long want_tiles = epi_items;
    if (want_tiles > want) want = want_tiles;
    int grid = (c->sm_count - c->opt_reserve_sms) * occ;
    if (want < grid) grid = static_cast<int>(want < 1 ? 1 : want);
    // pipelining across frame boundaries pays when a CTA has many chunks per frame; with few, the late
    // publication of a frame's count (in the back half of the NEXT frame's first chunk) delays its tiles
    bp.hard_frames = static_cast<long long>(items) < 8LL * n * grid ? 1 : 0;
    // scatter maps in rotation: a frame holds its map from its first chunk to the end of its epilogue's reads; with small
    // frames that latency, not the work, sets the pace, so more frames are kept in flight
    {
        long long ev_total = 0;
        for (int f = 0; f < n; ++f) ev_total += a[f].n_events;
        const int auto_maps = ev_total <= 1500000LL * n ? 6 : (bp.hard_frames ? 4 : xm::kBatchMaps);  // (measured: profiles/EXPERIMENTS_r02.md)
        bp.n_maps = c->opt_batch_maps > 0 ? c->opt_batch_maps : auto_maps;
    }
    for (int i = 1; i < bp.n_maps; ++i)
        if (!c->d_map_ring[i]) {
            XM_CUDA(cudaMalloc(&c->d_map_ring[i], static_cast<size_t>(c->map_cells) * 8));
            XM_CUDA(cudaMemsetAsync(c->d_map_ring[i], 0, static_cast<size_t>(c->map_cells) * 8, s));
        }
    for (int i = 0; i < bp.n_maps; ++i) bp.maps[i] = c->d_map_ring[i];
    if (c->opt_profile) {
        int rc = profile_mark(c, s);
        if (rc) return rc;
    }
    // Dense streams (nearly every event an inlier: what a scanning projector produces) take the instantiation that
    // aggregates the scatter of dense chunks per warp round.  "scatter_aggregate" = 2 (auto) decides from the inlier
    // fraction of the last batch whose statistics have arrived on the host (read back asynchronously, never waited for).
    if (c->agg_event && c->agg_pending && cudaEventQuery(c->agg_event) == cudaSuccess) {
        c->agg_pending = false;
        c->agg_dense = c->h_agg_stats[0] > 0 && c->h_agg_stats[1] * 10ULL > c->h_agg_stats[0] * 6ULL;
    }
    const bool agg = !cam && (c->opt_scatter_aggregate == 1 || (c->opt_scatter_aggregate == 2 && c->agg_dense));
    if (cam)
        xm::batch_kernel<true><<<grid, xm::kBatchThreads, c->batch_smem[1], s>>>(bp);
    else if (tile_warps == xm::kTileWarpsLarge && agg)
        xm::batch_kernel<false, true, xm::kTileWarpsLarge><<<grid, xm::kWsThreads + xm::kTileWarpsLarge * 32, c->batch_smem[0], s>>>(bp);
    else if (tile_warps == xm::kTileWarpsLarge)
        xm::batch_kernel<false, false, xm::kTileWarpsLarge><<<grid, xm::kWsThreads + xm::kTileWarpsLarge * 32, c->batch_smem[0], s>>>(bp);
    else if (agg)
        xm::batch_kernel<false, true><<<grid, xm::kBatchThreads, c->batch_smem[0], s>>>(bp);
    else
        xm::batch_kernel<false><<<grid, xm::kBatchThreads, c->batch_smem[0], s>>>(bp);
    XM_LAUNCHED();
    if (c->opt_scatter_aggregate == 2 && !cam && !c->agg_pending) {
        if (!c->agg_event) {
            XM_CUDA(cudaEventCreateWithFlags(&c->agg_event, cudaEventDisableTiming));
            XM_CUDA(cudaMallocHost(&c->h_agg_stats, 2 * sizeof(unsigned long long)));
        }
        // n_valid, n_inliers of the batch's first frame (adjacent 64-bit counters of its state block)
        XM_CUDA(cudaMemcpyAsync(c->h_agg_stats, &c->d_bstate[0].n_valid, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        XM_CUDA(cudaEventRecord(c->agg_event, s));
        c->agg_pending = true;
    }
    if (c->opt_profile) {
        int rc = profile_mark(c, s);
        if (rc) return rc;
        rc = profile_mark(c, s);
        if (rc) return rc;
        c->prof_batch_extra += n - 1;  // one event triple covers n frames
    }

    if (fixup) {
        xm::BatchRedoParams rp;
        memset(&rp, 0, sizeof(rp));
        xm::EventParams& p = rp.ev;
        p.polarity = polarity;
        p.lut_xy = c->d_lut_xy;
        p.cam_w = c->cam_w;
        p.cam_h = c->cam_h;
        p.xmap_t = c->d_xmap_t;
        p.xmap_w = c->xmap_w;
        p.xmap_h = c->xmap_h;
        p.col_stride = c->col_stride;
        p.t_px_scale = c->t_px_scale;
        p.x_offset = c->x_offset;
        p.rect_w = c->rect_w;
        p.rect_h = c->rect_h;
        p.view = a[0].view;
        p.map = c->d_map;
        p.cap_cols = c->cap_cols;
        p.lookahead = c->opt_lookahead;
        p.stages = c->opt_stages;
        p.win_stages = c->opt_win_stages;
        p.smem_bytes = c->ev_smem;
        rp.ep = bp.ep;
        rp.ep.map = c->d_map;
        rp.ep.region_cap = c->opt_region_cells;
        rp.epoch_redo0 = epoch0 + static_cast<unsigned>(n);
        rp.n_frames = n;
        rp.sm_count = c->sm_count;
        rp.k1_grid_max = c->sm_count * c->ev_occ_i64;
        rp.k1_smem = c->ev_smem;
        rp.tiles_x = tiles_x;
        rp.tiles_y = tiles_y;
        rp.k2_smem = c->opt_region_cells * 4;
        rp.view = a[0].view;
        rp.states = c->d_bstate;
        for (int f = 0; f < n; ++f) rp.frames[f] = bp.frames[f];
        xm::batch_redo_kernel<<<1, 32, 0, s>>>(rp);
        XM_LAUNCHED();
    }
    c->status_src = c->d_bstate + (n - 1);
    return XM_OK;
}

size_t output_bytes(const XmCtx* c, int view, int output) {
    const size_t px = view == XM_VIEW_CAMERA ? static_cast<size_t>(c->cam_w) * c->cam_h : static_cast<size_t>(c->proj_w) * c->proj_h;
    return output == XM_OUT_BGR ? px * 3 : px * 4;
}

}  // namespace

// =============================================================================================
extern "C" {

int xm_abi_version(void) { return XM_ABI_VERSION; }
const char* xm_last_error(void) { return g_last_error.c_str(); }
int64_t xm_launch_count(void) { return g_launches.load(); }

int xm_ctx_create(const XmTables* t, int device, XmCtx** out) {
    if (!t || !out) return fail(XM_ERR_INVALID_ARG, "null tables or out pointer");
    *out = nullptr;
    if (t->cam_w <= 0 || t->cam_h <= 0 || t->rect_w <= 0 || t->rect_h <= 0)
        return fail(XM_ERR_INVALID_ARG, "camera / rectified size must be positive");
    if (!t->lut_x || !t->lut_y) return fail(XM_ERR_INVALID_ARG, "lut_x / lut_y are required");
    if (t->remap_xy && (t->proj_w <= 0 || t->proj_h <= 0)) return fail(XM_ERR_INVALID_ARG, "projector size must be positive");
    if (t->dilate < 1 || (t->dilate & 1) == 0 || t->dilate > 31) return fail(XM_ERR_INVALID_ARG, "dilate must be odd, 1..31");
    if (t->rect_w > 32767 || t->rect_h > 32767) return fail(XM_ERR_TABLE_RANGE, "rectified image exceeds int16 coordinates");
    if (static_cast<long long>(t->proj_w) * t->proj_h > (1LL << 30) || static_cast<long long>(t->cam_w) * t->cam_h > (1LL << 30))
        return fail(XM_ERR_UNSUPPORTED, "frames are limited to 2^30 pixels");
    int n_dev = 0;
    XM_CUDA(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return fail(XM_ERR_INVALID_ARG, "device %d of %d", device, n_dev);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(XM_ERR_CUDA, "cannot select device %d", device);

    XmCtx* c = new XmCtx();
    c->device = device;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete c;
        return fail(XM_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    }
    c->sm_count = prop.multiProcessorCount;
    c->cam_w = t->cam_w;
    c->cam_h = t->cam_h;
    c->rect_w = t->rect_w;
    c->rect_h = t->rect_h;
    c->proj_w = t->proj_w;
    c->proj_h = t->proj_h;
    c->dilate = t->dilate;
    c->depth_scale = t->depth_scale;

    int rc = XM_OK;
    auto bail = [&](int code) {
        xm_ctx_destroy(c);
        return code;
    };
    const size_t cam_px = static_cast<size_t>(t->cam_w) * t->cam_h;
    {
        std::vector<int> packed(cam_px);
        int min_x = 32767, max_x = -32768;
        for (size_t i = 0; i < cam_px; ++i) {
            min_x = t->lut_x[i] < min_x ? t->lut_x[i] : min_x;
            max_x = t->lut_x[i] > max_x ? t->lut_x[i] : max_x;
        }
        c->lut_x_max = max_x;
        c->lut_x_min = min_x;
        c->lut_safe = min_x > -t->x_offset && t->x_offset > 0;
        for (size_t i = 0; i < cam_px; ++i)
            packed[i] = static_cast<int>((static_cast<unsigned>(static_cast<unsigned short>(t->lut_y[i])) << 16) |
                                         static_cast<unsigned short>(t->lut_x[i]));
        if (cudaMalloc(&c->d_lut_xy, cam_px * 4) != cudaSuccess ||
            cudaMemcpy(c->d_lut_xy, packed.data(), cam_px * 4, cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(XM_ERR_CUDA, "uploading the rectification LUT failed: %s", cudaGetErrorString(cudaGetLastError())));
        c->h_lut_xy.swap(packed);
    }
    if (t->lut_x_f32 && t->lut_y_f32) {
        if (cudaMalloc(&c->d_lut_x_f32, cam_px * 4) != cudaSuccess || cudaMalloc(&c->d_lut_y_f32, cam_px * 4) != cudaSuccess ||
            cudaMemcpy(c->d_lut_x_f32, t->lut_x_f32, cam_px * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(c->d_lut_y_f32, t->lut_y_f32, cam_px * 4, cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(XM_ERR_CUDA, "uploading the float LUT failed: %s", cudaGetErrorString(cudaGetLastError())));
    }
    if (t->remap_xy) {
        const size_t bytes = static_cast<size_t>(t->proj_w) * t->proj_h * 4;
        if (cudaMalloc(&c->d_remap_xy, bytes) != cudaSuccess ||
            cudaMemcpy(c->d_remap_xy, t->remap_xy, bytes, cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(XM_ERR_CUDA, "uploading the remap table failed: %s", cudaGetErrorString(cudaGetLastError())));
        const int tx = (t->proj_w + xm::kTile - 1) / xm::kTile, ty = (t->proj_h + xm::kTile - 1) / xm::kTile;
        std::vector<short4> boxes(static_cast<size_t>(tx) * ty);
        for (int by = 0; by < ty; ++by)
            for (int bx = 0; bx < tx; ++bx) {
                int x0 = 32767, y0 = 32767, x1 = -1, y1 = -1;
                for (int v = by * xm::kTile; v < (by + 1) * xm::kTile && v < t->proj_h; ++v)
                    for (int u = bx * xm::kTile; u < (bx + 1) * xm::kTile && u < t->proj_w; ++u) {
                        const int mx = t->remap_xy[(static_cast<size_t>(v) * t->proj_w + u) * 2];
                        const int my = t->remap_xy[(static_cast<size_t>(v) * t->proj_w + u) * 2 + 1];
                        if (mx < 0 || mx >= t->rect_w || my < 0 || my >= t->rect_h) continue;
                        x0 = mx < x0 ? mx : x0;
                        x1 = mx > x1 ? mx : x1;
                        y0 = my < y0 ? my : y0;
                        y1 = my > y1 ? my : y1;
                    }
                boxes[static_cast<size_t>(by) * tx + bx] = make_short4(static_cast<short>(x0), static_cast<short>(y0), static_cast<short>(x1), static_cast<short>(y1));
                if (x1 >= 0) {  // shared-memory cells this tile's region needs (tile_region(), dilate radius 3)
                    const int rx0 = (x0 - 3) & ~1, rw = (x1 + 3 - rx0 + 2) & ~1, rh = y1 - y0 + 1 + 6;
                    const long long cells = static_cast<long long>(rw + xm::kRowExtra) * rh;
                    if (cells > c->region_need) c->region_need = cells;
                }
            }
        // the projector epilogue's regions are sized to the largest tile of THIS remap table (rounded up), so that
        // no tile falls back to direct 49-tap reads and no shared memory is spent on cells no tile uses
        if (c->region_need > 0) {
            long long cells = (c->region_need + 127) & ~127LL;
            if (cells < 1024) cells = 1024;
            if (cells > 48 * 1024 / 4) cells = 48 * 1024 / 4;
            c->opt_region_cells = static_cast<int>(cells);
        }
        if (cudaMalloc(&c->d_tile_box, boxes.size() * sizeof(short4)) != cudaSuccess ||
            cudaMemcpy(c->d_tile_box, boxes.data(), boxes.size() * sizeof(short4), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(XM_ERR_CUDA, "uploading the tile boxes failed: %s", cudaGetErrorString(cudaGetLastError())));
        // every output pixel's cell inside its tile's region (the layout of xm::tile_region(): even rx0, 3-cell halo,
        // kRowPad zero cells left of the data, row stride rw + kRowExtra); tiles whose region exceeds 65535 cells keep 0xffff
        // and never take the table path (such a region does not fit the shared-memory buffers either)
        {
            std::vector<unsigned short> offs(static_cast<size_t>(t->proj_w) * t->proj_h, 0xffffu);
            for (int by = 0; by < ty; ++by)
                for (int bx = 0; bx < tx; ++bx) {
                    const short4 b = boxes[static_cast<size_t>(by) * tx + bx];
                    if (b.z < 0) continue;
                    const int rx0 = (b.x - 3) & ~1, rw = (b.z + 3 - rx0 + 2) & ~1, ry0 = b.y - 3, rh = b.w - b.y + 1 + 6;
                    const int stride = rw + xm::kRowExtra;
                    if (static_cast<long long>(stride) * rh >= 0xffff) continue;
                    for (int v = by * xm::kTile; v < (by + 1) * xm::kTile && v < t->proj_h; ++v)
                        for (int u = bx * xm::kTile; u < (bx + 1) * xm::kTile && u < t->proj_w; ++u) {
                            const int mx = t->remap_xy[(static_cast<size_t>(v) * t->proj_w + u) * 2];
                            const int my = t->remap_xy[(static_cast<size_t>(v) * t->proj_w + u) * 2 + 1];
                            if (mx < 0 || mx >= t->rect_w || my < 0 || my >= t->rect_h) continue;
                            offs[static_cast<size_t>(v) * t->proj_w + u] = static_cast<unsigned short>((my - ry0) * stride + (mx - rx0) + xm::kRowPad);
                        }
                }
            if (cudaMalloc(&c->d_tile_off, offs.size() * sizeof(unsigned short)) != cudaSuccess ||
                cudaMemcpy(c->d_tile_off, offs.data(), offs.size() * sizeof(unsigned short), cudaMemcpyHostToDevice) != cudaSuccess)
                return bail(fail(XM_ERR_CUDA, "uploading the tile offsets failed: %s", cudaGetErrorString(cudaGetLastError())));
        }
        // strip epilogue: every output pixel's cell in the rectified image (even rect_w only: cell pairs are read as one 16-byte word)
        if (!(t->rect_w & 1) && static_cast<long long>(t->rect_w) * t->rect_h < 0x7fffffffLL) {
            std::vector<unsigned> cells(static_cast<size_t>(t->proj_w) * t->proj_h, 0xffffffffu);
            int bx0 = 1 << 30, by0 = 1 << 30, bx1 = -1, by1 = -1;
            for (size_t i = 0; i < cells.size(); ++i) {
                const int mx = t->remap_xy[i * 2], my = t->remap_xy[i * 2 + 1];
                if (mx < 0 || mx >= t->rect_w || my < 0 || my >= t->rect_h) continue;
                cells[i] = static_cast<unsigned>(my) * static_cast<unsigned>(t->rect_w) + static_cast<unsigned>(mx);
                bx0 = mx < bx0 ? mx : bx0;
                bx1 = mx > bx1 ? mx : bx1;
                by0 = my < by0 ? my : by0;
                by1 = my > by1 ? my : by1;
            }
            if (bx1 < 0) bx0 = by0 = bx1 = by1 = 0;  // no pixel maps into the rectified image: one (empty) item
            c->strip_box[0] = bx0;
            c->strip_box[1] = by0;
            c->strip_box[2] = bx1;
            c->strip_box[3] = by1;
            if (cudaMalloc(&c->d_pix_cell, cells.size() * sizeof(unsigned)) != cudaSuccess ||
                cudaMemcpy(c->d_pix_cell, cells.data(), cells.size() * sizeof(unsigned), cudaMemcpyHostToDevice) != cudaSuccess)
                return bail(fail(XM_ERR_CUDA, "uploading the pixel cell table failed: %s", cudaGetErrorString(cudaGetLastError())));
        }
    }
    c->map_cells = static_cast<long long>(t->rect_w) * t->rect_h;
    if (static_cast<long long>(cam_px) > c->map_cells) c->map_cells = static_cast<long long>(cam_px);
    if (cudaMalloc(&c->d_map, static_cast<size_t>(c->map_cells) * 8) != cudaSuccess ||
        cudaMemset(c->d_map, 0, static_cast<size_t>(c->map_cells) * 8) != cudaSuccess ||
        cudaMalloc(&c->d_state, kStateSlots * sizeof(xm::FrameState)) != cudaSuccess ||
        cudaMemset(c->d_state, 0, kStateSlots * sizeof(xm::FrameState)) != cudaSuccess || cudaMalloc(&c->d_turbo, 768) != cudaSuccess ||
        cudaMemset(c->d_turbo, 0, 768) != cudaSuccess)
        return bail(fail(XM_ERR_CUDA, "allocating the scatter map failed: %s", cudaGetErrorString(cudaGetLastError())));
    if (cudaMalloc(&c->d_depth_lut, 32768 * sizeof(float)) != cudaSuccess)
        return bail(fail(XM_ERR_CUDA, "allocating the depth table failed: %s", cudaGetErrorString(cudaGetLastError())));
    xm::depth_lut_kernel<<<32768 / 256, 256>>>(c->d_depth_lut, 32768, c->depth_scale);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (cudaDeviceSynchronize() != cudaSuccess)
        return bail(fail(XM_ERR_CUDA, "building the depth table failed: %s", cudaGetErrorString(cudaGetLastError())));
    if (t->x_map) {
        if (t->xmap_w <= 0) return bail(fail(XM_ERR_INVALID_ARG, "xmap_w must be positive"));
        rc = upload_xmap(c, t->x_map, t->rect_h, t->xmap_w, t->t_px_scale, t->x_offset);
        if (rc) return bail(rc);
    }
    *out = c;
    // XMAPS_B200_OPTS="key=value,key=value": option overrides for every new context (tuning / A-B runs)
    if (const char* env = getenv("XMAPS_B200_OPTS")) {
        std::string e(env);
        size_t pos = 0;
        while (pos < e.size()) {
            size_t end = e.find(',', pos);
            if (end == std::string::npos) end = e.size();
            const std::string kv = e.substr(pos, end - pos);
            const size_t eq = kv.find('=');
            if (eq != std::string::npos) {
                rc = xm_ctx_set_option(c, kv.substr(0, eq).c_str(), atoll(kv.c_str() + eq + 1));
                if (rc) {
                    *out = nullptr;
                    return bail(rc);
                }
            }
            pos = end + 1;
        }
    }
    return XM_OK;
}

int xm_ctx_destroy(XmCtx* c) {
    if (!c) return XM_OK;
    DeviceGuard guard(c->device);
    cudaFree(c->d_lut_xy);
    cudaFree(c->d_alive);
    cudaFree(c->d_alive_ones);
    cudaFree(c->d_lut_x_f32);
    cudaFree(c->d_bil_map);
    cudaFree(c->d_bil_rect);
    cudaFree(c->d_bil_out);
    cudaFree(c->d_lut_y_f32);
    cudaFree(c->d_xmap_t);
    cudaFree(c->d_remap_xy);
    cudaFree(c->d_tile_box);
    cudaFree(c->d_tile_off);
    cudaFree(c->d_pix_cell);
    cudaFree(c->d_dil);
    if (c->agg_event) cudaEventDestroy(c->agg_event);
    if (c->h_agg_stats) cudaFreeHost(c->h_agg_stats);
    cudaFree(c->d_turbo);
    cudaFree(c->d_depth_lut);
    cudaFree(c->d_dbg);
    cudaFree(c->d_map);
    for (int i = 1; i < xm::kBatchMapsMax; ++i) cudaFree(c->d_map_ring[i]);
    cudaFree(c->d_bstate);
    cudaFree(c->d_state);
    cudaFree(c->d_counts);
    cudaFree(c->d_filter_first);
    cudaFree(c->d_filter_last);
    cudaFree(c->d_pause_idx);
    cudaFree(c->d_stream_scratch);
    cudaFree(c->d_act_first);
    cudaFree(c->d_act_last);
    cudaFree(c->d_act_plan);
    cudaFree(c->d_stage_ev);
    cudaFree(c->d_stage_out);
    for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
    delete c;
    return XM_OK;
}

int xm_ctx_set_xmap(XmCtx* c, const int16_t* h_x_map, int32_t rows, int32_t cols, int32_t t_px_scale, int32_t x_offset) {
    if (!c) return fail(XM_ERR_INVALID_ARG, "null context");
    DeviceGuard guard(c->device);
    XM_CUDA(cudaDeviceSynchronize());
    return upload_xmap(c, h_x_map, rows, cols, t_px_scale, x_offset);
}

int xm_ctx_set_colormap(XmCtx* c, const uint8_t* h_bgr256) {
    if (!c || !h_bgr256) return fail(XM_ERR_INVALID_ARG, "null context or table");
    DeviceGuard guard(c->device);
    XM_CUDA(cudaMemcpy(c->d_turbo, h_bgr256, 768, cudaMemcpyHostToDevice));
    c->have_turbo = true;
    return XM_OK;
}

int xm_ctx_set_option(XmCtx* c, const char* key, int64_t value) {
    if (!c || !key) return fail(XM_ERR_INVALID_ARG, "null context or key");
    DeviceGuard guard(c->device);
    const int v = static_cast<int>(value);
    if (!strcmp(key, "stage_xmap")) {
        c->opt_stage_xmap = v != 0;
        return c->d_xmap_t ? configure_event_kernels(c) : XM_OK;
    }
    if (!strcmp(key, "smem_cols_bytes")) {
        if (v < 0 || v > 200 * 1024) return fail(XM_ERR_INVALID_ARG, "smem_cols_bytes out of range");
        c->opt_smem_cols_bytes = v;
        return c->d_xmap_t ? configure_event_kernels(c) : XM_OK;
    }
    if (!strcmp(key, "stages")) {
        if (v < 1 || v > xm::kWsMaxStages) return fail(XM_ERR_INVALID_ARG, "stages must be 1..%d", xm::kWsMaxStages);
        c->opt_stages = v;
        return c->d_xmap_t ? configure_event_kernels(c) : XM_OK;
    }
    if (!strcmp(key, "debug")) {
        c->opt_debug = v;
        if ((v & 8) && !c->d_dbg) {
            XM_CUDA(cudaMalloc(&c->d_dbg, 4096 * 8 * sizeof(unsigned long long)));
            XM_CUDA(cudaMemset(c->d_dbg, 0, 4096 * 8 * sizeof(unsigned long long)));
        }
        return XM_OK;
    }
    if (!strcmp(key, "win_stages")) {
        if (v < 1 || v > xm::kWsMaxStages) return fail(XM_ERR_INVALID_ARG, "win_stages must be 1..%d", xm::kWsMaxStages);
        c->opt_win_stages = v;
        return c->d_xmap_t ? configure_event_kernels(c) : XM_OK;
    }
    if (!strcmp(key, "batch_win_stages")) {
        if (v < 1 || v > xm::kWsMaxStages) return fail(XM_ERR_INVALID_ARG, "win_stages must be 1..%d", xm::kWsMaxStages);
        c->opt_batch_win_stages = v;
        return c->d_xmap_t ? configure_event_kernels(c) : XM_OK;
    }
    if (!strcmp(key, "k1_variant")) {
        if (v < 1 || v > 2) return fail(XM_ERR_INVALID_ARG, "k1_variant must be 1 or 2");
        c->opt_k1_variant = v;
        return c->d_xmap_t ? configure_event_kernels(c) : XM_OK;
    }
    if (!strcmp(key, "k2_variant")) {
        c->opt_k2_variant = v != 0;
        return XM_OK;
    }
    if (!strcmp(key, "reserve_sms")) {
        if (v < 0 || v >= c->sm_count) return fail(XM_ERR_INVALID_ARG, "reserve_sms out of range");
        c->opt_reserve_sms = v;
        return XM_OK;
    }
    if (!strcmp(key, "batch")) { /* 0: per-frame kernels, 1: batch_kernel (one persistent kernel per <= 32 frames) */
        if (v < 0 || v > 1) return fail(XM_ERR_INVALID_ARG, "batch must be 0 or 1");
        c->opt_batch = v;
        return XM_OK;
    }
    if (!strcmp(key, "fused")) {
        c->opt_fused = v != 0;
        return XM_OK;
    }
    if (!strcmp(key, "coop")) { /* 1: cooperative launch of the fused frame kernel (co-residency guaranteed by the runtime) */
        c->opt_coop = v != 0;
        return XM_OK;
    }
    if (!strcmp(key, "scatter_aggregate")) { /* 0 / 1 / 2 = auto: warp-aggregated scatter of dense chunks in the batch kernel */
        if (v < 0 || v > 2) return fail(XM_ERR_INVALID_ARG, "scatter_aggregate must be 0, 1 or 2");
        c->opt_scatter_aggregate = v;
        return XM_OK;
    }
    if (!strcmp(key, "tile_off")) { /* 1: projector epilogue reads each pixel's region cell from the precomputed table */
        c->opt_tile_off = v != 0;
        return XM_OK;
    }
    if (!strcmp(key, "batch_strips")) { /* 1: the batch kernel's projector epilogue runs as two barrier-free passes (strip dilation, then remap) */
        c->opt_batch_strips = v != 0;
        return c->d_xmap_t ? configure_event_kernels(c) : XM_OK;  // the batch kernel's shared memory depends on it
    }
    if (!strcmp(key, "strip_rows") || !strcmp(key, "strip_blocks")) { /* item sizes of the strip epilogue (0 = auto) */
        if (v < 0 || v > 4096) return fail(XM_ERR_INVALID_ARG, "%s out of range", key);
        (key[6] == 'r' ? c->opt_strip_rows : c->opt_strip_blocks) = v;
        return XM_OK;
    }
    if (!strcmp(key, "tile_warps")) { /* epilogue warps per CTA of the batch kernel's strip epilogue: 0 = auto (2 for large frames), 2, 4 */
        if (v != 0 && v != xm::kTileWarpsLarge && v != xm::kTileWarps) return fail(XM_ERR_INVALID_ARG, "tile_warps must be 0, %d or %d", xm::kTileWarpsLarge, xm::kTileWarps);
        c->opt_tile_warps = v;
        return XM_OK;
    }
    if (!strcmp(key, "strip_lag")) { /* blocks the pass-2 items of a frame trail its pass-1 items by in the strip epilogue's item list */
        if (v < 1 || v >= xm::kBatchDilMaps) return fail(XM_ERR_INVALID_ARG, "strip_lag must be 1 ... %d", xm::kBatchDilMaps - 1);
        c->opt_strip_lag = v;
        return XM_OK;
    }
    if (!strcmp(key, "batch_maps")) { /* scatter maps the batch kernel rotates through (0 = auto) */
        if (v < 0 || v == 1 || v > xm::kBatchMapsMax) return fail(XM_ERR_INVALID_ARG, "batch_maps must be 0 or 2 ... %d", xm::kBatchMapsMax);
        c->opt_batch_maps = v;
        return XM_OK;
    }
    if (!strcmp(key, "alive")) { /* 1: the batch kernel skips the look-ups of events whose pixel block can never yield an inlier */
        c->opt_alive = v != 0;
        return XM_OK;
    }
    if (!strcmp(key, "pdl")) {
        c->opt_pdl = v != 0;
        return XM_OK;
    }
    if (!strcmp(key, "safe_tables")) {
        c->opt_safe_tables = v != 0;
        return XM_OK;
    }
    if (!strcmp(key, "auto_fixup")) {
        c->opt_auto_fixup = v != 0;
        return XM_OK;
    }
    if (!strcmp(key, "lookahead")) {
        if (v < 0) return fail(XM_ERR_INVALID_ARG, "lookahead must be >= 0");
        c->opt_lookahead = v;
        return XM_OK;
    }
    if (!strcmp(key, "ctas_per_sm")) {
        if (v < 0 || v > 32) return fail(XM_ERR_INVALID_ARG, "ctas_per_sm out of range");
        c->opt_ctas_per_sm = v;
        return XM_OK;
    }
    if (!strcmp(key, "profile")) { /* 1: time K1 / K2 of every frame with CUDA events; 0: stop; -1: reset */
        int rc = profile_drain(c);
        if (rc) return rc;
        if (v < 0) {
            c->prof_k1_ms = c->prof_k2_ms = 0.0;
            c->prof_frames = 0;
            c->prof_batch_extra = 0;
        } else {
            c->opt_profile = v != 0;
        }
        return XM_OK;
    }
    if (!strcmp(key, "epoch")) { /* test hook: jump the scatter-map epoch (e.g. next to the 16-bit wrap) */
        if (v < 0 || v > 0xffff) return fail(XM_ERR_INVALID_ARG, "epoch out of range");
        XM_CUDA(cudaDeviceSynchronize());
        XM_CUDA(cudaMemset(c->d_map, 0, static_cast<size_t>(c->map_cells) * 8));
        c->epoch = static_cast<unsigned>(v);
        return XM_OK;
    }
    if (!strcmp(key, "region_cells")) {
        if (v < 0 || v > 48 * 1024 / 4) return fail(XM_ERR_INVALID_ARG, "region_cells out of range");
        c->opt_region_cells = v;
        return c->d_xmap_t ? configure_event_kernels(c) : XM_OK;  // the batch kernel's shared memory depends on it
    }
    return fail(XM_ERR_INVALID_ARG, "unknown option '%s'", key);
}

int xm_ctx_get_option(XmCtx* c, const char* key, int64_t* value) {
    if (!c || !key || !value) return fail(XM_ERR_INVALID_ARG, "null argument");
    if (!strcmp(key, "stage_xmap")) *value = c->opt_stage_xmap;
    else if (!strcmp(key, "smem_cols_bytes")) *value = c->opt_smem_cols_bytes;
    else if (!strcmp(key, "debug_ptr")) *value = static_cast<int64_t>(reinterpret_cast<uintptr_t>(c->d_dbg));
    else if (!strcmp(key, "stages")) *value = c->opt_stages;
    else if (!strcmp(key, "win_stages")) *value = c->opt_win_stages;
    else if (!strcmp(key, "batch_win_stages")) *value = c->opt_batch_win_stages;
    else if (!strcmp(key, "k1_variant")) *value = c->opt_k1_variant;
    else if (!strcmp(key, "pdl")) *value = c->opt_pdl;
    else if (!strcmp(key, "fused")) *value = c->opt_fused;
    else if (!strcmp(key, "alive")) *value = c->opt_alive;
    else if (!strcmp(key, "tile_off")) *value = c->opt_tile_off;
    else if (!strcmp(key, "batch_strips")) *value = c->opt_batch_strips;
    else if (!strcmp(key, "batch_maps")) *value = c->opt_batch_maps;
    else if (!strcmp(key, "strip_rows")) *value = c->opt_strip_rows;
    else if (!strcmp(key, "strip_lag")) *value = c->opt_strip_lag;
    else if (!strcmp(key, "tile_warps")) *value = c->opt_tile_warps;
    else if (!strcmp(key, "strip_blocks")) *value = c->opt_strip_blocks;
    else if (!strcmp(key, "scatter_aggregate")) *value = c->opt_scatter_aggregate;
    else if (!strcmp(key, "scatter_aggregate_now")) *value = c->agg_dense ? 1 : 0;  /* read-only: what "auto" currently selects */
    else if (!strcmp(key, "coop")) *value = c->opt_coop;
    else if (!strcmp(key, "alive_px")) *value = c->alive_px;          /* read-only: camera pixels inside alive blocks */
    else if (!strcmp(key, "batch")) *value = c->opt_batch;
    else if (!strcmp(key, "reserve_sms")) *value = c->opt_reserve_sms;
    else if (!strcmp(key, "batch_occ")) *value = c->batch_occ;
    else if (!strcmp(key, "batch_smem")) *value = c->batch_smem[0];
    else if (!strcmp(key, "batch_cols")) *value = c->batch_cols[0];
    else if (!strcmp(key, "k2_variant")) *value = c->opt_k2_variant;
    else if (!strcmp(key, "safe_tables")) *value = c->opt_safe_tables && c->lut_safe && c->xmap_safe;
    else if (!strcmp(key, "auto_fixup")) *value = c->opt_auto_fixup;
    else if (!strcmp(key, "lookahead")) *value = c->opt_lookahead;
    else if (!strcmp(key, "ctas_per_sm")) *value = c->opt_ctas_per_sm;
    else if (!strcmp(key, "region_cells")) *value = c->opt_region_cells;
    else if (!strcmp(key, "region_need")) *value = c->region_need;    /* read-only: cells of the largest tile region */
    else if (!strcmp(key, "epoch")) *value = c->epoch;
    else if (!strcmp(key, "profile")) *value = c->opt_profile;
    else if (!strcmp(key, "profile_k1_ns") || !strcmp(key, "profile_k2_ns") || !strcmp(key, "profile_frames") || !strcmp(key, "profile_launches")) {
        DeviceGuard guard(c->device);
        int rc = profile_drain(c); /* synchronises on the recorded events */
        if (rc) return rc;
        if (key[9] == '1') *value = static_cast<int64_t>(c->prof_k1_ms * 1e6);
        else if (key[9] == '2') *value = static_cast<int64_t>(c->prof_k2_ms * 1e6);
        else if (key[8] == 'l') *value = c->prof_frames;  /* timed K1 launches (a batch launch covers many frames) */
        else *value = c->prof_frames + c->prof_batch_extra;
    }
    else if (!strcmp(key, "cap_cols")) *value = c->cap_cols;          /* read-only, derived */
    else if (!strcmp(key, "occupancy")) *value = c->ev_occ_i64;       /* read-only, derived */
    else if (!strcmp(key, "sm_count")) *value = c->sm_count;          /* read-only */
    else if (!strcmp(key, "event_smem_bytes")) *value = c->ev_smem;   /* read-only, derived */
    else return fail(XM_ERR_INVALID_ARG, "unknown option '%s'", key);
    return XM_OK;
}

// ---------------------------------------------------------------------------------------------
int xm_frame(XmCtx* c, const XmFrameArgs* a, void* stream) {
    int rc = check_frame_args(c, a);
    if (rc) return rc;
    DeviceGuard guard(c->device);
    return frame_impl(c, a, static_cast<cudaStream_t>(stream));
}

int xm_frame_batch(XmCtx* c, const XmFrameArgs* a, int32_t n_frames, void* stream) {
    if (n_frames < 0 || (n_frames > 0 && !a)) return fail(XM_ERR_INVALID_ARG, "bad batch");
    for (int i = 0; i < n_frames; ++i) {
        int rc = check_frame_args(c, a + i);
        if (rc) return rc;
    }
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (batch_applies(c, a, n_frames)) {
        // one persistent kernel per group of <= kBatchMax frames; a trailing single frame joins the last group
        int i = 0;
        while (i < n_frames) {
            int m = n_frames - i;
            if (m > xm::kBatchMax) m = (m - xm::kBatchMax == 1) ? xm::kBatchMax - 1 : xm::kBatchMax;
            int rc = batch_impl(c, a + i, m, s);
            if (rc) return rc;
            i += m;
        }
        return XM_OK;
    }
    for (int i = 0; i < n_frames; ++i) {
        int rc = frame_impl(c, a + i, s);
        if (rc) return rc;
    }
    return XM_OK;
}

int xm_frame_status(XmCtx* c, XmFrameStatus* h, void* stream) {
    if (!c || !h) return fail(XM_ERR_INVALID_ARG, "null context or status");
    DeviceGuard guard(c->device);
    xm::FrameState st;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    XM_CUDA(cudaMemcpyAsync(&st, c->status_src ? c->status_src : c->d_state + c->last_slot, sizeof(st), cudaMemcpyDeviceToHost, s));
    XM_CUDA(cudaStreamSynchronize(s));
    h->n_events = 0;
    h->n_valid = static_cast<int64_t>(st.n_valid);
    h->n_inliers = static_cast<int64_t>(st.n_inliers);
    h->t_min = st.t_lo_bits;
    h->t_max = st.t_hi_bits;
    h->flags = st.flags;
    h->epoch = c->epoch;
    h->fixup_ran = st.redo;
    return XM_OK;
}

int xm_frame_host(XmCtx* c, const XmFrameArgs* a, const void* h_events, void* h_out, XmFrameStatus* h_status, void* stream) {
    if (!c || !a) return fail(XM_ERR_INVALID_ARG, "null context or arguments");
    if (a->n_events < 0 || (a->n_events > 0 && !h_events) || !h_out) return fail(XM_ERR_INVALID_ARG, "null host buffer");
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long ev_bytes = static_cast<long long>(a->n_events) * 16;
    if (ev_bytes > c->stage_ev_cap) {
        XM_CUDA(cudaStreamSynchronize(s));
        cudaFree(c->d_stage_ev);
        c->d_stage_ev = nullptr;
        c->stage_ev_cap = 0;
        XM_CUDA(cudaMalloc(&c->d_stage_ev, static_cast<size_t>(ev_bytes)));
        c->stage_ev_cap = ev_bytes;
    }
    const long long out_bytes = static_cast<long long>(output_bytes(c, a->view == XM_VIEW_CAMERA ? XM_VIEW_CAMERA : XM_VIEW_PROJECTOR, a->output));
    if (out_bytes > c->stage_out_cap) {
        XM_CUDA(cudaStreamSynchronize(s));
        cudaFree(c->d_stage_out);
        c->d_stage_out = nullptr;
        c->stage_out_cap = 0;
        XM_CUDA(cudaMalloc(&c->d_stage_out, static_cast<size_t>(out_bytes)));
        c->stage_out_cap = out_bytes;
    }
    XmFrameArgs dev = *a;
    dev.d_events = c->d_stage_ev ? c->d_stage_ev : reinterpret_cast<const void*>(16);
    dev.d_out = c->d_stage_out;
    int rc = check_frame_args(c, &dev);
    if (rc) return rc;
    if (ev_bytes) XM_CUDA(cudaMemcpyAsync(c->d_stage_ev, h_events, static_cast<size_t>(ev_bytes), cudaMemcpyHostToDevice, s));
    rc = frame_impl(c, &dev, s);
    if (rc) return rc;
    XM_CUDA(cudaMemcpyAsync(h_out, c->d_stage_out, static_cast<size_t>(out_bytes), cudaMemcpyDeviceToHost, s));
    if (h_status) {
        rc = xm_frame_status(c, h_status, stream);
        if (rc) return rc;
        h_status->n_events = a->n_events;
    } else {
        XM_CUDA(cudaStreamSynchronize(s));
    }
    return XM_OK;
}

int xm_host_alloc(void** h_ptr, int64_t bytes) {
    if (!h_ptr || bytes < 0) return fail(XM_ERR_INVALID_ARG, "bad host allocation request");
    XM_CUDA(cudaHostAlloc(h_ptr, static_cast<size_t>(bytes > 0 ? bytes : 1), cudaHostAllocDefault));
    return XM_OK;
}
int xm_host_free(void* h_ptr) {
    if (h_ptr) XM_CUDA(cudaFreeHost(h_ptr));
    return XM_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU gather without a collective kernel: CUDA IPC mapping of the gathering rank's slab
// ---------------------------------------------------------------------------------------------
int xm_peer_alloc(int device, int64_t bytes, void** d_ptr, XmIpcHandle* handle) {
    if (!d_ptr || !handle || bytes <= 0) return fail(XM_ERR_INVALID_ARG, "xm_peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(XmIpcHandle), "IPC handle size");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(XM_ERR_CUDA, "cannot select device %d", device);
    *d_ptr = nullptr;
    XM_CUDA(cudaMalloc(d_ptr, static_cast<size_t>(bytes)));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *d_ptr);
    if (e != cudaSuccess) {
        cudaFree(*d_ptr);
        *d_ptr = nullptr;
        return fail(XM_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(handle->bytes, &h, sizeof(h));
    return XM_OK;
}

int xm_peer_free(int device, void* d_ptr) {
    DeviceGuard guard(device);
    if (d_ptr) XM_CUDA(cudaFree(d_ptr));
    return XM_OK;
}

int xm_peer_open(int device, int owner_device, const XmIpcHandle* handle, void** d_ptr) {
    if (!d_ptr || !handle) return fail(XM_ERR_INVALID_ARG, "xm_peer_open: bad arguments");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(XM_ERR_CUDA, "cannot select device %d", device);
    *d_ptr = nullptr;
    if (owner_device != device) {
        int can = 0;
        XM_CUDA(cudaDeviceCanAccessPeer(&can, device, owner_device));
        if (!can) return fail(XM_ERR_UNSUPPORTED, "device %d cannot access memory of device %d (no P2P path)", device, owner_device);
        cudaError_t e = cudaDeviceEnablePeerAccess(owner_device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled)
            cudaGetLastError();  // clear
        else if (e != cudaSuccess)
            return fail(XM_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", owner_device, cudaGetErrorString(e));
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle->bytes, sizeof(h));
    XM_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return XM_OK;
}

int xm_peer_close(int device, void* d_ptr) {
    DeviceGuard guard(device);
    if (d_ptr) XM_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return XM_OK;
}

int xm_peer_copy(void* d_dst, int dst_device, const void* d_src, int src_device, int64_t bytes, void* stream) {
    if (bytes < 0 || (bytes > 0 && (!d_dst || !d_src))) return fail(XM_ERR_INVALID_ARG, "xm_peer_copy: bad arguments");
    if (bytes == 0) return XM_OK;
    XM_CUDA(cudaMemcpyPeerAsync(d_dst, dst_device, d_src, src_device, static_cast<size_t>(bytes), static_cast<cudaStream_t>(stream)));
    return XM_OK;
}

int xm_peer_info(int device, int peer_device, int32_t* can_access, int32_t* perf_rank) {
    if (!can_access || !perf_rank) return fail(XM_ERR_INVALID_ARG, "xm_peer_info: null output");
    int can = 0, rank = -1;
    XM_CUDA(cudaDeviceCanAccessPeer(&can, device, peer_device));
    if (can) XM_CUDA(cudaDeviceGetP2PAttribute(&rank, cudaDevP2PAttrPerformanceRank, device, peer_device));
    *can_access = can;
    *perf_rank = rank;
    return XM_OK;
}

// ---------------------------------------------------------------------------------------------
// stage-by-stage entry points
// ---------------------------------------------------------------------------------------------
int xm_rectify_i16(XmCtx* c, const void* d_events, int64_t n, int16_t* d_x, int16_t* d_y, void* stream) {
    if (!c || n < 0 || (n > 0 && (!d_events || !d_x || !d_y))) return fail(XM_ERR_INVALID_ARG, "rectify_i16: bad arguments");
    if (n == 0) return XM_OK;
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    xm::FrameState* st;
    int rc = staged_state(c, s, true, &st);
    if (rc) return rc;
    xm::rectify_i16_kernel<<<grid_for(n, 256, 4, c->sm_count * 8), 256, 0, s>>>(static_cast<const int4*>(d_events), n, c->d_lut_xy,
                                                                                c->cam_w, c->cam_h, d_x, d_y, st);
    XM_LAUNCHED();
    return XM_OK;
}

int xm_rectify_f32(XmCtx* c, const void* d_events, int64_t n, float* d_x, float* d_y, void* stream) {
    if (!c || n < 0 || (n > 0 && (!d_events || !d_x || !d_y))) return fail(XM_ERR_INVALID_ARG, "rectify_f32: bad arguments");
    if (!c->d_lut_x_f32) return fail(XM_ERR_INVALID_ARG, "rectify_f32 needs XmTables.lut_x_f32 / lut_y_f32");
    if (n == 0) return XM_OK;
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    xm::FrameState* st;
    int rc = staged_state(c, s, true, &st);
    if (rc) return rc;
    xm::rectify_f32_kernel<<<grid_for(n, 256, 4, c->sm_count * 8), 256, 0, s>>>(
        static_cast<const int4*>(d_events), n, c->d_lut_x_f32, c->d_lut_y_f32, c->cam_w, c->cam_h, d_x, d_y, st);
    XM_LAUNCHED();
    return XM_OK;
}

int xm_event_disparity(XmCtx* c, const XmFrameArgs* a, const int16_t* d_x_rect, const int16_t* d_y_rect, int16_t* d_disp_full,
                       uint8_t* d_mask, void* stream) {
    if (!c || !a) return fail(XM_ERR_INVALID_ARG, "null context or arguments");
    if (a->n_events < 0 || (a->n_events > 0 && (!a->d_events || !d_disp_full || !d_mask)))
        return fail(XM_ERR_INVALID_ARG, "event_disparity: bad arguments");
    if ((d_x_rect == nullptr) != (d_y_rect == nullptr)) return fail(XM_ERR_INVALID_ARG, "give both or neither of d_x_rect / d_y_rect");
    if (a->time_bounds < XM_TBOUNDS_REDUCE || a->time_bounds > XM_TBOUNDS_GIVEN) return fail(XM_ERR_INVALID_ARG, "unknown time_bounds");
    if (!c->d_xmap_t) return fail(XM_ERR_NO_XMAP, "no X-map");
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    xm::FrameState* st;
    int rc = staged_state(c, s, false, &st);
    if (rc) return rc;
    rc = launch_bounds(c, st, a->d_events, a->n_events, a->flags, a->time_bounds, a->t_min, a->t_max, c->epoch, s);
    if (rc) return rc;
    if (a->n_events == 0) return XM_OK;
    xm::DisparityParams p;
    p.events = static_cast<const int4*>(a->d_events);
    p.n = a->n_events;
    p.polarity = (a->flags & XM_FLAG_POLARITY) ? 1 : 0;
    p.lut_xy = c->d_lut_xy;
    p.cam_w = c->cam_w;
    p.cam_h = c->cam_h;
    p.x_rect = d_x_rect;
    p.y_rect = d_y_rect;
    p.xmap_t = c->d_xmap_t;
    p.xmap_w = c->xmap_w;
    p.xmap_h = c->xmap_h;
    p.col_stride = c->col_stride;
    p.t_px_scale = c->t_px_scale;
    p.x_offset = c->x_offset;
    p.disp_full = d_disp_full;
    p.mask = d_mask;
    p.state = st;
    p.verify = a->time_bounds != XM_TBOUNDS_REDUCE;
    const int grid = grid_for(a->n_events, 256, 4, c->sm_count * 8);
    if (a->flags & XM_FLAG_TIME_F64)
        xm::event_disparity_kernel<true><<<grid, 256, 0, s>>>(p);
    else
        xm::event_disparity_kernel<false><<<grid, 256, 0, s>>>(p);
    XM_LAUNCHED();
    return XM_OK;
}

int xm_compact_i16(XmCtx* c, const int16_t* d_vals, const uint8_t* d_mask, int64_t n, int16_t* d_out, int64_t* d_count, void* stream) {
    if (!c || n < 0 || !d_count || (n > 0 && (!d_vals || !d_mask || !d_out))) return fail(XM_ERR_INVALID_ARG, "compact: bad arguments");
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        XM_CUDA(cudaMemsetAsync(d_count, 0, 8, s));
        return XM_OK;
    }
    const long long blocks = (n + xm::kCompactBlock - 1) / xm::kCompactBlock;
    if (blocks > c->counts_cap) {
        XM_CUDA(cudaStreamSynchronize(s));
        cudaFree(c->d_counts);
        c->d_counts = nullptr;
        c->counts_cap = 0;
        XM_CUDA(cudaMalloc(&c->d_counts, static_cast<size_t>(blocks) * 4));
        c->counts_cap = blocks;
    }
    xm::compact_count_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(d_mask, n, c->d_counts);
    XM_LAUNCHED();
    xm::compact_scan_kernel<<<1, 1024, 0, s>>>(c->d_counts, blocks, reinterpret_cast<long long*>(d_count));
    XM_LAUNCHED();
    xm::compact_write_kernel<short><<<static_cast<unsigned>(blocks), 256, 0, s>>>(d_vals, d_mask, n, c->d_counts, d_out);
    XM_LAUNCHED();
    return XM_OK;
}

int xm_scatter_last_wins(XmCtx* c, const int16_t* d_rows, const int16_t* d_cols, const int16_t* d_vals, int64_t n, int32_t h,
                         int32_t w, float* d_map, void* stream) {
    if (!c || n < 0 || h <= 0 || w <= 0 || !d_map || (n > 0 && (!d_rows || !d_cols || !d_vals)))
        return fail(XM_ERR_INVALID_ARG, "scatter: bad arguments");
    if (static_cast<long long>(h) * w > c->map_cells) return fail(XM_ERR_UNSUPPORTED, "scatter target %d x %d exceeds the context's map", h, w);
    if (n > 0xffffffffLL) return fail(XM_ERR_UNSUPPORTED, "more than 2^32 - 1 values");
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t err;
    const unsigned epoch = next_epoch(c, 1, s, &err);
    XM_CUDA(err);
    xm::FrameState* st;
    int rc = staged_state(c, s, true, &st);
    if (rc) return rc;
    if (n > 0) {
        xm::scatter_keys_kernel<<<grid_for(n, 256, 4, c->sm_count * 8), 256, 0, s>>>(d_rows, d_cols, d_vals, n, h, w, c->d_map, epoch,
                                                                                    st);
        XM_LAUNCHED();
    }
    const long long cells = static_cast<long long>(h) * w;
    xm::decode_map_kernel<<<grid_for(cells, 256, 4, c->sm_count * 8), 256, 0, s>>>(c->d_map, cells, epoch, d_map);
    XM_LAUNCHED();
    return XM_OK;
}

int xm_dilate_remap(XmCtx* c, const float* d_rect_map, float* d_proj_map, void* stream) {
    if (!c || !d_rect_map || !d_proj_map) return fail(XM_ERR_INVALID_ARG, "dilate_remap: bad arguments");
    if (!c->d_remap_xy) return fail(XM_ERR_INVALID_ARG, "dilate_remap needs XmTables.remap_xy");
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long n = static_cast<long long>(c->proj_w) * c->proj_h;
    xm::dilate_remap_kernel<<<grid_for(n, 256, 1, c->sm_count * 16), 256, 0, s>>>(d_rect_map, c->rect_w, c->rect_h, c->d_remap_xy,
                                                                                 c->proj_w, c->proj_h, c->dilate / 2, d_proj_map);
    XM_LAUNCHED();
    return XM_OK;
}

int xm_disp_to_depth(XmCtx* c, const float* d_disp, int64_t n, double depth_scale, float* d_depth, void* stream) {
    if (!c || n < 0 || (n > 0 && (!d_disp || !d_depth))) return fail(XM_ERR_INVALID_ARG, "disp_to_depth: bad arguments");
    if (n == 0) return XM_OK;
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    xm::convert_kernel<<<grid_for(n, 256, 4, c->sm_count * 8), 256, 0, s>>>(d_disp, n, make_output(c, XM_OUT_DEPTH, depth_scale, 0.f, 0.f),
                                                                            d_depth);
    XM_LAUNCHED();
    return XM_OK;
}

int xm_colorize(XmCtx* c, const float* d_disp, int64_t n, double depth_scale, float z_near, float z_far, uint8_t* d_bgr, void* stream) {
    if (!c || n < 0 || (n > 0 && (!d_disp || !d_bgr))) return fail(XM_ERR_INVALID_ARG, "colorize: bad arguments");
    if (!c->have_turbo) return fail(XM_ERR_INVALID_ARG, "colorize needs a colour map (xm_ctx_set_colormap)");
    if (n == 0) return XM_OK;
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    xm::convert_kernel<<<grid_for(n, 256, 4, c->sm_count * 8), 256, 0, s>>>(d_disp, n, make_output(c, XM_OUT_BGR, depth_scale, z_near, z_far),
                                                                            d_bgr);
    XM_LAUNCHED();
    return XM_OK;
}

int xm_point_cloud(XmCtx* c, const float* d_x, const float* d_y, const float* d_disp, int64_t n, const double* h_Q, float* d_xyz,
                   void* stream) {
    if (!c || n < 0 || !h_Q || (n > 0 && (!d_x || !d_y || !d_disp || !d_xyz))) return fail(XM_ERR_INVALID_ARG, "point_cloud: bad arguments");
    if (n == 0) return XM_OK;
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    xm::Mat4f q;
    for (int i = 0; i < 16; ++i) q.m[i] = static_cast<float>(h_Q[i]);
    xm::point_cloud_kernel<<<grid_for(n, 256, 4, c->sm_count * 8), 256, 0, s>>>(d_x, d_y, d_disp, n, q, d_xyz);
    XM_LAUNCHED();
    return XM_OK;
}

namespace {
int ensure_counts(XmCtx* c, long long blocks, cudaStream_t s) {
    if (blocks > c->counts_cap) {
        XM_CUDA(cudaStreamSynchronize(s));
        cudaFree(c->d_counts);
        c->d_counts = nullptr;
        c->counts_cap = 0;
        XM_CUDA(cudaMalloc(&c->d_counts, static_cast<size_t>(blocks) * 4));
        c->counts_cap = blocks;
    }
    return XM_OK;
}
int ensure_stream_scratch(XmCtx* c) {
    if (!c->d_stream_scratch) XM_CUDA(cudaMalloc(&c->d_stream_scratch, 64));
    return XM_OK;
}
}  // namespace

int xm_polarity_filter(XmCtx* c, const void* d_events, int64_t n, void* d_out, int64_t* d_count, void* stream) {
    if (!c || n < 0 || !d_count || (n > 0 && (!d_events || !d_out))) return fail(XM_ERR_INVALID_ARG, "polarity_filter: bad arguments");
    if (reinterpret_cast<uintptr_t>(d_events) & 15 || reinterpret_cast<uintptr_t>(d_out) & 15)
        return fail(XM_ERR_INVALID_ARG, "event buffers must be 16-byte aligned");
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        XM_CUDA(cudaMemsetAsync(d_count, 0, 8, s));
        return XM_OK;
    }
    const long long blocks = (n + xm::kCompactBlock - 1) / xm::kCompactBlock;
    int rc = ensure_counts(c, blocks, s);
    if (rc) return rc;
    const xm::PolarityPred pred{static_cast<const int4*>(d_events)};
    const xm::CopyEventEmit emit{static_cast<const int4*>(d_events), static_cast<int4*>(d_out)};
    xm::flag_count_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(pred, n, c->d_counts);
    XM_LAUNCHED();
    xm::compact_scan_kernel<<<1, 1024, 0, s>>>(c->d_counts, blocks, reinterpret_cast<long long*>(d_count));
    XM_LAUNCHED();
    xm::flag_write_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(pred, emit, n, c->d_counts);
    XM_LAUNCHED();
    return XM_OK;
}

namespace {
constexpr int kActPlanCap = 4096;  // sub-packets per call (a packet of kActPlanCap thresholds' length; the last one takes the rest)

int ensure_activity(XmCtx* c) {
    const size_t cells = static_cast<size_t>(c->cam_w) * c->cam_h;
    if (!c->d_act_first) XM_CUDA(cudaMalloc(&c->d_act_first, cells * sizeof(unsigned)));
    if (!c->d_act_last) {
        XM_CUDA(cudaMalloc(&c->d_act_last, cells * sizeof(long long)));
        XM_CUDA(cudaMemset(c->d_act_last, 0, cells * sizeof(long long)));
    }
    if (!c->d_act_plan) XM_CUDA(cudaMalloc(&c->d_act_plan, (kActPlanCap + 1 + 4) * sizeof(long long)));
    return XM_OK;
}
}  // namespace

int xm_activity_reset(XmCtx* c, void* stream) {
    if (!c) return fail(XM_ERR_INVALID_ARG, "activity_reset: NULL context");
    DeviceGuard guard(c->device);
    if (c->d_act_last)
        XM_CUDA(cudaMemsetAsync(c->d_act_last, 0, static_cast<size_t>(c->cam_w) * c->cam_h * sizeof(long long), static_cast<cudaStream_t>(stream)));
    return XM_OK;
}

int xm_activity_filter(XmCtx* c, const void* d_events, int64_t n, int64_t threshold_us, void* d_out, int64_t* d_count, void* stream) {
    if (!c || n < 0 || !d_count || (n > 0 && (!d_events || !d_out))) return fail(XM_ERR_INVALID_ARG, "activity_filter: bad arguments");
    if (threshold_us <= 0) return fail(XM_ERR_INVALID_ARG, "activity_filter: threshold must be positive");
    if (n > 0xfffffffeLL) return fail(XM_ERR_UNSUPPORTED, "more than 2^32 - 2 events");
    if (reinterpret_cast<uintptr_t>(d_events) & 15 || reinterpret_cast<uintptr_t>(d_out) & 15)
        return fail(XM_ERR_INVALID_ARG, "event buffers must be 16-byte aligned");
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    XM_CUDA(cudaMemsetAsync(d_count, 0, 8, s));
    if (n == 0) return XM_OK;
    int rc = ensure_activity(c);
    if (rc) return rc;
    const int4* ev = static_cast<const int4*>(d_events);
    long long* bounds = c->d_act_plan;
    int* d_nsub = reinterpret_cast<int*>(c->d_act_plan + kActPlanCap + 1);
    unsigned* d_unsorted = reinterpret_cast<unsigned*>(c->d_act_plan + kActPlanCap + 2);
    long long* d_part = c->d_act_plan + kActPlanCap + 3;
    XM_CUDA(cudaMemsetAsync(d_unsorted, 0, 8, s));
    // sub-packets that each span less than the threshold (normally one: a packet is a slice of a frame)
    xm::activity_plan_kernel<<<1, 1, 0, s>>>(ev, n, threshold_us, bounds, kActPlanCap, d_nsub);
    XM_LAUNCHED();
    std::vector<long long> h_bounds(kActPlanCap + 2);
    XM_CUDA(cudaMemcpyAsync(h_bounds.data(), bounds, (kActPlanCap + 2) * sizeof(long long), cudaMemcpyDeviceToHost, s));
    XM_CUDA(cudaStreamSynchronize(s));
    const int nsub = *reinterpret_cast<const int*>(&h_bounds[kActPlanCap + 1]);
    const long long cells = static_cast<long long>(c->cam_w) * c->cam_h;
    for (int k = 0; k < nsub; ++k) {
        const long long lo = h_bounds[k], cnt = h_bounds[k + 1] - lo;
        if (cnt <= 0) continue;
        xm::ActivityParams p;
        p.events = ev + lo;
        p.n = cnt;
        p.threshold = threshold_us;
        p.cols = c->cam_w;
        p.rows = c->cam_h;
        p.first = c->d_act_first;
        p.last = c->d_act_last;
        p.unsorted = d_unsorted;
        const long long blocks = (cnt + xm::kCompactBlock - 1) / xm::kCompactBlock;
        rc = ensure_counts(c, blocks, s);
        if (rc) return rc;
        xm::activity_clear_kernel<<<grid_for(cells, 256, 4, c->sm_count * 8), 256, 0, s>>>(p.first, cells);
        XM_LAUNCHED();
        xm::activity_mark_kernel<<<grid_for(cnt, 256, 4, c->sm_count * 8), 256, 0, s>>>(p);
        XM_LAUNCHED();
        const xm::ActivityPred pred{p};
        const xm::ActivityEmit emit{p.events, static_cast<int4*>(d_out), reinterpret_cast<const long long*>(d_count)};
        xm::flag_count_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(pred, cnt, c->d_counts);
        XM_LAUNCHED();
        xm::compact_scan_kernel<<<1, 1024, 0, s>>>(c->d_counts, blocks, d_part);
        XM_LAUNCHED();
        xm::flag_write_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(pred, emit, cnt, c->d_counts);
        XM_LAUNCHED();
        xm::activity_update_kernel<<<grid_for(cnt, 256, 4, c->sm_count * 8), 256, 0, s>>>(p);
        XM_LAUNCHED();
        xm::activity_advance_kernel<<<1, 1, 0, s>>>(reinterpret_cast<long long*>(d_count), d_part);
        XM_LAUNCHED();
    }
    unsigned h_unsorted = 0;
    XM_CUDA(cudaMemcpyAsync(&h_unsorted, d_unsorted, 4, cudaMemcpyDeviceToHost, s));
    XM_CUDA(cudaStreamSynchronize(s));
    if (h_unsorted) return fail(XM_ERR_INVALID_ARG, "activity_filter: timestamps are not sorted (the filter's state is now undefined: xm_activity_reset)");
    return XM_OK;
}

int xm_filter_events(XmCtx* c, const void* d_events, int64_t n, int32_t mode, const int16_t* d_x_rect, int32_t as_reference,
                     void* d_out, int64_t* d_count, void* stream) {
    if (!c || n < 0 || !d_count || (n > 0 && (!d_events || !d_out))) return fail(XM_ERR_INVALID_ARG, "filter: bad arguments");
    if (mode < XM_FILTER_FIRST_YT || mode > XM_FILTER_MEAN_XY) return fail(XM_ERR_INVALID_ARG, "filter: unknown mode %d", mode);
    if (mode == XM_FILTER_FIRST_YT && n > 0 && !d_x_rect) return fail(XM_ERR_INVALID_ARG, "XM_FILTER_FIRST_YT needs d_x_rect");
    if (n > 0xfffffffeLL) return fail(XM_ERR_UNSUPPORTED, "more than 2^32 - 2 events");
    if (reinterpret_cast<uintptr_t>(d_events) & 15 || reinterpret_cast<uintptr_t>(d_out) & 15)
        return fail(XM_ERR_INVALID_ARG, "event buffers must be 16-byte aligned");
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    xm::FrameState* st;
    int rc = staged_state(c, s, true, &st);
    if (rc) return rc;
    if (n == 0) {
        XM_CUDA(cudaMemsetAsync(d_count, 0, 8, s));
        return XM_OK;
    }
    const int yt_stride = c->lut_x_max + 1 > 1 ? c->lut_x_max + 1 : 1;
    const int max_stride = yt_stride > c->cam_w ? yt_stride : c->cam_w;
    if (!c->d_filter_first) {
        const size_t bytes = static_cast<size_t>(c->cam_h) * max_stride * 4;
        XM_CUDA(cudaMalloc(&c->d_filter_first, bytes));
        XM_CUDA(cudaMalloc(&c->d_filter_last, bytes));
    }
    rc = ensure_stream_scratch(c);
    if (rc) return rc;
    xm::FilterParams p;
    p.events = static_cast<const int4*>(d_events);
    p.n = n;
    p.xp = d_x_rect;
    p.mode = mode;
    p.rows = c->cam_h;
    p.cols = c->cam_w;
    p.stride = mode == XM_FILTER_FIRST_YT ? yt_stride : c->cam_w;
    p.as_reference = as_reference ? 1 : 0;
    p.first = c->d_filter_first;
    p.last = c->d_filter_last;
    p.xp_max = reinterpret_cast<int*>(static_cast<char*>(c->d_stream_scratch) + 32);
    p.state = st;
    const long long cells = static_cast<long long>(p.rows) * p.stride;
    const long long blocks = (cells + xm::kCompactBlock - 1) / xm::kCompactBlock;
    rc = ensure_counts(c, blocks, s);
    if (rc) return rc;
    xm::filter_prepare_kernel<<<grid_for(cells, 256, 4, c->sm_count * 8), 256, 0, s>>>(p);
    XM_LAUNCHED();
    if (mode == XM_FILTER_FIRST_YT) {
        xm::filter_xpmax_kernel<<<grid_for(n, 256, 8, c->sm_count * 8), 256, 0, s>>>(p);
        XM_LAUNCHED();
    }
    xm::filter_mark_kernel<<<grid_for(n, 256, 4, c->sm_count * 8), 256, 0, s>>>(p);
    XM_LAUNCHED();
    const xm::FilterPred pred{c->d_filter_last};
    const xm::FilterEmit emit{p, static_cast<int4*>(d_out)};
    xm::flag_count_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(pred, cells, c->d_counts);
    XM_LAUNCHED();
    xm::compact_scan_kernel<<<1, 1024, 0, s>>>(c->d_counts, blocks, reinterpret_cast<long long*>(d_count));
    XM_LAUNCHED();
    xm::flag_write_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(pred, emit, cells, c->d_counts);
    XM_LAUNCHED();
    return XM_OK;
}

int xm_find_trigger(XmCtx* c, const void* d_events, int64_t n, int64_t pause_thresh_us, double frame_us, int64_t min_events,
                    int64_t* d_result, void* stream) {
    if (!c || n < 0 || !d_result || (n > 0 && !d_events)) return fail(XM_ERR_INVALID_ARG, "find_trigger: bad arguments");
    if (n > 0xfffffffeLL) return fail(XM_ERR_UNSUPPORTED, "more than 2^32 - 2 events");
    if (reinterpret_cast<uintptr_t>(d_events) & 15) return fail(XM_ERR_INVALID_ARG, "d_events must be 16-byte aligned");
    if (!(frame_us > 0.0)) return fail(XM_ERR_INVALID_ARG, "frame_us must be positive");
    DeviceGuard guard(c->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = ensure_stream_scratch(c);
    if (rc) return rc;
    xm::TriggerScratch* sc = static_cast<xm::TriggerScratch*>(c->d_stream_scratch);
    const long long blocks = n > 0 ? (n + xm::kCompactBlock - 1) / xm::kCompactBlock : 1;
    rc = ensure_counts(c, blocks, s);
    if (rc) return rc;
    if (n > c->pause_cap) {
        XM_CUDA(cudaStreamSynchronize(s));
        cudaFree(c->d_pause_idx);
        c->d_pause_idx = nullptr;
        c->pause_cap = 0;
        XM_CUDA(cudaMalloc(&c->d_pause_idx, static_cast<size_t>(n) * 4));
        c->pause_cap = n;
    }
    XM_CUDA(cudaMemsetAsync(sc, 0xff, sizeof(xm::TriggerScratch), s));  // first_pair = none
    const int4* ev = static_cast<const int4*>(d_events);
    const xm::PausePred pred{ev, n, pause_thresh_us};
    const xm::PauseEmit emit{c->d_pause_idx};
    xm::flag_count_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(pred, n, c->d_counts);
    XM_LAUNCHED();
    xm::compact_scan_kernel<<<1, 1024, 0, s>>>(c->d_counts, blocks, &sc->n_pauses);
    XM_LAUNCHED();
    xm::flag_write_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(pred, emit, n, c->d_counts);
    XM_LAUNCHED();
    xm::trigger_pairs_kernel<<<c->sm_count * 2, 256, 0, s>>>(ev, c->d_pause_idx, sc, frame_us / 2.0);
    XM_LAUNCHED();
    xm::trigger_decide_kernel<<<1, 32, 0, s>>>(ev, c->d_pause_idx, sc, frame_us, min_events, reinterpret_cast<long long*>(d_result));
    XM_LAUNCHED();
    return XM_OK;
}

int xm_build_xmap(int device, const float* d_time_map, int32_t h, int32_t w, int32_t x_map_width, int32_t t_px_scale, int32_t x_offset,
                  int32_t num_scanlines, int16_t* d_x_map, float* d_t_diffs, void* stream) {
    if (!d_time_map || !d_x_map || h <= 0 || w <= 0 || x_map_width <= 0 || t_px_scale <= 0 || num_scanlines <= 0)
        return fail(XM_ERR_INVALID_ARG, "build_xmap: bad arguments");
    // asserts of the reference (x_maps_disparity.py:52-53)
    if (h > 32767 || w + x_offset > 32767) return fail(XM_ERR_TABLE_RANGE, "time map %d x %d (+%d) exceeds int16", h, w, x_offset);
    if (static_cast<size_t>(w) * 4 > 200 * 1024) return fail(XM_ERR_UNSUPPORTED, "time-map row of %d floats exceeds shared memory", w);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(XM_ERR_CUDA, "cannot select device %d", device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int optin = 0;
    XM_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    cudaFuncAttributes fa;
    XM_CUDA(cudaFuncGetAttributes(&fa, xm::build_xmap_kernel));
    optin -= static_cast<int>(fa.sharedSizeBytes);
    if (w * 4 > optin) return fail(XM_ERR_UNSUPPORTED, "time-map row of %d floats exceeds shared memory", w);
    XM_CUDA(cudaFuncSetAttribute(xm::build_xmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    xm::build_xmap_kernel<<<h, 256, static_cast<size_t>(w) * 4, s>>>(d_time_map, h, w, x_map_width, t_px_scale, x_offset, num_scanlines,
                                                                    d_x_map, d_t_diffs);
    XM_LAUNCHED();
    return XM_OK;
}

int xm_build_inverse_lut(int device, const double* h_K, const double* h_D, int32_t n_dist, const double* h_RR, int32_t w, int32_t h,
                         float* d_mapx, float* d_mapy, int16_t* d_xy_i16, void* stream) {
    if (!h_K || !h_RR || w <= 0 || h <= 0 || n_dist < 0 || (n_dist > 0 && !h_D) || (!d_mapx && !d_mapy && !d_xy_i16))
        return fail(XM_ERR_INVALID_ARG, "build_inverse_lut: bad arguments");
    if (n_dist != 0 && n_dist != 4 && n_dist != 5 && n_dist != 8 && n_dist != 12 && n_dist != 14)
        return fail(XM_ERR_INVALID_ARG, "build_inverse_lut: %d distortion coefficients (OpenCV takes 4, 5, 8, 12 or 14)", n_dist);
    if (n_dist == 14 && (h_D[12] != 0.0 || h_D[13] != 0.0)) return fail(XM_ERR_UNSUPPORTED, "build_inverse_lut: tilted sensor model (tau_x, tau_y) is not supported");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(XM_ERR_CUDA, "cannot select device %d", device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    xm::InverseLutParams p;
    memset(&p, 0, sizeof(p));
    p.ifx = 1.0 / h_K[0];
    p.ify = 1.0 / h_K[4];
    p.cx = h_K[2];
    p.cy = h_K[5];
    for (int i = 0; i < n_dist && i < 12; ++i) p.k[i] = h_D[i];
    p.has_dist = n_dist > 0 ? 1 : 0;  // (OpenCV runs the iteration whenever a coefficient vector is given, zeros included)
    for (int i = 0; i < 9; ++i) p.rr[i] = h_RR[i];
    p.w = w;
    p.h = h;
    unsigned* d_flag = nullptr;
    XM_CUDA(cudaMalloc(&d_flag, 4));
    XM_CUDA(cudaMemsetAsync(d_flag, 0, 4, s));
    const long long n = static_cast<long long>(w) * h;
    xm::build_inverse_lut_kernel<<<grid_for(n, 256, 1, 148 * 16), 256, 0, s>>>(p, d_mapx, d_mapy, d_xy_i16, d_flag);
    XM_LAUNCHED();
    unsigned h_flag = 0;
    cudaError_t e = cudaMemcpyAsync(&h_flag, d_flag, 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_flag);
    XM_CUDA(e);
    if (h_flag) return fail(XM_ERR_TABLE_RANGE, "rectification map does not fit int16");
    return XM_OK;
}

}  // extern "C"
