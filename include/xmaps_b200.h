/*
 * xmaps_b200.h — C ABI of the B200-native X-maps per-event depth path.
 *
 * The reference (fraunhoferhhi/X-maps) is pure Python and has no FFI of its own:
 * its boundary is a set of Python call signatures.  Each entry point below is therefore annotated
 * with the reference *Python* function it replaces (paths relative to the reference checkout);
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add to call it.
 *
 * Conventions
 *   - every function returns an int status (XM_OK == 0); no exception ever crosses the ABI;
 *     xm_last_error() returns a thread-local description of the last failure;
 *   - pointers named d_* are DEVICE pointers owned by the caller (e.g. torch tensor storage),
 *     pointers named h_* are HOST pointers; the context owns only its own tables and scratch;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     asynchronous with respect to the host and ordered on that stream unless stated otherwise;
 *   - one context per GPU per stream; a context is not thread-safe.
 *
 * Event record (Metavision EventCD, call sites python/trigger_finder.py:2,21 and
 * python/depth_reprojection_pipe.py:114): 16-byte array-of-structs
 *     offset 0  uint16 x      offset 2  uint16 y      offset 4  int16 p      offset 6  pad
 *     offset 8  int64  t  (microseconds)   — or float64 t when XM_FLAG_TIME_F64 is set
 *                                            (python/eval/compute_depth_x_maps.py:91).
 */
#ifndef XMAPS_B200_H
#define XMAPS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XM_ABI_VERSION 1

/* status codes */
enum {
    XM_OK = 0,
    XM_ERR_INVALID_ARG = 1, /* NULL pointer, negative size, unknown enum value             */
    XM_ERR_CUDA = 2,        /* a CUDA runtime call failed (see xm_last_error)               */
    XM_ERR_NO_XMAP = 3,     /* entry point needs the X-map and none was supplied            */
    XM_ERR_TABLE_RANGE = 4, /* table violates an assert of the reference (int16 capacity)   */
    XM_ERR_UNSUPPORTED = 5  /* shape not supported by this build                            */
};

/* output view: python/depth_reprojection_pipe.py:148-162 (params.camera_perspective) */
enum { XM_VIEW_PROJECTOR = 0, XM_VIEW_CAMERA = 1 };

/* what xm_frame writes */
enum {
    XM_OUT_DEPTH = 0,     /* float32 [H, W]     metric depth   (disparity_to_depth_rectified)   */
    XM_OUT_DISPARITY = 1, /* float32 [H, W]     disparity map  (compute_disp_map_* [+ remap])  */
    XM_OUT_BGR = 2        /* uint8   [H, W, 3]  colourised     (colorize_depth_from_disp)      */
};

/* how t.min() / t.max() of the frame (python/x_maps_disparity.py:12-13) are obtained */
enum {
    XM_TBOUNDS_REDUCE = 0, /* full device reduction over the (polarity-masked) events: always exact  */
    XM_TBOUNDS_SORTED = 1, /* first / last valid event; the main kernel verifies every event against */
                           /* them, raises XM_STATUS_TBOUNDS_VIOLATED if the stream was unsorted and  */
                           /* (option "auto_fixup") redoes the frame with reduced bounds              */
    XM_TBOUNDS_GIVEN = 2   /* caller supplies t_min / t_max; verified the same way                    */
};

/* XmFrameArgs.flags */
#define XM_FLAG_POLARITY 0x1u /* keep only p == 1 (PolarityFilterAlgorithm(1), ...pipe.py:43,114)  */
#define XM_FLAG_TIME_F64 0x2u /* the 8-byte time field holds a float64                              */
/* xm_frame only, opt-in, NOT reference behaviour (the reference's lookup is nearest, x_maps_disparity.py:19,25):
 * bilinear X-map lookup at the un-rounded (y_rect, time column) of the float32 rectification LUT (needs
 * XmTables.lut_x_f32 / lut_y_f32); undefined (0) cells are left out of the blend; float32 disparities, last event
 * per cell wins; 7x7 dilate / remap / depth as usual on the float map.  Staged kernels, no batch path.  */
#define XM_FLAG_BILINEAR 0x4u

/* XmFrameStatus.flags (device-side findings of the last frame) */
#define XM_STATUS_TBOUNDS_VIOLATED 0x1u /* an event lies outside the assumed [t_min, t_max]          */
#define XM_STATUS_PIXEL_OOB 0x2u        /* event pixel outside the camera image (ref: IndexError)   */
#define XM_STATUS_SCATTER_OOB 0x4u      /* scatter target outside the map       (ref: IndexError)   */
#define XM_STATUS_FILTER_POLARITY 0x8u  /* xm_filter_events YT: an event with p != 1 (ref: index arrays would not line up) */
#define XM_STATUS_FILTER_INDEX 0x10u    /* xm_filter_events YT: column outside the (y, x_rect) image (ref: IndexError)     */

/* XmFilterMode: the reference's frame_event_filter.py classes */
#define XM_FILTER_FIRST_YT 1 /* FirstEventPerYTFilter          :70-98   */
#define XM_FILTER_FIRST_XY 2 /* FirstEventPerXYFilter          :44-66   */
#define XM_FILTER_LAST_XY 3  /* LastEventPerXYFilter           :19-41   */
#define XM_FILTER_MEAN_XY 4  /* MeanFirstLastEventPerXYFilter  :101-128 */

typedef struct XmCtx XmCtx;

/* Host-side description of one calibration's read-only tables; copied to the device by
 * xm_ctx_create.  Field meaning follows the reference objects that own them. */
typedef struct XmTables {
    int32_t cam_w, cam_h;   /* camera image                (RuntimeParams.camera_*)            */
    int32_t rect_w, rect_h; /* rectified image             (calib.rect_image_*)                */
    int32_t proj_w, proj_h; /* projector-view output       (RuntimeParams.projector_*)         */
    int32_t xmap_w;         /* X_MAP_WIDTH                 (x_maps_disparity.py:58)            */
    int32_t t_px_scale;     /* T_PX_SCALE = X_MAP_WIDTH-1  (x_maps_disparity.py:59)            */
    int32_t x_offset;       /* X_OFFSET = 4242             (x_maps_disparity.py:49)            */
    int32_t dilate;         /* dilate kernel size, 7       (disp_to_depth.py:74)               */
    double depth_scale;     /* P2[0,3]                     (disp_to_depth.py:104-107)          */
    const int16_t* lut_x;   /* [cam_h, cam_w]  disp_cam_mapx_i16 (cam_proj_calibration.py:255) */
    const int16_t* lut_y;   /* [cam_h, cam_w]  disp_cam_mapy_i16                               */
    const int16_t* x_map;   /* [rect_h, xmap_w] proj_x_map, may be NULL (set later)            */
    const int16_t* remap_xy; /* [proj_h, proj_w, 2] disp_proj_mapxy_i16, may be NULL (camera view only) */
    const float* lut_x_f32; /* [cam_h, cam_w]  disp_cam_mapx_f32, may be NULL                  */
    const float* lut_y_f32; /* [cam_h, cam_w]  disp_cam_mapy_f32, may be NULL                  */
} XmTables;

typedef struct XmFrameArgs {
    const void* d_events; /* device, n_events * 16 bytes, 16-byte aligned                     */
    int64_t n_events;
    uint32_t flags;      /* XM_FLAG_*                                                        */
    int32_t view;        /* XM_VIEW_*                                                        */
    int32_t time_bounds; /* XM_TBOUNDS_*                                                     */
    int32_t output;      /* XM_OUT_*                                                         */
    int64_t t_min, t_max; /* XM_TBOUNDS_GIVEN: int64 values, or float64 bit patterns with XM_FLAG_TIME_F64 */
    void* d_out;         /* device output, see XM_OUT_*                                      */
    float z_near, z_far; /* XM_OUT_BGR only (DisparityToDepth.z_near / z_far)                */
} XmFrameArgs;

typedef struct XmFrameStatus {
    int64_t n_events;  /* records looked at                                                  */
    int64_t n_valid;   /* after the polarity mask                                            */
    int64_t n_inliers; /* events that produced a disparity (inlier_mask.sum())               */
    int64_t t_min, t_max; /* bounds used (int64, or float64 bit patterns)                     */
    uint32_t flags;    /* XM_STATUS_*                                                        */
    uint32_t epoch;    /* internal frame counter of the scatter map                          */
    uint32_t fixup_ran; /* 1: assumed bounds were wrong and the exact second pass produced the frame */
    uint32_t reserved;
} XmFrameStatus;

/* ---- library --------------------------------------------------------------------------- */
int xm_abi_version(void);
const char* xm_last_error(void);
/* number of kernel launches issued through this library since load (bench.py "gpu_launches") */
int64_t xm_launch_count(void);

/* ---- context --------------------------------------------------------------------------- */
/* Replaces the table ownership of CamProjMaps / XMapsDisparity / DisparityToDepth
 * (python/cam_proj_calibration.py:143-172, python/x_maps_disparity.py:35-67,
 * python/disp_to_depth.py:66-74): uploads and re-packs the tables for the kernels. */
int xm_ctx_create(const XmTables* tables, int device, XmCtx** out);
int xm_ctx_destroy(XmCtx* ctx);
/* (Re)place the X-map after creation (XMapsDisparity.__post_init__, x_maps_disparity.py:44-67). */
int xm_ctx_set_xmap(XmCtx* ctx, const int16_t* h_x_map, int32_t rows, int32_t cols, int32_t t_px_scale,
                    int32_t x_offset);
/* 256 x 3 uint8 BGR colour table used by XM_OUT_BGR / xm_colorize (the reference uses
 * cv2.COLORMAP_TURBO, disp_to_depth.py:36; the host side passes OpenCV's own table). */
int xm_ctx_set_colormap(XmCtx* ctx, const uint8_t* h_bgr256);
/* Tuning knobs; unknown keys return XM_ERR_INVALID_ARG.  Keys:
 *   "stage_xmap"      0/1  stage X-map time columns into shared memory with 1-D bulk (TMA) copies [1]
 *   "smem_cols_bytes" shared-memory budget per CTA for that window                              [12288]
 *   "stages"          depth of K1's shared-memory event ring (16 KB per stage, TMA-filled)      [2]
 *   "win_stages"      depth of the X-map window ring of the per-frame kernels                   [2]
 *   "batch_win_stages" ... of the batch kernel                                                  [3]
  *   "safe_tables"     0/1  allow the check-free scatter of K1 when the tables were verified at upload
 *                     (every defined X-map cell in [x_offset, x_offset + rect_w), LUT x > -x_offset) [1]
    *   "fused"           1: one fused kernel per frame (events + grid barrier + epilogue) where the lean path
 *                     applies (integer time, verified tables, 7x7 dilate); 0: separate K1 / K2 kernels [1]
 *   "pdl"             1: programmatic dependent launch between K1 / K2 / the next frame's K1 (used only
 *                     when "auto_fixup" is 0 or the bounds are exact: it cannot be combined with the
 *                     device-side fix-up launch)                                             [1]
 *   "k2_variant"      1: sliding-window projector epilogue (7x7 dilate, even rect_w), 0: per-tap [1]
 *   "lookahead"       extra columns fetched ahead of a time-sorted stream                       [1]
 *   "auto_fixup"      0/1  with XM_TBOUNDS_SORTED / _GIVEN: when an event lies outside the assumed
 *                     bounds, redo the frame on the device with exact (reduced) bounds         [1]
 *   "batch"           xm_frame_batch on uniform batches: 1 = batch_kernel (one persistent kernel per <= 32
 *                     frames: staged event pipeline + epilogue warp groups), 0 = the per-frame kernels
 *                     back to back                                                              [1]
 *   "scatter_aggregate" batch kernel, projector view: 1 = chunks in which nearly every event is live keep one RED per
 *                     distinct cell and warp round (pays on a scanning projector's stream, where consecutive events
 *                     share pixels; costs 25 % on a uniform stream), 0 = plain scatter, 2 = decide from the inlier
 *                     fraction of the last batch whose statistics reached the host (asynchronous)       [2]
 *   "batch_strips"    batch kernel, projector view: 1 = strip epilogue (two barrier-free passes, one warp per item:
 *                     decode + 7x7 dilation of the remap targets' window into a u16 map, then one gather per output
 *                     pixel), 0 = tile epilogue (32x32 output tiles through shared memory).  Identical results    [1]
 *   "batch_maps"      scatter maps the batch kernel rotates through, 2 ... 8; 0 = by events per frame (3 / 4 / 6)  [0]
 *   "tile_warps"      epilogue warps per CTA of the strip epilogue: 2, 4; 0 = two for large frames, four for small [0]
 *   "strip_rows", "strip_blocks"  item sizes of the strip epilogue (rows per pass-1 item, 256-pixel blocks per pass-2
 *                     item); 0 = 42 / 8 for small frames, 90 / 16 for large ones                                [0]
 *   "strip_lag"       blocks by which the pass-2 items of a frame trail its pass-1 items in the strip epilogue's
 *                     item list, 1 ... 3                                                                       [1]
 *   "reserve_sms"     SMs the persistent batch kernel leaves free (e.g. for NCCL's copy kernels)  [0]
 *   "ctas_per_sm"     resident CTAs per SM for the event kernel, 0 = occupancy query            [0]
 *   "region_cells"    shared-memory cells per buffer of the projector-view epilogue [largest tile region of
 *                     the remap table, rounded up to 128; read-only "region_need" = the exact figure]
 *   "profile"         1 / 0: record CUDA events around K1 (per-event kernel) and K2 (per-pixel
 *                     epilogue) of every frame; -1 resets the accumulators.  Read back with
 *                     "profile_k1_ns", "profile_k2_ns", "profile_frames", "profile_launches" (these
 *                     synchronise; a batch launch counts once in profile_launches, its frames in profile_frames)
 *   "alive"           1: the batch kernel consults a shared-memory table (per 8x8 block of camera pixels the hull of
 *                     the time columns at which an event of the block can be an inlier, derived exactly from the LUT
 *                     and the X-map at upload) and only counts / bounds-checks events outside it -- no LUT gather,
 *                     X-map lookup or scatter for them; 0: every event is looked up                             [1]
 *   "epoch"           test hook: clear the scatter map and set its 16-bit frame counter
 * read-only (xm_ctx_get_option): "cap_cols", "occupancy", "sm_count", "event_smem_bytes", "batch_occ",
 * "batch_smem", "batch_cols", "alive_px" (camera pixels inside alive blocks).  The environment variable XMAPS_B200_OPTS="key=value,key=value" applies options to
 * every context at creation (A/B runs without code changes). */
int xm_ctx_set_option(XmCtx* ctx, const char* key, int64_t value);
int xm_ctx_get_option(XmCtx* ctx, const char* key, int64_t* value);

/* ---- the hot path ---------------------------------------------------------------------- */
/* One projector frame, fused: polarity mask -> rectify LUT -> X-map lookup -> disparity ->
 * last-write-wins scatter -> [7x7 dilate + nearest remap] -> depth.  Replaces the body of
 * DepthReprojectionPipe.process_ev_frame (python/depth_reprojection_pipe.py:121-167), i.e. the
 * chain rectify_cam_coords_i16 (cam_proj_calibration.py:277-281), compute_event_disparity
 * (x_maps_disparity.py:9-32,69-82), compute_disp_map_{projector,camera}_view
 * (cam_proj_calibration.py:299-303,312-317), remap_rectified_disp_map_to_proj
 * (disp_to_depth.py:76-97) and disparity_to_depth_rectified (disp_to_depth.py:46-63);
 * with XM_OUT_BGR also colorize_depth_from_disp (disp_to_depth.py:99-115). */
int xm_frame(XmCtx* ctx, const XmFrameArgs* args, void* stream);
/* Same for `n_frames` independent frames on one stream (process_ev_frame keeps no state across frames).
 * Batches that are uniform -- one view / output / flag set / z range, integer timestamps, bounds SORTED or
 * GIVEN, verified tables -- are rendered by ONE persistent kernel per <= 32 frames (option "batch"): the
 * epilogue of a frame overlaps the event stream of the next ones, and frames whose assumed bounds turn out
 * wrong are re-rendered exactly on the device.  Anything else falls back to xm_frame per frame.  Results are
 * identical either way.  xm_frame_status afterwards reports the LAST frame of the batch. */
int xm_frame_batch(XmCtx* ctx, const XmFrameArgs* args, int32_t n_frames, void* stream);
/* Copies the status block of the most recent frame to the host; synchronises `stream`. */
int xm_frame_status(XmCtx* ctx, XmFrameStatus* h_status, void* stream);
/* Host-buffer variant (the call a CPU-side user of the reference makes): copies `h_events` to the
 * device, runs xm_frame, copies the result back into `h_out`, and synchronises.  Pinned host
 * memory (xm_host_alloc, cudaHostAlloc, torch pin_memory) gives full PCIe speed. */
int xm_frame_host(XmCtx* ctx, const XmFrameArgs* args /* d_events/d_out ignored */, const void* h_events,
                  void* h_out, XmFrameStatus* h_status /* may be NULL */, void* stream);
int xm_host_alloc(void** h_ptr, int64_t bytes);
int xm_host_free(void* h_ptr);

/* ---- the same path, stage by stage (materialised intermediates) -------------------------- */
/* CamProjMaps.rectify_cam_coords_i16 / _f32 (cam_proj_calibration.py:272-281). */
int xm_rectify_i16(XmCtx* ctx, const void* d_events, int64_t n, int16_t* d_x_rect, int16_t* d_y_rect, void* stream);
int xm_rectify_f32(XmCtx* ctx, const void* d_events, int64_t n, float* d_x_rect, float* d_y_rect, void* stream);
/* compute_disparity (x_maps_disparity.py:9-32): per-event, un-compacted.  d_disp_full[i] is the
 * disparity of event i or -1; d_mask[i] is the inlier mask (over the polarity-masked events when
 * XM_FLAG_POLARITY is set, masked-out events get 0).  d_x_rect / d_y_rect may be NULL (computed
 * from the LUT) or the arrays returned by xm_rectify_i16.  Only d_events, n_events, flags,
 * time_bounds, t_min, t_max of `args` are used. */
int xm_event_disparity(XmCtx* ctx, const XmFrameArgs* args, const int16_t* d_x_rect, const int16_t* d_y_rect,
                       int16_t* d_disp_full, uint8_t* d_mask, void* stream);
/* Order-preserving compaction out[k] = vals[i] for the k-th i with mask[i] != 0 (the
 * `disp[disp_inlier_mask]` of x_maps_disparity.py:32); *d_count receives the number kept. */
int xm_compact_i16(XmCtx* ctx, const int16_t* d_vals, const uint8_t* d_mask, int64_t n, int16_t* d_out,
                   int64_t* d_count, void* stream);
/* `m = zeros(h, w); m[rows, cols] = vals` with NumPy's last-write-wins rule
 * (cam_proj_calibration.py:301-302,315-316).  (h, w) must be the rectified or the camera size. */
int xm_scatter_last_wins(XmCtx* ctx, const int16_t* d_rows, const int16_t* d_cols, const int16_t* d_vals,
                         int64_t n, int32_t h, int32_t w, float* d_map, void* stream);
/* DisparityToDepth.remap_rectified_disp_map_to_proj (disp_to_depth.py:76-97) on a materialised
 * float32 [rect_h, rect_w] map -> float32 [proj_h, proj_w]. */
int xm_dilate_remap(XmCtx* ctx, const float* d_rect_map, float* d_proj_map, void* stream);
/* disparity_to_depth_rectified (disp_to_depth.py:46-63): float32 [n] -> float32 [n]. */
int xm_disp_to_depth(XmCtx* ctx, const float* d_disp, int64_t n, double depth_scale, float* d_depth, void* stream);
/* colorize_depth_from_disp (disp_to_depth.py:99-115): float32 disparity [n] -> uint8 BGR [n, 3]. */
int xm_colorize(XmCtx* ctx, const float* d_disp, int64_t n, double depth_scale, float z_near, float z_far,
                uint8_t* d_bgr, void* stream);
/* CamProjMaps.construct_point_cloud (cam_proj_calibration.py:319-331): float32 [n] x3 -> [n, 3]. */
int xm_point_cloud(XmCtx* ctx, const float* d_x, const float* d_y, const float* d_disp, int64_t n,
                   const double* h_Q /* 4x4 row-major */, float* d_xyz, void* stream);

/* ---- the rows either side of the path ("next" rows N4, N2) -------------------------------- */
/* PolarityFilterAlgorithm(1).process_events (Metavision; call site depth_reprojection_pipe.py:43,114) ==
 * events[events["p"] == 1] (frame_event_filter.py:21), order preserved; d_out needs room for n records.
 * (xm_frame fuses this mask; this entry point materialises it for the stream in front of the trigger finder.) */
int xm_polarity_filter(XmCtx* ctx, const void* d_events, int64_t n, void* d_out, int64_t* d_count, void* stream);
/* ActivityNoiseFilterAlgorithm(width, height, threshold_us).process_events (Metavision, closed binary; constructed at
 * depth_reprojection_pipe.py:65-67 with threshold = 1e6 / projector_fps, applied to every packet at :116-117, between the
 * polarity filter and the trigger finder).  Restated from its published semantics ("parity unpinned", see the oracle): an
 * event passes iff one of the 8 neighbours of its pixel saw an event less than threshold_us earlier; per-pixel
 * timestamps start at 0 and are carried from call to call (xm_activity_reset clears them).  Events must be sorted by
 * time (int64 timestamps); survivors keep their order; d_out needs room for n records.  The call synchronises the
 * stream (it reads the packet's sub-division and the sortedness flag back). */
int xm_activity_filter(XmCtx* ctx, const void* d_events, int64_t n, int64_t threshold_us, void* d_out, int64_t* d_count, void* stream);
int xm_activity_reset(XmCtx* ctx, void* stream);
/* FrameEventFilter.filter_events (python/frame_event_filter.py:19-128): one survivor per key -- pixel
 * (x, y), or (y, rectified x) for XM_FILTER_FIRST_YT, which needs d_x_rect = rectify_cam_coords_i16 of the
 * frame -- written as EventCD records (p = 1, t wrapped to int32 as the reference's int32 images do) in
 * row-major key order.  d_out needs room for min(n, cam_h * key-image width) records; *d_count receives
 * the number written.  as_reference != 0 reproduces the reference as it runs (its "first" filters assign
 * reversed NumPy views, which NumPy walks in memory order again, so they keep the LAST event per key);
 * as_reference == 0 keeps the first event, the documented intent.  Only events with p == 1 take part. */
int xm_filter_events(XmCtx* ctx, const void* d_events, int64_t n, int32_t mode, const int16_t* d_x_rect,
                     int32_t as_reference, void* d_out, int64_t* d_count, void* stream);
/* RobustTriggerFinder.find_trigger (python/trigger_finder.py:146-189) on one time-ordered buffer:
 * pauses = events followed by a gap >= pause_thresh_us; the first two consecutive pauses further apart
 * than frame_us / 2 decide.  d_result[0..5] (device, int64) = status, prev_idx, next_idx, n_pauses,
 * t[prev_idx + 2], t[next_idx - 2];  status 1: frame = events[prev_idx + 2 : next_idx - 2], the caller
 * keeps events[next_idx - 2 :];  0: candidate rejected, keep events[next_idx :];  -1: none, keep nothing. */
int xm_find_trigger(XmCtx* ctx, const void* d_events, int64_t n, int64_t pause_thresh_us, double frame_us,
                    int64_t min_events, int64_t* d_result /* [8] */, void* stream);

/* ---- multi-GPU: depth frames gathered without a collective kernel ---------------
 * One process per GPU, frames sharded round-robin; the path itself exchanges nothing.  The only transfer is the
 * final gather of finished depth frames on one rank.  The reference has no multi-GPU code (nothing to mirror);
 * these calls let rank r place its frames straight into the gathering rank's memory over NVLink / NVSwitch:
 *   - the gathering rank allocates its slab with xm_peer_alloc (a plain cudaMalloc, so the IPC handle refers to
 *     the allocation base) and sends the 64-byte handle to its peers (any host channel, e.g. torch.distributed);
 *   - a peer maps it with xm_peer_open (peer access to `owner_device` is enabled explicitly and checked) and
 *     either passes the mapped pointer as XmFrameArgs.d_out -- the epilogue warps of the frame / batch kernels
 *     then store finished tiles directly into the gathering rank's HBM, a fused compute + gather with no extra
 *     kernel and no SM taken from the persistent kernel -- or copies finished frames with xm_peer_copy
 *     (cudaMemcpyPeerAsync: copy engines, no SM either). */
typedef struct XmIpcHandle {
    unsigned char bytes[64]; /* cudaIpcMemHandle_t */
} XmIpcHandle;
int xm_peer_alloc(int device, int64_t bytes, void** d_ptr, XmIpcHandle* handle);
int xm_peer_free(int device, void* d_ptr);
int xm_peer_open(int device, int owner_device, const XmIpcHandle* handle, void** d_ptr);
int xm_peer_close(int device, void* d_ptr);
int xm_peer_copy(void* d_dst, int dst_device, const void* d_src, int src_device, int64_t bytes, void* stream);
/* can_access[0] = 1 if `device` can map memory of `peer_device`, perf_rank[0] = cudaDevP2PAttrPerformanceRank */
int xm_peer_info(int device, int peer_device, int32_t* can_access, int32_t* perf_rank);

/* ---- set-up time ("next" row N3) --------------------------------------------------------- */
/* compute_x_map_from_time_map (python/x_map.py:5-55): float32 time map [h, w] (device) ->
 * int16 X-map [h, x_map_width] (+ optional float32 t_diffs). */
int xm_build_xmap(int device, const float* d_time_map, int32_t h, int32_t w, int32_t x_map_width,
                  int32_t t_px_scale, int32_t x_offset, int32_t num_scanlines, int16_t* d_x_map,
                  float* d_t_diffs /* may be NULL */, void* stream);

/* initUndistortRectifyMapInverse (python/cam_proj_calibration.py:31-41; builds disp_cam_map{x,y} at :246-254 and
 * disp_proj_mapxy at :262-270) = cv2.undistortPoints over the whole w x h pixel grid, on the device: float32 maps
 * [h, w] (either may be NULL) and/or the interleaved int16 (x, y) table [h, w, 2] = mapf_to_i16 of them (:44-48).
 * h_K: 3x3 camera matrix (row-major), h_D: n_dist distortion coefficients in OpenCV order (0, 4, 5, 8, 12 or 14; the
 * tilt terms must be 0), h_RR: the 3x3 product P[:, :3] * R the way OpenCV forms it (cv2.gemm).  Bit-identical to
 * OpenCV's result (same float64 operations in the same order, 5 iterations).  Synchronises the stream. */
int xm_build_inverse_lut(int device, const double* h_K, const double* h_D, int32_t n_dist, const double* h_RR, int32_t w,
                         int32_t h, float* d_mapx, float* d_mapy, int16_t* d_xy_i16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XMAPS_B200_H */
