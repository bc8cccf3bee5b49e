// spano_internal.h -- shared declarations of the sm_100a compositing library (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/spano.h"

#define SPANO_BLUR_RADIUS_MAX 32 /* ceil(3*sigma) supported by the blend kernels (sigma <= 10) */

// Derived projector state handed to kernels (== cv::detail::ProjectorBase after setCameraParams).
struct SpanoProjector {
    int kind;
    float scale;
    float k[9], rinv[9], r_kinv[9], k_rinv[9];
};

// host, bit-exact with OpenCV (projector_host.cpp; compiled with -ffp-contract=off)
void spano_host_set_camera(SpanoProjector *p, int kind, float scale, const float *K, const float *R);
void spano_host_map_forward(const SpanoProjector *p, float x, float y, float *u, float *v);
void spano_host_roi(const SpanoProjector *p, int src_w, int src_h, int roi[4]);
// finish a ROI from float extremes found elsewhere (stereographic GPU scan + exact host refinement)
void spano_host_gaussian_taps(int n, double sigma, float *taps);

struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
};

struct spano_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    // host-buffer fused path: uploads run on their own stream, double-buffered against the compute stream
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_start = nullptr;
    // fused path: warp + mask of the next image run on this stream while the current image is blended
    cudaStream_t d2h_stream = nullptr; // host-buffer fused path: finished canvas columns are downloaded while blending continues
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_warped[2] = {nullptr, nullptr}, ev_blended[2] = {nullptr, nullptr}, ev_start2 = nullptr;
    std::mutex mu;
    std::string err;
    long long launches = 0;
    // grow-only scratch buffers, indexed by role
    enum { BUF_LABELS = 0, BUF_DARK, BUF_MASK0, BUF_TABLES, BUF_ACC, BUF_TILE, BUF_TILEMASK, BUF_CUTMASK, BUF_SRC,
           BUF_CANVAS, BUF_MISC, BUF_SRC2, BUF_CUT2, BUF_RESIZE, BUF_CUTSMALL, BUF_CUTSMALL2, BUF_FIELD, BUF_FIELD2, BUF_BLENDPLAN, BUF_BLENDPLAN2, BUF_DT_TMP, BUF_DT_JOBS, BUF_DT_JOBS2, BUF_DT_ARENA, BUF_PREP_CUT, BUF_PREP_PLAN, BUF_RESIZE_AUX, BUF_TABLES_AUX, BUF_PRETILE, BUF_EQUALIZE, BUF_COUNT };
    DeviceBuffer buf[BUF_COUNT];
    std::vector<void *> owned; // extra allocations freed at destroy / end of call
    unsigned long long *blend_stats = nullptr; // device: [0] tile pixels the blend processed, [1] tile pixels offered
    // incremental band blend (spano_dev_blend_begin / add / finish)
    struct BlendSession {
        bool open = false;
        int cw = 0, mx = 0, my = 0, row0 = 0, row1 = 0, bands = 0, radius = 0;
        float4 *acc = nullptr;
        // host variant: preview-scale masks uploaded at begin (host pointer -> device copy, step)
        struct Staged { const uint8_t *host; const uint8_t *dev; size_t step; };
        std::vector<Staged> staged;
        // host variant with the canvas announced at begin: canvas columns that no later image touches are
        // normalised and downloaded on the download stream while the remaining images are still blended
        struct ColRun { int c0, c1, idx; bool flushed; };
        std::vector<ColRun> runs;
        const spano_image_desc *images = nullptr;
        int n_images = 0;
        uint8_t *h_canvas = nullptr;
        size_t h_step = 0;
        uint8_t *d_canvas = nullptr;
        size_t d_step = 0;
        std::vector<cudaEvent_t> events;
        // spano_*_blend_prepare: per announced image, mask_cut resized for the slice rows and the blend's sparsity plan,
        // made ahead of time on the auxiliary stream (ready = event recorded there)
        struct Prepared { const spano_image_desc *im; const uint8_t *cut_v; size_t cut_step; const int *plan; cudaEvent_t ready; };
        std::vector<Prepared> prepared;
    } bs;
    std::vector<cudaEvent_t> event_pool;   // reusable events of the prepare step
    int tap_bands = 0;                     // Gaussian taps this context blends with (launch_blend_setup), per band and
    double tap_sigma = 0.0;                // distance from the centre; copied into the parameters of every blend launch
    float taps[SPANO_MAX_BANDS][SPANO_BLUR_RADIUS_MAX + 1] = {};
    int tap_slot = -1;                     // marching kernel: constant-memory slot holding these taps (-1: none)
    // spano_set_option
    int opt_blend_dense = 0;               // 1: ignore the mask_cut sparsity (every tile pixel is filtered)
    int opt_warp_kernel = 0;               // 1: the TMA-staged warp kernel instead of the plain (un-staged) one
    int opt_flag_wait = 0;                 // 1: wait for readiness flags with a polling kernel instead of cuStreamWaitValue32
    int opt_blend_kernel = 0;              // 1: always the generic-radius blend kernel (cross-check of the marching one)
    // timers
    bool timers_on = false;
    float stage_ms[4] = {0, 0, 0, 0};
    long long stage_launches[4] = {0, 0, 0, 0};
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending_events;
};

int spano_fail(spano_ctx *ctx, int code, const char *fmt, ...);
// Small per-launch table buffers (resize axis tables, warp trig tables) are rewritten by every launch on whatever stream
// the context currently enqueues on; work on the auxiliary stream (spano_*_blend_prepare, the warp-ahead of the fused
// path) gets its own copy so that it cannot race with launches of the same kind on the main stream.
inline int spano_table_buffer(const spano_ctx *ctx, int main_id, int aux_id) { return (ctx->aux_stream && ctx->stream == ctx->aux_stream) ? aux_id : main_id; }
int spano_reserve(spano_ctx *ctx, int which, size_t bytes, void **out);

#define SPANO_CUDA(ctx, call)                                                                      \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return spano_fail(ctx, SPANO_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                              __LINE__);                                                           \
    } while (0)

// Row-range scatter table of the tile-sharded multi-GPU path: a kernel that produces tile rows writes row v to
// every entry with row0 <= v < row1, at base + v * step (base is the VIRTUAL address of tile row 0, i.e. the
// slice's storage minus row0 * step).  The bases may be peer-GPU memory (NVLink P2P stores).
#define SPANO_MAX_SLICES 16
struct SpanoScatter {
    int n = 0;
    int row0[SPANO_MAX_SLICES], row1[SPANO_MAX_SLICES];
    int col0[SPANO_MAX_SLICES], col1[SPANO_MAX_SLICES];   // tile columns the slice holds (multiples of 32, or the tile width)
    uint8_t *base[SPANO_MAX_SLICES];                       // VIRTUAL address of tile pixel (0, 0) in the slice's storage
    size_t step[SPANO_MAX_SLICES];
};

// ---- kernel launchers (each returns the number of kernels launched, or <0) -------------
// warp_kernels.cu
int launch_warp(spano_ctx *ctx, const SpanoProjector &P, const uint8_t *src, int src_w, int src_h, size_t src_step,
                double gain, int tl_x, int tl_y, int dst_w, int dst_h, int row_begin, int row_end, uint8_t *dst,
                size_t dst_step, uint8_t *dark, size_t dark_step, float *xmap = nullptr, float *ymap = nullptr,
                const SpanoScatter *scatter = nullptr);
int launch_remap(spano_ctx *ctx, const uint8_t *src, int src_w, int src_h, size_t src_step, const float *xmap,
                 const float *ymap, int dst_w, int dst_h, uint8_t *dst, size_t dst_step);
int launch_gain(spano_ctx *ctx, uint8_t *img, int w, int h, size_t step, double gain);
int launch_dark_flags(spano_ctx *ctx, const uint8_t *bgr, int w, int h, size_t step, uint8_t *dark, size_t dark_step);
// mask_kernels.cu
int launch_valid_mask(spano_ctx *ctx, const uint8_t *dark, int w, int h, size_t dark_step, int erode_iters,
                      uint8_t *mask, size_t mask_step, const SpanoScatter *scatter = nullptr);
// blend_kernels.cu
struct BlendTile {
    const uint8_t *tile;   size_t tile_step;   // 8UC3, gain applied
    const uint8_t *cut;    size_t cut_step;    // mask_cut 8UC1
    const uint8_t *valid;  size_t valid_step;  // validity mask 8UC1
    int w, h;                                  // full tile extent (reflect borders refer to it)
    int cx, cy;                                // tile corner in canvas coordinates
};
int launch_blend_setup(spano_ctx *ctx, int bands, double sigma);
int launch_blend_clear(spano_ctx *ctx, float4 *acc, int canvas_w, int rows);
// `plan`: device plan made beforehand with launch_blend_plan for the same tile, band and row range (marching
// kernel only; blend_plan_bytes() == 0 when another kernel will run); nullptr = plan at launch.
int launch_blend_tile(spano_ctx *ctx, const BlendTile &t, int bands, int radius, float4 *acc, int canvas_w, int row0,
                      int row1, const int *plan = nullptr);
size_t blend_plan_bytes(spano_ctx *ctx, int w, int bands, int radius);
// canvas_w: width of the accumulator the tile will be blended into (tile columns outside [0, canvas_w) - cx are clipped, as
// launch_blend_tile clips them: the plan must be made for the same window)
int launch_blend_plan(spano_ctx *ctx, const BlendTile &t, int bands, int radius, int row0, int row1, int *plan, int canvas_w);
// mask_cut still at preview scale (small: sw x sh, device): the plan is made from the preview mask and only the rows of the
// tile-size mask that the plan's pieces read are up-scaled, into t.cut (pitch t.cut_step; the rest of that buffer is left
// untouched and is never read by a blend launched with this plan).  Returns 0 when no plan applies (the caller then
// up-scales the whole mask with launch_resize_mask).
int launch_blend_plan_preview(spano_ctx *ctx, const BlendTile &t, const uint8_t *small, int sw, int sh, size_t sstep, int bands, int radius,
                              int row0, int row1, int *plan, int canvas_w);
int launch_normalise(spano_ctx *ctx, const float4 *acc, int canvas_w, int rows, int bands, int out_kind, void *out,
                     size_t out_step, int col0 = 0, int col1 = -1);
int launch_fp32_peak(spano_ctx *ctx, int variant, double *tflops);
// resize_kernels.cu: cv::resize(CV_8UC1, INTER_LINEAR) of the preview-scale seam masks
// rows [row_begin,row_end) of the dw x dh result are written at dst + row * dstep (row_end < 0: all rows)
int launch_resize_mask(spano_ctx *ctx, const uint8_t *src, int sw, int sh, size_t sstep, uint8_t *dst, int dw, int dh,
                       size_t dstep, int row_begin = 0, int row_end = -1);
int launch_adjust_intensity(spano_ctx *ctx, uint8_t *bgr, int w, int h, size_t step, const float *field, int fw, int fh,
                            size_t fpitch_elems);
// The same up-scaling split in two, for the blend's sparsity plan (launch_blend_plan_preview):
// (a) per strip of strip_w result columns the first / last result row in [ra, rb) that has a non-zero source tap -- a superset
//     of the rows where the up-scaled mask is non-zero (its coefficients are >= 0) -- merged into ymin / ymax (device, one int
//     per strip, initialised to INT_MAX / -1 by the caller);
// (b) the up-scaling itself, restricted per strip to the result rows [need0[s], need1[s]) (device arrays): nothing else is
//     written.
int launch_resize_activity(spano_ctx *ctx, const uint8_t *src, int sw, int sh, size_t sstep, int dw, int dh, int strip_w, int ra, int rb,
                           int *ymin, int *ymax);
int launch_resize_mask_rows(spano_ctx *ctx, const uint8_t *src, int sw, int sh, size_t sstep, uint8_t *dst, int dw, int dh, size_t dstep,
                            int strip_w, const int *need0, const int *need1);
// dist_kernels.cu: 5x5 chamfer distance transform (cv::distanceTransform DIST_L2 / DIST_MASK_5) and dcut::dist_cut
int launch_distance_transform(spano_ctx *ctx, int n, const uint8_t *const *masks, const size_t *msteps, const int *w, const int *h,
                              float *const *dist, const size_t *dsteps);
int launch_dist_cut(spano_ctx *ctx, int n, const uint8_t *const *masks, const size_t *msteps, const float *const *dist,
                    const size_t *dsteps, const int *tl_x, const int *tl_y, const int *w, const int *h, uint8_t *const *cut,
                    const size_t *csteps);
int launch_simple_blend(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tsteps, const float *const *dist,
                        const size_t *dsteps, const int *ax, const int *ay, const int *w, const int *h, float4 *acc, int canvas_w,
                        int canvas_h, uint8_t *out, size_t ostep);
int launch_no_blend(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tsteps, const uint8_t *const *masks,
                    const size_t *msteps, const int *ax, const int *ay, const int *w, const int *h, int canvas_w, int canvas_h,
                    uint8_t *out, size_t ostep);
int launch_gray(spano_ctx *ctx, const uint8_t *bgr, size_t step, int w, int h, uint8_t *gray, size_t gstep);
int launch_overlap_sums(spano_ctx *ctx, const uint8_t *gi, size_t gis, const uint8_t *mi, size_t mis, const uint8_t *gj, size_t gjs,
                        const uint8_t *mj, size_t mjs, int xi, int yi, int xj, int yj, int ow, int oh, unsigned long long *acc);
// mask_kernels.cu: readiness flags of the tile-sharded path (stores to possibly peer-GPU counters; polling wait)
#define SPANO_MAX_FLAG_TARGETS 16
int launch_flag_signal(spano_ctx *ctx, uint32_t *const *targets, int n, uint32_t value);
int launch_flag_wait_kernel(spano_ctx *ctx, const uint32_t *flag, uint32_t value);
// equalize_kernels.cu: test::equalizeIntensities (device buffers at preview scale)
void spano_equalize_field_size(int w, int h, float ratio, int *fw, int *fh);
int launch_equalize_intensities(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tsteps, const uint8_t *const *masks,
                                const size_t *msteps, const int *tl_x, const int *tl_y, const int *w, const int *h, float ratio,
                                float *const *out, const size_t *opitch);
// disk_kernels.cu: stereographic centre fix (util::RadialNormalizer state + normalised radius)
struct SpanoDiskParams {
    float cx, cy, scale;
    float radius_n;
    int quadratic;
};
int spano_disk_plan(int n, const int *tl_x, const int *tl_y, const int *w, const int *h, int ansatz_x, int ansatz_y,
                    float radius, int quadratic, SpanoDiskParams *P, int *org_x, int *org_y, int *new_x, int *new_y,
                    int *new_w, int *new_h);
int launch_disk_gather(spano_ctx *ctx, const SpanoDiskParams &P, const uint8_t *src, int sw, int sh, size_t sstep, int ox,
                       int oy, uint8_t *dst, int dw, int dh, size_t dstep, int dx0, int dy0);
