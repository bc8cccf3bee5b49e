/*
 * spano.h -- C ABI of the B200-native compositing path for SimplePanorama.
 *
 * This is the drop-in boundary: the reference's unchanged C++ entry points in
 * src/classes (stitch_parameters::return_full / get_preview / blend,
 * src/classes/_panorama.cpp:161-354) keep calling their callees
 *   proj::get_proj_parameters / projection::project   src/math/_projection.h:53-60,155-161
 *   blnd::createSurroundingMask + cv::erode            src/math/_blending.h, _projection.cpp:441-443
 *   imgs[i] / gain[i]                                  src/classes/_panorama.cpp:321-327
 *   blnd::multi_blend                                  src/math/_blending.h:24
 *   util::get_pan_dimension                            src/system/_util.cpp:204-231
 * and those callee bodies forward to the functions below (shim/ shows the cv::Mat glue,
 * INTEGRATION.md the build wiring).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no C++/torch/OpenCV types.
 *   - every function returns 0 on success, <0 on error (SPANO_E_*); it never throws.
 *     spano_last_error(ctx) gives the message of the last failure on that context.
 *   - pointers are borrowed for the duration of the call; outputs are caller-allocated
 *     after a size query (spano_warp_roi / spano_pan_dimension).
 *   - images are row-major, interleaved BGR uint8 ("8UC3"), `step` = bytes per row.
 *   - one spano_ctx per panorama object / thread; calls on one ctx are serialised by an
 *     internal mutex, different contexts may be used concurrently and share no mutable state
 *     (the Gaussian taps of a blend travel in the parameters of its kernel launches).
 *   - functions named spano_dev_* take DEVICE pointers and enqueue on the context's
 *     stream without synchronising (the caller owns the stream, see spano_set_stream);
 *     the others take HOST pointers and return when the result is in host memory.
 *   - there is no CPU fallback: without a CUDA device spano_create fails.
 */
#ifndef SPANO_H
#define SPANO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 100: round 1.  200: spano_slice grew by col0/col1 (column bands); spano_set_option, spano_*_composite_fixed, spano_shard_step_*,
 * spano_equalize_intensities added.  A host compiled against another major (hundreds) must not load this library. */
#define SPANO_VERSION 200

/* projection kinds == pan::Projection order used by the reference's projector classes */
#define SPANO_SPHERICAL 0     /* proj::spherical_proj   -> cv::detail::SphericalWarper    */
#define SPANO_CYLINDRICAL 1   /* proj::cylindrical_proj -> cv::detail::CylindricalWarper  */
#define SPANO_STEREOGRAPHIC 2 /* proj::sten_proj        -> cv::detail::StereographicWarper */

#define SPANO_OK 0
#define SPANO_E_INVALID (-1)  /* bad argument (what the reference would CV_Assert / throw on) */
#define SPANO_E_CUDA (-2)     /* CUDA runtime error                                            */
#define SPANO_E_NOMEM (-3)    /* device or host allocation failed                              */
#define SPANO_E_NODEVICE (-4) /* no usable CUDA device: the library has no CPU path            */
#define SPANO_E_LIMIT (-5)    /* exceeds a limit of the reference path (e.g. remap's 32767 px) */

#define SPANO_MAX_BANDS 10    /* UI caps bands at one digit; divisor 255/bands needs bands <= 255 */

/* output kinds of the blend */
#define SPANO_OUT_F32 0 /* CV_32FC3 as returned by blnd::multi_blend                          */
#define SPANO_OUT_U8 1  /* CV_8UC3 as returned by stitch_parameters::blend (x255, convertTo)  */

typedef struct spano_ctx spano_ctx;

/* ---- lifetime --------------------------------------------------------------------- */
int spano_version(void);
int spano_create(spano_ctx **out, int device);
void spano_destroy(spano_ctx *ctx);
const char *spano_last_error(spano_ctx *ctx);
/* Use an external CUDA stream (cudaStream_t as void*) for all work of this context;
 * NULL restores the context's own stream. */
int spano_set_stream(spano_ctx *ctx, void *cuda_stream);
/* Block until everything enqueued on the context's stream has finished. */
int spano_sync(spano_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
long long spano_launch_count(spano_ctx *ctx);
/* Per-context options (measurement and cross-check switches; they never change a result beyond the tolerances
 * stated: BLEND_DENSE is bit-identical, BLEND_KERNEL selects a kernel with a different summation order).
 *   SPANO_OPT_BLEND_DENSE  1: the multiband blend filters every tile pixel instead of skipping those whose whole
 *                             window of mask_cut is zero (see spano_blend_stats); default 0
 *   SPANO_OPT_BLEND_KERNEL 0 (default): the warp-specialised marching-strip kernel for sigma = 7, the generic-radius
 *                             kernel otherwise; 1: always the generic-radius kernel; 2: the 8-warp marching-strip kernel
 *                             of round 1; 3: the warp-specialised kernel with 12 instead of 8 consumer warps; 4: the
 *                             warp-specialised kernel without the register re-split (2, 3 and 4 are measurement
 *                             variants: same arithmetic in the same order as the default, bit-identical canvas) */
#define SPANO_OPT_BLEND_DENSE 1
#define SPANO_OPT_BLEND_KERNEL 2
#define SPANO_OPT_WARP_KERNEL 4 /* 1: the warp kernel stages the source footprint of each block of destination pixels with TMA
                                   (cp.async.bulk.tensor.2d) and samples from shared memory; 0 (default): global loads only.
                                   Same integer arithmetic: bit-identical tiles.                                           */
#define SPANO_OPT_FLAG_WAIT 3 /* 0 (default): stream memory operations; 1: a one-thread polling kernel (spano_shard_step_*) */
int spano_set_option(spano_ctx *ctx, int option, int value);

/* ---- a3: projector geometry (host arithmetic, bit-exact with OpenCV's warpers) -----
 * Replaces cv::detail::RotationWarperBase::detectResultRoi[ByBorder] and
 * SphericalWarper::detectResultRoi as reached from projection::project
 * (src/math/_projection.cpp:51,81,321).  K and R are the float32 3x3 matrices the
 * reference hands to OpenCV, i.e. K is already K_adj (principal point flipped,
 * _projection.cpp:41-44).  The warped tile is (roi + 1) in both dimensions and its
 * corner is the truncated ROI top-left, exactly like RotationWarperBase::warp.
 * Pure host arithmetic: ctx may be NULL.                                                 */
int spano_warp_roi(spano_ctx *ctx, int proj, float scale, const float K[9], const float R[9], int src_w, int src_h,
                   int *tl_x, int *tl_y, int *dst_w, int *dst_h);

/* a9: util::get_pan_dimension (src/system/_util.cpp:204-231). */
int spano_pan_dimension(int n, const int *tl_x, const int *tl_y, const int *w, const int *h, int *canvas_w,
                        int *canvas_h, int *min_x, int *min_y);

/* ---- a3 + a4 + a6: warp one image (HOST buffers) -------------------------------------
 * = projection::project + createSurroundingMask(warped,true,1) + cv::erode(3 iterations)
 *   + `warped / gain` (src/math/_projection.cpp:422-454, src/classes/_panorama.cpp:321-327).
 * dst must be dst_w x dst_h from spano_warp_roi.  The validity mask is computed from the
 * un-gained warp, as in the reference; gain is applied afterwards (gain = 1.0: none).
 * dst_valid_mask may be NULL (get_masks = false).                                        */
int spano_warp(spano_ctx *ctx, int proj, float scale, const float K[9], const float R[9], const uint8_t *src_bgr,
               int src_w, int src_h, size_t src_step, double gain, uint8_t *dst_bgr, size_t dst_step,
               uint8_t *dst_valid_mask, size_t mask_step);

/* The two halves of the warp on their own (HOST buffers), for unit parity and for callers that
 * use them directly: cv::detail::RotationWarperBase::buildMaps (float maps of a w x h tile whose
 * corner is (tl_x, tl_y)) and cv::remap(src, dst, xmap, ymap, INTER_LINEAR, BORDER_CONSTANT)
 * with OpenCV's 8-bit fixed-point sampler (the reference calls cv::remap itself at
 * src/math/_projection.cpp:278).                                                         */
int spano_build_maps(spano_ctx *ctx, int proj, float scale, const float K[9], const float R[9], int tl_x, int tl_y,
                     int w, int h, float *xmap, float *ymap);
int spano_remap(spano_ctx *ctx, const uint8_t *src_bgr, int src_w, int src_h, size_t src_step, const float *xmap,
                const float *ymap, int dst_w, int dst_h, uint8_t *dst_bgr, size_t dst_step);

/* a4 alone: blnd::createSurroundingMask(img,true,1) followed by `erode_iters` 3x3 erosions. */
int spano_surrounding_mask(spano_ctx *ctx, const uint8_t *bgr, int w, int h, size_t step, int erode_iters,
                           uint8_t *mask, size_t mask_step);

/* cv::resize(src, dst, Size(dst_w, dst_h)) for CV_8UC1 with the default INTER_LINEAR -- what return_full
 * applies to every mask_cut (src/classes/_panorama.cpp:329-335; the INTER_CUBIC there lands in `fx`).
 * OpenCV's fixed-point arithmetic, bit-exact.  HOST buffers.                                          */
int spano_resize_mask(spano_ctx *ctx, const uint8_t *src, int src_w, int src_h, size_t src_step, uint8_t *dst,
                      int dst_w, int dst_h, size_t dst_step);

/* test::adjust_intensity for one image, in place on CV_8UC3 (src/test/_test.cpp:110-122): the CV_32FC1
 * correction field (field_w x field_h, field_step in bytes) is resized to the image with cv::resize's
 * float INTER_LINEAR and the image becomes  sat_u8(rint((v/255 / clamp(field)) * 255)).  HOST buffers. */
int spano_adjust_intensity(spano_ctx *ctx, uint8_t *bgr, int w, int h, size_t step, const float *field, int field_w,
                           int field_h, size_t field_step);

/* test::equalizeIntensities(images, masks, top_lefts, ratio) (src/test/_test.cpp:9-106), the solve that produces the
 * correction fields spano_adjust_intensity consumes: called by stitch_parameters::set_config on the preview-size warps
 * when conf.blend_intensity is on (src/classes/_panorama.cpp:131-133).  tiles 8UC3, masks 8UC1 (validity masks), one
 * CV_32FC1 field per image of the size spano_equalize_intensities_size gives (cv::resize(.., Size(), ratio, ratio)).
 * field_steps in BYTES.  HOST buffers.  The integer stages are bit-exact with OpenCV, the fields within 1e-5 relative. */
int spano_equalize_intensities_size(int w, int h, float ratio, int *field_w, int *field_h);
int spano_equalize_intensities(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tile_steps,
                               const uint8_t *const *masks, const size_t *mask_steps, const int *tl_x, const int *tl_y,
                               const int *w, const int *h, float ratio, float *const *fields, const size_t *field_steps);

/* ---- seam search by distance (preview scale) ---------------------------------
 * spano_distance_transform = cv::distanceTransform(mask, dist, cv::DIST_L2, cv::DIST_MASK_5, CV_32F) on CV_8UC1
 * (reference src/math/_distance_cut.cpp:63, src/math/_blending.cpp:110): distance of every pixel to the nearest
 * zero pixel with the 5x5 chamfer metrics (1, 1.4, 2.1969), float32 two-pass arithmetic, FLT_MAX where the mask
 * has no zero pixel at all.  dist_step in BYTES.  Bit-exact with the pinned OpenCV build.
 * spano_dist_cut = dcut::dist_cut(masks, top_lefts) (src/math/_distance_cut.cpp:7-51 with distance_transform
 * :57-73): cut[i] = masks[i], zeroed where an overlapping image j has a strictly larger distance/255.
 * HOST buffers.                                                                                            */
int spano_distance_transform(spano_ctx *ctx, const uint8_t *mask, int w, int h, size_t step, float *dist, size_t dist_step);
int spano_dist_cut(spano_ctx *ctx, int n, const uint8_t *const *masks, const size_t *mask_steps, const int *tl_x,
                   const int *tl_y, const int *w, const int *h, uint8_t *const *cut, const size_t *cut_steps);

/* ---- the other two branches of stitch_parameters::blend (src/classes/_panorama.cpp:220-240) ---------------
 * spano_simple_blend = blnd::simple_blend(images, masks, top_lefts) (src/math/_blending.cpp:83-153): feathering with
 * alpha = normalize(distanceTransform(mask)), images over-composited in order, -> CV_8UC3 canvas.
 * spano_no_blend = blnd::no_blend (src/math/_blending.cpp:157-182): masked copy in order -> CV_8UC3 canvas.
 * tiles 8UC3, masks 8UC1 (validity masks, or mask_cut for no_blend when conf.cut); out = canvas_w x canvas_h from
 * spano_pan_dimension.  HOST buffers.                                                                      */
int spano_simple_blend(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tile_steps, const uint8_t *const *masks,
                       const size_t *mask_steps, const int *tl_x, const int *tl_y, const int *w, const int *h, uint8_t *out,
                       size_t out_step);
int spano_no_blend(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tile_steps, const uint8_t *const *masks,
                   const size_t *mask_steps, const int *tl_x, const int *tl_y, const int *w, const int *h, uint8_t *out,
                   size_t out_step);

/* ---- gain::get_overlapp_intensity (src/math/_gain_compensation.cpp:7-75), the reduction that feeds the gain solve:
 * for every pair i <= j with adj[i*n+j] > 0 or i == j (the reference adds the identity): the number of pixels of the
 * overlap rectangle that are valid in both createSurroundingMask(img, true, 1) masks, and the sums of both 8-bit
 * gray images (cv::cvtColor BGR2GRAY) over those pixels.  Exact integers, returned as doubles like OverlapInfo.
 * `out` must hold n (n + 1) / 2 entries; *n_out receives the number written (the reference's push_back order).
 * HOST buffers.                                                                                             */
typedef struct spano_overlap_info {
    int i, j;
    double area, I_i, I_j;
} spano_overlap_info;
int spano_overlap_intensity(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tile_steps, const int *tl_x,
                            const int *tl_y, const int *w, const int *h, const double *adj, spano_overlap_info *out, int *n_out);

/* a6 alone: `img / gain` on CV_8UC3, in place. */
int spano_apply_gain(spano_ctx *ctx, uint8_t *bgr, int w, int h, size_t step, double gain);

/* ---- a5: sten_proj::disk_reproj, the little-planet centre fix (HOST buffers) -----------
 * (src/math/_projection.cpp:193-294 with get_bounding_box/create_border :87-190 and
 * util::RadialNormalizer src/system/_util.h:172-200).  Inputs: the stereographic tiles and their
 * corners as proj::get_proj_parameters returned them, and the circle (ansatz, radius) that
 * sten_proj::estimate_circle found (that connected-component search stays in the reference).
 * spano_disk_reproj_size gives the new corners (relative to the canvas centre, exactly the values
 * the reference stores back into proj.corners) and sizes; spano_disk_reproj fills the resampled
 * tiles and their recomputed validity masks (createSurroundingMask + 3 erosions).
 * quadratic != 0 selects QUADRATIC_SCALING.  The size query is host arithmetic: ctx may be NULL. */
int spano_disk_reproj_size(spano_ctx *ctx, int n, const int *tl_x, const int *tl_y, const int *w, const int *h,
                           int ansatz_x, int ansatz_y, float radius, int quadratic, int *out_tl_x, int *out_tl_y,
                           int *out_w, int *out_h);
int spano_disk_reproj(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tile_steps, const int *tl_x,
                      const int *tl_y, const int *w, const int *h, int ansatz_x, int ansatz_y, float radius,
                      int quadratic, uint8_t *const *out_tiles, const size_t *out_steps, uint8_t *const *out_masks,
                      const size_t *out_mask_steps);

/* ---- a7 + a8 + a10: blnd::multi_blend (+ the x255/convertTo tail of blend()) ----------
 * tiles[j]  : CV_8UC3 w[j] x h[j]      (gain already applied)
 * masks[j]  : CV_8UC1 mask_cut, 0..255 (seam mask resized to the tile, _panorama.cpp:329-335)
 * masks_orig: CV_8UC1 validity mask {0,255}
 * out       : canvas_w x canvas_h from spano_pan_dimension; float32x3 (SPANO_OUT_F32) or
 *             uint8x3 (SPANO_OUT_U8); out_step in bytes.                                  */
int spano_multiblend(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tile_steps,
                     const uint8_t *const *masks, const size_t *mask_steps, const uint8_t *const *masks_orig,
                     const size_t *orig_steps, const int *tl_x, const int *tl_y, const int *w, const int *h,
                     int bands, double sigma, int out_kind, void *out, size_t out_step);

/* ---- a1: the fused path of stitch_parameters::return_full (MULTI_BLEND) --------------
 * sources -> warp + validity masks + gain -> multi_blend -> 8-bit canvas, tiles stay on
 * the device between the stages.  Geometry (tile corners/sizes, canvas size) must have been
 * queried with spano_warp_roi / spano_pan_dimension; masks_cut[j] is w[j] x h[j].
 * gains may be NULL (conf.gain_compensation == false).
 * row0/row1 select a band of canvas rows [row0,row1) (row-band sharding, one band per GPU);
 * `canvas` then receives only those rows (row 0 of `canvas` is canvas row row0).          */
typedef struct spano_image_desc {
    const uint8_t *src_bgr; /* source image, 8UC3 */
    int src_w, src_h;
    size_t src_step;
    float K[9]; /* K_adj, float32 */
    float R[9];
    double gain;               /* 1.0 = none */
    const uint8_t *mask_cut;   /* w x h, 8UC1 */
    size_t mask_cut_step;
    int tl_x, tl_y, w, h;      /* from spano_warp_roi */
    /* Optional (spano_dev_composite only): the tile's validity mask computed beforehand with
     * spano_dev_tile_mask (w x h, 8UC1, device pointer).  When set, the flood fill is skipped and only
     * the tile rows the band needs are warped -- this is what lets the warp + mask stages shard across
     * GPUs: masks are a whole-tile property, so ranks compute them for disjoint tiles and all-gather them. */
    const uint8_t *valid_mask;
    size_t valid_mask_step;
    /* Size of `mask_cut` when it is still at PREVIEW scale, as stitch_parameters::return_full receives it
     * (src/classes/_panorama.cpp:329-335): the fused path then up-scales it to w x h on the device with
     * cv::resize's 8-bit INTER_LINEAR arithmetic.  0, 0 = mask_cut is already w x h.                    */
    int mask_cut_w, mask_cut_h;
    /* Optional intensity-correction field of test::adjust_intensity (conf.blend_intensity; CV_32FC1,
     * low resolution; src/test/_test.cpp:110-122, src/classes/_panorama.cpp:337-339): applied to the
     * gained tile before the blend.  NULL = blend_intensity off.  intensity_step in BYTES.            */
    const float *intensity;
    int intensity_w, intensity_h;
    size_t intensity_step;
} spano_image_desc;

int spano_composite(spano_ctx *ctx, int proj, float scale, int n, const spano_image_desc *images, int bands,
                    double sigma, int row0, int row1, uint8_t *canvas, size_t canvas_step);

/* Same, with every pointer in `images` and `canvas` a DEVICE pointer; asynchronous on the
 * context's stream.  src_step must be a multiple of 8 and src_bgr 8-byte aligned for the
 * fast sampling path (any layout is accepted; others take the byte path).                */
int spano_dev_composite(spano_ctx *ctx, int proj, float scale, int n, const spano_image_desc *images, int bands,
                        double sigma, int row0, int row1, uint8_t *canvas, size_t canvas_step);

/* The fused path with the little-planet centre fix (conf.proj == STEREOGRAPHIC && conf.fix_center,
 * src/classes/_panorama.cpp:292-311): between the warp and the gain every tile goes through sten_proj::disk_reproj
 * (src/math/_projection.cpp:193-294) for the circle (ansatz, radius) that sten_proj::estimate_circle found -- the
 * radial re-projection, the new corners / sizes and the recomputed validity masks all stay on the device.
 * images[j].tl_x / tl_y / w / h describe the tiles as spano_warp_roi gives them (the warp needs those); the tiles that
 * are blended have the corners and sizes of spano_disk_reproj_size, so the canvas is spano_pan_dimension of THOSE, and a
 * tile-sized mask_cut[j] must have the new size (preview-scale masks are up-scaled to it).  fix == NULL: no centre fix
 * (identical to spano_composite / spano_dev_composite).  spano_composite_fixed: HOST buffers; spano_dev_composite_fixed:
 * DEVICE buffers, asynchronous.                                                                               */
typedef struct spano_center_fix {
    int ansatz_x, ansatz_y; /* circle centre in canvas pixel coordinates (sten_proj::estimate_circle) */
    float radius;
    int quadratic;          /* conf.stretching == QUADRATIC_SCALING */
} spano_center_fix;
int spano_composite_fixed(spano_ctx *ctx, int proj, float scale, int n, const spano_image_desc *images, int bands, double sigma,
                          const spano_center_fix *fix, int row0, int row1, uint8_t *canvas, size_t canvas_step);
int spano_dev_composite_fixed(spano_ctx *ctx, int proj, float scale, int n, const spano_image_desc *images, int bands,
                              double sigma, const spano_center_fix *fix, int row0, int row1, uint8_t *canvas, size_t canvas_step);

/* ---- device-pointer stage entry points (asynchronous on the context's stream) -------- */
/* Validity mask of one warped tile without keeping the tile (a3 sampling + a4), DEVICE pointers. */
int spano_dev_tile_mask(spano_ctx *ctx, int proj, float scale, const float K[9], const float R[9], const uint8_t *src_bgr,
                        int src_w, int src_h, size_t src_step, int tl_x, int tl_y, int w, int h, uint8_t *valid_mask,
                        size_t mask_step);
int spano_dev_warp(spano_ctx *ctx, int proj, float scale, const float K[9], const float R[9], const uint8_t *src_bgr,
                   int src_w, int src_h, size_t src_step, double gain, int tl_x, int tl_y, int dst_w, int dst_h,
                   uint8_t *dst_bgr, size_t dst_step, uint8_t *dst_valid_mask, size_t mask_step);
int spano_dev_multiblend(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tile_steps,
                         const uint8_t *const *masks, const size_t *mask_steps, const uint8_t *const *masks_orig,
                         const size_t *orig_steps, const int *tl_x, const int *tl_y, const int *w, const int *h,
                         int bands, double sigma, int row0, int row1, int out_kind, void *out, size_t out_step);

/* ---- tile-sharded multi-GPU path (one process per GPU, one box) -----------------------------
 * The reference composites serially on one CPU (stitch_parameters::return_full,
 * src/classes/_panorama.cpp:259-354: the loop over images of proj::get_proj_parameters followed by the
 * loop over images inside blnd::multi_blend).  Across G GPUs the two loops shard differently:
 *   - warp + validity mask (src/math/_projection.cpp:422-454) are per-IMAGE work and the mask's flood
 *     fill is a whole-tile property: image j is uploaded, warped and masked ONCE, by its owner rank;
 *   - the blend (src/math/_blending.cpp:186-252) shards by canvas BAND -- row bands, or column bands where that
 *     leaves the bands closer to square (a wide one-row panorama): rank k accumulates and normalises its rows
 *     [row0,row1) (or columns) from the part of every tile that its band reads (its rows / columns plus the blur
 *     radius; BORDER_REFLECT is resolved inside the tile, so nothing else is needed).
 * The exchange between the two is fused into the producing kernels: the warp kernel and the mask kernel
 * store every tile row straight into the memory of the GPU(s) whose band reads it (peer-to-peer stores
 * over NVLink / NVSwitch through pointers opened with spano_peer_open); no staging copy, no collective
 * on the data path -- the caller only needs a barrier between "all owners have written" and "blend".
 *
 * A slice = rows [row0,row1) (and, optionally, columns [col0,col1)) of one warped tile as stored at its band owner:
 * `tile` / `valid` point at the storage of tile pixel (row0, col0) (8UC3 / 8UC1, steps in bytes; 16-byte aligned rows
 * recommended).  col1 <= col0 (e.g. both 0): all columns.  A column range starts on a multiple of 32 and ends on one
 * or at the tile's right edge.  Column ranges serve bands that are COLUMN ranges of the canvas: open the band's blend
 * session with canvas_w = the width of the range and min_x = the panorama's min_x + its first column (the blend
 * clips every tile to the session's columns); such a band reads the 32-column strips of a tile that intersect its
 * columns, plus the blur radius on either side.                                                          */
typedef struct spano_slice {
    int row0, row1;
    uint8_t *tile;
    size_t tile_step;
    uint8_t *valid;
    size_t valid_step;
    int col0, col1;
} spano_slice;

/* Peer memory: cudaMalloc + cudaIpcGetMemHandle on the owner, cudaIpcOpenMemHandle on the peers
 * (handle = the 64 opaque bytes of cudaIpcMemHandle_t, exchanged by the caller, e.g. with
 * torch.distributed.all_gather_object).  spano_peer_close for opened pointers, spano_peer_free for owned. */
int spano_peer_alloc(spano_ctx *ctx, size_t bytes, void **dptr, unsigned char handle[64]);
int spano_peer_open(spano_ctx *ctx, const unsigned char handle[64], void **dptr);
int spano_peer_close(spano_ctx *ctx, void *dptr);
int spano_peer_free(spano_ctx *ctx, void *dptr);

/* Owner side: warp image `im` (a3 + a6), build its validity mask (a4) and scatter tile and mask rows to the
 * `n_slices` band slices (device pointers, local or peer).  Every tile row must be covered by at least one
 * slice only if somebody reads it; rows covered by no slice are simply not stored.  im->mask_cut,
 * im->valid_mask and im->intensity are ignored (the intensity field is not supported on this path).
 * spano_dev_warp_scatter: im->src_bgr is a DEVICE pointer.  spano_warp_scatter: im->src_bgr is a HOST pointer
 * (pinned for a truly asynchronous upload); the image is uploaded on the context's stream first.
 * Both are asynchronous on the context's stream.                                                         */
int spano_dev_warp_scatter(spano_ctx *ctx, int proj, float scale, const spano_image_desc *im, int n_slices,
                           const spano_slice *slices);
int spano_warp_scatter(spano_ctx *ctx, int proj, float scale, const spano_image_desc *im, int n_slices,
                       const spano_slice *slices);

/* Band side: blnd::multi_blend restricted to canvas rows [row0,row1), one image at a time
 * (begin -> add x n -> finish == the loop body of multi_blend + the normalisation + blend()'s x255/convertTo).
 * canvas_w/min_x/min_y are spano_pan_dimension's results for the WHOLE panorama.
 * add: im gives the geometry (tl_x, tl_y, w, h) and mask_cut (+ mask_cut_w/h when at preview scale; it is
 * up-scaled for the slice rows only); `slice` holds the tile rows this band reads, which must include
 * [max(0, first-R), min(h, last+R)) of the tile rows [first,last) that fall into the band (R = ceil(3 sigma));
 * the whole tile when h < 4R.  Tiles that do not touch the band are skipped.
 * spano_dev_*: mask_cut and canvas are DEVICE pointers, asynchronous.  spano_blend_begin / spano_blend_add /
 * spano_blend_finish: mask_cut / canvas are HOST pointers; finish returns when the band is in host memory.
 * spano_blend_begin additionally takes the n images that will be added and uploads their preview-scale
 * mask_cut right away, so that these small copies are queued on the copy engine AHEAD of the owners' large
 * source uploads instead of behind them (spano_blend_add finds them by their mask_cut pointer; an image that
 * was not announced, or a tile-sized mask, is uploaded on demand).  `canvas` (HOST, may be NULL) announces the
 * destination of spano_blend_finish: when it is given and the images are then added in array order, canvas
 * columns that no later image touches are normalised and downloaded while the remaining images are still being
 * blended (spano_blend_finish must then be called with the same canvas).                                    */
int spano_dev_blend_begin(spano_ctx *ctx, int canvas_w, int min_x, int min_y, int row0, int row1, int bands, double sigma);
int spano_blend_begin(spano_ctx *ctx, int canvas_w, int min_x, int min_y, int row0, int row1, int bands, double sigma,
                      int n, const spano_image_desc *images, uint8_t *canvas, size_t canvas_step);
/* Optional, between begin and the first add: everything about the n images that does not depend on the owners' data
 * -- the up-scaling of mask_cut for the slice rows and the blend's sparsity plan -- is done now, on an auxiliary
 * stream, while the band waits for the tile rows to arrive; spano_*_blend_add then only launches the blend kernel
 * of an image it recognises (by address: pass elements of the same `images` array).  slices[j].row1 <= row0 marks an
 * image that does not touch the band.  spano_blend_prepare: mask_cut are HOST pointers (those announced at
 * spano_blend_begin are already on the device).                                                              */
int spano_dev_blend_prepare(spano_ctx *ctx, int n, const spano_image_desc *images, const spano_slice *slices);
int spano_blend_prepare(spano_ctx *ctx, int n, const spano_image_desc *images, const spano_slice *slices);
int spano_dev_blend_add(spano_ctx *ctx, const spano_image_desc *im, const spano_slice *slice);
int spano_dev_blend_finish(spano_ctx *ctx, uint8_t *canvas, size_t canvas_step);
int spano_blend_add(spano_ctx *ctx, const spano_image_desc *im, const spano_slice *slice);
int spano_blend_finish(spano_ctx *ctx, uint8_t *canvas, size_t canvas_step);

/* ---- one step of the tile-sharded path, enqueued by the library -------------------------------------------
 * The owner side and the band side of ONE rank for one pass over the image set (= the loop of
 * proj::get_proj_parameters and the loop of blnd::multi_blend of stitch_parameters::return_full,
 * src/classes/_panorama.cpp:259-354, as sharded above).  There is no collective and no host rendezvous inside a
 * step: ranks order their work through READINESS FLAGS in device memory --
 *   flags[k]        base of rank k's flag block as addressable from this process (own allocation or peer mapping,
 *                   spano_peer_alloc / spano_peer_open): n + world 32-bit counters, zero-initialised once
 *   ready[j]  = flags[k][j]          written by the owner of image j (a store from its stream, after the warp / mask
 *                                    kernels whose peer stores put tile j's rows into rank k's arena): "the rows of
 *                                    image j that band k reads have landed for step `step`"
 *   done[r]   = flags[k][n + r]      written by band r after its last blend of the step: "band r no longer reads its
 *                                    arena for step `step`" (owner k may overwrite it in step + 1)
 * and each stream waits for the counters it depends on with a stream memory operation (cuStreamWaitValue32, >=),
 * which occupies no SM; `step` must increase by one per step, starting at 1.
 * spano_shard_step_owner (context / stream of the owner side): waits done[*] >= step - 1, then for every image j
 * with owner[j] == rank, in `order`: upload (host variant) + warp + validity mask scattered to the slices
 * slices[k * n + j] of every band k that reads it, then ready[j] = step at those bands.
 * spano_shard_step_band (context / stream of the band side; a DIFFERENT context than the owner side's): blend_begin,
 * prepare (mask up-scaling + sparsity plans on the auxiliary stream), then for j = 0..n-1 in array order (the
 * accumulation order of the single-GPU path: bit-identical canvas): wait ready[j] >= step, blend; normalise into
 * `canvas` (device pointer, local or peer; or, host variant, download to host_canvas); done[rank] = step at every owner.
 * images: n descriptors, identical on every rank except the pointers (src_bgr is only dereferenced by the owner,
 * mask_cut only by the bands that read the tile); host != 0: src_bgr / mask_cut are HOST pointers.              */
typedef struct spano_shard_plan {
    int world, rank, n;
    int proj;
    float scale;
    int bands;
    double sigma;
    int canvas_w, min_x, min_y; /* row bands: spano_pan_dimension of the whole panorama.  Column band [c0,c1): canvas_w = c1 - c0,
                                   min_x = the panorama's min_x + c0 (the band's blend session covers just those columns) */
    int row0, row1;             /* this rank's canvas rows (a column band: 0 and the canvas height) */
    const spano_image_desc *images;
    const int *owner;           /* [n] */
    const int *order;           /* [n] permutation: the order in which owners process images */
    const spano_slice *slices;  /* [world * n]; row1 <= row0: image j does not touch band k */
    uint32_t *const *flags;     /* [world] */
    uint8_t *canvas;            /* first pixel of this band in the destination canvas (device pointer; NULL with host != 0) */
    size_t canvas_step;
    /* 1 (or 0): the owners of step s wait until every band has finished step s - 1 (one set of slice arenas).
     * 2: the caller alternates between TWO sets of arenas (`slices` of even / odd steps point into different memory), so the
     * owners of step s only wait for step s - 2 and their warp + mask work overlaps the blends of step s - 1.     */
    int done_lag;
} spano_shard_plan;
int spano_shard_step_owner(spano_ctx *ctx, const spano_shard_plan *plan, unsigned step, int host);
int spano_shard_step_band(spano_ctx *ctx, const spano_shard_plan *plan, unsigned step, int host, uint8_t *host_canvas,
                          size_t host_canvas_step);

/* ---- measurement helpers ---------------------------------------------------------------
 * Time (ms, CUDA events on the context's stream) spent in the kernels of each stage since
 * the last spano_timers_reset: [0] warp  [1] validity mask  [2] blend  [3] normalise.
 * Enabled with spano_timers_enable(ctx, 1); adds event records around each stage.         */
int spano_timers_enable(spano_ctx *ctx, int on);
int spano_timers_reset(spano_ctx *ctx);
int spano_timers_read(spano_ctx *ctx, float ms[4], long long launches[4]);
/* Sparsity of the blend: tile pixels whose whole (2R+1)^2 window of mask_cut is zero have zero weight in every
 * band and add exactly 0 to colour and alpha, so the blend kernel skips them (bit-identical canvas).
 * processed_px = tile pixels the blend kernels actually filtered since the last reset, offered_px = tile pixels
 * of the launches (w x rows in the band).  Synchronises the context's stream.                              */
int spano_blend_stats(spano_ctx *ctx, unsigned long long *processed_px, unsigned long long *offered_px, int reset);
/* FP32 FMA-pipe microbenchmark (the blend kernel's roofline denominator; MEASURED_PEAKS.json
 * carries no fp32 entry): returns achieved TFLOP/s of a register-resident FFMA loop.      */
int spano_fp32_peak(spano_ctx *ctx, int variant, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* SPANO_H */
