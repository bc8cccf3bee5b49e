// mask_kernels.cu -- validity mask of a warped tile (sm_100a).
//
// Replaces blnd::createSurroundingMask(img, true, 1) + cv::erode(mask, Mat(), (-1,-1), 3)
// (reference src/math/_blending.cpp:278-324, src/math/_projection.cpp:441-443, 287-288):
//   dark  = gray(BGR) <= 1                      (flag produced by the warp kernel)
//   out   = 4-connected component of `dark` that touches the tile border (the reference
//           flood-fills from every border pixel with cv::floodFill)
//   mask  = 255 everywhere except `out`; then three 3x3 erosions with OpenCV's default
//           morphology border (+inf)  ==  one (2*3+1)^2 minimum that ignores out-of-image taps.
//           Done on a 1-bit-per-pixel image of `out` (ballot-packed): vertical OR of the window rows,
//           horizontal dilation by shifts, 32 output pixels per thread as two 16-byte stores.
//
// The flood fill is a whole-tile property, so it is computed as a parallel union-find
// (label equivalence) over the dark pixels with one virtual node for "the border":
//   node 0 = border, node p+1 = pixel p.  Links always point to a smaller id, so node 0 is
//   the root of everything that reaches the border.  Runs of dark pixels inside a 32-px warp
//   segment start out already linked to their run start (one ballot), and a vertical link is
//   attempted only where a run of vertical adjacency begins, which removes most atomics.
// Integer work; results are bit-exact with the reference by construction.
// HBM traffic: 1 B (dark) + 4 B labels written, ~2x4 B labels re-read, 1 B mask written per px.
#include "spano_internal.h"

namespace {


__device__ __forceinline__ uint32_t ld_label(const uint32_t *L, uint32_t i) { return __ldcg(L + i); }

// root of node a; compresses the path behind it (only ever lowers a link: safe under races)
__device__ __forceinline__ uint32_t find_root(uint32_t *L, uint32_t a)
{
    uint32_t p = ld_label(L, a);
    while (p != a) {
        const uint32_t g = ld_label(L, p);
        if (g != p) atomicMin(L + a, g);
        a = p;
        p = g;
    }
    return a;
}

__device__ __forceinline__ void unite(uint32_t *L, uint32_t a, uint32_t b)
{
    for (;;) {
        a = find_root(L, a);
        b = find_root(L, b);
        if (a == b) return;
        if (a < b) { const uint32_t t = a; a = b; b = t; }
        const uint32_t old = atomicMin(L + a, b); // link the larger root under the smaller
        if (old == a) return;
        a = old; // somebody re-linked a meanwhile: carry on with its new parent
    }
}

// one thread per pixel, blockDim.x == 32 so that a warp is one 32-px row segment
__global__ void ccl_init_kernel(const uint8_t *dark, size_t dark_step, int w, int h, uint32_t *L)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) L[0] = 0;
    const bool in = (x < w) && (y < h);
    const bool d = in && dark[(size_t)y * dark_step + x] != 0;
    const uint32_t bits = __ballot_sync(0xFFFFFFFFu, d);
    if (!in) return;
    const size_t p = (size_t)y * w + x;
    if (!d) return;   // labels of non-dark pixels are never read
    const uint32_t lane = threadIdx.x;
    const uint32_t zeros_below = ~bits & ((1u << lane) - 1u);
    const uint32_t start = zeros_below ? (32u - __clz(zeros_below)) : 0u; // first lane of my run
    L[p + 1] = (uint32_t)(p + 1) - (lane - start);
}

__global__ void ccl_merge_kernel(const uint8_t *dark, size_t dark_step, int w, int h, uint32_t *L)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t *row = dark + (size_t)y * dark_step;
    if (!row[x]) return;
    const uint32_t id = (uint32_t)((size_t)y * w + x) + 1u;
    if (x == 0 || y == 0 || x == w - 1 || y == h - 1) unite(L, id, 0u);
    const bool left = x > 0 && row[x - 1];
    if (left && (x & 31) == 0) unite(L, id, id - 1u); // runs are pre-linked only inside a segment
    if (y > 0) {
        const uint8_t *up = row - dark_step;
        if (up[x]) {
            const bool chained = left && (x & 31) != 0 && up[x - 1];
            if (!chained) unite(L, id, id - (uint32_t)w);
        }
    }
}

// "outside" bit of every pixel (dark and connected to the border), 32 pixels per word
__global__ void resolve_bits_kernel(const uint8_t *dark, size_t dark_step, int w, int h, uint32_t *L, uint32_t *bits,
                                    int words_per_row)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    bool out = false;
    if (x < w && y < h && dark[(size_t)y * dark_step + x]) {
        const uint32_t id = (uint32_t)((size_t)y * w + x) + 1u;
        out = find_root(L, id) == 0u;
    }
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, out);
    if (threadIdx.x == 0 && y < h && blockIdx.x < words_per_row) bits[(size_t)y * words_per_row + blockIdx.x] = word;
}

// mask = 0 where any outside pixel lies in the (2r+1)^2 window (out-of-image taps ignored), else 255:
// r erosions with a 3x3 box == one (2r+1)^2 minimum.  One thread = one 32-pixel word of one row.
__global__ void erode_bits_kernel(const uint32_t *bits, int words_per_row, int w, int h, int r, uint8_t *mask, size_t mask_step)
{
    const int wx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (wx >= words_per_row || y >= h) return;
    uint32_t lo = 0, mid = 0, hi = 0;   // vertical OR over the window rows of words wx-1, wx, wx+1
    for (int dy = -r; dy <= r; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= h) continue;
        const uint32_t *row = bits + (size_t)yy * words_per_row;
        mid |= row[wx];
        if (wx > 0) lo |= row[wx - 1];
        if (wx + 1 < words_per_row) hi |= row[wx + 1];
    }
    uint32_t o = mid;
    for (int d = 1; d <= r; ++d) {
        o |= (mid << d) | (lo >> (32 - d));   // outside pixel d to the left
        o |= (mid >> d) | (hi << (32 - d));   // outside pixel d to the right
    }
    const int x0 = wx * 32;
    uint8_t *dst = mask + (size_t)y * mask_step + x0;
    if (x0 + 32 <= w && (((uintptr_t)dst) & 15) == 0) {
        uint32_t v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t nib = (o >> (4 * q)) & 15u;
            // 4 pixels -> 4 bytes: 0x00 where the bit is set, 0xFF otherwise
            v[q] = ((nib & 1u) ? 0u : 0x000000FFu) | ((nib & 2u) ? 0u : 0x0000FF00u) | ((nib & 4u) ? 0u : 0x00FF0000u) |
                   ((nib & 8u) ? 0u : 0xFF000000u);
        }
        reinterpret_cast<uint4 *>(dst)[0] = make_uint4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<uint4 *>(dst)[1] = make_uint4(v[4], v[5], v[6], v[7]);
    } else {
        for (int i = 0; i < 32 && x0 + i < w; ++i) dst[i] = ((o >> i) & 1u) ? 0 : 255;
    }
}

} // namespace

int launch_valid_mask(spano_ctx *ctx, const uint8_t *dark, int w, int h, size_t dark_step, int erode_iters,
                      uint8_t *mask, size_t mask_step)
{
    if (w <= 0 || h <= 0) return 0;
    if (erode_iters < 0 || erode_iters > 15)
        return spano_fail(ctx, SPANO_E_INVALID, "erode iterations %d not in [0,15]", erode_iters);
    uint32_t *L = nullptr, *bits = nullptr;
    const int wpr = (w + 31) / 32;
    int rc = spano_reserve(ctx, spano_ctx::BUF_LABELS, ((size_t)w * h + 1) * sizeof(uint32_t), (void **)&L);
    if (rc) return rc;
    rc = spano_reserve(ctx, spano_ctx::BUF_MASK0, (size_t)wpr * h * sizeof(uint32_t), (void **)&bits);
    if (rc) return rc;
    dim3 block(32, 8), grid(wpr, (h + 7) / 8);
    ccl_init_kernel<<<grid, block, 0, ctx->stream>>>(dark, dark_step, w, h, L);
    ccl_merge_kernel<<<grid, block, 0, ctx->stream>>>(dark, dark_step, w, h, L);
    resolve_bits_kernel<<<grid, block, 0, ctx->stream>>>(dark, dark_step, w, h, L, bits, wpr);
    dim3 eblock(64), egrid((wpr + 63) / 64, h);
    erode_bits_kernel<<<egrid, eblock, 0, ctx->stream>>>(bits, wpr, w, h, erode_iters, mask, mask_step);
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 4;
    return 4;
}
