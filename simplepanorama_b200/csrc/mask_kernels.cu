// mask_kernels.cu -- validity mask of a warped tile (sm_100a).
//
// Replaces blnd::createSurroundingMask(img, true, 1) + cv::erode(mask, Mat(), (-1,-1), 3)
// (reference src/math/_blending.cpp:278-324, src/math/_projection.cpp:441-443, 287-288):
//   dark  = gray(BGR) <= 1                      (flag produced by the warp kernel, 1 byte/px)
//   out   = 4-connected component of `dark` that touches the tile border (the reference
//           flood-fills from every border pixel with cv::floodFill)
//   mask  = 255 everywhere except `out`; then three 3x3 erosions with OpenCV's default
//           morphology border (+inf)  ==  one (2*3+1)^2 minimum that ignores out-of-image taps.
//
// The flood fill is a whole-tile property, so it is computed as a parallel union-find
// (label equivalence) over the dark pixels with one virtual node for "the border":
//   node 0 = border, node p+1 = pixel p.  Links always point to a smaller id, so node 0 is
//   the root of everything that reaches the border.
// Every thread owns a group of 16 consecutive pixels of one row (one 16-byte load of flags, packed
// to a 16-bit word with a multiply); bright groups -- almost all of a tile -- cost one load and one
// compare.  Runs of dark pixels inside a group start out linked to their run start, horizontal
// links are only needed across group boundaries, and a vertical link is attempted only where a run of
// vertical adjacency begins, which removes most atomics.
// The result is packed to 1 bit/px; the erosion is a vertical OR of the window rows + shifts,
// 32 output pixels per thread written as two 16-byte stores.
// Integer work; results are bit-exact with the reference by construction.
// HBM traffic per tile pixel: 3 x 1 B flag reads, 1 B mask write, 4 B label traffic per DARK pixel.
#include "spano_internal.h"

namespace {

constexpr int GPX = 16; // pixels per thread group

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

// 16 flag bytes (0/1) -> 16-bit word, bit i = pixel i of the group; pixels >= w are cleared
__device__ __forceinline__ uint32_t load_group(const uint8_t *row, int x0, int w)
{
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(row + x0));
    auto nib = [](uint32_t q) { return (((q & 0x01010101u) * 0x01020408u) >> 24) & 0xFu; };
    uint32_t bits = nib(v.x) | (nib(v.y) << 4) | (nib(v.z) << 8) | (nib(v.w) << 12);
    const int valid = w - x0;
    if (valid < GPX) bits &= (1u << valid) - 1u;
    return bits;
}

// thread = (group gx, row y)
__global__ void ccl_init_kernel(const uint8_t *dark, size_t dark_step, int w, int h, int groups, uint32_t *L)
{
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (gx == 0 && y == 0) L[0] = 0;
    if (gx >= groups || y >= h) return;
    const int x0 = gx * GPX;
    uint32_t bits = load_group(dark + (size_t)y * dark_step, x0, w);
    if (!bits) return;
    const uint32_t id0 = (uint32_t)((size_t)y * w + x0) + 1u;
    uint32_t start = 0, prev = 0;
    for (uint32_t m = bits; m;) {
        const uint32_t i = __ffs(m) - 1;
        m &= m - 1;
        if (!(prev && i == prev)) start = i;     // a new run begins unless pixel i-1 was dark
        L[id0 + i] = id0 + start;
        prev = i + 1;
    }
}

// blockDim.x == 32: a warp = 32 consecutive groups (512 pixels) of ONE row.
__global__ void ccl_merge_kernel(const uint8_t *dark, size_t dark_step, int w, int h, int groups, uint32_t *L)
{
    const int lane = threadIdx.x;
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const bool in = gx < groups && y < h;
    const int x0 = gx * GPX;
    const uint8_t *row = dark + (size_t)(in ? y : 0) * dark_step;
    const uint32_t bits = in ? load_group(row, x0, w) : 0u;
    const uint32_t id0 = (uint32_t)((size_t)y * w + x0) + 1u;
    const uint32_t left_px = (in && (bits & 1u) && x0 > 0 && row[x0 - 1]) ? 1u : 0u;   // only matters when pixel 0 is dark

    // A run that enters a group from the left is linked straight to the first pixel of that run as far back as this
    // warp can see it, not to its left neighbour: a horizontal run of G groups then hangs off G/32 links instead of
    // a chain of G.  tail = first pixel of the run that leaves this group through its right edge; a fully dark group
    // that continues its left neighbour's run inherits the neighbour's tail (segmented copy scan over the lanes).
    const bool cont = left_px != 0u;
    uint32_t tail = 0u;
    if (bits & 0x8000u) tail = id0 + (16u - (uint32_t)__clz(~(bits << 16)));   // start of the run that holds pixel 15
    bool open = (bits == 0xFFFFu) && cont && lane > 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t2 = __shfl_up_sync(0xffffffffu, tail, d);
        const int o2 = __shfl_up_sync(0xffffffffu, (int)open, d);
        if (lane >= d && open) { tail = t2; open = o2 != 0; }
    }
    const uint32_t prev_tail = __shfl_up_sync(0xffffffffu, tail, 1);
    if (!bits) return;
    const uint32_t left_target = (lane > 0) ? prev_tail : id0 - 1u;

    uint32_t up_ext = 0;
    if (y > 0) {
        const uint8_t *up = row - dark_step;
        up_ext = (load_group(up, x0, w) << 1) | ((x0 > 0 && up[x0 - 1]) ? 1u : 0u);
    }
    const bool border_row = (y == 0 || y == h - 1);
    const uint32_t up = up_ext >> 1;                       // bit i = the pixel above pixel i is dark
    // one iteration per RUN of dark pixels (ccl_init linked every pixel of a run to the run's first pixel)
    for (uint32_t m = bits; m;) {
        const uint32_t i = __ffs(m) - 1;                   // first pixel of the run
        const uint32_t len = __ffs(~(m >> i)) - 1;         // m < 2^16, so a zero bit always follows
        const uint32_t runmask = ((1u << len) - 1u) << i;
        m &= ~runmask;
        const uint32_t id = id0 + i;
        const int xa = x0 + (int)i, xb = xa + (int)len - 1;
        if (border_row || xa == 0 || xb == w - 1) unite(L, id, 0u);
        if (i == 0 && cont) unite(L, id, left_target);     // runs are pre-linked only inside a group
        // every maximal segment of dark pixels above the run is (part of) one run of the row above; a segment that
        // starts at the group's first pixel with a dark pixel above-left belongs to the upper run the left
        // neighbour group already sees -- and this run continues to the left, so that link is made there
        for (uint32_t um = up & runmask; um;) {
            const uint32_t j = __ffs(um) - 1;
            const uint32_t ulen = __ffs(~(um >> j)) - 1;
            um &= ~(((1u << ulen) - 1u) << j);
            const bool chained = (j == 0) && cont && (up_ext & 1u);
            if (!chained) unite(L, id, id0 + j - (uint32_t)w);
        }
    }
}

// "outside" bit of every pixel (dark and connected to the border): one 16-bit half word per group
__global__ void resolve_bits_kernel(const uint8_t *dark, size_t dark_step, int w, int h, int groups, uint32_t *L,
                                    uint16_t *bits16, int halfwords_per_row)
{
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (gx >= halfwords_per_row || y >= h) return;
    uint32_t out = 0;
    if (gx < groups) {
        const int x0 = gx * GPX;
        const uint32_t bits = load_group(dark + (size_t)y * dark_step, x0, w);
        const uint32_t id0 = (uint32_t)((size_t)y * w + x0) + 1u;
        for (uint32_t m = bits; m;) {                      // one root lookup per run: its pixels share the run's first pixel
            const uint32_t i = __ffs(m) - 1;
            const uint32_t len = __ffs(~(m >> i)) - 1;
            const uint32_t runmask = ((1u << len) - 1u) << i;
            m &= ~runmask;
            if (find_root(L, id0 + i) == 0u) out |= runmask;
        }
    }
    bits16[(size_t)y * halfwords_per_row + gx] = (uint16_t)out;
}

// mask = 0 where any outside pixel lies in the (2r+1)^2 window (out-of-image taps ignored), else 255:
// r erosions with a 3x3 box == one (2r+1)^2 minimum.  One thread = one 32-pixel word of one row.
__device__ __forceinline__ void store_mask_word(uint8_t *dst, uint32_t o, int x0, int w)
{
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

__global__ void erode_bits_kernel(const uint32_t *bits, int words_per_row, int w, int h, int r, uint8_t *mask, size_t mask_step,
                                  const SpanoScatter sc)
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
    if (sc.n > 0) {
        // tile-sharded multi-GPU path: row y goes to every band slice that reads it (possibly peer-GPU memory)
        for (int d = 0; d < sc.n; ++d)
            if (y >= sc.row0[d] && y < sc.row1[d] && x0 >= sc.col0[d] && x0 < sc.col1[d])
                store_mask_word(sc.base[d] + (size_t)y * sc.step[d] + x0, o, x0, w);
    } else {
        store_mask_word(mask + (size_t)y * mask_step + x0, o, x0, w);
    }
}

} // namespace

// `dark` rows must be 16-byte aligned (dark_step % 16 == 0, base from cudaMalloc): every caller in this
// library allocates it that way.
int launch_valid_mask(spano_ctx *ctx, const uint8_t *dark, int w, int h, size_t dark_step, int erode_iters,
                      uint8_t *mask, size_t mask_step, const SpanoScatter *scatter)
{
    if (w <= 0 || h <= 0) return 0;
    if (erode_iters < 0 || erode_iters > 15)
        return spano_fail(ctx, SPANO_E_INVALID, "erode iterations %d not in [0,15]", erode_iters);
    if ((dark_step & 15) || (((uintptr_t)dark) & 15) || dark_step < (size_t)((w + GPX - 1) / GPX * GPX))
        return spano_fail(ctx, SPANO_E_INVALID, "dark-flag rows must be 16-byte aligned and padded to 16 pixels");
    uint32_t *L = nullptr, *bits = nullptr;
    const int groups = (w + GPX - 1) / GPX;
    const int wpr = (w + 31) / 32;
    int rc = spano_reserve(ctx, spano_ctx::BUF_LABELS, ((size_t)w * h + 1) * sizeof(uint32_t), (void **)&L);
    if (rc) return rc;
    rc = spano_reserve(ctx, spano_ctx::BUF_MASK0, (size_t)wpr * h * sizeof(uint32_t), (void **)&bits);
    if (rc) return rc;
    dim3 block(32, 8), grid((2 * wpr + 31) / 32, (h + 7) / 8);
    ccl_init_kernel<<<grid, block, 0, ctx->stream>>>(dark, dark_step, w, h, groups, L);
    ccl_merge_kernel<<<grid, block, 0, ctx->stream>>>(dark, dark_step, w, h, groups, L);
    resolve_bits_kernel<<<grid, block, 0, ctx->stream>>>(dark, dark_step, w, h, groups, L, reinterpret_cast<uint16_t *>(bits), 2 * wpr);
    dim3 eblock(64), egrid((wpr + 63) / 64, h);
    erode_bits_kernel<<<egrid, eblock, 0, ctx->stream>>>(bits, wpr, w, h, erode_iters, mask, mask_step, scatter ? *scatter : SpanoScatter());
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 4;
    return 4;
}

// ---- readiness flags of the tile-sharded multi-GPU path ---------------------------------------------------------------
namespace {
struct FlagTargets { uint32_t *p[SPANO_MAX_FLAG_TARGETS]; };

// Runs after the kernels whose (peer) stores it publishes, on the same stream: the kernel boundary orders those stores
// before this one's; the system-scope fence keeps that order on the way to another GPU's memory.
__global__ void flag_signal_kernel(FlagTargets T, int n, uint32_t value)
{
    if ((int)threadIdx.x < n) {
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t *>(T.p[threadIdx.x]) = value;
    }
}

__global__ void flag_wait_kernel(const uint32_t *flag, uint32_t value)
{
    const volatile uint32_t *p = flag;
    while ((int32_t)(*p - value) < 0) __nanosleep(256);
    __threadfence_system();
}
} // namespace

int launch_flag_signal(spano_ctx *ctx, uint32_t *const *targets, int n, uint32_t value)
{
    int launches = 0;
    for (int i = 0; i < n; i += SPANO_MAX_FLAG_TARGETS) {
        FlagTargets T;
        const int m = n - i < SPANO_MAX_FLAG_TARGETS ? n - i : SPANO_MAX_FLAG_TARGETS;
        for (int k = 0; k < m; ++k) T.p[k] = targets[i + k];
        flag_signal_kernel<<<1, 32, 0, ctx->stream>>>(T, m, value);
        ++launches;
    }
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += launches;
    return launches;
}

int launch_flag_wait_kernel(spano_ctx *ctx, const uint32_t *flag, uint32_t value)
{
    flag_wait_kernel<<<1, 1, 0, ctx->stream>>>(flag, value);
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return 1;
}
