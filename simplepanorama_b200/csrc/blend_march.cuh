// blend_march.cuh -- marching-strip multiband blend kernel (sm_100a), radius 21 (sigma = 7).
//
// Same arithmetic as the block-tiled kernel in blend_kernels.cu (see the header there for what the
// reference's blnd::multi_blend computes, src/math/_blending.cpp:186-252); different schedule:
//
//   one CTA (256 threads, 1 CTA per SM) owns a strip of SW tile columns and marches down a segment
//   of tile rows 8 rows at a time.  The horizontally filtered rows of ALL 4 channels x B sigmas live
//   in a 56-row (7 chunks of 8) circular buffer in shared memory, so every input row is filtered
//   horizontally exactly once: no halo recomputation in y.  In x only loads are redundant (the
//   21-px halo is staged, not filtered).
//
//   Every thread runs the same schedule per 8-row step s (3 CTA barriers; the loop body is kept small
//   enough -- one copy of each pass -- to stay resident in the instruction cache):
//     A  | stage: raw chunk s+7 (bytes prefetched last step) -> shared as float; prefetch chunk s+8
//        | combine of step s-1: weights, validity zeroing, band algebra, one float4 RMW of the canvas
//        |   (its global loads -- validity, centre pixel, accumulator -- were issued one step earlier)
//     A2 | vertical pass of step s: thread = (sigma group, channel, column); per sigma 50 LDS feed
//        |   8 accumulators each (344 FFMA, taps in registers); results published to shared (G)
//     B  | issue the global loads of the combine of step s
//        | horizontal pass of chunk s+7 for all B sigmas at once (shared pair sums) -> circular buffer
//   Steps -7..-1 run the same loop with the vertical pass and combine skipped (they fill the buffer).
//   Per tile pixel: 4*(21 + 22B) + 4*43B FMA-pipe instructions (B = 6: 1644), 37 B of HBM traffic.
#pragma once

namespace march {

constexpr int R = 21;              // blur radius
constexpr int STEP = 8;            // rows per step == rows per chunk
constexpr int NCHUNK = 7;          // chunks in the circular buffer (50 rows live = 6.25 chunks)
constexpr int NBUF = NCHUNK * STEP; // 56 rows
constexpr int THREADS = 256;

template <int B, int SW>
struct Cfg {
    static constexpr int PLANES = 4 * B;                       // plane = b * 4 + channel
    static constexpr int PLANE_STRIDE = NBUF * SW + (SW == 16 ? 16 : 0); // floats; pad keeps half-warps on distinct banks
    static constexpr int NC = SW + 2 * R;                      // staged columns
    static constexpr int RAW_PITCH = (SW == 32) ? 76 : 80;     // floats, multiple of 4 (LDS.128)
    static constexpr int GROUPS = SW / 4;                      // 4-column groups per row
    static constexpr int ROW_ITEMS = STEP * 4 * GROUPS;        // horizontal-pass items per chunk (<= THREADS)
    static constexpr int STAGE_ELEMS = STEP * NC * 4;          // raw elements per chunk
    static constexpr int PRE = (STAGE_ELEMS + THREADS - 1) / THREADS; // prefetch registers per thread
    static constexpr int NG = THREADS / (4 * SW);              // sigma groups in the vertical pass
    static constexpr int HB = (B + NG - 1) / NG;               // sigmas per group
    static constexpr int NPX = STEP * SW;                      // pixels per step (<= THREADS)
    static constexpr size_t SMEM = sizeof(float) * ((size_t)PLANES * PLANE_STRIDE + (size_t)PLANES * STEP * SW + 4 * STEP * RAW_PITCH) +
                                   sizeof(int) * NC;
};

// The Gaussian taps live in __constant__ memory (the FMAs take them as constant-bank / uniform-register operands;
// taps passed as kernel parameters end up in vector registers and cost the vertical pass its 2-source FFMA2 form).
// They are stored in write-once SLOTS keyed by (bands, sigma) (blend_kernels.cu: tap_slot): a launch names its slot,
// contexts with different parameters use different slots and nothing is ever rewritten while kernels may read it.
struct Params {
    int slot;               // tap slot of this launch's (bands, sigma): first row of its taps in c_taps / c_tap2
    const uint8_t *tile;  size_t tile_step;
    const uint8_t *cut;   size_t cut_step;
    const uint8_t *valid; size_t valid_step;
    int w, h;               // tile extent
    int ty_begin, ty_end;   // tile rows to produce
    float4 *acc;
    int canvas_w;
    int ax, ay;
    const int *plan;        // device plan of this launch (see Plan below)
    int wx0, wx1;           // tile columns that are accumulated (the tile clipped to the accumulator's columns)
};

// ---- sparsity plan -----------------------------------------------------------------------------------------
// A tile pixel whose whole (2R+1)^2 window of mask_cut is zero has weight G_s(mask_cut) = 0 for every band, so
// it adds exactly 0 to colour and alpha: skipping it leaves the canvas bit-identical.  Seam masks partition the
// canvas (dist_cut / graph_cut assign every pixel to one image), so most of a tile is such pixels.  Per launch:
//   activity_kernel : per strip of SW columns, the first / last tile row holding a non-zero mask_cut byte
//   plan_kernel     : per strip the output rows that can be non-zero (neighbour strips and +-R rows included,
//                     clipped to the band), their prefix sum in units of rows, and an even split of that total
//                     over the CTAs: CTA i produces the "virtual rows" [i*per, (i+1)*per) of the concatenation,
//                     i.e. at most a few (strip, row range) pieces -- perfectly balanced, no wave quantisation.
// Plan layout (ints): [0] total rows (each strip rounded up to 8)  [1] rows per CTA  [2] CTAs in use  [3] pieces
//   cta_start[max_cta + 1]   first piece of CTA i (pieces of CTA i: cta_start[i] .. cta_start[i+1])
//   pieces[3 * (S + max_cta)] {strip, y0, y1} tile rows to produce
//   scratch: a0[S], a1[S], prefix[S + 1], ymin[S], ymax[S]
struct PlanView {
    int S, C;
    __host__ __device__ static size_t ints(int S, int max_cta) { return 4 + (size_t)(max_cta + 1) + 3 * (size_t)(S + max_cta) + 5 * (size_t)S + 1; }
    __host__ __device__ PlanView(int strips, int max_cta) : S(strips), C(max_cta) {}
    __host__ __device__ int cta_start() const { return 4; }
    __host__ __device__ int pieces() const { return 4 + C + 1; }
    __host__ __device__ int a0() const { return pieces() + 3 * (S + C); }
    __host__ __device__ int a1() const { return a0() + S; }
    __host__ __device__ int prefix() const { return a1() + S; }
    __host__ __device__ int ymin() const { return prefix() + S + 1; }
    __host__ __device__ int ymax() const { return ymin() + S; }
};

// first / last row with a non-zero mask_cut byte per strip, over tile rows [ra, rb).  One warp row pass covers
// 512 columns (16 per lane); a block of 8 warps covers 64 rows.
template <int SW>
__global__ void activity_kernel(const uint8_t *cut, size_t cut_step, int w, int ra, int rb, int *ymin, int *ymax)
{
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    const int x0 = blockIdx.x * 512 + lane * 16;
    int lo = 0x7fffffff, hi = -1;
    if (x0 < w) {
        const int nx = min(16, w - x0);
        for (int i = 0; i < 8; ++i) {
            const int y = ra + blockIdx.y * 64 + i * 8 + wv;
            if (y >= rb) break;
            const uint8_t *p = cut + (size_t)y * cut_step + x0;
            uint32_t any = 0;
            if (nx == 16 && (((uintptr_t)p) & 15) == 0) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
                any = v.x | v.y | v.z | v.w;
            } else {
                for (int k = 0; k < nx; ++k) any |= __ldg(p + k);
            }
            if (any) { lo = min(lo, y); hi = max(hi, y); }
        }
    }
    if (SW == 32) {   // two lanes per strip
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, 1));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 1));
        if ((lane & 1) == 0 && hi >= 0) { atomicMin(ymin + x0 / 32, lo); atomicMax(ymax + x0 / 32, hi); }
    } else {
        if (hi >= 0) { atomicMin(ymin + x0 / 16, lo); atomicMax(ymax + x0 / 16, hi); }
    }
}

__global__ void plan_init_kernel(int *ymin, int *ymax, int S)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) { ymin[s] = 0x7fffffff; ymax[s] = -1; }
}

// one CTA; S <= 2048 strips, max_cta <= 1024
template <int SW>
// need_rows != 0: ymin / ymax are overwritten with, per strip, the tile rows [need0, need1) at which ANY launch piece reads
// mask_cut in that strip's columns (the strip itself and the neighbours whose windows reach into it, +-R rows, clipped to the
// tile of height need_rows; need1 <= need0: none) -- what an up-scaling of a preview-scale mask has to produce.
__global__ void plan_kernel(int *plan, int S, int max_cta, int ty_begin, int ty_end, int dense, unsigned long long *stats, int wx0, int wx1,
                            int need_rows)
{
    constexpr int NB = (R + SW - 1) / SW;   // neighbour strips whose mask_cut reaches into this strip's window
    const PlanView V(S, max_cta);
    const int *ymin = plan + V.ymin(), *ymax = plan + V.ymax();
    // prefix sums live in the plan itself (scratch `prefix`, and `cta_start` which ends up holding them): no shared memory,
    // so this one-CTA kernel fits on an SM next to a blend CTA of the previous image instead of waiting for it to end
    int *s_pref = plan + V.prefix();
    int *s_cnt = plan + V.cta_start();
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        int lo = 0x7fffffff, hi = -1;
        for (int t = max(0, s - NB); t <= min(S - 1, s + NB); ++t) { lo = min(lo, ymin[t]); hi = max(hi, ymax[t]); }
        int a0 = ty_begin, a1 = ty_end;
        if (!dense) {
            if (hi < 0) { a0 = a1 = ty_begin; }
            else { a0 = max(ty_begin, lo - R); a1 = min(ty_end, hi + R + 1); if (a1 < a0) a1 = a0; }
        }
        if ((s + 1) * SW <= wx0 || s * SW >= wx1) a1 = a0;   // the strip lies outside the columns that are accumulated
        plan[V.a0() + s] = a0;
        plan[V.a1() + s] = a1;
        s_pref[s + 1] = (a1 - a0 + STEP - 1) / STEP * STEP;
    }
    __syncthreads();
    if (need_rows > 0) {
        int *need0 = plan + V.ymin(), *need1 = plan + V.ymax();   // (activity no longer needed: every a0 / a1 is final)
        for (int s = threadIdx.x; s < S; s += blockDim.x) {
            int lo = 0x7fffffff, hi = -1;
            for (int t = max(0, s - NB); t <= min(S - 1, s + NB); ++t) {
                const int b0 = plan[V.a0() + t], b1 = plan[V.a1() + t];
                if (b1 > b0) { lo = min(lo, b0 - R); hi = max(hi, b1 + R); }
            }
            need0[s] = hi < 0 ? 0 : max(0, lo);
            need1[s] = hi < 0 ? 0 : min(need_rows, hi);
        }
    }
    if (threadIdx.x == 0) {
        s_pref[0] = 0;
        int run = 0;
        for (int s = 0; s < S; ++s) { run += s_pref[s + 1]; s_pref[s + 1] = run; }
        const int total = run;
        // CTAs in use: all of them once every CTA gets at least 64 rows; rows per CTA rounded up to the step
        int ncta = min(max_cta, max(1, total / 64));
        int per = ((total + ncta - 1) / ncta + STEP - 1) / STEP * STEP;
        if (per < STEP) per = STEP;
        ncta = total ? (total + per - 1) / per : 0;
        plan[0] = total;  plan[1] = per;  plan[2] = ncta;
        if (stats) {
            atomicAdd(stats, (unsigned long long)total * SW);                       // tile pixels processed (incl. rounding)
            atomicAdd(stats + 1, (unsigned long long)(ty_end - ty_begin) * (wx1 - wx0));      // tile pixels of the launch
        }
    }
    __syncthreads();
    const int total = plan[0], per = plan[1], ncta = plan[2];
    // CTA i produces the virtual rows [i*per, (i+1)*per): pass 0 counts its pieces, pass 1 writes them
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = threadIdx.x; i < ncta; i += blockDim.x) {
            int v = i * per;
            const int vend = min(total, v + per);
            int a = 0, b = S;        // strip with prefix[a] <= v < prefix[a+1]
            while (b - a > 1) {
                const int m = (a + b) >> 1;
                if (s_pref[m] <= v) a = m; else b = m;
            }
            int n = 0;
            int *out = plan + V.pieces() + 3 * (pass ? s_cnt[i] : 0);
            for (int strip = a; v < vend && strip < S; ++strip) {
                const int p0 = s_pref[strip], p1 = s_pref[strip + 1];
                if (p1 <= v) continue;
                const int sa0 = plan[V.a0() + strip], sa1 = plan[V.a1() + strip];
                const int y0 = sa0 + (v - p0), y1 = min(sa1, sa0 + (min(vend, p1) - p0));
                v = min(vend, p1);
                if (y1 <= y0) continue;
                if (pass) { out[3 * n] = strip; out[3 * n + 1] = y0; out[3 * n + 2] = y1; }
                ++n;
            }
            if (!pass) s_cnt[i + 1] = n;
        }
        __syncthreads();
        if (!pass && threadIdx.x == 0) {
            s_cnt[0] = 0;
            int run = 0;
            for (int i = 0; i < ncta; ++i) { run += s_cnt[i + 1]; s_cnt[i + 1] = run; }
            plan[3] = run;
        }
        __syncthreads();
    }
    const int last = s_cnt[ncta];
    for (int i = ncta + 1 + threadIdx.x; i <= max_cta; i += blockDim.x) s_cnt[i] = last;
}

// u8 global load that lands zero-extended in a 32-bit register: no dependent instruction (mask /
// convert) is scheduled behind the load, so its latency can be hidden behind the next phase
__device__ __forceinline__ uint32_t ldg_u8(const uint8_t *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// Blackwell packed FP32: one FFMA2 issues two FMAs.  acc.{lo,hi} += v.{lo,hi} * s   (s broadcast)
__device__ __forceinline__ void ffma2_vs(unsigned long long &acc, float v_lo, float v_hi, float s)
{
    asm("{ .reg .b64 vb, sb; mov.b64 vb, {%1,%2}; mov.b64 sb, {%3,%3}; fma.rn.f32x2 %0, vb, sb, %0; }"
        : "+l"(acc) : "f"(v_lo), "f"(v_hi), "f"(s));
}
__device__ __forceinline__ float2 unpack2(unsigned long long v)
{
    float2 r;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

__device__ __forceinline__ int reflect_idx(int p, int len)
{
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p - 1;
        else p = len - 1 - (p - len);
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

// raw element e of the chunk whose first row is tile row yrow0: (ch, i, c), c fastest
template <int B, int SW>
__device__ __forceinline__ uint32_t fetch_raw(const Params &P, const int *xtab, int e, int yrow0)
{
    using C = Cfg<B, SW>;
    const int c = e % C::NC;
    const int t = e / C::NC;
    const int i = t & (STEP - 1), ch = t >> 3;
    const int y = reflect_idx(yrow0 + i, P.h);
    const int x = xtab[c];
    // one load through a selected address (two predicated loads into one register would serialise
    // on the first one's latency)
    const uint8_t *pc = P.cut + (size_t)y * P.cut_step + x;
    const uint8_t *pt = P.tile + (size_t)y * P.tile_step + (size_t)x * 3 + (ch - 1);
    return ldg_u8(ch == 0 ? pc : pt);
}

// horizontal pass of one item (row i of channel ch, 4 columns) of the staged chunk -> circular buffer slot
template <int B, int SW>
__device__ __forceinline__ void row_pass_item(const Params &P, const float *raw, float *rowbuf, int item, int slot_row0)
{
    using C = Cfg<B, SW>;
    const int g = item % C::GROUPS;
    const int rc = item / C::GROUPS;
    const int i = rc & (STEP - 1), ch = rc >> 3;
    float v[48];
    const float4 *src = reinterpret_cast<const float4 *>(raw + (ch * STEP + i) * C::RAW_PITCH + 4 * g);
#pragma unroll
    for (int q = 0; q < 12; ++q) {
        const float4 s = src[q];
        v[4 * q] = s.x; v[4 * q + 1] = s.y; v[4 * q + 2] = s.z; v[4 * q + 3] = s.w;
    }
    // outputs j, j+1 share one packed FMA per tap: acc2 += {pair_j[k], pair_{j+1}[k]} * tap[b][k]
    unsigned long long acc2[B][2];
#pragma unroll
    for (int jp = 0; jp < 2; ++jp) {
        float plo[R + 1], phi[R + 1];
        plo[0] = v[2 * jp + R];
        phi[0] = v[2 * jp + 1 + R];
#pragma unroll
        for (int k = 1; k <= R; ++k) {
            plo[k] = v[2 * jp + R - k] + v[2 * jp + R + k];
            phi[k] = v[2 * jp + 1 + R - k] + v[2 * jp + 1 + R + k];
        }
#pragma unroll
        for (int b = 0; b < B; ++b) {
            unsigned long long a = 0ull;
#pragma unroll
            for (int k = 0; k <= R; ++k) ffma2_vs(a, plo[k], phi[k], c_taps[P.slot + b][k]);
            acc2[b][jp] = a;
        }
    }
    float out[B][4];
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const float2 q0 = unpack2(acc2[b][0]), q1 = unpack2(acc2[b][1]);
        out[b][0] = q0.x; out[b][1] = q0.y; out[b][2] = q1.x; out[b][3] = q1.y;
    }
    float *dst = rowbuf + (size_t)ch * C::PLANE_STRIDE + (slot_row0 + i) * SW + 4 * g;
#pragma unroll
    for (int b = 0; b < B; ++b)
        *reinterpret_cast<float4 *>(dst + (size_t)(b * 4) * C::PLANE_STRIDE) = make_float4(out[b][0], out[b][1], out[b][2], out[b][3]);
}

// vertical pass of one sigma: 50 rows of one plane column -> 8 outputs.  Output rows (o, o+1) take one
// packed FMA per input row: {res[o], res[o+1]} += {T[d], T[d-1]} * val with d = i - o; the tap pairs come
// from constant memory through uniform registers (tp is warp-uniform), val is the broadcast scalar.
template <int SW>
__device__ __forceinline__ void vertical_one(const float *col /* plane + x */, const float2 *tp, int chunk0, float (&res)[STEP])
{
    unsigned long long acc[STEP / 2];
#pragma unroll
    for (int p = 0; p < STEP / 2; ++p) acc[p] = 0ull;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        int slot = chunk0 + c;
        if (slot >= NCHUNK) slot -= NCHUNK;
        const float *p = col + slot * STEP * SW;
#pragma unroll
        for (int j = 0; j < STEP; ++j) {
            const int i = c * STEP + j;                 // relative row 0..55 (only 0..49 are used)
            if (i < STEP + 2 * R) {
                const float val = p[j * SW];
#pragma unroll
                for (int q = 0; q < STEP / 2; ++q) {
                    const int d = i - 2 * q;            // tap index of output row 2q (2q+1 uses d-1)
                    if (d >= 0 && d <= 2 * R + 1) {
                        const float2 t = tp[d];
                        ffma2_vs(acc[q], t.x, t.y, val);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < STEP / 2; ++q) {
        const float2 r = unpack2(acc[q]);
        res[2 * q] = r.x;
        res[2 * q + 1] = r.y;
    }
}

template <int B, int SW>
__global__ void __launch_bounds__(THREADS, 1) blend_march_kernel(const Params P)
{
    using C = Cfg<B, SW>;
    static_assert(C::ROW_ITEMS <= THREADS && C::NPX <= THREADS && C::HB <= 5 && THREADS == 256 && STEP == 8, "thread mapping");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *rowbuf = reinterpret_cast<float *>(smem_raw);                  // [PLANES][PLANE_STRIDE]
    float *G = rowbuf + (size_t)C::PLANES * C::PLANE_STRIDE;              // [PLANES][STEP][SW]
    float *raw = G + (size_t)C::PLANES * STEP * SW;                       // [4][STEP][RAW_PITCH]
    int *xtab = reinterpret_cast<int *>(raw + 4 * STEP * C::RAW_PITCH);   // [NC] reflected tile x of staged column c

    const int tid = threadIdx.x;
    // This CTA's pieces {strip, y0, y1} of the plan.  The piece loop is folded into the step loop below (one
    // loop level: a second one makes ptxas give up the uniform-register taps of the vertical pass).
    const int S = (P.w + SW - 1) / SW;
    const PlanView V(S, gridDim.x);
    int pi = __ldg(P.plan + V.cta_start() + blockIdx.x);
    const int pend = __ldg(P.plan + V.cta_start() + blockIdx.x + 1);
    int tx0 = 0, y0 = 0, y1 = 0, nsteps = -1, ybase = 0;

    // vertical-pass role
    const int vx = tid % SW, vch = (tid / SW) & 3, vg = tid / (4 * SW);
    // combine role
    const int po = tid / SW, px = tid % SW;

    // Staging role: warp wv owns channel wv>>1 and rows (wv&1)*4 .. +3 of every chunk; its lanes cover the
    // NC staged columns in NPASS passes of 32.  Column offsets are step-invariant and live in registers,
    // the row part of the address is warp-uniform: a staged byte costs ~3 instructions.
    constexpr int NPASS = (C::NC + 31) / 32;
    constexpr int NQ = (4 * STEP) / (THREADS / 32);        // (channel,row) pairs per warp = 4
    const int lane = tid & 31, wv = tid >> 5;
    const int sch = wv >> 1;                               // staged channel of this warp
    const uint8_t *sbase = (sch == 0) ? P.cut : P.tile + (sch - 1);
    const size_t sstep = (sch == 0) ? P.cut_step : P.tile_step;
    int xoff[NPASS];
    (void)xtab;
    auto prefetch = [&](uint32_t (&dst)[NQ * NPASS], int yrow0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int y = reflect_idx(yrow0 + (wv & 1) * NQ + q, P.h);
            const uint8_t *rowp = sbase + (size_t)y * sstep;
#pragma unroll
            for (int ps = 0; ps < NPASS; ++ps)
                if (lane + 32 * ps < C::NC) dst[q * NPASS + ps] = ldg_u8(rowp + xoff[ps]);
        }
    };
    auto stage = [&](const uint32_t (&src)[NQ * NPASS]) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float *rrow = raw + (sch * STEP + (wv & 1) * NQ + q) * C::RAW_PITCH + lane;
#pragma unroll
            for (int ps = 0; ps < NPASS; ++ps)
                if (lane + 32 * ps < C::NC) rrow[32 * ps] = (float)src[q * NPASS + ps];
        }
    };

    uint32_t pre[NQ * NPASS];                              // prefetch registers (raw bytes)
    int chunk0 = 0;                                        // circular slot of chunk s (== slot chunk s+7 will reuse)
    // registers of the combine of the previous step (loaded one step ahead so their latency is hidden)
    uint32_t vraw = 0, i0 = 0, i1 = 0, i2 = 0;
    float4 accv = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 *accp = nullptr;
    bool cdo = false;

    // steps -7..-1 only fill the circular buffer (chunks 0..6); step s >= 0 produces tile rows y0+8s..y0+8s+7;
    // step nsteps only runs the combine of the last rows
    int s = 0;
#pragma unroll 1
    for (;; ++s) {
        if (s > nsteps) {                                  // next piece of this CTA
            if (pi >= pend) break;
            const int *pc = P.plan + V.pieces() + 3 * pi;
            ++pi;
            tx0 = __ldg(pc) * SW;  y0 = __ldg(pc + 1);  y1 = __ldg(pc + 2);
            nsteps = (y1 - y0 + STEP - 1) / STEP;
            ybase = y0 - R;                                // tile row of relative row 0
#pragma unroll
            for (int ps = 0; ps < NPASS; ++ps) {
                const int c = lane + 32 * ps;
                const int x = reflect_idx(tx0 - R + min(c, C::NC - 1), P.w);
                xoff[ps] = (sch == 0) ? x : 3 * x;
            }
#pragma unroll
            for (int k = 0; k < NQ * NPASS; ++k) pre[k] = 0u;
            prefetch(pre, ybase);                          // chunk 0, consumed in step -7
            chunk0 = 0;
            cdo = false;
            s = -NCHUNK;
        }
        __syncthreads();                                   // A: chunks s..s+6 filtered, G(s-1) published, raw free
        const bool more_rows = s + 1 < nsteps;             // chunk s+7 is needed by step s+1
        if (more_rows) {
            stage(pre);
            if (s + 2 < nsteps) prefetch(pre, ybase + (s + 1 + NCHUNK) * STEP);
        }
        // ---- combine of step s-1 (its global loads were issued during step s-1) ----
        if (cdo) {
            const bool keep = vraw == 255u;
            const float I0 = (float)i0, I1 = (float)i1, I2 = (float)i2;
            const float *g = G + po * SW + px;              // + plane * STEP * SW
            float wsum = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
            float p0 = 0.f, p1 = 0.f, p2 = 0.f, wprev = 0.f;
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const float *gb = g + (size_t)(b * 4) * STEP * SW;
                const float wv = keep ? gb[0] * (float)(1.0 / 255.0) : 0.f;
                const float g0 = gb[STEP * SW], g1 = gb[2 * STEP * SW], g2 = gb[3 * STEP * SW];
                wsum += wv;
                if (b == 0) {
                    if (B > 1) { c0 = g0 * wv; c1 = g1 * wv; c2 = g2 * wv; }
                } else if (b >= 2) {                        // band b-1 = G_{b-1} - G_b
                    c0 = fmaf(p0 - g0, wprev, c0); c1 = fmaf(p1 - g1, wprev, c1); c2 = fmaf(p2 - g2, wprev, c2);
                }
                if (b == B - 1) {                           // band B-1 = I - G_{B-1}
                    c0 = fmaf(I0 - g0, wv, c0); c1 = fmaf(I1 - g1, wv, c1); c2 = fmaf(I2 - g2, wv, c2);
                }
                p0 = g0; p1 = g1; p2 = g2; wprev = wv;
            }
            accv.x += c0; accv.y += c1; accv.z += c2; accv.w += wsum;
            *accp = accv;
        }
        __syncthreads();                                   // A2: G(s-1) consumed
        // ---- vertical pass of step s, results straight into G ----
        const bool vdo = s >= 0 && s < nsteps;
        if (vdo) {
            // the loop counter b is uniform, so the taps c_taps[b][..] become uniform-register operands
            // (2 vector-register sources per FFMA); each thread group takes the sigmas of its share
#pragma unroll 1
            for (int b = 0; b < B; ++b) {
                if (b / C::HB == vg) {
                    float res[STEP];
                    vertical_one<SW>(rowbuf + (size_t)(b * 4 + vch) * C::PLANE_STRIDE + vx, c_tap2[P.slot + b], chunk0, res);
                    float *gp = G + (size_t)(b * 4 + vch) * STEP * SW + vx;
#pragma unroll
                    for (int o = 0; o < STEP; ++o) gp[o * SW] = res[o];
                }
            }
        }
        __syncthreads();                                   // B: raw visible; vertical(s) done with the circular buffer
        // issue the global loads of the combine of step s (consumed after the next barrier A)
        {
            const int cty = y0 + s * STEP + po, ctx_ = tx0 + px;
            cdo = vdo && (tid < C::NPX) && (ctx_ >= P.wx0) && (ctx_ < P.wx1) && (cty < y1);
            if (cdo) {
                vraw = ldg_u8(P.valid + (size_t)cty * P.valid_step + ctx_);
                const uint8_t *pp = P.tile + (size_t)cty * P.tile_step + (size_t)ctx_ * 3;
                i0 = ldg_u8(pp); i1 = ldg_u8(pp + 1); i2 = ldg_u8(pp + 2);
                accp = P.acc + (size_t)(P.ay + cty) * P.canvas_w + (P.ax + ctx_);
                accv = *accp;
            }
        }
        if (more_rows && tid < C::ROW_ITEMS) row_pass_item<B, SW>(P, raw, rowbuf, tid, chunk0 * STEP);   // chunk s+7 takes the slot chunk s vacates
        chunk0 = (chunk0 + 1 == NCHUNK) ? 0 : chunk0 + 1;
    }
}

} // namespace march
