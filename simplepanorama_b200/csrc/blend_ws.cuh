// blend_ws.cuh -- warp-specialised marching-strip multiband blend kernel (sm_100a), radius 21 (sigma = 7).
//
// Same arithmetic, per pixel and in the same order, as blend_march_kernel (blend_march.cuh: what blnd::multi_blend
// computes, reference src/math/_blending.cpp:186-252) -- the two kernels produce bit-identical canvases -- but a
// different schedule.  In the marching kernel all 8 warps walk through three barrier-separated phases per 8-row step
// (stage + combine, vertical pass, horizontal pass); each phase runs at ~63 % of the FMA pipe because 2 warps per
// scheduler cannot hide each other's shared-memory and dependency latencies, and the phases cannot overlap (ncu:
// FMA pipe 64 % active, 8 warps / SM).  Here the CTA has 12 warps with fixed roles:
//
//   H group (4 warps, one per scheduler): producer.  For every 8-row chunk of the strip: global u8 loads (issued one
//     chunk ahead) -> float staging buffer -> horizontal pass for all B sigmas (shared pair sums, packed FFMA2) ->
//     slot of the circular row buffer in shared memory.
//   V group (8 warps, two per scheduler): consumer.  For every 8-row step: vertical pass over the 50 buffered rows of
//     its (sigma, channel, column) items (packed FFMA2, taps from uniform registers) -> results through shared memory ->
//     weights, validity zeroing, band algebra, one float4 read-modify-write of the canvas accumulator.
//
// The circular buffer has 8 chunk slots (7 live + 1 being filled), so the horizontal pass of chunk s+7 runs WHILE the
// vertical pass of step s reads chunks s..s+6: every scheduler always has one H stream and two V streams of FFMA2 to
// pick from and no CTA-wide barrier is left in the loop.  The groups meet only through two mbarriers per slot
// (full: H -> V, empty: V -> H) and synchronise among themselves with named barriers.
// Shared memory (B = 6, SW = 32): 24 planes x 64 rows x 32 floats = 192 KB row buffer + 24 KB results + 9.5 KB staging.
#pragma once

namespace march {

constexpr int WS_SLOTS = 8;                    // chunk slots of the circular row buffer
constexpr int WS_HTHREADS = 128;
// V group size: 8 warps (2 per scheduler) or 12 warps (3 per scheduler; strips of 32 columns only).  Registers per thread
// after the role split: the launch allocates 128 x 384 = 49152 (8 V warps; then 256 x 96 + 128 x 192 = 49152) or
// 128 x 512 = 65536 (12 V warps; then 384 x 96 + 128 x 224 = 65536).
// With 8 V warps the launch is capped at 128 registers per thread (49152 per CTA) instead of the 168 it could take; a
// quarter of the register file (16 K) and 1.4 KB of shared memory stay free, i.e. room for ONE small CTA per SM: in the
// fused path the warp / mask kernels of the NEXT image (auxiliary stream) run next to the blend of the current one.  Every
// kernel of that chain has to fit (<= 64 registers x 256 threads, no shared memory beyond the 1 KB system reserve): one
// that does not waits for the blend to end and holds up the chain behind it (measured, profiles/r2_fused_step_timeline.txt;
// see resize_linear_u8_kernel and plan_kernel).
// (setmaxnreg is a warpgroup-wide instruction: both groups must be whole warpgroups, so VT is a multiple of 128 -- a 10-warp V
// group deadlocks.)  Strips of 16 columns with B = 9 or 10 use the 12-warp V group: six sigma groups of two sigmas instead of
// four groups of three with one of them idle or short.
template <int VT> struct WsRegs {
    static_assert(VT % 128 == 0, "setmaxnreg needs whole warpgroups");
    static constexpr int LAUNCH = 128, V = 96, H = (VT == 256) ? 192 : 224;
};

template <int B, int SW, int VT>
struct WsCfg {
    static constexpr int VTHREADS = VT, THREADS = VT + WS_HTHREADS;
    static constexpr int PLANES = 4 * B;
    static constexpr int PLANE_STRIDE = WS_SLOTS * STEP * SW + (SW == 16 ? 16 : 0);   // floats
    static constexpr int NC = SW + 2 * R;
    static constexpr int RAW_PITCH = (SW == 32) ? 76 : 80;
    static constexpr int GROUPS = SW / 4;
    static constexpr int ROW_ITEMS = STEP * 4 * GROUPS;                 // 256 (SW = 32) or 128 (SW = 16)
    static constexpr int ITEMS_PER_THREAD = ROW_ITEMS / WS_HTHREADS;    // 2 or 1
    static constexpr int NG = VT / (4 * SW);                   // sigma groups of the vertical pass
    static constexpr int HB = (B + NG - 1) / NG;
    static constexpr int NPX = STEP * SW;
    static constexpr int NPASS = (NC + 31) / 32;                        // staging passes of 32 lanes over the NC columns
    static constexpr size_t SMEM = sizeof(float) * ((size_t)PLANES * PLANE_STRIDE + (size_t)PLANES * STEP * SW + 4 * STEP * RAW_PITCH) +
                                   2 * WS_SLOTS * sizeof(unsigned long long);
};

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void group_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// vertical pass of one sigma over a ring of NSLOT chunk slots (same arithmetic and order as vertical_one)
template <int SW, int NSLOT>
__device__ __forceinline__ void vertical_ring(const float *col /* plane + x */, const float2 *tp, int chunk0, float (&res)[STEP])
{
    unsigned long long acc[STEP / 2];
#pragma unroll
    for (int p = 0; p < STEP / 2; ++p) acc[p] = 0ull;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        const int slot = (chunk0 + c) & (NSLOT - 1);
        const float *p = col + slot * STEP * SW;
#pragma unroll
        for (int j = 0; j < STEP; ++j) {
            const int i = c * STEP + j;
            if (i < STEP + 2 * R) {
                const float val = p[j * SW];
#pragma unroll
                for (int q = 0; q < STEP / 2; ++q) {
                    const int d = i - 2 * q;
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

// horizontal pass of one item (row i of channel ch, 4 columns) -> row buffer (same arithmetic as row_pass_item)
template <int B, int SW>
__device__ __forceinline__ void ws_row_item(const Params &P, const float *raw, float *rowbuf, int item, int slot_row0)
{
    using C = WsCfg<B, SW, 256>;
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
            for (int k = 0; k <= R; ++k) {
                // B >= 9: both passes read ONE table (T[R + k] is the .x of the vertical pass's pair R + k), so that the taps
                // of all bands stay inside the 4 KB constant cache (9 x (96 + 352) B would not: measured 2x slower)
                const float t = (B >= 9) ? c_tap2[P.slot + b][R + k].x : c_taps[P.slot + b][k];
                ffma2_vs(a, plo[k], phi[k], t);
            }
            acc2[b][jp] = a;
        }
    }
    float *dst = rowbuf + (size_t)ch * C::PLANE_STRIDE + (slot_row0 + i) * SW + 4 * g;
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const float2 q0 = unpack2(acc2[b][0]), q1 = unpack2(acc2[b][1]);
        *reinterpret_cast<float4 *>(dst + (size_t)(b * 4) * C::PLANE_STRIDE) = make_float4(q0.x, q0.y, q1.x, q1.y);
    }
}

// UNIFORM = true: every thread keeps the 128 registers of the launch (no setmaxnreg) and the H group loads the bytes of a
// chunk right before it stages them instead of one chunk ahead (no prefetch registers; the H group has the slack to sit out
// the load latency).  Measured 2x slower (the H warps, one per scheduler, are the critical path and now wait for their
// loads); kept as SPANO_OPT_BLEND_KERNEL = 4 for A/B runs.
template <int B, int SW, int VT, bool UNIFORM>
__global__ void __maxnreg__(WsRegs<VT>::LAUNCH) blend_ws_kernel(const Params P)
{
    using C = WsCfg<B, SW, VT>;
    constexpr int WS_VTHREADS = VT;
    static_assert(C::NPX <= WS_VTHREADS && C::HB <= 5 && STEP == 8 && (WS_SLOTS & (WS_SLOTS - 1)) == 0 && WS_SLOTS > NCHUNK, "mapping");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *rowbuf = reinterpret_cast<float *>(smem_raw);                  // [PLANES][PLANE_STRIDE]
    float *G = rowbuf + (size_t)C::PLANES * C::PLANE_STRIDE;              // [PLANES][STEP][SW]
    float *raw = G + (size_t)C::PLANES * STEP * SW;                       // [4][STEP][RAW_PITCH]
    unsigned long long *full = reinterpret_cast<unsigned long long *>(raw + 4 * STEP * C::RAW_PITCH);   // [WS_SLOTS]
    unsigned long long *empty = full + WS_SLOTS;                                                        // [WS_SLOTS]

    const int tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < WS_SLOTS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // this CTA's pieces {strip, y0, y1} of the plan; both groups walk the same list
    const int S = (P.w + SW - 1) / SW;
    const PlanView V(S, gridDim.x);
    const int pbeg = __ldg(P.plan + V.cta_start() + blockIdx.x);
    const int pend = __ldg(P.plan + V.cta_start() + blockIdx.x + 1);
    unsigned gc = 0;   // chunks this CTA has started so far (slot = gc % 8, use count of the slot = gc / 8)

    if (tid >= WS_VTHREADS) {
        // =========================== H group: stage + horizontal pass, one chunk at a time ===========================
        // (the horizontal pass holds a 48-float window, 44 pair sums and 2B packed accumulators per item plus the
        // prefetched bytes of the next chunk: it takes the registers the V warps hand back)
        if (!UNIFORM) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WsRegs<VT>::H));
        const int ht = tid - WS_VTHREADS;
        const int lane = ht & 31, hw = ht >> 5;                 // warp hw stages channel hw (all 8 rows of a chunk)
        const uint8_t *sbase = (hw == 0) ? P.cut : P.tile + (hw - 1);
        const size_t sstep = (hw == 0) ? P.cut_step : P.tile_step;
        int xoff[C::NPASS];
        uint32_t pre[STEP * C::NPASS];
        auto prefetch = [&](int yrow0) {
#pragma unroll
            for (int q = 0; q < STEP; ++q) {
                const int y = reflect_idx(yrow0 + q, P.h);
                const uint8_t *rowp = sbase + (size_t)y * sstep;
#pragma unroll
                for (int ps = 0; ps < C::NPASS; ++ps)
                    if (lane + 32 * ps < C::NC) pre[q * C::NPASS + ps] = ldg_u8(rowp + xoff[ps]);
            }
        };
#pragma unroll 1
        for (int pi = pbeg; pi < pend; ++pi) {
            const int *pc = P.plan + V.pieces() + 3 * pi;
            const int tx0 = __ldg(pc) * SW, y0 = __ldg(pc + 1), y1 = __ldg(pc + 2);
            const int nchunks = (y1 - y0 + STEP - 1) / STEP + NCHUNK - 1;   // steps + 6
            const int ybase = y0 - R;
#pragma unroll
            for (int ps = 0; ps < C::NPASS; ++ps) {
                const int c = lane + 32 * ps;
                const int x = reflect_idx(tx0 - R + min(c, C::NC - 1), P.w);
                xoff[ps] = (hw == 0) ? x : 3 * x;
            }
#pragma unroll
            for (int k = 0; k < STEP * C::NPASS; ++k) pre[k] = 0u;
            if (!UNIFORM) prefetch(ybase);
#pragma unroll 1
            for (int c = 0; c < nchunks; ++c, ++gc) {
                const int slot = gc & (WS_SLOTS - 1);
                if (UNIFORM) prefetch(ybase + c * STEP);
                // raw <- the bytes of chunk c (raw is free: the barrier at the end of the previous chunk)
#pragma unroll
                for (int q = 0; q < STEP; ++q) {
                    float *rrow = raw + (hw * STEP + q) * C::RAW_PITCH + lane;
#pragma unroll
                    for (int ps = 0; ps < C::NPASS; ++ps)
                        if (lane + 32 * ps < C::NC) rrow[32 * ps] = (float)pre[q * C::NPASS + ps];
                }
                if (!UNIFORM && c + 1 < nchunks) prefetch(ybase + (c + 1) * STEP);
                mbar_wait(empty + slot, ((gc / WS_SLOTS) & 1) ^ 1);      // the V group is done with the slot's previous chunk
                group_sync(2, WS_HTHREADS);                              // raw visible to the whole group
#pragma unroll
                for (int it = 0; it < C::ITEMS_PER_THREAD; ++it) ws_row_item<B, SW>(P, raw, rowbuf, ht + it * WS_HTHREADS, slot * STEP);
                group_sync(2, WS_HTHREADS);                              // slot complete, raw free
                if (ht == 0) mbar_arrive(full + slot);
            }
        }
        return;
    }

    // =============================== V group: vertical pass + combine, one step at a time ===============================
    if (!UNIFORM) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WsRegs<VT>::V));
    const int vx = tid % SW, vch = (tid / SW) & 3, vg = tid / (4 * SW);
    const int po = tid / SW, px = tid % SW;
#pragma unroll 1
    for (int pi = pbeg; pi < pend; ++pi) {
        const int *pc = P.plan + V.pieces() + 3 * pi;
        const int tx0 = __ldg(pc) * SW, y0 = __ldg(pc + 1), y1 = __ldg(pc + 2);
        const int nsteps = (y1 - y0 + STEP - 1) / STEP;
#pragma unroll 1
        for (int s = 0; s < nsteps; ++s) {
            // global loads of this step's combine (consumed after the vertical pass)
            uint32_t vraw = 0, i0 = 0, i1 = 0, i2 = 0;
            float4 accv = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 *accp = nullptr;
            const int cty = y0 + s * STEP + po, ctx_ = tx0 + px;
            const bool cdo = (tid < C::NPX) && (ctx_ >= P.wx0) && (ctx_ < P.wx1) && (cty < y1);
            if (cdo) {
                vraw = ldg_u8(P.valid + (size_t)cty * P.valid_step + ctx_);
                const uint8_t *pp = P.tile + (size_t)cty * P.tile_step + (size_t)ctx_ * 3;
                i0 = ldg_u8(pp); i1 = ldg_u8(pp + 1); i2 = ldg_u8(pp + 2);
                accp = P.acc + (size_t)(P.ay + cty) * P.canvas_w + (P.ax + ctx_);
                accv = *accp;
            }
            // chunks gc+s .. gc+s+6 must be filtered: the H group fills in order, the last one implies the others
            const unsigned last = gc + s + NCHUNK - 1;
            mbar_wait(full + (last & (WS_SLOTS - 1)), (last / WS_SLOTS) & 1);
            group_sync(1, WS_VTHREADS);                                  // everybody is done with G of the previous step
            const int chunk0 = (gc + s) & (WS_SLOTS - 1);
#pragma unroll 1
            for (int b = 0; b < B; ++b) {
                if (b / C::HB == vg) {
                    float res[STEP];
                    vertical_ring<SW, WS_SLOTS>(rowbuf + (size_t)(b * 4 + vch) * C::PLANE_STRIDE + vx, c_tap2[P.slot + b], chunk0, res);
                    float *gp = G + (size_t)(b * 4 + vch) * STEP * SW + vx;
#pragma unroll
                    for (int o = 0; o < STEP; ++o) gp[o * SW] = res[o];
                }
            }
            group_sync(1, WS_VTHREADS);                                  // G complete; nobody reads chunk gc+s any more
            if (tid == 0) {
                mbar_arrive(empty + chunk0);
                if (s == nsteps - 1)                                     // the piece's remaining chunks were only ever read
                    for (int k = 1; k < NCHUNK; ++k) mbar_arrive(empty + ((chunk0 + k) & (WS_SLOTS - 1)));
            }
            // (the loaded values must not be consumed before this point: without the pin the compiler hoists the first use
            // above the vertical pass and every V warp sits out the DRAM latency at the top of the step)
            asm volatile("" : "+r"(vraw), "+r"(i0), "+r"(i1), "+r"(i2), "+f"(accv.x), "+f"(accv.y), "+f"(accv.z), "+f"(accv.w));
            if (cdo) {
                const bool keep = vraw == 255u;
                const float I0 = (float)i0, I1 = (float)i1, I2 = (float)i2;
                const float *g = G + po * SW + px;
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
                    } else if (b >= 2) {
                        c0 = fmaf(p0 - g0, wprev, c0); c1 = fmaf(p1 - g1, wprev, c1); c2 = fmaf(p2 - g2, wprev, c2);
                    }
                    if (b == B - 1) {
                        c0 = fmaf(I0 - g0, wv, c0); c1 = fmaf(I1 - g1, wv, c1); c2 = fmaf(I2 - g2, wv, c2);
                    }
                    p0 = g0; p1 = g1; p2 = g2; wprev = wv;
                }
                accv.x += c0; accv.y += c1; accv.z += c2; accv.w += wsum;
                *accp = accv;
            }
        }
        gc += nsteps + NCHUNK - 1;
    }
}

} // namespace march
