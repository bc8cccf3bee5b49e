// tma_probe.cu -- minimal cp.async.bulk.tensor.2d load of a (72 words x 32 rows) box from a 2-D uint32 tensor, the exact
// sequence warp_tma_kernel uses; prints whether the staged bytes equal the source.  Finding (B200, driver 580): the box must
// start on a 16-byte boundary of the innermost dimension (coordinate 0 * 4 bytes a multiple of 16): an unaligned start -- e.g. word -3
// -- raises "illegal instruction"; negative and past-the-end coordinates are fine and read as zeros.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int BOX_W = 72, BOX_H = 32, BOX_PITCH = BOX_W * 4;
struct alignas(64) TmaDesc { unsigned long long opaque[16]; };

__global__ void k(const __grid_constant__ TmaDesc tmap, int w0, int row0, uint32_t *out, int variant)
{
    __shared__ __align__(128) uint8_t s_box[BOX_H * BOX_PITCH];
    __shared__ __align__(8) unsigned long long s_bar;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar), dst = (uint32_t)__cvta_generic_to_shared(s_box);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (variant == 1) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(BOX_H * BOX_PITCH) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(w0 + 4 * (int)blockIdx.x), "r"(row0), "r"(bar) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(bar) : "memory");
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < BOX_H * BOX_W; i += blockDim.x) out[i] = reinterpret_cast<uint32_t *>(s_box)[i];
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const int W = argc > 1 ? atoi(argv[1]) : 6000, H = argc > 2 ? atoi(argv[2]) : 4000, variant = argc > 3 ? atoi(argv[3]) : 0;
    const size_t step = (size_t)W * 3;
    std::vector<uint8_t> h(step * H);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 2654435761u >> 13);
    uint8_t *d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    uint32_t *out; cudaMalloc(&out, BOX_H * BOX_PITCH);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry %p q=%d\n", p, (int)q);
    TmaDesc tm;
    const cuuint64_t dims[2] = {(cuuint64_t)((step + 3) / 4), (cuuint64_t)H};
    const cuuint64_t strides[1] = {(cuuint64_t)step};
    const cuuint32_t box[2] = {BOX_W, BOX_H}, estr[2] = {1, 1};
    CUresult r = ((encode_tiled_fn)p)(reinterpret_cast<CUtensorMap *>(&tm), CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode W=%d H=%d step=%zu -> %d\n", W, H, step, (int)r);
    for (int test = 0; test < 4; ++test) {
        const int w0 = test == 0 ? 100 : test == 1 ? -4 : test == 2 ? (int)dims[0] - 20 - ((int)dims[0] & 3) : 1000, row0 = test == 0 ? 50 : test == 1 ? -2 : test == 2 ? H - 10 : 2000;
        cudaMemset(out, 0xff, BOX_H * BOX_PITCH);
        k<<<test == 3 ? 2000 : 1, 256>>>(tm, w0, row0, out, variant);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<uint32_t> o(BOX_H * BOX_W);
        cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (int y = 0; y < BOX_H; ++y)
            for (int x = 0; x < BOX_W; ++x) {
                const long long gx = w0 + x, gy = row0 + y;
                uint32_t exp = 0;
                if (gx >= 0 && gx < (long long)dims[0] && gy >= 0 && gy < H) {
                    for (int b = 0; b < 4; ++b) {
                        const size_t byte = (size_t)gy * step + (size_t)gx * 4 + b;
                        exp |= (uint32_t)(byte < h.size() ? h[byte] : 0) << (8 * b);
                    }
                }
                bad += o[y * BOX_W + x] != exp;
            }
        printf("test %d (w0=%d,row0=%d): %s, mismatching words %zu\n", test, w0, row0, cudaGetErrorString(e), bad);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
