// Development probe: what does the memory system deliver for a given TMA box shape?
// A CTA owns a band of `bh` rows (mode h: walks the W axis in steps of bw columns, the access
// pattern of a horizontal SGA scan in the standard layout) or a strip of bw columns (mode v:
// walks the H axis one row at a time, the pattern of the vertical scans).  Each step loads the
// box (bw, bh, D) of one slice into a shared-memory stage and stores it to the same place of a
// second volume: pure data movement, one elected thread, S-deep ring.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_copy_probe tma_copy_probe.cu
//   ./tma_copy_probe slices D H W   then one line per case on stdout
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../ganet_b200/csrc/tma_utils.cuh"
using namespace ganet;

static bool make_map(CUtensorMap *map, const void *base, long long planes, int H, int W, int bw, int bh,
                     int bp, CUtensorMapSwizzle sw, CUtensorMapL2promotion l2)
{
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bp};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// mode 0: band of bh rows, walk columns; mode 1: strip of bw columns, walk rows
__global__ void __launch_bounds__(32)
copy_kernel(const __grid_constant__ CUtensorMap src, const __grid_constant__ CUtensorMap dst, int mode,
            int bw, int bh, int D, int H, int W, int units, int S, int stage_bytes, int do_store)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * stage_bytes);
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; i++) mbar_init(&full[i], 1);
        fence_mbarrier_init();
        fence_proxy_async();
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    const long long s = blockIdx.x / units;
    const int u = blockIdx.x - (int)(s * units);
    const int steps = mode == 0 ? (W + bw - 1) / bw : (H + bh - 1) / bh;
    const int c2 = (int)(s * D);
    auto coords = [&](int t, int &c0, int &c1) {
        if (mode == 0) { c0 = t * bw; c1 = u * bh; } else { c0 = u * bw; c1 = t * bh; }
    };
    auto issue = [&](int t) {
        const int st = t % S;
        int c0, c1;
        coords(t, c0, c1);
        mbar_arrive_expect_tx(&full[st], (unsigned)stage_bytes);
        tma_load_3d(smem + (size_t)st * stage_bytes, &src, &full[st], c0, c1, c2);
    };
    for (int t = 0; t < S && t < steps; t++) issue(t);
    for (int t = 0; t < steps; t++) {
        const int st = t % S;
        mbar_wait(&full[st], (t / S) & 1);
        int c0, c1;
        coords(t, c0, c1);
        if (do_store) {
            tma_store_3d(&dst, smem + (size_t)st * stage_bytes, c0, c1, c2);
            tma_commit();
        }
        if (t + S < steps) {
            if (do_store) tma_wait_read_all();
            issue(t + S);
        }
    }
    tma_wait_all();
}

int main(int argc, char **argv)
{
    const int slices = argc > 1 ? atoi(argv[1]) : 32;
    const int D = argc > 2 ? atoi(argv[2]) : 192;
    const int H = argc > 3 ? atoi(argv[3]) : 240;
    const int W = argc > 4 ? atoi(argv[4]) : 624;
    const size_t n = (size_t)slices * D * H * W;
    float *a, *b;
    if (cudaMalloc(&a, n * 4) != cudaSuccess || cudaMalloc(&b, n * 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(a, 0, n * 4);
    cudaMemset(b, 0, n * 4);
    cudaFuncSetAttribute(copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    struct Case { int mode, bw, bh, S, sw, store; const char *note; };
    const Case cases[] = {
        {1, 32, 1, 4, 0, 1, "vertical strip, 128B rows (current design)"},
        {1, 32, 1, 8, 0, 1, "vertical strip, 128B rows, 8 stages"},
        {1, 32, 1, 2, 0, 1, "vertical strip, 128B rows, 2 stages (3 CTAs/SM)"},
        {1, 16, 1, 8, 0, 1, "vertical strip, 64B rows"},
        {1, 64, 1, 4, 0, 1, "vertical strip, 256B rows, 4 stages"},
        {1, 128, 1, 2, 0, 1, "vertical strip, 512B rows, 2 stages"},
        {1, 128, 1, 1, 0, 1, "vertical strip, 512B rows, 1 stage (2 CTAs/SM)"},
        {1, 32, 2, 4, 0, 1, "vertical strip, 128B rows, 2 rows per box, 4 stages"},
        {1, 32, 4, 2, 0, 1, "vertical strip, 128B rows, 4 rows per box, 2 stages"},
        {1, 16, 1, 16, 0, 1, "vertical strip, 64B rows, 16 stages"},
        {0, 4, 32, 2, 0, 1, "band 32 rows, 16B inner"},
        {0, 8, 16, 2, 0, 1, "band 16 rows, 32B inner"},
        {0, 8, 16, 2, 1, 1, "band 16 rows, 32B inner, swizzle32"},
        {0, 8, 8, 4, 0, 1, "band 8 rows, 32B inner, 4 stages"},
        {0, 8, 4, 8, 0, 1, "band 4 rows, 32B inner, 8 stages"},
        {0, 16, 8, 2, 0, 1, "band 8 rows, 64B inner"},
        {0, 16, 4, 4, 0, 1, "band 4 rows, 64B inner, 4 stages"},
        {0, 16, 4, 4, 2, 1, "band 4 rows, 64B inner, 4 stages, swizzle64"},
        {0, 16, 2, 8, 0, 1, "band 2 rows, 64B inner, 8 stages"},
        {0, 32, 4, 2, 0, 1, "band 4 rows, 128B inner"},
        {0, 32, 2, 4, 0, 1, "band 2 rows, 128B inner, 4 stages"},
        {0, 32, 2, 4, 3, 1, "band 2 rows, 128B inner, 4 stages, swizzle128"},
        {0, 32, 1, 8, 0, 1, "band 1 row, 128B inner, 8 stages"},
        {0, 32, 1, 4, 0, 1, "band 1 row, 128B inner, 4 stages (2 CTAs/SM)"},
        {0, 32, 1, 2, 0, 1, "band 1 row, 128B inner, 2 stages (4 CTAs/SM)"},
        {0, 8, 16, 2, 0, 0, "band 16 rows, 32B inner, loads only"},
        {0, 32, 2, 4, 0, 0, "band 2 rows, 128B inner, loads only"},
        {1, 32, 1, 4, 0, 0, "vertical strip, loads only"},
    };
    const CUtensorMapSwizzle sws[4] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B};
    printf("# volume %d slices x D=%d x %d x %d, %.2f GB per tensor\n", slices, D, H, W, n * 4 / 1e9);
    for (const Case &c : cases) {
        CUtensorMap ms, md;
        if (!make_map(&ms, a, (long long)slices * D, H, W, c.bw, c.bh, D, sws[c.sw], CU_TENSOR_MAP_L2_PROMOTION_L2_128B) ||
            !make_map(&md, b, (long long)slices * D, H, W, c.bw, c.bh, D, sws[c.sw], CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) {
            printf("%-60s encode rejected\n", c.note);
            continue;
        }
        const int stage_bytes = c.bw * c.bh * D * 4;
        const size_t smem = (size_t)c.S * stage_bytes + c.S * 8 + 64;
        if (smem > 227 * 1024) { printf("%-60s smem %zu too large\n", c.note, smem); continue; }
        const int units = c.mode == 0 ? (H + c.bh - 1) / c.bh : (W + c.bw - 1) / c.bw;
        const long long blocks = (long long)slices * units;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            copy_kernel<<<(unsigned)blocks, 32, smem>>>(ms, md, c.mode, c.bw, c.bh, D, H, W, units, c.S, stage_bytes, c.store);
            cudaEventRecord(e1);
            cudaError_t e = cudaEventSynchronize(e1);
            if (e != cudaSuccess) { printf("%-60s FAULT %s\n", c.note, cudaGetErrorString(e)); return 1; }
            float ms_;
            cudaEventElapsedTime(&ms_, e0, e1);
            if (ms_ < best) best = ms_;
        }
        const double bytes = (double)n * 4 * (c.store ? 2 : 1);
        printf("%-60s smem %6zu  %8.3f ms  %7.1f GB/s\n", c.note, smem, best, bytes / best / 1e6);
        fflush(stdout);
    }
    return 0;
}
