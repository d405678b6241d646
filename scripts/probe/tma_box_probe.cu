// Development probe: which TMA box shapes / start coordinates does the hardware accept?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_box_probe tma_box_probe.cu
//   ./tma_box_probe W H P  bw bh bp  cx cy cz
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../ganet_b200/csrc/tma_utils.cuh"
using namespace ganet;

__global__ void probe(const __grid_constant__ CUtensorMap map, int bytes, int cx, int cy, int cz, float *out, int n)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbarrier_init();
        fence_proxy_async();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, bytes);
        tma_load_3d(smem, &map, &bar, cx, cy, cz);
    }
    mbar_wait(&bar, 0);
    const float *t = reinterpret_cast<const float *>(smem);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = t[i];
}

int main(int argc, char **argv)
{
    if (argc < 10) return 2;
    int W = atoi(argv[1]), H = atoi(argv[2]), P = atoi(argv[3]);
    int bw = atoi(argv[4]), bh = atoi(argv[5]), bp = atoi(argv[6]);
    int cx = atoi(argv[7]), cy = atoi(argv[8]), cz = atoi(argv[9]);
    size_t n = (size_t)W * H * P;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; i++) h[i] = (float)(i % 9973) + 1.f;
    float *d, *o;
    cudaMalloc(&d, n * 4);
    cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    int bn = bw * bh * bp;
    cudaMalloc(&o, bn * 4);
    CUtensorMap map;
    if (!make_plane_map(&map, d, 4, P, H, W, bw, bp, bh)) { printf("encode rejected\n"); return 0; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    probe<<<1, 128, bn * 4 + 128>>>(map, bn * 4, cx, cy, cz, o, bn);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("FAULT: %s\n", cudaGetErrorString(e)); return 0; }
    std::vector<float> r(bn);
    cudaMemcpy(r.data(), o, bn * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int p = 0; p < bp; p++)
        for (int y = 0; y < bh; y++)
            for (int x = 0; x < bw; x++) {
                int gx = cx + x, gy = cy + y, gz = cz + p;
                float want = (gx < 0 || gx >= W || gy < 0 || gy >= H || gz < 0 || gz >= P) ? 0.f
                             : h[((size_t)gz * H + gy) * W + gx];
                if (r[(p * bh + y) * bw + x] != want) bad++;
            }
    printf("ok, %ld mismatches of %d\n", bad, bn);
    return 0;
}
