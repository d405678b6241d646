// Plane transposes (H <-> W) used to run the horizontal SGA scans as coalesced
// vertical scans: src is `planes` matrices of R rows x Cc columns, dst the same
// planes with Cc rows x R columns.  32x32 tiles through padded shared memory, both
// sides read/written in full rows.  ACC: dst += transpose(src) (fp32 only).
#pragma once
#include "common.cuh"

namespace ganet {

template <typename T, bool ACC>
__global__ void __launch_bounds__(256)
transpose_planes_kernel(const T *__restrict__ src, T *dst, int R, int Cc)
{
    __shared__ T tile[32][33];
    const long long plane = blockIdx.z;
    const T *sp = src + plane * (long long)R * Cc;
    T *dp = dst + plane * (long long)R * Cc;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;    // 32 x 8
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int r = r0 + ty + k, c = c0 + tx;
        if (r < R && c < Cc) tile[ty + k][tx] = sp[(long long)r * Cc + c];
    }
    T old[4];                                                  // ACC: fetched before the barrier so
    if (ACC) {                                                 // both load groups are in flight together
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
            const int c = c0 + ty + k, r = r0 + tx;
            old[k / 8] = (r < R && c < Cc) ? dp[(long long)c * R + r] : T(0);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int c = c0 + ty + k, r = r0 + tx;                // dst row = c, dst col = r
        if (r < R && c < Cc) {
            const long long e = (long long)c * R + r;
            if (ACC) dp[e] = old[k / 8] + tile[tx][ty + k];
            else dp[e] = tile[tx][ty + k];
        }
    }
}

template <typename T, bool ACC>
static int launch_transpose(const T *src, T *dst, long long planes, int R, int Cc, cudaStream_t st)
{
    if (planes <= 0) return GANET_OK;
    const long long zmax = 65535;
    for (long long z0 = 0; z0 < planes; z0 += zmax) {         // gridDim.z limit
        const long long nz = planes - z0 < zmax ? planes - z0 : zmax;
        dim3 grid((unsigned)((Cc + 31) / 32), (unsigned)((R + 31) / 32), (unsigned)nz);
        transpose_planes_kernel<T, ACC><<<grid, 256, 0, st>>>(src + z0 * (long long)R * Cc,
                                                              dst + z0 * (long long)R * Cc, R, Cc);
    }
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

// Byte planes (the direction mask): 32 x 128-byte tiles moved as 32-bit words on both
// sides (the generic kernel's 32-byte rows reach only ~1.6 TB/s).  Needs R % 4 == 0 and
// Cc % 4 == 0; the launcher falls back to the generic kernel otherwise.
__global__ void __launch_bounds__(256)
transpose_u8_planes_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int R, int Cc)
{
    __shared__ uint32_t tile[32][33];              // 32 rows x 128 bytes, padded
    const long long plane = blockIdx.z;
    const uint8_t *sp = src + plane * (long long)R * Cc;
    uint8_t *dp = dst + plane * (long long)R * Cc;
    const int c0 = blockIdx.x * 128, r0 = blockIdx.y * 32;
    const int tid = threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int row = (tid >> 5) + k * 8, word = tid & 31;
        const int r = r0 + row, c = c0 + word * 4;
        uint32_t v = 0;
        if (r < R && c < Cc) v = *reinterpret_cast<const uint32_t *>(sp + (long long)r * Cc + c);
        tile[row][word] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int c = (tid >> 3) + k * 32, q = tid & 7;     // output row c, output word q
        const int oc = c0 + c, orow = r0 + 4 * q;
        if (oc < Cc && orow < R) {
            const int wi = c >> 2, sh = (c & 3) * 8;
            const uint32_t b0 = (tile[4 * q + 0][wi] >> sh) & 0xff, b1 = (tile[4 * q + 1][wi] >> sh) & 0xff;
            const uint32_t b2 = (tile[4 * q + 2][wi] >> sh) & 0xff, b3 = (tile[4 * q + 3][wi] >> sh) & 0xff;
            *reinterpret_cast<uint32_t *>(dp + (long long)oc * R + orow) = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        }
    }
}

static int launch_transpose_u8(const uint8_t *src, uint8_t *dst, long long planes, int R, int Cc,
                               cudaStream_t st)
{
    if ((R % 4) != 0 || (Cc % 4) != 0 || ((uintptr_t)src & 3) || ((uintptr_t)dst & 3))
        return launch_transpose<uint8_t, false>(src, dst, planes, R, Cc, st);
    if (planes <= 0) return GANET_OK;
    const long long zmax = 65535;
    for (long long z0 = 0; z0 < planes; z0 += zmax) {
        const long long nz = planes - z0 < zmax ? planes - z0 : zmax;
        dim3 grid((unsigned)((Cc + 127) / 128), (unsigned)((R + 31) / 32), (unsigned)nz);
        transpose_u8_planes_kernel<<<grid, 256, 0, st>>>(src + z0 * (long long)R * Cc,
                                                         dst + z0 * (long long)R * Cc, R, Cc);
    }
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

// out = max of the four directional aggregates, mask = the first direction attaining it
// (strict < in the order down, up, right, left: the reference's Max chain, GANet_kernel.cu
// :23-36, :975-994), with the two horizontal aggregates still TRANSPOSED (planes of W x H): they
// are read through padded 32x32 shared-memory tiles, so no separate back-transpose pass (and
// no intermediate volumes) is needed.  a0, a1, out, mask: planes of H x W.
__global__ void __launch_bounds__(256)
merge4_transposed_kernel(const float *__restrict__ a0, const float *__restrict__ a1,
                         const float *__restrict__ a2t, const float *__restrict__ a3t,
                         float *__restrict__ out, uint8_t *__restrict__ mask, int H, int W)
{
    __shared__ float t2[32][33], t3[32][33];
    const long long plane = blockIdx.z;
    const long long pb = plane * (long long)H * W;
    const int w0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 32 x 8
#pragma unroll
    for (int k = 0; k < 32; k += 8) {                              // rows of the transposed planes = w
        const int w = w0 + ty + k, h = h0 + tx;
        if (w < W && h < H) {
            t2[ty + k][tx] = __ldg(a2t + pb + (long long)w * H + h);
            t3[ty + k][tx] = __ldg(a3t + pb + (long long)w * H + h);
        }
    }
    float p0[4], p1[4];                                            // standard-layout operands, fetched
#pragma unroll                                                     // before the barrier
    for (int k = 0; k < 32; k += 8) {
        const int h = h0 + ty + k, w = w0 + tx;
        const bool ok = h < H && w < W;
        p0[k / 8] = ok ? __ldg(a0 + pb + (long long)h * W + w) : 0.f;
        p1[k / 8] = ok ? __ldg(a1 + pb + (long long)h * W + w) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int h = h0 + ty + k, w = w0 + tx;
        if (h < H && w < W) {
            const long long e = pb + (long long)h * W + w;
            float b = p0[k / 8];
            uint8_t id = 0;
            const float v1 = p1[k / 8], v2 = t2[tx][ty + k], v3 = t3[tx][ty + k];
            if (b < v1) { b = v1; id = 1; }
            if (b < v2) { b = v2; id = 2; }
            if (b < v3) { b = v3; id = 3; }
            out[e] = b;
            mask[e] = id;
        }
    }
}

// Wide variant for W % 4 == 0: a CTA covers 128 columns x 32 rows, so the 1-byte mask leaves as
// 128-byte rows of 32-bit words (the 32-byte rows of the kernel above are what held it at
// 4.6 TB/s; cf. the byte-plane transpose) and 64 loads per thread are in flight before the barrier.
__global__ void __launch_bounds__(256)
merge4_transposed_wide_kernel(const float *__restrict__ a0, const float *__restrict__ a1,
                              const float *__restrict__ a2t, const float *__restrict__ a3t,
                              float *__restrict__ out, uint8_t *__restrict__ mask, int H, int W)
{
    __shared__ float t2[4][32][33], t3[4][32][33];
    __shared__ __align__(16) uint8_t mk[32][128];
    const long long plane = blockIdx.z;
    const long long pb = plane * (long long)H * W;
    const int w0 = blockIdx.x * 128, h0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 32 x 8
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int k = 0; k < 32; k += 8) {                          // rows of the transposed planes = w
            const int w = w0 + 32 * q + ty + k, h = h0 + tx;
            if (w < W && h < H) {
                t2[q][ty + k][tx] = __ldg(a2t + pb + (long long)w * H + h);
                t3[q][ty + k][tx] = __ldg(a3t + pb + (long long)w * H + h);
            }
        }
    float p0[16], p1[16];                                          // standard-layout operands
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
            const int h = h0 + ty + k, w = w0 + 32 * q + tx;
            const bool ok = h < H && w < W;
            p0[q * 4 + k / 8] = ok ? __ldg(a0 + pb + (long long)h * W + w) : 0.f;
            p1[q * 4 + k / 8] = ok ? __ldg(a1 + pb + (long long)h * W + w) : 0.f;
        }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
            const int h = h0 + ty + k, w = w0 + 32 * q + tx;
            float b = p0[q * 4 + k / 8];
            uint8_t id = 0;
            const float v1 = p1[q * 4 + k / 8], v2 = t2[q][tx][ty + k], v3 = t3[q][tx][ty + k];
            if (b < v1) { b = v1; id = 1; }
            if (b < v2) { b = v2; id = 2; }
            if (b < v3) { b = v3; id = 3; }
            if (h < H && w < W) out[pb + (long long)h * W + w] = b;
            mk[ty + k][32 * q + tx] = id;
        }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {                              // one 32-bit word per lane: 128-byte rows
        const int h = h0 + ty + k, w = w0 + 4 * tx;
        if (h < H && w < W)                                        // W % 4 == 0: a word is inside or outside
            *reinterpret_cast<uint32_t *>(mask + pb + (long long)h * W + w) =
                *reinterpret_cast<const uint32_t *>(&mk[ty + k][4 * tx]);
    }
}

static int launch_merge4_transposed(const float *a0, const float *a1, const float *a2t, const float *a3t,
                                    float *out, uint8_t *mask, long long planes, int H, int W,
                                    cudaStream_t st)
{
    if (planes <= 0) return GANET_OK;
    const bool wide = (W % 4) == 0 && (((uintptr_t)mask) & 3) == 0 && ((long long)H * W) % 4 == 0;
    const long long zmax = 65535;
    for (long long z0 = 0; z0 < planes; z0 += zmax) {
        const long long nz = planes - z0 < zmax ? planes - z0 : zmax;
        const long long o = z0 * (long long)H * W;
        if (wide) {
            dim3 grid((unsigned)((W + 127) / 128), (unsigned)((H + 31) / 32), (unsigned)nz);
            merge4_transposed_wide_kernel<<<grid, 256, 0, st>>>(a0 + o, a1 + o, a2t + o, a3t + o, out + o, mask + o, H, W);
        } else {
            dim3 grid((unsigned)((W + 31) / 32), (unsigned)((H + 31) / 32), (unsigned)nz);
            merge4_transposed_kernel<<<grid, 256, 0, st>>>(a0 + o, a1 + o, a2t + o, a3t + o, out + o, mask + o, H, W);
        }
    }
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

// first arg-max over depth per pixel of a TRANSPOSED aggregate aT[D][W][H], written in
// the standard layout idx[H][W] (MaxDepth, GANet_kernel.cu:50-64)
__global__ void __launch_bounds__(256)
max_depth_from_transposed_kernel(const float *__restrict__ aT, int32_t *__restrict__ idx, int D,
                                 int H, int W)
{
    const long long s = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;      // q = w * H + h
    if (q >= H * W) return;
    const float *ap = aT + s * (long long)D * H * W + q;
    float best = ld_nc(ap);
    int k = 0;
    for (int d = 1; d < D; d++) {
        const float v = ld_nc(ap + (long long)d * H * W);
        if (best < v) { best = v; k = d; }
    }
    const int w = q / H, h = q - w * H;
    idx[s * (long long)H * W + (long long)h * W + w] = k;
}

}  // namespace ganet
