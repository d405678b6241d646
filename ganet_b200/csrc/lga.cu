// LGA (local guided aggregation) for sm_100a.
//
// Reference being replaced: libs/GANet/src/GANet_kernel.cu
//   lga_filtering_forward :1131-1175, lga_filter_backward :1177-1216,
//   lga_data_backward :1218-1269, hosts lga_forward/backward :1271-1322 and the
//   lga3d variants :1324-1364 (same kernels, leading dims folded into B).
//
// y[b,d,h,w] = sum over taps (dd,r,c) in {-1,0,1} x [-R,R]^2 of
//              f[b,loc,h,w] * x[b,d',h',w'],  (d',h',w') = (d+dd,h+r,w+c) when all
//              three are in range, else the CENTRE voxel (d,h,w)   (:1162-1165)
// with loc = (dd+1)(2R+1)^2 + (r+R)(2R+1) + (c+R).
//
// Kernel layout: one thread per pixel (h,w) keeps its 3(2R+1)^2 filter taps in
// registers and walks the depth axis, so filters are read from HBM once per
// pixel instead of once per voxel, the x neighbourhood comes from L1 (adjacent
// threads share it), and every output is accumulated in a register and written
// once (the reference does 75 global read-modify-writes per voxel).
#include <stdlib.h>

#include "common.cuh"
#include "lga_tile.cuh"

namespace ganet {

constexpr int kLgaThreads = 128;

// All three kernels iterate over INPUT depth planes: the (2R+1)^2 neighbourhood of plane dp
// is loaded once (25 loads for R=2) and used for the three depth taps it feeds, so a voxel
// costs 25 loads + 75 FMAs instead of 75 + 75.  Taps whose (r,c) leaves the image are given
// weight 0 (their load is redirected to the pixel itself) and their true contribution --
// always "centre voxel times tap weight" (:1162-1165) -- is added as one term per output.

template <int R>
struct LgaGeom {
    static constexpr int WS = 2 * R + 1, P2 = WS * WS, F = 3 * P2;
};

// in-plane validity + clamped neighbour offsets of this pixel
template <int R>
__device__ __forceinline__ void lga_taps(int h, int w, int H, int W, bool (&ok)[LgaGeom<R>::P2],
                                         int (&noff)[LgaGeom<R>::P2])
{
    constexpr int WS = LgaGeom<R>::WS;
#pragma unroll
    for (int r = -R; r <= R; r++)
#pragma unroll
        for (int c = -R; c <= R; c++) {
            const int t = (r + R) * WS + (c + R);
            ok[t] = (h + r >= 0) && (h + r < H) && (w + c >= 0) && (w + c < W);
            noff[t] = ok[t] ? r * W + c : 0;
        }
}

// the (2R+1)^2 neighbourhood of one plane.  Interior pixels: one pointer per filter row and
// compile-time column offsets (LDG [R + imm]); rebuilding a 64-bit address from a runtime
// offset for each of the 25 loads cost ~4 integer instructions per load.
template <int R>
__device__ __forceinline__ void lga_load_plane(const float *p, int W, bool interior,
                                               const int (&noff)[LgaGeom<R>::P2],
                                               float (&v)[LgaGeom<R>::P2])
{
    constexpr int WS = LgaGeom<R>::WS;
    if (interior) {
#pragma unroll
        for (int r = -R; r <= R; r++) {
            const float *row = p + r * W;
#pragma unroll
            for (int c = -R; c <= R; c++) v[(r + R) * WS + (c + R)] = ld_nc(row + c);
        }
    } else {
#pragma unroll
        for (int t = 0; t < LgaGeom<R>::P2; t++) v[t] = ld_nc(p + noff[t]);
    }
}

// The plane walk has no register room for a software pipeline (75 tap weights live in
// registers), so the rows of a plane a few steps ahead are pulled towards L1 instead.
constexpr int kLgaPrefetch = 3;
template <int R>
__device__ __forceinline__ void lga_prefetch_plane(const float *p, int W, bool interior)
{
    if (interior) {
#pragma unroll
        for (int r = -R; r <= R; r++)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(p + r * W));
    } else {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    }
}

// ---- forward ---------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kLgaThreads)
lga_fwd_kernel(const float *__restrict__ x, const float *__restrict__ f, float *__restrict__ y,
               int D, int H, int W, int d_chunk)
{
    constexpr int P2 = LgaGeom<R>::P2, F = LgaGeom<R>::F;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int h = blockIdx.y;
    const int nchunk = (D + d_chunk - 1) / d_chunk;
    const long long b = blockIdx.z / nchunk;
    const int dc = blockIdx.z % nchunk;
    if (w >= W) return;
    const int HW = H * W;
    const float *xb = x + b * (long long)D * HW + h * W + w;
    float *yb = y + b * (long long)D * HW + h * W + w;
    const float *fb = f + b * (long long)F * HW + h * W + w;

    bool ok[P2];
    int noff[P2];
    lga_taps<R>(h, w, H, W, ok, noff);
    const bool interior = (h >= R) && (h + R < H) && (w >= R) && (w + R < W);
    float wz[F];                 // tap weights, 0 where (r,c) leaves the image
    float cval[3] = {0.f, 0.f, 0.f};   // per depth tap: sum of the in-image weights
    float coob = 0.f;            // sum of all out-of-image weights (all three depth taps)
#pragma unroll
    for (int dd = 0; dd < 3; dd++)
#pragma unroll
        for (int t = 0; t < P2; t++) {
            const float v = ld_nc(fb + (long long)(dd * P2 + t) * HW);
            wz[dd * P2 + t] = ok[t] ? v : 0.f;
            if (ok[t]) cval[dd] += v; else coob += v;
        }

    const int dbeg = dc * d_chunk, dend = min(D, dbeg + d_chunk);
    // rolling outputs: a0 = y[dp-1] (complete after plane dp), a1 = y[dp], a2 = y[dp+1]
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    float cprev = 0.f;           // centre value of plane dp-1
    for (int dp = max(dbeg - 1, 0); dp < min(dend + 1, D); dp++) {
        const float *xp = xb + (long long)dp * HW;
        const float ctr = ld_nc(xp);
        // contributions of this plane to y[dp+1], y[dp], y[dp-1]; one partial sum per filter
        // row so that 3*(2R+1) FMA chains are in flight instead of 3 (latency-bound at 8 warps/SM)
        float q0[2 * R + 1], q1[2 * R + 1], q2[2 * R + 1];
        float v[P2];
        if (dp + kLgaPrefetch < D) lga_prefetch_plane<R>(xp + (long long)kLgaPrefetch * HW, W, interior);
        lga_load_plane<R>(xp, W, interior, noff, v);
#pragma unroll
        for (int r = 0; r < 2 * R + 1; r++) {
            q0[r] = 0.f; q1[r] = 0.f; q2[r] = 0.f;
#pragma unroll
            for (int c = 0; c < 2 * R + 1; c++) {
                const int t = r * (2 * R + 1) + c;
                q0[r] = fmaf(v[t], wz[0 * P2 + t], q0[r]);    // depth tap -1 of output dp+1
                q1[r] = fmaf(v[t], wz[1 * P2 + t], q1[r]);    // depth tap  0 of output dp
                q2[r] = fmaf(v[t], wz[2 * P2 + t], q2[r]);    // depth tap +1 of output dp-1
            }
        }
        float n0 = 0.f, n1 = 0.f, n2 = 0.f;
#pragma unroll
        for (int r = 0; r < 2 * R + 1; r++) { n0 += q0[r]; n1 += q1[r]; n2 += q2[r]; }
        a0 += n2; a1 += n1; a2 += n0;
        // centre-fallback terms of output dp (its own centre): out-of-image taps always,
        // plus the whole -1 / +1 depth tap at the volume faces
        float fb_w = coob;
        if (dp == 0) fb_w += cval[0];
        if (dp == D - 1) fb_w += cval[2];
        a1 = fmaf(ctr, fb_w, a1);
        if (dp - 1 >= dbeg && dp - 1 < dend) yb[(long long)(dp - 1) * HW] = a0;
        a0 = a1; a1 = a2; a2 = 0.f;
        cprev = ctr;
    }
    (void)cprev;
    const int last = min(dend + 1, D) - 1;       // plane index processed last
    if (last >= dbeg && last < dend) yb[(long long)last * HW] = a0;   // only when dend == D
}

// ---- backward: filter gradient (:1177-1216) ---------------------------------
// gf[dd][t] = sum_d go[d] * (x[d+dd] at neighbour t, or the centre x[d] when that voxel is
// outside).  One thread per pixel, all depths; 75 accumulators in registers.
template <int R>
__global__ void __launch_bounds__(kLgaThreads)
lga_bwd_filter_kernel(const float *__restrict__ x, const float *__restrict__ go,
                      float *__restrict__ gf, int accumulate, int D, int H, int W)
{
    constexpr int P2 = LgaGeom<R>::P2, F = LgaGeom<R>::F;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int h = blockIdx.y;
    const long long b = blockIdx.z;
    if (w >= W) return;
    const int HW = H * W;
    const float *xb = x + b * (long long)D * HW + h * W + w;
    const float *gb = go + b * (long long)D * HW + h * W + w;
    float *gfb = gf + b * (long long)F * HW + h * W + w;

    bool ok[P2];
    int noff[P2];
    lga_taps<R>(h, w, H, W, ok, noff);
    const bool interior = (h >= R) && (h + R < H) && (w >= R) && (w + R < W);
    float acc[F];
#pragma unroll
    for (int l = 0; l < F; l++) acc[l] = 0.f;

    // plane dp of x meets go[dp+1] (depth tap -1), go[dp] (tap 0), go[dp-1] (tap +1)
    float gm = 0.f;                               // go[dp-1]
    float gc = ld_nc(gb);                         // go[dp]
    float sgc = 0.f, e_first = 0.f, e_last = 0.f; // sum go*x centre; face terms
    for (int dp = 0; dp < D; dp++) {
        const float gp = (dp + 1 < D) ? ld_nc(gb + (long long)(dp + 1) * HW) : 0.f;   // go[dp+1]
        const float *xp = xb + (long long)dp * HW;
        const float ctr = ld_nc(xp);
        sgc = fmaf(gc, ctr, sgc);
        if (dp == 0) e_first = gc * ctr;
        if (dp == D - 1) e_last = gc * ctr;
        float v[P2];
        if (dp + kLgaPrefetch < D) {
            lga_prefetch_plane<R>(xp + (long long)kLgaPrefetch * HW, W, interior);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(gb + (long long)(dp + kLgaPrefetch) * HW));
        }
        lga_load_plane<R>(xp, W, interior, noff, v);
#pragma unroll
        for (int t = 0; t < P2; t++) {
            acc[0 * P2 + t] = fmaf(gp, v[t], acc[0 * P2 + t]);
            acc[1 * P2 + t] = fmaf(gc, v[t], acc[1 * P2 + t]);
            acc[2 * P2 + t] = fmaf(gm, v[t], acc[2 * P2 + t]);
        }
        gm = gc; gc = gp;
    }
#pragma unroll
    for (int dd = 0; dd < 3; dd++)
#pragma unroll
        for (int t = 0; t < P2; t++) {
            // out-of-image tap: every depth falls back to the centre; in-image tap: only the
            // face depth whose d+dd leaves the volume does
            float v = ok[t] ? acc[dd * P2 + t] + (dd == 0 ? e_first : dd == 2 ? e_last : 0.f) : sgc;
            float *dst = gfb + (long long)(dd * P2 + t) * HW;
            *dst = accumulate ? *dst + v : v;
        }
}

// ---- backward: data gradient (:1218-1269) -----------------------------------
// gx[v] = sum over taps: neighbour u = v + tap in range ? go[u] * f[loc(-tap) at u]
//                                                        : go[v] * f[loc(tap) at v]
// Same rolling structure as the forward with the mirrored weights of the neighbours.
template <int R>
__global__ void __launch_bounds__(kLgaThreads)
lga_bwd_data_kernel(const float *__restrict__ f, const float *__restrict__ go,
                    float *__restrict__ gx, int D, int H, int W, int d_chunk)
{
    constexpr int WS = LgaGeom<R>::WS, P2 = LgaGeom<R>::P2, F = LgaGeom<R>::F;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int h = blockIdx.y;
    const int nchunk = (D + d_chunk - 1) / d_chunk;
    const long long b = blockIdx.z / nchunk;
    const int dc = blockIdx.z % nchunk;
    if (w >= W) return;
    const int HW = H * W;
    const float *gb = go + b * (long long)D * HW + h * W + w;
    float *gxb = gx + b * (long long)D * HW + h * W + w;
    const float *fb = f + b * (long long)F * HW + h * W + w;

    bool ok[P2];
    int noff[P2];
    lga_taps<R>(h, w, H, W, ok, noff);
    const bool interior = (h >= R) && (h + R < H) && (w >= R) && (w + R < W);
    float wn[F];                 // mirrored weight of the neighbour pixel, 0 outside the image
    float cval[3] = {0.f, 0.f, 0.f};   // this pixel's own in-image weights per depth tap
    float coob = 0.f;                  // this pixel's own out-of-image weights
#pragma unroll
    for (int r = -R; r <= R; r++)
#pragma unroll
        for (int c = -R; c <= R; c++) {
            const int t = (r + R) * WS + (c + R);
#pragma unroll
            for (int dd = -1; dd <= 1; dd++) {
                const int loc_m = (-dd + 1) * P2 + (-r + R) * WS + (-c + R);
                wn[(dd + 1) * P2 + t] = ok[t] ? ld_nc(fb + r * W + c + (long long)loc_m * HW) : 0.f;
                const float own = ld_nc(fb + (long long)((dd + 1) * P2 + t) * HW);
                if (ok[t]) cval[dd + 1] += own; else coob += own;
            }
        }

    const int dbeg = dc * d_chunk, dend = min(D, dbeg + d_chunk);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int dp = max(dbeg - 1, 0); dp < min(dend + 1, D); dp++) {
        const float *gp = gb + (long long)dp * HW;
        const float ctr = ld_nc(gp);
        float q0[2 * R + 1], q1[2 * R + 1], q2[2 * R + 1];
        float v[P2];
        if (dp + kLgaPrefetch < D) lga_prefetch_plane<R>(gp + (long long)kLgaPrefetch * HW, W, interior);
        lga_load_plane<R>(gp, W, interior, noff, v);
#pragma unroll
        for (int r = 0; r < 2 * R + 1; r++) {
            q0[r] = 0.f; q1[r] = 0.f; q2[r] = 0.f;
#pragma unroll
            for (int c = 0; c < 2 * R + 1; c++) {
                const int t = r * (2 * R + 1) + c;
                q0[r] = fmaf(v[t], wn[0 * P2 + t], q0[r]);    // go plane dp is the dd=-1 neighbour of output dp+1
                q1[r] = fmaf(v[t], wn[1 * P2 + t], q1[r]);
                q2[r] = fmaf(v[t], wn[2 * P2 + t], q2[r]);    // ... and the dd=+1 neighbour of output dp-1
            }
        }
        float n0 = 0.f, n1 = 0.f, n2 = 0.f;
#pragma unroll
        for (int r = 0; r < 2 * R + 1; r++) { n0 += q0[r]; n1 += q1[r]; n2 += q2[r]; }
        a0 += n2; a1 += n1; a2 += n0;
        float fb_w = coob;
        if (dp == 0) fb_w += cval[0];
        if (dp == D - 1) fb_w += cval[2];
        a1 = fmaf(ctr, fb_w, a1);
        if (dp - 1 >= dbeg && dp - 1 < dend) gxb[(long long)(dp - 1) * HW] = a0;
        a0 = a1; a1 = a2; a2 = 0.f;
    }
    const int last = min(dend + 1, D) - 1;
    if (last >= dbeg && last < dend) gxb[(long long)last * HW] = a0;
}

static int pick_d_chunk(int64_t B, int64_t D, long long ctas_per_image)
{
    // enough CTAs to fill 148 SMs a few times over, otherwise keep chunks long so
    // the per-pixel filter load is amortised over many depths
    const long long ctas = B * ctas_per_image;
    long long split = (148ll * 8 + ctas - 1) / ctas;
    if (split < 1) split = 1;
    if (split > D) split = D;
    if (B * split > 65535) split = 65535 / B;     // gridDim.z limit
    if (split < 1) split = 1;
    return (int)((D + split - 1) / split);
}

// ---- TMA-tiled variant (lga_tile.cuh): radius 2, rows a multiple of 16 bytes ---------------
static bool lga_tile_enabled()
{
    static int v = -1;
    if (v < 0) v = getenv("GANET_LGA_NO_TILE") ? 0 : 1;
    return v != 0;
}

static bool lga_tile_shape_ok(int64_t B, int64_t D, int64_t H, int64_t W)
{
    return lga_tile_enabled() && (W % 4) == 0 && W >= kBW && H >= kBH && D >= kPD &&
           B * D + kPD < (1ll << 31) && (H + kTH - 1) / kTH <= 65535;
}

static const int kNoTile = -100;     // internal: use the per-pixel LDG kernels below

// forward (MODE 0: src = x) and data backward (MODE 1: src = gradOut)
template <int MODE>
static int lga_tile_launch(const float *src, const float *f, float *dst, int64_t B, int64_t D, int64_t H,
                           int64_t W, cudaStream_t st)
{
    if (!lga_tile_shape_ok(B, D, H, W)) return kNoTile;
    LgaTileMaps maps;
    if (!make_plane_map(&maps.src, src, 4, B * D, (int)H, (int)W, kBW, kPD, kBH)) return kNoTile;
    maps.go = maps.src;
    const long long tiles = ((W + kTW - 1) / kTW) * ((H + kTH - 1) / kTH);
    const int dck = pick_d_chunk(B, D, tiles);
    const int nchunk = (int)((D + dck - 1) / dck);
    dim3 grid((unsigned)((W + kTW - 1) / kTW), (unsigned)((H + kTH - 1) / kTH), (unsigned)(B * nchunk));
    lga_tile_kernel<MODE><<<grid, kTileThreads, 0, st>>>(maps, f, dst, (int)D, (int)H, (int)W, dck);
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

static int lga_tile_filter_launch(const float *x, const float *go, float *gf, int accumulate, int64_t B,
                                  int64_t D, int64_t H, int64_t W, cudaStream_t st)
{
    if (!lga_tile_shape_ok(B, D, H, W) || B > 65535) return kNoTile;
    LgaTileMaps maps;
    if (!make_plane_map(&maps.src, x, 4, B * D, (int)H, (int)W, kBW, kPD, kBH)) return kNoTile;
    if (!make_plane_map(&maps.go, go, 4, B * D, (int)H, (int)W, kTW, kPD, kTH)) return kNoTile;
    dim3 grid((unsigned)((W + kTW - 1) / kTW), (unsigned)((H + kTH - 1) / kTH), (unsigned)B);
    lga_tile_filter_kernel<<<grid, kTileThreads, 0, st>>>(maps, go, gf, accumulate, (int)D, (int)H, (int)W);
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

template <int R>
static int lga_forward_r(const float *x, const float *f, float *y, int64_t B, int64_t D,
                         int64_t H, int64_t W, cudaStream_t st)
{
    if (R == kTR) {
        const int rc = lga_tile_launch<0>(x, f, y, B, D, H, W, st);
        if (rc != kNoTile) return rc;
    }
    const int dck = pick_d_chunk(B, D, H * ((W + kLgaThreads - 1) / kLgaThreads));
    const int nchunk = (int)((D + dck - 1) / dck);
    dim3 grid((unsigned)((W + kLgaThreads - 1) / kLgaThreads), (unsigned)H, (unsigned)(B * nchunk));
    lga_fwd_kernel<R><<<grid, kLgaThreads, 0, st>>>(x, f, y, (int)D, (int)H, (int)W, dck);
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

template <int R>
static int lga_backward_r(const float *x, const float *f, const float *go, float *gx, float *gf,
                          int accumulate, int64_t B, int64_t D, int64_t H, int64_t W,
                          cudaStream_t st)
{
    int rc = (R == kTR) ? lga_tile_filter_launch(x, go, gf, accumulate, B, D, H, W, st) : kNoTile;
    if (rc == kNoTile) {
        dim3 gridf((unsigned)((W + kLgaThreads - 1) / kLgaThreads), (unsigned)H, (unsigned)B);
        lga_bwd_filter_kernel<R><<<gridf, kLgaThreads, 0, st>>>(x, go, gf, accumulate, (int)D, (int)H,
                                                                (int)W);
        GANET_RETURN_IF_LAUNCH_FAILED();
    } else if (rc) {
        return rc;
    }
    rc = (R == kTR) ? lga_tile_launch<1>(go, f, gx, B, D, H, W, st) : kNoTile;
    if (rc != kNoTile) return rc;
    const int dck = pick_d_chunk(B, D, H * ((W + kLgaThreads - 1) / kLgaThreads));
    const int nchunk = (int)((D + dck - 1) / dck);
    dim3 grid((unsigned)((W + kLgaThreads - 1) / kLgaThreads), (unsigned)H, (unsigned)(B * nchunk));
    lga_bwd_data_kernel<R><<<grid, kLgaThreads, 0, st>>>(f, go, gx, (int)D, (int)H, (int)W, dck);
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

static int lga_check(int64_t B, int64_t D, int64_t H, int64_t W, int radius)
{
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || radius < 0) return GANET_EINVAL;
    if (radius > 2) return GANET_EUNSUPPORTED;
    if (H > 65535 || B > 65535) return GANET_EUNSUPPORTED;
    if (D * H * W >= (1ll << 31)) return GANET_EUNSUPPORTED;
    return GANET_OK;
}

}  // namespace ganet

using namespace ganet;

GANET_API int ganet_lga_forward(const float *x, const float *f, float *y, int64_t B, int64_t D,
                                int64_t H, int64_t W, int radius, ganet_stream_t stream)
{
    if (!x || !f || !y || x == y) return GANET_EINVAL;
    int rc = lga_check(B, D, H, W, radius);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (radius) {
    case 0: return lga_forward_r<0>(x, f, y, B, D, H, W, st);
    case 1: return lga_forward_r<1>(x, f, y, B, D, H, W, st);
    default: return lga_forward_r<2>(x, f, y, B, D, H, W, st);
    }
}

GANET_API int ganet_lga_backward(const float *x, const float *f, const float *grad_out,
                                 float *grad_x, float *grad_f, int accumulate_f, int64_t B,
                                 int64_t D, int64_t H, int64_t W, int radius,
                                 ganet_stream_t stream)
{
    if (!x || !f || !grad_out || !grad_x || !grad_f || grad_x == grad_out || grad_x == x)
        return GANET_EINVAL;
    int rc = lga_check(B, D, H, W, radius);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (radius) {
    case 0: return lga_backward_r<0>(x, f, grad_out, grad_x, grad_f, accumulate_f, B, D, H, W, st);
    case 1: return lga_backward_r<1>(x, f, grad_out, grad_x, grad_f, accumulate_f, B, D, H, W, st);
    default: return lga_backward_r<2>(x, f, grad_out, grad_x, grad_f, accumulate_f, B, D, H, W, st);
    }
}
