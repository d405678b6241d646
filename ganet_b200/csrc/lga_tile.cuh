// LGA for sm_100a, radius 2, TMA-tiled variant of the kernels in lga.cu.
//
// Same arithmetic as lga.cu (same summation order, bit-identical results), different data
// movement.  lga.cu gives every pixel thread 25 global loads per depth plane (its 5x5
// neighbourhood, served by L1); at 8-12 resident warps per SM (75 tap weights live in
// registers) those loads are latency-bound (profiles/r01_ncu_full_lga_kernels.txt: issue
// slots 28-48 % busy).  Here a CTA owns a 32 x 4 pixel tile; a producer warp streams the
// (32+8) x (4+4) halo tile of kPD consecutive depth planes per stage into shared memory with
// one TMA box (out-of-image elements arrive as zeros -- they are only ever multiplied by a
// zero weight), and the four consumer warps read their neighbourhoods with LDS at
// compile-time offsets: 2.5 global elements per output instead of 25, no address arithmetic
// in the plane loop, no bounds predicates.
//
// Reference semantics: GANet_kernel.cu:1131-1269 (see lga.cu for the derivation of the
// centre-fallback terms).
#pragma once
#include "common.cuh"
#include "tma_utils.cuh"

namespace ganet {

constexpr int kTR = 2;                       // radius served by this variant
constexpr int kTW = 32, kTH = 4;             // pixel tile of a CTA (one warp per tile row)
constexpr int kPadL = 4;                     // the box starts 4 columns left of the tile: TMA wants the
                                             // start of a box row 16-byte aligned, so the 2-column halo
                                             // is rounded up to 4 (measured: a start at w0-2 faults)
constexpr int kBW = kTW + 2 * kPadL;         // 40 floats = 160 B box rows
constexpr int kBH = kTH + 2 * kTR;           // 8 box rows
constexpr int kPD = 4;                       // depth planes per stage
constexpr int kTS = 4;                       // stages in the ring
constexpr int kTileThreads = (kTH + 1) * 32; // consumers + one producer warp
constexpr int kPlaneFloats = kBH * kBW;      // 288
constexpr int kStageFloats = kPD * kPlaneFloats;
constexpr int kGoStageFloats = kPD * kTH * kTW;
static_assert((kStageFloats * 4) % 128 == 0 && (kGoStageFloats * 4) % 128 == 0, "TMA destinations are 128-byte aligned");

struct LgaTileMaps { CUtensorMap src, go; };  // go: filter backward only

struct LgaTilePos {
    int w0, h0, w, h, ty, tx;
    bool active;
};

__device__ __forceinline__ LgaTilePos lga_tile_pos(int H, int W)
{
    LgaTilePos p;
    p.tx = threadIdx.x & 31;
    p.ty = threadIdx.x >> 5;
    p.w0 = blockIdx.x * kTW;
    p.h0 = blockIdx.y * kTH;
    p.w = p.w0 + p.tx;
    p.h = p.h0 + p.ty;
    p.active = p.ty < kTH && p.w < W && p.h < H;
    return p;
}

// ---- forward (MODE 0) and data backward (MODE 1) ------------------------------------------
// MODE 0: src = x,       weights = this pixel's taps                       (:1131-1175)
// MODE 1: src = gradOut, weights = the mirrored tap of each neighbour pixel (:1218-1269)
template <int MODE>
__global__ void __launch_bounds__(kTileThreads, 3)
lga_tile_kernel(const __grid_constant__ LgaTileMaps maps, const float *__restrict__ f,
                float *__restrict__ dst, int D, int H, int W, int d_chunk)
{
    constexpr int R = kTR, WS = 2 * R + 1, P2 = WS * WS, F = 3 * P2;
    __shared__ __align__(128) float tiles[kTS][kStageFloats];
    __shared__ __align__(8) uint64_t full[kTS], empty[kTS];

    const LgaTilePos q = lga_tile_pos(H, W);
    const int nchunk = (D + d_chunk - 1) / d_chunk;
    const long long b = blockIdx.z / nchunk;
    const int dc = blockIdx.z % nchunk;
    const int dbeg = dc * d_chunk, dend = min(D, dbeg + d_chunk);
    const int p_lo = max(dbeg - 1, 0), p_hi = min(dend + 1, D);     // input planes [p_lo, p_hi)
    const int nst = (p_hi - p_lo + kPD - 1) / kPD;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kTS; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], kTH); }
        fence_mbarrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    if (q.ty == kTH) {                                    // ---------------- producer warp
        if (q.tx == 0) {
            const int c2 = (int)(b * D) + p_lo;
            for (int i = 0; i < nst; i++) {
                const int s = i % kTS;
                if (i >= kTS) mbar_wait(&empty[s], ((i / kTS) - 1) & 1);
                mbar_arrive_expect_tx(&full[s], kStageFloats * 4);
                tma_load_3d(tiles[s], &maps.src, &full[s], q.w0 - kPadL, q.h0 - R, c2 + i * kPD);
            }
        }
        return;
    }

    // ---------------- consumers: one pixel per thread
    const int HW = H * W;
    const int wc = min(q.w, W - 1), hc = min(q.h, H - 1);          // surplus threads shadow a real pixel
    const float *fb = f + b * (long long)F * HW + hc * W + wc;
    float *yb = dst + b * (long long)D * HW + hc * W + wc;

    float wz[F];                      // tap weights, 0 where (r,c) leaves the image
    float cval[3] = {0.f, 0.f, 0.f};  // this pixel's own in-image weights per depth tap
    float coob = 0.f;                 // this pixel's own out-of-image weights
    if (MODE == 0) {
#pragma unroll
        for (int dd = 0; dd < 3; dd++)
#pragma unroll
            for (int t = 0; t < P2; t++) {
                const int r = t / WS - R, c = t % WS - R;
                const bool ok = (hc + r >= 0) && (hc + r < H) && (wc + c >= 0) && (wc + c < W);
                const float v = ld_nc(fb + (long long)(dd * P2 + t) * HW);
                wz[dd * P2 + t] = ok ? v : 0.f;
                if (ok) cval[dd] += v; else coob += v;
            }
    } else {
#pragma unroll
        for (int t = 0; t < P2; t++) {
            const int r = t / WS - R, c = t % WS - R;
            const bool ok = (hc + r >= 0) && (hc + r < H) && (wc + c >= 0) && (wc + c < W);
#pragma unroll
            for (int dd = 0; dd < 3; dd++) {
                // the neighbour at (+r,+c) one depth tap away reaches back with its mirrored tap
                const int loc_m = (2 - dd) * P2 + (-r + R) * WS + (-c + R);
                wz[dd * P2 + t] = ok ? ld_nc(fb + r * W + c + (long long)loc_m * HW) : 0.f;
                const float own = ld_nc(fb + (long long)(dd * P2 + t) * HW);
                if (ok) cval[dd] += own; else coob += own;
            }
        }
    }

    // rolling outputs: a0 = y[dp-1] (complete after plane dp), a1 = y[dp], a2 = y[dp+1]
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int i = 0; i < nst; i++) {
        const int s = i % kTS;
        mbar_wait(&full[s], (i / kTS) & 1);
        const float *tp = &tiles[s][q.ty * kBW + q.tx + (kPadL - R)];   // top-left of this pixel's window
#pragma unroll
        for (int pp = 0; pp < kPD; pp++) {
            const int dp = p_lo + i * kPD + pp;
            if (dp < p_hi) {
                const float *pl = tp + pp * kPlaneFloats;
                const float ctr = pl[R * kBW + R];
                // 75 FMAs per plane as 35 packed fp32 FMAs (fma.rn.f32x2, FFMA2) + 5 scalar ones.  Each
                // lane of a packed FMA is an ordinary IEEE fma and every accumulator chain keeps its
                // order (per window row, columns left to right), so the bits are those of lga.cu.
                //   q01[r] = (depth tap -1 of output dp+1, depth tap 0 of output dp), same input voxel
                //   q2     = depth tap +1 of output dp-1; rows (0,1) and (2,3) share an instruction
                float2 q01[WS], q2a, q2b;
                float q2c = 0.f;
                q2a = make_float2(0.f, 0.f); q2b = make_float2(0.f, 0.f);
#pragma unroll
                for (int r = 0; r < WS; r++) q01[r] = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < WS; c++) {
                    float v[WS];
#pragma unroll
                    for (int r = 0; r < WS; r++) v[r] = pl[r * kBW + c];
#pragma unroll
                    for (int r = 0; r < WS; r++) {
                        const int t = r * WS + c;
                        q01[r] = __ffma2_rn(make_float2(v[r], v[r]), make_float2(wz[0 * P2 + t], wz[1 * P2 + t]), q01[r]);
                    }
                    q2a = __ffma2_rn(make_float2(v[0], v[1]), make_float2(wz[2 * P2 + 0 * WS + c], wz[2 * P2 + 1 * WS + c]), q2a);
                    q2b = __ffma2_rn(make_float2(v[2], v[3]), make_float2(wz[2 * P2 + 2 * WS + c], wz[2 * P2 + 3 * WS + c]), q2b);
                    q2c = fmaf(v[4], wz[2 * P2 + 4 * WS + c], q2c);
                }
                float n0 = 0.f, n1 = 0.f, n2 = 0.f;
#pragma unroll
                for (int r = 0; r < WS; r++) { n0 += q01[r].x; n1 += q01[r].y; }
                n2 += q2a.x; n2 += q2a.y; n2 += q2b.x; n2 += q2b.y; n2 += q2c;
                a0 += n2; a1 += n1; a2 += n0;
                // centre-fallback terms of output dp: out-of-image taps always, plus the whole
                // -1 / +1 depth tap at the volume faces
                float fb_w = coob;
                if (dp == 0) fb_w += cval[0];
                if (dp == D - 1) fb_w += cval[2];
                a1 = fmaf(ctr, fb_w, a1);
                if (q.active && dp - 1 >= dbeg && dp - 1 < dend) yb[(long long)(dp - 1) * HW] = a0;
                a0 = a1; a1 = a2; a2 = 0.f;
            }
        }
        __syncwarp();
        if (q.tx == 0) mbar_arrive(&empty[s]);
    }
    const int last = p_hi - 1;                   // plane processed last
    if (q.active && last >= dbeg && last < dend) yb[(long long)last * HW] = a0;   // only when dend == D
}

// ---- filter backward (:1177-1216) ------------------------------------------------------------
// gf[dd][t] = sum_d go[d] * (x[d+dd] at neighbour t, or the centre x[d] when that voxel is
// outside).  One thread per pixel, all depths, 75 accumulators in registers; x planes as halo
// tiles, gradOut planes as plain 32 x 4 tiles (one plane ahead: plane dp of x meets go[dp+1]).
__global__ void __launch_bounds__(kTileThreads, 3)
lga_tile_filter_kernel(const __grid_constant__ LgaTileMaps maps, const float *__restrict__ go,
                       float *__restrict__ gf, int accumulate, int D, int H, int W)
{
    constexpr int R = kTR, WS = 2 * R + 1, P2 = WS * WS, F = 3 * P2;
    __shared__ __align__(128) float tiles[kTS][kStageFloats];
    __shared__ __align__(128) float gtiles[kTS][kGoStageFloats];
    __shared__ __align__(8) uint64_t full[kTS], empty[kTS];

    const LgaTilePos q = lga_tile_pos(H, W);
    const long long b = blockIdx.z;
    const int nst = (D + kPD - 1) / kPD;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kTS; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], kTH); }
        fence_mbarrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    if (q.ty == kTH) {                                    // ---------------- producer warp
        if (q.tx == 0) {
            const int c2 = (int)(b * D);
            for (int i = 0; i < nst; i++) {
                const int s = i % kTS;
                if (i >= kTS) mbar_wait(&empty[s], ((i / kTS) - 1) & 1);
                mbar_arrive_expect_tx(&full[s], (kStageFloats + kGoStageFloats) * 4);
                tma_load_3d(tiles[s], &maps.src, &full[s], q.w0 - kPadL, q.h0 - R, c2 + i * kPD);
                tma_load_3d(gtiles[s], &maps.go, &full[s], q.w0, q.h0, c2 + i * kPD + 1);   // go[dp+1]
            }
        }
        return;
    }

    const int HW = H * W;
    const int wc = min(q.w, W - 1), hc = min(q.h, H - 1);          // surplus threads shadow a real pixel
    float *gfb = gf + b * (long long)F * HW + hc * W + wc;

    float acc[F];
#pragma unroll
    for (int l = 0; l < F; l++) acc[l] = 0.f;

    // plane dp of x meets go[dp+1] (depth tap -1), go[dp] (tap 0), go[dp-1] (tap +1)
    float gm = 0.f;                                                   // go[dp-1]
    float gc = ld_nc(go + b * (long long)D * HW + hc * W + wc);       // go[dp], dp = 0
    float sgc = 0.f, e_first = 0.f, e_last = 0.f;                     // sum go*x centre; face terms
    for (int i = 0; i < nst; i++) {
        const int s = i % kTS;
        mbar_wait(&full[s], (i / kTS) & 1);
        const float *tp = &tiles[s][q.ty * kBW + q.tx + (kPadL - R)];
        const float *gt = &gtiles[s][q.ty * kTW + q.tx];
#pragma unroll
        for (int pp = 0; pp < kPD; pp++) {
            const int dp = i * kPD + pp;
            if (dp < D) {
                const float gp = (dp + 1 < D) ? gt[pp * kTH * kTW] : 0.f;    // go[dp+1]
                const float *pl = tp + pp * kPlaneFloats;
                const float ctr = pl[R * kBW + R];
                sgc = fmaf(gc, ctr, sgc);
                if (dp == 0) e_first = gc * ctr;
                if (dp == D - 1) e_last = gc * ctr;
                // packed fp32 FMAs: (tap -1, tap 0) of one neighbour share an instruction, the tap +1
                // accumulators of two neighbours share one; each accumulator still sums over depth in order
                const float2 gpc = make_float2(gp, gc), gmm = make_float2(gm, gm);
#pragma unroll
                for (int t = 0; t < P2; t++) {
                    const float v = pl[(t / WS) * kBW + (t % WS)];
                    const float2 r = __ffma2_rn(gpc, make_float2(v, v), make_float2(acc[0 * P2 + t], acc[1 * P2 + t]));
                    acc[0 * P2 + t] = r.x; acc[1 * P2 + t] = r.y;
                }
#pragma unroll
                for (int t = 0; t + 1 < P2; t += 2) {
                    const float v0 = pl[(t / WS) * kBW + (t % WS)], v1 = pl[((t + 1) / WS) * kBW + ((t + 1) % WS)];
                    const float2 r = __ffma2_rn(gmm, make_float2(v0, v1), make_float2(acc[2 * P2 + t], acc[2 * P2 + t + 1]));
                    acc[2 * P2 + t] = r.x; acc[2 * P2 + t + 1] = r.y;
                }
                {
                    const int t = P2 - 1;
                    acc[2 * P2 + t] = fmaf(gm, pl[(t / WS) * kBW + (t % WS)], acc[2 * P2 + t]);
                }
                gm = gc; gc = gp;
            }
        }
        __syncwarp();
        if (q.tx == 0) mbar_arrive(&empty[s]);
    }
    if (!q.active) return;
#pragma unroll
    for (int dd = 0; dd < 3; dd++)
#pragma unroll
        for (int t = 0; t < P2; t++) {
            const int r = t / WS - R, c = t % WS - R;
            const bool ok = (hc + r >= 0) && (hc + r < H) && (wc + c >= 0) && (wc + c < W);
            // out-of-image tap: every depth falls back to the centre; in-image tap: only the
            // face depth whose d+dd leaves the volume does
            const float v = ok ? acc[dd * P2 + t] + (dd == 0 ? e_first : dd == 2 ? e_last : 0.f) : sgc;
            float *dstp = gfb + (long long)(dd * P2 + t) * HW;
            *dstp = accumulate ? *dstp + v : v;
        }
}

}  // namespace ganet
