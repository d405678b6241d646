// Vertical (down / up) SGA scans: the coalesced kernels.
//
// A CTA owns a strip of 32 adjacent pixel columns of one (n,c) slice and walks
// the rows in scan order.  Lane = column (so every global access of a warp is one
// 128-byte row segment of x / out / A / grad, and 32 contiguous mask bytes), warp
// j = the depth chunk [K*j, K*j+K) kept in registers.  Per row the warps exchange
// only what couples depth chunks -- the two chunk-edge values and the chunk max
// (forward) or the chunk sums / arg-max (backward) -- through a few KB of
// double-buffered shared memory, one __syncthreads per row.
//
// Reference: sga_down_forward / sga_up_forward (GANet_kernel.cu:66-127, :285-346),
// sga_{down,up}_data_backward (:129-208, :348-426), *_weight_backward (:210-281,
// :428-505), Max (:23-36), get_temp_grad (:38-48), MaxDepth (:50-64).
#pragma once
#include <math.h>

#include "common.cuh"
#include "sga_step.cuh"

namespace ganet {

enum { VMODE_FIRST = 0, VMODE_SECOND = 1, VMODE_COMBINE = 2, VMODE_RAW = 3 };

struct MaskIds { int first, mine; };   // direction ids written to the mask

// ---------------------------------------------------------------------------
// forward.  dir 0 = down (rows 0..H-1), 1 = up (rows H-1..0).
//   VMODE_FIRST   : out = A                                  (mask untouched)
//   VMODE_SECOND  : m = out < A; out = m ? A : out; mask = m ? ids.mine : ids.first
//                   (every mask byte written: no memset needed anywhere)
//   VMODE_COMBINE : replace when out < A, or out == A and ids.mine < mask -- the
//                   reference's "ties keep the lower direction id" (Max, :31) made
//                   independent of the order in which directions are merged
//   VMODE_RAW     : out = A
// `dir` is the scan direction in the layout at hand (0 = rows ascending); the mask ids
// are separate because the horizontal scans run as vertical scans on transposed data.
// smem: float[2][3][NW][32]
// ---------------------------------------------------------------------------
template <int K, int MAXW, int MODE>
__global__ void __launch_bounds__(MAXW * 32)
sga_vert_fwd_kernel(const float *__restrict__ x, const float *__restrict__ g, float *out,
                    uint8_t *mask, int dir, MaskIds ids, int D, int H, int W, int strips)
{
    static_assert(K % 2 == 0, "depth parity must be a compile-time property");
    extern __shared__ float ex[];
    const int lane = threadIdx.x & 31, j = threadIdx.x >> 5, NW = blockDim.x >> 5;
    const long long s = blockIdx.x / strips;
    const int strip = blockIdx.x - (int)(s * strips);
    const int wcol = strip * 32 + lane;
    const bool wok = wcol < W;
    const int wc = wok ? wcol : W - 1;
    const int HW = H * W;
    const long long S = (long long)D * HW;
    const float *xs = x + s * S;
    const float *gs = g + s * 5ll * HW;
    float *os = out + s * S;
    uint8_t *ms = (MODE == VMODE_RAW || MODE == VMODE_FIRST) ? nullptr : mask + s * S;

    const int d0 = K * j;
    int off[K];
#pragma unroll
    for (int i = 0; i < K; i++) off[i] = min(d0 + i, D - 1) * HW;

    int p = (dir == 0) ? wc : (H - 1) * W + wc;
    const int ps = (dir == 0) ? W : -W;
    const int plane = NW * 32;               // one exchange array
    float P[K], xc[K], w[5];
#pragma unroll
    for (int i = 0; i < K; i++) { xc[i] = ld_nc(xs + off[i] + p); P[i] = 0.f; }
#pragma unroll
    for (int k = 0; k < 5; k++) w[k] = ld_nc(gs + k * HW + p);

    for (int t = 0; t < H; t++) {
        const int pn = p + ps;
        float xn[K], wn[5], oc[K];
        uint8_t mc[K];
        if (t + 1 < H) {
#pragma unroll
            for (int i = 0; i < K; i++) xn[i] = ld_nc(xs + off[i] + pn);
#pragma unroll
            for (int k = 0; k < 5; k++) wn[k] = ld_nc(gs + k * HW + pn);
        } else {
#pragma unroll
            for (int i = 0; i < K; i++) xn[i] = 0.f;
#pragma unroll
            for (int k = 0; k < 5; k++) wn[k] = 0.f;
        }
        if (MODE == VMODE_SECOND || MODE == VMODE_COMBINE) {
#pragma unroll
            for (int i = 0; i < K; i++) oc[i] = os[off[i] + p];
        }
        if (MODE == VMODE_COMBINE) {
#pragma unroll
            for (int i = 0; i < K; i++) mc[i] = ms[off[i] + p];
        }

        float A[K];
        if (t == 0) {
            sga_first_step<K>(xc, w, A);
        } else {
            const float *eb = ex + ((t - 1) & 1) * 3 * plane;
            const float up = (j > 0) ? eb[plane + (j - 1) * 32 + lane] : 0.f;       // P[d0-1]
            const float dn = (j + 1 < NW) ? eb[(j + 1) * 32 + lane] : 0.f;          // P[d0+K]
            float pmax = eb[2 * plane + lane];
            for (int jj = 1; jj < NW; jj++) pmax = fmaxf(pmax, eb[2 * plane + jj * 32 + lane]);
            sga_next_step<K>(P, xc, w, up, dn, pmax, d0, D, A);
        }
        {
            float *wb = ex + (t & 1) * 3 * plane;
            wb[j * 32 + lane] = A[0];
            wb[plane + j * 32 + lane] = A[K - 1];
            wb[2 * plane + j * 32 + lane] = chunk_max<K>(A, d0, D);
        }
        if (wok) {
#pragma unroll
            for (int i = 0; i < K; i++) {
                if (d0 + i < D) {
                    const int e = off[i] + p;
                    if (MODE == VMODE_FIRST || MODE == VMODE_RAW) {
                        os[e] = A[i];
                    } else if (MODE == VMODE_SECOND) {
                        const bool m = oc[i] < A[i];
                        if (m) os[e] = A[i];
                        ms[e] = m ? (uint8_t)ids.mine : (uint8_t)ids.first;
                    } else {
                        const bool m = oc[i] < A[i] || (oc[i] == A[i] && ids.mine < (int)mc[i]);
                        if (m) { os[e] = A[i]; ms[e] = (uint8_t)ids.mine; }
                    }
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < K; i++) { P[i] = A[i]; xc[i] = xn[i]; }
#pragma unroll
        for (int k = 0; k < 5; k++) w[k] = wn[k];
        p = pn;
    }
}

// ---------------------------------------------------------------------------
// backward of one vertical direction, reverse sweep.
// `a` holds this direction's aggregate (VMODE_RAW pass).  gg is overwritten.
// smem: float[2][NBW][NW][32]   (NBW exchange arrays, see BX_* below)
// ---------------------------------------------------------------------------
enum { BX_TLO = 0, BX_THI, BX_ST, BX_S0, BX_S1, BX_S2, BX_S3, BX_AMAX, BX_AIDX, NBW };

template <int K, int MAXW>
__global__ void __launch_bounds__(MAXW * 32)
sga_vert_bwd_kernel(const float *__restrict__ x, const float *__restrict__ g,
                    const float *__restrict__ a, const uint8_t *__restrict__ mask,
                    const float *__restrict__ go, float *gi, float *__restrict__ gg, int dir,
                    int mask_id, int accumulate, int D, int H, int W, int strips)
{
    extern __shared__ float ex[];
    const int lane = threadIdx.x & 31, j = threadIdx.x >> 5, NW = blockDim.x >> 5;
    const long long s = blockIdx.x / strips;
    const int strip = blockIdx.x - (int)(s * strips);
    const int wcol = strip * 32 + lane;
    const bool wok = wcol < W;
    const int wc = wok ? wcol : W - 1;
    const int HW = H * W;
    const long long S = (long long)D * HW;
    const float *xs = x + s * S;
    const float *as = a + s * S;
    const float *gos = go + s * S;
    const uint8_t *ms = mask + s * S;
    float *gis = gi + s * S;
    const float *gs = g + s * 5ll * HW;
    float *ggs = gg + s * 5ll * HW;

    const int d0 = K * j;
    int off[K];
#pragma unroll
    for (int i = 0; i < K; i++) off[i] = min(d0 + i, D - 1) * HW;
    const int off_m1 = max(d0 - 1, 0) * HW;            // depth d0-1 (edge of the previous chunk)
    const int off_pK = min(d0 + K, D - 1) * HW;        // depth d0+K (edge of the next chunk)

    const int ps = (dir == 0) ? W : -W;                // forward scan step
    int p = ((dir == 0) ? wc : (H - 1) * W + wc) + (H - 1) * ps;   // last scan position
    const int plane = NW * 32;
    const int bufsz = NBW * plane;

    float Tn[K], wnx[5];
#pragma unroll
    for (int i = 0; i < K; i++) Tn[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 5; k++) wnx[k] = 0.f;
    int p_next = p;                                    // pixel of scan position t+1

    for (int t = H - 1; t >= 0; t--) {
        const int pq = p - ps;                         // scan position t-1
        float xv[K], t0[K], ap[K], w[5], gold[K];
        float aup = 0.f, adn = 0.f;
#pragma unroll
        for (int i = 0; i < K; i++) {
            const int e = off[i] + p;
            xv[i] = ld_nc(xs + e);
            const float gv = ld_nc(gos + e);
            const uint8_t mv = ms[e];
            t0[i] = (d0 + i < D && mv == mask_id) ? gv : 0.f;        // get_temp_grad :38-48
            ap[i] = (t >= 1) ? ld_nc(as + off[i] + pq) : 0.f;
            gold[i] = accumulate ? gis[e] : 0.f;
        }
        if (t >= 1) {
            aup = ld_nc(as + off_m1 + pq);             // A[d0-1, t-1]
            adn = ld_nc(as + off_pK + pq);             // A[d0+K, t-1]
        }
#pragma unroll
        for (int k = 0; k < 5; k++) w[k] = ld_nc(gs + k * HW + p);

        float tc[K];
        if (t + 1 < H) {
            // everything the previous iteration (scan position t+1) published
            const float *eb = ex + ((t + 1) & 1) * bufsz;
            const float up = (j > 0) ? eb[BX_THI * plane + (j - 1) * 32 + lane] : 0.f;
            const float dn = (j + 1 < NW) ? eb[BX_TLO * plane + (j + 1) * 32 + lane] : 0.f;
            float sum_tn = 0.f, amax = -INFINITY;
            int idx_cur = 0x7fffffff;
            for (int jj = 0; jj < NW; jj++) {
                sum_tn += eb[BX_ST * plane + jj * 32 + lane];
                const float v = eb[BX_AMAX * plane + jj * 32 + lane];
                const int vi = __float_as_int(eb[BX_AIDX * plane + jj * 32 + lane]);
                if (v > amax) { amax = v; idx_cur = vi; }          // strict >: first maximum
            }
            // guidance gradients of scan position t+1 are complete now: one warp per weight
            for (int k = j; k < 5; k += NW) {
                float tot;
                if (k == 4) {
                    tot = sum_tn * amax;                           // (:265-272)
                } else {
                    tot = 0.f;
                    for (int jj = 0; jj < NW; jj++) tot += eb[(BX_S0 + k) * plane + jj * 32 + lane];
                }
                if (wok) ggs[k * HW + p_next] = tot;
            }
            const float inj = sum_tn * wnx[4];
#pragma unroll
            for (int i = 0; i < K; i++) {
                const int d = d0 + i;
                const float tm = (i == 0) ? up : Tn[i == 0 ? 0 : i - 1];
                const float tp = (i == K - 1) ? dn : Tn[i == K - 1 ? K - 1 : i + 1];
                float v = t0[i];
                v += Tn[i] * wnx[1];
                if (d + 1 < D) v += tp * wnx[2];
                if (d >= 1) v += tm * wnx[3];
                if (d == idx_cur) v += inj;
                tc[i] = (d < D) ? v : 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < K; i++) tc[i] = t0[i];
        }

        if (wok) {
#pragma unroll
            for (int i = 0; i < K; i++) {
                const int d = d0 + i;
                if (d < D) {
                    float v = tc[i] * w[0];
                    if (d == 0) v += tc[i] * w[2];
                    if (d == D - 1) v += tc[i] * w[3];
                    gis[off[i] + p] = gold[i] + v;
                }
            }
        }

        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, st = 0.f;
        float best = -INFINITY;
        int bi = d0;
#pragma unroll
        for (int i = 0; i < K; i++) {
            const int d = d0 + i;
            s0 += tc[i] * xv[i];
            st += tc[i];
            if (t >= 1) {
                const float am = (i == 0) ? aup : ap[i == 0 ? 0 : i - 1];
                const float apn = (i == K - 1) ? adn : ap[i == K - 1 ? K - 1 : i + 1];
                s1 += tc[i] * ap[i];
                s2 += tc[i] * ((d >= 1) ? am : xv[i]);
                s3 += tc[i] * ((d + 1 < D) ? apn : xv[i]);
                if (d < D && ap[i] > best) { best = ap[i]; bi = d; }
            }
        }
        {
            float *wb = ex + (t & 1) * bufsz;
            const int o = j * 32 + lane;
            wb[BX_TLO * plane + o] = tc[0];
            wb[BX_THI * plane + o] = tc[K - 1];
            wb[BX_ST * plane + o] = st;
            wb[BX_S0 * plane + o] = s0;
            wb[BX_S1 * plane + o] = s1;
            wb[BX_S2 * plane + o] = s2;
            wb[BX_S3 * plane + o] = s3;
            wb[BX_AMAX * plane + o] = best;
            wb[BX_AIDX * plane + o] = __int_as_float(bi);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < K; i++) Tn[i] = tc[i];
#pragma unroll
        for (int k = 0; k < 5; k++) wnx[k] = w[k];
        p_next = p;
        p = pq;
    }
    // scan position 0: only w0 receives a gradient (Appendix A.3 quirk)
    {
        const float *eb = ex + 0 * bufsz;
        for (int k = j; k < 5; k += NW) {
            float tot = 0.f;
            if (k == 0)
                for (int jj = 0; jj < NW; jj++) tot += eb[BX_S0 * plane + jj * 32 + lane];
            if (wok) ggs[k * HW + p_next] = tot;
        }
    }
}

}  // namespace ganet
