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

enum { VMODE_FIRST = 0, VMODE_SECOND = 1, VMODE_COMBINE = 2, VMODE_RAW = 3, VMODE_FIRST3 = 4 };

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
template <int K, int MAXW, int MODE, bool FULL>
__global__ void __launch_bounds__(MAXW * 32)
sga_vert_fwd_kernel(const float *__restrict__ x, const float *__restrict__ g, float *out,
                    uint8_t *mask, int dir, MaskIds ids, int D, int H, int W, int strips)
{
    // FULL: NW * K == D, so no thread owns a depth >= D and every depth predicate
    // disappears.  Addresses are walked as pointers (plane stride HW per depth, row
    // stride ps per scan step): rebuilding base + (offset + pixel) * 4 for every
    // access cost 6 integer instructions per load in the first version of this kernel.
    static_assert(K % 2 == 0, "depth parity must be a compile-time property");
    extern __shared__ float ex[];
    const int lane = threadIdx.x & 31, j = threadIdx.x >> 5, NW = blockDim.x >> 5;
    const long long s = blockIdx.x / strips;
    const int strip = blockIdx.x - (int)(s * strips);
    const int wcol = strip * 32 + lane;
    const bool wok = wcol < W;
    const int wc = wok ? wcol : W - 1;
    const long long HW = (long long)H * W;
    const long long S = (long long)D * HW;
    const int d0 = K * j;
    const int dfirst = FULL ? d0 : min(d0, D - 1);
    const long long ps = (dir == 0) ? W : -W;
    const long long e0 = s * S + dfirst * HW + ((dir == 0) ? wc : (long long)(H - 1) * W + wc);
    // row base addresses (depth d0, current scan row) as integers, walked by `ps` per row
    addr_t xrow = (addr_t)(x + e0);
    addr_t orow = (addr_t)(out + e0);
    addr_t mrow = (MODE == VMODE_RAW || MODE == VMODE_FIRST) ? 0 : (addr_t)(mask + e0);
    addr_t grow = (addr_t)(g + s * 5 * HW + ((dir == 0) ? wc : (long long)(H - 1) * W + wc));
    const long long psb = ps * 4;                 // row step in bytes (fp32 tensors)

    const int plane = NW * 32;               // one exchange array
    unsigned offb[K];                        // byte offset of depth d0+i within the row base
#pragma unroll
    for (int i = 0; i < K; i++) {
        offb[i] = (unsigned)(i * (int)HW) * 4u;
        asm volatile("" : "+r"(offb[i]));    // per-thread register, not re-derived each use
    }
    float P[K], xc[K], w[5];
#pragma unroll
    for (int i = 0; i < K; i++) {
        xc[i] = (FULL || d0 + i < D) ? ld_nc(at<const float>(xrow, offb[i])) : 0.f;
        P[i] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 5; k++) w[k] = ld_nc(at<const float>(grow, (unsigned)(k * (int)HW) * 4u));

    for (int t = 0; t < H; t++) {
        float xn[K], wn[5], oc[K];
        uint8_t mc[K];
        if (t + 1 < H) {
            const addr_t xnext = xrow + psb, gnext = grow + psb;
#pragma unroll
            for (int i = 0; i < K; i++)
                xn[i] = (FULL || d0 + i < D) ? ld_nc(at<const float>(xnext, offb[i])) : 0.f;
#pragma unroll
            for (int k = 0; k < 5; k++) wn[k] = ld_nc(at<const float>(gnext, (unsigned)(k * (int)HW) * 4u));
        } else {
#pragma unroll
            for (int i = 0; i < K; i++) xn[i] = 0.f;
#pragma unroll
            for (int k = 0; k < 5; k++) wn[k] = 0.f;
        }
        if (MODE == VMODE_SECOND || MODE == VMODE_COMBINE) {
#pragma unroll
            for (int i = 0; i < K; i++) oc[i] = (FULL || d0 + i < D) ? *at<const float>(orow, offb[i]) : 0.f;
        }
        if (MODE == VMODE_COMBINE) {
#pragma unroll
            for (int i = 0; i < K; i++)
                mc[i] = (FULL || d0 + i < D) ? *at<const uint8_t>(mrow, offb[i] >> 2) : (uint8_t)0;
        }

        float A[K];
        if (t == 0) {
            sga_first_step<K>(xc, w, A);
        } else {
            const float *eb = ex + ((t - 1) & 1) * 3 * plane;
            const float up = (j > 0) ? eb[plane + (j - 1) * 32 + lane] : 0.f;       // P[d0-1]
            const float dn = (j + 1 < NW) ? eb[(j + 1) * 32 + lane] : 0.f;          // P[d0+K]
            const float *mx = eb + 2 * plane + lane;
            float pmax = mx[0];
            for (int jj = 1; jj < NW; jj++) pmax = fmaxf(pmax, mx[jj * 32]);
            sga_next_step<K, FULL>(P, xc, w, up, dn, pmax, d0, D, A);
        }
        {
            float *wb = ex + (t & 1) * 3 * plane + j * 32 + lane;
            wb[0] = A[0];
            wb[plane] = A[K - 1];
            wb[2 * plane] = FULL ? chunk_max<K>(A, 0, K) : chunk_max<K>(A, d0, D);
        }
        if (wok) {
#pragma unroll
            for (int i = 0; i < K; i++) {
                if (FULL || d0 + i < D) {
                    if (MODE == VMODE_FIRST || MODE == VMODE_RAW) {
                        *at<float>(orow, offb[i]) = A[i];
                    } else if (MODE == VMODE_SECOND) {
                        const bool m = oc[i] < A[i];
                        if (m) *at<float>(orow, offb[i]) = A[i];
                        *at<uint8_t>(mrow, offb[i] >> 2) = m ? (uint8_t)ids.mine : (uint8_t)ids.first;
                    } else {
                        const bool m = oc[i] < A[i] || (oc[i] == A[i] && ids.mine < (int)mc[i]);
                        if (m) {
                            *at<float>(orow, offb[i]) = A[i];
                            *at<uint8_t>(mrow, offb[i] >> 2) = (uint8_t)ids.mine;
                        }
                    }
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < K; i++) { P[i] = A[i]; xc[i] = xn[i]; }
#pragma unroll
        for (int k = 0; k < 5; k++) w[k] = wn[k];
        xrow += psb; orow += psb; grow += psb;
        if (MODE == VMODE_SECOND || MODE == VMODE_COMBINE) mrow += ps;      // bytes: u8 tensor
    }
}

// ---------------------------------------------------------------------------
// backward of one vertical direction, reverse sweep.
// `a` holds this direction's aggregate (VMODE_RAW pass).  gg is overwritten.
// smem: float[2][NBW][NW][32]   (NBW exchange arrays, see BX_* below)
// ---------------------------------------------------------------------------
enum { BX_TLO = 0, BX_THI, BX_ST, BX_S0, BX_S1, BX_S2, BX_S3, BX_AMAX, BX_AIDX, NBW };

template <int K, int MAXW, bool FULL>
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
    const long long HW = (long long)H * W;
    const long long S = (long long)D * HW;
    const int d0 = K * j;
    const int dfirst = FULL ? d0 : min(d0, D - 1);
    const long long ps = (dir == 0) ? W : -W;                 // forward scan step
    // last scan position of this column
    const long long pix = ((dir == 0) ? wc : (long long)(H - 1) * W + wc) + (H - 1) * ps;
    const long long e0 = s * S + dfirst * HW + pix;
    const long long psb = ps * 4;
    addr_t xrow = (addr_t)(x + e0);
    addr_t arow = (addr_t)(a + e0 - ps);                      // aggregate at scan position t-1
    addr_t gorow = (addr_t)(go + e0);
    addr_t mrow = (addr_t)(mask + e0);
    addr_t girow = (addr_t)(gi + e0);
    addr_t grow = (addr_t)(g + s * 5 * HW + pix);
    float *ggrow = gg + s * 5 * HW + pix;
    // neighbours across the chunk edges come straight from global memory
    addr_t auprow = (addr_t)(a + e0 - ps + ((d0 >= 1) ? -HW : 0));                        // depth d0-1
    addr_t adnrow = (addr_t)(a + e0 - ps + (long long)(min(d0 + K, D - 1) - dfirst) * HW);   // depth d0+K
    unsigned offb[K];
#pragma unroll
    for (int i = 0; i < K; i++) {
        offb[i] = (unsigned)(i * (int)HW) * 4u;
        asm volatile("" : "+r"(offb[i]));
    }

    const int plane = NW * 32;
    const int bufsz = NBW * plane;
    float Tn[K], wnx[5];
#pragma unroll
    for (int i = 0; i < K; i++) Tn[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 5; k++) wnx[k] = 0.f;
    float *gg_next = ggrow;                                   // guidance-gradient pixel of position t+1

    for (int t = H - 1; t >= 0; t--) {
        float xv[K], t0[K], ap[K], w[5], gold[K];
        float aup = 0.f, adn = 0.f;
        {
#pragma unroll
            for (int i = 0; i < K; i++) {
                const bool ok = FULL || d0 + i < D;
                xv[i] = ok ? ld_nc(at<const float>(xrow, offb[i])) : 0.f;
                const float gv = ok ? ld_nc(at<const float>(gorow, offb[i])) : 0.f;
                const uint8_t mv = ok ? *at<const uint8_t>(mrow, offb[i] >> 2) : (uint8_t)255;
                t0[i] = (mv == mask_id) ? gv : 0.f;                       // get_temp_grad :38-48
                ap[i] = (ok && t >= 1) ? ld_nc(at<const float>(arow, offb[i])) : 0.f;
                gold[i] = (ok && accumulate) ? *at<const float>(girow, offb[i]) : 0.f;
            }
            if (t >= 1) {
                aup = ld_nc((const float *)auprow);                       // A[d0-1, t-1]
                adn = ld_nc((const float *)adnrow);                       // A[d0+K, t-1]
            }
#pragma unroll
            for (int k = 0; k < 5; k++) w[k] = ld_nc(at<const float>(grow, (unsigned)(k * (int)HW) * 4u));
        }

        float tc[K];
        if (t + 1 < H) {
            // everything the previous iteration (scan position t+1) published
            const float *eb = ex + ((t + 1) & 1) * bufsz + lane;
            const float up = (j > 0) ? eb[BX_THI * plane + (j - 1) * 32] : 0.f;
            const float dn = (j + 1 < NW) ? eb[BX_TLO * plane + (j + 1) * 32] : 0.f;
            float sum_tn = 0.f, amax = -INFINITY;
            int idx_cur = 0x7fffffff;
            for (int jj = 0; jj < NW; jj++) {
                sum_tn += eb[BX_ST * plane + jj * 32];
                const float v = eb[BX_AMAX * plane + jj * 32];
                const int vi = __float_as_int(eb[BX_AIDX * plane + jj * 32]);
                if (v > amax) { amax = v; idx_cur = vi; }          // strict >: first maximum
            }
            // guidance gradients of scan position t+1 are complete now: one warp per weight
            for (int k = j; k < 5; k += NW) {
                float tot;
                if (k == 4) {
                    tot = sum_tn * amax;                           // (:265-272)
                } else {
                    tot = 0.f;
                    for (int jj = 0; jj < NW; jj++) tot += eb[(BX_S0 + k) * plane + jj * 32];
                }
                if (wok) gg_next[k * (int)HW] = tot;
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
                tc[i] = (FULL || d < D) ? v : 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < K; i++) tc[i] = t0[i];
        }

        if (wok) {
#pragma unroll
            for (int i = 0; i < K; i++) {
                const int d = d0 + i;
                if (FULL || d < D) {
                    float v = tc[i] * w[0];
                    if (d == 0) v += tc[i] * w[2];
                    if (d == D - 1) v += tc[i] * w[3];
                    *at<float>(girow, offb[i]) = gold[i] + v;
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
                if ((FULL || d < D) && ap[i] > best) { best = ap[i]; bi = d; }
            }
        }
        {
            float *wb = ex + (t & 1) * bufsz + j * 32 + lane;
            wb[BX_TLO * plane] = tc[0];
            wb[BX_THI * plane] = tc[K - 1];
            wb[BX_ST * plane] = st;
            wb[BX_S0 * plane] = s0;
            wb[BX_S1 * plane] = s1;
            wb[BX_S2 * plane] = s2;
            wb[BX_S3 * plane] = s3;
            wb[BX_AMAX * plane] = best;
            wb[BX_AIDX * plane] = __int_as_float(bi);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < K; i++) Tn[i] = tc[i];
#pragma unroll
        for (int k = 0; k < 5; k++) wnx[k] = w[k];
        gg_next = ggrow;
        xrow -= psb; arow -= psb; gorow -= psb; girow -= psb; grow -= psb; auprow -= psb; adnrow -= psb;
        mrow -= ps;                                            // bytes: u8 tensor
        ggrow -= ps;
    }
    // scan position 0: only w0 receives a gradient (Appendix A.3 quirk)
    {
        const float *eb = ex + lane;
        for (int k = j; k < 5; k += NW) {
            float tot = 0.f;
            if (k == 0)
                for (int jj = 0; jj < NW; jj++) tot += eb[BX_S0 * plane + jj * 32];
            if (wok) gg_next[k * (int)HW] = tot;
        }
    }
}

}  // namespace ganet
