// The SGA recurrence for K consecutive depths held in registers, shared by every
// scan kernel.  See sga.cu for the rounding contract.
#pragma once
#include "common.cuh"

namespace ganet {

// First scan position: all five terms use the raw input (GANet_kernel.cu:99-119,
// else-branches); nvcc contracts the reference into five chained FFMAs from +0.
template <int K>
__device__ __forceinline__ void sga_first_step(const float (&x)[K], const float (&w)[5],
                                               float (&A)[K])
{
#pragma unroll
    for (int i = 0; i < K; i++) {
        float a = __fmaf_rn(x[i], w[0], 0.f);
        a = __fmaf_rn(x[i], w[1], a);
        a = __fmaf_rn(x[i], w[2], a);
        a = __fmaf_rn(x[i], w[3], a);
        A[i] = __fmaf_rn(x[i], w[4], a);
    }
}

// Later positions (GANet_kernel.cu:97-120).  P is the previous row of this chunk,
// `up` = P[d0-1], `dn` = P[d0+K] (ignored where out of range), pmax = max_d P[d].
// d0 must be even (K even everywhere), so parity(d) == parity(i).
template <int K, bool FULL = false>
__device__ __forceinline__ void sga_next_step(const float (&P)[K], const float (&x)[K],
                                              const float (&w)[5], float up, float dn,
                                              float pmax, int d0, int D, float (&A)[K])
{
    // FULL: every depth of the chunk is < D, so only the two chunk edges can miss a
    // neighbour (d = 0 and d = D-1); interior taps need no select at all.
    const float up_eff = (d0 >= 1) ? up : x[0];
    const float dn_eff = (d0 + K < D) ? dn : x[K - 1];
#pragma unroll
    for (int i = 0; i < K; i++) {
        const int d = d0 + i;
        float pm, s3;
        if (FULL) {
            pm = (i == 0) ? up_eff : P[i == 0 ? 0 : i - 1];
            s3 = (i == K - 1) ? dn_eff : P[i == K - 1 ? K - 1 : i + 1];
        } else {
            pm = (i == 0) ? up : P[i == 0 ? 0 : i - 1];
            const float pp = (i == K - 1) ? dn : P[i == K - 1 ? K - 1 : i + 1];
            s3 = (d + 1 < D) ? pp : x[i];
        }
        float a = __fmaf_rn(x[i], w[0], 0.f);
        a = __fmaf_rn(P[i], w[1], a);
        if (i & 1) {                                  // odd d: d-1 exists, fused
            a = __fmaf_rn(pm, w[2], a);
        } else {                                      // even d: select, then mul + add
            const float s2 = FULL ? pm : ((d >= 1) ? pm : x[i]);
            a = __fadd_rn(a, __fmul_rn(s2, w[2]));
        }
        a = __fadd_rn(a, __fmul_rn(s3, w[3]));
        A[i] = __fmaf_rn(pmax, w[4], a);
    }
}

template <int K>
__device__ __forceinline__ float chunk_max(const float (&A)[K], int d0, int D)
{
    float lm = -INFINITY;
#pragma unroll
    for (int i = 0; i < K; i++)
        if (d0 + i < D) lm = fmaxf(lm, A[i]);
    return lm;
}

}  // namespace ganet
