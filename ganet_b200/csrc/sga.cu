// SGA (semi-global guided aggregation) for sm_100a -- scan kernels and C ABI.
//
// Reference being replaced: libs/GANet/src/GANet_kernel.cu
//   sga_{down,up,right,left}_forward        :66-127, :285-346, :507-565, :720-778
//   Max / get_temp_grad / MaxDepth          :23-64
//   sga_*_data_backward, *_weight_backward  :129-281, :348-505, :567-718, :780-933
//   host sequences sga_kernel_forward/backward :935-1129
//
// Design (DESIGN.md has the full story).  The reference gives every scan LINE to
// one thread that walks H*D (or W*D) dependent global read-modify-writes.  Here a
// scan line belongs to a GROUP of L lanes of one warp: lane j of the group keeps
// the contiguous depth chunk d = K*j .. K*j+K-1 of the running row in registers,
// chunk-edge neighbours travel by warp shuffle, and the running max over depth
// -- the only cross-depth reduction of the forward recurrence, since
// P[argmax P] == max P -- is one `redux.sync.max.f32` (CREDUX) for L == 32.
// The four directions share one kernel: a direction is just (first pixel,
// pixel step) of the line.
//
// Rounding.  `out` and `mask` must be bit-identical to the reference CUDA
// build, so the forward step spells out the exact FMA-contraction pattern nvcc
// 12.9 chose for the reference on sm_100a (audited again from
// the SASS of the unmodified reference build):
//     first scan step : five chained fma(x, w_k, acc), acc = +0
//     later, even d   : fma, fma, mul+add, mul+add, fma
//     later, odd  d   : fma, fma, fma,     mul+add, fma
// K is always even, so the parity of d = K*j + i is the parity of the unrolled
// index i and the choice is made at compile time.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "sga_step.cuh"
#include "sga_vert.cuh"
#include "sga_tma.cuh"
#include "sga_hscan.cuh"
#include "transpose.cuh"

namespace ganet {

enum { MODE_FIRST = 0, MODE_COMBINE = 1, MODE_RAW = 2 };

constexpr int kWarpsPerBlock = 4;

struct LineGeom {
    int T;       // scan length
    int NL;      // number of lines per slice
    int lmul;    // first pixel = line * lmul + padd
    int padd;
    int pstep;   // pixel step per scan step
};

__host__ __device__ inline LineGeom line_geom(int dir, int H, int W)
{
    LineGeom q;
    switch (dir) {
    case 0: q.T = H; q.NL = W; q.lmul = 1; q.padd = 0; q.pstep = W; break;                 // down
    case 1: q.T = H; q.NL = W; q.lmul = 1; q.padd = (H - 1) * W; q.pstep = -W; break;      // up
    case 2: q.T = W; q.NL = H; q.lmul = W; q.padd = 0; q.pstep = 1; break;                 // right
    default: q.T = W; q.NL = H; q.lmul = W; q.padd = W - 1; q.pstep = -1; break;           // left
    }
    return q;
}

// ---------------------------------------------------------------------------
// forward scan of one direction
//   MODE_FIRST   : out = A, mask = 0           (sga_kernel_forward :962-968)
//   MODE_COMBINE : if (out < A) { out = A; mask = dir; }   (Max, :23-36)
//   MODE_RAW     : out = A                     (recompute for backward / debug)
// ---------------------------------------------------------------------------
template <int K, int L, int MODE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sga_scan_fwd_kernel(const float *__restrict__ x, const float *__restrict__ g, float *out,
                    uint8_t *mask, int dir, int D, int H, int W, long long n_slices,
                    int groups_per_slice)
{
    static_assert(K % 2 == 0, "depth parity must be a compile-time property");
    constexpr int G = 32 / L;
    const int lane = threadIdx.x & 31;
    const long long gwarp = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const long long s = gwarp / groups_per_slice;
    if (s >= n_slices) return;
    const int grp = (int)(gwarp - s * groups_per_slice);
    const int j = lane / G, gl = lane - j * G;
    const int HW = H * W;
    const LineGeom q = line_geom(dir, H, W);
    int line = grp * G + gl;
    const bool line_ok = line < q.NL;
    line = line_ok ? line : q.NL - 1;      // surplus lanes shadow the last line, never store

    const long long S = (long long)D * HW;
    const float *xs = x + s * S;
    const float *gs = g + s * 5ll * HW;
    float *os = out + s * S;
    uint8_t *ms = (MODE == MODE_RAW) ? nullptr : mask + s * S;

    const int d0 = K * j;
    int off[K];
#pragma unroll
    for (int i = 0; i < K; i++) off[i] = min(d0 + i, D - 1) * HW;   // clamped: loads stay in range

    int p = line * q.lmul + q.padd;
    float P[K], xc[K], w[5];
#pragma unroll
    for (int i = 0; i < K; i++) { xc[i] = ld_nc(xs + off[i] + p); P[i] = 0.f; }
#pragma unroll
    for (int k = 0; k < 5; k++) w[k] = ld_nc(gs + k * HW + p);
    float pmax = 0.f;

    for (int t = 0; t < q.T; t++) {
        // software prefetch of the next scan position
        const int pn = p + q.pstep;
        float xn[K], wn[5];
        if (t + 1 < q.T) {
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

        float A[K];
        if (t == 0) {
            sga_first_step<K>(xc, w, A);
        } else {
            const float up = from_prev_chunk<L>(P[K - 1]);   // P[d0 - 1]
            const float dn = from_next_chunk<L>(P[0]);       // P[d0 + K]
            sga_next_step<K>(P, xc, w, up, dn, pmax, d0, D, A);
        }
        pmax = group_max<L>(chunk_max<K>(A, d0, D));

        if (line_ok) {
#pragma unroll
            for (int i = 0; i < K; i++) {
                if (d0 + i < D) {
                    const int e = off[i] + p;
                    if (MODE == MODE_FIRST) {
                        os[e] = A[i];
                        ms[e] = 0;
                    } else if (MODE == MODE_COMBINE) {
                        if (os[e] < A[i]) { os[e] = A[i]; ms[e] = (uint8_t)dir; }
                    } else {
                        os[e] = A[i];
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < K; i++) { P[i] = A[i]; xc[i] = xn[i]; }
#pragma unroll
        for (int k = 0; k < 5; k++) w[k] = wn[k];
        p = pn;
    }
}

// ---------------------------------------------------------------------------
// backward of one direction: reverse scan that
// propagates T, emits gradInput and the five guidance gradients per pixel.
// `a` is the recomputed aggregate of this direction (MODE_RAW pass above).
// ---------------------------------------------------------------------------
template <int K, int L>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sga_scan_bwd_kernel(const float *__restrict__ x, const float *__restrict__ g,
                    const float *__restrict__ a, const uint8_t *__restrict__ mask,
                    const float *__restrict__ go, float *gi, float *__restrict__ gg,
                    int32_t *__restrict__ max_idx, int dir, int accumulate, int D, int H, int W,
                    long long n_slices, int groups_per_slice)
{
    static_assert(L >= 8, "five guidance gradients are written by lanes j = 0..4");
    constexpr int G = 32 / L;
    const int lane = threadIdx.x & 31;
    const long long gwarp = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const long long s = gwarp / groups_per_slice;
    if (s >= n_slices) return;
    const int grp = (int)(gwarp - s * groups_per_slice);
    const int j = lane / G, gl = lane - j * G;
    const int HW = H * W;
    const LineGeom q = line_geom(dir, H, W);
    int line = grp * G + gl;
    const bool line_ok = line < q.NL;
    line = line_ok ? line : q.NL - 1;

    const long long S = (long long)D * HW;
    const float *xs = x + s * S;
    const float *as = a + s * S;
    const float *gos = go + s * S;
    const uint8_t *ms = mask + s * S;
    float *gis = gi + s * S;
    const float *gs = g + s * 5ll * HW;
    float *ggs = gg + s * 5ll * HW;
    int32_t *mis = max_idx ? max_idx + s * (long long)HW : nullptr;

    const int d0 = K * j;
    int off[K];
#pragma unroll
    for (int i = 0; i < K; i++) off[i] = min(d0 + i, D - 1) * HW;

    // first arg-max over depth of one row held as K values per lane
    auto row_argmax = [&](const float (&r)[K], float &vmax) -> int {
        float best = (d0 < D) ? r[0] : -INFINITY;
        int bi = d0;
#pragma unroll
        for (int i = 1; i < K; i++)
            if (d0 + i < D && r[i] > best) { best = r[i]; bi = d0 + i; }
        vmax = group_max<L>(best);
        return group_min<L>(best == vmax ? bi : 0x7fffffff);
    };

    int p = line * q.lmul + q.padd + (q.T - 1) * q.pstep;   // last scan position

    if (mis) {   // the reference's MaxDepth covers every pixel, including the last row
        float r[K], vm;
#pragma unroll
        for (int i = 0; i < K; i++) r[i] = ld_nc(as + off[i] + p);
        const int k = row_argmax(r, vm);
        if (line_ok && j == 0) mis[p] = k;
    }

    float Tn[K], wnx[5];
#pragma unroll
    for (int i = 0; i < K; i++) Tn[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 5; k++) wnx[k] = 0.f;
    float sum_tn = 0.f;
    int idx_cur = -1;            // arg-max of A at the current position (set one step earlier)

    for (int t = q.T - 1; t >= 0; t--) {
        const int pq = p - q.pstep;      // previous scan position (t - 1)
        float xv[K], t0[K], ap[K], w[5];
#pragma unroll
        for (int i = 0; i < K; i++) {
            const int e = off[i] + p;
            xv[i] = ld_nc(xs + e);
            const float gv = ld_nc(gos + e);
            const uint8_t mv = ms[e];
            t0[i] = (d0 + i < D && mv == dir) ? gv : 0.f;       // get_temp_grad :38-48
            ap[i] = (t >= 1) ? ld_nc(as + off[i] + pq) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 5; k++) w[k] = ld_nc(gs + k * HW + p);

        float tc[K];
        if (t + 1 < q.T) {
            const float up = from_prev_chunk<L>(Tn[K - 1]);     // T[d0-1, t+1]
            const float dn = from_next_chunk<L>(Tn[0]);         // T[d0+K, t+1]
            const float inj = sum_tn * wnx[4];                  // max-path term (:167-178)
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

        // gradInput (:164, :177, :200-207)
        if (line_ok) {
#pragma unroll
            for (int i = 0; i < K; i++) {
                const int d = d0 + i;
                if (d < D) {
                    float v = tc[i] * w[0];
                    if (d == 0) v += tc[i] * w[2];
                    if (d == D - 1) v += tc[i] * w[3];
                    const int e = off[i] + p;
                    gis[e] = accumulate ? gis[e] + v : v;
                }
            }
        }

        // guidance gradients (:210-281)
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, st = 0.f, amax = 0.f;
        int idx_prev = -1;
        if (t >= 1) {
            const float aup = from_prev_chunk<L>(ap[K - 1]);    // A[d0-1, t-1]
            const float adn = from_next_chunk<L>(ap[0]);        // A[d0+K, t-1]
#pragma unroll
            for (int i = 0; i < K; i++) {
                const int d = d0 + i;
                const float am = (i == 0) ? aup : ap[i == 0 ? 0 : i - 1];
                const float apn = (i == K - 1) ? adn : ap[i == K - 1 ? K - 1 : i + 1];
                s0 += tc[i] * xv[i];
                st += tc[i];
                s1 += tc[i] * ap[i];
                s2 += tc[i] * ((d >= 1) ? am : xv[i]);
                s3 += tc[i] * ((d + 1 < D) ? apn : xv[i]);
            }
            idx_prev = row_argmax(ap, amax);
            if (mis && line_ok && j == 0) mis[pq] = idx_prev;
        } else {
#pragma unroll
            for (int i = 0; i < K; i++) { s0 += tc[i] * xv[i]; st += tc[i]; }
        }
        s0 = group_sum<L>(s0);
        st = group_sum<L>(st);
        if (t >= 1) {
            s1 = group_sum<L>(s1);
            s2 = group_sum<L>(s2);
            s3 = group_sum<L>(s3);
        }
        if (line_ok && j < 5) {
            const float s4 = st * amax;
            const float v = j == 0 ? s0 : (t >= 1 ? (j == 1 ? s1 : j == 2 ? s2 : j == 3 ? s3 : s4) : 0.f);
            ggs[j * HW + p] = v;
        }

#pragma unroll
        for (int i = 0; i < K; i++) Tn[i] = tc[i];
#pragma unroll
        for (int k = 0; k < 5; k++) wnx[k] = w[k];
        sum_tn = st;
        idx_cur = idx_prev;
        p = pq;
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct Cfg { int K, L; };

// compiled (K, L) pairs; K even.  L = 8 puts 4 adjacent lines in one warp,
// L = 32 one line per warp.
#define GANET_SGA_CFGS(X) \
    X(2, 8) X(4, 8) X(6, 8) X(10, 8) X(12, 8) X(24, 8) \
    X(2, 32) X(4, 32) X(6, 32) X(10, 32) X(16, 32) X(24, 32)

static const Cfg kCfgs[] = {
#define X(K_, L_) {K_, L_},
    GANET_SGA_CFGS(X)
#undef X
};

static bool pick_cfg(int D, bool vertical, Cfg *out)
{
    double best = -1;
    Cfg pick{0, 0};
    for (const Cfg &c : kCfgs) {
        if (c.K * c.L < D) continue;
        double eff = (double)D / (c.K * c.L);
        // vertical scans like 4 adjacent pixels per warp (16-byte segments);
        // horizontal scans like many independent warps
        if (vertical == (c.L == 8)) eff *= 1.2;
        if (eff > best) { best = eff; pick = c; }
    }
    if (best < 0) return false;
    *out = pick;
    return true;
}

template <int MODE>
static int launch_fwd(Cfg c, const float *x, const float *g, float *out, uint8_t *mask, int dir,
                      int D, int H, int W, long long n_slices, cudaStream_t st)
{
    const LineGeom q = line_geom(dir, H, W);
    const int G = 32 / c.L;
    const int gps = (q.NL + G - 1) / G;
    const long long warps = n_slices * gps;
    const long long blocks = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks <= 0) return GANET_OK;
    if (blocks > 0x7fffffffll) return GANET_EUNSUPPORTED;
#define X(K_, L_)                                                                           \
    if (c.K == K_ && c.L == L_) {                                                           \
        sga_scan_fwd_kernel<K_, L_, MODE><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>( \
            x, g, out, mask, dir, D, H, W, n_slices, gps);                                  \
    } else
    GANET_SGA_CFGS(X) { return GANET_EUNSUPPORTED; }
#undef X
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

static int launch_bwd(Cfg c, const float *x, const float *g, const float *a, const uint8_t *mask,
                      const float *go, float *gi, float *gg, int32_t *max_idx, int dir,
                      int accumulate, int D, int H, int W, long long n_slices, cudaStream_t st)
{
    const LineGeom q = line_geom(dir, H, W);
    const int G = 32 / c.L;
    const int gps = (q.NL + G - 1) / G;
    const long long warps = n_slices * gps;
    const long long blocks = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks <= 0) return GANET_OK;
    if (blocks > 0x7fffffffll) return GANET_EUNSUPPORTED;
#define X(K_, L_)                                                                        \
    if (c.K == K_ && c.L == L_) {                                                        \
        sga_scan_bwd_kernel<K_, L_><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(   \
            x, g, a, mask, go, gi, gg, max_idx, dir, accumulate, D, H, W, n_slices, gps); \
    } else
    GANET_SGA_CFGS(X) { return GANET_EUNSUPPORTED; }
#undef X
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

// ---- vertical (coalesced) kernels: (K, MAXW) instantiations -------------------------
#define GANET_VERT_CFGS(X) X(2, 16) X(4, 16) X(6, 16) X(6, 32) X(12, 16) X(24, 12)

struct VCfg { int K, NW; };

static bool pick_vert_cfg(int D, VCfg *out)
{
    static const int ks[] = {2, 4, 6, 12, 24};
    static const int mw[] = {16, 16, 16, 16, 12};
    static int forced_k = -1;                                   // tuning aid, read once
    if (forced_k < 0) { const char *env = getenv("GANET_VERT_K"); forced_k = env ? atoi(env) : 0; }
    if (forced_k > 0) {
        const int k = forced_k;
        for (int i = 0; i < 5; i++)
            if (ks[i] == k && (D + k - 1) / k <= (k == 6 ? 32 : mw[i])) { out->K = k; out->NW = (D + k - 1) / k; return true; }
    }
    for (int i = 0; i < 5; i++) {
        const int nw = (D + ks[i] - 1) / ks[i];
        if (nw <= mw[i]) { out->K = ks[i]; out->NW = nw; return true; }
    }
    return false;       // D > 288: the line kernels take over
}

template <int MODE>
static int launch_vert_fwd(VCfg c, const float *x, const float *g, float *out, uint8_t *mask,
                           int dir, MaskIds ids, int D, int H, int W, long long n_slices,
                           cudaStream_t st)
{
    const int strips = (W + 31) / 32;
    const long long blocks = n_slices * strips;
    if (blocks <= 0) return GANET_OK;
    if (blocks > 0x7fffffffll) return GANET_EUNSUPPORTED;
    const size_t smem = (size_t)2 * 3 * c.NW * 32 * sizeof(float);
    const bool full = c.K * c.NW == D;
#define X(K_, W_)                                                                            \
    if (c.K == K_ && c.NW <= W_) {                                                           \
        if (full)                                                                            \
            sga_vert_fwd_kernel<K_, W_, MODE, true><<<(unsigned)blocks, c.NW * 32, smem, st>>>( \
                x, g, out, mask, dir, ids, D, H, W, strips);                                 \
        else                                                                                 \
            sga_vert_fwd_kernel<K_, W_, MODE, false><<<(unsigned)blocks, c.NW * 32, smem, st>>>( \
                x, g, out, mask, dir, ids, D, H, W, strips);                                 \
    } else
    GANET_VERT_CFGS(X) { return GANET_EUNSUPPORTED; }
#undef X
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

static int launch_vert_bwd(VCfg c, const float *x, const float *g, const float *a,
                           const uint8_t *mask, const float *go, float *gi, float *gg, int dir,
                           int mask_id, int accumulate, int D, int H, int W, long long n_slices,
                           cudaStream_t st)
{
    const int strips = (W + 31) / 32;
    const long long blocks = n_slices * strips;
    if (blocks <= 0) return GANET_OK;
    if (blocks > 0x7fffffffll) return GANET_EUNSUPPORTED;
    const size_t smem = (size_t)2 * NBW * c.NW * 32 * sizeof(float);
    const bool full = c.K * c.NW == D;
#define X(K_, W_)                                                                            \
    if (c.K == K_ && c.NW <= W_) {                                                           \
        if (smem > 48 * 1024) {                                                              \
            cudaFuncSetAttribute(sga_vert_bwd_kernel<K_, W_, true>,                          \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
            cudaFuncSetAttribute(sga_vert_bwd_kernel<K_, W_, false>,                         \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        }                                                                                    \
        if (full)                                                                            \
            sga_vert_bwd_kernel<K_, W_, true><<<(unsigned)blocks, c.NW * 32, smem, st>>>(    \
                x, g, a, mask, go, gi, gg, dir, mask_id, accumulate, D, H, W, strips);       \
        else                                                                                 \
            sga_vert_bwd_kernel<K_, W_, false><<<(unsigned)blocks, c.NW * 32, smem, st>>>(   \
                x, g, a, mask, go, gi, gg, dir, mask_id, accumulate, D, H, W, strips);       \
    } else
    GANET_VERT_CFGS(X) { return GANET_EUNSUPPORTED; }
#undef X
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

// ---- TMA-staged variants of the vertical kernels -----------------------------------------
constexpr int kNotApplicable = -100;          // internal: fall back to the LDG kernels
constexpr int kSmemBudget = 227 * 1024 - 2048;

static bool tma_enabled()
{
    static int v = -1;
    if (v < 0) v = getenv("GANET_NO_TMA") ? 0 : 1;
    return v != 0;
}

// tuning aid: cap the depth of the TMA stage ring (fewer stages = less shared memory per CTA =
// more resident CTAs per SM)
static int stage_cap(int dflt)
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("GANET_TMA_STAGES"); v = e ? atoi(e) : 0; }
    return (v >= 2 && v < dflt) ? v : dflt;
}

// L2 evict-first hints on the TMA traffic.  Measured on B200 at 1x32x192x240x624: WORSE
// (forward 12.6 -> 15.1 ms, backward 28.1 -> 29.2 ms) -- neighbouring strips and consecutive
// passes do profit from L2 -- so the hint is off unless GANET_L2_HINT=1 asks for it.
static int stream_hint(long long n_slices, int D, int H, int W)
{
    (void)n_slices; (void)D; (void)H; (void)W;
    static int forced = -2;
    if (forced == -2) { const char *e = getenv("GANET_L2_HINT"); forced = e ? atoi(e) : 0; }
    return forced > 0;
}

// gradInput accumulation by TMA reduce-add instead of load + add + store.  Same bits
// (test_tma_and_ldg_kernels_agree_bitwise); measured on B200 at 2x32x192x240x624: backward
// 38.3-39.3 -> 36.1-36.2 ms with kept aggregates, 54.1 -> 51.1 ms without.  GANET_TMA_REDUCE=0
// restores the load + add + store form.
// L2 prefetch distance (rows beyond the shared-memory ring) of the vertical TMA kernels
static int vert_prefetch()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("GANET_VERT_PREFETCH"); v = e ? atoi(e) : 0; }
    return v;
}

static bool tma_reduce_enabled()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("GANET_TMA_REDUCE"); v = e ? atoi(e) : 1; }
    return v != 0;
}

template <int MODE>
static int launch_tma_fwd(VCfg c, const float *x, const float *g, float *out, uint8_t *mask, int dir,
                          MaskIds ids, int D, int H, int W, long long n_slices, cudaStream_t st,
                          const float *a2 = nullptr, const float *a3 = nullptr)
{
    constexpr bool kThree = (MODE == VMODE_FIRST3);
    constexpr bool kCombine = (MODE == VMODE_SECOND || MODE == VMODE_COMBINE || kThree);
    if (!tma_enabled() || D > 256 || (W % 4) != 0 || (kCombine && (W % 16) != 0)) return kNotApplicable;
    const FwdPlan pl = fwd_plan(D, kCombine, kThree);
    const int ex_bytes = 2 * 3 * c.NW * 32 * 4;
    int S = (kSmemBudget - ex_bytes - 128) / pl.stage_bytes;
    if (S > stage_cap(6)) S = stage_cap(6);
    if (S > H) S = H;
    if (S < 2 && H >= 2) return kNotApplicable;
    const size_t smem = (size_t)S * pl.stage_bytes + ex_bytes + 2 * S * sizeof(uint64_t);
    TmaFwdMaps maps;
    if (!make_plane_map(&maps.x, x, 4, n_slices * D, H, W, 32, D)) return kNotApplicable;
    if (!make_plane_map(&maps.g, g, 4, n_slices * 5, H, W, 32, 5)) return kNotApplicable;
    if (!make_plane_map(&maps.out, out, 4, n_slices * D, H, W, 32, D)) return kNotApplicable;
    if (kCombine) {
        if (!make_plane_map(&maps.mask, mask, 1, n_slices * D, H, W, 32, D)) return kNotApplicable;
    } else {
        maps.mask = maps.out;
    }
    maps.a2 = maps.out;
    maps.a3 = maps.out;
    if (kThree) {
        if (!a2 || !a3) return GANET_EINVAL;
        if (!make_plane_map(&maps.a2, a2, 4, n_slices * D, H, W, 32, D)) return kNotApplicable;
        if (!make_plane_map(&maps.a3, a3, 4, n_slices * D, H, W, 32, D)) return kNotApplicable;
    }
    const int strips = (W + 31) / 32;
    const long long blocks = n_slices * strips;
    if (blocks <= 0) return GANET_OK;
    if (blocks > 0x7fffffffll) return GANET_EUNSUPPORTED;
    const bool full = c.K * c.NW == D;
#define X(K_, W_)                                                                              \
    if (c.K == K_ && c.NW <= W_) {                                                             \
        auto kf = sga_tma_fwd_kernel<K_, W_, MODE, true>;                                      \
        auto kp = sga_tma_fwd_kernel<K_, W_, MODE, false>;                                     \
        auto k = full ? kf : kp;                                                               \
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != \
            cudaSuccess) { cudaGetLastError(); return kNotApplicable; }                        \
        k<<<(unsigned)blocks, (c.NW + 1) * 32, smem, st>>>(maps, dir, ids, D, H, strips, S,    \
                                                           stream_hint(n_slices, D, H, W), vert_prefetch()); \
    } else
    GANET_VERT_CFGS(X) { return kNotApplicable; }
#undef X
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

static int launch_tma_bwd(VCfg c, const float *x, const float *g, const float *a, const uint8_t *mask,
                          const float *go, float *gi, float *gg, int dir, int mask_id, int accumulate,
                          int D, int H, int W, long long n_slices, cudaStream_t st)
{
    if (!tma_enabled() || D > 256 || (W % 16) != 0) return kNotApplicable;
    const BwdPlan pl = bwd_plan(D);
    const int ex_bytes = 2 * NBW * c.NW * 32 * 4;
    int S = (kSmemBudget - ex_bytes - 128) / pl.stage_bytes;
    if (S > 4) S = 4;
    if (S > H) S = H;
    if (S < 2 && H >= 2) return kNotApplicable;
    const size_t smem = (size_t)S * pl.stage_bytes + ex_bytes + 2 * S * sizeof(uint64_t);
    TmaBwdMaps maps;
    if (!make_plane_map(&maps.x, x, 4, n_slices * D, H, W, 32, D)) return kNotApplicable;
    if (!make_plane_map(&maps.g, g, 4, n_slices * 5, H, W, 32, 5)) return kNotApplicable;
    if (!make_plane_map(&maps.a, a, 4, n_slices * D, H, W, 32, D)) return kNotApplicable;
    if (!make_plane_map(&maps.go, go, 4, n_slices * D, H, W, 32, D)) return kNotApplicable;
    if (!make_plane_map(&maps.gi, gi, 4, n_slices * D, H, W, 32, D)) return kNotApplicable;
    if (!make_plane_map(&maps.mask, mask, 1, n_slices * D, H, W, 32, D)) return kNotApplicable;
    const int strips = (W + 31) / 32;
    const long long blocks = n_slices * strips;
    if (blocks <= 0) return GANET_OK;
    if (blocks > 0x7fffffffll) return GANET_EUNSUPPORTED;
    const bool full = c.K * c.NW == D;
    const int acc_mode = accumulate ? (tma_reduce_enabled() ? 2 : 1) : 0;
#define X(K_, W_)                                                                              \
    if (c.K == K_ && c.NW <= W_) {                                                             \
        auto kf = sga_tma_bwd_kernel<K_, W_, true>;                                            \
        auto kp = sga_tma_bwd_kernel<K_, W_, false>;                                           \
        auto k = full ? kf : kp;                                                               \
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != \
            cudaSuccess) { cudaGetLastError(); return kNotApplicable; }                        \
        k<<<(unsigned)blocks, (c.NW + 1) * 32, smem, st>>>(maps, gi, gg, dir, mask_id,         \
                                                           acc_mode, D, H, W, strips, S,       \
                                                           stream_hint(n_slices, D, H, W), vert_prefetch()); \
    } else
    GANET_VERT_CFGS(X) { return kNotApplicable; }
#undef X
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

// horizontal aggregate in the standard layout (no transposes), DIR 0 = right, 1 = left
template <int DIR>
static int launch_tma_hraw(VCfg c, const float *x, const float *g, float *out, int D, int H, int W,
                           long long n_slices, cudaStream_t st)
{
    if (!tma_enabled() || D > 256 || (W % 4) != 0 || H < 32) return kNotApplicable;
    const int stage_bytes = D * 512 + 2560;
    const int ex_bytes = 2 * 3 * c.NW * 32 * 4;
    int S = (kSmemBudget - ex_bytes - 128) / stage_bytes;
    if (S > 4) S = 4;
    if (S > W / 4) S = W / 4;
    if (S < 2) return kNotApplicable;
    const size_t smem = (size_t)S * stage_bytes + ex_bytes + 2 * S * sizeof(uint64_t);
    TmaHrawMaps maps;
    if (!make_plane_map(&maps.x, x, 4, n_slices * D, H, W, 4, D, 32)) return kNotApplicable;
    if (!make_plane_map(&maps.g, g, 4, n_slices * 5, H, W, 4, 5, 32)) return kNotApplicable;
    if (!make_plane_map(&maps.out, out, 4, n_slices * D, H, W, 4, D, 32)) return kNotApplicable;
    const int strips = (H + 31) / 32;
    const long long blocks = n_slices * strips;
    if (blocks <= 0) return GANET_OK;
    if (blocks > 0x7fffffffll) return GANET_EUNSUPPORTED;
    const bool full = c.K * c.NW == D;
#define X(K_, W_)                                                                              \
    if (c.K == K_ && c.NW <= W_) {                                                             \
        auto kf = sga_tma_hraw_kernel<K_, W_, DIR, true>;                                      \
        auto kp = sga_tma_hraw_kernel<K_, W_, DIR, false>;                                     \
        auto k = full ? kf : kp;                                                               \
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != \
            cudaSuccess) { cudaGetLastError(); return kNotApplicable; }                        \
        k<<<(unsigned)blocks, (c.NW + 1) * 32, smem, st>>>(maps, D, W, strips, S);             \
    } else
    GANET_VERT_CFGS(X) { return kNotApplicable; }
#undef X
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}


// ---- horizontal scans in the standard layout (sga_hscan.cuh) ---------------------------------
// K = depths per lane (even), 32 lanes cover D <= 32*K; the TMA box holds all D planes (<= 256).
#define GANET_HSCAN_KS(X) X(2) X(4) X(6) X(8)

static bool hscan_enabled()
{
    static int v = -1;
    if (v < 0) v = getenv("GANET_NO_HSCAN") ? 0 : 1;
    return v != 0 && tma_enabled();
}

// L2 prefetch distance (tiles beyond the shared-memory ring) of the horizontal kernels
static int hscan_prefetch(bool backward)
{
    static int vf = -1, vb = -1;
    // measured on B200 at the headline size: every distance > 0 is SLOWER (right/left raw 1.46 ->
    // 1.67 ms at 4 tiles, backward 17.7 -> 18.1 ms; the vertical kernels likewise): a prefetch is one
    // more request per row and the memory system is not short of requests.  Off unless asked for.
    if (vf < 0) { const char *e = getenv("GANET_HSCAN_PREFETCH"); vf = e ? atoi(e) : 0; }
    if (vb < 0) { const char *e = getenv("GANET_HSCAN_BWD_PREFETCH"); vb = e ? atoi(e) : 0; }
    return backward ? vb : vf;
}

static int hscan_k(int D)
{
    for (int k = 2; k <= 8; k += 2)
        if (32 * k >= D) return k;
    return 0;
}

// Shapes the horizontal kernels take: every tensor row 16-byte aligned for fp32 AND for the uint8
// mask (W % 16 == 0), all planes of a slice in one TMA box (D <= 256).  A pure function of the
// shape (plus the process-wide switches), so forward and backward agree on the layout of the kept
// aggregates without passing a flag.
static bool hscan_ok(int D, int H, int W)
{
    (void)H;
    return hscan_enabled() && D <= 256 && (W % 16) == 0 && get_encode_tiled() != nullptr;
}

// A row's recurrence is a latency chain, so the horizontal kernels want as many rows (CTAs) per SM as
// shared memory allows: two stages per CTA are enough (ring depths 2-4 measured equal at D = 192), and a
// CTA is kept under a quarter of the SM's shared memory wherever two stages fit in that (D <= 128:
// D = 96 ran at 24 instead of 38 Gvoxel/s with four-stage rings and two rows per SM).
constexpr int kHscanCtaBudget = 56 * 1024;

// tuning aids: cap the stage rings of the horizontal kernels (fewer stages = more CTAs per SM)
static int hscan_stage_cap(bool backward, int dflt)
{
    static int vf = -1, vb = -1;
    if (vf < 0) { const char *e = getenv("GANET_HSCAN_STAGES"); vf = e ? atoi(e) : 0; }
    if (vb < 0) { const char *e = getenv("GANET_HSCAN_BWD_STAGES"); vb = e ? atoi(e) : 0; }
    const int v = backward ? vb : vf;
    return v >= 1 ? v : dflt;
}

template <int DIR>
static int launch_hscan_fwd(const float *x, const float *g, float *out, int D, int H, int W,
                            long long n_slices, cudaStream_t st)
{
    constexpr int BW = 32;
    const int K = hscan_k(D);
    if (!K || D > 256 || (W % 4) != 0) return kNotApplicable;
    const int stage = hfwd_stage_bytes(D, BW);
    const int nb = (W + BW - 1) / BW;
    int S = (kHscanCtaBudget - 2048) / stage;
    if (S < 2) S = 2;
    if (S > hscan_stage_cap(false, 4)) S = hscan_stage_cap(false, 4);
    if (S > nb) S = nb;
    if (S < 2 && nb >= 2) return kNotApplicable;
    if (S < 1) S = 1;
    const size_t smem = (size_t)S * stage + 2 * S * sizeof(uint64_t) + 1024;      // + alignment slack
    HFwdMaps maps;
    if (!make_plane_map(&maps.x, x, 4, n_slices * D, H, W, BW, D, 1, CU_TENSOR_MAP_SWIZZLE_128B)) return kNotApplicable;
    if (!make_plane_map(&maps.out, out, 4, n_slices * D, H, W, BW, D, 1, CU_TENSOR_MAP_SWIZZLE_128B)) return kNotApplicable;
    if (!make_plane_map(&maps.g, g, 4, n_slices * 5, H, W, BW, 5, 1)) return kNotApplicable;
    const long long blocks = n_slices * H;
    if (blocks <= 0) return GANET_OK;
    if (blocks > 0x7fffffffll) return GANET_EUNSUPPORTED;
    const bool full = (32 * K == D);
#define X(K_)                                                                                  \
    if (K == K_) {                                                                             \
        auto kf = sga_hscan_fwd_kernel<K_, BW, DIR, true>;                                     \
        auto kp = sga_hscan_fwd_kernel<K_, BW, DIR, false>;                                    \
        auto k = full ? kf : kp;                                                               \
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != \
            cudaSuccess) { cudaGetLastError(); return kNotApplicable; }                        \
        k<<<(unsigned)blocks, 64, smem, st>>>(maps, D, H, W, S, hscan_prefetch(false));                               \
    } else
    GANET_HSCAN_KS(X) { return kNotApplicable; }
#undef X
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

template <int DIR>
static int launch_hscan_bwd(const float *x, const float *g, const float *a, const uint8_t *mask,
                            const float *go, float *gi, float *gg, int32_t *max_idx, int mask_id,
                            int accumulate, int D, int H, int W, long long n_slices, cudaStream_t st)
{
    constexpr int BW = 16;
    const int K = hscan_k(D);
    if (!K || D > 256 || (W % 16) != 0) return kNotApplicable;
    const HBwdPlan pl = hbwd_plan(D, BW);
    const int nb = (W + BW - 1) / BW;
    int S = (kHscanCtaBudget - 2048) / pl.stage_bytes;
    if (S < 2) S = 2;
    if (S > hscan_stage_cap(true, 4)) S = hscan_stage_cap(true, 4);
    if (S > nb) S = nb;
    if (S < 2 && nb >= 2) return kNotApplicable;
    if (S < 1) S = 1;
    const size_t smem = (size_t)S * pl.stage_bytes + 2 * S * sizeof(uint64_t) + 64 + 1024;
    HBwdMaps maps;
    const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_64B;
    if (!make_plane_map(&maps.x, x, 4, n_slices * D, H, W, BW, D, 1, sw)) return kNotApplicable;
    if (!make_plane_map(&maps.go, go, 4, n_slices * D, H, W, BW, D, 1, sw)) return kNotApplicable;
    if (!make_plane_map(&maps.a, a, 4, n_slices * D, H, W, BW, D, 1, sw)) return kNotApplicable;
    if (!make_plane_map(&maps.gi, gi, 4, n_slices * D, H, W, BW, D, 1, sw)) return kNotApplicable;
    if (!make_plane_map(&maps.mask, mask, 1, n_slices * D, H, W, BW, D, 1)) return kNotApplicable;
    if (!make_plane_map(&maps.g, g, 4, n_slices * 5, H, W, BW, 5, 1)) return kNotApplicable;
    const long long blocks = n_slices * H;
    if (blocks <= 0) return GANET_OK;
    if (blocks > 0x7fffffffll) return GANET_EUNSUPPORTED;
    const bool full = (32 * K == D);
#define X(K_)                                                                                  \
    if (K == K_) {                                                                             \
        auto kf = sga_hscan_bwd_kernel<K_, BW, DIR, true>;                                     \
        auto kp = sga_hscan_bwd_kernel<K_, BW, DIR, false>;                                    \
        auto k = full ? kf : kp;                                                               \
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != \
            cudaSuccess) { cudaGetLastError(); return kNotApplicable; }                        \
        k<<<(unsigned)blocks, 96, smem, st>>>(maps, gg, max_idx, mask_id, accumulate, D, H, W, S, hscan_prefetch(true)); \
    } else
    GANET_HSCAN_KS(X) { return kNotApplicable; }
#undef X
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

static int launch_merge4(const float *a0, const float *a1, const float *a2, const float *a3, float *out,
                         uint8_t *mask, long long total, cudaStream_t st)
{
    if (total % 4) return GANET_EUNSUPPORTED;
    const long long n4 = total / 4;
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks <= 0) return GANET_OK;
    merge4_kernel<<<(unsigned)blocks, 256, 0, st>>>((const float4 *)a0, (const float4 *)a1, (const float4 *)a2,
                                                    (const float4 *)a3, (float4 *)out, (uint32_t *)mask, n4);
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

// Should the transpose-free forward (hraw x2 + FIRST3 + COMBINE, 4 launches instead of 9) run?
// Its boxes have a 16-byte inner extent: 6144 tiny TMA requests per box, which is request-rate
// bound once the volume no longer sits in L2.  Measured on B200 (profiles/): faster up to
// ~35M voxels per call (1x32x65x80x208: 0.72 vs 0.85 ms; 1x48x33x64x208: 0.38 vs 0.52 ms),
// slower beyond (1x32x192x240x624: 33.8 vs 12.8 ms) -- so it is used for small calls only.
static bool fwd_direct_ok(VCfg c, int D, int H, int W, long long voxels)
{
    static int no_direct = -1, force_direct = -1;               // read once
    if (no_direct < 0) no_direct = getenv("GANET_NO_DIRECT") ? 1 : 0;
    if (force_direct < 0) force_direct = getenv("GANET_FORCE_DIRECT") ? 1 : 0;
    if (!tma_enabled() || no_direct || D > 256 || (W % 16) != 0 || H < 32) return false;
    if (voxels > 48ll * 1000 * 1000 && !force_direct) return false;
    const int ex_bytes = 2 * 3 * c.NW * 32 * 4;
    if ((kSmemBudget - ex_bytes - 128) / (D * 512 + 2560) < 2) return false;
    if ((kSmemBudget - ex_bytes - 128) / fwd_plan(D, true, true).stage_bytes < 2) return false;
    return get_encode_tiled() != nullptr;
}

// front doors: TMA when the shape allows it, else the LDG kernels
template <int MODE>
static int run_vert_fwd(VCfg c, const float *x, const float *g, float *out, uint8_t *mask, int dir,
                        MaskIds ids, int D, int H, int W, long long n_slices, cudaStream_t st)
{
    const int rc = launch_tma_fwd<MODE>(c, x, g, out, mask, dir, ids, D, H, W, n_slices, st);
    if (rc != kNotApplicable) return rc;
    return launch_vert_fwd<MODE>(c, x, g, out, mask, dir, ids, D, H, W, n_slices, st);
}

static int run_vert_bwd(VCfg c, const float *x, const float *g, const float *a, const uint8_t *mask,
                        const float *go, float *gi, float *gg, int dir, int mask_id, int accumulate,
                        int D, int H, int W, long long n_slices, cudaStream_t st)
{
    const int rc = launch_tma_bwd(c, x, g, a, mask, go, gi, gg, dir, mask_id, accumulate, D, H, W,
                                  n_slices, st);
    if (rc != kNotApplicable) return rc;
    return launch_vert_bwd(c, x, g, a, mask, go, gi, gg, dir, mask_id, accumulate, D, H, W, n_slices, st);
}

static int check_dims(int64_t N, int64_t C, int64_t D, int64_t H, int64_t W)
{
    if (N <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return GANET_EINVAL;
    if (D > 768) return GANET_EUNSUPPORTED;
    if (D * H * W >= (1ll << 31)) return GANET_EUNSUPPORTED;   // one slice is indexed with int
    return GANET_OK;
}

}  // namespace ganet

using namespace ganet;

// ---- workspace carving ------------------------------------------------------------
static inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

struct FwdWs { size_t xT, outT, maskT, t3, gT2, gT3, total; };
static FwdWs fwd_ws(long long n, long long S, long long HW, bool keep = false, bool hs = false)
{
    FwdWs w; size_t o = 0;
    if (hs) {       // horizontal scans in the standard layout: two aggregates of scratch unless kept
        w.xT = o;    o += keep ? 0 : align_up((size_t)n * S * 4);
        w.outT = o;  o += keep ? 0 : align_up((size_t)n * S * 4);
        w.maskT = w.t3 = w.gT2 = w.gT3 = o;
        w.total = o ? o : 256;
        return w;
    }
    // keep: xT and the aggregates live in the caller's buffer; only the guidance is staged here
    w.xT = o;    o += keep ? 0 : align_up((size_t)n * S * 4);
    w.outT = o;  o += keep ? 0 : align_up((size_t)n * S * 4);
    w.maskT = o; o += keep ? 0 : align_up((size_t)n * S);
    w.t3 = o;
    w.gT2 = o;   o += align_up((size_t)n * 5 * HW * 4);
    w.gT3 = o;   o += align_up((size_t)n * 5 * HW * 4);
    w.total = o;
    return w;
}

struct BwdWs { size_t a, xT, goT, maskT, giT, gT, ggT, total; };
static BwdWs bwd_ws(long long n, long long S, long long HW, bool kept = false, bool hs = false)
{
    BwdWs w; size_t o = 0;
    if (hs) {       // only the recompute variant needs scratch: one aggregate at a time
        w.a = o;     o += kept ? 0 : align_up((size_t)n * S * 4);
        w.xT = w.goT = w.maskT = w.giT = w.gT = w.ggT = o;
        w.total = o ? o : 256;
        return w;
    }
    w.a = o;     o += kept ? 0 : align_up((size_t)n * S * 4);    // kept aggregates: no recompute scratch
    w.xT = o;    o += kept ? 0 : align_up((size_t)n * S * 4);    // ... and xT is kept as well
    w.goT = o;   o += align_up((size_t)n * S * 4);
    w.maskT = o; o += align_up((size_t)n * S);
    w.giT = o;   o += align_up((size_t)n * S * 4);
    w.gT = o;    o += align_up((size_t)n * 5 * HW * 4);
    w.ggT = o;   o += align_up((size_t)n * 5 * HW * 4);
    w.total = o;
    return w;
}

// largest slice count whose workspace fits (0 if not even one)
template <class F>
static long long fit_slices(F sizer, size_t bytes, long long ns)
{
    if (sizer(1) > bytes) return 0;
    long long lo = 1, hi = ns;
    while (lo < hi) {
        const long long mid = (lo + hi + 1) / 2;
        if (sizer(mid) <= bytes) lo = mid; else hi = mid - 1;
    }
    return lo;
}

GANET_API size_t ganet_sga_forward_workspace_min(int64_t N, int64_t C, int64_t D, int64_t H, int64_t W)
{
    (void)N; (void)C;
    const bool hs = hscan_ok((int)D, (int)H, (int)W);
    const size_t a = fwd_ws(1, D * H * W, H * W, false, hs).total, b = fwd_ws(1, D * H * W, H * W, true, hs).total;
    return a > b ? a : b;          // one size serves both forward variants (the plain one is larger)
}
GANET_API size_t ganet_sga_forward_workspace_best(int64_t N, int64_t C, int64_t D, int64_t H, int64_t W)
{
    const bool hs = hscan_ok((int)D, (int)H, (int)W);
    const size_t a = fwd_ws(N * C, D * H * W, H * W, false, hs).total, b = fwd_ws(N * C, D * H * W, H * W, true, hs).total;
    return a > b ? a : b;
}
GANET_API size_t ganet_sga_backward_workspace_min(int64_t N, int64_t C, int64_t D, int64_t H, int64_t W)
{
    (void)N; (void)C;
    return bwd_ws(1, D * H * W, H * W, false, hscan_ok((int)D, (int)H, (int)W)).total;
}
GANET_API size_t ganet_sga_backward_workspace_best(int64_t N, int64_t C, int64_t D, int64_t H, int64_t W)
{
    return bwd_ws(N * C, D * H * W, H * W, false, hscan_ok((int)D, (int)H, (int)W)).total;
}

/* How many (N,C,D,H,W) volumes the `aggregates` buffer of ganet_sga_forward / _backward holds for
 * this shape: 4 (the four aggregates, standard layout) when the horizontal scans run in the standard
 * layout, 5 (down, up, right^T, left^T, x^T) on the transposed path. */
GANET_API int ganet_sga_aggregate_volumes(int64_t N, int64_t C, int64_t D, int64_t H, int64_t W)
{
    (void)N; (void)C;
    return hscan_ok((int)D, (int)H, (int)W) ? 4 : 5;
}

// slow generic path: one warp group per scan line, any D <= 768, no workspace
static int sga_forward_lines(const float *x, const float *const g[4], float *out, uint8_t *mask,
                             int D, int H, int W, long long ns, cudaStream_t st)
{
    Cfg cv, ch;
    if (!pick_cfg(D, true, &cv) || !pick_cfg(D, false, &ch)) return GANET_EUNSUPPORTED;
    int rc = launch_fwd<MODE_FIRST>(cv, x, g[0], out, mask, 0, D, H, W, ns, st);
    if (rc) return rc;
    rc = launch_fwd<MODE_COMBINE>(cv, x, g[1], out, mask, 1, D, H, W, ns, st);
    if (rc) return rc;
    rc = launch_fwd<MODE_COMBINE>(ch, x, g[2], out, mask, 2, D, H, W, ns, st);
    if (rc) return rc;
    return launch_fwd<MODE_COMBINE>(ch, x, g[3], out, mask, 3, D, H, W, ns, st);
}

GANET_API int ganet_sga_forward(const float *x, const float *g_down, const float *g_up,
                                const float *g_right, const float *g_left, float *out,
                                uint8_t *mask, float *aggregates, void *workspace,
                                size_t workspace_bytes, int64_t N, int64_t C, int64_t D, int64_t H,
                                int64_t W, ganet_stream_t stream)
{
    if (!x || !g_down || !g_up || !g_right || !g_left || !out || !mask) return GANET_EINVAL;
    int rc = check_dims(N, C, D, H, W);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long S = D * H * W, HW = H * W, ns = N * C;
    const float *g[4] = {g_down, g_up, g_right, g_left};
    VCfg vc;
    if (!pick_vert_cfg((int)D, &vc)) {
        if (aggregates) return GANET_EUNSUPPORTED;       // D > 288: the generic path keeps nothing
        return sga_forward_lines(x, g, out, mask, (int)D, (int)H, (int)W, ns, st);
    }
    const bool keep = aggregates != nullptr;
    const int iD = (int)D, iH = (int)H, iW = (int)W;
    const bool hs = hscan_ok(iD, iH, iW);
    const long long chunk = workspace
        ? fit_slices([&](long long n) { return fwd_ws(n, S, HW, keep, hs).total; }, workspace_bytes, ns) : 0;
    if (chunk < 1) return GANET_EWORKSPACE;
    char *ws = (char *)workspace;
    for (long long s0 = 0; s0 < ns; s0 += chunk) {
        const long long n = (ns - s0 < chunk) ? ns - s0 : chunk;
        const FwdWs w = fwd_ws(n, S, HW, keep, hs);
        if (hs) {
            // Horizontal scans in the standard layout (sga_hscan.cuh): no transposes anywhere.
            const float *xs = x + s0 * S;
            float *os = out + s0 * S;
            uint8_t *ms = mask + s0 * S;
            if (keep) {
                // the four raw aggregates go to the caller's buffer, one streaming kernel merges them
                float *A[4];
                for (int k = 0; k < 4; k++) A[k] = aggregates + k * ns * S + s0 * S;
                if ((rc = run_vert_fwd<VMODE_RAW>(vc, xs, g_down + s0 * 5 * HW, A[0], nullptr, 0, MaskIds{0, 0}, iD, iH, iW, n, st))) return rc;
                if ((rc = run_vert_fwd<VMODE_RAW>(vc, xs, g_up + s0 * 5 * HW, A[1], nullptr, 1, MaskIds{0, 0}, iD, iH, iW, n, st))) return rc;
                if ((rc = launch_hscan_fwd<0>(xs, g_right + s0 * 5 * HW, A[2], iD, iH, iW, n, st))) return rc == kNotApplicable ? GANET_EUNSUPPORTED : rc;
                if ((rc = launch_hscan_fwd<1>(xs, g_left + s0 * 5 * HW, A[3], iD, iH, iW, n, st))) return rc == kNotApplicable ? GANET_EUNSUPPORTED : rc;
                if ((rc = launch_merge4(A[0], A[1], A[2], A[3], os, ms, n * S, st))) return rc;
                continue;
            }
            // recompute variant: right and left aggregates into scratch, `down` merges all three
            // (strict <, lower id wins), `up` merges on top
            float *a2 = (float *)(ws + w.xT), *a3 = (float *)(ws + w.outT);
            if ((rc = launch_hscan_fwd<0>(xs, g_right + s0 * 5 * HW, a2, iD, iH, iW, n, st))) return rc == kNotApplicable ? GANET_EUNSUPPORTED : rc;
            if ((rc = launch_hscan_fwd<1>(xs, g_left + s0 * 5 * HW, a3, iD, iH, iW, n, st))) return rc == kNotApplicable ? GANET_EUNSUPPORTED : rc;
            rc = launch_tma_fwd<VMODE_FIRST3>(vc, xs, g_down + s0 * 5 * HW, os, ms, 0, MaskIds{0, 0}, iD, iH, iW, n, st, a2, a3);
            if (rc) return rc == kNotApplicable ? GANET_EUNSUPPORTED : rc;   // D <= 256 always fits two stages
            if ((rc = run_vert_fwd<VMODE_COMBINE>(vc, xs, g_up + s0 * 5 * HW, os, ms, 1, MaskIds{0, 1}, iD, iH, iW, n, st))) return rc;
            continue;
        }
        if (keep) {
            // Memory-for-bandwidth variant: every direction's raw aggregate goes to the caller's
            // buffer (down, up in the standard layout; right, left transposed -- the layouts
            // backward consumes), one streaming kernel merges the four.  Backward then skips
            // its four recompute passes.
            float *gT2 = (float *)(ws + w.gT2), *gT3 = (float *)(ws + w.gT3);
            const float *xs = x + s0 * S;
            float *A0 = aggregates + 0 * ns * S + s0 * S, *A1 = aggregates + 1 * ns * S + s0 * S;
            float *A2T = aggregates + 2 * ns * S + s0 * S, *A3T = aggregates + 3 * ns * S + s0 * S;
            float *xT = aggregates + 4 * ns * S + s0 * S;      // kept too: backward skips its T(x)
            if ((rc = run_vert_fwd<VMODE_RAW>(vc, xs, g_down + s0 * 5 * HW, A0, nullptr, 0, MaskIds{0, 0}, iD, iH, iW, n, st))) return rc;
            if ((rc = run_vert_fwd<VMODE_RAW>(vc, xs, g_up + s0 * 5 * HW, A1, nullptr, 1, MaskIds{0, 0}, iD, iH, iW, n, st))) return rc;
            if ((rc = launch_transpose<float, false>(xs, xT, n * D, iH, iW, st))) return rc;
            if ((rc = launch_transpose<float, false>(g_right + s0 * 5 * HW, gT2, n * 5, iH, iW, st))) return rc;
            if ((rc = launch_transpose<float, false>(g_left + s0 * 5 * HW, gT3, n * 5, iH, iW, st))) return rc;
            if ((rc = run_vert_fwd<VMODE_RAW>(vc, xT, gT2, A2T, nullptr, 0, MaskIds{0, 0}, iD, iW, iH, n, st))) return rc;
            if ((rc = run_vert_fwd<VMODE_RAW>(vc, xT, gT3, A3T, nullptr, 1, MaskIds{0, 0}, iD, iW, iH, n, st))) return rc;
            if ((rc = launch_merge4_transposed(A0, A1, A2T, A3T, out + s0 * S, mask + s0 * S, n * D, iH, iW, st))) return rc;
            continue;
        }
        float *xT = (float *)(ws + w.xT), *outT = (float *)(ws + w.outT);
        uint8_t *maskT = (uint8_t *)(ws + w.maskT);
        float *gT2 = (float *)(ws + w.gT2), *gT3 = (float *)(ws + w.gT3);
        const float *xs = x + s0 * S;
        float *os = out + s0 * S;
        uint8_t *ms = mask + s0 * S;
        if (fwd_direct_ok(vc, iD, iH, iW, n * S)) {
            // transpose-free: right and left aggregates straight from the standard layout into
            // scratch, then `down` merges all three, then `up` merges on top
            float *a2 = xT, *a3 = outT;
            rc = launch_tma_hraw<0>(vc, xs, g_right + s0 * 5 * HW, a2, iD, iH, iW, n, st);
            if (rc == GANET_OK) rc = launch_tma_hraw<1>(vc, xs, g_left + s0 * 5 * HW, a3, iD, iH, iW, n, st);
            if (rc == GANET_OK)
                rc = launch_tma_fwd<VMODE_FIRST3>(vc, xs, g_down + s0 * 5 * HW, os, ms, 0, MaskIds{0, 0}, iD, iH,
                                                  iW, n, st, a2, a3);
            if (rc == GANET_OK)
                rc = run_vert_fwd<VMODE_COMBINE>(vc, xs, g_up + s0 * 5 * HW, os, ms, 1, MaskIds{0, 1}, iD, iH, iW, n, st);
            if (rc != kNotApplicable) {
                if (rc) return rc;
                continue;
            }
        }
        // horizontal scans = vertical scans of the H<->W transposed slices
        if ((rc = launch_transpose<float, false>(xs, xT, n * D, iH, iW, st))) return rc;
        if ((rc = launch_transpose<float, false>(g_right + s0 * 5 * HW, gT2, n * 5, iH, iW, st))) return rc;
        if ((rc = launch_transpose<float, false>(g_left + s0 * 5 * HW, gT3, n * 5, iH, iW, st))) return rc;
        if ((rc = run_vert_fwd<VMODE_FIRST>(vc, xT, gT2, outT, maskT, 0, MaskIds{2, 2}, iD, iW, iH, n, st))) return rc;
        if ((rc = run_vert_fwd<VMODE_SECOND>(vc, xT, gT3, outT, maskT, 1, MaskIds{2, 3}, iD, iW, iH, n, st))) return rc;
        if ((rc = launch_transpose<float, false>(outT, os, n * D, iW, iH, st))) return rc;
        if ((rc = launch_transpose_u8(maskT, ms, n * D, iW, iH, st))) return rc;
        // vertical scans merge on top; the tie rule keeps the lower direction id
        if ((rc = run_vert_fwd<VMODE_COMBINE>(vc, xs, g_down + s0 * 5 * HW, os, ms, 0, MaskIds{0, 0}, iD, iH, iW, n, st))) return rc;
        if ((rc = run_vert_fwd<VMODE_COMBINE>(vc, xs, g_up + s0 * 5 * HW, os, ms, 1, MaskIds{0, 1}, iD, iH, iW, n, st))) return rc;
    }
    return GANET_OK;
}

GANET_API int ganet_sga_direction(const float *x, const float *g, float *a, int dir, int64_t N,
                                  int64_t C, int64_t D, int64_t H, int64_t W,
                                  ganet_stream_t stream)
{
    if (!x || !g || !a || dir < 0 || dir > 3) return GANET_EINVAL;
    int rc = check_dims(N, C, D, H, W);
    if (rc) return rc;
    Cfg c;
    VCfg vc;
    if (dir < 2 && pick_vert_cfg((int)D, &vc))
        return run_vert_fwd<VMODE_RAW>(vc, x, g, a, nullptr, dir, MaskIds{0, 0}, (int)D, (int)H,
                                          (int)W, N * C, (cudaStream_t)stream);
    if (dir >= 2 && hscan_ok((int)D, (int)H, (int)W)) {
        rc = dir == 2 ? launch_hscan_fwd<0>(x, g, a, (int)D, (int)H, (int)W, N * C, (cudaStream_t)stream)
                      : launch_hscan_fwd<1>(x, g, a, (int)D, (int)H, (int)W, N * C, (cudaStream_t)stream);
        if (rc != kNotApplicable) return rc;
    }
    if (!pick_cfg((int)D, dir < 2, &c)) return GANET_EUNSUPPORTED;
    return launch_fwd<MODE_RAW>(c, x, g, a, nullptr, dir, (int)D, (int)H, (int)W, N * C,
                                (cudaStream_t)stream);
}

// slow generic backward: `a` scratch only
static int sga_backward_lines(const float *x, const float *const g[4], const uint8_t *mask,
                              const float *go, float *gi, float *const gg[4], int32_t *max_idx,
                              float *a, long long chunk, int D, int H, int W, long long ns,
                              cudaStream_t st)
{
    Cfg cv, ch;
    if (!pick_cfg(D, true, &cv) || !pick_cfg(D, false, &ch)) return GANET_EUNSUPPORTED;
    const long long S = (long long)D * H * W, HW = (long long)H * W;
    static const int order[4] = {3, 0, 1, 2};    // the reference's order (:1040, :1061, :1084, :1106)
    for (long long s0 = 0; s0 < ns; s0 += chunk) {
        const long long n = (ns - s0 < chunk) ? ns - s0 : chunk;
        for (int o = 0; o < 4; o++) {
            const int dir = order[o];
            const Cfg c = dir < 2 ? cv : ch;
            int rc = launch_fwd<MODE_RAW>(c, x + s0 * S, g[dir] + s0 * 5 * HW, a, nullptr, dir, D, H, W, n, st);
            if (rc) return rc;
            rc = launch_bwd(c, x + s0 * S, g[dir] + s0 * 5 * HW, a, mask + s0 * S, go + s0 * S,
                            gi + s0 * S, gg[dir] + s0 * 5 * HW,
                            (max_idx && dir == 2) ? max_idx + s0 * HW : nullptr, dir, o > 0, D, H, W, n, st);
            if (rc) return rc;
        }
    }
    return GANET_OK;
}

GANET_API int ganet_sga_backward(const float *x, const float *g_down, const float *g_up,
                                 const float *g_right, const float *g_left, const uint8_t *mask,
                                 const float *aggregates, const float *grad_out, float *grad_in,
                                 float *gg_down, float *gg_up, float *gg_right, float *gg_left,
                                 int32_t *max_idx, void *workspace, size_t workspace_bytes,
                                 int64_t N, int64_t C, int64_t D, int64_t H, int64_t W,
                                 ganet_stream_t stream)
{
    if (!x || !g_down || !g_up || !g_right || !g_left || !mask || !grad_out || !grad_in ||
        !gg_down || !gg_up || !gg_right || !gg_left || !workspace)
        return GANET_EINVAL;
    int rc = check_dims(N, C, D, H, W);
    if (rc) return rc;
    const long long S = D * H * W, HW = H * W, ns = N * C;
    cudaStream_t st = (cudaStream_t)stream;
    const float *g[4] = {g_down, g_up, g_right, g_left};
    float *gg[4] = {gg_down, gg_up, gg_right, gg_left};
    const int iD = (int)D, iH = (int)H, iW = (int)W;
    VCfg vc;
    if (!pick_vert_cfg(iD, &vc)) {
        if (aggregates) return GANET_EUNSUPPORTED;       // the generic path recomputes
        const long long fit = (long long)(workspace_bytes / ((size_t)S * sizeof(float)));
        if (fit < 1) return GANET_EWORKSPACE;
        return sga_backward_lines(x, g, mask, grad_out, grad_in, gg, max_idx, (float *)workspace,
                                  fit < ns ? fit : ns, iD, iH, iW, ns, st);
    }
    const bool kept = aggregates != nullptr;     // forward kept the four aggregates: no recompute
    const bool hs = hscan_ok(iD, iH, iW);
    const long long chunk =
        fit_slices([&](long long n) { return bwd_ws(n, S, HW, kept, hs).total; }, workspace_bytes, ns);
    if (chunk < 1) return GANET_EWORKSPACE;
    char *ws = (char *)workspace;
    for (long long s0 = 0; s0 < ns; s0 += chunk) {
        const long long n = (ns - s0 < chunk) ? ns - s0 : chunk;
        const BwdWs w = bwd_ws(n, S, HW, kept, hs);
        if (hs) {
            // all four reverse sweeps in the standard layout; gradInput: `down` stores, the other
            // three accumulate with TMA reduce-adds
            float *a = (float *)(ws + w.a);
            const float *xs = x + s0 * S, *gos = grad_out + s0 * S;
            const uint8_t *ms = mask + s0 * S;
            float *gis = grad_in + s0 * S;
            for (int dir = 0; dir < 4; dir++) {
                const float *gd = g[dir] + s0 * 5 * HW;
                float *ggd = gg[dir] + s0 * 5 * HW;
                const float *ak = kept ? aggregates + dir * ns * S + s0 * S : a;
                if (dir < 2) {
                    if (!kept)
                        if ((rc = run_vert_fwd<VMODE_RAW>(vc, xs, gd, a, nullptr, dir, MaskIds{0, 0}, iD, iH, iW, n, st))) return rc;
                    if ((rc = run_vert_bwd(vc, xs, gd, ak, ms, gos, gis, ggd, dir, dir, dir > 0, iD, iH, iW, n, st))) return rc;
                } else if (dir == 2) {
                    if (!kept)
                        if ((rc = launch_hscan_fwd<0>(xs, gd, a, iD, iH, iW, n, st))) return rc == kNotApplicable ? GANET_EUNSUPPORTED : rc;
                    rc = launch_hscan_bwd<0>(xs, gd, ak, ms, gos, gis, ggd, max_idx ? max_idx + s0 * HW : nullptr, 2, 1, iD, iH, iW, n, st);
                    if (rc) return rc == kNotApplicable ? GANET_EUNSUPPORTED : rc;
                } else {
                    if (!kept)
                        if ((rc = launch_hscan_fwd<1>(xs, gd, a, iD, iH, iW, n, st))) return rc == kNotApplicable ? GANET_EUNSUPPORTED : rc;
                    rc = launch_hscan_bwd<1>(xs, gd, ak, ms, gos, gis, ggd, nullptr, 3, 1, iD, iH, iW, n, st);
                    if (rc) return rc == kNotApplicable ? GANET_EUNSUPPORTED : rc;
                }
            }
            continue;
        }
        float *a = (float *)(ws + w.a), *xT = (float *)(ws + w.xT), *goT = (float *)(ws + w.goT);
        float *giT = (float *)(ws + w.giT), *gT = (float *)(ws + w.gT), *ggT = (float *)(ws + w.ggT);
        uint8_t *maskT = (uint8_t *)(ws + w.maskT);
        const float *xs = x + s0 * S, *gos = grad_out + s0 * S;
        const uint8_t *ms = mask + s0 * S;
        float *gis = grad_in + s0 * S;
        // vertical directions in place
        for (int dir = 0; dir < 2; dir++) {
            const float *ak = kept ? aggregates + dir * ns * S + s0 * S : a;
            if (!kept)
                if ((rc = run_vert_fwd<VMODE_RAW>(vc, xs, g[dir] + s0 * 5 * HW, a, nullptr, dir, MaskIds{0, 0}, iD, iH, iW, n, st))) return rc;
            if ((rc = run_vert_bwd(vc, xs, g[dir] + s0 * 5 * HW, ak, ms, gos, gis, gg[dir] + s0 * 5 * HW, dir, dir, dir > 0, iD, iH, iW, n, st))) return rc;
        }
        // horizontal directions on the transposed slices
        if (kept) xT = const_cast<float *>(aggregates) + 4 * ns * S + s0 * S;
        else if ((rc = launch_transpose<float, false>(xs, xT, n * D, iH, iW, st))) return rc;
        if ((rc = launch_transpose<float, false>(gos, goT, n * D, iH, iW, st))) return rc;
        if ((rc = launch_transpose_u8(ms, maskT, n * D, iH, iW, st))) return rc;
        for (int dir = 2; dir < 4; dir++) {
            const float *ak = kept ? aggregates + dir * ns * S + s0 * S : a;
            if ((rc = launch_transpose<float, false>(g[dir] + s0 * 5 * HW, gT, n * 5, iH, iW, st))) return rc;
            if (!kept)
                if ((rc = run_vert_fwd<VMODE_RAW>(vc, xT, gT, a, nullptr, dir - 2, MaskIds{0, 0}, iD, iW, iH, n, st))) return rc;
            if (max_idx && dir == 2) {
                for (long long z0 = 0; z0 < n; z0 += 65535) {          // gridDim.y limit
                    const long long nz = n - z0 < 65535 ? n - z0 : 65535;
                    dim3 grid((unsigned)((HW + 255) / 256), (unsigned)nz);
                    max_depth_from_transposed_kernel<<<grid, 256, 0, st>>>(ak + z0 * S, max_idx + (s0 + z0) * HW,
                                                                           iD, iH, iW);
                }
                GANET_RETURN_IF_LAUNCH_FAILED();
            }
            if ((rc = run_vert_bwd(vc, xT, gT, ak, maskT, goT, giT, ggT, dir - 2, dir, dir > 2, iD, iW, iH, n, st))) return rc;
            if ((rc = launch_transpose<float, false>(ggT, gg[dir] + s0 * 5 * HW, n * 5, iW, iH, st))) return rc;
        }
        if ((rc = launch_transpose<float, true>(giT, gis, n * D, iW, iH, st))) return rc;   // gi += giT^T
    }
    return GANET_OK;
}
