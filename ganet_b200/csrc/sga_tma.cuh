// TMA-staged vertical SGA scans (sm_100a): the same algorithm and thread mapping as
// sga_vert.cuh (lane = pixel column, warp = depth chunk, one exchange barrier per row),
// but every global transfer is a bulk tensor copy issued by a dedicated producer warp:
//
//   producer warp (lane 0)                      NW consumer warps
//   ------------------------------------       ------------------------------------------
//   TMA-load row tiles [D][32 cols] of x,       wait full[stage]
//   guidance, (out, mask | gradOut, A)  --->    read tiles from smem (LDS, immediate offsets)
//   into a ring of S stages                     run the recurrence in registers
//                                               write results back INTO the stage (in place)
//   wait done[stage]                    <---    fence.proxy.async ; arrive done[stage]
//   TMA-store the result tiles                  exchange chunk edges / max ; bar.sync (consumers)
//   wait until the store has read smem,
//   re-arm full[stage], load row t+S
//
// The compute warps execute no global loads/stores, no 64-bit address arithmetic and no
// bounds predicates (TMA clips partial strips and zero-fills), which is what made the
// LDG version issue-bound (profiles/r01_ncu_full_vertical_kernels.txt: 36-99 executed
// instructions per voxel).  Rows arrive as full 128-byte lines.
//
// Reference semantics: identical to sga_vert.cuh (same step functions, same exchange).
#pragma once
#include <math.h>

#include "common.cuh"
#include "sga_step.cuh"
#include "sga_vert.cuh"
#include "tma_utils.cuh"

namespace ganet {

struct TmaFwdMaps { CUtensorMap x, g, out, mask, a2, a3; };   // a2/a3: FIRST3 inputs
struct TmaHrawMaps { CUtensorMap x, g, out; };
struct TmaBwdMaps { CUtensorMap x, g, a, mask, go, gi; };

__host__ __device__ inline int align128(int v) { return (v + 127) & ~127; }

// ---- shared-memory plan (host and device agree through these helpers) ------------------
struct FwdPlan { int off_x, off_o, off_m, off_a3, off_g, stage_bytes; };
__host__ __device__ inline FwdPlan fwd_plan(int D, bool combine, bool three = false)
{
    FwdPlan p;
    const int xb = D * 128;
    p.off_x = 0;
    p.off_o = xb;
    p.off_m = 2 * xb;
    p.off_a3 = 2 * xb + align128(D * 32);
    p.off_g = three ? p.off_a3 + xb : (combine ? p.off_a3 : xb);
    p.stage_bytes = p.off_g + 640;            // 5 guidance rows of 32 floats
    return p;
}
struct BwdPlan { int off_x, off_go, off_a, off_m, off_g, stage_bytes; };
__host__ __device__ inline BwdPlan bwd_plan(int D)
{
    BwdPlan p;
    const int xb = D * 128;
    p.off_x = 0;
    p.off_go = xb;
    p.off_a = 2 * xb;
    p.off_m = 3 * xb;
    p.off_g = 3 * xb + align128(D * 32);
    p.stage_bytes = p.off_g + 640;
    return p;
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <int K, int MAXW, int MODE, bool FULL>
__global__ void __launch_bounds__(MAXW * 32 + 32)
sga_tma_fwd_kernel(const __grid_constant__ TmaFwdMaps maps, int dir, MaskIds ids, int D, int H,
                   int strips, int S, int stream_hint, int PF)
{
    static_assert(K % 2 == 0, "depth parity must be a compile-time property");
    constexpr bool kThree = (MODE == VMODE_FIRST3);       // merge down with the two horizontal aggregates
    constexpr bool kCombine = (MODE == VMODE_SECOND || MODE == VMODE_COMBINE || kThree);
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, j = tid >> 5;
    const int NW = (blockDim.x >> 5) - 1;                 // consumer warps; warp NW is the producer
    const long long s = blockIdx.x / strips;
    const int strip = blockIdx.x - (int)(s * strips);
    const int w0 = strip * 32;
    const FwdPlan pl = fwd_plan(D, kCombine, kThree);
    const int plane = NW * 32;
    float *ex = reinterpret_cast<float *>(smem + (size_t)S * pl.stage_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * pl.stage_bytes + 2 * 3 * plane * 4);
    uint64_t *done = full + S;
    const int c2x = (int)(s * D), c2g = (int)(s * 5);

    if (tid == 0) {
        for (int i = 0; i < S; i++) { mbar_init(&full[i], 1); mbar_init(&done[i], NW * 32); }
        fence_mbarrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    if (j == NW) {                                        // ---------------- producer
        if (lane == 0) {
            const uint64_t pol = l2_evict_first_policy();
            auto tma_load_3d = [&](void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2) {
                if (stream_hint) ganet::tma_load_3d_hint(dst, m, bar, c0, c1, c2, pol);
                else ganet::tma_load_3d(dst, m, bar, c0, c1, c2);
            };
            auto tma_store_3d = [&](const CUtensorMap *m, const void *src, int c0, int c1, int c2) {
                if (stream_hint) ganet::tma_store_3d_hint(m, src, c0, c1, c2, pol);
                else ganet::tma_store_3d(m, src, c0, c1, c2);
            };
            const unsigned tx = D * 128 + 640 + (kCombine ? D * 128 : 0) + (MODE == VMODE_COMBINE ? D * 32 : 0) +
                                (kThree ? D * 128 : 0);
            auto issue = [&](int t) {
                const int st = t % S;
                const int h = (dir == 0) ? t : H - 1 - t;
                unsigned char *b = smem + (size_t)st * pl.stage_bytes;
                mbar_arrive_expect_tx(&full[st], tx);
                tma_load_3d(b + pl.off_x, &maps.x, &full[st], w0, h, c2x);
                tma_load_3d(b + pl.off_g, &maps.g, &full[st], w0, h, c2g);
                if (kThree) {
                    tma_load_3d(b + pl.off_o, &maps.a2, &full[st], w0, h, c2x);
                    tma_load_3d(b + pl.off_a3, &maps.a3, &full[st], w0, h, c2x);
                } else if (kCombine) {
                    tma_load_3d(b + pl.off_o, &maps.out, &full[st], w0, h, c2x);
                }
                if (MODE == VMODE_COMBINE) tma_load_3d(b + pl.off_m, &maps.mask, &full[st], w0, h, c2x);
            };
            // rows beyond the ring are pulled into L2 (PF rows ahead; 0 = off)
            auto prefetch = [&](int t) {
                if (t < H) {
                    const int h = (dir == 0) ? t : H - 1 - t;
                    tma_prefetch_3d(&maps.x, w0, h, c2x);
                    if (kThree) { tma_prefetch_3d(&maps.a2, w0, h, c2x); tma_prefetch_3d(&maps.a3, w0, h, c2x); }
                    else if (kCombine) tma_prefetch_3d(&maps.out, w0, h, c2x);
                    if (MODE == VMODE_COMBINE) tma_prefetch_3d(&maps.mask, w0, h, c2x);
                }
            };
            for (int t = 0; t < S && t < H; t++) issue(t);
            for (int t = S; t < S + PF; t++) prefetch(t);
            for (int t = 0; t < H; t++) {
                const int st = t % S;
                if (PF > 0) prefetch(t + S + PF);
                mbar_wait(&done[st], (t / S) & 1);
                const int h = (dir == 0) ? t : H - 1 - t;
                unsigned char *b = smem + (size_t)st * pl.stage_bytes;
                if (kCombine) {
                    tma_store_3d(&maps.out, b + pl.off_o, w0, h, c2x);
                    tma_store_3d(&maps.mask, b + pl.off_m, w0, h, c2x);
                } else {
                    tma_store_3d(&maps.out, b + pl.off_x, w0, h, c2x);   // A written over the x tile
                }
                tma_commit();
                if (t + S < H) {
                    tma_wait_read_all();                  // the stores have read the stage
                    issue(t + S);
                }
            }
            tma_wait_all();
        }
        return;
    }

    // ---------------- consumers
    const int d0 = K * j;
    const int toff = (d0 * 32 + lane) * 4;                // this thread's first element in an f32 tile
    const int moff = d0 * 32 + lane;
    float P[K];
#pragma unroll
    for (int i = 0; i < K; i++) P[i] = 0.f;

    for (int t = 0; t < H; t++) {
        const int st = t % S;
        unsigned char *b = smem + (size_t)st * pl.stage_bytes;
        mbar_wait(&full[st], (t / S) & 1);
        float *xt = reinterpret_cast<float *>(b + pl.off_x + toff);
        float *ot = reinterpret_cast<float *>(b + pl.off_o + toff);
        uint8_t *mt = b + pl.off_m + moff;
        const float *gt = reinterpret_cast<const float *>(b + pl.off_g) + lane;
        float xc[K], w[5], oc[K], o3[K];
        uint8_t mc[K];
#pragma unroll
        for (int i = 0; i < K; i++) xc[i] = (FULL || d0 + i < D) ? xt[i * 32] : 0.f;
#pragma unroll
        for (int k = 0; k < 5; k++) w[k] = gt[k * 32];
        if (kCombine) {
#pragma unroll
            for (int i = 0; i < K; i++) oc[i] = (FULL || d0 + i < D) ? ot[i * 32] : 0.f;
        }
        if (kThree) {
            const float *a3t = reinterpret_cast<const float *>(b + pl.off_a3 + toff);
#pragma unroll
            for (int i = 0; i < K; i++) o3[i] = (FULL || d0 + i < D) ? a3t[i * 32] : 0.f;
        }
        if (MODE == VMODE_COMBINE) {
#pragma unroll
            for (int i = 0; i < K; i++) mc[i] = (FULL || d0 + i < D) ? mt[i * 32] : (uint8_t)0;
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
#pragma unroll
        for (int i = 0; i < K; i++) {
            if (FULL || d0 + i < D) {
                if (!kCombine) {
                    xt[i * 32] = A[i];
                } else if (MODE == VMODE_SECOND) {
                    const bool m = oc[i] < A[i];
                    ot[i * 32] = m ? A[i] : oc[i];
                    mt[i * 32] = m ? (uint8_t)ids.mine : (uint8_t)ids.first;
                } else if (kThree) {
                    // this direction (lowest id) first, then right (2), then left (3): strict <
                    // keeps the lower id on ties, as the reference's Max chain does (:23-36)
                    float best = A[i];
                    uint8_t id = (uint8_t)ids.mine;
                    if (best < oc[i]) { best = oc[i]; id = 2; }
                    if (best < o3[i]) { best = o3[i]; id = 3; }
                    ot[i * 32] = best;
                    mt[i * 32] = id;
                } else {
                    const bool m = oc[i] < A[i] || (oc[i] == A[i] && ids.mine < (int)mc[i]);
                    ot[i * 32] = m ? A[i] : oc[i];
                    mt[i * 32] = m ? (uint8_t)ids.mine : mc[i];
                }
            }
        }
        fence_proxy_async();
        mbar_arrive(&done[st]);
        {
            float *wb = ex + (t & 1) * 3 * plane + j * 32 + lane;
            wb[0] = A[0];
            wb[plane] = A[K - 1];
            wb[2 * plane] = FULL ? chunk_max<K>(A, 0, K) : chunk_max<K>(A, d0, D);
        }
        named_barrier(1, NW * 32);
#pragma unroll
        for (int i = 0; i < K; i++) P[i] = A[i];
    }
}

// ---------------------------------------------------------------------------
// horizontal scans (right / left) in the STANDARD layout, no transposes.
// A CTA owns 32 image rows of one slice; lane = row, warp = depth chunk.  TMA boxes of
// (4 columns x 32 rows x D planes) land as [d][row][4] in shared memory, so every thread
// fetches ITS four consecutive scan steps with one conflict-free 16-byte access, walks them
// in registers (exchange + consumer barrier per step as in the vertical kernel), writes the
// four aggregates back in place with one 16-byte store, and the producer stores the box.
// DIR 0 = right (columns ascending), 1 = left.  Output: the raw aggregate.
// ---------------------------------------------------------------------------
template <int K, int MAXW, int DIR, bool FULL>
__global__ void __launch_bounds__(MAXW * 32 + 32)
sga_tma_hraw_kernel(const __grid_constant__ TmaHrawMaps maps, int D, int W, int strips, int S)
{
    static_assert(K % 2 == 0, "depth parity must be a compile-time property");
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, j = tid >> 5;
    const int NW = (blockDim.x >> 5) - 1;
    const long long s = blockIdx.x / strips;
    const int strip = blockIdx.x - (int)(s * strips);
    const int h0 = strip * 32;
    const int xbytes = D * 512, stage_bytes = xbytes + 2560;      // [D][32][4] + [5][32][4] floats
    const int plane = NW * 32;
    float *ex = reinterpret_cast<float *>(smem + (size_t)S * stage_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * stage_bytes + 2 * 3 * plane * 4);
    uint64_t *done = full + S;
    const int c2x = (int)(s * D), c2g = (int)(s * 5);
    const int nb = W / 4;                                         // boxes along the scan

    if (tid == 0) {
        for (int i = 0; i < S; i++) { mbar_init(&full[i], 1); mbar_init(&done[i], NW * 32); }
        fence_mbarrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    if (j == NW) {                                        // ---------------- producer
        if (lane == 0) {
            const unsigned tx = xbytes + 2560;
            auto col_of = [&](int b) { return 4 * (DIR == 0 ? b : nb - 1 - b); };
            auto issue = [&](int b) {
                const int st = b % S;
                unsigned char *p = smem + (size_t)st * stage_bytes;
                mbar_arrive_expect_tx(&full[st], tx);
                tma_load_3d(p, &maps.x, &full[st], col_of(b), h0, c2x);
                tma_load_3d(p + xbytes, &maps.g, &full[st], col_of(b), h0, c2g);
            };
            for (int b = 0; b < S && b < nb; b++) issue(b);
            for (int b = 0; b < nb; b++) {
                const int st = b % S;
                mbar_wait(&done[st], (b / S) & 1);
                tma_store_3d(&maps.out, smem + (size_t)st * stage_bytes, col_of(b), h0, c2x);
                tma_commit();
                if (b + S < nb) {
                    tma_wait_read_all();
                    issue(b + S);
                }
            }
            tma_wait_all();
        }
        return;
    }

    // ---------------- consumers
    const int d0 = K * j;
    const int toff = (d0 * 32 + lane) * 16;               // bytes: this thread's first 16-byte cell
    float P[K];
#pragma unroll
    for (int i = 0; i < K; i++) P[i] = 0.f;

    for (int b = 0; b < nb; b++) {
        const int st = b % S;
        unsigned char *p = smem + (size_t)st * stage_bytes;
        mbar_wait(&full[st], (b / S) & 1);
        float4 *xt = reinterpret_cast<float4 *>(p + toff);
        const float4 *gt = reinterpret_cast<const float4 *>(p + xbytes) + lane;
        float xq[K][4], gq[5][4];
#pragma unroll
        for (int i = 0; i < K; i++) {
            const float4 v = (FULL || d0 + i < D) ? xt[i * 32] : make_float4(0.f, 0.f, 0.f, 0.f);
            xq[i][0] = v.x; xq[i][1] = v.y; xq[i][2] = v.z; xq[i][3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const float4 v = gt[k * 32];
            gq[k][0] = v.x; gq[k][1] = v.y; gq[k][2] = v.z; gq[k][3] = v.w;
        }
#pragma unroll
        for (int qq = 0; qq < 4; qq++) {
            constexpr int dummy = 0; (void)dummy;
            const int q = (DIR == 0) ? qq : 3 - qq;       // compile-time after unrolling
            const int t = b * 4 + qq;                     // scan step
            float xc[K], w[5], A[K];
#pragma unroll
            for (int i = 0; i < K; i++) xc[i] = xq[i][q];
#pragma unroll
            for (int k = 0; k < 5; k++) w[k] = gq[k][q];
            if (t == 0) {
                sga_first_step<K>(xc, w, A);
            } else {
                const float *eb = ex + ((t - 1) & 1) * 3 * plane;
                const float up = (j > 0) ? eb[plane + (j - 1) * 32 + lane] : 0.f;
                const float dn = (j + 1 < NW) ? eb[(j + 1) * 32 + lane] : 0.f;
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
#pragma unroll
            for (int i = 0; i < K; i++) { P[i] = A[i]; xq[i][q] = A[i]; }
            named_barrier(1, NW * 32);
        }
#pragma unroll
        for (int i = 0; i < K; i++)
            if (FULL || d0 + i < D) xt[i * 32] = make_float4(xq[i][0], xq[i][1], xq[i][2], xq[i][3]);
        fence_proxy_async();
        mbar_arrive(&done[st]);
    }
}

// ---------------------------------------------------------------------------
// backward (reverse sweep).  gradInput leaves through the gradOut tile (in place);
// accumulate = 1: the old gradInput row is read with plain coalesced loads and added here;
// accumulate = 2: the tile leaves as a TMA reduce-add, nothing is read back.
// The guidance gradients (5 floats per pixel) are written directly by warps 0..4.
// ---------------------------------------------------------------------------
template <int K, int MAXW, bool FULL>
__global__ void __launch_bounds__(MAXW * 32 + 32)
sga_tma_bwd_kernel(const __grid_constant__ TmaBwdMaps maps, const float *gi_old, float *__restrict__ gg,
                   int dir, int mask_id, int accumulate, int D, int H, int W, int strips, int S,
                   int stream_hint, int PF)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, j = tid >> 5;
    const int NW = (blockDim.x >> 5) - 1;
    const long long s = blockIdx.x / strips;
    const int strip = blockIdx.x - (int)(s * strips);
    const int w0 = strip * 32;
    const BwdPlan pl = bwd_plan(D);
    const int plane = NW * 32;
    const int bufsz = NBW * plane;
    float *ex = reinterpret_cast<float *>(smem + (size_t)S * pl.stage_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * pl.stage_bytes + 2 * bufsz * 4);
    uint64_t *done = full + S;
    const int c2x = (int)(s * D), c2g = (int)(s * 5);

    if (tid == 0) {
        for (int i = 0; i < S; i++) { mbar_init(&full[i], 1); mbar_init(&done[i], NW * 32); }
        fence_mbarrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    // iteration `it` handles scan position t = H-1-it, image row h(t)
    if (j == NW) {                                        // ---------------- producer
        if (lane == 0) {
            const uint64_t pol = l2_evict_first_policy();
            auto tma_load_3d = [&](void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2) {
                if (stream_hint) ganet::tma_load_3d_hint(dst, m, bar, c0, c1, c2, pol);
                else ganet::tma_load_3d(dst, m, bar, c0, c1, c2);
            };
            auto tma_store_3d = [&](const CUtensorMap *m, const void *src, int c0, int c1, int c2) {
                if (stream_hint) ganet::tma_store_3d_hint(m, src, c0, c1, c2, pol);
                else ganet::tma_store_3d(m, src, c0, c1, c2);
            };
            auto row_of = [&](int t) { return (dir == 0) ? t : H - 1 - t; };
            auto issue = [&](int it) {
                const int st = it % S;
                const int t = H - 1 - it;
                unsigned char *b = smem + (size_t)st * pl.stage_bytes;
                const unsigned tx = 2 * D * 128 + D * 32 + 640 + (t >= 1 ? D * 128 : 0);
                mbar_arrive_expect_tx(&full[st], tx);
                tma_load_3d(b + pl.off_x, &maps.x, &full[st], w0, row_of(t), c2x);
                tma_load_3d(b + pl.off_go, &maps.go, &full[st], w0, row_of(t), c2x);
                tma_load_3d(b + pl.off_m, &maps.mask, &full[st], w0, row_of(t), c2x);
                tma_load_3d(b + pl.off_g, &maps.g, &full[st], w0, row_of(t), c2g);
                if (t >= 1) tma_load_3d(b + pl.off_a, &maps.a, &full[st], w0, row_of(t - 1), c2x);
            };
            auto prefetch = [&](int it) {
                if (it < H) {
                    const int t = H - 1 - it;
                    tma_prefetch_3d(&maps.x, w0, row_of(t), c2x);
                    tma_prefetch_3d(&maps.go, w0, row_of(t), c2x);
                    tma_prefetch_3d(&maps.mask, w0, row_of(t), c2x);
                    if (t >= 1) tma_prefetch_3d(&maps.a, w0, row_of(t - 1), c2x);
                }
            };
            for (int it = 0; it < S && it < H; it++) issue(it);
            for (int it = S; it < S + PF; it++) prefetch(it);
            for (int it = 0; it < H; it++) {
                const int st = it % S;
                if (PF > 0) prefetch(it + S + PF);
                mbar_wait(&done[st], (it / S) & 1);
                unsigned char *b = smem + (size_t)st * pl.stage_bytes;
                if (accumulate == 2) tma_reduce_add_3d(&maps.gi, b + pl.off_go, w0, row_of(H - 1 - it), c2x);
                else tma_store_3d(&maps.gi, b + pl.off_go, w0, row_of(H - 1 - it), c2x);
                tma_commit();
                if (it + S < H) {
                    tma_wait_read_all();
                    issue(it + S);
                }
            }
            tma_wait_all();
        }
        return;
    }

    // ---------------- consumers
    const int wcol = w0 + lane;
    const bool wok = wcol < W;
    const int wc = wok ? wcol : W - 1;
    const long long HW = (long long)H * W;
    const int d0 = K * j;
    const int dfirst = FULL ? d0 : min(d0, D - 1);
    const int toff = (d0 * 32 + lane) * 4;
    const int moff = d0 * 32 + lane;
    const long long ps = (dir == 0) ? W : -W;
    const long long pix = ((dir == 0) ? wc : (long long)(H - 1) * W + wc) + (H - 1) * ps;   // last position
    addr_t girow = (addr_t)(gi_old + s * (long long)D * HW + dfirst * HW + pix);
    float *ggrow = gg + s * 5 * HW + pix;
    const long long psb = ps * 4;
    unsigned offb[K];
#pragma unroll
    for (int i = 0; i < K; i++) {
        offb[i] = (unsigned)(i * (int)HW) * 4u;
        asm volatile("" : "+r"(offb[i]));
    }

    float Tn[K], wnx[5];
#pragma unroll
    for (int i = 0; i < K; i++) Tn[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 5; k++) wnx[k] = 0.f;
    float *gg_next = ggrow;

    for (int it = 0; it < H; it++) {
        const int t = H - 1 - it;
        const int st = it % S;
        unsigned char *b = smem + (size_t)st * pl.stage_bytes;
        float gold[K];
        if (accumulate == 1) {                              // issued before the wait: overlaps it
#pragma unroll
            for (int i = 0; i < K; i++)
                gold[i] = (FULL || d0 + i < D) ? *at<const float>(girow, offb[i]) : 0.f;
        } else {
#pragma unroll
            for (int i = 0; i < K; i++) gold[i] = 0.f;
        }
        mbar_wait(&full[st], (it / S) & 1);
        const float *xt = reinterpret_cast<const float *>(b + pl.off_x + toff);
        float *got = reinterpret_cast<float *>(b + pl.off_go + toff);
        const float *at_ = reinterpret_cast<const float *>(b + pl.off_a + toff);
        const uint8_t *mt = b + pl.off_m + moff;
        const float *gt = reinterpret_cast<const float *>(b + pl.off_g) + lane;

        float xv[K], t0[K], ap[K], w[5];
        float aup = 0.f, adn = 0.f;
#pragma unroll
        for (int i = 0; i < K; i++) {
            const bool ok = FULL || d0 + i < D;
            xv[i] = ok ? xt[i * 32] : 0.f;
            const float gv = ok ? got[i * 32] : 0.f;
            const uint8_t mv = ok ? mt[i * 32] : (uint8_t)255;
            t0[i] = (mv == mask_id) ? gv : 0.f;                           // get_temp_grad :38-48
            ap[i] = (ok && t >= 1) ? at_[i * 32] : 0.f;
        }
        if (t >= 1) {
            if (d0 >= 1) aup = at_[-32];                                  // A[d0-1, t-1]
            if (d0 + K < D) adn = at_[K * 32];                            // A[d0+K, t-1]
        }
#pragma unroll
        for (int k = 0; k < 5; k++) w[k] = gt[k * 32];

        float tc[K];
        if (t + 1 < H) {
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
            for (int k = j; k < 5; k += NW) {
                float tot;
                if (k == 4) {
                    tot = sum_tn * amax;
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

#pragma unroll
        for (int i = 0; i < K; i++) {
            const int d = d0 + i;
            if (FULL || d < D) {
                float v = tc[i] * w[0];
                if (d == 0) v += tc[i] * w[2];
                if (d == D - 1) v += tc[i] * w[3];
                got[i * 32] = gold[i] + v;                                // gradInput tile, in place
            }
        }

        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, stt = 0.f;
        float best = -INFINITY;
        int bi = d0;
#pragma unroll
        for (int i = 0; i < K; i++) {
            const int d = d0 + i;
            s0 += tc[i] * xv[i];
            stt += tc[i];
            if (t >= 1) {
                const float am = (i == 0) ? aup : ap[i == 0 ? 0 : i - 1];
                const float apn = (i == K - 1) ? adn : ap[i == K - 1 ? K - 1 : i + 1];
                s1 += tc[i] * ap[i];
                s2 += tc[i] * ((d >= 1) ? am : xv[i]);
                s3 += tc[i] * ((d + 1 < D) ? apn : xv[i]);
                if ((FULL || d < D) && ap[i] > best) { best = ap[i]; bi = d; }
            }
        }
        fence_proxy_async();
        mbar_arrive(&done[st]);
        {
            float *wb = ex + (t & 1) * bufsz + j * 32 + lane;
            wb[BX_TLO * plane] = tc[0];
            wb[BX_THI * plane] = tc[K - 1];
            wb[BX_ST * plane] = stt;
            wb[BX_S0 * plane] = s0;
            wb[BX_S1 * plane] = s1;
            wb[BX_S2 * plane] = s2;
            wb[BX_S3 * plane] = s3;
            wb[BX_AMAX * plane] = best;
            wb[BX_AIDX * plane] = __int_as_float(bi);
        }
        named_barrier(1, NW * 32);
#pragma unroll
        for (int i = 0; i < K; i++) Tn[i] = tc[i];
#pragma unroll
        for (int k = 0; k < 5; k++) wnx[k] = w[k];
        gg_next = ggrow;
        girow -= psb;
        ggrow -= ps;
    }
    {
        const float *eb = ex + lane;                          // scan position 0 sits in buffer 0
        for (int k = j; k < 5; k += NW) {
            float tot = 0.f;
            if (k == 0)
                for (int jj = 0; jj < NW; jj++) tot += eb[BX_S0 * plane + jj * 32];
            if (wok) gg_next[k * (int)HW] = tot;
        }
    }
}

}  // namespace ganet
