// Horizontal SGA scans (right / left) in the STANDARD layout, one image row per CTA.
//
// The scan axis of `right` / `left` is the contiguous axis of the volume (sga_right_forward,
// GANet_kernel.cu:507-565, walks it with one thread per row and W*4-byte strides).  Round 1 ran
// these directions as vertical scans of H<->W transposed copies; the transposes and the transposed
// merge were 29 % of the SGA time and 48 GB of 151 GB DRAM traffic per 920M-voxel sample.  A row
// band is in fact the friendliest access pattern the volume offers -- consecutive tiles of a row
// are adjacent in every depth plane (profiles/r02_tma_copy_probe.txt: 5.5-6.0 TB/s for 1-2 row
// bands with many small stages, against 4.5-4.9 TB/s for the 32-column strips of the vertical
// scans) -- so these kernels walk it directly:
//
//   * one CTA = one image row of one (n,c) slice = one scan line; one consumer warp + one producer
//     warp; 2-4 CTAs per SM
//   * LANE = DEPTH CHUNK: lane l holds depths [K*l, K*l+K) of the running row in registers, chunk
//     edges travel by warp shuffle, the running max over depth is one `redux.sync.max.f32`
//     (CREDUX.MAX.F32) -- no shared-memory exchange, no CTA barrier per step
//   * the producer streams tiles of BW scan steps, TMA box (BW columns, 1 row, D planes), into a ring
//     of S stages; the box lands as [d][c] with the 128-byte swizzle, so that a lane fetches its
//     FOUR consecutive scan steps of one depth with one 16-byte access (2-way bank conflict at
//     most: K is even, see swz128 below), walks them in registers, writes the four results back
//     in place, and the producer stores the tile
//
// Forward semantics and rounding: sga_step.cuh (bit-exact vs the reference build).  Backward:
// SURVEY.md Appendix A.3 / GANet_kernel.cu:567-718 (right), :780-933 (left), same arithmetic as
// sga_scan_bwd_kernel in sga.cu, evaluated one scan step late for the terms that need the
// previous position's aggregate (so that every tile of a stage covers the SAME columns).
#pragma once
#include <math.h>

#include "common.cuh"
#include "sga_step.cuh"
#include "tma_utils.cuh"

namespace ganet {

struct HFwdMaps { CUtensorMap x, g, out; };
struct HBwdMaps { CUtensorMap x, g, a, mask, go, gi; };

// CU_TENSOR_MAP_SWIZZLE_128B: byte-address bits [4,7) ^= bits [7,10) (tile base 1024-aligned).
// A lane reads 16 bytes at logical offset (d * BW + c) * 4 with d = K*l + i: for BW = 32 the row
// index IS bits [7,..), so eight consecutive lanes (one shared-memory phase) hit the 16-byte
// columns (c/4) ^ ((K*l + i) & 7): 4 distinct ones for K = 2 mod 4, i.e. a 2-way conflict.
__device__ __forceinline__ unsigned swz128(unsigned o) { return o ^ ((o >> 3) & 0x70u); }
// SWIZZLE_64B: bits [4,6) ^= bits [7,9) (64-byte rows, BW = 16: also 2-way for K = 2 mod 4);
// SWIZZLE_32B: bit 4 ^= bit 7 (the uint8 mask tile, 16- or 32-byte rows)
__device__ __forceinline__ unsigned swz64(unsigned o) { return o ^ ((o >> 3) & 0x30u); }
__device__ __forceinline__ unsigned swz32(unsigned o) { return o ^ ((o >> 3) & 0x10u); }
// fp32 tile of BW columns per row: 128-byte rows use the 128-byte swizzle, 64-byte rows the 64-byte one
template <int BW>
__device__ __forceinline__ unsigned tile_swz(unsigned o) { return BW == 32 ? swz128(o) : swz64(o); }

__host__ __device__ inline int round1k(int v) { return (v + 1023) & ~1023; }

// stage = [x / A tile][guidance tile], both 1024-aligned
__host__ __device__ inline int hfwd_stage_bytes(int D, int BW) { return round1k(D * BW * 4) + 1024; }
// stage = [x][gradOut -> gradInput][A][mask u8][guidance]
struct HBwdPlan { int off_x, off_go, off_a, off_m, off_g, stage_bytes; };
__host__ __device__ inline HBwdPlan hbwd_plan(int D, int BW)
{
    HBwdPlan p;
    const int t = round1k(D * BW * 4);
    p.off_x = 0;
    p.off_go = t;
    p.off_a = 2 * t;
    p.off_m = 3 * t;
    p.off_g = 3 * t + round1k(D * BW);
    p.stage_bytes = p.off_g + 1024;
    return p;
}

__device__ __forceinline__ unsigned char *align1k(unsigned char *p)
{
    const unsigned a = smem_u32(p);
    return p + (((a + 1023u) & ~1023u) - a);
}

// ---------------------------------------------------------------------------
// forward: raw aggregate of one horizontal direction.  DIR 0 = right (columns ascending),
// 1 = left.  grid = n_slices * H, block = 64 (warp 0 consumer, warp 1 producer).
// ---------------------------------------------------------------------------
template <int K, int BW, int DIR, bool FULL>
__global__ void __launch_bounds__(64)
sga_hscan_fwd_kernel(const __grid_constant__ HFwdMaps maps, int D, int H, int W, int S, int PF)
{
    static_assert(K % 2 == 0, "depth parity must be a compile-time property");
    static_assert(BW == 16 || BW == 32, "tile width");
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = align1k(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long s = blockIdx.x / H;
    const int h = blockIdx.x - (int)(s * H);
    const int stage_bytes = hfwd_stage_bytes(D, BW);
    const int xbytes = round1k(D * BW * 4);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * stage_bytes);
    uint64_t *done = full + S;
    const int c2x = (int)(s * D), c2g = (int)(s * 5);
    const int nb = (W + BW - 1) / BW;

    if (threadIdx.x == 0) {
        for (int i = 0; i < S; i++) { mbar_init(&full[i], 1); mbar_init(&done[i], 32); }
        fence_mbarrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    if (warp == 1) {                                      // ---------------- producer
        if (lane == 0) {
            const unsigned tx = (unsigned)(D * BW * 4 + 5 * BW * 4);
            auto col_of = [&](int b) { return BW * (DIR == 0 ? b : nb - 1 - b); };
            auto issue = [&](int b) {
                const int st = b % S;
                unsigned char *p = smem + (size_t)st * stage_bytes;
                mbar_arrive_expect_tx(&full[st], tx);
                tma_load_3d(p, &maps.x, &full[st], col_of(b), h, c2x);
                tma_load_3d(p + xbytes, &maps.g, &full[st], col_of(b), h, c2g);
            };
            // tiles S .. S+PF-1 ahead of the ring are pulled into L2 (PF = 0: off)
            auto prefetch = [&](int b) {
                if (b < nb) {
                    tma_prefetch_3d(&maps.x, col_of(b), h, c2x);
                    tma_prefetch_3d(&maps.g, col_of(b), h, c2g);
                }
            };
            for (int b = 0; b < S && b < nb; b++) issue(b);
            for (int b = S; b < S + PF; b++) prefetch(b);
            for (int b = 0; b < nb; b++) {
                const int st = b % S;
                if (PF > 0) prefetch(b + S + PF);
                mbar_wait(&done[st], (b / S) & 1);
                tma_store_3d(&maps.out, smem + (size_t)st * stage_bytes, col_of(b), h, c2x);
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

    // ---------------- consumer warp: lane = depth chunk
    const int d0 = K * lane;
    float P[K];
#pragma unroll
    for (int i = 0; i < K; i++) P[i] = 0.f;
    float pmax = 0.f;
    bool first = true;
    int slot = 0, phase = 0;

    for (int b = 0; b < nb; b++) {
        unsigned char *p = smem + (size_t)slot * stage_bytes;
        const int c0 = BW * (DIR == 0 ? b : nb - 1 - b);
        mbar_wait(&full[slot], phase);
        const float *gt = reinterpret_cast<const float *>(p + xbytes);
#pragma unroll 1
        for (int qq = 0; qq < BW / 4; qq++) {
            const int qi = (DIR == 0) ? qq : BW / 4 - 1 - qq;
            if (c0 + 4 * qi >= W) continue;               // columns past the image (zero-filled)
            float xq[K][4], gq[5][4];
#pragma unroll
            for (int i = 0; i < K; i++) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (FULL || d0 + i < D)
                    v = *reinterpret_cast<const float4 *>(p + tile_swz<BW>((unsigned)(((d0 + i) * BW + 4 * qi) * 4)));
                xq[i][0] = v.x; xq[i][1] = v.y; xq[i][2] = v.z; xq[i][3] = v.w;
            }
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const float4 v = *reinterpret_cast<const float4 *>(gt + k * BW + 4 * qi);
                gq[k][0] = v.x; gq[k][1] = v.y; gq[k][2] = v.z; gq[k][3] = v.w;
            }
#pragma unroll
            for (int ss = 0; ss < 4; ss++) {
                const int e = (DIR == 0) ? ss : 3 - ss;   // compile-time after unrolling
                float xc[K], w[5], A[K];
#pragma unroll
                for (int i = 0; i < K; i++) xc[i] = xq[i][e];
#pragma unroll
                for (int k = 0; k < 5; k++) w[k] = gq[k][e];
                if (ss == 0 && first) {
                    sga_first_step<K>(xc, w, A);
                    first = false;
                } else {
                    const float up = __shfl_up_sync(kFullMask, P[K - 1], 1);     // P[d0-1]
                    const float dn = __shfl_down_sync(kFullMask, P[0], 1);       // P[d0+K]
                    sga_next_step<K, FULL>(P, xc, w, up, dn, pmax, d0, D, A);
                }
                pmax = group_max<32>(FULL ? chunk_max<K>(A, 0, K) : chunk_max<K>(A, d0, D));
#pragma unroll
                for (int i = 0; i < K; i++) { P[i] = A[i]; xq[i][e] = A[i]; }
            }
#pragma unroll
            for (int i = 0; i < K; i++)
                if (FULL || d0 + i < D)
                    *reinterpret_cast<float4 *>(p + tile_swz<BW>((unsigned)(((d0 + i) * BW + 4 * qi) * 4))) =
                        make_float4(xq[i][0], xq[i][1], xq[i][2], xq[i][3]);
        }
        fence_proxy_async();
        mbar_arrive(&done[slot]);
        if (++slot == S) { slot = 0; phase ^= 1; }
    }
}

// four-value warp sum in 6 shuffles: on return lanes 0-7 hold sum(a), 8-15 sum(b), 16-23 sum(c),
// 24-31 sum(d) over all 32 lanes
__device__ __forceinline__ float warp_sum4(float a, float b, float c, float d, int lane)
{
    const bool hi = (lane & 16) != 0;
    const float s0 = hi ? a : c, s1 = hi ? b : d;         // what this lane gives away
    float k0 = hi ? c : a, k1 = hi ? d : b;               // what it keeps
    k0 += __shfl_xor_sync(kFullMask, s0, 16);
    k1 += __shfl_xor_sync(kFullMask, s1, 16);
    const bool h8 = (lane & 8) != 0;
    float v = h8 ? k1 : k0;
    v += __shfl_xor_sync(kFullMask, h8 ? k0 : k1, 8);
    v += __shfl_xor_sync(kFullMask, v, 4);
    v += __shfl_xor_sync(kFullMask, v, 2);
    v += __shfl_xor_sync(kFullMask, v, 1);
    return v;
}

// ---------------------------------------------------------------------------
// backward (reverse sweep) of one horizontal direction.  DIR 0 = right (walks the columns
// DEscending), 1 = left (ascending).  One CTA = one image row; THREE warps:
//
//   warp 1  producer: TMA tiles (BW columns, 1 row, D planes) of x, gradOut, A, mask, guidance into a
//           ring of S stages; stores the gradOut tile, which by then holds gradInput (plain store, or
//           reduce-add when `accumulate`)
//   warp 0  "T warp": the sequential part.  Per quad of four scan steps: mask-select gradOut -> T0,
//           depth arg-max of the aggregate (max-path target, max_idx), then the recurrence
//           T[t+1] -> T[t]; the sum over depth of T that the max-path term needs follows from the
//           recurrence itself,
//               sum_d T[d,t] = sum_d T0[d,t] + (w1+w2+w3+w4)(t+1) * sum_d T[d,t+1]
//                              - w2(t+1) * T[0,t+1] - w3(t+1) * T[D-1,t+1],
//           so only two broadcasts sit on the critical path.  T is written over the gradOut tile.
//   warp 2  "gradient warp", one quad behind (one named barrier per quad): reads T, x, A; the five
//           guidance-gradient dot products of four steps side by side (one packed four-value warp
//           sum each, written straight to global memory: 80/D bytes per voxel), then
//           gradInput = T * w0 (+ boundary terms) over the T values in place.
//
// Why this shape (profiles/r02_*): a reverse step is ~250 instructions for one warp issuing in order, and
// shared memory holds few rows per SM (13 bytes per voxel staged: 41 KB per 16-column stage at D = 192).
// One warp per row, two rows per SM ran at 3.2 TB/s, 45 % of the time waiting for tiles and the rest
// issue-bound; 32-column tiles (one row per SM) were slower still.  Two warps per row, 16-column
// tiles, two rows per SM: four working warps per SM.
// ---------------------------------------------------------------------------
template <int K, int BW, int DIR, bool FULL>
__global__ void __launch_bounds__(96)
sga_hscan_bwd_kernel(const __grid_constant__ HBwdMaps maps, float *__restrict__ gg,
                     int32_t *__restrict__ max_idx, int mask_id, int accumulate, int D, int H, int W,
                     int S, int PF)
{
    static_assert(BW == 16 || BW == 32, "tile width");
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = align1k(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long s = blockIdx.x / H;
    const int h = blockIdx.x - (int)(s * H);
    const HBwdPlan pl = hbwd_plan(D, BW);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * pl.stage_bytes);
    uint64_t *done = full + S;
    float *svbuf = reinterpret_cast<float *>(done + S);   // [2][4]: sum over depth of T, per quad parity
    const int c2x = (int)(s * D), c2g = (int)(s * 5);
    const int nb = (W + BW - 1) / BW;

    if (threadIdx.x == 0) {
        for (int i = 0; i < S; i++) { mbar_init(&full[i], 1); mbar_init(&done[i], 32); }
        fence_mbarrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    // processing order = reverse scan order: right walks the tiles from the last to the first
    if (warp == 1) {                                      // ---------------- producer
        if (lane == 0) {
            const unsigned tx = (unsigned)(3 * D * BW * 4 + D * BW + 5 * BW * 4);
            auto col_of = [&](int b) { return BW * (DIR == 0 ? nb - 1 - b : b); };
            auto issue = [&](int b) {
                const int st = b % S;
                unsigned char *p = smem + (size_t)st * pl.stage_bytes;
                mbar_arrive_expect_tx(&full[st], tx);
                tma_load_3d(p + pl.off_go, &maps.go, &full[st], col_of(b), h, c2x);
                tma_load_3d(p + pl.off_m, &maps.mask, &full[st], col_of(b), h, c2x);
                tma_load_3d(p + pl.off_a, &maps.a, &full[st], col_of(b), h, c2x);
                tma_load_3d(p + pl.off_g, &maps.g, &full[st], col_of(b), h, c2g);
                tma_load_3d(p + pl.off_x, &maps.x, &full[st], col_of(b), h, c2x);
            };
            auto prefetch = [&](int b) {
                if (b < nb) {
                    tma_prefetch_3d(&maps.x, col_of(b), h, c2x);
                    tma_prefetch_3d(&maps.go, col_of(b), h, c2x);
                    tma_prefetch_3d(&maps.a, col_of(b), h, c2x);
                    tma_prefetch_3d(&maps.mask, col_of(b), h, c2x);
                }
            };
            for (int b = 0; b < S && b < nb; b++) issue(b);
            for (int b = S; b < S + PF; b++) prefetch(b);
            for (int b = 0; b < nb; b++) {
                const int st = b % S;
                if (PF > 0) prefetch(b + S + PF);
                mbar_wait(&done[st], (b / S) & 1);
                unsigned char *p = smem + (size_t)st * pl.stage_bytes;
                if (accumulate) tma_reduce_add_3d(&maps.gi, p + pl.off_go, col_of(b), h, c2x);
                else tma_store_3d(&maps.gi, p + pl.off_go, col_of(b), h, c2x);
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

    const int d0 = K * lane;
    const long long HW = (long long)H * W;
    int slot = 0, phase = 0;
    int quad_parity = 0;

    if (warp == 0) {
        // ---------------- T warp: lane = depth chunk
        int32_t *mirow = max_idx ? max_idx + s * HW + (long long)h * W : nullptr;
        const int lane_last = (D - 1) / K, i_last = (D - 1) - K * lane_last;      // owner of depth D-1
        float Tn[K], wn[5];                               // T and guidance of position t+1
#pragma unroll
        for (int i = 0; i < K; i++) Tn[i] = 0.f;
#pragma unroll
        for (int k = 0; k < 5; k++) wn[k] = 0.f;
        float sum_tn = 0.f;                               // sum over depth of Tn
        bool has_next = false;

        for (int b = 0; b < nb; b++) {
            unsigned char *p = smem + (size_t)slot * pl.stage_bytes;
            const int c0 = BW * (DIR == 0 ? nb - 1 - b : b);
            mbar_wait(&full[slot], phase);
            const float *gt = reinterpret_cast<const float *>(p + pl.off_g);
            uint4 mq[K];                                  // mask bytes of 16 columns (four quads)
#pragma unroll 1
            for (int qq = 0; qq < BW / 4; qq++) {
                const int qi = (DIR == 0) ? BW / 4 - 1 - qq : qq;
                if (c0 + 4 * qi >= W) continue;           // columns past the image (both warps skip them)
                if ((DIR == 0) ? (qi & 3) == 3 : (qi & 3) == 0) {     // a new group of 16 columns (W % 16 == 0)
#pragma unroll
                    for (int i = 0; i < K; i++)
                        mq[i] = (FULL || d0 + i < D)
                                    ? *reinterpret_cast<const uint4 *>(p + pl.off_m + (unsigned)((d0 + i) * BW + 16 * (qi >> 2)))
                                    : make_uint4(~0u, ~0u, ~0u, ~0u);
                }
                float tq[K][4], aq[K][4], gq[5][4];
                const int wsel = qi & 3;
#pragma unroll
                for (int i = 0; i < K; i++) {
                    float4 vg = make_float4(0.f, 0.f, 0.f, 0.f), va = vg;
                    if (FULL || d0 + i < D) {
                        const unsigned o = tile_swz<BW>((unsigned)(((d0 + i) * BW + 4 * qi) * 4));
                        vg = *reinterpret_cast<const float4 *>(p + pl.off_go + o);
                        va = *reinterpret_cast<const float4 *>(p + pl.off_a + o);
                    }
                    const unsigned mw = wsel == 0 ? mq[i].x : wsel == 1 ? mq[i].y : wsel == 2 ? mq[i].z : mq[i].w;
                    aq[i][0] = va.x; aq[i][1] = va.y; aq[i][2] = va.z; aq[i][3] = va.w;
                    // get_temp_grad (:38-48): gradOut where this direction won the max
                    tq[i][0] = ((mw & 0xffu) == (unsigned)mask_id) ? vg.x : 0.f;
                    tq[i][1] = (((mw >> 8) & 0xffu) == (unsigned)mask_id) ? vg.y : 0.f;
                    tq[i][2] = (((mw >> 16) & 0xffu) == (unsigned)mask_id) ? vg.z : 0.f;
                    tq[i][3] = ((mw >> 24) == (unsigned)mask_id) ? vg.w : 0.f;
                }
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const float4 v = *reinterpret_cast<const float4 *>(gt + k * BW + 4 * qi);
                    gq[k][0] = v.x; gq[k][1] = v.y; gq[k][2] = v.z; gq[k][3] = v.w;
                }

                // ---- per step, independent of T: arg-max of the aggregate, sum of T0 ----------
                float st0[4];
                int idx[4];
                {
                    float best[4], amax[4];
                    int bi[4];
#pragma unroll
                    for (int ss = 0; ss < 4; ss++) {
                        const int e = (DIR == 0) ? 3 - ss : ss;
                        float bv = (FULL || d0 < D) ? aq[0][e] : -INFINITY;   // strict >: first maximum
                        int bd = d0;
                        float t0s = tq[0][e];
#pragma unroll
                        for (int i = 1; i < K; i++) {
                            if ((FULL || d0 + i < D) && aq[i][e] > bv) { bv = aq[i][e]; bd = d0 + i; }
                            t0s += tq[i][e];
                        }
                        best[ss] = bv; bi[ss] = bd; st0[ss] = t0s;
                    }
#pragma unroll
                    for (int ss = 0; ss < 4; ss++) amax[ss] = group_max<32>(best[ss]);
#pragma unroll
                    for (int ss = 0; ss < 4; ss++)
                        idx[ss] = __reduce_min_sync(kFullMask, best[ss] == amax[ss] ? bi[ss] : 0x7fffffff);
                    const float v = warp_sum4(st0[0], st0[1], st0[2], st0[3], lane);
#pragma unroll
                    for (int ss = 0; ss < 4; ss++) st0[ss] = __shfl_sync(kFullMask, v, 8 * ss);
                }
                if (mirow && lane < 4) {
                    const int ss = lane;
                    const int e = (DIR == 0) ? 3 - ss : ss;
                    mirow[c0 + 4 * qi + e] = ss == 0 ? idx[0] : ss == 1 ? idx[1] : ss == 2 ? idx[2] : idx[3];
                }

                // ---- the recurrence (t0 -> T in place in tq) -----------------------------------
                float sv[4];
#pragma unroll
                for (int ss = 0; ss < 4; ss++) {
                    const int e = (DIR == 0) ? 3 - ss : ss;
                    float scur = st0[ss];
                    if (ss > 0 || has_next) {
                        const float up = __shfl_up_sync(kFullMask, Tn[K - 1], 1);     // T[d0-1, t+1]
                        const float dn = __shfl_down_sync(kFullMask, Tn[0], 1);       // T[d0+K, t+1]
                        float tl = Tn[K - 1];
                        if (!FULL) {
#pragma unroll
                            for (int i = 0; i < K; i++)
                                if (i == i_last) tl = Tn[i];
                        }
                        const float t_first = __shfl_sync(kFullMask, Tn[0], 0);       // T[0, t+1]
                        const float t_last = __shfl_sync(kFullMask, tl, FULL ? 31 : lane_last);   // T[D-1, t+1]
                        const float inj = sum_tn * wn[4];                             // max-path term (:167-178)
#pragma unroll
                        for (int i = 0; i < K; i++) {
                            const int d = d0 + i;
                            const float tm = (i == 0) ? up : Tn[i == 0 ? 0 : i - 1];
                            const float tp = (i == K - 1) ? dn : Tn[i == K - 1 ? K - 1 : i + 1];
                            float v = tq[i][e];
                            v += Tn[i] * wn[1];
                            if (d + 1 < D) v += tp * wn[2];
                            if (d >= 1) v += tm * wn[3];
                            if (d == idx[ss]) v += inj;
                            tq[i][e] = (FULL || d < D) ? v : 0.f;
                        }
                        scur += sum_tn * (wn[1] + wn[2] + wn[3] + wn[4]) - wn[2] * t_first - wn[3] * t_last;
                    }
                    sv[ss] = scur;
                    sum_tn = scur;
#pragma unroll
                    for (int i = 0; i < K; i++) Tn[i] = tq[i][e];
#pragma unroll
                    for (int k = 0; k < 5; k++) wn[k] = gq[k][e];
                }
                has_next = true;

                // T over the gradOut tile, the four sums beside it; the gradient warp takes over
#pragma unroll
                for (int i = 0; i < K; i++)
                    if (FULL || d0 + i < D)
                        *reinterpret_cast<float4 *>(p + pl.off_go + tile_swz<BW>((unsigned)(((d0 + i) * BW + 4 * qi) * 4))) =
                            make_float4(tq[i][0], tq[i][1], tq[i][2], tq[i][3]);
                if (lane == 0)
                    *reinterpret_cast<float4 *>(svbuf + 4 * quad_parity) = make_float4(sv[0], sv[1], sv[2], sv[3]);
                quad_parity ^= 1;
                named_barrier(1, 64);
            }
            if (++slot == S) { slot = 0; phase ^= 1; }
        }
        return;
    }

    // ---------------- gradient warp: lane = depth chunk, one quad behind the T warp
    float *ggrow = gg + s * 5 * HW + (long long)h * W;    // + k * HW + column
    float Tin[K], xin[K];                                 // T and x of position t+1
#pragma unroll
    for (int i = 0; i < K; i++) { Tin[i] = 0.f; xin[i] = 0.f; }
    float sum_in = 0.f;
    bool has_in = false;
    int col_in = 0;

    for (int b = 0; b < nb; b++) {
        unsigned char *p = smem + (size_t)slot * pl.stage_bytes;
        const int c0 = BW * (DIR == 0 ? nb - 1 - b : b);
        mbar_wait(&full[slot], phase);
        const float *gt = reinterpret_cast<const float *>(p + pl.off_g);
#pragma unroll 1
        for (int qq = 0; qq < BW / 4; qq++) {
            const int qi = (DIR == 0) ? BW / 4 - 1 - qq : qq;
            if (c0 + 4 * qi >= W) continue;
            float xq[K][4], aq[K][4], w0q[4], w2q[4], w3q[4];
#pragma unroll
            for (int i = 0; i < K; i++) {
                float4 vx = make_float4(0.f, 0.f, 0.f, 0.f), va = vx;
                if (FULL || d0 + i < D) {
                    const unsigned o = tile_swz<BW>((unsigned)(((d0 + i) * BW + 4 * qi) * 4));
                    vx = *reinterpret_cast<const float4 *>(p + pl.off_x + o);
                    va = *reinterpret_cast<const float4 *>(p + pl.off_a + o);
                }
                xq[i][0] = vx.x; xq[i][1] = vx.y; xq[i][2] = vx.z; xq[i][3] = vx.w;
                aq[i][0] = va.x; aq[i][1] = va.y; aq[i][2] = va.z; aq[i][3] = va.w;
            }
            {
                const float4 a = *reinterpret_cast<const float4 *>(gt + 0 * BW + 4 * qi);
                const float4 c = *reinterpret_cast<const float4 *>(gt + 2 * BW + 4 * qi);
                const float4 d = *reinterpret_cast<const float4 *>(gt + 3 * BW + 4 * qi);
                w0q[0] = a.x; w0q[1] = a.y; w0q[2] = a.z; w0q[3] = a.w;
                w2q[0] = c.x; w2q[1] = c.y; w2q[2] = c.z; w2q[3] = c.w;
                w3q[0] = d.x; w3q[1] = d.y; w3q[2] = d.z; w3q[3] = d.w;
            }
            float amax[4];
            {
                float best[4];
#pragma unroll
                for (int ss = 0; ss < 4; ss++) {
                    const int e = (DIR == 0) ? 3 - ss : ss;
                    float bv = (FULL || d0 < D) ? aq[0][e] : -INFINITY;
#pragma unroll
                    for (int i = 1; i < K; i++)
                        if (FULL || d0 + i < D) bv = fmaxf(bv, aq[i][e]);
                    best[ss] = bv;
                }
#pragma unroll
                for (int ss = 0; ss < 4; ss++) amax[ss] = group_max<32>(best[ss]);
            }

            named_barrier(1, 64);                         // the T warp has written this quad
            float tq[K][4];
#pragma unroll
            for (int i = 0; i < K; i++) {
                float4 vt = make_float4(0.f, 0.f, 0.f, 0.f);
                if (FULL || d0 + i < D)
                    vt = *reinterpret_cast<const float4 *>(p + pl.off_go + tile_swz<BW>((unsigned)(((d0 + i) * BW + 4 * qi) * 4)));
                tq[i][0] = vt.x; tq[i][1] = vt.y; tq[i][2] = vt.z; tq[i][3] = vt.w;
            }
            const float4 svv = *reinterpret_cast<const float4 *>(svbuf + 4 * quad_parity);
            const float sv[4] = {svv.x, svv.y, svv.z, svv.w};
            quad_parity ^= 1;

            // ---- guidance gradients of the four steps (:210-281), then gradInput ----------------
            float s0[4], s1[4], s2[4], s3[4];
#pragma unroll
            for (int ss = 0; ss < 4; ss++) {
                const int e = (DIR == 0) ? 3 - ss : ss;
                const int ep = (DIR == 0) ? e + 1 : e - 1;     // the step processed before (position t+1)
                s0[ss] = 0.f; s1[ss] = 0.f; s2[ss] = 0.f; s3[ss] = 0.f;
                const float aup = __shfl_up_sync(kFullMask, aq[K - 1][e], 1);     // A[d0-1, t]
                const float adn = __shfl_down_sync(kFullMask, aq[0][e], 1);       // A[d0+K, t]
#pragma unroll
                for (int i = 0; i < K; i++) {
                    const int d = d0 + i;
                    const float tp_ = (ss == 0) ? Tin[i] : tq[i][ss == 0 ? e : ep];      // T[d, t+1]
                    const float xp_ = (ss == 0) ? xin[i] : xq[i][ss == 0 ? e : ep];      // x[d, t+1]
                    const float am = (i == 0) ? aup : aq[i == 0 ? 0 : i - 1][e];
                    const float apn = (i == K - 1) ? adn : aq[i == K - 1 ? K - 1 : i + 1][e];
                    s0[ss] += tq[i][e] * xq[i][e];
                    s1[ss] += tp_ * aq[i][e];
                    s2[ss] += tp_ * ((d >= 1) ? am : xp_);
                    s3[ss] += tp_ * ((d + 1 < D) ? apn : xp_);
                }
            }
            float tot[4];
#pragma unroll
            for (int ss = 0; ss < 4; ss++) tot[ss] = warp_sum4(s0[ss], s1[ss], s2[ss], s3[ss], lane);
#pragma unroll
            for (int ss = 0; ss < 4; ss++) {
                const int e = (DIR == 0) ? 3 - ss : ss;
                const int col = c0 + 4 * qi + e;
                const int colp = (ss == 0) ? col_in : ((DIR == 0) ? col + 1 : col - 1);
                const bool hasp = (ss > 0) || has_in;
                const float sp = (ss == 0) ? sum_in : sv[ss == 0 ? 0 : ss - 1];
                // lanes 0 / 8 / 16 / 24 hold the sums s0 (position t) / s1 / s2 / s3 (position t+1)
                if ((lane & 7) == 0) {
                    const int k = lane >> 3;
                    if (k == 0) ggrow[col] = tot[ss];
                    else if (hasp) ggrow[k * HW + colp] = tot[ss];
                }
                if (lane == 1 && hasp) ggrow[4 * HW + colp] = sp * amax[ss];
            }
            {
                const int e_last = (DIR == 0) ? 0 : 3;
#pragma unroll
                for (int i = 0; i < K; i++) { Tin[i] = tq[i][e_last]; xin[i] = xq[i][e_last]; }
                sum_in = sv[3];
                has_in = true;
                col_in = c0 + 4 * qi + e_last;
            }
            // gradInput (:164, :177, :200-207), in place over the T values
#pragma unroll
            for (int i = 0; i < K; i++) {
                const int d = d0 + i;
                if (FULL || d < D) {
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float t = tq[i][e];
                        float r = t * w0q[e];
                        if (d == 0) r += t * w2q[e];
                        if (d == D - 1) r += t * w3q[e];
                        v[e] = r;
                    }
                    *reinterpret_cast<float4 *>(p + pl.off_go + tile_swz<BW>((unsigned)(((d0 + i) * BW + 4 * qi) * 4))) =
                        make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
        fence_proxy_async();
        mbar_arrive(&done[slot]);
        if (++slot == S) { slot = 0; phase ^= 1; }
    }
    // scan position 0 has no predecessor: its guidance gradients 1..4 are zero (A.3 quirk)
    if (has_in && lane >= 1 && lane <= 4) ggrow[lane * HW + col_in] = 0.f;
}

// ---------------------------------------------------------------------------
// out = max of the four aggregates, mask = winning direction, ties keep the lower id
// (the reference's Max chain, GANet_kernel.cu:23-36, :964-994); all in the standard layout.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
merge4_kernel(const float4 *__restrict__ a0, const float4 *__restrict__ a1, const float4 *__restrict__ a2,
              const float4 *__restrict__ a3, float4 *__restrict__ out, uint32_t *__restrict__ mask,
              long long n4)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v0 = __ldcs(a0 + i), v1 = __ldcs(a1 + i), v2 = __ldcs(a2 + i), v3 = __ldcs(a3 + i);
        float o[4] = {v0.x, v0.y, v0.z, v0.w};
        const float b1[4] = {v1.x, v1.y, v1.z, v1.w}, b2[4] = {v2.x, v2.y, v2.z, v2.w},
                    b3[4] = {v3.x, v3.y, v3.z, v3.w};
        unsigned m = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            unsigned id = 0;
            if (o[e] < b1[e]) { o[e] = b1[e]; id = 1; }
            if (o[e] < b2[e]) { o[e] = b2[e]; id = 2; }
            if (o[e] < b3[e]) { o[e] = b3[e]; id = 3; }
            m |= id << (8 * e);
        }
        __stcs(out + i, make_float4(o[0], o[1], o[2], o[3]));
        __stcs(mask + i, m);
    }
}

}  // namespace ganet
