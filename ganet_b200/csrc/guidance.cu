// SGABlock prologue / epilogue: the split + L1-normalisation of the guidance that
// every SGABlock of the reference's models performs in front of its SGA call, fused.
//
// Reference being replaced: models/GANet_deep.py:264-268 (GANet11.py:246-250)
//     k1, k2, k3, k4 = torch.split(g, (C*5, C*5, C*5, C*5), 1)
//     k_i = F.normalize(k_i.view(N, C, 5, H, W), p=1, dim=2)            (x4)
// i.e. per direction a norm reduction, a clamp, an expand and a division -- four tensors read and
// three written per direction, a dozen launches -- plus the autograd graph of all that in backward.
// Here: ONE streaming pass reads the raw (N, 4*C*5, H, W) guidance once and writes the four
// (N, C, 5, H, W) weight tensors the SGA kernels consume; ONE pass turns the four guidance
// gradients SGA's backward returns into the gradient of the raw guidance.
//
// Arithmetic of F.normalize(p=1, eps=1e-12): y = x / max(sum_j |x_j|, eps).  The sum over the five
// weights follows the order of torch's strided reduction (four accumulators round-robin, then
// combined left to right): (((|x0| + |x4|) + |x1|) + |x2|) + |x3|, so that the weights -- and with them
// SGA's max / arg-max decisions -- are the unfused path's, bit for bit
// (tests/test_gpu_fused.py::test_guidance_prologue_matches_f_normalize).  Backward, for S > eps:
// dx_j = (g_j - sign(x_j) * sum_i g_i y_i) / S; for S <= eps the clamp cuts the norm's gradient: dx = g / eps.
#include "common.cuh"

namespace ganet {

constexpr float kNormEps = 1e-12f;

__device__ __forceinline__ float l1_of_five(float a0, float a1, float a2, float a3, float a4)
{
    float s = fabsf(a0) + fabsf(a4);
    s = s + fabsf(a1);
    s = s + fabsf(a2);
    return s + fabsf(a3);
}

// one thread = 4 adjacent pixels of one (n, direction, c): five float4 in, five float4 out
__global__ void __launch_bounds__(256)
guidance_fwd_kernel(const float4 *__restrict__ raw, float4 *__restrict__ g0, float4 *__restrict__ g1,
                    float4 *__restrict__ g2, float4 *__restrict__ g3, int C, long long hw4, long long total)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long p = i % hw4;                  // pixel quad
        const long long t = i / hw4;                  // (n * 4 + k) * C + c
        const int c = (int)(t % C);
        const long long nk = t / C;
        const int k = (int)(nk & 3);
        const long long n = nk >> 2;
        const float4 *src = raw + t * 5 * hw4 + p;
        float4 *dst = (k == 0 ? g0 : k == 1 ? g1 : k == 2 ? g2 : g3) + (n * C + c) * 5 * hw4 + p;
        float4 v[5];
#pragma unroll
        for (int j = 0; j < 5; j++) v[j] = __ldcs(src + j * hw4);
        float4 s;
        s.x = fmaxf(l1_of_five(v[0].x, v[1].x, v[2].x, v[3].x, v[4].x), kNormEps);
        s.y = fmaxf(l1_of_five(v[0].y, v[1].y, v[2].y, v[3].y, v[4].y), kNormEps);
        s.z = fmaxf(l1_of_five(v[0].z, v[1].z, v[2].z, v[3].z, v[4].z), kNormEps);
        s.w = fmaxf(l1_of_five(v[0].w, v[1].w, v[2].w, v[3].w, v[4].w), kNormEps);
#pragma unroll
        for (int j = 0; j < 5; j++)
            dst[j * hw4] = make_float4(__fdiv_rn(v[j].x, s.x), __fdiv_rn(v[j].y, s.y), __fdiv_rn(v[j].z, s.z),
                                       __fdiv_rn(v[j].w, s.w));
    }
}

__device__ __forceinline__ void norm_bwd5(const float (&x)[5], const float (&g)[5], float (&dx)[5])
{
    const float s = l1_of_five(x[0], x[1], x[2], x[3], x[4]);
    if (s > kNormEps) {
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < 5; j++) dot += g[j] * (x[j] / s);
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const float sg = x[j] > 0.f ? 1.f : (x[j] < 0.f ? -1.f : 0.f);
            dx[j] = (g[j] - sg * dot) / s;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 5; j++) dx[j] = g[j] / kNormEps;
    }
}

__global__ void __launch_bounds__(256)
guidance_bwd_kernel(const float4 *__restrict__ raw, const float4 *__restrict__ gg0,
                    const float4 *__restrict__ gg1, const float4 *__restrict__ gg2,
                    const float4 *__restrict__ gg3, float4 *__restrict__ grad_raw, int C, long long hw4,
                    long long total)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long p = i % hw4;
        const long long t = i / hw4;
        const int c = (int)(t % C);
        const long long nk = t / C;
        const int k = (int)(nk & 3);
        const long long n = nk >> 2;
        const float4 *src = raw + t * 5 * hw4 + p;
        const float4 *gsrc = (k == 0 ? gg0 : k == 1 ? gg1 : k == 2 ? gg2 : gg3) + (n * C + c) * 5 * hw4 + p;
        float4 xv[5], gv[5], out[5];
#pragma unroll
        for (int j = 0; j < 5; j++) { xv[j] = __ldcs(src + j * hw4); gv[j] = __ldcs(gsrc + j * hw4); }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            float x[5], g[5], dx[5];
#pragma unroll
            for (int j = 0; j < 5; j++) {
                x[j] = reinterpret_cast<const float *>(&xv[j])[e];
                g[j] = reinterpret_cast<const float *>(&gv[j])[e];
            }
            norm_bwd5(x, g, dx);
#pragma unroll
            for (int j = 0; j < 5; j++) reinterpret_cast<float *>(&out[j])[e] = dx[j];
        }
#pragma unroll
        for (int j = 0; j < 5; j++) __stcs(grad_raw + t * 5 * hw4 + p + j * hw4, out[j]);
    }
}

}  // namespace ganet

using namespace ganet;

static int guidance_dims_ok(int64_t N, int64_t C, int64_t H, int64_t W)
{
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return GANET_EINVAL;
    if ((H * W) % 4 != 0) return GANET_EUNSUPPORTED;          // 16-byte accesses along the pixel axis
    return GANET_OK;
}

GANET_API int ganet_sga_guidance_forward(const float *raw, float *g_down, float *g_up, float *g_right,
                                         float *g_left, int64_t N, int64_t C, int64_t H, int64_t W,
                                         ganet_stream_t stream)
{
    if (!raw || !g_down || !g_up || !g_right || !g_left) return GANET_EINVAL;
    const int rc = guidance_dims_ok(N, C, H, W);
    if (rc) return rc;
    const long long hw4 = H * W / 4, total = N * 4 * C * hw4;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    guidance_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const float4 *)raw, (float4 *)g_down, (float4 *)g_up, (float4 *)g_right, (float4 *)g_left, (int)C, hw4, total);
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

GANET_API int ganet_sga_guidance_backward(const float *raw, const float *gg_down, const float *gg_up,
                                          const float *gg_right, const float *gg_left, float *grad_raw,
                                          int64_t N, int64_t C, int64_t H, int64_t W, ganet_stream_t stream)
{
    if (!raw || !gg_down || !gg_up || !gg_right || !gg_left || !grad_raw) return GANET_EINVAL;
    const int rc = guidance_dims_ok(N, C, H, W);
    if (rc) return rc;
    const long long hw4 = H * W / 4, total = N * 4 * C * hw4;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    guidance_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const float4 *)raw, (const float4 *)gg_down, (const float4 *)gg_up, (const float4 *)gg_right,
        (const float4 *)gg_left, (float4 *)grad_raw, (int)C, hw4, total);
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}
