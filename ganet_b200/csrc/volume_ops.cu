// GetCostVolume and DisparityRegression for sm_100a.
//
// Reference being replaced: libs/GANet/modules/GANet.py
//   GetCostVolume.forward       :119-134  (a Python loop of 2*(maxdisp+1)
//                               strided slice copies into a zero-filled volume)
//   DisparityRegression.forward :142-148  (materialises a repeated index volume
//                               and a product volume, then torch.sum over dim 1)
// Both are pure HBM streaming: one pass, every byte touched once.
#include "common.cuh"

namespace ganet {

constexpr int kThreads = 256;

// cost[n, c,   i, h, w] = w >= i ? x[n, c, h, w]     : 0
// cost[n, C+c, i, h, w] = w >= i ? y[n, c, h, w - i] : 0
// One CTA row per (n, c2, h): the source row sits in registers/L1 and is
// replayed Dm times; stores are coalesced along w.
__global__ void __launch_bounds__(kThreads)
cost_volume_fwd_kernel(const float *__restrict__ x, const float *__restrict__ y,
                       float *__restrict__ cost, int C, int Dm, int H, int W)
{
    const int h = blockIdx.x;
    const int c2 = blockIdx.y;
    const long long n = blockIdx.z;
    const bool left = c2 < C;
    const float *src = (left ? x : y) + ((n * C + (left ? c2 : c2 - C)) * (long long)H + h) * W;
    float *dst = cost + (((n * 2 * C + c2) * (long long)Dm) * H + h) * W;
    const long long plane = (long long)H * W;
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
        const float xv = left ? ld_nc(src + w) : 0.f;
        for (int i = 0; i < Dm; i++) {
            float v = 0.f;
            if (w >= i) v = left ? xv : ld_nc(src + w - i);
            dst[i * plane + w] = v;
        }
    }
}

// gx[n,c,h,w] = sum_{i <= w, i < Dm} gcost[n,c,i,h,w]
// gy[n,c,h,w] = sum_{i < Dm, w+i < W} gcost[n,C+c,i,h,w+i]
__global__ void __launch_bounds__(kThreads)
cost_volume_bwd_kernel(const float *__restrict__ gcost, float *__restrict__ gx,
                       float *__restrict__ gy, int C, int Dm, int H, int W)
{
    const int h = blockIdx.x;
    const int c2 = blockIdx.y;
    const long long n = blockIdx.z;
    const bool left = c2 < C;
    float *dst = (left ? gx : gy) + ((n * C + (left ? c2 : c2 - C)) * (long long)H + h) * W;
    const float *src = gcost + (((n * 2 * C + c2) * (long long)Dm) * H + h) * W;
    const long long plane = (long long)H * W;
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
        float acc = 0.f;
        if (left) {
            const int lim = min(Dm, w + 1);
            for (int i = 0; i < lim; i++) acc += ld_nc(src + i * plane + w);
        } else {
            const int lim = min(Dm, W - w);
            for (int i = 0; i < lim; i++) acc += ld_nc(src + i * plane + w + i);
        }
        dst[w] = acc;
    }
}

// disp[n,h,w] = sum_d d * p[n,d,h,w]; four independent partial sums per thread
__global__ void __launch_bounds__(kThreads)
disp_regression_fwd_kernel(const float *__restrict__ p, float *__restrict__ disp, int Dm,
                           long long HW)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = blockIdx.y;
    if (i >= HW) return;
    const float *src = p + n * Dm * HW + i;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int d = 0;
    for (; d + 3 < Dm; d += 4) {
        a0 = fmaf(ld_nc(src + (long long)d * HW), (float)d, a0);
        a1 = fmaf(ld_nc(src + (long long)(d + 1) * HW), (float)(d + 1), a1);
        a2 = fmaf(ld_nc(src + (long long)(d + 2) * HW), (float)(d + 2), a2);
        a3 = fmaf(ld_nc(src + (long long)(d + 3) * HW), (float)(d + 3), a3);
    }
    for (; d < Dm; d++) a0 = fmaf(ld_nc(src + (long long)d * HW), (float)d, a0);
    disp[n * HW + i] = (a0 + a1) + (a2 + a3);
}

__global__ void __launch_bounds__(kThreads)
disp_regression_bwd_kernel(const float *__restrict__ gdisp, float *__restrict__ gp, int Dm,
                           long long HW)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = blockIdx.y;
    if (i >= HW) return;
    const float g = ld_nc(gdisp + n * HW + i);
    float *dst = gp + n * Dm * HW + i;
    for (int d = 0; d < Dm; d++) dst[(long long)d * HW] = g * (float)d;
}

// DispAgg tail (partial): F.normalize(x, p=1, dim=1) followed by DisparityRegression
// (models/GANet_deep.py:245-247) in one pass over x:
//     disp = sum_d d * x_d / max(sum_d |x_d|, 1e-12)
// The unfused tail reads x, writes the normalised volume, reads it again (and in backward walks the autograd
// graph of the normalisation); here x is read once each way and nothing of its size is written in forward.
__global__ void __launch_bounds__(kThreads)
norm_regression_fwd_kernel(const float *__restrict__ x, float *__restrict__ disp, float *__restrict__ norm,
                           int Dm, long long HW)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = blockIdx.y;
    if (i >= HW) return;
    const float *src = x + n * Dm * HW + i;
    float a0 = 0.f, a1 = 0.f, s0 = 0.f, s1 = 0.f;
    int d = 0;
    for (; d + 1 < Dm; d += 2) {
        const float v0 = ld_nc(src + (long long)d * HW), v1 = ld_nc(src + (long long)(d + 1) * HW);
        a0 = fmaf(v0, (float)d, a0); s0 += fabsf(v0);
        a1 = fmaf(v1, (float)(d + 1), a1); s1 += fabsf(v1);
    }
    for (; d < Dm; d++) {
        const float v0 = ld_nc(src + (long long)d * HW);
        a0 = fmaf(v0, (float)d, a0); s0 += fabsf(v0);
    }
    const float s = fmaxf(s0 + s1, 1e-12f);
    disp[n * HW + i] = (a0 + a1) / s;
    norm[n * HW + i] = s0 + s1;                       // unclamped: backward needs to know whether the clamp cut
}

// gx_d = g * (d - sign(x_d) * disp) / S for S > eps (disp = sum_k k x_k / S), g * d / eps otherwise
__global__ void __launch_bounds__(kThreads)
norm_regression_bwd_kernel(const float *__restrict__ x, const float *__restrict__ disp,
                           const float *__restrict__ norm, const float *__restrict__ gdisp,
                           float *__restrict__ gx, int Dm, long long HW)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = blockIdx.y;
    if (i >= HW) return;
    const float g = ld_nc(gdisp + n * HW + i), dv = ld_nc(disp + n * HW + i), s = ld_nc(norm + n * HW + i);
    const float *src = x + n * Dm * HW + i;
    float *dst = gx + n * Dm * HW + i;
    if (s > 1e-12f) {
        const float gs = g / s;
        for (int d = 0; d < Dm; d++) {
            const float v = ld_nc(src + (long long)d * HW);
            const float sg = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);
            dst[(long long)d * HW] = gs * ((float)d - sg * dv);
        }
    } else {
        const float ge = g / 1e-12f;
        for (int d = 0; d < Dm; d++) dst[(long long)d * HW] = ge * (float)d;
    }
}

}  // namespace ganet

using namespace ganet;

GANET_API int ganet_norm_disp_regression_forward(const float *x, float *disp, float *norm, int64_t N,
                                                 int64_t Dm, int64_t H, int64_t W, ganet_stream_t stream)
{
    if (!x || !disp || !norm || N <= 0 || Dm <= 0 || H <= 0 || W <= 0) return GANET_EINVAL;
    if (N > 65535) return GANET_EUNSUPPORTED;
    const long long HW = H * W;
    dim3 grid((unsigned)((HW + kThreads - 1) / kThreads), (unsigned)N);
    norm_regression_fwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, disp, norm, (int)Dm, HW);
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

GANET_API int ganet_norm_disp_regression_backward(const float *x, const float *disp, const float *norm,
                                                  const float *grad_disp, float *grad_x, int64_t N,
                                                  int64_t Dm, int64_t H, int64_t W, ganet_stream_t stream)
{
    if (!x || !disp || !norm || !grad_disp || !grad_x || N <= 0 || Dm <= 0 || H <= 0 || W <= 0) return GANET_EINVAL;
    if (N > 65535) return GANET_EUNSUPPORTED;
    const long long HW = H * W;
    dim3 grid((unsigned)((HW + kThreads - 1) / kThreads), (unsigned)N);
    norm_regression_bwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, disp, norm, grad_disp, grad_x,
                                                                            (int)Dm, HW);
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

GANET_API int ganet_cost_volume_forward(const float *x, const float *y, float *cost, int64_t N,
                                        int64_t C, int64_t Dm, int64_t H, int64_t W,
                                        ganet_stream_t stream)
{
    if (!x || !y || !cost || N <= 0 || C <= 0 || Dm <= 0 || H <= 0 || W <= 0) return GANET_EINVAL;
    if (2 * C > 65535 || N > 65535) return GANET_EUNSUPPORTED;
    dim3 grid((unsigned)H, (unsigned)(2 * C), (unsigned)N);
    cost_volume_fwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, y, cost, (int)C, (int)Dm,
                                                                        (int)H, (int)W);
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

GANET_API int ganet_cost_volume_backward(const float *grad_cost, float *grad_x, float *grad_y,
                                         int64_t N, int64_t C, int64_t Dm, int64_t H, int64_t W,
                                         ganet_stream_t stream)
{
    if (!grad_cost || !grad_x || !grad_y || N <= 0 || C <= 0 || Dm <= 0 || H <= 0 || W <= 0)
        return GANET_EINVAL;
    if (2 * C > 65535 || N > 65535) return GANET_EUNSUPPORTED;
    dim3 grid((unsigned)H, (unsigned)(2 * C), (unsigned)N);
    cost_volume_bwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(grad_cost, grad_x, grad_y,
                                                                        (int)C, (int)Dm, (int)H,
                                                                        (int)W);
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

GANET_API int ganet_disp_regression_forward(const float *p, float *disp, int64_t N, int64_t Dm,
                                            int64_t H, int64_t W, ganet_stream_t stream)
{
    if (!p || !disp || N <= 0 || Dm <= 0 || H <= 0 || W <= 0) return GANET_EINVAL;
    if (N > 65535) return GANET_EUNSUPPORTED;
    const long long HW = H * W;
    dim3 grid((unsigned)((HW + kThreads - 1) / kThreads), (unsigned)N);
    disp_regression_fwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(p, disp, (int)Dm, HW);
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

GANET_API int ganet_disp_regression_backward(const float *grad_disp, float *grad_p, int64_t N,
                                             int64_t Dm, int64_t H, int64_t W,
                                             ganet_stream_t stream)
{
    if (!grad_disp || !grad_p || N <= 0 || Dm <= 0 || H <= 0 || W <= 0) return GANET_EINVAL;
    if (N > 65535) return GANET_EUNSUPPORTED;
    const long long HW = H * W;
    dim3 grid((unsigned)((HW + kThreads - 1) / kThreads), (unsigned)N);
    disp_regression_bwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(grad_disp, grad_p,
                                                                            (int)Dm, HW);
    GANET_RETURN_IF_LAUNCH_FAILED();
    return GANET_OK;
}

GANET_API int ganet_abi_version(void) { return GANET_B200_ABI_VERSION; }

GANET_API const char *ganet_error_string(int code)
{
    switch (code) {
    case GANET_OK: return "ok";
    case GANET_EINVAL: return "invalid argument (null pointer, aliasing or non-positive dimension)";
    case GANET_EUNSUPPORTED: return "shape outside the compiled kernel range";
    case GANET_EWORKSPACE: return "workspace too small";
    case GANET_EALIGN: return "pointer not sufficiently aligned";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown ganet error";
    }
}
