// The reference's native module surface -- a pybind11 module named `GANet` with the six functions
// of libs/GANet/src/GANet_cuda.cpp:67-75 -- implemented over the C ABI of libganet_b200.so.
//
// This is what a maintainer's rewritten GANet_cuda.cpp looks like, compiled and
// tested (tests/test_gpu_parity.py::test_pybind_module_matches_reference_extension): the reference's own
// libs/GANet/functions/GANet.py can import it unchanged (`from ..build.lib import GANet`, :3).
// Contract of the reference's Python layer (functions/GANet.py:10-48, :51-263): the CALLER allocates and
// zero-fills every output and scratch tensor, gradients are accumulated into them, every call returns 1,
// `mask` and `max_idx` are fp32 volumes, `temp_out` leaves forward holding the `left` aggregate
// (GANet_kernel.cu:989-994).  Differences: work goes to the CURRENT stream (the reference: legacy
// stream 0), and failures raise instead of being dropped.
//
// Only tensor plumbing lives here (torch is the device-memory layer); all arithmetic is behind ganet_b200.h.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "../../include/ganet_b200.h"

namespace {

void check_rc(int rc, const char *what)
{
    TORCH_CHECK(rc == 0, "ganet_b200: ", what, " failed (", rc, "): ", ganet_error_string(rc));
}

const float *fptr(const at::Tensor &t, const char *name)
{
    TORCH_CHECK(t.is_cuda(), "ganet_b200: ", name, " must be a CUDA tensor (there is no CPU path)");
    TORCH_CHECK(t.scalar_type() == at::kFloat, "ganet_b200: ", name, " must be float32");
    TORCH_CHECK(t.is_contiguous(), "ganet_b200: ", name, " must be contiguous");
    return t.data_ptr<float>();
}

float *fptr_mut(at::Tensor &t, const char *name) { return const_cast<float *>(fptr(t, name)); }

at::Tensor scratch(const at::Tensor &like, size_t bytes)
{
    return at::empty({(int64_t)(bytes ? bytes : 256)}, like.options().dtype(at::kByte));
}

// lga_cuda_forward / lga3d_cuda_forward (GANet_cuda.cpp:14-20, :31-37): output += LGA(input, filters)
int lga_forward_any(at::Tensor input, at::Tensor filters, at::Tensor output, const int radius)
{
    c10::cuda::CUDAGuard guard(input.device());
    auto st = at::cuda::getCurrentCUDAStream().stream();
    const int nd = (int)input.dim();
    TORCH_CHECK(nd == 4 || nd == 5, "ganet_b200: LGA input must be (N,D,H,W) or (N,C,D,H,W)");
    int64_t lead = 1;
    for (int i = 0; i < nd - 3; i++) lead *= input.size(i);
    auto y = at::empty_like(input);
    check_rc(ganet_lga_forward(fptr(input, "input"), fptr(filters, "filters"), y.data_ptr<float>(), lead,
                               input.size(nd - 3), input.size(nd - 2), input.size(nd - 1), radius, st),
             "ganet_lga_forward");
    output.add_(y);
    return 1;
}

// lga_cuda_backward / lga3d_cuda_backward (GANet_cuda.cpp:5-12, :22-29): gradFilters +=, gradInput
// overwritten; gradInput may alias input (functions/GANet.py:197), so it is produced in a fresh buffer
int lga_backward_any(at::Tensor input, at::Tensor filters, at::Tensor gradOutput, at::Tensor gradInput,
                     at::Tensor gradFilters, const int radius)
{
    c10::cuda::CUDAGuard guard(input.device());
    auto st = at::cuda::getCurrentCUDAStream().stream();
    const int nd = (int)input.dim();
    TORCH_CHECK(nd == 4 || nd == 5, "ganet_b200: LGA input must be (N,D,H,W) or (N,C,D,H,W)");
    int64_t lead = 1;
    for (int i = 0; i < nd - 3; i++) lead *= input.size(i);
    auto gx = at::empty_like(input);
    check_rc(ganet_lga_backward(fptr(input, "input"), fptr(filters, "filters"), fptr(gradOutput, "gradOutput"),
                                gx.data_ptr<float>(), fptr_mut(gradFilters, "gradFilters"), /*accumulate_f=*/1,
                                lead, input.size(nd - 3), input.size(nd - 2), input.size(nd - 1), radius, st),
             "ganet_lga_backward");
    gradInput.copy_(gx);
    return 1;
}

// sga_cuda_forward (GANet_cuda.cpp:39-48)
int sga_forward(at::Tensor input, at::Tensor g_down, at::Tensor g_up, at::Tensor g_right, at::Tensor g_left,
                at::Tensor temp_out, at::Tensor output, at::Tensor mask)
{
    c10::cuda::CUDAGuard guard(input.device());
    auto st = at::cuda::getCurrentCUDAStream().stream();
    TORCH_CHECK(input.dim() == 5, "ganet_b200: SGA input must be (N,C,D,H,W)");
    const int64_t N = input.size(0), C = input.size(1), D = input.size(2), H = input.size(3), W = input.size(4);
    auto m8 = at::empty(input.sizes(), input.options().dtype(at::kByte));
    auto ws = scratch(input, ganet_sga_forward_workspace_best(N, C, D, H, W));
    check_rc(ganet_sga_forward(fptr(input, "input"), fptr(g_down, "guidance_down"), fptr(g_up, "guidance_up"),
                               fptr(g_right, "guidance_right"), fptr(g_left, "guidance_left"),
                               fptr_mut(output, "output"), m8.data_ptr<uint8_t>(), /*aggregates=*/nullptr,
                               ws.data_ptr(), (size_t)ws.numel(), N, C, D, H, W, st),
             "ganet_sga_forward");
    mask.copy_(m8);                                         // the reference's fp32 mask volume
    check_rc(ganet_sga_direction(fptr(input, "input"), fptr(g_left, "guidance_left"), fptr_mut(temp_out, "temp_out"),
                                 3, N, C, D, H, W, st),
             "ganet_sga_direction");                        // temp_out = the `left` aggregate
    return 1;
}

// sga_cuda_backward (GANet_cuda.cpp:50-64): accumulates into the caller's zero-filled buffers; max_idx
// receives the depth arg-max of the `right` aggregate as float (GANet_kernel.cu:1119); temp_out / temp_grad
// are scratch upstream and are left untouched here
int sga_backward(at::Tensor input, at::Tensor g_down, at::Tensor g_up, at::Tensor g_right, at::Tensor g_left,
                 at::Tensor temp_out, at::Tensor mask, at::Tensor max_idx, at::Tensor gradOutput,
                 at::Tensor temp_grad, at::Tensor gradInput, at::Tensor grad_down, at::Tensor grad_up,
                 at::Tensor grad_right, at::Tensor grad_left)
{
    (void)temp_out; (void)temp_grad;
    c10::cuda::CUDAGuard guard(input.device());
    auto st = at::cuda::getCurrentCUDAStream().stream();
    TORCH_CHECK(input.dim() == 5, "ganet_b200: SGA input must be (N,C,D,H,W)");
    const int64_t N = input.size(0), C = input.size(1), D = input.size(2), H = input.size(3), W = input.size(4);
    auto m8 = mask.to(at::kByte);
    auto gi = at::empty_like(input);
    at::Tensor gg[4] = {at::empty_like(g_down), at::empty_like(g_up), at::empty_like(g_right), at::empty_like(g_left)};
    auto idx = at::empty({N, C, H, W}, input.options().dtype(at::kInt));
    auto ws = scratch(input, ganet_sga_backward_workspace_best(N, C, D, H, W));
    check_rc(ganet_sga_backward(fptr(input, "input"), fptr(g_down, "guidance_down"), fptr(g_up, "guidance_up"),
                                fptr(g_right, "guidance_right"), fptr(g_left, "guidance_left"),
                                m8.data_ptr<uint8_t>(), /*aggregates=*/nullptr, fptr(gradOutput, "gradOutput"),
                                gi.data_ptr<float>(), gg[0].data_ptr<float>(), gg[1].data_ptr<float>(),
                                gg[2].data_ptr<float>(), gg[3].data_ptr<float>(), idx.data_ptr<int32_t>(),
                                ws.data_ptr(), (size_t)ws.numel(), N, C, D, H, W, st),
             "ganet_sga_backward");
    gradInput.add_(gi);
    grad_down.add_(gg[0]);
    grad_up.add_(gg[1]);
    grad_right.add_(gg[2]);
    grad_left.add_(gg[3]);
    max_idx.copy_(idx);
    return 1;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
    m.def("lga_cuda_forward", &lga_forward_any, "LGA forward (CUDA)");
    m.def("lga_cuda_backward", &lga_backward_any, "LGA backward (CUDA)");
    m.def("lga3d_cuda_forward", &lga_forward_any, "LGA3D forward (CUDA)");
    m.def("lga3d_cuda_backward", &lga_backward_any, "LGA3D backward (CUDA)");
    m.def("sga_cuda_forward", &sga_forward, "SGA forward (CUDA)");
    m.def("sga_cuda_backward", &sga_backward, "SGA backward (CUDA)");
}
