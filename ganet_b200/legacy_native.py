"""The reference's native module surface, re-implemented over the C ABI.

The reference's Python layer binds a pybind11 module named `GANet` with six
functions (libs/GANet/src/GANet_cuda.cpp:67-75) and the buffer contract of
libs/GANet/functions/GANet.py: the CALLER allocates and zero-fills every output
and scratch tensor, results are accumulated into them, every call returns 1.
This module offers the same six names with the same argument order and buffer
semantics, so the reference's own functions/GANet.py can run unmodified on the
B200 kernels (`from ..build.lib import GANet`, functions/GANet.py:3) -- the
cheapest A/B harness: swap one module.

New code should call ganet_b200.ops / the C ABI directly: the legacy contract
costs extra copies (fp32 mask volume, temp_out, accumulate-into-zeros).
"""
import torch

from . import ops


def sga_cuda_forward(input, guidance_down, guidance_up, guidance_right, guidance_left,
                     temp_out, output, mask):
    """GANet_cuda.cpp:39-48.  output = max over the four aggregates, mask = winning
    direction as float, temp_out = the `left` aggregate (GANet_kernel.cu:989-994)."""
    out, m = ops.sga_forward(input, guidance_down, guidance_up, guidance_right, guidance_left)
    output.copy_(out)
    mask.copy_(m)
    temp_out.copy_(ops.sga_direction(input, guidance_left, 3))
    return 1


def sga_cuda_backward(input, guidance_down, guidance_up, guidance_right, guidance_left,
                      temp_out, mask, max_idx, gradOutput, temp_grad, gradInput,
                      grad_down, grad_up, grad_right, grad_left):
    """GANet_cuda.cpp:50-64.  Gradients are accumulated (+=) into the caller's
    zero-filled buffers; max_idx receives the depth arg-max of the `right`
    aggregate as float (GANet_kernel.cu:1119); temp_out / temp_grad are scratch
    upstream and are left untouched here."""
    gi, gg, idx = ops.sga_backward(input, guidance_down, guidance_up, guidance_right,
                                   guidance_left, mask.to(torch.uint8), gradOutput,
                                   want_max_idx=True)
    gradInput.add_(gi)
    for dst, src in zip((grad_down, grad_up, grad_right, grad_left), gg):
        dst.add_(src)
    max_idx.copy_(idx)
    return 1


def lga_cuda_forward(input, filters, output, radius):
    """GANet_cuda.cpp:14-20: output += LGA(input, filters)."""
    output.add_(ops.lga_forward(input, filters, int(radius)))
    return 1


def lga_cuda_backward(input, filters, gradOutput, gradInput, gradFilters, radius):
    """GANet_cuda.cpp:5-12: gradFilters +=, gradInput overwritten.  gradInput may
    alias input (Lga2Function passes temp_out as both, functions/GANet.py:197)."""
    gx, _ = ops.lga_backward(input, filters, gradOutput, int(radius), grad_f=gradFilters)
    gradInput.copy_(gx)
    return 1


# the 5-D variants run the same kernels with (N, C) folded into the batch
# (GANet_kernel.cu:1324-1364)
lga3d_cuda_forward = lga_cuda_forward
lga3d_cuda_backward = lga_cuda_backward
