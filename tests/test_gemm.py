"""Bridge GEMM (tcgen05 / TMEM / TMA, csrc/gemm.cu) against PyTorch: C = act(A . B^T + bias) with bf16
operands and fp32 accumulation.  Tolerances: against an fp32 matmul of the SAME bf16-rounded operands
only the summation order differs (1e-3 relative to the output scale); against the fp32 layer
(reference RCNN_top, vgg16_rpn.py:56-61) the bf16 rounding of inputs and weights gives ~1e-2."""
import numpy as np
import pytest
import torch

gpu = pytest.mark.gpu


def _gemm(A, B, bias, relu, out_bf16):
    from nafae_b200 import _C
    M, K = A.shape
    N = B.shape[0]
    C = torch.full((M, N), float("nan"), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=A.device)
    flags = (1 if relu else 0) | (2 if out_bf16 else 0)
    st = _C.lib.nafae_gemm_bf16_tn(_C.ptr(A), _C.ptr(B), _C.ptr(bias), _C.ptr(C), M, N, K, flags, _C.stream())
    assert st == 1, _C.last_error()
    return C


@gpu
@pytest.mark.parametrize("M,N,K,relu,bias,out_bf16", [
    (128, 256, 64, False, False, False),     # one tile, one K step
    (128, 256, 512, False, True, False),
    (100, 40, 72, True, True, False),        # ragged everywhere: M, N tails and a partial K block
    (300, 512, 4096, True, True, True),      # fc7-like, bf16 output
    (800, 4096, 25088, True, True, False),   # fc6 at the benchmark size (R = 800 RoIs)
    (800, 512, 4096, False, True, False),
    (257, 96, 1000, False, False, False),
])
def test_gemm_matches_torch(M, N, K, relu, bias, out_bf16):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev).to(torch.bfloat16)
    B = (torch.randn(N, K, generator=g) / np.sqrt(K)).to(dev).to(torch.bfloat16)
    b = torch.randn(N, generator=g).to(dev) * 0.1 if bias else None
    got = _gemm(A, B, b, relu, out_bf16).float()
    torch.cuda.synchronize()
    want = A.float() @ B.float().t()
    if bias:
        want = want + b
    if relu:
        want = want.clamp_min(0)
    assert torch.isfinite(got).all()
    scale = float(want.abs().max())
    tol = (8e-3 if out_bf16 else 1e-3) * scale
    assert float((got - want).abs().max()) <= tol, (float((got - want).abs().max()), scale)


@gpu
def test_fc6_fc7_chain_close_to_fp32_layers():
    """RCNN_top as the reference runs it (fp32 Linear + ReLU twice) vs two tensor-core launches."""
    from nafae_b200.bridge import RCNNTop
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    fc6 = torch.nn.Linear(2048, 1024).to(dev)
    fc7 = torch.nn.Linear(1024, 512).to(dev)
    pooled = torch.randn(200, 2048, device=dev).clamp_min(0)
    want = torch.relu(fc7(torch.relu(fc6(pooled))))
    top = RCNNTop(fc6, fc7)
    got = top(pooled)
    assert got.dtype == torch.float32 and got.shape == want.shape
    err = float((got - want).abs().max()) / float(want.abs().max())
    assert err < 2e-2, err
    # a bf16 A operand (what RoIAlign can emit directly) goes in without a cast
    got2 = top(pooled.to(torch.bfloat16))
    assert torch.equal(got, got2)
