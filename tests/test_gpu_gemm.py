"""The tcgen05/TMA GEMM kernel on its own (odevit_gemm_bf16 diagnostic entry) against a torch
fp32 matmul of the same bf16-rounded operands, and against the FFMA kernel.  Tolerance: fp32
accumulation of exact bf16 products -> only summation-order noise (1e-5 relative to |C|max)."""
import ctypes

import pytest
import torch

from odevit_b200 import _lib

pytestmark = pytest.mark.gpu


def _run(M, N, K, mn, engine, accumulate=False, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    ref = a.float() @ b.float().t()
    A = a.t().contiguous() if mn else a
    B = b.t().contiguous() if mn else b
    c = torch.full((M, N), 0.5, device="cuda") if accumulate else torch.empty(M, N, device="cuda")
    st = _lib.lib().odevit_gemm_bf16(M, N, K, 1 if mn else 0, A.data_ptr(), B.data_ptr(), c.data_ptr(),
                                     1 if accumulate else 0, engine,
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(st, "odevit_gemm_bf16")
    torch.cuda.synchronize()
    if accumulate:
        ref = ref + 0.5
    return c, ref


SHAPES = [(128, 128, 64), (128, 128, 256), (256, 384, 192), (414, 320, 64), (1000, 1344, 192),
          (13248, 3072, 768), (207, 768, 1536), (69, 192, 960), (64, 16, 8), (130, 144, 72)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_gemm_k_major(M, N, K):
    c, ref = _run(M, N, K, mn=False, engine=1)
    assert float((c - ref).abs().max() / ref.abs().max()) < 1e-5


@pytest.mark.parametrize("M,N,K", [(768, 1536, 13248), (3072, 768, 13248), (192, 960, 552), (64, 192, 57 * 8),
                                   (1344, 192, 4416), (128, 128, 64), (320, 64, 1000)])
@pytest.mark.parametrize("accumulate", [False, True])
def test_tc_gemm_mn_major_split_k(M, N, K, accumulate):
    c, ref = _run(M, N, K, mn=True, engine=1, accumulate=accumulate)
    assert float((c - ref).abs().max() / ref.abs().max()) < 2e-5


def test_tc_matches_ffma_kernel():
    c1, _ = _run(414, 320, 192, mn=False, engine=1)
    c0, _ = _run(414, 320, 192, mn=False, engine=0)
    assert float((c1 - c0).abs().max() / c0.abs().max()) < 1e-5


def test_tc_gemm_rejects_unsupported_shapes():
    a = torch.zeros(64, 20, device="cuda", dtype=torch.bfloat16)
    c = torch.zeros(64, 64, device="cuda")
    st = _lib.lib().odevit_gemm_bf16(64, 64, 20, 0, a.data_ptr(), a.data_ptr(), c.data_ptr(), 0, 1, None)
    assert st == -4
