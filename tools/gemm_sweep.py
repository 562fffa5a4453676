"""Developer tool (GPU): times the tcgen05 GEMM (odevit_gemm_bf16 diagnostic entry) over the hot path's
shapes for each tile configuration (ODEVIT_GEMM_TILE=<cg>x<bn>) and checks it against torch.matmul."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odevit_b200 import _lib  # noqa: E402

SHAPES = [  # (M, N, K, mn_major, accumulate)
    (13248, 3072, 768, 0, 0), (13248, 768, 1536, 0, 0), (13248, 1536, 768, 0, 0), (13248, 768, 3072, 0, 0),
    (3072, 768, 13248, 1, 1), (768, 1536, 13248, 1, 1), (35328, 1344, 192, 0, 0), (35328, 192, 960, 0, 0),
    (8192, 8192, 8192, 0, 0),
]
TILES = ["1x128", "2x128", "2x192", "2x256", ""]


def main():
    L = _lib.lib()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for (M, N, K, mn, acc) in SHAPES:
        g = torch.Generator(device="cuda").manual_seed(0)
        a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
        b = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
        ref = a.float() @ b.float().t() if M * N <= 13248 * 3072 else None
        A = a.t().contiguous() if mn else a
        B = b.t().contiguous() if mn else b
        c = torch.zeros(M, N, device="cuda")
        for tile in TILES:
            if mn and tile == "2x192":
                continue
            os.environ["ODEVIT_GEMM_TILE"] = tile
            c.zero_()
            st = L.odevit_gemm_bf16(M, N, K, mn, A.data_ptr(), B.data_ptr(), c.data_ptr(), acc, 1, stream)
            if st != 0:
                print(f"M={M} N={N} K={K} mn={mn} tile={tile or 'auto'}: status {st} {L.odevit_last_error_string().decode()}")
                continue
            torch.cuda.synchronize()
            err = float((c - ref).abs().max() / ref.abs().max()) if ref is not None else float("nan")
            for _ in range(3):
                L.odevit_gemm_bf16(M, N, K, mn, A.data_ptr(), B.data_ptr(), c.data_ptr(), acc, 1, stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for _ in range(n):
                L.odevit_gemm_bf16(M, N, K, mn, A.data_ptr(), B.data_ptr(), c.data_ptr(), acc, 1, stream)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / n * 1e3
            print(f"M={M:6d} N={N:5d} K={K:6d} mn={mn} tile={tile or 'auto':6s} err={err:.1e} {us:8.1f} us "
                  f"{2.0 * M * N * K / us / 1e6:7.1f} TF/s", flush=True)
        if (M, N, K) == (8192, 8192, 8192):
            bt = b.t().contiguous()
            for _ in range(3):
                torch.matmul(a, bt)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                torch.matmul(a, bt)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 10 * 1e3
            print(f"cuBLAS bf16 8192^3: {us:.1f} us {2.0 * M * N * K / us / 1e6:.1f} TF/s")
    os.environ["ODEVIT_GEMM_TILE"] = ""


if __name__ == "__main__":
    main()
