"""The C-ABI boundary (include/odevit.h <-> odevit_b200/csrc/libodevit.so).

CPU part: the library loads without a GPU, exports every symbol the header declares, and its
argument validation / workspace arithmetic (no kernel launches) behaves.  GPU part: error
behaviour of the compute entry points (host pointers, short workspaces, bad grids)."""
import ctypes
import os
import re

import pytest
import torch

from odevit_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "odevit.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(odevit_[a-z0-9_]+)\s*\(", src)))


def _desc(**kw):
    d = _lib.Desc()
    d.abi_version = _lib.ABI_VERSION
    d.batch, d.tokens, d.dim, d.heads, d.hidden = 2, 19, 64, 2, 128
    d.variant, d.precision, d.scaler = _lib.FIELD_PARALLEL, _lib.FP32, 12.0
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _declared_functions()
    assert set(declared) == set(_lib.DECLARED_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.odevit_abi_version() == _lib.ABI_VERSION
    assert b"sm_100a" in L.odevit_build_info()


def test_struct_layouts_match_header():
    # 8 int32 + float + 7 int32 reserved = 64 bytes; 19 + 5 pointers; 15 + 9 pointers
    assert ctypes.sizeof(_lib.Desc) == 64
    assert ctypes.sizeof(_lib.Weights) == 24 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_lib.WeightGrads) == 24 * ctypes.sizeof(ctypes.c_void_p)


def test_workspace_bytes_and_validation():
    L = _lib.lib()
    d = _desc()
    sizes = {}
    for kind in (_lib.WS_FIELD, _lib.WS_SOLVE_FWD, _lib.WS_SOLVE_BWD):
        for m in (_lib.EULER, _lib.MIDPOINT, _lib.RK4_38):
            n = L.odevit_workspace_bytes(ctypes.byref(d), kind, m)
            assert n > 0
            sizes[(kind, m)] = n
    assert sizes[(_lib.WS_SOLVE_BWD, _lib.RK4_38)] > sizes[(_lib.WS_SOLVE_BWD, _lib.EULER)]
    assert sizes[(_lib.WS_SOLVE_BWD, _lib.EULER)] > sizes[(_lib.WS_SOLVE_FWD, _lib.EULER)]
    big = _desc(batch=64)
    assert L.odevit_workspace_bytes(ctypes.byref(big), _lib.WS_SOLVE_FWD, _lib.EULER) > sizes[(_lib.WS_SOLVE_FWD, _lib.EULER)]
    for bad in (_desc(abi_version=99), _desc(heads=3), _desc(batch=0), _desc(precision=7), _desc(variant=9)):
        assert L.odevit_workspace_bytes(ctypes.byref(bad), _lib.WS_FIELD, _lib.EULER) == 0
        assert L.odevit_last_error_string() != b"ok"
    assert L.odevit_workspace_bytes(ctypes.byref(d), _lib.WS_SOLVE_FWD, 17) == 0
    assert b"method" in L.odevit_last_error_string()


def test_null_arguments_are_errors_not_crashes():
    L = _lib.lib()
    d = _desc()
    w = _lib.Weights()
    st = L.odevit_field_fwd(ctypes.byref(d), ctypes.byref(w), None, None, None, None, 0, None)
    assert st == -1          # ODEVIT_ERR_INVALID_ARG: weights missing
    st = L.odevit_field_fwd(None, ctypes.byref(w), None, None, None, None, 0, None)
    assert st == -1


def test_cpu_tensor_is_rejected_by_the_module_surface():
    import odevit_b200 as ob
    f = ob.ViT_ODEFunc(dim=64, num_heads=2, mlp_ratio=2.0, l2_attention=False)
    with pytest.raises(ob.OdevitError, match="no CPU fallback"):
        f(torch.tensor(0.0), torch.randn(1, 5, 64))
    with pytest.raises(TypeError):
        ob.odeint(lambda t, y: y, torch.randn(1, 5, 64), torch.linspace(0, 1, 3))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libodevit.so")
    with pytest.raises(_lib.OdevitError, match="not built"):
        _lib.lib()


# ---- GPU: error behaviour of the compute entry points --------------------------------------------

def _gpu_field_args():
    import odevit_b200 as ob
    f = ob.ViT_ODEFunc(dim=64, num_heads=2, mlp_ratio=2.0, l2_attention=False).cuda()
    w, keep = ob.ops._pack_weights(*zip(*[(k, v) for k, v in f.block.field_weights().items() if v is not None]))
    return f, w, keep


@pytest.mark.gpu
def test_host_pointer_and_short_workspace_are_errors():
    L = _lib.lib()
    f, w, keep = _gpu_field_args()
    d = _desc()
    x = torch.randn(2, 19, 64, device="cuda")
    dx = torch.empty_like(x)
    need = L.odevit_workspace_bytes(ctypes.byref(d), _lib.WS_FIELD, _lib.EULER)
    ws = torch.empty(need + 2048, dtype=torch.uint8, device="cuda")
    base = (ws.data_ptr() + 1023) & ~1023
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ok = L.odevit_field_fwd(ctypes.byref(d), ctypes.byref(w), x.data_ptr(), dx.data_ptr(), None, base, need, stream)
    assert ok == 0, L.odevit_last_error_string()
    host = torch.randn(2, 19, 64)
    st = L.odevit_field_fwd(ctypes.byref(d), ctypes.byref(w), host.data_ptr(), dx.data_ptr(), None, base, need, stream)
    assert st == -5 and b"not device memory" in L.odevit_last_error_string() or st == -5
    st = L.odevit_field_fwd(ctypes.byref(d), ctypes.byref(w), x.data_ptr(), dx.data_ptr(), None, base, 4096, stream)
    assert st == -2
    st = L.odevit_field_fwd(ctypes.byref(d), ctypes.byref(w), x.data_ptr(), dx.data_ptr(), None, base + 4, need, stream)
    assert st == -2
    st = L.odevit_field_fwd(ctypes.byref(d), ctypes.byref(w), x.data_ptr(), x.data_ptr(), None, base, need, stream)
    assert st == -1
    torch.cuda.synchronize()


@pytest.mark.gpu
def test_non_monotonic_grid_is_rejected():
    import odevit_b200 as ob
    f, _, _ = _gpu_field_args()
    x = torch.randn(2, 19, 64, device="cuda")
    with pytest.raises(ob.OdevitError, match="strictly"):
        ob.odeint(f, x, torch.tensor([0.0, 0.5, 0.4]), method="euler")
    with pytest.raises(ValueError):
        ob.odeint(f, x, torch.linspace(0, 1, 3), method="dopri5")
    # decreasing grids are legal (torchdiffeq integrates backwards in time)
    s = ob.odeint(f, x, torch.tensor([1.0, 0.5, 0.0]), method="rk4")
    assert s.shape == (3, 2, 19, 64) and torch.isfinite(s).all()


@pytest.mark.gpu
def test_launch_counter_counts_kernels():
    import odevit_b200 as ob
    f, _, _ = _gpu_field_args()
    x = torch.randn(2, 19, 64, device="cuda")
    ob.reset_launch_count()
    with torch.no_grad():
        ob.odeint(f, x, torch.linspace(0, 1, 4), method="euler", record_attention=False)
    n = ob.launch_count()
    assert n >= 3          # at least one launch per step
    ob.reset_launch_count()
    assert ob.launch_count() == 0
