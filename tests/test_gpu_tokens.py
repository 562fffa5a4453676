"""Row (f)1: token assembly in the patch GEMM's epilogue and head + cross-entropy in one launch, against the plain
PyTorch composition of the reference (ode_transformer_gpt.py:148-182, :588-589, :626)."""
import pytest
import torch
import torch.nn.functional as F

from _util import max_rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dist,R,pos_reg,img,patch,D", [(False, 4, False, 32, 4, 192), (True, 4, True, 32, 4, 192),
                                                       (False, 10, True, 224, 16, 768), (True, 0, False, 32, 8, 128),
                                                       (True, 10, False, 224, 16, 768)])
def test_token_assembly_matches_reference_composition(dist, R, pos_reg, img, patch, D):
    import odevit_b200 as ob
    torch.manual_seed(0)
    pe = ob.vit_ode.PatchEmbed(img, patch, 3, D, dist, register_tokens=R, pos_embed_register_tokens=pos_reg).cuda()
    with torch.no_grad():
        pe.pos_embed.normal_(0, 0.5)
        pe.cls_token.normal_(0, 0.5)
    pe.precision = "bf16"
    x = torch.randn(3, 3, img, img, device="cuda", requires_grad=True)
    w = torch.randn(3, pe.num_patches + 1 + int(dist) + R, D, device="cuda")

    def run(fused):
        pe.fused_assembly = fused
        pe.zero_grad(set_to_none=True)
        x.grad = None
        t = pe(x)
        (t * w).sum().backward()
        return t.detach().clone(), x.grad.clone(), {k: p.grad.clone() for k, p in pe.named_parameters()
                                                    if p.grad is not None and p.numel() > 0}

    t1, gx1, g1 = run(True)
    t0, gx0, g0 = run(False)          # im2col GEMM + torch.cat + pos add (the path before the fusion)
    assert t1.shape == t0.shape
    assert max_rel(t1, t0) < 1e-5
    assert max_rel(gx1, gx0) < 1e-3
    assert set(g1) == set(g0)
    for k in g0:
        assert max_rel(g1[k], g0[k]) < 1e-3, k


@pytest.mark.parametrize("B,N,D,C,bias", [(64, 207, 768, 100, True), (5, 69, 192, 10, True), (3, 7, 64, 1000, False)])
def test_head_ce_matches_linear_and_cross_entropy(B, N, D, C, bias):
    from odevit_b200 import ops
    g = torch.Generator().manual_seed(B)
    final = torch.randn(B, N, D, generator=g).cuda().requires_grad_(True)
    W = (torch.randn(C, D, generator=g) * 0.1).cuda().requires_grad_(True)
    b = (torch.randn(C, generator=g) * 0.1).cuda().requires_grad_(True) if bias else None
    labels = torch.randint(0, C, (B,), generator=g).cuda()
    wl = torch.randn(B, C, generator=g).cuda()
    logits, loss = ops.head_ce(final, W, b, labels, 0.05)
    (loss * 1.7 + (logits * wl).sum()).backward()
    fr, Wr = final.detach().clone().requires_grad_(True), W.detach().clone().requires_grad_(True)
    br = b.detach().clone().requires_grad_(True) if bias else None
    lr = F.linear(fr[:, 0], Wr, br)
    lossr = F.cross_entropy(lr, labels, label_smoothing=0.05)
    (lossr * 1.7 + (lr * wl).sum()).backward()
    assert max_rel(logits, lr) < 1e-5
    assert float(loss) == pytest.approx(float(lossr), rel=1e-5)
    assert max_rel(final.grad, fr.grad) < 1e-5
    assert max_rel(W.grad, Wr.grad) < 1e-5
    if bias:
        assert max_rel(b.grad, br.grad) < 1e-5
    # no labels: logits only, cotangent on the logits only
    final.grad = None
    logits2, none = ops.head_ce(final, W, b, None, 0.05)
    assert none is None and torch.equal(logits2, logits)
    (logits2 * wl).sum().backward()
    fr.grad = None
    (F.linear(fr[:, 0], Wr, br) * wl).sum().backward()
    assert max_rel(final.grad, fr.grad) < 1e-5
