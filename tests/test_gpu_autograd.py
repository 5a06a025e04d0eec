"""the Lorentz distance under autograd (multimodal/contrastive_loss.py) and the batched distance
of eval_hierarchy.compute_distortion, against the oracle's torch-CPU restatement run under torch autograd.
Tolerance: 1e-5 relative (north_star), scaled by the largest gradient entry of the row."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL = 1e-5


def _points(n, d, scale, seed):
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(n, d, generator=g) * scale
    x0 = torch.sqrt(1.0 + (v * v).sum(-1, keepdim=True))
    return torch.cat([x0, v], dim=-1)


def _close(got, want, rel=REL):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    scale = want.abs().amax(dim=-1, keepdim=True).clamp_min(1e-30) if want.dim() else want.abs().clamp_min(1e-30)
    return bool(((got - want).abs() <= rel * scale + 1e-12).all())


@pytest.mark.parametrize("sem", ["lorentz", "reference"])
@pytest.mark.parametrize("n,d,scale", [(64, 100, 0.3), (33, 7, 1.0), (5, 3, 0.05)])
def test_distance_gradients_match_autograd(sem, n, d, scale):
    from hyptokenizer_b200.embedding import lorentz_model as LM
    from oracle import lorentz as OL
    x, y = _points(n, d, scale, 1), _points(n, d, scale, 2)
    g = torch.randn(n, generator=torch.Generator().manual_seed(3))
    xo, yo = x.clone().requires_grad_(), y.clone().requires_grad_()
    OL.distance(xo, yo, 1.3, sem).backward(g)
    xg, yg = x.cuda().requires_grad_(), y.cuda().requires_grad_()
    out = LM.distance(xg, yg, 1.3, semantics=sem)
    out.backward(g.cuda())
    assert _close(xg.grad, xo.grad) and _close(yg.grad, yo.grad)
    if sem == "reference":                      # as shipped: every distance is clamped, the gradient is exactly 0
        assert torch.count_nonzero(xg.grad).item() == 0


def test_distance_gradient_broadcast_and_partial():
    """One row against many (the reference's `expand`ed row, contrastive_loss.py:41-45): the broadcast operand
    receives the sum; an operand that does not require grad receives none."""
    from hyptokenizer_b200.embedding import lorentz_model as LM
    from oracle import lorentz as OL
    x, y = _points(1, 20, 0.4, 5), _points(50, 20, 0.4, 6)
    xo = x.clone().requires_grad_()
    OL.distance(xo.expand(50, -1), y, 1.0, "lorentz").sum().backward()
    xg = x.cuda().requires_grad_()
    yg = y.cuda()
    LM.distance(xg.expand(50, -1), yg, 1.0, semantics="lorentz").sum().backward()
    assert _close(xg.grad, xo.grad) and yg.grad is None
    xg2 = x.cuda().requires_grad_()
    LM.distance(xg2, yg, 1.0, semantics="lorentz").sum().backward()      # implicit broadcasting
    assert _close(xg2.grad, xo.grad)


@pytest.mark.parametrize("sem", ["lorentz", "reference"])
@pytest.mark.parametrize("n1,n2,d", [(70, 70, 100), (129, 65, 50), (3, 200, 7)])
def test_batch_distance_gradients_match_autograd(sem, n1, n2, d):
    from hyptokenizer_b200.embedding import lorentz_model as LM
    from oracle import lorentz as OL
    x, y = _points(n1, d, 0.3, 11), _points(n2, d, 0.3, 12)
    g = torch.randn(n1, n2, generator=torch.Generator().manual_seed(13))
    xo, yo = x.clone().requires_grad_(), y.clone().requires_grad_()
    OL.batch_distance(xo, yo, 0.7, sem).backward(g)
    xg, yg = x.cuda().requires_grad_(), y.cuda().requires_grad_()
    LM.batch_distance(xg, yg, 0.7, semantics=sem).backward(g.cuda())
    assert _close(xg.grad, xo.grad, 5e-5) and _close(yg.grad, yo.grad, 5e-5)   # sums of n2 (n1) terms, fp32 GEMM order


def _oracle_contrastive(z_text, z_img, temp, reduction, sem):
    """multimodal/contrastive_loss.py:17-60, row by row as shipped, on the oracle's distance."""
    import torch.nn.functional as F
    from oracle import lorentz as OL
    B = z_text.size(0)
    rows = [OL.distance(z_text[i].unsqueeze(0).expand(B, -1), z_img, 1.0, sem) for i in range(B)]
    sim = -torch.stack(rows) / temp
    labels = torch.arange(B)
    return (F.cross_entropy(sim, labels, reduction=reduction) + F.cross_entropy(sim.t(), labels, reduction=reduction)) / 2.0


@pytest.mark.parametrize("sem", ["lorentz", "reference"])
@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
def test_contrastive_loss_and_gradients(sem, reduction):
    from hyptokenizer_b200.multimodal import HyperbolicInfoNCE, hyperbolic_contrastive_loss
    B, d = 48, 32
    zt, zi = _points(B, d, 0.5, 21), _points(B, d, 0.5, 22)
    to, io = zt.clone().requires_grad_(), zi.clone().requires_grad_()
    want = _oracle_contrastive(to, io, 0.07, reduction, sem)
    want.sum().backward()
    tg, ig = zt.cuda().requires_grad_(), zi.cuda().requires_grad_()
    got = hyperbolic_contrastive_loss(tg, ig, temp=0.07, reduction=reduction, semantics=sem)
    got.sum().backward()
    assert got.shape == want.shape
    assert torch.allclose(got.detach().cpu(), want.detach(), rtol=1e-5, atol=1e-6)
    assert _close(tg.grad, to.grad, 1e-4) and _close(ig.grad, io.grad, 1e-4)   # softmax at temp 0.07 amplifies 1e-5
    if reduction == "mean":
        mod = HyperbolicInfoNCE(0.07, semantics=sem)
        assert torch.equal(mod(zt.cuda(), zi.cuda()), got.detach())
        if sem == "reference":                  # all distances 0: the loss is log(B), whatever the inputs
            assert abs(float(got.detach()) - float(np.log(B))) < 1e-6


def test_triplet_loss_and_gradients():
    import torch.nn.functional as F
    from hyptokenizer_b200.multimodal import hyperbolic_triplet_loss
    from oracle import lorentz as OL
    a, p_, n_ = _points(40, 16, 0.6, 31), _points(40, 16, 0.6, 32), _points(40, 16, 0.6, 33)
    ao = a.clone().requires_grad_()
    want = F.relu(OL.distance(ao, p_, 1.0, "lorentz") - OL.distance(ao, n_, 1.0, "lorentz") + 1.0).mean()
    want.backward()
    ag = a.cuda().requires_grad_()
    got = hyperbolic_triplet_loss(ag, p_.cuda(), n_.cuda(), margin=1.0, semantics="lorentz")
    got.backward()
    assert abs(float(got.detach()) - float(want.detach())) <= 1e-5 * abs(float(want.detach())) + 1e-7
    assert _close(ag.grad, ao.grad)
    none = hyperbolic_triplet_loss(a.cuda(), p_.cuda(), n_.cuda(), reduction="none", semantics="lorentz")
    assert none.shape == (40,)


def test_distortion_ratios_match_per_pair_loop():
    """eval_hierarchy.py:139-170: per-pair `distance(...).item() / graph_dist`, then numpy statistics."""
    from hyptokenizer_b200.eval import distortion_ratios
    from oracle import lorentz as OL
    E = _points(500, 50, 0.4, 41)
    rng = np.random.default_rng(0)
    pairs = [(int(a), int(b), float(g)) for a, b, g in zip(rng.integers(0, 500, 300), rng.integers(0, 500, 300),
                                                          rng.integers(1, 12, 300))]
    want = np.array([OL.distance(E[i].unsqueeze(0), E[j].unsqueeze(0), 1.0, "lorentz").item() / g for i, j, g in pairs])
    ratios, stats = distortion_ratios(E.cuda(), pairs, 1.0, semantics="lorentz")
    assert np.all(np.abs(ratios - want) <= REL * np.abs(want) + 1e-9)
    assert stats["num_pairs"] == 300 and abs(stats["mean"] - float(np.mean(want))) <= 1e-5 * abs(float(np.mean(want)))
    with pytest.raises(IndexError):
        distortion_ratios(E.cuda(), [(0, 500, 1.0)])


def test_non_differentiable_ops_fail_loudly():
    """The reference computes every Lorentz function with torch ops; here only distance / batch_distance have a
    backward kernel.  The others must refuse inputs that require grad instead of silently returning detached results."""
    from hyptokenizer_b200.embedding import lorentz_model as LM
    from hyptokenizer_b200.synth import synthetic_embeddings
    x = synthetic_embeddings(4, 6, scale=0.3, seed=1, device="cuda").requires_grad_(True)
    y = synthetic_embeddings(4, 6, scale=0.3, seed=2, device="cuda")
    for fn, args in ((LM.minkowski_dot, (x, y)), (LM.exp_map, (y, x)), (LM.log_map, (x, y)), (LM.project_to_hyperboloid, (x,))):
        with pytest.raises(RuntimeError, match="no backward pass"):
            fn(*args)
        with torch.no_grad():
            fn(*args)                                   # fine when no graph is being recorded
        fn(*[a.detach() for a in args])
    LM.distance(x, y, semantics="lorentz").sum().backward()    # the differentiable ones still are
    assert x.grad is not None
