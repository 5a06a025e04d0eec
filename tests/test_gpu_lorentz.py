"""GPU parity: K1/K2/K3 kernels (through the Python mirror of embedding/lorentz_model.py, i.e.
through the C ABI) against the golden vectors of the unmodified reference and against the oracle.

Bars: the Minkowski product, norm/projection and everything without a transcendental are BIT-EXACT;
acosh/cosh/sinh-dependent outputs are within 1e-5 relative (north_star), in practice a few ulp.
"""
import numpy as np
import pytest
import torch

from helpers import from_bits, same_bits, to_bits, ulp_diff

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5      # north_star: distances within 1e-5 relative in fp32
ULP_TOL = 8         # what the log1p-form acosh / CUDA cosh/sinh actually achieve vs torch CPU


def g(b):
    return np.asarray(b, dtype=np.uint32).view(np.float32)


def rows_close(got, want, D, rel=REL_TOL):
    """Vector outputs (tangent vectors, points): error measured against the largest component of the
    row -- component-wise relative error is meaningless where y - <x,y>x cancels to ~0."""
    got = got.detach().cpu().numpy().reshape(-1, D)
    want = np.asarray(want, dtype=np.float32).reshape(-1, D)
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern differs"
    inf = np.isinf(want)
    assert np.array_equal(got[inf], want[inf]), "infinities differ"
    fin = np.isfinite(want)
    scale = np.where(fin, np.abs(want), 0).max(axis=1, keepdims=True) * np.ones_like(want)
    err = np.abs(np.where(fin, got, 0).astype(np.float64) - np.where(fin, want, 0))
    assert np.all(err <= rel * scale + 1e-30)


def close(got, want, rel=REL_TOL, ulps=None):
    got = got.detach().cpu().numpy().reshape(-1) if isinstance(got, torch.Tensor) else np.asarray(got).reshape(-1)
    want = np.asarray(want, dtype=np.float32).reshape(-1)
    assert got.shape == want.shape
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), "NaN pattern differs"
    ok = ~nan_w
    err = np.abs(got[ok].astype(np.float64) - want[ok].astype(np.float64))
    bound = rel * np.abs(want[ok].astype(np.float64)) + 1e-12
    assert np.all(err <= bound), f"max rel err {np.max(err / (np.abs(want[ok]) + 1e-30))}"
    if ulps is not None:
        assert ulp_diff(got[ok], want[ok]).max(initial=0) <= ulps


@pytest.fixture(scope="module")
def LM():
    from hyptokenizer_b200.embedding import lorentz_model
    return lorentz_model


def test_golden_lorentz_ops(golden, LM):
    dev = torch.device("cuda:0")
    for case in golden("lorentz_ops.json")["cases"]:
        d, n, c = case["d"], case["n"], case["c"]
        X = from_bits(case["X"], n, d + 1).to(dev)
        P = from_bits(case["P"], n, d + 1).to(dev)
        ia, ib = torch.tensor(case["ia"], device=dev), torch.tensor(case["ib"], device=dev)
        # bit-exact family
        assert same_bits(LM.minkowski_dot(X.unsqueeze(1), X.unsqueeze(0)), g(case["mdot"]))
        # torch's CPU sqrt is a Sleef routine that is NOT correctly rounded (0.7 % of inputs are 1 ulp off
        # IEEE sqrt, probed); the device uses IEEE sqrt, so sqrt outputs are pinned to 1 ulp, the rest exactly
        assert ulp_diff(LM.minkowski_norm(X), g(case["mnorm"])).max() <= 1
        pr = LM.project_to_hyperboloid(P, c)
        want = torch.from_numpy(g(case["project"]).reshape(n, d + 1).copy())
        assert same_bits(pr[:, 1:], want[:, 1:]) and ulp_diff(pr[:, 0], want[:, 0]).max() <= 1
        pr = LM.project_to_hyperboloid(P[2], c)
        want = torch.from_numpy(g(case["project_row"]).copy())
        assert same_bits(pr[1:], want[1:]) and ulp_diff(pr[:1], want[:1]).max() <= 1
        # transcendental family
        V = from_bits(case["V"], len(ia), d + 1).to(dev)
        rows_close(LM.exp_map(X[ia], V, c), g(case["exp_map"]), d + 1)
        for sem in ("reference", "lorentz"):
            r = case[sem]
            got_d = LM.distance(X[ia], X[ib], c, semantics=sem)
            got_b = LM.batch_distance(X, X, c, semantics=sem)
            got_l = LM.log_map(X[ia], X[ib], c, semantics=sem)
            if sem == "reference":
                # shipped arithmetic: every distance is exactly 0.0 and every log map NaN
                assert same_bits(got_d, g(r["distance"]))
                assert same_bits(got_b, g(r["batch_distance"]))
                assert same_bits(got_l, g(r["log_map"]))
            else:
                close(got_d, g(r["distance"]), ulps=ULP_TOL)
                close(got_b, g(r["batch_distance"]), ulps=ULP_TOL)
                rows_close(got_l, g(r["log_map"]), d + 1)


def test_golden_midpoint(golden):
    from hyptokenizer_b200 import _lib
    from hyptokenizer_b200._lib import SEM, check, ptr, stream_ptr
    dev = torch.device("cuda:0")
    for case in golden("lorentz_ops.json")["cases"]:
        d, n, c = case["d"], case["n"], case["c"]
        X = from_bits(case["X"], n, d + 1).to(dev)
        m = len(case["ia"])
        ia = torch.tensor(case["ia"], dtype=torch.int32, device=dev)
        ib = torch.tensor(case["ib"], dtype=torch.int32, device=dev)
        li = torch.full((m,), 1, dtype=torch.int32, device=dev)
        lj = torch.full((m,), 3, dtype=torch.int32, device=dev)
        for sem in ("reference", "lorentz"):
            out = torch.empty((m, d + 1), device=dev)
            check(_lib.lib().hyp_midpoint(ptr(X), d + 1, ptr(ia), ptr(ib), ptr(li), ptr(lj), ptr(out), d + 1, m, d + 1,
                                          c, SEM[sem], 1, stream_ptr()))
            rows_close(out, g(case[sem]["midpoint_1_3"]), d + 1)


@pytest.mark.parametrize("d,scale", [(50, 0.01), (100, 0.01), (100, 0.3), (37, 0.1), (8, 0.2), (5, 0.5), (130, 0.05)])
def test_u_bitexact_vs_oracle_numpy(LM, d, scale):
    """Pre-clamp Minkowski product against the plain-numpy statement of ATen's order (host independent)."""
    from oracle import sumorder as S
    g_ = torch.Generator().manual_seed(d)
    n = 40
    X = torch.randn(n, d + 1, generator=g_) * scale
    X[:, 0] = torch.sqrt(1 + (X[:, 1:] ** 2).sum(-1))
    got = LM.minkowski_dot(X.cuda().unsqueeze(1), X.cuda().unsqueeze(0)).cpu().numpy()
    Xn = X.numpy()
    for i in range(0, n, 3):
        for j in range(n):
            assert to_bits(got[i, j])[0] == to_bits(S.mdot_fp32(Xn[i], Xn[j]))[0], (i, j)


@pytest.mark.parametrize("n1,n2,d", [(1, 1, 100), (65, 130, 100), (257, 63, 50), (300, 300, 7), (129, 64, 3)])
def test_batch_distance_vs_oracle(LM, n1, n2, d):
    from oracle import lorentz as OL
    from oracle import sumorder as S
    torch.manual_seed(n1 * 1000 + n2)
    X = OL.initialize_embeddings(n1, d, scale=0.2)
    Y = OL.initialize_embeddings(n2, d, scale=0.2)
    for sem in ("reference", "lorentz"):
        want = OL.batch_distance(X, Y, 1.3, sem)
        got = LM.batch_distance(X.cuda(), Y.cuda(), 1.3, semantics=sem)
        close(got, want.numpy(), ulps=ULP_TOL)
    if S.host_matches_torch():
        # pairwise kernel and tile kernel agree bit for bit with each other and with torch's product
        u_pair = LM.minkowski_dot(X.cuda().unsqueeze(1), Y.cuda().unsqueeze(0))
        assert same_bits(u_pair, OL.gram_u(X, Y))


def test_rescore_pairs_matches_distance(LM):
    from hyptokenizer_b200 import _lib
    from hyptokenizer_b200._lib import check, ptr, stream_ptr
    from oracle import lorentz as OL
    torch.manual_seed(5)
    E = OL.initialize_embeddings(500, 100, scale=0.1).cuda()
    ii = torch.randint(0, 500, (4096,), dtype=torch.int32, device="cuda")
    jj = torch.randint(0, 500, (4096,), dtype=torch.int32, device="cuda")
    dd = torch.empty(4096, device="cuda")
    uu = torch.empty(4096, device="cuda")
    check(_lib.lib().hyp_rescore_pairs(ptr(E), 101, ptr(ii), ptr(jj), ptr(dd), ptr(uu), 4096, 101, 1.0, 1, stream_ptr()))
    assert same_bits(dd, LM.distance(E[ii.long()], E[jj.long()], 1.0, semantics="lorentz"))
    assert same_bits(uu, LM.minkowski_dot(E[ii.long()], E[jj.long()]))
    # and against the dense tile kernel (different code path, same summation order)
    full = LM.batch_distance(E, E, 1.0, semantics="lorentz")
    assert same_bits(dd, full[ii.long(), jj.long()])


def test_edge_cases(LM):
    dev = "cuda"
    # empty batch
    e = torch.empty(0, 11, device=dev)
    assert LM.distance(e, e).shape == (0,)
    assert LM.batch_distance(e, torch.randn(4, 11, device=dev)).shape == (0, 4)
    # NaN rows propagate, never turn into zeros
    x = torch.randn(3, 11, device=dev)
    x[1] = float("nan")
    d = LM.distance(x, x.flip(0), semantics="lorentz")
    assert torch.isnan(d[1]) and torch.isnan(LM.batch_distance(x, x, semantics="lorentz")[1]).all()
    # 1-D operands (scripts/train_hyperbolic_tokenizer.py:104 calls exp_map on vectors)
    o = torch.zeros(11, device=dev)
    o[0] = 1
    t = torch.zeros(11, device=dev)
    assert torch.allclose(LM.exp_map(o, t), o, atol=1e-5)      # reference test_exp_map
    # reference test_project_to_hyperboloid / test_minkowski_norm known answers
    p = LM.project_to_hyperboloid(torch.randn(10, 4, device=dev))
    assert torch.allclose(LM.minkowski_dot(p, p), torch.ones(10, device=dev), atol=1e-5)
    assert (p[:, 0] > 0).all()
    assert torch.allclose(LM.minkowski_norm(p), torch.ones(10, device=dev), atol=1e-5)
    # wrong device / dtype fail loudly
    with pytest.raises(RuntimeError):
        LM.distance(torch.randn(2, 5), torch.randn(2, 5))
    with pytest.raises(TypeError):
        LM.distance(torch.randn(2, 5, device=dev).double(), torch.randn(2, 5, device=dev).double())
