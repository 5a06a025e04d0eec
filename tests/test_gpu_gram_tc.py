"""GPU parity: the tcgen05 Gram path (K2, tensor cores) must return exactly what the exact CUDA-core
kernel returns -- it is a certified filter followed by fp32 re-scoring, never an approximation."""
import pytest
import torch

from helpers import same_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["ss", "ts"])
def tc_engine(request, monkeypatch):
    """Every test of this file runs on both tensor-core engines: "ss" (both MMA operands in shared memory, 128-column
    tiles) and "ts" (A operands in tensor memory, 64-column tiles); HYP_TC_ENGINE is read at every call."""
    monkeypatch.setenv("HYP_TC_ENGINE", request.param)
    return request.param


@pytest.mark.parametrize("n,d,k,scale,nrows,row0", [(1024, 100, 32, 0.05, None, 0), (4500, 100, 32, 0.01, None, 0),
                                                    (4500, 50, 16, 0.3, 1000, 777), (9000, 100, 32, 0.01, 3000, 6000),
                                                    (2100, 124, 8, 0.1, None, 0), (1500, 7, 4, 0.2, None, 0),
                                                    (3000, 48, 8, 0.1, None, 0), (2000, 16, 4, 0.2, 700, 1300),
                                                    (2500, 90, 8, 0.1, None, 0), (2500, 64, 8, 0.1, None, 0)])
@pytest.mark.parametrize("pair,seg", [("1", None), ("2", None), ("2", "1"), ("2", "5"), ("1", "16")])
def test_tc_equals_exact(monkeypatch, pair, seg, n, d, k, scale, nrows, row0):
    from hyptokenizer_b200 import knn
    from hyptokenizer_b200.knn import lorentz_topk
    from hyptokenizer_b200.synth import synthetic_embeddings
    monkeypatch.setenv("HYP_TC_PAIR", pair)      # one or two row blocks per work item
    if seg is not None:
        monkeypatch.setenv("HYP_TC_SEG", seg)    # column segments per row pair (default: the library's cost model)
    E = synthetic_embeddings(n, d, scale=scale, seed=n + d, device="cuda")
    nrows = n - row0 if nrows is None else nrows
    ei, ed = lorentz_topk(E, k, 1.0, "lorentz", n, row0, nrows, engine="exact")
    ti, td = lorentz_topk(E, k, 1.0, "lorentz", n, row0, nrows, engine="tc")
    torch.cuda.synchronize()
    assert torch.equal(ti, ei) and same_bits(td, ed)
    # with at least ~1.5 k column tiles the filter, not the fallback, does the work on non-degenerate data
    # (fewer than k tiles cannot bound the k-th best: every row is flagged and redone exactly)
    if n >= 128 * k * 3 // 2:
        assert knn.last_flagged() <= max(2, nrows // 100)
    print("flagged", knn.last_flagged(), "of", nrows)


def test_tc_degenerate_inputs_fall_back():
    """Shipped semantics (every distance 0.0), duplicated rows and NaN rows: the certificate fails or the
    buffers overflow, rows are flagged and recomputed exactly -- same output as the exact kernel."""
    from hyptokenizer_b200.knn import lorentz_topk
    from hyptokenizer_b200.synth import synthetic_embeddings
    n, k = 3000, 32
    E = synthetic_embeddings(n, 100, scale=0.05, seed=1, device="cuda")
    ei, ed = lorentz_topk(E, k, 1.0, "reference", engine="exact")
    ti, td = lorentz_topk(E, k, 1.0, "reference", engine="tc")
    assert torch.equal(ti, ei) and same_bits(td, ed)
    E[100:200] = E[5]                    # 100 duplicates: a 100-way tie at distance 0 for those rows
    E[777] = float("nan")
    ei, ed = lorentz_topk(E, k, 1.0, "lorentz", engine="exact")
    ti, td = lorentz_topk(E, k, 1.0, "lorentz", engine="tc")
    assert torch.equal(ti, ei) and same_bits(td, ed)


def test_tc_full_size_shard_equals_exact():
    """BASELINE configs[2] size: a 2048-row shard of V=100k, d=100, k=32 through the tensor-core path is bit-identical to
    the exact CUDA-core kernel, with no row needing the fallback."""
    from hyptokenizer_b200.knn import lorentz_topk
    from hyptokenizer_b200.synth import synthetic_embeddings
    n, k, row0, nrows = 100000, 32, 77000, 2048
    E = synthetic_embeddings(n, 100, scale=0.01, seed=42, device="cuda")
    ei, ed = lorentz_topk(E, k, 1.0, "lorentz", n, row0, nrows, engine="exact")
    ti, td = lorentz_topk(E, k, 1.0, "lorentz", n, row0, nrows, engine="tc")
    assert torch.equal(ti, ei) and same_bits(td, ed)
    from hyptokenizer_b200 import knn
    assert knn.last_flagged() == 0


def test_tc_dimension_limit_and_large_scale():
    """d + 4 operand columns (the time-like term rides inside the MMA as hi/lo TF32 parts) must fit 128: d = 125 is
    refused by the tensor-core engine and `auto` takes the exact kernel.  Points far from the origin (x0 up to ~50,
    where one TF32 slot for x0 would be off by 1) still give the exact lists."""
    from hyptokenizer_b200.knn import lorentz_topk
    from hyptokenizer_b200.synth import synthetic_embeddings
    E = synthetic_embeddings(9000, 125, scale=0.1, seed=3, device="cuda")
    with pytest.raises(ValueError):
        lorentz_topk(E, 8, 1.0, "lorentz", engine="tc")
    ai, ad = lorentz_topk(E, 8, 1.0, "lorentz", engine="auto")
    ei, ed = lorentz_topk(E, 8, 1.0, "lorentz", engine="exact")
    assert torch.equal(ai, ei) and same_bits(ad, ed)
    for scale in (0.5, 1.5):
        F = synthetic_embeddings(6000, 40, scale=scale, seed=17, device="cuda")
        assert float(F[:, 0].max()) > (3.0 if scale == 0.5 else 30.0)
        ti, td = lorentz_topk(F, 16, 1.0, "lorentz", engine="tc")
        xi, xd = lorentz_topk(F, 16, 1.0, "lorentz", engine="exact")
        assert torch.equal(ti, xi) and same_bits(td, xd)
