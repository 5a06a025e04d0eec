"""GPU (>= 2 devices, skipped on a 1-GPU box; run with `gpurun --gpus 2`): kernels that opt into large dynamic shared
memory must work on EVERY device of a process, not only the first one used (cudaFuncSetAttribute is per device)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_large_smem_kernels_on_second_device():
    from hyptokenizer_b200.embedding import lorentz_model as LM
    from hyptokenizer_b200.knn import lorentz_topk
    from hyptokenizer_b200.pair_count import count_pairs
    from hyptokenizer_b200.synth import synthetic_embeddings
    from oracle import merge as OM
    text = ("ab cd\n  hello world \n\nxyz\n" * 400).encode()
    want = OM.count_pairs_py(text.decode().splitlines(True))
    res = {}
    for dev in (0, 1):
        with torch.cuda.device(dev):
            assert count_pairs(text, device=torch.device("cuda", dev)) == want           # ~216 KB of dynamic smem
            E = synthetic_embeddings(600, 300, scale=0.1, seed=2, device=f"cuda:{dev}")   # D = 301: midpoint needs > 48 KB
            i = torch.arange(0, 64, dtype=torch.int32, device=E.device)
            j = i + 100
            ln = torch.ones(64, dtype=torch.int32, device=E.device)
            from hyptokenizer_b200._lib import check, lib, ptr, stream_ptr
            out = torch.empty((64, 301), device=E.device)
            check(lib().hyp_midpoint(ptr(E), 301, ptr(i), ptr(j), ptr(ln), ptr(ln), ptr(out), 301, 64, 301, 1.0, 1, 1,
                                     stream_ptr()))
            gi, gd = lorentz_topk(E, 8, 1.0, "lorentz")
            res[dev] = (out.cpu(), gi.cpu(), gd.cpu(), LM.batch_distance(E[:5], E[:7], 1.0, semantics="lorentz").cpu())
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a, b)
