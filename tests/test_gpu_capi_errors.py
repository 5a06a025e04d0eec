"""Error behaviour of the C ABI: negative status + hyp_last_error(), mapped by the Python layer to
ValueError (bad arguments / table full) or RuntimeError; no call may touch memory when its arguments are refused."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_status_codes_and_messages():
    from hyptokenizer_b200 import _lib
    from hyptokenizer_b200._lib import ptr, stream_ptr
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    E = torch.zeros((64, 9), dtype=torch.float32, device=dev)
    E[:, 0] = 1.0
    out = torch.empty(64, device=dev)
    # D < 2 and a non-positive curvature are refused before any launch
    assert L.hyp_distance(ptr(E), 9, ptr(E), 9, ptr(out), 64, 1, 1.0, 0, stream_ptr()) == _lib.HYP_ERR_ARG
    assert L.hyp_merge_steps(ptr(E), 9, None, 9, 1.0, 0, None, None, 1, 0, 0, 1.0, 64, None, 0, stream_ptr()) == _lib.HYP_ERR_ARG
    assert b"hyp_merge_steps" in L.hyp_last_error()
    # a workspace that is too small names the size it needs
    ws = torch.empty(64, dtype=torch.uint8, device=dev)
    best = torch.empty(32, dtype=torch.uint8, device=dev)
    rc = L.hyp_allpairs_min(ptr(E), 9, 64, 9, 1.0, 1, 1.0, ptr(best), ptr(ws), 8, stream_ptr())
    assert rc == _lib.HYP_ERR_WORKSPACE and b"workspace" in L.hyp_last_error()
    with pytest.raises(RuntimeError, match="workspace"):
        _lib.check(rc)
    # the tensor-core top-k: k > 32, an unaligned workspace, and d + 4 > 128 columns
    idx = torch.empty((64, 40), dtype=torch.int32, device=dev)
    dd = torch.empty((64, 40), dtype=torch.float32, device=dev)
    fl = torch.empty(64, dtype=torch.int32, device=dev)
    big = torch.empty(L.hyp_gram_topk_workspace_bytes(64, 64, 9) + 512, dtype=torch.uint8, device=dev)
    base = big.data_ptr() + ((-big.data_ptr()) % 256)
    assert L.hyp_gram_topk(ptr(E), 9, 64, 0, 64, 9, 1.0, 1, 33, ptr(idx), ptr(dd), ptr(fl), base, big.numel() - 512,
                           stream_ptr()) == _lib.HYP_ERR_ARG
    assert L.hyp_gram_topk(ptr(E), 9, 64, 0, 64, 9, 1.0, 1, 8, ptr(idx), ptr(dd), ptr(fl), base + 4, big.numel() - 512,
                           stream_ptr()) == _lib.HYP_ERR_ARG
    assert L.hyp_gram_topk_workspace_bytes(64, 64, 126) == -1
    # pair counting: the table capacity must be a power of two, the text 16-byte aligned
    text = torch.zeros(64, dtype=torch.uint8, device=dev)
    asc = torch.empty(128 * 128, dtype=torch.int64, device=dev)
    keys = torch.empty(24, dtype=torch.int64, device=dev)
    vals = torch.empty(24, dtype=torch.int64, device=dev)
    ovf = torch.empty(1, dtype=torch.int32, device=dev)
    assert L.hyp_pair_count(ptr(text), 64, ptr(asc), ptr(keys), ptr(vals), 24, ptr(ovf), stream_ptr()) == _lib.HYP_ERR_ARG
    keys16 = torch.empty(16, dtype=torch.int64, device=dev)
    assert L.hyp_pair_count(text.data_ptr() + 1, 32, ptr(asc), ptr(keys16), ptr(keys16), 16, ptr(ovf),
                            stream_ptr()) == _lib.HYP_ERR_ARG
    torch.cuda.synchronize()                      # nothing above launched anything that could fault
    assert _lib.HYP_OK == 0 and L.hyp_abi_version() == 1


def test_python_layer_maps_errors():
    from hyptokenizer_b200.embedding import lorentz_model as LM
    from hyptokenizer_b200.knn import lorentz_topk
    x = torch.ones((4, 6), device="cuda")
    with pytest.raises(ValueError):
        LM.distance(x, torch.ones((4, 5), device="cuda"))          # last dimensions differ
    with pytest.raises(ValueError):
        LM.batch_distance(x, torch.ones((6,), device="cuda"))      # not 2-D
    with pytest.raises(ValueError):
        lorentz_topk(x, 2, engine="nope")
    with pytest.raises(RuntimeError):
        lorentz_topk(x.cpu(), 2)
