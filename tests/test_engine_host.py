"""Host-side pieces of the engine that need no GPU."""
import numpy as np
import torch

from surface_sampling_b200 import engine


def test_workspace_cache_reuses_and_grows_geometrically():
    """Regression: the cache once compared torch.device('cuda') with cuda:0, never matched, and re-allocated a
    GB-sized workspace on every relax call."""
    ws = engine._Workspace()
    a = ws.get(1000, torch.device("cpu"))
    assert a.numel() >= 1500                      # 1.5x head room
    assert ws.get(900, torch.device("cpu")) is a and ws.get(1400, "cpu") is a
    b = ws.get(5000, torch.device("cpu"))
    assert b is not a and b.numel() >= 7500 and ws.get(5000, "cpu") is b


def test_pack_painn_weights_tf32_split_is_exact_to_22_bits():
    """[exact | hi | lo] weight block: hi is TF32-representable and hi + lo reproduces the weight to 2^-21."""
    from oracle.painn import init_random_weights
    w = engine.pack_painn_weights(init_random_weights(0))
    n = w.size // 3
    exact, hi, lo = w[:n], w[n:2 * n], w[2 * n:]
    assert np.all((hi.view(np.uint32) & 0x1FFF) == 0) and np.all((lo.view(np.uint32) & 0x1FFF) == 0)
    err = np.abs((hi.astype(np.float64) + lo.astype(np.float64)) - exact.astype(np.float64))
    assert np.all(err <= np.abs(exact.astype(np.float64)) * 2.0 ** -21 + 1e-45)


def test_relax_outputs_are_caller_owned_or_fresh():
    """relax() results are never recycled behind the caller's back (round 1 used a 16-deep ring): fresh tensors per
    call, or the caller's own buffers after validation."""
    import pytest
    from surface_sampling_b200 import _lib
    eng = engine.PainnEngine.__new__(engine.PainnEngine)
    eng.device = torch.device("cpu")
    a, b = eng._result_buffers(4, 100, None), eng._result_buffers(4, 100, None)
    assert a[0].shape == (4, 8) and a[1].shape == (100, 3) and a[3].dtype == torch.int32
    assert a[0].data_ptr() != b[0].data_ptr() and a[1].data_ptr() != b[1].data_ptr()
    mine = {"out": torch.empty((4, 8), dtype=torch.float64), "forces": torch.empty((100, 3)),
            "forces_std": torch.empty((100, 3)), "status": torch.zeros(1, dtype=torch.int32)}
    got = eng._result_buffers(4, 100, mine)
    assert got[0] is mine["out"] and got[3] is mine["status"]
    mine["forces"] = torch.empty((99, 3))
    with pytest.raises(_lib.VssrError):
        eng._result_buffers(4, 100, mine)
