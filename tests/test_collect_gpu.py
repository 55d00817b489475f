"""Receiver collection kernels against a numpy restatement of `_collect_receiver_energy`
(reference RadiosityFast.py:1148-1185: out[r, b] += roll(E[k, rdir, b] * scale, shift)) and
against each other: `k_collect_staged` (rows staged in shared memory for a group of receivers,
the default for diffuse scenes with many receivers) vs `k_collect_partial` (one row read per
receiver).  Operator level, through the C ABI; the class-level parity with the oracle is in
test_class_gpu.py (ground plane, config-3 analogue)."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float64: 1e-12, torch.float32: 2e-5}


def make_case(n_rcv, n_patches, n_bands, n_samples, dtype, seed=0, n_alloc=None, invisible=0.2):
    from sparrowpy_b200 import exchange
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(seed)
    n_alloc = n_alloc or n_patches
    pad = 32
    ld = pad + -(-n_samples // 256) * 256
    data = torch.zeros((n_bands * n_alloc, ld), dtype=dtype)
    data[:, pad:pad + n_samples] = torch.rand((n_bands * n_alloc, n_samples), generator=g,
                                              dtype=torch.float64).to(dtype)
    # what lies beyond the T bins of a row must never reach a result
    data[:, pad + n_samples:] = 1e30
    data[:, :pad] = -1e30
    shift = torch.randint(0, n_samples, (n_rcv, n_patches), generator=g, dtype=torch.int32)
    shift[:, 0] = 0
    shift[:, -1] = n_samples - 1
    scale = torch.rand((n_rcv, n_patches, n_bands), generator=g, dtype=torch.float64)
    scale[torch.rand((n_rcv, n_patches), generator=g) < invisible] = 0.0
    rdir = torch.zeros((n_rcv, n_patches), dtype=torch.int32)
    hist = exchange.EnergyHistogram(data.to(dev), n_patches, 1, n_bands, n_samples, pad,
                                    n_alloc=n_alloc)
    e = data.view(n_bands, n_alloc, ld)[:, :n_patches, pad:pad + n_samples].double().numpy()
    ref = np.zeros((n_rcv, n_bands, n_samples))
    sh, sc = shift.numpy(), scale.to(dtype).double().numpy()
    for r in range(n_rcv):
        for k in range(n_patches):
            for b in range(n_bands):
                if sc[r, k, b] != 0.0:
                    ref[r, b] += np.roll(e[b, k] * sc[r, k, b], sh[r, k])
    return hist, rdir.to(dev), shift.to(dev), scale.to(dev), ref


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("kind", ["direct", "staged", "staged:2", "staged:3", "staged:4"])
@pytest.mark.parametrize("n_rcv,n_patches,n_bands,n_samples", [
    (13, 70, 2, 1000),      # receivers not a multiple of the group, T a multiple of 4
    (17, 33, 3, 333),       # odd T: copy A of the doubled row goes element by element
    (1, 40, 1, 2050),       # three time chunks
    (40, 300, 1, 64),       # T shorter than one chunk; several patch splits
])
def test_collect_kernels_match_numpy(kind, dtype, n_rcv, n_patches, n_bands, n_samples,
                                     monkeypatch):
    from sparrowpy_b200 import exchange
    monkeypatch.setenv("SPB_COLLECT", kind)
    hist, rdir, shift, scale, ref = make_case(n_rcv, n_patches, n_bands, n_samples, dtype,
                                              n_alloc=n_patches + 3)
    for n_split in (None, 1, 3):
        mono = exchange.collect_mono(hist, rdir, shift, scale, n_split=n_split)
        assert mono.shape == (n_rcv, n_bands, n_samples)
        assert torch.isfinite(mono).all()
        assert rel_err(mono.double().cpu().numpy(), ref) < TOL[dtype]


def test_default_kernel_choice(monkeypatch):
    """Staged for diffuse scenes with >= 4 receivers whose rows fit shared memory, direct
    otherwise (directional histograms read a receiver-dependent row)."""
    from sparrowpy_b200 import exchange
    monkeypatch.delenv("SPB_COLLECT", raising=False)
    hist, *_ = make_case(1, 4, 1, 1000, torch.float64)
    assert exchange.collect_kind(hist, 64) == ("staged", 0)
    assert exchange.collect_kind(hist, 4) == ("staged", 0)
    assert exchange.collect_kind(hist, 3) == ("direct",)
    hist.n_samples = 2000
    assert exchange.collect_kind(hist, 64) == ("staged", 0)
    hist.n_samples = 40000
    assert exchange.collect_kind(hist, 64) == ("direct",)
    hist.n_samples, hist.n_dirs = 1000, 4
    assert exchange.collect_kind(hist, 64) == ("direct",)
    monkeypatch.setenv("SPB_COLLECT", "staged")
    with pytest.raises(Exception):
        exchange.collect_kind(hist, 64)


def test_staged_equals_direct_for_a_source_batch(monkeypatch):
    """Sources ride as extra bands (hist.n_sources): both kernels give the same (S, R, B, T)."""
    from sparrowpy_b200 import exchange
    hist, rdir, shift, scale, _ = make_case(9, 120, 6, 500, torch.float64)
    hist.n_sources = 3
    scale = scale[:, :, :2].contiguous()
    out = {}
    for kind in ("direct", "staged"):
        monkeypatch.setenv("SPB_COLLECT", kind)
        out[kind] = exchange.collect_mono(hist, rdir, shift, scale).cpu().numpy()
    assert out["staged"].shape == (3, 9, 2, 500)
    assert rel_err(out["staged"], out["direct"]) < 1e-13
