"""CPU suite, part 2: the C-ABI library loads and exports every symbol include/b2c.h declares,
the host-side mirror keeps the reference's interface, and nothing computes without a GPU."""
import os
import re

import pytest
import torch

import multimodal_vqvae_compression_audio_tactile_b200 as pkg
from multimodal_vqvae_compression_audio_tactile_b200 import _lib, engine, modules
from oracle import cases, proposed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(pkg.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return pkg.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "b2c.h")).read()
    declared = set(re.findall(r"\b(b2c_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b2c_abi_version() == _lib.ABI_VERSION == int(re.search(r"#define B2C_ABI_VERSION (\d+)", hdr).group(1))


def test_no_cpu_fallback(lib):
    net = pkg.build_proposed(1, 128)
    x = torch.zeros(1, 1, 4800)
    for call in (lambda: net.forward_eval(x, x), lambda: net.A_ENC(x), lambda: net.T_DEC(torch.zeros(1, 1024, 4)),
                 lambda: pkg.nearest_code(torch.zeros(4, 8), torch.zeros(16, 8)),
                 lambda: net.vq(torch.zeros(1, 96, 4)), lambda: pkg.metrics.psnr_batch(x, x),
                 lambda: pkg.metrics.stsim_batch(x, x), lambda: pkg.metrics.psnr_3k_aligned_batch(x, x)):
        with pytest.raises(pkg.B2CError):
            call()


def test_state_dict_keys_match_reference_layout():
    ref = cases.build_reference_style_model(proposed.ProposedEval, dict(books=2, K=128))
    net = pkg.build_proposed(2, 128)
    res = net.load_state_dict(ref.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in ref.state_dict().items():
        assert net.state_dict()[k].shape == v.shape, k
    assert net.out_len(24000) == 23992 and net.latent_len(24000) == 75
    assert net.out_len(8000) == 7992


def test_arena_reuses_and_merges():
    a = engine.Arena()
    x = a.alloc(1000)
    y = a.alloc(5000)
    z = a.alloc(300)
    a.free(y)
    w = a.alloc(4000)
    assert w == y
    a.free(x); a.free(w); a.free(z)
    assert a.top == 0 and not a.live
    assert a.peak >= 6300


def test_interface_errors():
    net = pkg.build_proposed(1, 128)
    with pytest.raises(RuntimeError):
        net.A_ENC.block[0](torch.zeros(1, 1, 8))   # inner layers are fused, not callable
    assert modules.encoder_out_len(net.A_ENC, 24000) == 75


def test_bitstream_round_trip_and_rate():
    """Packed code indices (the payload of ProposedEval.decode_indices): exact round trip, ceil(log2 K) bits per
    index, and the reference's analytic bitrate (Training/compare_dacvsproposal_5.py:372-373) for power-of-two K."""
    import torch
    import multimodal_vqvae_compression_audio_tactile_b200 as pkg
    g = torch.Generator().manual_seed(5)
    for K, books in ((128, 10), (512, 8), (1024, 3), (300, 2), (2, 1), (1, 1)):
        idx = torch.randint(0, K, (3, books, 75), generator=g)
        buf = pkg.pack_indices(idx, K)
        assert len(buf) == pkg.packed_bytes(idx.shape, K) == (idx.numel() * pkg.bits_per_index(K) + 7) // 8
        back = pkg.unpack_indices(buf, idx.shape, K)
        assert back.dtype == torch.int32 and torch.equal(back.long(), idx)
        if K & (K - 1) == 0 and K > 1:
            per_frame_bits = 8 * len(pkg.pack_indices(idx[:1], K))
            assert abs(per_frame_bits / 1000.0 - pkg.estimated_kbps(books, K)) < 8 / 1000.0   # byte padding only
    assert pkg.pack_indices(torch.zeros(0, dtype=torch.int64), 512) == b""
    with pytest.raises(ValueError):
        pkg.pack_indices(torch.tensor([512]), 512)
    with pytest.raises(ValueError):
        pkg.unpack_indices(b"\x00", (3,), 512)
