"""CPU: the oracle restatement against the golden vectors the REFERENCE's own modules produced
(tests/golden/make_golden.py) — this is the oracle's pin."""
import os

import numpy as np
import pytest
import torch

from auralis_b200.config import XTTSDims
from oracle import xtts_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _state(name, request):
    return request.getfixturevalue(f"state_{name}"), (XTTSDims.small() if name == "small" else XTTSDims.full())


@pytest.mark.parametrize("name", ["small", "full"])
def test_vocoder_matches_reference_golden(name, request):
    (gs, cs), dims = _state(name, request)
    z = np.load(os.path.join(GOLD, f"vocoder_{name}.npz"))
    wav = O.vocoder(torch.from_numpy(z["latents"]), torch.from_numpy(z["g"]), cs, dims).numpy()
    assert wav.shape == z["wav"].shape == (dims.voc.n_samples(z["latents"].shape[0]),)
    assert np.abs(z["wav"]).max() > 0.05                      # a live signal, not silence / saturation
    np.testing.assert_allclose(wav, z["wav"], atol=2e-6, rtol=0)


@pytest.mark.parametrize("name", ["small", "full"])
def test_conditioning_matches_reference_golden(name, request):
    (gs, cs), dims = _state(name, request)
    z = np.load(os.path.join(GOLD, f"conditioning_{name}.npz"))
    wav22 = torch.from_numpy(z["wav22"])
    mel = O.mel_cloning(wav22, cs["mel_stats"], dims.cond.n_mels)
    np.testing.assert_allclose(mel.numpy(), z["mel"], atol=2e-5, rtol=0)
    h = O.cond_encoder(torch.from_numpy(z["mel"]), cs, dims)
    ref_h = z["cond_enc"]
    np.testing.assert_allclose(h.numpy()[:, : ref_h.shape[1]], ref_h, atol=2e-5, rtol=0)
    lat = O.perceiver(h.t(), cs, dims)
    np.testing.assert_allclose(lat.numpy(), z["perceiver"], atol=5e-5, rtol=0)
    dv = O.speaker_embedding(torch.from_numpy(z["wav16"]), cs, dims)
    np.testing.assert_allclose(dv.numpy(), z["dvector"], atol=2e-6, rtol=0)
    assert abs(float(np.linalg.norm(z["dvector"])) - 1.0) < 1e-5


def test_interpolate_length_rule():
    """T latents -> floor(floor(4T)*24000/22050) z-frames -> x256 samples (605 -> 2634 -> 674304)."""
    v = XTTSDims.full().voc
    assert v.z_frames(605) == 2634 and v.n_samples(605) == 674304
    for T in (1, 2, 7, 64, 320):
        lat = torch.zeros(T, 8)
        assert O.interp_latents(lat, v).shape[1] == v.z_frames(T)


def test_vocoder_lookahead_covers_receptive_field(state_small):
    """The engine's first-audio early emit vocodes a chunk's first n + 6 latent frames and keeps the samples of the first n
    (kEarlyLookahead in engine.cu).  With the oracle: prefix and whole-chunk waveforms agree to 1e-6 on those samples — the
    vocoder's receptive field (conv_pre + the k = 11, d = 5 resblock of the first stage) reaches ~3 frames ahead, so 3 frames
    are the minimum that passes and 6 leave margin.  Same kernels and dilations as the full geometry."""
    dims = XTTSDims.small()
    cs = state_small[1]
    g = torch.Generator().manual_seed(3)
    lat = torch.randn(60, dims.voc.in_dim, generator=g)
    dv = torch.nn.functional.normalize(torch.randn(dims.voc.d_vector, generator=g), dim=0)
    full = O.vocoder(lat, dv, cs, dims).numpy()
    for n in (10, 16, 24):
        keep = dims.voc.n_samples(n)
        part6 = O.vocoder(lat[: n + 6], dv, cs, dims).numpy()
        np.testing.assert_allclose(part6[:keep], full[:keep], atol=1e-6, rtol=0)
        part3 = O.vocoder(lat[: n + 3], dv, cs, dims).numpy()
        np.testing.assert_allclose(part3[:keep], full[:keep], atol=1e-6, rtol=0)
        part1 = O.vocoder(lat[: n + 1], dv, cs, dims).numpy()
        assert np.abs(part1[:keep] - full[:keep]).max() > 1e-4          # one frame of lookahead is NOT enough
