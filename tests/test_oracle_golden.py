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
