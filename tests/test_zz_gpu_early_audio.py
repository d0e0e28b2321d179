"""GPU, EXPERIMENTAL (written after the round's GPU budget was spent: never run on a B200 yet).  Skipped unless
XTTS_TEST_EXPERIMENTAL=1.  First-audio early emit (`xtts_sampling.early_tokens`): the partial piece plus the
final piece must reproduce the unsplit chunk — token ids identical, waveform bit-identical in fp32 mode and in bf16/fp16
tensor-core mode (same arithmetic on the same inputs; only positions beyond the vocoder's receptive field are kept)."""
import os

import numpy as np
import pytest

from auralis_b200.native import Sampling
from conftest import text_ids

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("XTTS_TEST_EXPERIMENTAL") != "1", reason="experimental: set XTTS_TEST_EXPERIMENTAL=1")]


def _run(eng, dims, early, n_seq=3, max_tokens=40):
    jobs = []
    for i in range(n_seq):
        sp = Sampling(temperature=0.0, repetition_penalty=5.0, max_tokens=max_tokens, stop_token=dims.gpt.stop_audio_token,
                      seed=7, seq_seed=i, early_tokens=early if i != 1 else 0)         # sequence 1 never asks for it
        jobs.append((100 + i, text_ids(dims, 9 + 3 * i, i), i % 2, sp))
    res = eng.run_batch(jobs, timeout_s=120)
    return res, dict(eng.last_partials)


@pytest.mark.parametrize("which", ["engine_small", "engine_small_bf16"])
def test_partial_plus_final_equals_unsplit(request, dims_small, which):
    eng = request.getfixturevalue(which)
    ref, none = _run(eng, dims_small, early=0)
    assert not none
    got, parts = _run(eng, dims_small, early=10)
    assert sorted(parts) == [100, 102]                                   # only the sequences that asked
    for sid in (100, 101, 102):
        r_ref, toks_ref, wav_ref, _ = ref[sid]
        r, toks, wav, _ = got[sid]
        np.testing.assert_array_equal(toks, toks_ref)
        if sid in parts:
            pr, ptoks, pwav = parts[sid]
            assert pr.status == 1 and pr.n_tokens == 10 and pr.n_samples == dims_small.voc.n_samples(10)
            np.testing.assert_array_equal(ptoks, toks_ref[:10])
            assert pr.t_done <= r.t_done
            np.testing.assert_array_equal(np.concatenate([pwav, wav]), wav_ref)
        else:
            np.testing.assert_array_equal(wav, wav_ref)


def test_short_sequences_never_split(engine_small, dims_small):
    """a chunk that ends before early_tokens + lookahead frames exist is delivered whole"""
    got, parts = _run(engine_small, dims_small, early=10, max_tokens=12)
    assert not parts and all(got[s][0].n_tokens == 12 for s in got)


@pytest.mark.parametrize("which,dims_name,T,n", [("engine_small", "dims_small", 40, 20), ("engine_small_bf16", "dims_small", 40, 20),
                                                   ("engine_full_bf16", "dims_full", 605, 58)])
def test_vocoder_prefix_property(request, which, dims_name, T, n):
    """Size-independent property the early emit rests on, checked through xtts_vocode alone: the waveform of the first n
    latent frames does not depend on frames more than 6 ahead — vocode(lat[:n+6])[:samples(n)] == vocode(lat)[:samples(n)]."""
    import torch
    eng, dims = request.getfixturevalue(which), request.getfixturevalue(dims_name)
    g = torch.Generator().manual_seed(77)
    lat = torch.randn(T, dims.voc.in_dim, generator=g).numpy()
    full = eng.vocode(lat, 0)
    part = eng.vocode(lat[: n + 6], 0)
    keep = dims.voc.n_samples(n)
    assert part.shape[0] == dims.voc.n_samples(n + 6) and keep < part.shape[0]
    np.testing.assert_array_equal(part[:keep], full[:keep])
