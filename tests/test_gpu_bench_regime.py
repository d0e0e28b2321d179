"""GPU: parity in the regime bench.py measures (VERDICT r1 "what's weak" 2) — full geometry, full-length chunks, M >= 163
decode rows (two GEMM m-tiles, two decode branches, graph replay), batched T = 605 tensor-core vocoder — against the CPU
oracle.  The oracle decodes ~20 tokens/s at this geometry, so whole-chunk checks verify the ENGINE's token sequence with one
teacher-forced oracle prefill over it (1.7 s) instead of 605 oracle decode steps: if at every position the oracle's greedy
choice given the engine's prefix equals the engine's token, the oracle's own free-running decode yields the same ids."""
import numpy as np
import pytest
import torch

from auralis_b200.native import Sampling
from oracle import xtts_oracle as O
from conftest import text_ids, _make_engine

pytestmark = pytest.mark.gpu


NO_STOP = 4095          # an id outside the audio vocabulary: the chunk runs to max_tokens (random-init weights stop at random)


def _greedy_replay(orc, dims, cond, ids, toks, penalty=5.0):
    """oracle logits [n,V] for the prefix `toks` (one prefill) -> the oracle's greedy pick at every position + latents"""
    lg, lat = orc.teacher_forced(cond, ids, list(toks))
    sp = O.SamplingParams(temperature=0.0, repetition_penalty=penalty, max_tokens=len(toks), stop_token=NO_STOP)
    seen = O.prompt_seen_set(dims.gpt)
    picks = []
    for k in range(len(toks)):
        picks.append(O.sample_token(lg[k].clone(), seen, sp, 0, k))
        seen.add(int(toks[k]))
    return np.array(picks), lg.numpy(), lat.numpy()


@pytest.mark.parametrize("stop", ["model", "never"])
def test_cfg1_full_chunk_fp32_greedy_bit_exact(engine_full, dims_full, state_full, speakers_full, stop):
    """BASELINE cfg1 at its real size: one 64-char-like request, greedy, fp32 parity mode, a complete chunk — once ending
    where the (random-init) model emits the stop token, once with the stop token disabled so that all 605 tokens are decoded.
    Token ids bit-exact against the oracle, latents 1e-3, waveform within the fp32 tolerance (2e-4)."""
    orc = O.GPTOracle(state_full[0], state_full[1], dims_full)
    g = dims_full.gpt
    ids = text_ids(dims_full, 22, 64)                       # ~64 characters of BPE ids
    sp = Sampling(temperature=0.0, repetition_penalty=5.0, max_tokens=g.max_audio_tokens,
                  stop_token=g.stop_audio_token if stop == "model" else NO_STOP)
    r, toks, wav, lat = engine_full.run_batch([(1, ids, 0, sp)], timeout_s=600, want_latents=True)[1]
    assert r.n_tokens == len(toks) and 1 <= len(toks) <= g.max_audio_tokens
    assert stop == "model" or len(toks) == g.max_audio_tokens
    picks, lg, olat = _greedy_replay(orc, dims_full, speakers_full[0][0], ids, toks)
    bad = [(k, int(a), int(b), float(lg[k][b] - lg[k][a])) for k, (a, b) in enumerate(zip(toks, picks)) if a != b]
    print("cfg1 full chunk:", len(toks), "tokens, disagreements", bad[:4], "latent max err", float(np.abs(lat - olat).max()))
    assert bad == []
    assert np.abs(lat - olat).max() < 1e-3
    ewav = O.vocoder(torch.from_numpy(olat), speakers_full[0][1], state_full[1], dims_full).numpy()
    assert wav.shape == ewav.shape == (dims_full.voc.n_samples(len(toks)),)
    err = float(np.abs(wav - ewav).max())
    print("cfg1 full chunk waveform max-abs err", err)
    assert err < 2e-4


@pytest.fixture(scope="module")
def engine_wide_bf16(dims_full, state_full, speakers_full):
    eng = _make_engine(dims_full, state_full, speakers_full, 1, max_batch=192)
    yield eng
    eng.close()


def test_decode_at_bench_batch_rows_in_both_m_tiles(engine_wide_bf16, dims_full, state_full, speakers_full):
    """168 concurrent chunks (the bench runs 163): decode rows span two 128-row GEMM tiles (A box + OOB fill), the step runs
    as two branches and from its CUDA graph after two eager steps.  Rows on both sides of the tile and branch boundaries are
    compared with the oracle: first greedy divergence (if any) must be a near-tie, latents within the bf16 tolerance."""
    eng = engine_wide_bf16
    orc = O.GPTOracle(state_full[0], state_full[1], dims_full)
    g = dims_full.gpt
    n_seq, n_tok = 168, 12
    jobs = []
    for i in range(n_seq):
        sp = Sampling(temperature=0.0, repetition_penalty=5.0, max_tokens=n_tok, stop_token=NO_STOP)
        jobs.append((i, text_ids(dims_full, 10 + (7 * i) % 70, 1000 + i), i % 3, sp))
    res = eng.run_batch(jobs, timeout_s=300, want_latents=True)
    assert eng.stats().decode_steps >= n_tok - 1
    worst = 0.0
    for i in (0, 1, 83, 84, 127, 128, 129, 167):
        sid, ids, spk, _ = jobs[i]
        r, toks, wav, lat = res[sid]
        assert r.n_tokens == n_tok and np.isfinite(wav).all()
        picks, lg, olat = _greedy_replay(orc, dims_full, speakers_full[spk][0], ids, toks)
        # teacher-forced on the engine's own tokens: every position is comparable, any flip must be a near-tie
        for k, (a, b) in enumerate(zip(toks, picks)):
            if a != b:
                assert float(lg[k][b] - lg[k][a]) < 0.15, (i, k, int(a), int(b), float(lg[k][b] - lg[k][a]))
        worst = max(worst, float(np.abs(lat - olat).max()))
        assert np.abs(lat - olat).max() < 0.08, (i, float(np.abs(lat - olat).max()))
    print("M=168 decode: worst latent error over the sampled rows", worst)


def test_batched_full_length_tc_vocoder_vs_oracle(engine_wide_bf16, dims_full, state_full, speakers_full):
    """8 chunks x 605 tokens finish together and are vocoded in one batched tensor-core launch sequence (the bench's vocoder
    shape): every waveform equals the single-chunk xtts_vocode of its latents bit for bit, and two of them are compared with
    the fp32 oracle vocoder (tolerance 2e-2 abs, mse 1e-5: fp16 operands, the precision the reference runs at on GPU)."""
    eng = engine_wide_bf16
    g = dims_full.gpt
    jobs = []
    for i in range(8):
        sp = Sampling(temperature=0.75, top_p=0.85, top_k=50, repetition_penalty=5.0, max_tokens=g.max_audio_tokens,
                      stop_token=NO_STOP, seed=11, seq_seed=i)
        jobs.append((i, text_ids(dims_full, 30 + i, 50 + i), i % 3, sp))
    res = eng.run_batch(jobs, timeout_s=300, want_latents=True)
    for i in range(8):
        r, toks, wav, lat = res[i]
        assert wav.shape == (dims_full.voc.n_samples(r.n_tokens),)
        np.testing.assert_array_equal(wav, eng.vocode(lat, i % 3))
    for i in (0, 7):
        r, toks, wav, lat = res[i]
        ewav = O.vocoder(torch.from_numpy(lat), speakers_full[i % 3][1], state_full[1], dims_full).numpy()
        err, mse = float(np.abs(wav - ewav).max()), float(np.mean((wav - ewav) ** 2))
        print(f"batched tc vocoder, chunk {i}: {r.n_tokens} tokens, max-abs err {err:.3e}, mse {mse:.3e}")
        assert err < 2e-2 and mse < 1e-5


def test_mixed_length_batch_tc_vocoder_vs_oracle(engine_wide_bf16, dims_full, state_full, speakers_full):
    """ragged vocoder batches at full geometry: chunk lengths ~ U(150, 605) (what real weights produce: the stop token lands
    at a different step per chunk) — every waveform equals the single-chunk result; the shortest and longest also go against
    the oracle."""
    eng = engine_wide_bf16
    g = dims_full.gpt
    rng = np.random.RandomState(5)
    lens = rng.randint(150, 606, size=24).tolist()
    jobs = []
    for i, n in enumerate(lens):
        sp = Sampling(temperature=0.75, top_p=0.85, top_k=50, repetition_penalty=5.0, max_tokens=n, stop_token=NO_STOP,
                      seed=12, seq_seed=i)
        jobs.append((i, text_ids(dims_full, 20 + i, 80 + i), i % 3, sp))
    res = eng.run_batch(jobs, timeout_s=300, want_latents=True)
    for i, n in enumerate(lens):
        r, toks, wav, lat = res[i]
        assert r.n_tokens == n and wav.shape == (dims_full.voc.n_samples(n),)
        np.testing.assert_array_equal(wav, eng.vocode(lat, i % 3))
    for i in (int(np.argmin(lens)), int(np.argmax(lens))):
        r, toks, wav, lat = res[i]
        ewav = O.vocoder(torch.from_numpy(lat), speakers_full[i % 3][1], state_full[1], dims_full).numpy()
        err = float(np.abs(wav - ewav).max())
        print(f"ragged tc vocoder, chunk of {lens[i]} tokens: max-abs err {err:.3e}")
        assert err < 2e-2
