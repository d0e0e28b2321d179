"""GPU: GPT-2 acoustic model — prefill, paged decode, latent capture, end-to-end chunks vs the oracle."""
import numpy as np
import pytest
import torch

from auralis_b200.native import Sampling
from oracle import xtts_oracle as O
from conftest import text_ids

pytestmark = pytest.mark.gpu


def _orc(dims, state):
    return O.GPTOracle(state[0], state[1], dims)


def _margin_report(ref_logits, got_tokens, ref_tokens):
    bad = []
    for k, (a, b) in enumerate(zip(got_tokens, ref_tokens)):
        if a != b:
            bad.append((k, int(a), int(b), float(ref_logits[k][b] - ref_logits[k][a])))
    return bad


@pytest.mark.parametrize("n_text,n_audio", [(3, 0), (12, 9), (30, 33)])
def test_prefill_small(engine_small, dims_small, state_small, speakers_small, n_text, n_audio):
    orc = _orc(dims_small, state_small)
    ids = text_ids(dims_small, n_text, n_text)
    rng = np.random.RandomState(n_audio)
    aud = rng.randint(0, dims_small.gpt.start_audio_token, size=n_audio).tolist()
    hid, logits, lat = engine_small.gpt_prefill(ids, 2, aud, want_hidden=True)
    rows = [orc.prompt_rows(speakers_small[2][0], ids)]
    for k, t in enumerate(aud[:-1], start=1):
        rows.append(orc.audio_row(t, k)[None])
    with torch.no_grad():
        h, _ = orc.forward_rows(torch.cat(rows, 0))
        n = max(1, n_audio)
        lg, lt = orc.head(h[-n:])
    np.testing.assert_allclose(hid, h.numpy(), atol=2e-4, rtol=0)
    np.testing.assert_allclose(logits, lg.numpy(), atol=2e-4, rtol=0)
    np.testing.assert_allclose(lat, lt.numpy(), atol=2e-4, rtol=0)


def test_teacher_forced_decode_small(engine_small, dims_small, state_small, speakers_small):
    """paged-KV decode path, step by step, against the oracle's full-prefill logits (no feedback drift)."""
    orc = _orc(dims_small, state_small)
    g = dims_small.gpt
    ids = text_ids(dims_small, 17, 4)
    rng = np.random.RandomState(8)
    forced = rng.randint(0, g.start_audio_token, size=40).tolist()      # crosses a 32-token page boundary
    sp = Sampling(temperature=0.0, repetition_penalty=5.0, max_tokens=40, stop_token=g.stop_audio_token)
    logits, lat, sampled = engine_small.gpt_teacher_forced(ids, 0, forced, sp)
    lg, lt = orc.teacher_forced(speakers_small[0][0], ids, forced)
    np.testing.assert_allclose(logits, lg.numpy(), atol=3e-4, rtol=0)
    np.testing.assert_allclose(lat, lt.numpy(), atol=3e-4, rtol=0)
    # sampler inside the loop: greedy + penalty over prompt ∪ forced history
    seen = O.prompt_seen_set(g)
    exp = []
    for k in range(40):
        z = O.apply_repetition_penalty(lg[k].clone(), seen, 5.0)
        exp.append(int(torch.argmax(z)))
        seen.add(forced[k])
    assert _margin_report(lg.numpy(), sampled, exp) == []


def test_e2e_greedy_small(engine_small, dims_small, state_small, speakers_small):
    """cfg1-style: chunks of different lengths/speakers decoded concurrently (continuous batching), greedy:
    token ids bit-exact vs the oracle, waveform within fp32 tolerance."""
    orc = _orc(dims_small, state_small)
    g = dims_small.gpt
    jobs, exp = [], {}
    for i, (n_text, spk, mt) in enumerate([(5, 0, 48), (11, 1, 20), (30, 2, 48), (2, 0, 7), (19, 1, 33), (8, 2, 48),
                                           (14, 0, 41), (3, 1, 48), (27, 2, 5), (9, 0, 48)]):
        ids = text_ids(dims_small, n_text, 100 + i)
        sp = Sampling(temperature=0.0, repetition_penalty=5.0, max_tokens=mt, stop_token=g.stop_audio_token, seq_seed=i)
        jobs.append((1000 + i, ids, spk, sp))
        osp = O.SamplingParams(temperature=0.0, repetition_penalty=5.0, max_tokens=mt, stop_token=g.stop_audio_token)
        toks, lats, lg = orc.generate(speakers_small[spk][0], ids, osp, return_logits=True)
        wav = O.vocoder(lats, speakers_small[spk][1], state_small[1], dims_small)
        exp[1000 + i] = (toks, lats.numpy(), wav.numpy(), lg.numpy())
    res = engine_small.run_batch(jobs, timeout_s=120, want_latents=True)
    for sid, (r, toks, wav, lat) in res.items():
        etoks, elat, ewav, elg = exp[sid]
        assert r.n_tokens == len(etoks)
        bad = _margin_report(elg, toks, etoks)
        assert bad == [], (sid, bad)
        np.testing.assert_allclose(lat, elat, atol=5e-4, rtol=0)
        assert wav.shape == ewav.shape
        assert np.abs(wav - ewav).max() < 5e-4, (sid, np.abs(wav - ewav).max())


def test_stop_token_ends_sequence(engine_small, dims_small, state_small, speakers_small):
    orc = _orc(dims_small, state_small)
    g = dims_small.gpt
    ids = text_ids(dims_small, 6, 77)
    osp = O.SamplingParams(temperature=0.0, repetition_penalty=5.0, max_tokens=30, stop_token=g.stop_audio_token)
    toks, _ = orc.generate(speakers_small[0][0], ids, osp)
    stop = toks[6]                                 # make the 7th greedy token the stop token
    sp = Sampling(temperature=0.0, repetition_penalty=5.0, max_tokens=30, stop_token=stop)
    res = engine_small.run_batch([(5, ids, 0, sp)], timeout_s=60)
    r, got, wav, _ = res[5]
    first = toks.index(stop)
    assert list(got) == toks[: first + 1]
    assert wav.shape[0] == dims_small.voc.n_samples(first + 1)


def test_seeded_sampling_small(engine_small, dims_small, state_small, speakers_small):
    """T=0.75/top_p=0.85/top_k=50 with the shared Philox stream: same tokens as the oracle (teacher-forced so a
    single ulp-level flip cannot cascade)."""
    orc = _orc(dims_small, state_small)
    g = dims_small.gpt
    ids = text_ids(dims_small, 10, 5)
    osp = O.SamplingParams(temperature=0.75, top_p=0.85, top_k=50, repetition_penalty=5.0, max_tokens=48,
                           stop_token=g.stop_audio_token, seed=99)
    toks, _ = orc.generate(speakers_small[1][0], ids, osp, seq_seed=3)
    sp = Sampling(temperature=0.75, top_p=0.85, top_k=50, repetition_penalty=5.0, max_tokens=len(toks),
                  stop_token=g.stop_audio_token, seed=99, seq_seed=3)
    _, _, sampled = engine_small.gpt_teacher_forced(ids, 1, toks, sp)
    agree = int((np.array(toks) == sampled).sum())
    assert agree >= len(toks) - 1, (agree, len(toks))


def test_prefill_full_size(engine_full, dims_full, state_full, speakers_full):
    """30 x 1024 geometry, short prompt + 6 audio rows."""
    orc = _orc(dims_full, state_full)
    ids = text_ids(dims_full, 10, 1)
    aud = [17, 900, 3, 511, 42, 640]
    _, logits, lat = engine_full.gpt_prefill(ids, 0, aud)
    lg, lt = orc.teacher_forced(speakers_full[0][0], ids, aud)
    np.testing.assert_allclose(logits, lg.numpy(), atol=1e-3, rtol=0)
    np.testing.assert_allclose(lat, lt.numpy(), atol=1e-3, rtol=0)


def test_e2e_greedy_full_size(engine_full, dims_full, state_full, speakers_full):
    """cfg1 shape at full geometry with a bounded token budget so the CPU oracle finishes in seconds."""
    orc = _orc(dims_full, state_full)
    g = dims_full.gpt
    jobs, exp = [], {}
    for i, (n_text, spk) in enumerate([(20, 0), (7, 1)]):
        ids = text_ids(dims_full, n_text, 300 + i)
        sp = Sampling(temperature=0.0, repetition_penalty=5.0, max_tokens=24, stop_token=g.stop_audio_token)
        jobs.append((i, ids, spk, sp))
        osp = O.SamplingParams(temperature=0.0, repetition_penalty=5.0, max_tokens=24, stop_token=g.stop_audio_token)
        toks, lats, lg = orc.generate(speakers_full[spk][0], ids, osp, return_logits=True)
        exp[i] = (toks, O.vocoder(lats, speakers_full[spk][1], state_full[1], dims_full).numpy(), lg.numpy())
    res = engine_full.run_batch(jobs, timeout_s=300)
    for sid, (r, toks, wav, _) in res.items():
        etoks, ewav, elg = exp[sid]
        assert _margin_report(elg, toks, etoks) == [], sid
        assert np.abs(wav - ewav).max() < 1e-3, np.abs(wav - ewav).max()


@pytest.mark.parametrize("top_k", [50, 0])
def test_sampler_full_vocab(engine_full, dims_full, top_k):
    """V = 1026 at the reference's default sampling (top_k 50 -> radix-select path; 0 -> full 2048-wide sort)."""
    V = dims_full.gpt.n_audio_tokens
    rng = np.random.RandomState(11)
    agree = total = 0
    for step in range(4):
        logits = (rng.randn(4, V) * 3.0).astype(np.float32)
        seen = (rng.rand(4, V) < 0.05).astype(np.uint8)
        sp = Sampling(temperature=0.75, top_p=0.85, top_k=top_k, repetition_penalty=5.0, seed=99 + step, seq_seed=3,
                      stop_token=dims_full.gpt.stop_audio_token)
        got = engine_full.debug_sample(logits, seen, sp, step=step)
        osp = O.SamplingParams(temperature=sp.temperature, top_p=sp.top_p, top_k=sp.top_k,
                               repetition_penalty=sp.repetition_penalty, seed=sp.seed)
        exp = np.array([O.sample_token(torch.from_numpy(logits[b].copy()), set(np.nonzero(seen[b])[0].tolist()),
                                       osp, sp.seq_seed + b, step) for b in range(4)])
        agree += int((got == exp).sum()); total += got.size
    assert agree >= total - 1, (agree, total)


# ------------------------------------------------------------------------------------------------
# bf16 / tcgen05 fast mode: same graph, looser tolerance (bf16 operands, fp32 accumulate)
# ------------------------------------------------------------------------------------------------
def test_bf16_prefill_small(engine_small_bf16, dims_small, state_small, speakers_small):
    orc = _orc(dims_small, state_small)
    ids = text_ids(dims_small, 12, 12)
    aud = [5, 100, 77, 3, 64, 9, 31, 2]
    _, logits, lat = engine_small_bf16.gpt_prefill(ids, 2, aud)
    lg, lt = orc.teacher_forced(speakers_small[2][0], ids, aud)
    err = np.abs(logits - lg.numpy()).max()
    print("bf16 logits max err", err, "logit std", float(lg.std()))
    assert err < 0.05 * max(1.0, float(lg.abs().max()))
    assert np.abs(lat - lt.numpy()).max() < 0.08


def test_bf16_teacher_forced_and_e2e_small(engine_small_bf16, dims_small, state_small, speakers_small):
    orc = _orc(dims_small, state_small)
    g = dims_small.gpt
    ids = text_ids(dims_small, 17, 4)
    osp = O.SamplingParams(temperature=0.0, repetition_penalty=5.0, max_tokens=40, stop_token=g.stop_audio_token)
    toks, lats, lg = orc.generate(speakers_small[0][0], ids, osp, return_logits=True)
    sp = Sampling(temperature=0.0, repetition_penalty=5.0, max_tokens=40, stop_token=g.stop_audio_token)
    logits, lat, sampled = engine_small_bf16.gpt_teacher_forced(ids, 0, toks, sp)
    err = np.abs(logits - lg.numpy()).max()
    agree = int((sampled == np.array(toks)).sum())
    print("bf16 decode logits max err", err, "greedy agreement", agree, "/", len(toks))
    assert err < 0.05 * max(1.0, float(lg.abs().max()))
    # every disagreement must be a near-tie in the fp32 oracle
    for k, a, b, margin in _margin_report(lg.numpy(), sampled, toks):
        assert margin < 0.1, (k, a, b, margin)
    res = engine_small_bf16.run_batch([(1, ids, 0, sp)], timeout_s=60)
    assert res[1][0].n_tokens == 40 and np.isfinite(res[1][2]).all()


def test_bf16_full_size_decode_paths(engine_full_bf16, dims_full, state_full, speakers_full):
    """full geometry in fast mode: exercises the split-K out-proj/down-proj GEMMs + fused reduce-LayerNorm and the
    CUDA-graph replay of the decode step (graphs kick in after two eager steps)."""
    orc = _orc(dims_full, state_full)
    g = dims_full.gpt
    ids = text_ids(dims_full, 14, 9)
    osp = O.SamplingParams(temperature=0.0, repetition_penalty=5.0, max_tokens=20, stop_token=g.stop_audio_token)
    toks, lats, lg = orc.generate(speakers_full[1][0], ids, osp, return_logits=True)
    sp = Sampling(temperature=0.0, repetition_penalty=5.0, max_tokens=20, stop_token=g.stop_audio_token)
    logits, lat, sampled = engine_full_bf16.gpt_teacher_forced(ids, 1, toks, sp)          # eager path
    err = np.abs(logits - lg.numpy()).max()
    print("bf16 full teacher-forced logits max err", err, "logit abs max", float(lg.abs().max()))
    assert err < 0.05 * max(1.0, float(lg.abs().max()))
    for k, a, b, margin in _margin_report(lg.numpy(), sampled, toks):
        assert margin < 0.15, (k, a, b, margin)
    # free-running batch (graph path): same tokens as the eager teacher-forced run unless a near-tie flips
    res = engine_full_bf16.run_batch([(i, ids, 1, sp) for i in range(3)], timeout_s=120, want_latents=True)
    for i in range(3):
        r, got, wav, glat = res[i]
        assert r.n_tokens == 20 and np.isfinite(wav).all()
        np.testing.assert_array_equal(got, res[0][1])                        # deterministic across batch slots
        bad = _margin_report(lg.numpy(), got, toks)
        # free-running: after the first flip the contexts differ, so only the first divergence is meaningful and it
        # must be a near-tie of the fp32 oracle
        assert not bad or bad[0][3] < 0.15, bad[:3]
    engine_full_bf16.set_option("splitk", 0)
    engine_full_bf16.set_option("cuda_graphs", 0)
    res2 = engine_full_bf16.run_batch([(7, ids, 1, sp)], timeout_s=120)
    engine_full_bf16.set_option("splitk", 1)
    engine_full_bf16.set_option("cuda_graphs", 1)
    agree = int((res2[7][1] == res[0][1]).sum())
    print("split-K+graphs vs plain bf16 token agreement", agree, "/ 20")
    first = next((k for k in range(20) if res2[7][1][k] != res[0][1][k]), 20)
    assert first >= 5 or _margin_report(lg.numpy(), res2[7][1][: first + 1], toks[: first + 1])[-1][3] < 0.15


def test_bf16_decode_microbatch_branches_match_single_branch(engine_full_bf16, dims_full):
    """The decode step may run its rows as concurrent branches on forked streams (engine option "microbatches").
    Rows are independent, so tokens and waveforms must not depend on how the step was split."""
    g = dims_full.gpt
    jobs = []
    for i in range(4):
        sp = Sampling(temperature=0.75, top_p=0.85, top_k=50, repetition_penalty=5.0, max_tokens=14, seed=77, seq_seed=i,
                      stop_token=g.stop_audio_token)
        jobs.append((i, text_ids(dims_full, 9 + 3 * i, 40 + i), i % 3, sp))
    out = {}
    try:
        engine_full_bf16.set_option("decode_chain", 0)         # the fused chain kernel always runs the rows as one branch
        engine_full_bf16.set_option("microbatch_min_rows", 2)
        for nmb in (1, 2, 3):
            engine_full_bf16.set_option("microbatches", nmb)
            res = engine_full_bf16.run_batch(jobs, timeout_s=120)
            out[nmb] = {sid: (list(toks), wav) for sid, (_, toks, wav, _) in res.items()}
    finally:
        engine_full_bf16.set_option("microbatches", 2)
        engine_full_bf16.set_option("microbatch_min_rows", 48)
    for nmb in (2, 3):
        for sid in out[1]:
            assert out[nmb][sid][0] == out[1][sid][0], (nmb, sid)
            np.testing.assert_allclose(out[nmb][sid][1], out[1][sid][1], rtol=0, atol=1e-6)


def test_bf16_decode_chain_matches_unfused(engine_full_bf16, dims_full, state_full, speakers_full):
    """Engine option "decode_chain": the persistent kernel that fuses out-proj, residual+LN2, fc+gelu, down-proj,
    residual+LN1 and the next QKV projection must reproduce the one-launch-per-op decode step (same bf16 operands, same
    split-K order; only the LayerNorm reduction tree differs) and stay within the fast-mode tolerance of the oracle."""
    orc = _orc(dims_full, state_full)
    g = dims_full.gpt
    ids = text_ids(dims_full, 21, 77)
    osp = O.SamplingParams(temperature=0.0, repetition_penalty=5.0, max_tokens=16, stop_token=g.stop_audio_token)
    toks, lats, lg = orc.generate(speakers_full[2][0], ids, osp, return_logits=True)
    sp = Sampling(temperature=0.0, repetition_penalty=5.0, max_tokens=16, stop_token=g.stop_audio_token)
    got = {}
    try:
        for chain in (1, 0):
            engine_full_bf16.set_option("decode_chain", chain)
            logits, lat, sampled = engine_full_bf16.gpt_teacher_forced(ids, 2, toks, sp)
            jobs = [(i, text_ids(dims_full, 8 + 5 * i, 60 + i), i % 3,
                     Sampling(temperature=0.75, top_p=0.85, top_k=50, repetition_penalty=5.0, max_tokens=12, seed=5, seq_seed=i,
                              stop_token=g.stop_audio_token)) for i in range(4)]
            res = engine_full_bf16.run_batch(jobs, timeout_s=120, want_latents=True)
            got[chain] = (logits, lat, {sid: (list(t), l) for sid, (_, t, _, l) in res.items()})
    finally:
        engine_full_bf16.set_option("decode_chain", 0)
    err = np.abs(got[1][0] - lg.numpy()).max()
    d_logits = np.abs(got[1][0] - got[0][0]).max()
    d_lat = np.abs(got[1][1] - got[0][1]).max()
    print("chain vs oracle logits max err", err, "| chain vs unfused: logits", d_logits, "latents", d_lat)
    assert err < 0.05 * max(1.0, float(lg.abs().max()))
    # two bf16 evaluation orders (down-proj split 4 vs 8, LayerNorm reduction tree): each is ~0.05 from the fp32 oracle
    assert d_logits < 0.1 and d_lat < 0.1
    same = sum(got[1][2][sid][0] == got[0][2][sid][0] for sid in got[1][2])
    print("sampled sequences identical with / without the chain kernel:", same, "/ 4")
    assert same >= 2          # seeded sampling at T=0.75: a bf16 near-tie may flip one chain and everything after it


def test_capped_attention_grid_and_forced_gemm_tile(engine_small, engine_full_bf16, dims_small, dims_full):
    """Engine options "attn_ctas_per_sm" (decode attention walks several (row, head) items per CTA) and "gemm_bn" (tile
    width of the decode GEMMs) change scheduling only: tokens and waveforms must be unchanged."""
    for eng, dims, opts in ((engine_small, dims_small, [("attn_ctas_per_sm", -3)]),
                            (engine_full_bf16, dims_full, [("attn_ctas_per_sm", -5), ("gemm_bn", 64), ("gemm_deep_ring", 1), ("gemm_l2_prefetch", 1)])):
        g = dims.gpt
        jobs = [(i, text_ids(dims, 6 + 4 * i, 90 + i), i % 3,
                 Sampling(temperature=0.75, top_p=0.85, top_k=50, repetition_penalty=5.0, max_tokens=10, seed=3, seq_seed=i,
                          stop_token=g.stop_audio_token)) for i in range(4)]
        ref = eng.run_batch(jobs, timeout_s=120)
        for key, val in opts:
            try:
                eng.set_option(key, val)
                got = eng.run_batch(jobs, timeout_s=120)
            finally:
                eng.set_option(key, 0)
            for sid in ref:
                if key == "attn_ctas_per_sm":
                    assert list(got[sid][1]) == list(ref[sid][1]), (key, sid)
                    np.testing.assert_allclose(got[sid][2], ref[sid][2], rtol=0, atol=1e-6)
                else:                                   # another tile width = another bf16 summation order inside the MMA? no:
                    assert list(got[sid][1]) == list(ref[sid][1]), (key, sid)     # K order per output is unchanged


def test_bulk_copy_attention_matches_register_attention(engine_small_bf16, engine_full_bf16, engine_full_fp16, dims_small, dims_full):
    """Engine option "attn_bulk": the decode attention streams cache pages with cp.async.bulk into shared-memory sub-rings (one
    per consumer warp; a producer warp, mbarriers, L2 prefetch one item ahead, persistent CTAs that walk several (row, head)
    items, streaming starts before griddepcontrol.wait) instead of loading them into registers.  Same page -> warp assignment
    and operation order as the register kernel with as many warps ("attn_warps" = 4 / 8 / 16): tokens AND latents must be
    bit-identical, for every grid size / ring depth, across page boundaries (>= 3 pages of context) and in both 16-bit modes."""
    for eng, dims, n_tok in ((engine_small_bf16, dims_small, 40), (engine_full_bf16, dims_full, 70), (engine_full_fp16, dims_full, 70)):
        jobs = [(i, text_ids(dims, 6 + 4 * i, 120 + i), i % 3,
                 Sampling(temperature=0.75, top_p=0.85, top_k=50, repetition_penalty=5.0, max_tokens=n_tok, seed=11, seq_seed=i,
                          stop_token=4095, vocode=False)) for i in range(7)]
        try:
            for nw in (4, 8, 16):
                eng.set_option("attn_warps", nw); eng.set_option("attn_bulk", 0); eng.set_option("attn_ctas_per_sm", 0)
                ref = eng.run_batch(jobs, timeout_s=180, want_wav=False, want_latents=True)
                for l2_pages in (1, 3):                  # register kernel + cp.async.bulk.prefetch.L2 of later pages: a hint only
                    eng.set_option("attn_l2_pages", l2_pages)
                    got = eng.run_batch(jobs, timeout_s=180, want_wav=False, want_latents=True)
                    eng.set_option("attn_l2_pages", 0)
                    for sid in ref:
                        assert list(got[sid][1]) == list(ref[sid][1]), (nw, l2_pages, sid)
                        np.testing.assert_array_equal(got[sid][3], ref[sid][3])
                for bulk, stages, grid, l2 in ((1, 8, 0, 1), (2, 16, 0, 1), (1, 4, -3, 1), (1, 16, -5, 1), (1, 24, -1000, 1), (1, 16, 0, 0)):
                    eng.set_option("attn_bulk", bulk); eng.set_option("attn_stages", stages); eng.set_option("attn_ctas_per_sm", grid)
                    eng.set_option("attn_l2_ahead", l2)
                    got = eng.run_batch(jobs, timeout_s=180, want_wav=False, want_latents=True)
                    for sid in ref:
                        assert list(got[sid][1]) == list(ref[sid][1]), (nw, bulk, stages, grid, sid)
                        assert len(got[sid][1]) == n_tok
                        np.testing.assert_array_equal(got[sid][3], ref[sid][3])
        finally:
            for k, v in (("attn_bulk", 0), ("attn_stages", 8), ("attn_ctas_per_sm", 0), ("attn_l2_ahead", 1), ("attn_warps", 4), ("attn_l2_pages", 0)):
                eng.set_option(k, v)


def test_kernel_profile_graph_events(engine_small_bf16, dims_small):
    """Option "profile": the decode step is replayed from a graph that carries an event-record node on either side of
    every kernel; the family table must account for every decode launch and the tokens must not change."""
    g = dims_small.gpt
    jobs = [(i, text_ids(dims_small, 7 + i, 30 + i), i % 3,
             Sampling(temperature=0.0, repetition_penalty=5.0, max_tokens=12, stop_token=g.stop_audio_token, vocode=False))
            for i in range(3)]
    ref = engine_small_bf16.run_batch(jobs, timeout_s=60, want_wav=False)
    engine_small_bf16.set_option("profile", 1)
    try:
        got = engine_small_bf16.run_batch(jobs, timeout_s=60, want_wav=False)
        prof = engine_small_bf16.kernel_profile()
    finally:
        engine_small_bf16.set_option("profile", 0)
    for sid in ref:
        assert list(got[sid][1]) == list(ref[sid][1])
    L = g.layers
    att = prof["attn_decode_paged"]
    assert att["launches"] >= 9 * L and att["ms"] > 0 and att["bytes"] > 0, att
    assert prof["sample"]["launches"] >= 10 and prof["gemm_bf16_tcgen05"]["ms"] > 0


def test_device_timer_brackets_a_batch(engine_small, dims_small):
    """xtts_device_timer (bench.py's stopwatch): two CUDA events on the engine stream; the device-clock span of a batch is
    positive and does not exceed the host wall clock around the same calls by more than scheduling noise."""
    import time
    g = dims_small.gpt
    jobs = [(i, text_ids(dims_small, 6 + i, 40 + i), i % 3,
             Sampling(temperature=0.0, repetition_penalty=5.0, max_tokens=10, stop_token=g.stop_audio_token)) for i in range(3)]
    engine_small.run_batch(jobs, timeout_s=60)                       # warm
    t0 = time.perf_counter()
    engine_small.timer_start()
    engine_small.run_batch(jobs, timeout_s=60)
    ms = engine_small.timer_stop_ms()
    wall_ms = (time.perf_counter() - t0) * 1e3
    assert 0.0 < ms <= wall_ms + 5.0, (ms, wall_ms)


# ------------------------------------------------------------------------------------------------
# precision "fp16": the tcgen05 fast path with IEEE fp16 operands / KV instead of bf16 (a tensor-core mode
# closer to the parity mode — kind::f16 takes fp16 at the same rate, 11 significand bits instead of 8)
# ------------------------------------------------------------------------------------------------
def test_fp16_operand_mode_small(engine_small_fp16, dims_small, state_small, speakers_small):
    orc = _orc(dims_small, state_small)
    g = dims_small.gpt
    ids = text_ids(dims_small, 17, 4)
    osp = O.SamplingParams(temperature=0.0, repetition_penalty=5.0, max_tokens=40, stop_token=g.stop_audio_token)
    toks, lats, lg = orc.generate(speakers_small[0][0], ids, osp, return_logits=True)
    sp = Sampling(temperature=0.0, repetition_penalty=5.0, max_tokens=40, stop_token=g.stop_audio_token)
    logits, lat, sampled = engine_small_fp16.gpt_teacher_forced(ids, 0, toks, sp)
    err = np.abs(logits - lg.numpy()).max()
    print("fp16 decode logits max err", err, "greedy agreement", int((sampled == np.array(toks)).sum()), "/", len(toks))
    assert err < 0.01 * max(1.0, float(lg.abs().max()))                     # bf16 mode: 0.05
    for k, a, b, margin in _margin_report(lg.numpy(), sampled, toks):
        assert margin < 0.02, (k, a, b, margin)
    res = engine_small_fp16.run_batch([(1, ids, 0, sp)], timeout_s=60, want_latents=True)
    r, got, wav, glat = res[1]
    assert np.isfinite(wav).all() and r.n_tokens >= 1
    if list(got) == toks:                                   # same tokens (no near-tie flipped): the captured latents agree too
        assert np.abs(glat - lats.numpy()).max() < 0.02


def test_fp16_operand_mode_full_size_vs_bf16(engine_full_fp16, engine_full_bf16, dims_full, state_full, speakers_full):
    """full geometry, 24 greedy tokens teacher-forced: the fp16-operand mode's logits are several times closer to the fp32
    oracle than the bf16 mode's, through the same kernels (split-K GEMMs, paged attention, graph replay in run_batch)."""
    orc = _orc(dims_full, state_full)
    g = dims_full.gpt
    ids = text_ids(dims_full, 14, 9)
    osp = O.SamplingParams(temperature=0.0, repetition_penalty=5.0, max_tokens=24, stop_token=g.stop_audio_token)
    toks, lats, lg = orc.generate(speakers_full[1][0], ids, osp, return_logits=True)
    sp = Sampling(temperature=0.0, repetition_penalty=5.0, max_tokens=24, stop_token=g.stop_audio_token)
    l16, lat16, s16 = engine_full_fp16.gpt_teacher_forced(ids, 1, toks, sp)
    lbf, latbf, sbf = engine_full_bf16.gpt_teacher_forced(ids, 1, toks, sp)
    e16, ebf = float(np.abs(l16 - lg.numpy()).max()), float(np.abs(lbf - lg.numpy()).max())
    a16, abf = int((s16 == np.array(toks)).sum()), int((sbf == np.array(toks)).sum())
    print(f"full geometry logits max err: fp16 operands {e16:.4f}, bf16 operands {ebf:.4f} (|logit| max {float(lg.abs().max()):.2f}); "
          f"greedy agreement {a16} / {abf} of {len(toks)}; latents {float(np.abs(lat16 - lats.numpy()).max()):.4f} / {float(np.abs(latbf - lats.numpy()).max()):.4f}")
    assert e16 < 0.5 * ebf and e16 < 0.01 * max(1.0, float(lg.abs().max()))
    assert a16 >= abf
    for k, a, b, margin in _margin_report(lg.numpy(), s16, toks):
        assert margin < 0.03, (k, a, b, margin)
    res = engine_full_fp16.run_batch([(i, ids, 1, sp) for i in range(3)], timeout_s=120)
    for i in range(3):
        assert res[i][0].n_tokens == 24 and np.isfinite(res[i][2]).all()
        np.testing.assert_array_equal(res[i][1], res[0][1])
