"""CPU: the model-directory format on the input side of the hot path.  The reference's converter
(`utils/checkpoint_converter.py`) turns a Coqui XTTSv2 checkpoint into `core_xttsv2/` + `gpt/` (two safetensors files
and two config.json); its own test pins the key names and config integers
(`/root/reference/tests/integration/test_checkpoint_converter.py:140-326`).  Here: what that converter writes is what
`auralis_b200.weights.load_model_dir` reads — by golden record (tests/golden/converter_layout.json, produced by the
reference converter itself) and, when /root/reference is mounted, by running the converter end to end."""
import json
import os

import pytest
import torch

from auralis_b200.config import XTTSDims
from auralis_b200.weights import check_state_shapes, load_model_dir
from converter_fixture import _shapes, coqui_checkpoint, dims_for, load_reference_converter
from oracle import ref_import

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "converter_layout.json")


def test_reference_converter_output_is_our_input_format():
    rec = json.load(open(GOLD))
    dims = XTTSDims.from_reference_configs(rec["core_config"], rec["gpt_config"])
    assert dims == dims_for(2)
    # the integers the reference's own test pins (test_checkpoint_converter.py:146-154,292-306)
    g = rec["gpt_config"]
    assert (g["vocab_size"], g["hidden_size"], g["num_hidden_layers"], g["num_attention_heads"], g["n_inner"]) == (6153, 1024, 2, 16, 4096)
    assert (g["num_audio_tokens"], g["max_audio_tokens"], g["start_audio_token"], g["stop_audio_token"]) == (1026, 605, 1024, 1025)
    assert (dims.gpt.n_text_tokens, dims.gpt.max_audio_tokens, dims.gpt.start_audio_token, dims.gpt.stop_audio_token) == (6153, 605, 1024, 1025)
    assert rec["core_config"]["model_type"] == "xtts" and rec["core_config"]["gpt"]["model_type"] == "xtts_gpt"
    # tensor names and shapes: exactly the state the native loader consumes for that geometry
    want_g, want_c = _shapes(dims)
    assert {k: tuple(v) for k, v in rec["gpt_tensors"].items()} == want_g
    assert {k: tuple(v) for k, v in rec["core_tensors"].items()} == want_c
    # only the core config (gpt_config embedded) is enough, as when `gpt_model` points at a bare safetensors file
    assert XTTSDims.from_reference_configs(rec["core_config"], None) == dims


def test_config_and_shape_errors_are_named():
    rec = json.load(open(GOLD))
    bad = dict(rec["gpt_config"], num_attention_heads=1)
    with pytest.raises(ValueError, match="64-wide heads"):
        XTTSDims.from_reference_configs(rec["core_config"], bad)
    with pytest.raises(ValueError, match="gelu_new"):
        XTTSDims.from_reference_configs(rec["core_config"], dict(rec["gpt_config"], activation_function="relu"))
    dims = XTTSDims.from_reference_configs(rec["core_config"], rec["gpt_config"])
    gs = {k: torch.zeros(()).expand(v) for k, v in rec["gpt_tensors"].items()}
    cs = {k: torch.zeros(()).expand(v) for k, v in rec["core_tensors"].items()}
    check_state_shapes(dims, gs, cs)
    with pytest.raises(ValueError, match="text_embedding.weight"):
        check_state_shapes(dims, gs, dict(cs, **{"text_embedding.weight": torch.zeros(6681, 1024)}))
    with pytest.raises(KeyError, match="mel_head.weight"):
        check_state_shapes(dims, {k: v for k, v in gs.items() if k != "mel_head.weight"}, cs)
    with pytest.raises(ValueError, match="more than num_hidden_layers"):
        check_state_shapes(dims, dict(gs, **{"gpt.h.2.ln_1.weight": torch.zeros(1024)}), cs)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
def test_live_reference_converter_to_loader_round_trip(tmp_path):
    """Coqui-format checkpoint -> the reference's converter (unmodified) -> load_model_dir: same geometry, every tensor
    bit-identical, training-only tensors dropped, final_norm present in both files."""
    conv = load_reference_converter()
    torch.manual_seed(0)
    ckpt = coqui_checkpoint(layers=1, seed=9, fill="random")
    gpt_w, xtts_w = conv.convert_model_weights(ckpt["model"])
    conv.save_model_weights({k: v.contiguous() for k, v in gpt_w.items()}, {k: v.contiguous() for k, v in xtts_w.items()}, str(tmp_path))
    conv.save_configs(str(tmp_path), ckpt)
    dims, gs, cs = load_model_dir(str(tmp_path / "core_xttsv2"), gpt_model=str(tmp_path / "gpt"))
    assert dims == dims_for(1)
    src = ckpt["model"]
    assert torch.equal(gs["gpt.wte.weight"], src["gpt.mel_embedding.weight"])
    assert torch.equal(gs["gpt.wpe.emb.weight"], src["gpt.mel_pos_embedding.emb.weight"])
    assert torch.equal(gs["gpt.h.0.attn.c_attn.weight"], src["gpt.gpt.h.0.attn.c_attn.weight"])
    assert torch.equal(gs["gpt.ln_f.bias"], src["gpt.gpt.ln_f.bias"]) and torch.equal(gs["mel_head.weight"], src["gpt.mel_head.weight"])
    assert torch.equal(gs["final_norm.weight"], src["gpt.final_norm.weight"]) and torch.equal(cs["final_norm.weight"], src["gpt.final_norm.weight"])
    assert torch.equal(cs["text_embedding.weight"], src["gpt.text_embedding.weight"])
    assert torch.equal(cs["hifigan_decoder.waveform_decoder.conv_pre.weight"], src["hifigan_decoder.waveform_decoder.conv_pre.weight"])
    assert not any("dvae" in k for k in list(gs) + list(cs))
    want_g, want_c = _shapes(dims)
    assert {k: tuple(v.shape) for k, v in gs.items()} == want_g and {k: tuple(v.shape) for k, v in cs.items()} == want_c


def test_hub_repo_ids_resolve_like_the_reference(tmp_path, monkeypatch):
    """`from_pretrained("org/repo", gpt_model="org/gpt-repo")`: names that are not local directories are fetched file by file
    through huggingface_hub (XTTSv2.py:262-298) — here a fake hub that serves files from two local folders."""
    import huggingface_hub
    from auralis_b200.weights import resolve_model_file, save_model_dir, synth_state
    dims = XTTSDims.small()
    gs, cs = synth_state(dims, 3)
    save_model_dir(str(tmp_path / "m"), dims, gs, cs)
    repos = {"Org/xtts-core": tmp_path / "m", "Org/xtts-gpt": tmp_path / "m" / "gpt"}
    asked = []

    def fake_download(repo_id, filename, **kw):
        asked.append((repo_id, filename))
        p = repos[repo_id] / filename
        if not p.exists():
            raise FileNotFoundError(filename)
        return str(p)
    monkeypatch.setattr(huggingface_hub, "hf_hub_download", fake_download)
    d2, g2, c2 = load_model_dir("Org/xtts-core", gpt_model="Org/xtts-gpt")
    assert d2 == dims and set(g2) == set(gs) and set(c2) == set(cs)
    assert ("Org/xtts-core", "xtts-v2.safetensors") in asked and ("Org/xtts-gpt", "gpt2_model.safetensors") in asked
    assert resolve_model_file("Org/xtts-gpt", "tokenizer.json", required=False) is None          # optional file: no error
    with pytest.raises(ValueError, match="neither locally or online"):
        resolve_model_file("Org/xtts-gpt", "nope.bin")
    with pytest.raises(ValueError, match="neither locally or online"):
        from auralis_b200 import TTS
        monkeypatch.setattr(huggingface_hub, "hf_hub_download", lambda **kw: (_ for _ in ()).throw(OSError("offline")))
        TTS().from_pretrained("Org/unknown")
