"""Weight containers for the XTTSv2 hot path.

* ``synth_state(dims, seed)`` builds deterministic random-init weights in the
  reference's checkpoint naming (the contract written by
  `/root/reference/src/auralis/models/xttsv2/utils/checkpoint_converter.py:230-272`
  for the GPT file and by ``XTTSv2Engine.state_dict()``
  `/root/reference/src/auralis/models/xttsv2/XTTSv2.py:91-140,300-301` for the core file).
  There is no network in this build, so no real checkpoint exists; every parity
  test and the bench run on these.
* ``load_model_dir`` / ``save_model_dir`` read and write the two safetensors files
  + config.json layout ``TTS.from_pretrained`` expects
  (`/root/reference/src/auralis/core/tts.py:53-89`, `XTTSv2.py:235-310`).

Tensors are fp32 torch CPU tensors: containers only — the native library copies
them to HBM and never sees a torch type.
"""
from __future__ import annotations

import hashlib
import json
import math
import os
from typing import Dict, Tuple

import numpy as np
import torch

from .config import XTTSDims

State = Dict[str, torch.Tensor]


def _gen(seed: int, name: str) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)
    return g


class _Maker:
    def __init__(self, seed: int):
        self.seed = seed
        self.out: State = {}

    def normal(self, name, shape, std=0.02, mean=0.0):
        t = torch.empty(*shape, dtype=torch.float32)
        t.normal_(mean, std, generator=_gen(self.seed, name))
        self.out[name] = t
        return t

    def uniform(self, name, shape, lo, hi):
        t = torch.empty(*shape, dtype=torch.float32)
        t.uniform_(lo, hi, generator=_gen(self.seed, name))
        self.out[name] = t
        return t

    def const(self, name, t):
        self.out[name] = t.to(torch.float32).contiguous()
        return self.out[name]

    def ln(self, prefix, n):
        self.normal(prefix + ".weight", (n,), std=0.1, mean=1.0)
        self.normal(prefix + ".bias", (n,), std=0.05)

    def conv(self, prefix, shape, bias=True, gain=1.0, fan_in=None):
        """Conv/linear weight ~ N(0, gain/sqrt(fan_in)); small non-zero bias."""
        if fan_in is None:
            fan_in = int(np.prod(shape[1:]))
        self.normal(prefix + ".weight", shape, std=gain / math.sqrt(fan_in))
        if bias:
            self.normal(prefix + ".bias", (shape[0],), std=0.02)

    def wn_conv(self, prefix, shape, gain=1.0, fan_in=None, bias_n=None):
        """weight-normed conv: parametrizations.weight.original0 (g, over dim 0) / original1 (v)
        (`hifigan_decoder.py:44-73,189-202`; torch parametrizations naming)."""
        if fan_in is None:
            fan_in = int(np.prod(shape[1:]))
        v = self.normal(prefix + ".parametrizations.weight.original1", shape, std=1.0 / math.sqrt(fan_in))
        # g = gain * ||v|| * (1 +- 10%) so that the fold actually rescales
        nrm = v.reshape(shape[0], -1).norm(dim=1).reshape(shape[0], *([1] * (len(shape) - 1)))
        j = torch.empty(shape[0], *([1] * (len(shape) - 1))).uniform_(0.9, 1.1, generator=_gen(self.seed, prefix + ".g"))
        self.out[prefix + ".parametrizations.weight.original0"] = (gain * nrm * j).contiguous()
        self.normal(prefix + ".bias", (bias_n if bias_n is not None else shape[0],), std=0.02)

    def bn(self, prefix, n):
        self.normal(prefix + ".weight", (n,), std=0.1, mean=1.0)
        self.normal(prefix + ".bias", (n,), std=0.05)
        self.normal(prefix + ".running_mean", (n,), std=0.1)
        self.uniform(prefix + ".running_var", (n,), 0.5, 1.5)
        self.const(prefix + ".num_batches_tracked", torch.zeros(()))


def synth_gpt_state(dims: XTTSDims, seed: int = 1234) -> State:
    g = dims.gpt
    m = _Maker(seed)
    H, F = g.hidden, g.ff
    m.normal("gpt.wte.weight", (g.n_audio_tokens, H), std=0.02)
    m.normal("gpt.wpe.emb.weight", (g.n_wpe, H), std=0.02)
    for i in range(g.layers):
        p = f"gpt.h.{i}."
        m.ln(p + "ln_1", H)
        m.ln(p + "ln_2", H)
        # HF Conv1D layout [in, out] (checkpoint_converter.py:236-262; transposed at load vllm_mm_gpt.py:723-725)
        m.normal(p + "attn.c_attn.weight", (H, 3 * H), std=0.02)
        m.normal(p + "attn.c_attn.bias", (3 * H,), std=0.01)
        m.normal(p + "attn.c_proj.weight", (H, H), std=0.02)
        m.normal(p + "attn.c_proj.bias", (H,), std=0.01)
        m.normal(p + "mlp.c_fc.weight", (H, F), std=0.02)
        m.normal(p + "mlp.c_fc.bias", (F,), std=0.01)
        m.normal(p + "mlp.c_proj.weight", (F, H), std=0.02)
        m.normal(p + "mlp.c_proj.bias", (H,), std=0.01)
    m.ln("gpt.ln_f", H)
    m.normal("mel_head.weight", (g.n_audio_tokens, H), std=0.05)
    m.normal("mel_head.bias", (g.n_audio_tokens,), std=0.02)
    m.ln("final_norm", H)
    return m.out


def synth_core_state(dims: XTTSDims, seed: int = 1234, gpt_state: State | None = None) -> State:
    g, v, c = dims.gpt, dims.voc, dims.cond
    m = _Maker(seed + 1)
    H = g.hidden
    m.uniform("mel_stats", (c.n_mels,), 0.8, 1.25)
    # --- ConditioningEncoder (latent_encoder.py:209-253)
    m.conv("conditioning_encoder.init", (H, c.n_mels, 1))
    for i in range(c.cond_blocks):
        p = f"conditioning_encoder.attn.{i}."
        m.ln(p + "norm", H)
        m.conv(p + "qkv", (3 * H, H, 1))
        # proj_out is zero-init in the reference (latent_encoder.py:178); re-drawn so the path is exercised
        m.conv(p + "proj_out", (H, H, 1))
    m.normal("text_embedding.weight", (g.n_text_tokens, H), std=1.0)
    m.normal("text_pos_embedding.emb.weight", (g.n_text_pos, H), std=0.02)
    # --- PerceiverResampler (perceiver_encoder.py:363-442)
    inner = c.perceiver_heads * c.perceiver_dim_head
    ffi = int(H * c.perceiver_ff_mult * 2 / 3)
    m.normal("conditioning_perceiver.latents", (g.n_cond_latents, H), std=0.5)
    for l in range(c.perceiver_depth):
        p = f"conditioning_perceiver.layers.{l}."
        m.conv(p + "0.to_q", (inner, H), bias=False)
        m.conv(p + "0.to_kv", (2 * inner, H), bias=False)
        m.conv(p + "0.to_out", (H, inner), bias=False)
        m.conv(p + "1.0", (2 * ffi, H))
        m.conv(p + "1.2", (H, ffi))
    m.normal("conditioning_perceiver.norm.gamma", (H,), std=0.1, mean=1.0)
    # --- HiFi-GAN generator (hifigan_decoder.py:145-226, HifiDecoder :692-773)
    w = "hifigan_decoder.waveform_decoder."
    m.conv(w + "conv_pre", (v.init_ch, v.in_dim, 7))
    m.conv(w + "cond_layer", (v.init_ch, v.d_vector, 1))
    ch = v.init_ch
    nk = len(v.rb_kernels)
    for i, (u, k) in enumerate(zip(v.up_rates, v.up_kernels)):
        cin, cout = v.init_ch // (2 ** i), v.init_ch // (2 ** (i + 1))
        # ConvTranspose1d weight [Cin, Cout, K]; weight_norm dim 0 -> g [Cin,1,1]
        # each output sample sums Cin * K/u taps
        m.wn_conv(w + f"ups.{i}", (cin, cout, k), fan_in=cin * k // u, bias_n=cout)
        m.conv(w + f"conds.{i}", (cout, v.d_vector, 1))
        for j, kk in enumerate(v.rb_kernels):
            for t in range(len(v.rb_dilations)):
                # small residual-branch gain keeps the 9-deep residual chain O(1)
                m.wn_conv(w + f"resblocks.{i * nk + j}.convs1.{t}", (cout, cout, kk), gain=0.7)
                m.wn_conv(w + f"resblocks.{i * nk + j}.convs2.{t}", (cout, cout, kk), gain=0.7)
        ch = cout
    m.conv(w + "conv_post", (1, ch, 7), bias=False, gain=0.5)
    # --- ResNet speaker encoder (hifigan_decoder.py:485-646)
    s = "hifigan_decoder.speaker_encoder."
    nf = c.spk_filters
    m.conv(s + "conv1", (nf[0], 1, 3, 3), gain=math.sqrt(2.0))
    m.bn(s + "bn1", nf[0])
    inpl = nf[0]
    for li, (planes, nb) in enumerate(zip(nf, c.spk_layers)):
        stride = 1 if li == 0 else 2
        for b in range(nb):
            p = s + f"layer{li + 1}.{b}."
            m.conv(p + "conv1", (planes, inpl if b == 0 else planes, 3, 3), bias=False, gain=math.sqrt(2.0))
            m.bn(p + "bn1", planes)
            m.conv(p + "conv2", (planes, planes, 3, 3), bias=False)
            m.bn(p + "bn2", planes)
            m.conv(p + "se.fc.0", (planes // 8, planes))
            m.conv(p + "se.fc.2", (planes, planes // 8))
            if b == 0 and (stride != 1 or inpl != planes):
                m.conv(p + "downsample.0", (planes, inpl, 1, 1), bias=False)
                m.bn(p + "downsample.1", planes)
        inpl = planes
    om = nf[3] * (c.spk_mels // 8)
    m.conv(s + "attention.0", (128, om, 1))
    m.bn(s + "attention.2", 128)
    m.conv(s + "attention.3", (om, 128, 1))
    m.conv(s + "fc", (c.spk_proj, om * 2))
    # persisted buffers of torch_spec (PreEmphasis + torchaudio MelSpectrogram, hifigan_decoder.py:556-569)
    m.const(s + "torch_spec.0.filter", torch.tensor([-0.97, 1.0]).view(1, 1, 2))
    m.const(s + "torch_spec.1.spectrogram.window", torch.hamming_window(400))
    m.const(s + "torch_spec.1.mel_scale.fb", mel_filterbank(257, 0.0, 8000.0, c.spk_mels, 16000, norm=None))
    # engine-level final_norm is the same tensor as the GPT's (checkpoint_converter.py:270-273)
    if gpt_state is not None:
        m.const("final_norm.weight", gpt_state["final_norm.weight"])
        m.const("final_norm.bias", gpt_state["final_norm.bias"])
    else:
        m.ln("final_norm", H)
    m.conv("text_head", (g.n_text_tokens, H))   # unused by inference (XTTSv2.py:139-140)
    return m.out


def synth_state(dims: XTTSDims, seed: int = 1234) -> Tuple[State, State]:
    gs = synth_gpt_state(dims, seed)
    cs = synth_core_state(dims, seed, gs)
    return gs, cs


# ----------------------------------------------------------------------------------------------
# mel filterbank (restates torchaudio.functional.melscale_fbanks, mel_scale="htk";
# used by MelSpectrogram at common/utilities.py:53-64 with norm="slaney" and by the speaker
# encoder's torch_spec at hifigan_decoder.py:559-568 with norm=None)
# ----------------------------------------------------------------------------------------------
def mel_filterbank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int,
                   norm: str | None) -> torch.Tensor:
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    if norm == "slaney":
        enorm = 2.0 / (f_pts[2: n_mels + 2] - f_pts[:n_mels])
        fb = fb * enorm.unsqueeze(0)
    return fb.contiguous()


# ----------------------------------------------------------------------------------------------
# model directory I/O
# ----------------------------------------------------------------------------------------------
def save_model_dir(path: str, dims: XTTSDims, gpt_state: State, core_state: State) -> None:
    from safetensors.torch import save_file
    os.makedirs(os.path.join(path, "gpt"), exist_ok=True)
    cfg = {"model_type": "xtts", "b200_dims": dims.to_json(),
           "gpt_config": {"hidden_size": dims.gpt.hidden, "num_hidden_layers": dims.gpt.layers,
                          "num_attention_heads": dims.gpt.heads, "n_inner": dims.gpt.ff}}
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f, indent=1)
    save_file({k: v.contiguous() for k, v in core_state.items()}, os.path.join(path, "xtts-v2.safetensors"))
    save_file({k: v.contiguous() for k, v in gpt_state.items()}, os.path.join(path, "gpt", "gpt2_model.safetensors"))
    with open(os.path.join(path, "gpt", "config.json"), "w") as f:
        json.dump(cfg["gpt_config"], f)


def resolve_model_file(path_or_repo: str, filename: str, required: bool = True) -> str | None:
    """A file of a model: `<dir>/<filename>` when the directory exists, else the Hugging Face Hub repo of that name —
    the rule of the reference (`XTTSv2.py:262-298`, `core/tts.py:72-84`: `TTS().from_pretrained("AstraMindAI/xttsv2",
    gpt_model="AstraMindAI/xtts2-gpt")`)."""
    if os.path.isdir(path_or_repo):
        p = os.path.join(path_or_repo, filename)
        if os.path.exists(p) or required:
            return p
        return None
    try:
        from huggingface_hub import hf_hub_download
        return hf_hub_download(repo_id=path_or_repo, filename=filename)
    except Exception as e:      # noqa: BLE001 — offline, unknown repo, missing file
        if required:
            raise ValueError(f"Could not load {filename} from {path_or_repo} neither locally or online: {e}") from e
        return None


def load_model_dir(path: str, gpt_model: str | None = None) -> Tuple[XTTSDims, State, State]:
    """Reads a model directory in either layout:
      * the reference converter's (`utils/checkpoint_converter.py:286-334`): `path` = `.../core_xttsv2` holding
        `config.json` + `xtts-v2.safetensors`, `gpt_model` = `.../gpt` holding `config.json` + `gpt2_model.safetensors`
        (what `TTS().from_pretrained("AstraMindAI/xttsv2", gpt_model="AstraMindAI/xtts2-gpt")` resolves to);
      * this package's own (`save_model_dir`): one directory with a `gpt/` sub-directory and the geometry under "b200_dims".
    Tensor names are the converter's in both (`checkpoint_converter.py:225-284`); shapes are checked against the geometry."""
    from safetensors.torch import load_file
    with open(resolve_model_file(path, "config.json")) as f:
        cfg = json.load(f)
    if gpt_model is not None and gpt_model.endswith(".safetensors"):
        gfile, gcfg_path = gpt_model, os.path.join(os.path.dirname(gpt_model), "config.json")
    else:
        gsrc = gpt_model if gpt_model is not None else os.path.join(path, "gpt")
        gfile = resolve_model_file(gsrc, "gpt2_model.safetensors")
        gcfg_path = resolve_model_file(gsrc, "config.json", required=False)
    if "b200_dims" in cfg:
        dims = XTTSDims.from_json(cfg["b200_dims"])
    else:
        gcfg = None
        if gcfg_path and os.path.exists(gcfg_path):
            with open(gcfg_path) as f:
                gcfg = json.load(f)
        dims = XTTSDims.from_reference_configs(cfg, gcfg)
    core = {k: v.float() for k, v in load_file(resolve_model_file(path, "xtts-v2.safetensors")).items()}
    gpt = {k: v.float() for k, v in load_file(gfile).items()}
    check_state_shapes(dims, gpt, core)
    return dims, gpt, core


def check_state_shapes(dims: XTTSDims, gpt: State, core: State) -> None:
    """Fails with the offending tensor's name when a checkpoint does not match the geometry of its config
    (the native loader would otherwise report a bare shape mismatch much later)."""
    g = dims.gpt
    want = {
        ("gpt", "gpt.wte.weight"): (g.n_audio_tokens, g.hidden),
        ("gpt", "mel_head.weight"): (g.n_audio_tokens, g.hidden),
        ("gpt", "gpt.h.0.attn.c_attn.weight"): (g.hidden, 3 * g.hidden),
        ("gpt", f"gpt.h.{g.layers - 1}.mlp.c_fc.weight"): (g.hidden, g.ff),
        ("gpt", "final_norm.weight"): (g.hidden,),
        ("core", "text_embedding.weight"): (g.n_text_tokens, g.hidden),
    }
    for (which, name), shape in want.items():
        st = gpt if which == "gpt" else core
        if name not in st:
            raise KeyError(f"{which} checkpoint has no tensor {name!r}")
        if tuple(st[name].shape) != tuple(shape):
            raise ValueError(f"{name}: checkpoint shape {tuple(st[name].shape)} != {tuple(shape)} implied by config.json")
    if f"gpt.h.{g.layers}.ln_1.weight" in gpt:
        raise ValueError(f"checkpoint has more than num_hidden_layers = {g.layers} GPT layers")
    if gpt["gpt.wpe.emb.weight"].shape[0] < g.max_audio_tokens + 1:
        raise ValueError("gpt.wpe.emb.weight has fewer rows than max_audio_tokens + 1")
    if core["text_pos_embedding.emb.weight"].shape[0] < g.max_text_tokens + 2:
        raise ValueError("text_pos_embedding.emb.weight has fewer rows than max_text_tokens + 2")
