"""TEST INFRASTRUCTURE: a synthetic Coqui-format XTTSv2 checkpoint (the INPUT of the reference's converter) and a loader
for the reference converter module.  Coqui key names follow what `convert_model_weights` consumes
(`/root/reference/src/auralis/models/xttsv2/utils/checkpoint_converter.py:225-284`) and the reference's own fixture
(`/root/reference/tests/integration/test_checkpoint_converter.py:18-52`)."""
import importlib
import os

import torch

from auralis_b200.config import GPTDims, XTTSDims
from auralis_b200.weights import synth_state
from oracle import ref_import


def load_reference_converter():
    if not ref_import.available():
        raise RuntimeError("reference tree not mounted")
    ref_import.load()
    base = os.path.join(ref_import.REF_SRC, "auralis", "models", "xttsv2")
    ref_import._stub("auralis.models.xttsv2.utils", os.path.join(base, "utils"))
    return importlib.import_module("auralis.models.xttsv2.utils.checkpoint_converter")


def dims_for(layers: int) -> XTTSDims:
    d = XTTSDims.full()
    d.gpt = GPTDims(layers=layers, n_text_tokens=6153)          # a finetune-sized text vocabulary, as in the reference's test
    return d


def to_coqui_name(which: str, name: str) -> str:
    """Our (= the converter's OUTPUT) tensor name -> the Coqui checkpoint name it came from."""
    if which == "gpt":
        if name == "gpt.wte.weight":
            return "gpt.mel_embedding.weight"
        if name == "gpt.wpe.emb.weight":
            return "gpt.mel_pos_embedding.emb.weight"
        if name.startswith("mel_head.") or name.startswith("final_norm."):
            return "gpt." + name
        return "gpt." + name                                     # gpt.h.N.* / gpt.ln_f.* -> gpt.gpt.h.N.* / gpt.gpt.ln_f.*
    if name.startswith("hifigan_decoder.") or name == "mel_stats":
        return name
    return "gpt." + name                                         # text_embedding, conditioning_*, text_head, ...


def coqui_checkpoint(layers: int = 2, seed: int = 5, fill: str = "random") -> dict:
    """{'model': Coqui-named state, 'config': ..., 'model_args': ...}.  fill="meta": zero-stride placeholders (names and
    shapes only, no 1.5 GB of random numbers)."""
    dims = dims_for(layers)
    if fill == "meta":
        gs, cs = synth_state(XTTSDims.small(), seed)              # names only ...
        full_g, full_c = _shapes(dims)
        gs = {k: torch.zeros(()).expand(full_g[k]) for k in gs if k in full_g}
        cs = {k: torch.zeros(()).expand(full_c[k]) for k in full_c}
    else:
        gs, cs = synth_state(dims, seed)
    model = {}
    for k, v in gs.items():
        model[to_coqui_name("gpt", k)] = v
    for k, v in cs.items():
        if k.startswith("final_norm."):
            continue                                             # one tensor in Coqui (gpt.final_norm.*), copied to both files
        model[to_coqui_name("core", k)] = v
    model["gpt.mel_embedding.weight"] = gs["gpt.wte.weight"]
    # training-only tensors the converter must drop (checkpoint_converter.py:248-252)
    model["dvae.codebook.embed"] = torch.zeros(4, 4)
    model["torch_mel_spectrogram_dvae.mel_stft.spectrogram.window"] = torch.zeros(8)
    config = {"gpt_max_text_tokens": dims.gpt.max_text_tokens, "output_hop_length": 256, "input_sample_rate": 22050,
              "output_sample_rate": 24000, "gpt_code_stride_len": 1024, "d_vector_dim": 512, "speaker_dim": 512,
              "languages": ["en", "es", "fr", "de", "it", "pt", "pl", "tr", "ru", "nl", "cs", "ar", "zh-cn", "hu", "ko", "ja"],
              "audio_config": {"sample_rate": 22050, "output_sample_rate": 24000}}
    return {"model": model, "config": config, "model_args": {"use_masking_gt_prompt_approach": True, "use_perceiver_resampler": True}}


def _shapes(dims: XTTSDims):
    """Tensor shapes of the full geometry without materialising it: build on the meta device."""
    with torch.device("meta"):
        gs, cs = synth_state(dims, 0)
    return {k: tuple(v.shape) for k, v in gs.items()}, {k: tuple(v.shape) for k, v in cs.items()}
