"""Generates the golden vectors under tests/golden/ by running the REFERENCE's own modules
(imported unmodified from /root/reference/src via oracle/ref_import.py) on the synthetic weights
of auralis_b200.weights.synth_state.  Container-only (needs /root/reference); the .npz files it
writes are committed and are what the GPU box checks the oracle and the CUDA path against.

    python tests/golden/make_golden.py

Pin: torch version, thread count and seeds are recorded inside every file.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from auralis_b200.config import XTTSDims            # noqa: E402
from auralis_b200.weights import synth_state        # noqa: E402
from oracle import ref_import                       # noqa: E402
from oracle import xtts_oracle as O                 # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SEED = 1234


def ref_vocoder(ns, dims, cs):
    v = dims.voc
    gen = ns.HifiganGenerator(v.in_dim, 1, "1", [list(v.rb_dilations)] * len(v.rb_kernels), list(v.rb_kernels),
                              list(v.up_kernels), v.init_ch, list(v.up_rates), inference_padding=0,
                              cond_channels=v.d_vector, conv_pre_weight_norm=False, conv_post_weight_norm=False,
                              conv_post_bias=False, cond_in_each_up_layer=True)
    pre = "hifigan_decoder.waveform_decoder."
    gen.load_state_dict({k[len(pre):]: t for k, t in cs.items() if k.startswith(pre)}, strict=True)
    gen.eval()
    hd = ns.HifiDecoder.__new__(ns.HifiDecoder)
    torch.nn.Module.__init__(hd)
    hd.input_sample_rate, hd.output_sample_rate = v.input_sample_rate, v.output_sample_rate
    hd.output_hop_length, hd.ar_mel_length_compression = v.output_hop_length, v.code_stride
    hd.waveform_decoder = gen
    return hd


def main():
    torch.set_num_threads(8)
    ns = ref_import.load()
    meta = dict(torch=torch.__version__, threads=8, seed=SEED, reference="astramind-ai/Auralis@e54a8de")
    for name, dims in (("small", XTTSDims.small()), ("full", XTTSDims.full())):
        gs, cs = synth_state(dims, SEED)
        g = torch.Generator().manual_seed(99)
        # ---- vocoder: HifiDecoder.forward (hifigan_decoder.py:776-802)
        T = 9 if name == "small" else 5
        lat = torch.randn(T, dims.voc.in_dim, generator=g)
        gv = torch.nn.functional.normalize(torch.randn(dims.voc.d_vector, generator=g), dim=0)
        hd = ref_vocoder(ns, dims, cs)
        with torch.no_grad():
            wav = ns.HifiDecoder.forward(hd, lat[None], g=gv.reshape(1, -1, 1)).reshape(-1)
        np.savez_compressed(os.path.join(OUT, f"vocoder_{name}.npz"), latents=lat.numpy(), g=gv.numpy(),
                            wav=wav.numpy(), meta=str(meta))
        # ---- conditioning: wav_to_mel_cloning -> ConditioningEncoder -> PerceiverResampler
        wav22 = O.synthetic_reference_wav(0.6, 22050, 130.0, 7)
        mel = ns.wav_to_mel_cloning(wav22[None], mel_norms=cs["mel_stats"], n_fft=2048, hop_length=256, win_length=1024,
                                    power=2, normalized=False, sample_rate=22050, f_min=0, f_max=8000, n_mels=80)
        ce = ns.ConditioningEncoder(dims.cond.n_mels, dims.gpt.hidden, attn_blocks=dims.cond.cond_blocks,
                                    num_attn_heads=dims.gpt.heads)
        ce.load_state_dict({k[len("conditioning_encoder."):]: t for k, t in cs.items() if k.startswith("conditioning_encoder.")})
        pr = ns.PerceiverResampler(dim=dims.gpt.hidden, depth=dims.cond.perceiver_depth, dim_context=dims.gpt.hidden,
                                   num_latents=dims.gpt.n_cond_latents, dim_head=dims.cond.perceiver_dim_head,
                                   heads=dims.cond.perceiver_heads, ff_mult=dims.cond.perceiver_ff_mult, use_flash_attn=False)
        pr.load_state_dict({k[len("conditioning_perceiver."):]: t for k, t in cs.items() if k.startswith("conditioning_perceiver.")})
        with torch.no_grad():
            h = ce(mel)
            latp = pr(h.permute(0, 2, 1))
        # ---- speaker encoder (hifigan_decoder.py:602-646)
        se = ns.ResNetSpeakerEncoder(input_dim=dims.cond.spk_mels, proj_dim=dims.cond.spk_proj, layers=list(dims.cond.spk_layers),
                                     num_filters=list(dims.cond.spk_filters), log_input=True, use_torch_spec=True,
                                     audio_config={"fft_size": 512, "win_length": 400, "hop_length": 160,
                                                   "sample_rate": 16000, "preemphasis": 0.97, "num_mels": 64})
        pre = "hifigan_decoder.speaker_encoder."
        se.load_state_dict({k[len(pre):]: t for k, t in cs.items() if k.startswith(pre)})
        se.eval()
        wav16 = O.synthetic_reference_wav(0.6, 16000, 130.0, 7)
        with torch.no_grad():
            dvec = se.forward(wav16[None].clone(), l2_norm=True)[0]
        keep_h = h[0].numpy() if name == "small" else h[0, :, :8].numpy()      # keep the full-size file small
        np.savez_compressed(os.path.join(OUT, f"conditioning_{name}.npz"), wav22=wav22.numpy(), wav16=wav16.numpy(),
                            mel=mel[0].numpy(), cond_enc=keep_h, perceiver=latp[0].numpy(), dvector=dvec.numpy(),
                            meta=str(meta))
        print(name, "wav", tuple(wav.shape), "mel", tuple(mel.shape), "perceiver", tuple(latp.shape))


if __name__ == "__main__":
    main()
