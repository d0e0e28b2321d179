"""Generates tests/golden/converter_layout.json by running the REFERENCE's own checkpoint converter
(/root/reference/src/auralis/models/xttsv2/utils/checkpoint_converter.py, imported unmodified) on a synthetic
Coqui-format XTTSv2 checkpoint (full tensor shapes, 2 GPT layers — the shape of the reference's own converter test,
tests/integration/test_checkpoint_converter.py:18-52).  Records what the converter wrote: both config.json files and
every tensor name with its shape.  tests/test_converter_layout.py replays them against auralis_b200's loader.

    python tests/golden/make_converter_golden.py          (container only: needs /root/reference)
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from converter_fixture import coqui_checkpoint, load_reference_converter  # noqa: E402


def main():
    conv = load_reference_converter()
    ckpt = coqui_checkpoint(layers=2, seed=5, fill="meta")
    with tempfile.TemporaryDirectory() as out:
        gpt_w, xtts_w = conv.convert_model_weights(ckpt["model"])
        paths = conv.save_configs(out, ckpt)
        core_cfg = json.load(open(paths[2]))
        gpt_cfg = json.load(open(paths[0]))
    rec = {"source": "reference checkpoint_converter.py (convert_model_weights + save_configs), imported unmodified",
           "core_config": core_cfg, "gpt_config": gpt_cfg,
           "gpt_tensors": {k: list(v.shape) for k, v in sorted(gpt_w.items())},
           "core_tensors": {k: list(v.shape) for k, v in sorted(xtts_w.items())}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "converter_layout.json")
    with open(path, "w") as f:
        json.dump(rec, f, indent=0)
    print(f"{len(rec['gpt_tensors'])} gpt + {len(rec['core_tensors'])} core tensors -> {path}")


if __name__ == "__main__":
    main()
