#!/usr/bin/env python
"""Run one of the REFERENCE's own scripts (train.py / eval.py / predict_video.py) unmodified, with the
drop-in `model` package of this repo shadowing the reference's `model` (SURVEY section 8b "Launcher").

    python tools/run_reference.py <reference_root> <script.py> [script args...]

Only `model` is shadowed: `utils`, `dataloader`, `eval`, `submodules` resolve to the reference.  Offline
shims needed in this image: numpy.Inf, stub tensorboardX / clip, and -- when the
config's tokenizer is not a local directory -- a generated BertTokenizer directory.  The reference tree is
read-only here, so relative output paths in the JSON (./checkpoint, ./log) are rewritten into VCT_RUN_DIR.
"""
import json
import os
import runpy
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PRODUCT = os.path.join(ROOT, "video-captioning-transformer_b200")


def main():
    ref_root, script = os.path.abspath(sys.argv[1]), sys.argv[2]
    args = sys.argv[3:]
    import numpy as np
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    for name in ("tensorboardX", "clip"):
        try:
            __import__(name)
        except Exception:
            mod = types.ModuleType(name)
            if name == "tensorboardX":
                class SummaryWriter:
                    def __init__(self, *a, **k): pass
                    def add_scalar(self, *a, **k): pass
                    def close(self): pass
                mod.SummaryWriter = SummaryWriter
            sys.modules[name] = mod
    run_dir = os.environ.get("VCT_RUN_DIR", os.path.join(ROOT, "gpurun_out", "run"))
    os.makedirs(run_dir, exist_ok=True)
    # rewrite the config: offline tokenizer + writable output dirs
    if "-c" in args:
        i = args.index("-c") + 1
        with open(args[i]) as f:
            cfg = json.load(f)
        tok = cfg["model"].get("tokenizer", "")
        if not os.path.isdir(tok):
            sys.path.insert(0, PRODUCT)
            from vct.synthetic import make_tokenizer_dir
            cfg["model"]["tokenizer"] = make_tokenizer_dir(os.path.join(run_dir, "_tok"))
        if "train" in cfg:
            for k in ("save_dir", "log_dir"):
                if k in cfg["train"] and not os.path.isabs(cfg["train"][k]):
                    cfg["train"][k] = os.path.join(run_dir, os.path.basename(cfg["train"][k]))
                if k in cfg["train"]:
                    os.makedirs(cfg["train"][k], exist_ok=True)      # the reference assumes ./checkpoint and ./log exist
        patched = os.path.join(run_dir, f"config_{os.getpid()}.json")
        with open(patched, "w") as f:
            json.dump(cfg, f)
        args[i] = patched
    sys.path[:0] = [PRODUCT, ref_root]          # `model` -> ours, everything else -> the reference
    sys.argv = [os.path.join(ref_root, script)] + args
    runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
