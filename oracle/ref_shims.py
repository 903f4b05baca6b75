"""Offline shims that let the REAL reference (``/root/reference``) be imported in this
container.  TEST INFRASTRUCTURE ONLY (see oracle/vct_oracle.py header).

The reference needs: a local BERT-style tokenizer directory
(``bert-base-uncased`` is not cached, no network), a stub ``clip`` module
(``TextEncoder("CLIP")`` would download ViT-B/32, model/TextEncoder.py:12-16), and --
for train.py itself -- ``numpy.Inf`` (utils.py:31) and a stub ``tensorboardX``
(train.py:15).  Nothing here is copied from the reference; the shims only stand in for
the third-party packages it imports.
"""
from __future__ import annotations

import json
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference() -> str:
    """/root/reference in the build container; on the GPU box the unmodified copy that baseline/install_reference.py
    placed under the git-ignored baseline/_ref/ (it travels with the gpurun snapshot)."""
    env = os.environ.get("VCT_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if os.path.isfile(os.path.join(cand, "model", "MMT4Caption.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_reference()
VOCAB_SIZE = 30522


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "MMT4Caption.py"))


def make_tokenizer_dir(path: str, vocab_size: int = VOCAB_SIZE) -> str:
    """Write a BertTokenizer directory with the bert-base-uncased special-token ids
    ([PAD]=0, [UNK]=100, [CLS]=101, [SEP]=102, [MASK]=103) and synthetic word entries."""
    os.makedirs(path, exist_ok=True)
    vocab_file = os.path.join(path, "vocab.txt")
    if not os.path.isfile(vocab_file):
        special = {0: "[PAD]", 100: "[UNK]", 101: "[CLS]", 102: "[SEP]", 103: "[MASK]"}
        with open(vocab_file, "w") as f:
            for i in range(vocab_size):
                f.write(special.get(i, f"[unused{i}]" if i < 1000 else f"w{i}") + "\n")
        with open(os.path.join(path, "tokenizer_config.json"), "w") as f:
            json.dump({"tokenizer_class": "BertTokenizer", "do_lower_case": True}, f)
    return path


def install_stub_modules() -> None:
    """Stub ``clip`` and ``tensorboardX``; restore ``numpy.Inf``."""
    import numpy as np
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    if "clip" not in sys.modules:
        clip = types.ModuleType("clip")

        class _Tower:
            def eval(self):
                return self

        clip.load = lambda name, device=None: (_Tower(), None)
        clip.tokenize = lambda captions: (_ for _ in ()).throw(RuntimeError("clip stub: no text tower offline"))
        sys.modules["clip"] = clip
    if "tensorboardX" not in sys.modules:
        tbx = types.ModuleType("tensorboardX")

        class SummaryWriter:
            def __init__(self, *a, **k):
                pass

            def add_scalar(self, *a, **k):
                pass

            def close(self):
                pass

        tbx.SummaryWriter = SummaryWriter
        sys.modules["tensorboardX"] = tbx


def import_reference_model():
    """Return the reference's ``model`` package (model.MMT4Caption etc.), imported from
    REFERENCE_ROOT under the private name ``_vct_ref_model`` so it never collides with the
    drop-in package that is also called ``model``."""
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    install_stub_modules()
    if "_vct_ref_model" in sys.modules:
        return sys.modules["_vct_ref_model"]
    import importlib.util
    # the reference does `from utils import generate_square_subsequent_mask` (model/CapDecoder.py:7):
    # expose its top-level utils.py under that name only while its package is being imported.
    saved_utils = sys.modules.get("utils")
    spec_u = importlib.util.spec_from_file_location("utils", os.path.join(REFERENCE_ROOT, "utils.py"))
    ref_utils = importlib.util.module_from_spec(spec_u)
    spec_u.loader.exec_module(ref_utils)
    sys.modules["utils"] = ref_utils
    try:
        # the reference has no model/__init__.py: build a namespace-style package by hand
        pkg = types.ModuleType("_vct_ref_model")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "model")]
        sys.modules["_vct_ref_model"] = pkg
        import importlib
        for sub in ("Embedding", "loss", "CapDecoder", "MMEncoder", "CapPreprocessor", "TextEncoder", "Matching",
                    "MMT4Caption"):
            setattr(pkg, sub, importlib.import_module(f"_vct_ref_model.{sub}"))
    finally:
        if saved_utils is not None:
            sys.modules["utils"] = saved_utils
        else:
            sys.modules.pop("utils", None)
    pkg.ref_utils = ref_utils
    return pkg


def shipped_model_config(tokenizer_dir: str, which: str = "msrvtt") -> dict:
    """The ``model`` block of the reference's shipped JSON with the tokenizer path rewritten."""
    name = {"msrvtt": "caption-task_baseline_modal_clip4clip_config.json",
            "msvd": "caption-task_baseline_modal_clip4clip_msvd_config.json"}[which]
    with open(os.path.join(REFERENCE_ROOT, "configs", name)) as f:
        cfg = json.load(f)["model"]
    cfg["tokenizer"] = tokenizer_dir
    return cfg
