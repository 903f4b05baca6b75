"""pytest configuration: registers the ``gpu`` marker and puts the repo root and the
product directory (``video-captioning-transformer_b200/``, which holds the drop-in ``model``
package and the ``vct`` host glue) on sys.path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT = os.path.join(ROOT, "video-captioning-transformer_b200")
for p in (PRODUCT, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def tokenizer_dir(tmp_path_factory):
    from vct.synthetic import make_tokenizer_dir
    return make_tokenizer_dir(str(tmp_path_factory.mktemp("tok")))
