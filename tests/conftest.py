"""Shared pytest plumbing: the `gpu` marker, repo imports, golden-fixture loaders."""

import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(GOLDEN / name, map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def golden_fusion():
    return load_golden("fusion.pt")


@pytest.fixture(scope="session")
def golden_rope():
    return load_golden("rope_merge.pt")


@pytest.fixture(scope="session")
def golden_qwen():
    return load_golden("qwen_tower_tiny.pt")


@pytest.fixture(scope="session")
def golden_vit():
    return load_golden("vit_tiny.pt")


@pytest.fixture(scope="session")
def golden_text_attention():
    return torch.load(GOLDEN / "text_attention.pt", map_location="cpu", weights_only=True)


@pytest.fixture(scope="session")
def golden_part2():
    return load_golden("part2_tiny.pt")
