import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


SAMPLER_GOLDENS = ["s_nh0_T6_L20", "s_nh1_T7_L24", "s_nh2_T5_L150", "s_nh5_T9_L40"]
TRAIN_GOLDENS = ["t_nh2_T7_L24", "t_nh0_T6_L20"]
