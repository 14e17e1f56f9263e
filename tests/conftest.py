import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["snelson_like_init", "road_like_trained", "kin_like_rbf", "house_like_warmstart",
                "ragged_rbf_init", "wide_d_matern", "song_like_wide", "restart_path", "kin_like_m256", "house_like_m256"]
GOLDEN_FP32_CASES = ["kin_like_rbf_fp32", "house_like_matern_fp32"]     # the reference's code run on float32 tensors
GRAD_NAMES = ["raw_noise", "mean_constant", "inducing_points", "raw_outputscale", "raw_lengthscale"]


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR
