import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run under gpurun)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices")


@pytest.fixture(scope="session")
def small_dataset(tmp_path_factory):
    """4 Parquet files, DATA_SPEC schema, globally unique key."""
    from ray_shuffling_data_loader_b200.data_generation import generate_data
    d = tmp_path_factory.mktemp("data_small")
    filenames, nbytes = generate_data(10_003, 4, 2, 0.0, str(d), seed=1234)
    return filenames, 10_003


@pytest.fixture(scope="session")
def float_dataset(tmp_path_factory):
    """3 files x 16 float32 columns + key."""
    import numpy as np
    from ray_shuffling_data_loader_b200.data_generation import generate_data, float_spec
    d = tmp_path_factory.mktemp("data_float")
    filenames, nbytes = generate_data(6_001, 3, 2, 0.0, str(d),
                                      data_spec=float_spec(16, np.float32), seed=7)
    return filenames, 6_001
