import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: tens of seconds (full-size benchmark configuration against the CPU oracle)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cfg_path(tmp_path_factory):
    from singleshotpose_b200.cfgs import write_cfg
    return write_cfg(str(tmp_path_factory.mktemp("cfg") / "yolo-pose.cfg"))


@pytest.fixture(scope="session")
def cfg_multi_path(tmp_path_factory):
    from singleshotpose_b200.cfgs import write_cfg
    return write_cfg(str(tmp_path_factory.mktemp("cfg") / "yolo-pose-multi.cfg"), multi=True)
