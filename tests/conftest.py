import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_tiny():
    return torch.load(os.path.join(GOLDEN_DIR, "ref_tiny.pt"), weights_only=False)


def oracle_from_golden(g):
    from oracle.path_oracle import OraclePath
    from articulated_point_nerf_b200.scene import CONFIGS
    cfg = CONFIGS[g["config"]]
    state = {k: v.clone() for k, v in g["state_dict"].items()}
    return OraclePath(state, g["canonical_pcd"], g["bones"], stepsize=cfg.stepsize, voxel_size=g["voxel_size"],
                      fast_color_thres=cfg.fast_color_thres, act_shift=g["act_shift"],
                      voxel_size_ratio=g["voxel_size_ratio"], mean_min_distance=g["mean_min_distance"],
                      pose_embedding_dim=cfg.pose_embedding_dim), cfg


@pytest.fixture(scope="session")
def oracle_tiny(golden_tiny):
    return oracle_from_golden(golden_tiny)
