import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# BASELINE.json north_star: rgb/alpha and gradients within 1e-4 relative in fp32
RTOL = 1e-4


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def rel_err(a, b):
    """max |a-b| / max |b|: the scale-relative error used for every fp32 comparison."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if b.numel() == 0:
        return 0.0 if a.numel() == 0 else float("inf")
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)


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


def model_from_golden(g, device="cuda"):
    """The product's TemporalPoints carrying the reference's parameters from the golden file."""
    from articulated_point_nerf_b200.scene import make_scene, build_model
    scene = make_scene(g["config"])
    model = build_model(scene)
    missing, unexpected = model.load_state_dict(g["state_dict"], strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("tineuvox.") for k in missing), missing
    return model.to(device), scene


def oracle_for_scene(scene, model):
    """Oracle carrying the product model's parameters (for seeded cases beyond the golden file)."""
    from oracle.path_oracle import OraclePath
    state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items() if not k.startswith("tineuvox.")}
    cfg = scene.cfg
    return OraclePath(state, scene.canonical_pcd, scene.bones, stepsize=cfg.stepsize, voxel_size=scene.voxel_size,
                      fast_color_thres=cfg.fast_color_thres, act_shift=float(model.tineuvox.act_shift),
                      voxel_size_ratio=float(model.tineuvox.voxel_size_ratio), pose_embedding_dim=cfg.pose_embedding_dim)
