"""Stage-1 -> stage-2 hand-over: `export_point_cloud` (run.py:1081-1240), host side.

The reference samples its stage-1 voxel model on a lattice (`TiNeuVox.get_grid_as_point_cloud`, lib/tineuvox.py:253-372),
bisects the lattice frequency until the thresholded, cleaned density volume holds ~`canonical_pcd_num` points, writes
`pcds/canonical.tar`, runs the skeletoniser on the same volume and writes `pcds/skeleton.tar`.  The voxel model is NOT part
of the point-cloud hot path, so it enters here duck-typed as `field` — anything with the two methods the reference calls
(the reference's own loaded TiNeuVox works unchanged):

    field.get_grid_as_point_cloud(stepsize=, time_sel=, viewdir=, threshold=, sampling_freq=, N_batch=, alpha_xyz_only=,
                                  grid_xyz=) -> (points, alphas, rgbs, feat, raw_feat, binary_volume, grid_xyz, alpha_volume)
    field.get_grid_xyz(sampling_freq) -> (X, Y, Z, 3)          field.voxel_size

What is restated here: the control flow of run.py:1128-1191 (coarse +-0.1 search for a bracket, ten bisection steps), the volume
clean-up `preprocess_volume` (run.py:1135-1142: optional gaussian, threshold, remove_small_holes(area_threshold=256), largest
26-connected component) on scipy.ndimage instead of skimage / cc3d (not installed here), and the two file layouts
(run.py:1091-1103, 1229-1235).  The skeletoniser (`create_skeleton`, skeletonizer.py:209-327: skimage's 3-D thinning + a
graph walk) is a plug-in: pass `create_skeleton=`; by default the reference's own module is imported if it is on the path.
Parity: UNPINNED against the reference (its function needs open3d / skimage / cc3d, none of which exist in this image);
tests/test_cpu_host.py checks the behaviour on an analytic field and that the written files feed `model_from_pcds`.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import numpy as np
import torch


def preprocess_volume(alpha_volume: np.ndarray, threshold: float, sigma: float = 0.0) -> np.ndarray:
    """run.py:1135-1142.  alpha (X,Y,Z) -> bool (X,Y,Z): [gaussian] -> > threshold -> holes smaller than 2^8 voxels filled
    (skimage.morphology.remove_small_holes: background components by face connectivity) -> largest 26-connected component."""
    from scipy import ndimage
    vol = np.asarray(alpha_volume, dtype=np.float64)
    if sigma > 0:
        vol = ndimage.gaussian_filter(vol, sigma=sigma, mode="nearest")       # skimage.filters.gaussian(..., preserve_range=True)
    binary = vol > threshold
    holes, n = ndimage.label(~binary)                                          # connectivity 1, as remove_small_holes uses
    if n:
        sizes = np.bincount(holes.ravel())
        small = sizes < 2 ** 8
        small[0] = False
        binary = binary | small[holes]
    comp, n = ndimage.label(binary, structure=np.ones((3, 3, 3), dtype=bool))  # cc3d.largest_k(connectivity=26, k=1)
    if n > 1:
        sizes = np.bincount(comp.ravel())
        sizes[0] = 0
        binary = comp == int(sizes.argmax())
    return binary.astype(bool)


def _sample(field, stepsize, t, viewdir, threshold, freq, n_batch):
    out = field.get_grid_as_point_cloud(stepsize=stepsize, time_sel=t, viewdir=viewdir, threshold=threshold, sampling_freq=freq,
                                        N_batch=n_batch, alpha_xyz_only=True)
    grid_xyz, alpha_volume = out[6], out[7]
    return grid_xyz, alpha_volume


def find_sampling_frequency(field, stepsize, t, viewdir, threshold: float, target: int, sigma: float = 0.0, n_batch: int = 2 ** 21,
                            verbose: bool = False):
    """run.py:1128-1191: -> (freq, grid_xyz, alpha_volume, mask) with mask.sum() as close to `target` as ten bisection steps get."""
    def count(freq):
        grid_xyz, alpha = _sample(field, stepsize, t, viewdir, threshold, freq, n_batch)
        mask = preprocess_volume(alpha.cpu().numpy(), threshold, sigma)
        return grid_xyz, alpha, mask, int(mask.sum())

    freq = 1.0
    grid_xyz, alpha, mask, n = count(freq)
    up = low = None
    if n > target:
        up, step = freq, -0.1
    elif n < target:
        low, step = freq, +0.1
    else:
        return freq, grid_xyz, alpha, mask
    while up is None or low is None:                       # coarse search for the other end of the bracket
        freq = freq + step
        if freq <= 0:
            raise ValueError("the density volume holds more points than the target at every sampling frequency")
        grid_xyz, alpha, mask, n = count(freq)
        if n > target:
            up = freq
        elif n < target:
            low = freq
        else:
            return freq, grid_xyz, alpha, mask
    for _ in range(10):                                    # bisection
        freq = (up + low) / 2
        grid_xyz, alpha, mask, n = count(freq)
        if verbose:
            print(f"Canonical sampling freq: {freq}, num points: {n}")
        if n > target:
            up = freq
        elif n < target:
            low = freq
        else:
            break
    return freq, grid_xyz, alpha, mask


def export_point_cloud(field, path: str, viewdir, stepsize: float, canonical_t: float = 0.0, threshold: float = 0.2,
                       bone_length: float = 4.0, canonical_pcd_num: float = 3e4, skeleton_density_threshold: float = 0.2,
                       create_skeleton: Optional[Callable] = None, verbose: bool = False):
    """run.py:1081-1240.  Writes `<path>/pcds/canonical.tar` (+ `skeleton.tar` when a skeletoniser is available); returns the
    dict stored in canonical.tar.  `viewdir` (1,3): the mean view direction of the first training camera (run.py:1147-1151).
    Existing exports are left alone, as in the reference (run.py:1087-1089)."""
    folder = os.path.join(path, 'pcds')
    os.makedirs(folder, exist_ok=True)
    can_path, skel_path = os.path.join(folder, 'canonical.tar'), os.path.join(folder, 'skeleton.tar')
    if os.path.isfile(can_path) and os.path.isfile(skel_path):
        if verbose:
            print('PCD and skeleton already exists, skipping export.')
        return torch.load(can_path, map_location='cpu', weights_only=False)
    t = torch.tensor([canonical_t])
    viewdir = torch.as_tensor(viewdir, dtype=torch.float32).reshape(1, 3)
    freq, grid_xyz, alpha_volume, mask = find_sampling_frequency(field, stepsize, t, viewdir, threshold, int(canonical_pcd_num),
                                                                 verbose=verbose)
    sel = grid_xyz[torch.as_tensor(mask)]
    points, alphas, rgbs, feat, raw_feat, _, _, _ = field.get_grid_as_point_cloud(
        stepsize=stepsize, time_sel=t, viewdir=viewdir, threshold=threshold, sampling_freq=freq, N_batch=2 ** 21,
        alpha_xyz_only=False, grid_xyz=sel)
    can = {'pcd': points, 'rgbs': rgbs, 'feat': feat, 'raw_feat': raw_feat, 'alphas': alphas, 't': float(canonical_t),
           'xyz_min': points.min(dim=0)[0], 'xyz_max': points.max(dim=0)[0], 'voxel_size': field.voxel_size}     # run.py:1091-1103
    torch.save(can, can_path)
    if create_skeleton is None:
        try:
            from skeletonizer import create_skeleton          # the reference's own module, if the caller has it on the path
        except Exception:
            create_skeleton = None
    if create_skeleton is not None:
        binary_volume = preprocess_volume(alpha_volume.cpu().numpy(), skeleton_density_threshold, 0.0)
        res = create_skeleton(binary_volume, field.get_grid_xyz(freq).cpu().numpy(), bone_length=bone_length)    # run.py:1206-1230
        torch.save(res, skel_path)
        if verbose:
            print(f"{len(res['bones'])} bones extracted.")
    elif verbose:
        print("no skeletoniser available (skimage / the reference's skeletonizer.py): wrote canonical.tar only")
    return can
