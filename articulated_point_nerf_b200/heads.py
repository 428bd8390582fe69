"""The sub-networks TemporalPoints borrows from the stage-1 model (`tineuvox.rgbnet`,
`tineuvox.densitynet`, `tineuvox.timenet`; lib/temporalpoints.py:133-147) and the constants it reads
from it (pos_poc / view_poc / time_poc, voxel_size_ratio, act_shift, no_view_dir).

The stage-1 TiNeuVox voxel backbone itself is out of scope; `TiNeuVoxHeads` carries
exactly the attributes the point-cloud path touches, under the same names, so a `TemporalPoints`
state dict keeps the reference's `rgbnet.*`, `densitynet.*`, `timenet.*` keys.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn


def poc_fre(input_data: torch.Tensor, poc_buf: torch.Tensor) -> torch.Tensor:
    """lib/tineuvox.py:872-878: [x, sin(x*2^i), cos(x*2^i)], coordinate-major then frequency."""
    input_data_emb = (input_data.unsqueeze(-1) * poc_buf).flatten(-2)
    return torch.cat([input_data, input_data_emb.sin(), input_data_emb.cos()], -1)


class RGBNet(nn.Module):
    """lib/tineuvox.py:65-88 (parameter names kept: feature_linears, views_linears.{0,2})."""

    def __init__(self, D=3, W=128, h_ch=128, views_ch=27, pts_ch=63, times_ch=17, output_ch=3):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = h_ch, views_ch
        self.input_ch_pts, self.input_ch_times, self.output_ch = pts_ch, times_ch, output_ch
        self.feature_linears = nn.Linear(self.input_ch, W)
        self.views_linears = nn.Sequential(nn.Linear(W + self.input_ch_views, W // 2), nn.ReLU(),
                                           nn.Linear(W // 2, self.output_ch))

    def forward(self, input_h, input_views=None):
        feature = self.feature_linears(input_h)
        if input_views is not None:
            feature = torch.cat([feature, input_views], dim=-1)
        else:
            assert self.input_ch_views == 0
        return self.views_linears(feature)


class TiNeuVoxHeads(nn.Module):
    """What `tineuvox` must provide to TemporalPoints (lib/temporalpoints.py:118-147,498-500)."""

    def __init__(self, xyz_min, xyz_max, num_voxels=160 ** 3, num_voxels_base=160 ** 3, alpha_init=1e-3,
                 net_width=128, voxel_dim=4, posbase_pe=10, viewbase_pe=4, timebase_pe=8, gridbase_pe=2,
                 no_view_dir=False, **kwargs):
        super().__init__()
        self.posbase_pe, self.viewbase_pe, self.timebase_pe, self.gridbase_pe = posbase_pe, viewbase_pe, timebase_pe, gridbase_pe
        self.no_view_dir = no_view_dir
        self.net_width = net_width
        self.register_buffer('xyz_min', torch.as_tensor(np.asarray(xyz_min)).float())
        self.register_buffer('xyz_max', torch.as_tensor(np.asarray(xyz_max)).float())
        self.num_voxels_base, self.num_voxels = num_voxels_base, num_voxels
        self.voxel_size_base = ((self.xyz_max - self.xyz_min).prod() / num_voxels_base).pow(1 / 3)
        self.voxel_size = ((self.xyz_max - self.xyz_min).prod() / num_voxels).pow(1 / 3)
        self.voxel_size_ratio = self.voxel_size / self.voxel_size_base      # lib/tineuvox.py:172-175
        self.alpha_init = alpha_init
        self.act_shift = np.log(1 / (1 - alpha_init) - 1)                    # lib/tineuvox.py:126
        times_ch = 2 * timebase_pe + 1
        views_ch = 0 if no_view_dir else 3 + 3 * viewbase_pe * 2
        timenet_output = voxel_dim + voxel_dim * 2 * gridbase_pe
        self.timenet = nn.Sequential(nn.Linear(times_ch, net_width), nn.ReLU(inplace=True),
                                     nn.Linear(net_width, timenet_output))
        self.densitynet = nn.Linear(net_width, 1)
        self.rgbnet = RGBNet(W=net_width, h_ch=net_width, views_ch=views_ch, pts_ch=3 + 3 * posbase_pe * 2,
                             times_ch=times_ch)
        self.register_buffer('time_poc', torch.FloatTensor([(2 ** i) for i in range(timebase_pe)]))
        self.register_buffer('grid_poc', torch.FloatTensor([(2 ** i) for i in range(gridbase_pe)]))
        self.register_buffer('pos_poc', torch.FloatTensor([(2 ** i) for i in range(posbase_pe)]))
        self.register_buffer('view_poc', torch.FloatTensor([(2 ** i) for i in range(viewbase_pe)]))
