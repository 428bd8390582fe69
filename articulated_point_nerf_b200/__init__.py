"""B200-native (sm_100a) point-cloud render path of Articulated-Point-NeRF.

Python surface mirrors the reference's lib/temporalpoints.py, lib/pointwarper.py, lib/masked_adam.py and
the pybind modules render_utils_cuda / adam_upd_cuda; compute lives in libapn_sm100.so (include/apn.h).
"""
from .heads import RGBNet, TiNeuVoxHeads, poc_fre                      # noqa: F401
from .masked_adam import MaskedAdam                                    # noqa: F401
from .pointwarper import PointWarper, TransformNet                     # noqa: F401
from .render_utils import (Alphas2Weights, Raw2Alpha, adam_upd_cuda,   # noqa: F401
                           render_utils_cuda)
from .temporalpoints import NoPointsException, TemporalPoints          # noqa: F401
from .render import (PoseCache, load_checkpoint, model_from_pcds,      # noqa: F401
                     render_repose, render_viewpoints, save_checkpoint,
                     save_pcds)
from .export import export_point_cloud                                 # noqa: F401

__version__ = "0.1.0"
