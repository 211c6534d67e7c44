"""stereo_toolbox_b200 -- Blackwell-native (sm_100a) cost-volume hot path behind the model API of
xxxupeng/stereo_toolbox.  See DESIGN.md for scope, INTEGRATION.md for the drop-in recipe."""
from .functional import (build_gwc_volume, build_concat_volume, build_concat_volume_unmasked,
                         groupwise_correlation, disparity_regression, disparityregression,
                         upsample_softargmin, CorrBlock1D, Combined_Geo_Encoding_Volume)
from .gwcnet import GwcNet_G, GwcNet_GC
from .psmnet import PSMNet
from .raft_stereo import RAFTStereo
from .acvnet import ACVNet
from .cfnet import CFNet
from .pcwnet import PCWNet_G, PCWNet_GC
from .igev_stereo import IGEVStereo
from .checkpoint import load_checkpoint_flexible
from .evaluation import speed_and_memory_test
from . import ops as _ops

_ops.register_torch_ops()      # torch.ops.stb200.* (dispatcher entries + Meta kernels + autograd formulas); no CUDA needed to register

__all__ = ["build_gwc_volume", "build_concat_volume", "build_concat_volume_unmasked", "groupwise_correlation",
           "disparity_regression", "disparityregression", "upsample_softargmin", "CorrBlock1D",
           "Combined_Geo_Encoding_Volume", "GwcNet_G", "GwcNet_GC", "PSMNet", "RAFTStereo", "ACVNet", "CFNet", "PCWNet_G", "PCWNet_GC", "IGEVStereo",
           "load_checkpoint_flexible", "speed_and_memory_test"]
