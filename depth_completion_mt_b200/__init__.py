"""depth_completion_mt_b200 -- B200 (sm_100a) implementation of the depth-completion hot path of
PatrizioPerugini/depth_completion_MT behind the reference's own call surface.

The compute lives in ``libdcmt.so`` (hand-written CUDA behind the C ABI of include/dcmt.h); this
package is the thin Python host side.  There is no CPU fallback: without the built library the
compute functions raise.
"""
from .api import (  # noqa: F401
    calculateMeasuementDerivatives,
    get_initial_disparity,
    img_completion,
    interpolate_with_superpixels,
    optimize_IG,
    retrieve_optimized_depth,
    stereo_params,
    stereo_refine,
)
from ._lib import DcmtError, StereoParams  # noqa: F401

__version__ = "0.1.0"
