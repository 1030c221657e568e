"""raytrace_cpu_b200 -- B200-native Kerr null-geodesic ray tracing behind the raytrace_cpu class API.

The CUDA engine lives in libkerrtrace.so (raytrace_cpu_b200/csrc, C ABI in include/kerrtrace.h).  This
package is the Python mirror of the reference's host interface (see api.py); importing it loads no shared
library -- that happens on first use and fails loudly if the library or a GPU is missing.
"""
from . import _abi  # noqa: F401
from .api import (  # noqa: F401
    Context, DiscWithISCODestination, FlatDiscDestination, FlatPlaneDestination, ImagePlane, ImagePlaneBundles,
    Integrator, PointSource, RayDestination, Raytracer, TextOutput, caustic_discplane, disc_image, emissivity, emissivity_bins, emissivity_dev,
    get_context, image_bins, imageplane_desc, kerr_horizon, kerr_isco, plane_rays, pointsource_desc,
)
from .lib import KerrtraceError  # noqa: F401
