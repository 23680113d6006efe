"""B200-native multiscale Generalized-ICP refinement: drop-in for the reference's Multiscale_GICP.

The directory name of this package is not a Python identifier; import it through the ``mgicp_b200``
shim at the repository root.
"""
from . import global_refinement, pcd_io, poses, shard, synthetic  # noqa: F401
from . import pose_graph  # noqa: F401,E402
from .engine import BatchResult, Engine, MgicpError  # noqa: F401
from .stream import BatchStream  # noqa: F401
from .registration import (Multiscale_GICP, RegistrationResult, create_scales, create_scales_script2,  # noqa: F401
                           max_correspondence_distances, multiscale_gicp, multiscale_gicp_batch, radius_from_cloud_pair,
                           evaluate_registration, get_information_matrix_from_point_clouds, calculate_RMSE_and_fitness,
                           Coarse_to_fine_M_GICP, registro_FGR, Coarse_to_fine_FGR_M_GICP)
