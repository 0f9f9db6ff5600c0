from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import *  # noqa: F401,F403
from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys  # noqa: F401
