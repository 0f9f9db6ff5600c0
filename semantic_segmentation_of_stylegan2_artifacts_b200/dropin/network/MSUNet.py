from semantic_segmentation_of_stylegan2_artifacts_b200.network.MSUNet import *  # noqa: F401,F403
from semantic_segmentation_of_stylegan2_artifacts_b200.network.MSUNet import MSUNet  # noqa: F401
