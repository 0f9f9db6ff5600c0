from semantic_segmentation_of_stylegan2_artifacts_b200.scripts.validation_functions import *  # noqa: F401,F403
