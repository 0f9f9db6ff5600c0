from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss  # noqa: F401
