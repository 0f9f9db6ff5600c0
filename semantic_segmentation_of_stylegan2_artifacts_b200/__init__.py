"""B200-native MS-UNet hot path: MSUNet forward/backward, DynamicLoss and Dice/IoU counting as
hand-written sm_100a CUDA kernels behind a C ABI (include/msunet_b200.h), with the reference's
Python signatures kept (network.MSUNet.MSUNet, loss.DynamicLoss.DynamicLoss,
scripts.validation_functions.*).  See DESIGN.md / INTEGRATION.md."""
from ._lib import LIB_PATH, launch_count, lib  # noqa: F401

__all__ = ["MSUNet", "MSUNetSys", "DynamicLoss", "lib", "launch_count", "LIB_PATH"]


def __getattr__(name):  # lazy: importing the package must not require a GPU or the .so
    if name == "MSUNet":
        from .network.MSUNet import MSUNet
        return MSUNet
    if name == "MSUNetSys":
        from .network.model_parts import MSUNetSys
        return MSUNetSys
    if name == "DynamicLoss":
        from .loss.DynamicLoss import DynamicLoss
        return DynamicLoss
    raise AttributeError(name)
