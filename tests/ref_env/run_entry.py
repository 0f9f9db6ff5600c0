"""Run one of the reference's entry points (train.py / test.py, copied UNMODIFIED under baseline/_ref by
tools/install_reference.py) on the drop-in overlay:  python run_entry.py <script> <report.json> [script args...]

sys.path = [overlay (network / loss / scripts -> this repo), repo root, stand-ins for absent third-party packages, reference].
MSU_PATCH_ADAMW=1 makes `torch.optim.AdamW` the fused one-launch step."""
import json
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = "semantic_segmentation_of_stylegan2_artifacts_b200"
REF = os.environ.get("MSU_REFERENCE_DIR", os.path.join(ROOT, "baseline", "_ref"))
script, report = sys.argv[1], sys.argv[2]
sys.path[:0] = [os.path.join(ROOT, PKG, "dropin"), ROOT, HERE, REF]
sys.argv = [os.path.join(REF, script)] + sys.argv[3:]
if os.environ.get("MSU_PATCH_ADAMW") == "1":
    from semantic_segmentation_of_stylegan2_artifacts_b200.optim import patch_torch_adamw
    patch_torch_adamw()
runpy.run_path(os.path.join(REF, script), run_name="__main__")
import network.MSUNet as M  # noqa: E402
import loss.DynamicLoss as DL  # noqa: E402
import scripts.validation_functions as VF  # noqa: E402
import scripts.csv_handler as CH  # noqa: E402
import semantic_segmentation_of_stylegan2_artifacts_b200 as pkg  # noqa: E402
import torch  # noqa: E402
with open(report, "w") as f:
    json.dump({"msunet": M.MSUNet.__module__, "loss": DL.DynamicLoss.__module__, "metrics": VF.calculate_metrics.__module__,
               "csv_handler_file": os.path.abspath(CH.__file__), "launches": int(pkg.launch_count()), "lib": pkg.LIB_PATH,
               "adamw": torch.optim.AdamW.__module__}, f)
