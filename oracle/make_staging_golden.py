"""Generate tests/golden/staging.npz by RUNNING THE REFERENCE's per-sample transform (build container only).

Imports dataset/dataset.py:RandomGenerator from /root/reference (cv2 / albumentations / matplotlib are not installed and are not
touched with transform=False, so empty stand-in modules satisfy the import lines dataset.py:8-10), feeds it seeded uint8
images / labels and stores inputs, the flip decisions (replayed from the same `random` seed: dataset.py:45 draws once, :50 once)
and the tensors it returns.  The fixture travels; the reference does not.

    python oracle/make_staging_golden.py
"""
import os
import random
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for name in ("cv2", "albumentations", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, "/root/reference")
from dataset.dataset import RandomGenerator  # noqa: E402  (reference)

H, W, N = 12, 20, 6
rng = np.random.default_rng(20240607)
images = rng.integers(0, 256, size=(N, H, W, 3), dtype=np.uint8)
labels = rng.integers(0, 256, size=(N, H, W), dtype=np.uint8)
labels[0, 0, :4] = (126, 127, 128, 129)                       # the threshold edge of dataset.py:63
gen = RandomGenerator((H, W), random_flip_flag=True, transform=False)
out_img, out_lab, flips = [], [], []
for i in range(N):
    random.seed(100 + i)
    random.random()
    flips.append(random.random() > 0.5)
    random.seed(100 + i)
    s = gen({"image": images[i], "label": labels[i]})
    out_img.append(s["image"].numpy())
    out_lab.append(s["label"].numpy())
assert any(flips) and not all(flips)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "staging.npz"), images=images, labels=labels,
                    flips=np.array(flips, dtype=np.uint8), out_image=np.stack(out_img), out_label=np.stack(out_lab))
print("flips", flips, "image", np.stack(out_img).shape, np.stack(out_img).dtype, "label", np.stack(out_lab).dtype)
