"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/crop.npz by running the UNMODIFIED reference `CropResizePad`
(/root/reference/src/utils/crop.py:11-61; build container only) on seeded images and boxes.

    python -m oracle.make_golden_crop
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ref_import

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "crop.npz")


def make_inputs(seed, n, C, H, W):
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(n, C, H, W, generator=g)
    x1 = torch.randint(0, W - 20, (n,), generator=g)
    y1 = torch.randint(0, H - 20, (n,), generator=g)
    x2 = x1 + torch.randint(8, W, (n,), generator=g)
    y2 = y1 + torch.randint(8, H, (n,), generator=g)
    x2[::2] = x2[::2].clamp(max=W)                       # every other box stays inside the image
    y2[::2] = y2[::2].clamp(max=H)
    side = torch.minimum(x2[-1] - x1[-1], y2[-1] - y1[-1])
    x2[-1], y2[-1] = x1[-1] + side, y1[-1] + side        # one square box (no padding branch)
    return images, torch.stack([x1, y1, x2, y2], -1)


CASES = {"t224": dict(seed=51, n=6, C=2, H=120, W=160, T=224), "t128": dict(seed=52, n=5, C=1, H=97, W=61, T=128),
         "t56": dict(seed=53, n=5, C=1, H=50, W=70, T=56)}      # t56: ATen's small-output kernel (oracle only)


def main():
    with ref_import._ReferenceImports() as ctx:
        crop = ctx.import_reference("src.utils.crop")
    arrays = {}
    for name, c in CASES.items():
        images, boxes = make_inputs(c["seed"], c["n"], c["C"], c["H"], c["W"])
        out = crop.CropResizePad(target_size=c["T"])(boxes, images)
        arrays[f"{name}_boxes"] = boxes.numpy()
        arrays[f"{name}_images"] = out["images"].numpy()
        arrays[f"{name}_M"] = out["M"].numpy()
        arrays[f"{name}_cfg"] = np.array([c["seed"], c["n"], c["C"], c["H"], c["W"], c["T"]])
        arrays[f"{name}_ck_in"] = np.float64(images.double().sum().item())
        print(name, tuple(out["images"].shape), boxes.tolist())
    np.savez_compressed(GOLDEN, **arrays)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    main()
