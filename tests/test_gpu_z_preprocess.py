"""Row f3: `gp_crop_resize_pad` (csrc/preprocess.cu) through the C ABI against the reference-generated golden crops, the
CPU oracle on random boxes, and the fused dataloader steps (reference crop.py:16-61, dataloader/train.py:80-123,
transform.yaml:2-7).  Images bit-exact (a gather), M to 1e-6."""
import os

import numpy as np
import pytest
import torch

from gigapose_b200 import preprocess
from oracle import make_golden_crop, port

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "crop.npz")


@pytest.mark.parametrize("case", ["t128", "t224"])
def test_kernel_reproduces_reference_crops(case):
    from src.utils.crop import CropResizePad
    g = np.load(GOLDEN)
    seed, n, C, H, W, T = (int(v) for v in g[f"{case}_cfg"])
    images, boxes = make_golden_crop.make_inputs(seed, n, C, H, W)
    out = CropResizePad(target_size=T)(boxes.to(DEV), images.to(DEV))
    assert torch.equal(out["images"].cpu(), torch.from_numpy(g[f"{case}_images"]))
    assert torch.allclose(out["M"].cpu(), torch.from_numpy(g[f"{case}_M"]), rtol=1e-6, atol=1e-6)


def _random_boxes(g, n, H, W):
    x1 = torch.randint(0, W - 8, (n,), generator=g); y1 = torch.randint(0, H - 8, (n,), generator=g)
    x2 = x1 + torch.randint(3, W, (n,), generator=g); y2 = y1 + torch.randint(3, H, (n,), generator=g)
    x2[::3] = x2[::3].clamp(max=W); y2[::3] = y2[::3].clamp(max=H)
    side = torch.minimum(x2[1::4] - x1[1::4], y2[1::4] - y1[1::4])
    x2[1::4], y2[1::4] = x1[1::4] + side, y1[1::4] + side
    return torch.stack([x1, y1, x2, y2], -1)


@pytest.mark.parametrize("T,H,W", [(224, 480, 640), (224, 97, 61), (128, 33, 200), (160, 240, 320)])
def test_kernel_matches_oracle_on_random_boxes(T, H, W):
    g = torch.Generator().manual_seed(T + H)
    n = 24
    images = torch.rand(n, 4, H, W, generator=g)
    boxes = _random_boxes(g, n, H, W)
    want = port.crop_resize_pad(boxes, images, target_size=T)
    got = preprocess.crop_resize_pad(boxes.to(DEV), images.to(DEV), target_size=T)
    assert torch.equal(got["images"].cpu(), want["images"])
    assert torch.allclose(got["M"].cpu(), want["M"], rtol=1e-6, atol=1e-6)


def test_fused_query_preprocessing_equals_the_dataloader_sequence():
    """rgb/255 -> x mask -> crop(rgba) -> Normalize(rgb), with detections indexing shared images (process_real)."""
    g = torch.Generator().manual_seed(9)
    m_img, n, H, W = 3, 10, 120, 160
    rgb = torch.randint(0, 256, (m_img, 3, H, W), generator=g, dtype=torch.uint8)
    masks = (torch.rand(n, H, W, generator=g) > 0.4).float()
    batch_im_id = torch.randint(0, m_img, (n,), generator=g)
    boxes = _random_boxes(g, n, H, W)
    # reference order of operations on the CPU (dataloader/train.py:80-113, transform.yaml:2-7)
    full = (rgb.float() / 255.0)[batch_im_id]
    m_rgba = torch.cat([full * masks[:, None], masks[:, None]], dim=1)
    cropped = port.crop_resize_pad(boxes, m_rgba, target_size=224)
    mean = torch.tensor(preprocess.CLIP_MEAN).view(1, 3, 1, 1); std = torch.tensor(preprocess.CLIP_STD).view(1, 3, 1, 1)
    want_img = (cropped["images"][:, :3] - mean) / std
    got = preprocess.preprocess_queries(rgb.to(DEV), masks.to(DEV), boxes.to(DEV), batch_im_id.to(DEV))
    assert torch.equal(got["tar_mask"].cpu(), cropped["images"][:, 3])
    assert torch.equal(got["tar_img"].cpu(), want_img)
    assert torch.allclose(got["tar_M"].cpu(), cropped["M"], rtol=1e-6, atol=1e-6)


def test_small_targets_are_refused():
    with pytest.raises(Exception, match="target_size"):
        preprocess.crop_resize_pad(torch.tensor([[0, 0, 8, 8]], device=DEV), torch.zeros(1, 1, 16, 16, device=DEV), target_size=56)
