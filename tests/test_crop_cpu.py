"""Row f3 oracle: `oracle.port.crop_resize_pad` against outputs of the UNMODIFIED reference `CropResizePad`
(tests/golden/crop.npz from `python -m oracle.make_golden_crop`; reference src/utils/crop.py:11-61).  Bit-exact images."""
import os

import numpy as np
import pytest
import torch

from oracle import make_golden_crop, port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "crop.npz")


@pytest.mark.parametrize("case", sorted(make_golden_crop.CASES))
def test_port_reproduces_reference_crops(case):
    g = np.load(GOLDEN)
    seed, n, C, H, W, T = (int(v) for v in g[f"{case}_cfg"])
    images, boxes = make_golden_crop.make_inputs(seed, n, C, H, W)
    assert np.isclose(images.double().sum().item(), float(g[f"{case}_ck_in"]), rtol=0, atol=1e-9)     # same RNG stream
    assert np.array_equal(boxes.numpy(), g[f"{case}_boxes"])
    out = port.crop_resize_pad(boxes, images, target_size=T)
    assert torch.equal(out["images"], torch.from_numpy(g[f"{case}_images"]))
    assert torch.allclose(out["M"], torch.from_numpy(g[f"{case}_M"]), rtol=1e-6, atol=1e-6)


def test_known_answer_square_box_is_a_pure_scale():
    """A 32x32 box scaled to 128: no padding, M = scale 4 about the box corner, pixels repeat 4x4."""
    img = torch.arange(64 * 64, dtype=torch.float32).reshape(1, 1, 64, 64)
    out = port.crop_resize_pad(torch.tensor([[8, 16, 40, 48]]), img, target_size=128)
    assert torch.equal(out["M"][0], torch.tensor([[4.0, 0, -32.0], [0, 4.0, -64.0], [0, 0, 1.0]]))
    assert torch.equal(out["images"][0, 0, ::4, ::4], img[0, 0, 16:48, 8:40])
    assert torch.equal(out["images"][0, 0, 3::4, 3::4], img[0, 0, 16:48, 8:40])


def test_known_answer_wide_box_is_centred_with_zero_rows():
    img = torch.ones(1, 1, 40, 100)
    out = port.crop_resize_pad(torch.tensor([[10, 10, 74, 26]]), img, target_size=128)     # 64 x 16 -> 128 x 32
    rows = out["images"][0, 0].sum(1)
    assert int((rows > 0).sum()) == 32 and bool((rows[:48] == 0).all()) and bool((rows[80:] == 0).all())
    assert out["M"][0, 1, 2].item() == 48 - 2 * 10 and out["M"][0, 0, 2].item() == -2 * 10
