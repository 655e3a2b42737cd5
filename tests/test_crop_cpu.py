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


def test_M_maps_sampled_source_pixels_onto_their_output_pixels():
    """Size-independent property (any image / box): the matrix M returned with a crop sends the source pixel an output
    pixel was copied from back to within one output pixel of it (nearest sampling: < scale + 1), at the shipped size."""
    g = torch.Generator().manual_seed(5)
    H, W, T, n = 240, 320, 224, 12
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    coords = torch.stack([xx, yy]).unsqueeze(0).expand(n, 2, H, W) + 1.0          # +1: zero stays "padding"
    x1 = torch.randint(0, W - 40, (n,), generator=g); y1 = torch.randint(0, H - 40, (n,), generator=g)
    x2 = (x1 + torch.randint(20, 200, (n,), generator=g)).clamp(max=W); y2 = (y1 + torch.randint(20, 200, (n,), generator=g)).clamp(max=H)
    boxes = torch.stack([x1, y1, x2, y2], -1)
    out = port.crop_resize_pad(boxes, coords, target_size=T)
    oy, ox = torch.meshgrid(torch.arange(T, dtype=torch.float32), torch.arange(T, dtype=torch.float32), indexing="ij")
    for i in range(n):
        sx, sy = out["images"][i, 0] - 1.0, out["images"][i, 1] - 1.0             # source coordinates of every output pixel
        inside = out["images"][i, 0] > 0
        assert inside.any()
        M = out["M"][i]
        px = M[0, 0] * sx + M[0, 2]
        py = M[1, 1] * sy + M[1, 2]
        scale = M[0, 0].item()
        assert ((px - ox)[inside].abs() < scale + 1).all() and ((py - oy)[inside].abs() < scale + 1).all()
        assert ((px - ox)[inside] <= 1e-3).all() and ((py - oy)[inside] <= 1e-3).all()   # floor sampling never overshoots
