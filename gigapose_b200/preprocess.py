"""Query pre-processing on the GPU (row f3): `crop_resize_pad` is `CropResizePad.__call__` (reference
src/utils/crop.py:16-61) as one gather kernel, optionally fused with the dataloader's element-wise steps
(`process_real` dataloader/train.py:80-123: /255 and x mask; CLIP normalisation configs/data/transform.yaml:2-7).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


@torch.no_grad()
def crop_resize_pad(xyxy_boxes: torch.Tensor, images: torch.Tensor, target_size: int = 224,
                    image_index: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None, in_div: float = 1.0,
                    mean: Optional[Sequence[float]] = None, std: Optional[Sequence[float]] = None):
    """xyxy_boxes [n,4], images [n,C,H,W] (or [m,C,H,W] with image_index [n]) on a CUDA device ->
    dict(images [n,C,T,T], M [n,3,3][, mask [n,T,T]]).  mask [n,H,W] is multiplied in before the crop and returned
    cropped; in_div / mean / std apply `(x / in_div * mask - mean) / std` in the reference's operation order."""
    if not images.is_cuda:
        raise _lib.GigaPoseNativeError("crop_resize_pad runs on CUDA tensors only (no CPU fallback)")
    lib = _lib.load()
    dev = images.device
    images = images.to(torch.float32).contiguous()
    boxes = torch.as_tensor(xyxy_boxes, device=dev).long().contiguous()      # BoundingBox.convert_long (bbox.py:18-22)
    n = boxes.shape[0]
    _, C, H, W = images.shape
    idx = None
    if image_index is not None:
        idx = torch.as_tensor(image_index, device=dev).to(torch.int32).contiguous()
        assert idx.shape == (n,)
    else:
        assert images.shape[0] == n, "one image per box unless image_index is given"
    m = None
    if mask is not None:
        m = mask.to(device=dev, dtype=torch.float32).contiguous()
        assert m.shape == (n, H, W), tuple(m.shape)
    sub = torch.tensor(list(mean), dtype=torch.float32, device=dev) if mean is not None else None
    div = torch.tensor(list(std), dtype=torch.float32, device=dev) if std is not None else None
    assert sub is None or sub.numel() == C
    assert div is None or div.numel() == C
    T = int(target_size)
    out = torch.empty(n, C, T, T, device=dev)
    out_mask = torch.empty(n, T, T, device=dev) if m is not None else None
    M = torch.empty(n, 3, 3, device=dev)
    ptr = lambda t: t.data_ptr() if t is not None else None
    with torch.cuda.device(dev):
        check(lib.gp_crop_resize_pad(n, C, H, W, T, images.data_ptr(), ptr(idx), boxes.data_ptr(), ptr(m), float(in_div),
                                     ptr(sub), ptr(div), out.data_ptr(), ptr(out_mask), M.data_ptr(),
                                     torch.cuda.current_stream(dev).cuda_stream))
    res = {"M": M, "images": out}
    if out_mask is not None:
        res["mask"] = out_mask
    return res


@torch.no_grad()
def preprocess_queries(rgb_u8: torch.Tensor, masks: torch.Tensor, xyxy_boxes: torch.Tensor, batch_im_id: torch.Tensor,
                       target_size: int = 224):
    """Detections -> network inputs in one launch: rgb_u8 [m,3,H,W] (0..255), masks [n,H,W] {0,1}, boxes [n,4],
    batch_im_id [n] -> tar_img [n,3,T,T] (masked, CLIP-normalised), tar_mask [n,T,T], tar_M [n,3,3]."""
    r = crop_resize_pad(xyxy_boxes, rgb_u8.to(torch.float32), target_size, image_index=batch_im_id, mask=masks, in_div=255.0,
                        mean=CLIP_MEAN, std=CLIP_STD)
    return {"tar_img": r["images"], "tar_mask": r["mask"], "tar_M": r["M"]}
