"""`CropResizePad` -- drop-in for reference `src/utils/crop.py:11-61` (Hydra target
configs/data/transform.yaml:10-12) on CUDA tensors: the per-detection python loop of crop / interpolate / pad /
interpolate becomes one gather kernel (`gigapose_b200/csrc/preprocess.cu`, row f3 of SURVEY.md §8)."""
from gigapose_b200.preprocess import crop_resize_pad


class CropResizePad:
    def __init__(self, target_size=224, patch_size=14):
        self.target_size = target_size
        self.patch_size = patch_size

    def __call__(self, xyxy_boxes, images):
        """xyxy_boxes [n,4], images [n,C,H,W] (CUDA) -> {"M": [n,3,3], "images": [n,C,target,target]}."""
        out = crop_resize_pad(xyxy_boxes, images, self.target_size)
        return {"M": out["M"], "images": out["images"]}


# names this file does not provide resolve from a reference checkout's copy of the same file (see src/__init__.py)
import src as _src  # noqa: E402

__getattr__ = _src.fallback_getattr(__name__)
