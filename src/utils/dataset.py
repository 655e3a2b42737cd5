"""Dataset constants used by the result writers (reference `src/utils/dataset.py:17-19`)."""

# occlusionLINEMOD numbers its 8 objects with their LINEMOD ids: index (1-based, as in the predictions) -> BOP object id
LMO_index_to_ID = ["1", "5", "6", "8", "9", "10", "11", "12"]
LMO_ID_to_index = {int(obj_id): idx + 1 for idx, obj_id in enumerate(LMO_index_to_ID)}


# names this file does not provide resolve from a reference checkout's copy of the same file (see src/__init__.py)
import src as _src  # noqa: E402

__getattr__ = _src.fallback_getattr(__name__)
