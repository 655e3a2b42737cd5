"""Result writers of the inference path (row f4 of SURVEY.md §8): BOP-challenge csv export of the per-batch prediction
files `GigaPose.filter_and_save` writes.  Drop-in for the functions `GigaPose.on_test_epoch_end` and the refiner call
in the reference `src/utils/inout.py` (`save_bop_results` :126, `load_bop_results` :156, `averaging_runtime_bop_results`
:200, `calculate_runtime_per_image` :222, `save_predictions_from_batched_predictions` :273); output files are
byte-identical to the reference's (tests/test_writers_cpu.py against reference-generated fixtures).

Host-side code: nothing here touches the GPU.
"""
from __future__ import annotations

import os
import os.path as osp
from collections import OrderedDict

import numpy as np

from src.utils.dataset import LMO_index_to_ID
from src.utils.logging import get_logger

logger = get_logger(__name__)

_HEADER = "scene_id,im_id,obj_id,score,R,t,time"


def _image_key(result):
    return f"{int(result['scene_id']):06d}_{int(result['im_id']):06d}"


def _floats(values):
    """Space-separated shortest-repr decimals of the (float32 or float64) values widened to Python floats."""
    return " ".join(str(v) for v in np.asarray(values).flatten().tolist())


def save_bop_results(path, results, additional_name=None):
    """One line per estimate: scene_id,im_id,obj_id,score,R (9 values),t (3 values),time[,<additional_name>]
    (the BOP toolkit's results format; `time` = -1 when absent).  No trailing newline."""
    lines = [_HEADER if additional_name is None else f"{_HEADER},{additional_name}"]
    for res in results:
        fields = [res["scene_id"], res["im_id"], res["obj_id"], res["score"], _floats(res["R"]), _floats(res["t"]),
                  res.get("time", -1)]
        if additional_name is not None:
            fields.append(res[additional_name])
        lines.append(",".join("{}".format(f) for f in fields))
    with open(path, "w") as f:
        f.write("\n".join(lines))


def load_bop_results(path, additional_name=None):
    """Inverse of `save_bop_results`: list of dicts with R [3,3] and t [3,1] as float64."""
    header = _HEADER if additional_name is None else f"{_HEADER},{additional_name}"
    n_fields = 7 if additional_name is None else 8
    results = []
    with open(path, "r") as f:
        for line_id, line in enumerate(f):
            if line_id == 0 and header in line:
                continue
            elems = line.split(",")
            if len(elems) != n_fields:
                raise ValueError("A line does not have {} comma-sep. elements: {}".format(n_fields, line))
            res = {"scene_id": int(elems[0]), "im_id": int(elems[1]), "obj_id": int(elems[2]), "score": float(elems[3]),
                   "R": np.array([float(x) for x in elems[4].split()], np.float64).reshape(3, 3),
                   "t": np.array([float(x) for x in elems[5].split()], np.float64).reshape(3, 1),
                   "time": float(elems[6])}
            if additional_name is not None:
                res[additional_name] = float(elems[7])
            results.append(res)
    return results


def averaging_runtime_bop_results(path, has_instance_id=False):
    """Rewrites `path` with every estimate's time replaced by the mean time of its image."""
    results = load_bop_results(path, has_instance_id)
    per_image = OrderedDict()
    for res in results:
        per_image.setdefault(_image_key(res), []).append(res["time"])
    mean_time = {key: np.mean(times) for key, times in per_image.items()}
    for res in results:
        res["time"] = mean_time[_image_key(res)]
    save_bop_results(path, results, has_instance_id)


def calculate_runtime_per_image(results, is_refined):
    """BOP run time of an image = its detection time + the time of every batch that contained one of its detections
    (coarse stage), or the batch times + the refinement times (refined stage).  Each (image, batch) pair counts once;
    the image's detection time is the one carried by the last batch that introduced it.  Every estimate's `time`
    becomes its image's total; the bookkeeping fields `additional_time` / `batch_id` are dropped."""
    batch_times, extra, seen = OrderedDict(), {}, {}
    for res in results:
        key = _image_key(res)
        assert "batch_id" in res, f"batch_id is not in {res}"
        batches = seen.setdefault(key, [])
        if res["batch_id"] not in batches:
            batches.append(res["batch_id"])
            batch_times.setdefault(key, []).append(res["time"])
            if is_refined:
                extra.setdefault(key, []).append(res["additional_time"])
            else:
                extra[key] = res["additional_time"]
        del res["additional_time"], res["batch_id"]
    totals = {}
    for key, times in batch_times.items():
        totals[key] = (np.sum(extra[key]) if is_refined else extra[key]) + np.sum(times)
    for res in results:
        res["time"] = totals[_image_key(res)]
    if results:
        logger.info(f"Average runtime per image: {np.mean([res['time'] for res in results]):.3f} s")
    return results


def save_predictions_from_batched_predictions(prediction_dir, dataset_name, model_name, run_id, is_refined):
    """Collects `{prediction_dir}/*.npz` (sorted by file name; keys scene_id, im_id, object_id, time,
    detection_time | refinement_time, poses [n,k,4,4] or [n,4,4], scores [n,k] or [n]) into
    `{model}-pbrreal-rgb-mmodel_{dataset}-test_{run_id}.csv` (best hypothesis per detection) and, when the files hold
    k hypotheses, `...MultiHypothesis.csv` (all k, with an `instance_id` column)."""
    files = sorted(f for f in os.listdir(prediction_dir) if f.endswith(".npz"))
    extra_name = "refinement_time" if is_refined else "detection_time"
    remap_lmo = (not is_refined) and "lmo" in dataset_name
    top1, topk = [], []
    has_hypotheses = False
    instance_id = 0
    for batch_id, name in enumerate(files):
        data = np.load(osp.join(prediction_dir, name))
        poses, scores = data["poses"], data["scores"]
        assert poses.ndim in (3, 4)
        if poses.ndim == 3:                              # top-1 only files: treat as one hypothesis
            poses, scores = poses[:, None], scores[:, None]
        else:
            has_hypotheses = True
        for i in range(len(data["im_id"])):
            obj_id = int(data["object_id"][i])
            if remap_lmo:
                obj_id = LMO_index_to_ID[obj_id - 1]
            for j in range(poses.shape[1]):
                est = dict(scene_id=int(data["scene_id"][i]), im_id=int(data["im_id"][i]), obj_id=obj_id,
                           score=scores[i][j], t=poses[i][j][:3, 3].reshape(-1), R=poses[i][j][:3, :3].reshape(-1),
                           time=data["time"][i], additional_time=data[extra_name][i], batch_id=batch_id,
                           instance_id=instance_id)
                topk.append(est)
                if j == 0:
                    top1.append(dict(est))
            instance_id += 1
    stem = f"{model_name}-pbrreal-rgb-mmodel_{dataset_name}-test_{run_id}"
    path = osp.join(prediction_dir, f"{stem}.csv")
    save_bop_results(path, calculate_runtime_per_image(top1, is_refined=is_refined))
    logger.info(f"Saved predictions to {path}")
    if has_hypotheses:
        path = osp.join(prediction_dir, f"{stem}MultiHypothesis.csv")
        save_bop_results(path, calculate_runtime_per_image(topk, is_refined=is_refined), additional_name="instance_id")
        logger.info(f"Saved predictions to {path}")



# everything this file does not provide (the CNOS-detection / test-list loaders the reference's dataloaders import from
# `src.utils.inout`) comes from the reference checkout's own `src/utils/inout.py` when one is on the path
import src as _src  # noqa: E402

__getattr__ = _src.fallback_getattr(__name__)
