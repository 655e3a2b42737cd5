"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/bop_csv/** by running the UNMODIFIED reference result writer
(`save_predictions_from_batched_predictions`, /root/reference/src/utils/inout.py:273-367; build container only) on small
seeded per-batch prediction files of the schema `GigaPose.filter_and_save` writes (gigaPose.py:439-448).

    python -m oracle.make_golden_csv
"""
from __future__ import annotations

import os
import shutil
import tempfile

import numpy as np

from . import ref_import

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bop_csv")

# name -> (dataset_name, hypotheses per detection (None = top-1 only files), detections per batch file)
CASES = {
    "lmo_k5": ("lmo", 5, [5, 4, 3]),
    "ycbv_k5": ("ycbv", 5, [3, 6]),
    "tless_top1": ("tless", None, [4, 2, 2]),
}
MODEL_NAME, RUN_ID = "large", "golden"


def random_pose(rng):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = q.astype(np.float32)
    pose[:3, 3] = (rng.normal(size=3) * np.array([80.0, 80.0, 150.0]) + np.array([0.0, 0.0, 900.0])).astype(np.float32)
    return pose


def make_batches(name, seed):
    """Per-batch dicts: images span batch files (the run-time accounting counts every batch of an image once)."""
    dataset, k, sizes = CASES[name]
    rng = np.random.default_rng(seed)
    images = [(2, 3), (2, 7), (48, 1), (48, 12)]                     # (scene_id, im_id)
    det_time = {im: float(rng.uniform(0.05, 0.4)) for im in images}
    batches = []
    cursor = 0
    for n in sizes:
        batch_time = float(rng.uniform(0.02, 0.08))
        ims = [images[(cursor + i) // 3 % len(images)] for i in range(n)]
        cursor += n
        poses = np.stack([np.stack([random_pose(rng) for _ in range(k or 1)]) for _ in range(n)])
        scores = np.sort(rng.integers(0, 257, size=(n, k or 1)), axis=1)[:, ::-1].astype(np.float32) / 256.0
        if k is None:
            poses, scores = poses[:, 0], scores[:, 0]
        batches.append(dict(scene_id=np.array([im[0] for im in ims]), im_id=np.array([im[1] for im in ims]),
                            object_id=rng.integers(1, 9, size=n), time=np.full(n, batch_time),
                            detection_time=np.array([det_time[im] for im in ims]), poses=poses,
                            scores=np.ascontiguousarray(scores)))
    return batches


def main():
    io = ref_import.load_inout()
    for seed, name in enumerate(CASES, start=41):
        dataset, k, _ = CASES[name]
        out_dir = os.path.join(GOLDEN_DIR, name)
        shutil.rmtree(out_dir, ignore_errors=True)
        os.makedirs(out_dir)
        with tempfile.TemporaryDirectory() as tmp:
            for i, b in enumerate(make_batches(name, seed)):
                np.savez(os.path.join(tmp, f"{i}.npz"), **b)
                np.savez(os.path.join(out_dir, f"{i}.npz"), **b)
            io.save_predictions_from_batched_predictions(tmp, dataset_name=dataset, model_name=MODEL_NAME, run_id=RUN_ID,
                                                         is_refined=False)
            for f in sorted(os.listdir(tmp)):
                if f.endswith(".csv"):
                    shutil.copy(os.path.join(tmp, f), os.path.join(out_dir, f))
                    print(name, f, sum(1 for _ in open(os.path.join(tmp, f))), "lines")


if __name__ == "__main__":
    main()
