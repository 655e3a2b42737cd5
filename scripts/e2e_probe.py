"""Diagnostic: where does the end-to-end step (pinned host batch -> poses on the host) spend its time?
Times the H2D copy alone, the resident step, and the pipelined stage / retrieve / fetch_async loop with and without the
CUDA graph.  python scripts/e2e_probe.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    cfg = bench.WORKLOADS["c2"]
    model = bench.build_models(dev)
    templates = bench.SyntheticTemplates(cfg["O"], cfg["T"], dev)
    model.template_datasets = {"synthetic": templates}
    model.test_dataset_name = "synthetic"
    model.set_template_data("synthetic")
    batch_host, labels, views = bench.make_queries(templates, cfg["B"])
    batch_dev = batch_host.clone().to(dev)
    print("pinned:", batch_host.tar_img.is_pinned(), batch_host.tar_mask.is_pinned())
    # H2D alone
    torch.cuda.synchronize()
    for name in ("tar_img", "tar_mask"):
        t = getattr(batch_host, name)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t.to(dev, non_blocking=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            t.to(dev, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"H2D {name}: {t.numel() * 4 / 1e6:.1f} MB in {ms:.3f} ms = {t.numel() * 4 / 1e6 / ms:.1f} GB/s")
    for graph in (False, True):
        model.use_cuda_graph = graph
        for _ in range(3):
            model.retrieve(batch_dev, "synthetic")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            model.retrieve(batch_dev, "synthetic")
        torch.cuda.synchronize()
        print(f"graph={graph}: resident step {(time.perf_counter() - t0) * 200:.2f} ms (wall)")
        # unpipelined e2e
        for _ in range(2):
            p = model.retrieve(batch_host, "synthetic")
            p.pred_poses.cpu()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            p = model.retrieve(batch_host, "synthetic")
            p.pred_poses.cpu(); p.scores.cpu()
        torch.cuda.synchronize()
        print(f"graph={graph}: unpipelined e2e step {(time.perf_counter() - t0) * 200:.2f} ms")
        # pipelined e2e, with per-phase host timing
        steps = 6
        tt = {"stage": 0.0, "retrieve": 0.0, "fetch": 0.0, "result": 0.0}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a = time.perf_counter()
        staged = model.stage(batch_host, "synthetic")
        tt["stage"] += time.perf_counter() - a
        pending = None
        for i in range(steps):
            cur = staged
            if i + 1 < steps:
                a = time.perf_counter(); staged = model.stage(batch_host, "synthetic"); tt["stage"] += time.perf_counter() - a
            a = time.perf_counter(); pred = model.retrieve(cur, "synthetic"); tt["retrieve"] += time.perf_counter() - a
            a = time.perf_counter(); handle = model.fetch_async(pred); tt["fetch"] += time.perf_counter() - a
            if pending is not None:
                a = time.perf_counter(); pending.result(); tt["result"] += time.perf_counter() - a
            pending = handle
        a = time.perf_counter(); pending.result(); tt["result"] += time.perf_counter() - a
        torch.cuda.synchronize()
        total = (time.perf_counter() - t0) * 1e3 / steps
        print(f"graph={graph}: pipelined e2e step {total:.2f} ms; host time per step (ms): "
              + ", ".join(f"{k} {v * 1e3 / steps:.2f}" for k, v in tt.items()))


if __name__ == "__main__":
    main()
