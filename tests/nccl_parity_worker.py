"""Worker of tests/test_gpu_nccl.py (launched with torch.distributed.run, one rank per GPU): the sharded pipeline
(`ShardedRetriever`: descriptor shards, in-library NCCL all-gathers, query-sharded tail) must reproduce the single-GPU
`GigaPose.retrieve` result bit for bit, including an uneven split of the batch and of the templates."""
import json
import os
import sys
from datetime import timedelta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device, timeout=timedelta(seconds=120))
    import bench
    from gigapose_b200.multigpu import ShardedRetriever, window
    O, T, B = 2, 9, 5                                     # 9 templates over 2 ranks: 5 + 4; 5 detections: 3 + 2
    model = bench.build_models(device)
    templates = bench.SyntheticTemplates(O, T, device)
    batch, labels, views = bench.make_queries(templates, B, seed=11)
    retr = ShardedRetriever(model, templates, rank, world, device, max_batch=B)
    lo, hi = window(B, rank, world)
    dev = lambda t: t.to(device)
    names = ("id_src", "scores", "pred_poses", "src_pts", "tar_pts", "ransac_scores", "relScale", "M")
    for _ in range(2):                                    # twice: buffers are reused across steps
        out = retr.retrieve(dev(batch.tar_img[lo:hi]), dev(batch.tar_mask), (labels - 1).to(device), dev(batch.tar_K[lo:hi]),
                            dev(batch.tar_M[lo:hi]))
    full = retr.gather_results(out, B, names=names)
    # pipelined form (stage / retrieve_staged / fetch_async) gives the same window results
    staged = retr.stage(batch)
    o2 = retr.retrieve_staged(staged)
    poses_host, scores_host = retr.fetch_async(o2).result()
    ok_pipe = bool(torch.equal(poses_host, out["pred_poses"].cpu()) and torch.equal(scores_host, out["scores"].cpu()))
    torch.cuda.synchronize()
    result = {"rank": rank, "pipelined_equal": ok_pipe}
    if rank == 0:
        model.template_datasets = {"synthetic": templates}
        model.test_dataset_name = "synthetic"
        single = model.retrieve(batch, "synthetic")
        diffs = {n: int((getattr(single, n) != full[n]).sum()) for n in names}
        result.update(diffs=diffs, planted=float((full["id_src"].cpu() == views[:, None]).any(dim=1).float().mean()))
    gathered = [None] * world
    dist.all_gather_object(gathered, result)
    if rank == 0:
        print("NCCL_PARITY " + json.dumps(gathered))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
