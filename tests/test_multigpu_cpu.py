"""world_size-2 gloo test (CPU) of the multi-GPU host logic: template-interleaved shard maps, candidate record
packing, the single all-gather and the merge ordering.  The similarity search of each shard is done by the CPU
oracle here (no GPU); on the GPU box the same host code drives the CUDA kernels (tests/test_gpu_multi.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gigapose_b200 import multigpu, synth
from oracle import port


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port_no, T, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    case = synth.make_feature_case(B=3, O=2, T=T, seed=17)
    ri = synth.to_reference_layout(case)
    k = 5
    ids = multigpu.shard_template_ids(T, rank, world)
    sim = port.similarity_search(ri["src_feats"][:, ids], ri["tar_feat"], ri["src_masks"][:, ids], ri["tar_mask"], k=k)
    B = 3
    flat, total = multigpu.alloc_packed(B, k, "cpu")
    mine = multigpu.field_views(flat, B, k, total)
    mine["score"].copy_(sim["score_src"])
    mine["id"].copy_(torch.tensor([[multigpu.local_to_global(int(j), rank, world) for j in row] for row in sim["id_src"]]))
    mine["pts_score"].copy_(sim["score_pts"])
    valid = sim["src_pts"][..., 0] != -1
    mine["valid"].copy_(valid.to(torch.uint8))
    mine["idx"].copy_(torch.where(valid, sim["src_pts"][..., 1] * 16 + sim["src_pts"][..., 0], 0).to(torch.uint8))
    gathered = multigpu.all_gather_packed(flat, world)
    views = [multigpu.field_views(gathered, B, k, total, rank_slot=g) for g in range(world)]
    gid, score, order = multigpu.merge_reference(views, k)
    if rank == 0:
        full = port.similarity_search(ri["src_feats"], ri["tar_feat"], ri["src_masks"], ri["tar_mask"], k=k)
        ret["ids_equal"] = bool(torch.equal(gid, full["id_src"]))
        ret["score_equal"] = bool(torch.allclose(score, full["score_src"], atol=1e-6))
        # the records of the winners carry the right correspondences
        pts = torch.cat([v["idx"] for v in views], dim=1)
        val = torch.cat([v["valid"] for v in views], dim=1)
        win_idx = torch.gather(pts, 1, order[..., None].expand(-1, -1, 256)).long()
        win_val = torch.gather(val, 1, order[..., None].expand(-1, -1, 256)).bool()
        sx = torch.where(win_val, win_idx % 16, -1)
        ret["pts_equal"] = bool(torch.equal(sx, full["src_pts"][..., 0]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("T", [12, 13])          # even and uneven shards
def test_sharded_candidates_merge_to_the_single_gpu_result(T):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), T, ret), nprocs=world, join=True)
    assert ret["ids_equal"] and ret["score_equal"] and ret["pts_equal"], dict(ret)


def test_shard_maps_and_record_layout():
    for T, G in [(162, 4), (162, 8), (576, 8), (16, 1)]:
        seen = sorted(i for r in range(G) for i in multigpu.shard_template_ids(T, r, G))
        assert seen == list(range(T))
        for r in range(G):
            ids = multigpu.shard_template_ids(T, r, G)
            assert [multigpu.local_to_global(j, r, G) for j in range(len(ids))] == ids
            assert abs(len(ids) - T / G) < 1
    lay, total = multigpu.record_layout(128, 5)
    assert total % 16 == 0 and all(off % 16 == 0 for off, _, _ in lay.values())
    # 4616 B per record (SURVEY.md §8e estimated 0.4-1.8 KB without the IST outputs)
    assert total >= 128 * 5 * 4616
    flat, total = multigpu.alloc_packed(2, 5, "cpu", world=3)
    v = multigpu.field_views(flat, 2, 5, total, rank_slot=2)
    v["id"].fill_(7)
    assert int(flat.view(torch.int32).sum()) == 7 * 10
