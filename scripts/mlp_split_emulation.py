"""GPU evidence for keeping row a5 (the per-correspondence IST MLP) on fp32 SIMT.

Runs the c2 feature-level case through the product path (fp32 SIMT MLP), then replaces the two hidden layers by an
emulation of the split-bf16 tensor-core arithmetic (operands rounded to bf16 hi + lo planes, hi*hi + hi*lo + lo*hi, fp32
accumulation -- products are exact in fp32, so this is the tensor-core result up to accumulation order; plain-bf16 too)
and feeds BOTH sets of (relScale, relInplane) to the same RANSAC kernel.  Reports how many of the B*k hypotheses change
their inlier count / inlier set / winning candidate.  python scripts/mlp_split_emulation.py [--out gpurun_out/...json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gigapose_b200 import synth  # noqa: E402
from helpers import engine_from_case  # noqa: E402
from oracle import port  # noqa: E402  (weights container only)


def split(x, dt=torch.bfloat16):
    hi = x.to(dt).float()
    return hi, (x - hi).to(dt).float()


def linear_split(a, w, b, passes, dt=torch.bfloat16):
    ah, al = split(a, dt)
    wh, wl = split(w, dt)
    y = ah @ wh.t()
    if passes == 3:
        y = y + ah @ wl.t() + al @ wh.t()
    return y + b


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "mlp_split_emulation.json"))
    a = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    res = {}
    for name, (B, O, T, seed) in {"c2": (32, 8, 162, 42), "c2_seed2": (32, 8, 162, 43), "t576": (16, 4, 576, 61)}.items():
        case = synth.make_feature_case(B=B, O=O, T=T, seed=seed)
        reg = port.RegressorPort(seed=9).to(dev)
        eng = engine_from_case(case, regressor=reg)
        eng.set_queries(case.q_feat, case.q_mask16.reshape(-1, 16, 16), case.q_label - 1)
        m = eng.sim_topk()
        rs, ri = eng.ist_mlp(case.q_ist, m)
        base = eng.ransac(m, rs, ri)
        # gather the MLP inputs exactly like the kernel: cat(query IST at tar_pt, template IST at src_pt)
        valid = m["src_pts"][..., 0] != -1                                           # [B,k,256]
        b_idx, k_idx, t_idx = torch.nonzero(valid, as_tuple=True)
        q_ist = case.q_ist.to(dev)                                                   # [B,256,16,16]
        bank_ist = case.bank_ist.to(dev)                                             # [O,T,256,16,16]
        tp, sp = m["tar_pts"][b_idx, k_idx, t_idx], m["src_pts"][b_idx, k_idx, t_idx]
        obj = (case.q_label.to(dev) - 1)[b_idx]
        tid = m["id_src"][b_idx, k_idx]
        fq = q_ist[b_idx, :, tp[:, 1], tp[:, 0]]
        ft = bank_ist[obj, tid, :, sp[:, 1], sp[:, 0]]
        x = torch.cat([fq, ft], dim=1)                                               # [rows,512]
        out = {"hypotheses": int(B * 5), "valid_rows": int(x.shape[0])}
        for mode, passes, dt in (("split_bf16_x3", 3, torch.bfloat16), ("split_f16_x3", 3, torch.float16), ("plain_bf16", 1, torch.bfloat16)):
            heads = []
            for head in (reg.scale_predictor, reg.inplane_predictor):
                h = torch.relu(linear_split(x, head[0].weight, head[0].bias, passes, dt))
                h = torch.relu(linear_split(h, head[2].weight, head[2].bias, passes, dt))
                heads.append(h @ head[4].weight.t() + head[4].bias)                  # fp32 head, as the kernel would keep it
            rs2 = torch.full_like(rs, -1000.0)
            ri2 = torch.full_like(ri, -1000.0)
            rs2[b_idx, k_idx, t_idx] = heads[0][:, 0]
            ri2[b_idx, k_idx, t_idx] = torch.tanh(heads[1])
            alt = eng.ransac(m, rs2, ri2)
            d_cnt = (alt["inlier_count"] != base["inlier_count"])
            d_set = (alt["ransac_src_pts"] != base["ransac_src_pts"]).flatten(2).any(-1)
            d_M = ((alt["M"] - base["M"]).abs().flatten(2).max(-1).values > 1e-2)
            out[mode] = {"relScale_err_max": float((rs2 - rs)[valid].abs().max()), "relInplane_err_max": float((ri2 - ri)[valid].abs().max()),
                         "hypotheses_with_other_inlier_count": int(d_cnt.sum()), "hypotheses_with_other_inlier_set": int(d_set.sum()),
                         "hypotheses_with_other_winning_candidate": int(d_M.sum())}
        res[name] = out
        del eng
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
