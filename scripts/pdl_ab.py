"""Same-process A/B of programmatic dependent launch (GIGAPOSE_PDL is read at every launch): ViT forward over 32 crops,
eager and as a CUDA graph, alternating the setting so that clock drift hits both arms alike."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gigapose_b200 import synth  # noqa: E402
from gigapose_b200.vit import DinoVisionTransformer  # noqa: E402
from gigapose_b200.vit_engine import NativeViT  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    vit = DinoVisionTransformer(init_seed=7).to(dev)
    eng = NativeViT(vit, dev, max_crops=32)
    x = synth.make_crops(32, seed=1, device=dev)[0]
    res = {}
    for rep in range(3):
        for pdl in ("0", "1"):
            os.environ["GIGAPOSE_PDL"] = pdl
            for _ in range(3):
                eng.forward(x)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(10):
                eng.forward(x)
            ev[1].record()
            torch.cuda.synchronize()
            eager = ev[0].elapsed_time(ev[1]) / 10
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                eng.forward(x)
            torch.cuda.current_stream().wait_stream(side)
            with torch.cuda.graph(g):
                eng.forward(x)
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(10):
                g.replay()
            ev[1].record()
            torch.cuda.synchronize()
            graph = ev[0].elapsed_time(ev[1]) / 10
            res.setdefault(pdl, []).append((round(eager, 3), round(graph, 3)))
            print(f"rep {rep} PDL={pdl}: eager {eager:.3f} ms, graph {graph:.3f} ms per 32-crop ViT forward", flush=True)
    print(res)


if __name__ == "__main__":
    main()
