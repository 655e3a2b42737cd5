import ctypes as C, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gigapose_b200 import _lib, synth
from gigapose_b200.vit import DinoVisionTransformer
from gigapose_b200.vit_engine import NativeViT
lib = _lib.load()
m = DinoVisionTransformer(depth=2, init_seed=1).cuda()
eng = NativeViT(m, "cuda:0", max_crops=32)
rgb, _ = synth.make_crops(32, seed=1, device="cuda")
for _ in range(3): eng.forward(rgb)
arr = (C.c_longlong * 32)()
_lib.check(lib.gp_debug_attention_timeline(arr))
t = list(arr); t0 = t[0]
names = {0:'start',1:'kv landed',2:'S0 issued',3:'P0 ready',4:'PV0 issued',6:'S1 issued',7:'P1 ready',8:'PV1 issued',12:'sm S0 ready',13:'sm max0',14:'sm P0 written',15:'sm O0 ready',16:'sm O0 stored',17:'sm S1 ready',18:'sm max1',19:'sm P1 written',20:'sm O1 ready',21:'sm O1 stored',24:'last-row kv',25:'last-row logits',26:'last-row done'}
for i in sorted(names): print(f"{names[i]:18s} {t[i]-t0:8d}")

g = (C.c_longlong * 64)()
_lib.check(lib.gp_debug_gemm_timeline(g))
g = list(g); g0 = g[63]
print("GEMM (QKV) CTA 0 timeline, cycles since kernel start:")
for t in range(7):
    print(f"  tile {t}: umma start {g[4*t]-g0:8d} issued {g[4*t+1]-g0:8d} | epi start {g[4*t+2]-g0:8d} end {g[4*t+3]-g0:8d}")
