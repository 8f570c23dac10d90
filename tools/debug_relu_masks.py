"""Not a test: relu activation patterns of the CUDA forward vs the fp64 oracle (a unit whose pre-activation is ~0 may
land on the other side of the kink in fp32: its gradient contribution then differs by construction)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import set_oracle as O
from sgrl_b200 import graph as G, morphologies as M, synth
import gpu_util

B = 100
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 51
actor, critic, pa, pc = gpu_util.make_modules(use_tc=int(os.environ.get("USE_TC", "1")))
par = M.ALL["3d_humanoid_9_full"]; N = len(par)
g = G.build_graph(par, device="cuda")
g64 = dict(g); g64["relation"] = g["relation"].double()
critic.change_morphology(g)
b = {k: v.cuda() for k, v in synth.make_batch(B, N, seed=seed).items()}
tb = critic._tables(B)
out, stash = critic.forward_raw(tb, b["obs"].contiguous(), b["action"].contiguous(), keep=True)
torch.cuda.synchronize()
x = torch.cat([b["obs"].view(B, N, 41), b["action"].view(B, N, 3)], 2)
for z, prefix in ((0, "critic1."), (1, "critic2.")):
    trace = {}
    p = {k: v.cuda().double() for k, v in O.sub(pc, prefix).items()}
    with torch.no_grad():
        O.transformer_model(p, x.double(), g64, trace=trace)
    for key, ref in trace.items():
        l, name = (int(key.split(".")[0]), key.split(".")[1]) if "." in key else (-1, key)
        if name not in ("A1", "A2", "T31", "AH", "BH"):
            continue
        got = gpu_util.stash_view(critic, stash, tb, 2, z, name, l)
        ref = ref.reshape(tb.T, -1)
        mism = torch.nonzero((got > 0) != (ref.to(got.device) > 0))
        for t, c in mism.tolist()[:5]:
            print(f"{prefix}{key}: token {t} (sample {t // N}) unit {c}: ours {got[t, c].item():.3e} oracle {ref[t, c].item():.3e}; row scale {ref[t].abs().max().item():.2e}")
print("done")
