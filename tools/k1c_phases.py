#!/usr/bin/env python
"""Phase timeline of one K1c step (a -DNPLANE_COOP_TIMING build named by NPLANE_LIB): per-warp clock64 stamps of CTA 0, in SM
cycles since kernel start.  0 prologue done | 1 pass-0 MLP share (+ owners' pre-tail) done | 2 barrier passed | 3 pass-0 tail
done | 4 exchange barrier passed | 5 pass-1 MLP share done | 8 barrier passed | 6 tail 1 + stores done | 7 final barrier | 9 end."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralplane_b200 import ControlEnv  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
env = ControlEnv(num_envs=n, config="heading", model="F16", random_seed=0, device="cuda:0")
env.reset()
a = torch.rand((n, 4), device="cuda") * 2 - 1
for k in range(20):
    env.step(a)
torch.cuda.synchronize()
nw = env.launch_info()["block"] // 32
t = env.last_obs.flatten()[:10 * nw].cpu().view(nw, 10).long()
order = [0, 1, 2, 3, 4, 5, 8, 6, 7, 9]
names = ["prologue", "mlp0", "bar0", "tail0", "xbar", "mlp1", "bar1", "tail1+st", "endbar", "end"]
print("warp " + " ".join(f"{x:>9}" for x in names))
for w in range(nw):
    print(f"  {w}  " + " ".join(f"{int(t[w, i]):9d}" for i in order))
