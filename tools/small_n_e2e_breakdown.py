#!/usr/bin/env python
"""Where the end-to-end time of a small-population step through the numpy boundary goes (n = 3 000, mapped boundary):
GPUVecEnv.step -> BaseEnv.step_mapped -> the bare ctypes call -> the kernel alone (CUDA events), plus the floor of
'launch one empty kernel and synchronise' on this box."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralplane_b200 import ControlEnv, GPUVecEnv  # noqa: E402
from neuralplane_b200 import _native as nv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
K = 2000
vec = GPUVecEnv([lambda: ControlEnv(num_envs=n, config="heading", model="F16", random_seed=0, device="cuda:0")])
vec.reset()
e = vec.gpu_vec_env
a3 = (np.random.default_rng(0).random((n, 1, 4), dtype=np.float32) * 2 - 1)
a2 = np.ascontiguousarray(a3.reshape(n, 4))
out = {"n": n, "boundary": vec.boundary}


def timeit(f, k=K):
    for _ in range(50):
        f()
    t = time.perf_counter()
    for _ in range(k):
        f()
    return round((time.perf_counter() - t) / k * 1e6, 2)


out["vec_step_us"] = timeit(lambda: vec.step(a3))
o = vec._out[0]
out["step_mapped_us"] = timeit(lambda: e.step_mapped(a2, vec._act_h, o["obs"], o["rew"], o["flags"], None))
lib = nv.lib()
args = (e._handle, a2.ctypes.data, vec._act_h.data_ptr(), None, o["obs"].data_ptr(), o["rew"].data_ptr(), o["flags"].data_ptr(), e.ld,
        torch.cuda.current_stream().cuda_stream)
out["ctypes_call_us"] = timeit(lambda: lib.np_env_step_mapped(*args))
# the kernel alone, writing to host memory / to device memory
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ts = []
for _ in range(200):
    ev[0].record(); lib.np_env_step_mapped(*args); ev[1].record(); torch.cuda.synchronize(); ts.append(ev[0].elapsed_time(ev[1]) * 1e3)
out["kernel_mapped_us_events"] = round(float(np.median(ts)), 2)
ad = torch.from_numpy(a2).cuda()
ts = []
for _ in range(200):
    ev[0].record(); e.step(ad); ev[1].record(); torch.cuda.synchronize(); ts.append(ev[0].elapsed_time(ev[1]) * 1e3)
out["kernel_device_us_events"] = round(float(np.median(ts)), 2)
out["device_step_plus_sync_us"] = timeit(lambda: (e.step(ad), torch.cuda.synchronize()))
x = torch.zeros(32, device="cuda")
out["empty_launch_plus_sync_us"] = timeit(lambda: (x.add_(1), torch.cuda.synchronize()))
out["python_noop_ctypes_us"] = timeit(lambda: lib.np_version())
print(json.dumps(out))
