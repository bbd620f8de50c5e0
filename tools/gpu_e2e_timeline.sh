#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' | tee gpurun_out/e2e_timeline.txt
import time, numpy as np, torch
from neuralplane_b200 import ControlEnv, GPUVecEnv
n=1_000_000
for pat, late_small in ((4, False), (4, True), ((1,2,3,4,4,4), True)):
    v=GPUVecEnv([lambda: ControlEnv(num_envs=n, config="heading", model="F16", random_seed=0, device="cuda:0")], pipeline_chunks=pat)
    v.reset()
    e=v.gpu_vec_env
    acts=[(np.random.rand(n,1,4).astype(np.float32)*2-1) for _ in range(2)]
    for i in range(5): v.step(acts[i%2])
    torch.cuda.synchronize()
    def step(a, log):
        a=np.asarray(a,dtype=np.float32).reshape(n,-1)
        o=v._out[v._flip]; v._flip^=1
        main=torch.cuda.current_stream(e.device)
        up,run,down=v._streams
        t0=time.perf_counter()
        start=torch.cuda.Event(enable_timing=True); start.record(main)
        for st in v._streams: st.wait_event(start)
        evs=[]
        E=lambda: torch.cuda.Event(enable_timing=True)
        for c,(i0,i1) in enumerate(v._chunks):
            v._act_h[i0:i1].copy_(torch.from_numpy(a[i0:i1,:4]))
            ev=[E() for _ in range(4)]
            with torch.cuda.stream(up):
                v._act_d[i0:i1].copy_(v._act_h[i0:i1],non_blocking=True); ev[0].record(up); run.wait_event(ev[0])
            with torch.cuda.stream(run):
                e.step_range(v._act_d,i0,i1-i0,advance=(c==0)); ev[1].record(run); down.wait_event(ev[1])
            with torch.cuda.stream(down):
                o["obs"][i0:i1].copy_(e.last_obs[i0:i1],non_blocking=True); ev[2].record(down)
                if not late_small:
                    o["rew"][i0:i1].copy_(e.last_reward[i0:i1],non_blocking=True)
                    for j in range(3): o["flags"][j,i0:i1].copy_(e._flags[j,i0:i1],non_blocking=True)
                ev[3].record(down)
            evs.append((ev, time.perf_counter()-t0))
        if late_small:
            with torch.cuda.stream(down):
                o["rew"].copy_(e.last_reward,non_blocking=True)
                o["flags"].copy_(e._flags[:, :n],non_blocking=True)
        t_issue=time.perf_counter()-t0
        main.wait_stream(down)
        main.synchronize()
        t_done=time.perf_counter()-t0
        if log:
            for c,(ev,tc) in enumerate(evs):
                print("chunk %d: cpu issued at %.3f ms | gpu: h2d done %.3f kernel done %.3f obs d2h done %.3f small d2h done %.3f"%(c,tc*1e3,*[start.elapsed_time(x) for x in ev]))
            print("cpu: all issued %.3f ms, done %.3f ms"%(t_issue*1e3,t_done*1e3))
        return t_done
    ts=[step(acts[i%2], False) for i in range(20)]
    print("pattern", pat, "late_small", late_small, "mean %.3f ms  min %.3f ms"%(np.mean(ts)*1e3, np.min(ts)*1e3))
    step(acts[0], True)
    del v
PY
