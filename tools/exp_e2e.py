import sys, os, time; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import llmapigateway_b200 as L
from llmapigateway_b200.synth import sse_batch
S,E=4096,512
b=sse_batch(n_streams=S,n_events=E,seed=3)
pin=lambda a: torch.from_numpy(a).pin_memory().numpy()
data,co,sc,ss=pin(b.data),pin(b.chunk_off),pin(b.seg_chunk),pin(b.seg_slot)
out=torch.empty(b.data.size,dtype=torch.uint8).pin_memory().numpy()
for k in (1,2,3,4,6,8):
    os.environ["LGW_SLICES"]=str(k)
    eng=L.Engine(max_streams=S,max_step_chunks=S*(E+2)+8,max_step_bytes=S*(E*64+512))
    ts=[]
    for it in range(6):
        eng.open(b.seg_slot)
        t0=time.perf_counter(); r=eng.step(data,co,sc,ss,out=out); t1=time.perf_counter()
        st=eng.close(b.seg_slot); t2=time.perf_counter()
        ts.append(((t1-t0)*1e3,(t2-t1)*1e3, eng.last_step_ms()["host_step"]))
    print("slices",k,"step ms %.3f close ms %.3f dev ms %.3f"%ts[-1], "ok", bool((r.out==b.data).all()), {a:round(v,3) for a,v in eng.last_step_ms().items()})
    eng.close_engine()
