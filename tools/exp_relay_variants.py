import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import llmapigateway_b200 as L
from llmapigateway_b200.synth import sse_batch
from llmapigateway_b200.engine import SEG_DTYPE
S,E=4096,512
eng=L.Engine(max_streams=S,max_step_chunks=S*(E+2)+8,max_step_bytes=S*(E*64+512))
for name,kw in [("full",{}),("no_usage_done",dict(with_usage=False,with_done=False)),("8_per_chunk",dict(events_per_chunk=8))]:
    b=sse_batch(n_streams=S,n_events=E,seed=3,**kw)
    d={k:torch.from_numpy(getattr(b,k)).cuda() for k in ("data","chunk_off","seg_chunk","seg_slot")}
    out=torch.empty_like(d["data"]); segs=torch.empty(S*SEG_DTYPE.itemsize,dtype=torch.uint8,device="cuda")
    ms=[]
    for it in range(6):
        eng.open(b.seg_slot)
        eng.step_device(d["data"].data_ptr(),int(b.data.size),d["chunk_off"].data_ptr(),b.n_chunks,d["seg_chunk"].data_ptr(),d["seg_slot"].data_ptr(),S,out.data_ptr(),segs.data_ptr())
        eng.sync(); ms.append(eng.last_step_ms())
    print(name, {k:round(v,4) for k,v in ms[-1].items()}, "equal", bool(torch.equal(out,d["data"])), "relay per iteration", [round(m["relay"],3) for m in ms])
    import ctypes as C
    st=(C.c_uint32*8)(); t0=C.create_string_buffer(64); t1=C.create_string_buffer(64)
    if hasattr(eng._lib,"lgw_debug_template_cache"):
        eng._lib.lgw_debug_template_cache(C.c_void_p(eng._h.value if hasattr(eng._h,"value") else eng._h), st, t0, t1, 64)
        print("   template cache: state", list(st)[:2], "len", list(st)[2:4], "slot0", t0.raw[:40], "slot1", t1.raw[:40])
