"""Where does the per-stream tail cost come from?  C3 variants: with/without usage event and [DONE]."""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import llmapigateway_b200 as L
from llmapigateway_b200.synth import sse_batch
from llmapigateway_b200.engine import SEG_DTYPE
S, E = 4096, 512
eng = L.Engine(max_streams=S, max_step_chunks=S * (E + 2) + 8, max_step_bytes=S * (E * 64 + 512))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, kw, reopen in [("usage+done", {}, True), ("usage only", dict(with_done=False), True), ("done only", dict(with_usage=False), True),
                         ("plain", dict(with_usage=False, with_done=False), True), ("usage+done, committed streams", {}, False), ("plain, committed", dict(with_usage=False, with_done=False), False)]:
    b = sse_batch(n_streams=S, n_events=E, seed=3, **kw)
    d = {k: torch.from_numpy(getattr(b, k)).cuda() for k in ("data", "chunk_off", "seg_chunk", "seg_slot")}
    out = torch.empty_like(d["data"]); segs = torch.empty(S * SEG_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    eng.open(b.seg_slot)
    res = {}
    for timing in (True, False):
        eng.set_kernel_timing(timing)
        ms = []
        for it in range(5):
            if reopen: eng.open(b.seg_slot)
            flush.fill_(it); torch.cuda.synchronize()
            eng.step_device(d["data"].data_ptr(), int(b.data.size), d["chunk_off"].data_ptr(), b.n_chunks, d["seg_chunk"].data_ptr(), d["seg_slot"].data_ptr(), S, out.data_ptr(), segs.data_ptr())
            eng.sync(); ms.append(eng.last_step_ms())
        res[timing] = ms
    parts = {k: round(v * 1e3, 1) for k, v in res[True][-1].items() if k != "host_step"}
    print(f"{name:32s} back-to-back {parts} sum {round(sum(parts.values()), 1)} | chained (PDL) step {[round(m['relay'] * 1e3, 1) for m in res[False][1:]]} us")
