"""One C3 step a few times (for ncu captures and quick timings)."""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import llmapigateway_b200 as L
from llmapigateway_b200.synth import sse_batch
from llmapigateway_b200.engine import SEG_DTYPE
S, E = 4096, 512
kw = dict(with_usage=False, with_done=False) if "plain" in sys.argv else {}
eng = L.Engine(max_streams=S, max_step_chunks=S * (E + 2) + 8, max_step_bytes=S * (E * 64 + 512))
b = sse_batch(n_streams=S, n_events=E, seed=3, **kw)
d = {k: torch.from_numpy(getattr(b, k)).cuda() for k in ("data", "chunk_off", "seg_chunk", "seg_slot")}
out = torch.empty_like(d["data"]); segs = torch.empty(S * SEG_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
for it in range(4):
    eng.open(b.seg_slot)
    eng.step_device(d["data"].data_ptr(), int(b.data.size), d["chunk_off"].data_ptr(), b.n_chunks, d["seg_chunk"].data_ptr(), d["seg_slot"].data_ptr(), S, out.data_ptr(), segs.data_ptr())
    eng.sync()
    print({k: round(v * 1e3, 1) for k, v in eng.last_step_ms().items()}, eng.debug_counters())
