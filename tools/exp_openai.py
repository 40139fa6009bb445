"""A few steps over the OpenAI-shaped probe (for ncu captures and timings)."""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import llmapigateway_b200 as L
from llmapigateway_b200.synth import openai_batch
from llmapigateway_b200.engine import SEG_DTYPE
S = 4096
b = openai_batch(n_streams=S, n_deltas=128, seed=5)
eng = L.Engine(max_streams=S, max_step_chunks=b.n_chunks + 8, max_step_bytes=int(b.data.size) + 4096)
d = {k: torch.from_numpy(getattr(b, k)).cuda() for k in ("data", "chunk_off", "seg_chunk", "seg_slot")}
out = torch.empty_like(d["data"]); segs = torch.empty(S * SEG_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
for it in range(5):
    eng.open(b.seg_slot)
    eng.step_device(d["data"].data_ptr(), int(b.data.size), d["chunk_off"].data_ptr(), b.n_chunks, d["seg_chunk"].data_ptr(), d["seg_slot"].data_ptr(), S, out.data_ptr(), segs.data_ptr())
    eng.sync()
    print(int(b.data.size), b.n_chunks, {k: round(v * 1e3, 1) for k, v in eng.last_step_ms().items()}, eng.debug_counters(), bool(torch.equal(out, d["data"])))
