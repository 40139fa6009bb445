"""C3 step: chained total and per-kernel times (for build variants of the commit kernel)."""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import llmapigateway_b200 as L
from llmapigateway_b200.synth import sse_batch
from llmapigateway_b200.engine import SEG_DTYPE
S, E = 4096, 512
eng = L.Engine(max_streams=S, max_step_chunks=S * (E + 2) + 8, max_step_bytes=S * (E * 64 + 512))
b = sse_batch(n_streams=S, n_events=E, seed=3)
d = {k: torch.from_numpy(getattr(b, k)).cuda() for k in ("data", "chunk_off", "seg_chunk", "seg_slot")}
out = torch.empty_like(d["data"]); segs = torch.empty(S * SEG_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for timing in (False, True):
    eng.set_kernel_timing(timing)
    ms = []
    for it in range(8):
        eng.open(b.seg_slot)
        flush.fill_(it); torch.cuda.synchronize()
        eng.step_device(d["data"].data_ptr(), int(b.data.size), d["chunk_off"].data_ptr(), b.n_chunks, d["seg_chunk"].data_ptr(), d["seg_slot"].data_ptr(), S, out.data_ptr(), segs.data_ptr())
        eng.sync(); ms.append(eng.last_step_ms())
    ms = ms[3:]
    res["per_kernel" if timing else "chained"] = {k: round(float(np.mean([m[k] for m in ms])) * 1e3, 1) for k in ("prime", "relay", "commit")}
st = eng.state(b.seg_slot[:3])
from llmapigateway_b200 import _abi
ok = all(_abi.usage_rec_to_dict(st[s].rec) == b.truths[s].expected_row() for s in range(3)) and bool(torch.equal(out, d["data"]))
print(res, "correct", ok)
