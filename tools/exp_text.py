"""Transcript tap pass over the C3 batch: device time of k_text_extract + k_text_scan + k_text_pack."""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import llmapigateway_b200 as L
from llmapigateway_b200.synth import sse_batch
from llmapigateway_b200.engine import SEG_DTYPE
S, E = 4096, 512
eng = L.Engine(max_streams=S, max_step_chunks=S * (E + 2) + 8, max_step_bytes=S * (E * 64 + 512))
eng.enable_transcripts()
b = sse_batch(n_streams=S, n_events=E, seed=3)
d = {k: torch.from_numpy(getattr(b, k)).cuda() for k in ("data", "chunk_off", "seg_chunk", "seg_slot")}
out = torch.empty_like(d["data"]); segs = torch.empty(S * SEG_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
ms = []
for it in range(5):
    eng.open(b.seg_slot)
    eng.step_device(d["data"].data_ptr(), int(b.data.size), d["chunk_off"].data_ptr(), b.n_chunks, d["seg_chunk"].data_ptr(), d["seg_slot"].data_ptr(), S, out.data_ptr(), segs.data_ptr())
    eng.sync()
    st = eng.step_transcript()
    ms.append(eng.transcript_last_ms())
raw = b.data.tobytes()[:E * 64]
want = b"".join(raw[i * 64 + 49:i * 64 + 57] for i in range(E))
print("transcript pass ms", [round(m, 3) for m in ms], "correct", st.segment(0) == want and int(st.seg_off[-1]) == S * E * 8)
