"""k_relay on realistic OpenAI-shaped streams (not the 64-byte BASELINE deltas): python tools/exp_openai_shape.py"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import llmapigateway_b200 as L  # noqa: E402
from llmapigateway_b200.engine import SEG_DTYPE  # noqa: E402
from llmapigateway_b200.synth import openai_batch  # noqa: E402

S, D = 4096, 128
b = openai_batch(n_streams=S, n_deltas=D, seed=9)
eng = L.Engine(max_streams=S, max_step_chunks=b.n_chunks + 8, max_step_bytes=int(b.data.size) + 4096)
d = {k: torch.from_numpy(getattr(b, k)).cuda() for k in ("data", "chunk_off", "seg_chunk", "seg_slot")}
out = torch.empty_like(d["data"]); segs = torch.empty(S * SEG_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
ms = []
for it in range(6):
    eng.open(b.seg_slot)
    eng.step_device(d["data"].data_ptr(), int(b.data.size), d["chunk_off"].data_ptr(), b.n_chunks, d["seg_chunk"].data_ptr(), d["seg_slot"].data_ptr(), S, out.data_ptr(), segs.data_ptr())
    eng.sync(); ms.append(eng.last_step_ms())
m = ms[-1]; tot = m["prime"] + m["relay"] + m["commit"]
print(f"streams={S} chunks={b.n_chunks} bytes={b.data.size} mean_chunk={b.data.size / b.n_chunks:.1f} B  ms={ {k: round(v, 4) for k, v in m.items()} }")
print(f"   chunks/s={b.n_chunks / tot * 1e3:.3e}  JSON GB/s={b.data.size / tot / 1e6:.1f}  (in+out)/HBM={2 * b.data.size / tot / 1e6 / 6569.6 * 100:.1f} %  equal={bool(torch.equal(out, d['data']))}")
