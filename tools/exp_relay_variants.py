"""Perf probe of the bulk path on a B200: per-kernel device times for C3 and its variants, template cache, counters."""
import sys; sys.path.insert(0, '/root/repo')
import ctypes as C
import numpy as np, torch
import llmapigateway_b200 as L
from llmapigateway_b200.synth import sse_batch, openai_batch
from llmapigateway_b200.engine import SEG_DTYPE
S, E = 4096, 512
eng = L.Engine(max_streams=S, max_step_chunks=S * (E + 2) + 8, max_step_bytes=S * (E * 64 + 512))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
cases = [("full", lambda: sse_batch(n_streams=S, n_events=E, seed=3)),
         ("no_usage_done", lambda: sse_batch(n_streams=S, n_events=E, seed=3, with_usage=False, with_done=False)),
         ("8_per_chunk", lambda: sse_batch(n_streams=S, n_events=E, seed=3, events_per_chunk=8)),
         ("openai_shaped", lambda: openai_batch(n_streams=S, n_deltas=128, seed=5))]
for name, mk in cases:
    b = mk()
    d = {k: torch.from_numpy(getattr(b, k)).cuda() for k in ("data", "chunk_off", "seg_chunk", "seg_slot")}
    out = torch.empty_like(d["data"]); segs = torch.empty(S * SEG_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    ms = []
    for it in range(6):
        eng.open(b.seg_slot)
        flush.fill_(it); torch.cuda.synchronize()
        eng.step_device(d["data"].data_ptr(), int(b.data.size), d["chunk_off"].data_ptr(), b.n_chunks, d["seg_chunk"].data_ptr(), d["seg_slot"].data_ptr(), S, out.data_ptr(), segs.data_ptr())
        eng.sync(); ms.append(eng.last_step_ms())
    tot = sum(v for k, v in ms[-1].items() if k != "host_step")
    print(name, "bytes", int(b.data.size), {k: round(v * 1e3, 1) for k, v in ms[-1].items() if k != "host_step"}, "us; sum", round(tot * 1e3, 1), "us;",
          round(2 * b.data.size / tot / 1e6, 0), "GB/s in+out; equal", bool(torch.equal(out, d["data"])), "relay per iteration", [round(m["relay"] * 1e3, 1) for m in ms])
    st = (C.c_uint32 * 16)()
    eng._lib.lgw_debug_template_cache(eng._h, st, None, 0)
    v = list(st)
    print("   templates: state", v[0:4], "len", v[4:8], "usage_ok", [x >> 31 for x in v[8:12]], "hits", v[12:16], "counters", eng.debug_counters())
