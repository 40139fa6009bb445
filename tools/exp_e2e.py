"""GPU experiment: where the end-to-end (host buffers) step time goes.  Pinned copy peaks, slices sweep, Python overheads."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import llmapigateway_b200 as L
from llmapigateway_b200 import numa
from llmapigateway_b200.synth import sse_batch

bind = "--bind" in sys.argv
info = numa.bind_to_gpu_node(0) if bind else {"bound": False, "gpu_numa_node": numa.gpu_numa_node(0)}
print("numa", info)
dev = torch.device("cuda", 0)
S, E = 4096, 512
b = sse_batch(S, E, 3)
n = int(b.data.size)
pin = lambda a: torch.from_numpy(a).pin_memory()
h_in, h_out = pin(b.data), torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
d2 = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

def h2d():
    with torch.cuda.stream(s1): d.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d2, non_blocking=True)
def both():
    h2d(); d2h()
res = {}
res["h2d_gbs"] = n / timeit(h2d) / 1e9
res["d2h_gbs"] = n / timeit(d2h) / 1e9
res["duplex_ms"] = timeit(both) * 1e3
res["duplex_gbs_each_way"] = n / (res["duplex_ms"] / 1e3) / 1e9
print(json.dumps(res))
eng = L.Engine(max_streams=S, max_step_chunks=S * (E + 2) + 8, max_step_bytes=S * (E * 64 + 512))
status = np.full(S, 200, np.int32)
hh = {"data": h_in.numpy(), "chunk_off": pin(b.chunk_off).numpy(), "seg_chunk": pin(b.seg_chunk).numpy(), "seg_slot": pin(b.seg_slot).numpy(), "out": h_out.numpy()}
for sl in [int(x) for x in os.environ.get('SL', '1,2,4,8,14').split(',')]:
    os.environ["LGW_SLICES"] = str(sl)
    ts = []
    parts = {"open": [], "step": [], "close": []}
    for k in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); eng.open(b.seg_slot, status); t1 = time.perf_counter()
        r = eng.step(hh["data"], hh["chunk_off"], hh["seg_chunk"], hh["seg_slot"], out=hh["out"]); t2 = time.perf_counter()
        st = eng.close(b.seg_slot); torch.cuda.synchronize(); t3 = time.perf_counter()
        if k >= 2:
            ts.append(t3 - t0); parts["open"].append(t1 - t0); parts["step"].append(t2 - t1); parts["close"].append(t3 - t2)
    print("slices", sl, "e2e ms %.3f" % (np.mean(ts) * 1e3), {k: round(float(np.mean(v)) * 1e3, 3) for k, v in parts.items()}, "engine host_step ms", eng.last_step_ms())
