"""Rank process of test_two_rank_gloo_sharded_path (CPU, gloo)."""
import sys
from datetime import datetime, timedelta

import numpy as np
import torch
import torch.distributed as dist

rank, world = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group("gloo", rank=rank, world_size=world)

import host_machine as hm  # noqa: E402
from golden_io import canon_rows, load_sse_cases  # noqa: E402
from llmapigateway_b200 import _abi, _native  # noqa: E402
from llmapigateway_b200.gateway import shard_of  # noqa: E402
from llmapigateway_b200.usage import PERIODS, ROLLUP_CELLS, synth_usage_columns  # noqa: E402
from stream_compare import emitted_from_segments, rows_from_result  # noqa: E402

cases = load_sse_cases()
mine = {}
for i, c in enumerate(cases):
    if shard_of(i, world) != rank:
        continue
    r = hm.run_stream(c["chunks"], [0], c["http_status"])
    st = r["state"]
    rows = None if st.n_exotic else canon_rows(rows_from_result(st, r["rows"]))
    mine[i] = (st.phase == _abi.PHASE_FAILED, emitted_from_segments(c["chunks"], r["step_chunk"], r["segs"]), rows)
gathered = [None] * world
dist.all_gather_object(gathered, mine)

# rollup: per-rank partial dense tables (host arithmetic of the kernel), merged with a SUM all-reduce
lib = _native.load()
n = 20000
ts, models, tok, cost = synth_usage_columns(n, seed=5, end=datetime(2026, 9, 21, 6, 57, 17, 47518))
names = sorted({m for m in models if m is not None}, key=lambda s: s.encode())
rank_of = {m: k + 1 for k, m in enumerate(names)}; rank_of[None] = 0
b = np.array([lib.lgw_rollup_bucket_of(int(t), PERIODS["day"]) for t in ts])
b0, nb, nm = int(b.min()), int(b.max() - b.min() + 1), len(names) + 1


def table_for(idx):
    t = np.zeros((nb, nm, 6), dtype=np.int64)
    for i in idx:
        g = t[b[i] - b0, rank_of[models[i]]]
        for k in range(5):
            g[k] += int(tok[k][i])
        g[5] += 1
    return t


part = torch.from_numpy(table_for([i for i in range(n) if shard_of(i, world) == rank]))
dist.all_reduce(part, op=dist.ReduceOp.SUM)
if rank == 0:
    union = {}
    for g in gathered:
        union.update(g)
    assert sorted(union) == list(range(len(cases)))
    for i, c in enumerate(cases):
        failed, emitted, rows = union[i]
        assert failed == c["failed"] and emitted == c["emitted"], c["name"]
        if rows is not None and not failed:
            assert rows == c["rows"], c["name"]
    assert np.array_equal(part.numpy(), table_for(range(n)))
    print("GLOO_OK")
dist.destroy_process_group()
