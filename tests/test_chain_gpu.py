"""GPU: config 4 end to end -- the fallback-chain walker over the real engine, against the goldens of the unmodified
`chat_completions` (chat.py:20-198) and, at BASELINE.json's full size (8 192 streams, 3-rule chain, 20 % injected attempt
failure), against the oracle walk plus size-independent properties."""
import numpy as np
import pytest

import chain_cases as cc
from llmapigateway_b200 import chat, rewrite, synth
from llmapigateway_b200.gateway import shard_of
from oracle import chain_oracle

pytestmark = pytest.mark.gpu


def test_chain_walker_goldens_on_the_engine():
    import llmapigateway_b200 as L
    doc, ups = cc.load()
    engines = []

    def factory():
        engines.append(L.Engine(max_streams=64, max_step_chunks=4096, max_step_bytes=1 << 22))
        return engines[-1]

    got = cc.walk_product(factory, doc["cases"], ups)
    for case, g in zip(doc["cases"], got):
        cc.check_against_golden(case, g)
    for e in engines:
        e.close_engine()


@pytest.mark.parametrize("n_streams,n_events,n_gpus", [(512, 64, 1), (8192, 512, 8)])
def test_chain_batch_full_size(n_streams, n_events, n_gpus):
    """One GPU's shard (stream -> GPU by shard_of) of config 4 through ChainBatch: every served stream relays exactly
    the bytes its serving attempt streamed, every 503 detail equals the oracle's, the attempt count equals the injected schedule's,
    and a sample of streams (all kinds) is replayed through the oracle walk byte for byte."""
    import llmapigateway_b200 as L
    providers, rules, fallback_provider = synth.chain_world()
    up = synth.ChainUpstream(n_streams, n_events, seed=4, p_fail=0.2)
    all_bodies = synth.chain_request_bodies(n_streams, seed=4)
    mine = np.array([i for i in range(n_streams) if shard_of(i, n_gpus) == 0])
    bodies = [all_bodies[i] for i in mine]
    n = len(mine)
    eng = L.Engine(max_streams=n, max_step_chunks=n * (n_events + 2) + 16, max_step_bytes=n * (n_events * 64 + 512) + 4096)
    plans = rewrite.RulePlans(rules, fallback_provider=fallback_provider, stream_mode=cc.stream_mode())
    eng.load_rules(plans)
    out = chat.ChainBatch(eng, plans, providers, rules).run(bodies, None, up, stream_ids=mine)
    kind = up.kind[:3, mine]                                            # [attempt, stream]
    first_ok = np.where((kind == 0).any(axis=0), (kind == 0).argmax(axis=0), -1)
    assert np.array_equal(out.served_round, first_ok.astype(np.int32))
    assert out.attempts == int(np.where(first_ok >= 0, first_ok + 1, 3).sum())
    b = up.batch
    co = b.chunk_off.astype(np.int64)
    for k in range(n):                                                  # relayed bytes == what the serving attempt streamed
        if first_ok[k] >= 0:
            s = int(mine[k])
            lo, hi = int(co[b.seg_chunk[s]]), int(co[b.seg_chunk[s + 1]])
            r = int(out.served_round[k])
            assert np.array_equal(out.round_out[r][int(out.spans[k, 0]):int(out.spans[k, 1])], b.data[lo:hi]), s
    assert out.chunks_relayed == int((first_ok >= 0).sum()) * (n_events + 2)
    rows = dict(out.usage_rows())
    for k in list(rows)[:64]:
        assert rows[k] == b.truths[int(mine[k])].expected_row()
    # the oracle walk: every exhausted request, and a sample of the served ones
    rot = chain_oracle.Rotation()
    sample = set(np.nonzero(first_ok < 0)[0].tolist()) | set(range(0, n, max(1, n // 48)))
    assert (first_ok < 0).sum() > 0 or n < 1000
    for k in sorted(sample):
        s = int(mine[k])
        want = chain_oracle.walk(bodies[k], {}, providers, rules, fallback_provider, lambda a: up.stream_chunks(s, a), rot, cc.stream_mode())
        if want["kind"] == "stream":
            assert out.emitted(k) == want["emitted"], s
        else:
            assert out.detail[k] == want["detail"], (s, out.detail[k], want["detail"])
    eng.close_engine()
