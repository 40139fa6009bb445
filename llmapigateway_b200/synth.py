"""Seeded synthetic upstreams (BASELINE.json north_star: "upstream providers are replaced
in-bench by a local synthetic SSE generator").  Workload shapes follow SURVEY.md section 8(d).

Everything here is host-side numpy; it produces the packed step layout the engine takes
(see include/llmgw_b200.h): one byte buffer, chunk offsets, and per-stream segments.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

DELTA_HEAD = b'data: {"choices":[{"index":0,"delta":{"content":"'
DELTA_TAIL = b'"}}]}\n\n'
DELTA_CONTENT = 8
DELTA_EVENT_BYTES = len(DELTA_HEAD) + DELTA_CONTENT + len(DELTA_TAIL)  # 64
DONE_EVENT = b"data: [DONE]\n\n"

# printable ASCII without '"' and '\\' 
_ALPHABET = np.array([c for c in range(0x20, 0x7F) if c not in (0x22, 0x5C)], dtype=np.uint8)

MODEL_NAMES = ["deepseek/deepseek-chat-v3-0324", "google/gemini-2.5-pro", "grok-3-mini-beta",
               "openai/gpt-4.1-mini", "deepseek-ai/DeepSeek-V3-0324", "anthropic/claude-sonnet"]
PROVIDER_NAMES = ["Chutes", "Targon", "DeepInfra", "Lambda", "Nebius AI Studio"]


@dataclass
class UsageTruth:
    prompt_tokens: int
    completion_tokens: int      # as reported upstream (before the reasoning subtraction)
    total_tokens: int
    reasoning_tokens: int
    cached_tokens: int
    cost_text: str
    model: str
    provider: str

    def expected_row(self) -> dict:
        """What chat_logging.py:233-272 makes of this usage event."""
        comp = self.completion_tokens - self.reasoning_tokens if self.reasoning_tokens > 0 else self.completion_tokens
        return {"prompt_tokens": self.prompt_tokens, "completion_tokens": comp,
                "total_tokens": self.total_tokens, "reasoning_tokens": self.reasoning_tokens,
                "cached_tokens": self.cached_tokens, "cost": float(self.cost_text),
                "provider": self.provider, "model": self.model}


def usage_event(u: UsageTruth) -> bytes:
    return (b'data: {"choices":[],"usage":{"prompt_tokens":%d,"completion_tokens":%d,"total_tokens":%d,'
            b'"cost":%s,"completion_tokens_details":{"reasoning_tokens":%d},'
            b'"prompt_tokens_details":{"cached_tokens":%d}},"model":"%s","provider":"%s"}\n\n'
            % (u.prompt_tokens, u.completion_tokens, u.total_tokens, u.cost_text.encode(),
               u.reasoning_tokens, u.cached_tokens, u.model.encode(), u.provider.encode()))


def _usage_truths(n_streams: int, seed: int) -> list[UsageTruth]:
    rng = np.random.default_rng([seed, 0xC3])
    p = rng.integers(1, 2**31 - 1, n_streams)
    r = rng.integers(0, 2**20, n_streams)
    c = r + rng.integers(1, 2**30, n_streams)
    h = rng.integers(0, 2**20, n_streams)
    k = rng.integers(0, 10**6, n_streams)
    mi = rng.integers(0, len(MODEL_NAMES), n_streams)
    pi = rng.integers(0, len(PROVIDER_NAMES), n_streams)
    out = []
    for i in range(n_streams):
        out.append(UsageTruth(int(p[i]), int(c[i]), int(min(p[i] + c[i], 2**31 - 1)), int(r[i]), int(h[i]),
                              "%.6f" % (int(k[i]) / 1e6), MODEL_NAMES[mi[i]], PROVIDER_NAMES[pi[i]]))
    return out


@dataclass
class PackedBatch:
    """One engine step worth of upstream bytes, grouped per stream (segment)."""
    data: np.ndarray            # uint8 [total_bytes]
    chunk_off: np.ndarray       # uint32 [n_chunks + 1]
    seg_chunk: np.ndarray       # uint32 [n_segs + 1]  chunk range of each segment
    seg_slot: np.ndarray        # uint32 [n_segs]      engine stream slot of each segment
    n_delta_events: int         # the events BASELINE.json's chunks/s counts
    truths: list[UsageTruth]

    @property
    def n_chunks(self) -> int:
        return int(self.chunk_off.shape[0] - 1)

    def stream_chunks(self, seg: int) -> list[bytes]:
        c0, c1 = int(self.seg_chunk[seg]), int(self.seg_chunk[seg + 1])
        raw = self.data
        return [raw[int(self.chunk_off[c]):int(self.chunk_off[c + 1])].tobytes() for c in range(c0, c1)]


def sse_batch(n_streams: int = 4096, n_events: int = 512, seed: int = 3,
              events_per_chunk: int = 1, with_usage: bool = True, with_done: bool = True,
              slot_base: int = 0) -> PackedBatch:
    """n_streams x n_events delta events of exactly 64 B, then (not counted)
    one usage event and `data: [DONE]`, one event per network chunk (or `events_per_chunk`)."""
    assert n_events % events_per_chunk == 0
    rng = np.random.default_rng([seed, n_streams, n_events])
    truths = _usage_truths(n_streams, seed)
    head = np.frombuffer(DELTA_HEAD, dtype=np.uint8)
    tail = np.frombuffer(DELTA_TAIL, dtype=np.uint8)
    tails = [usage_event(t) if with_usage else b"" for t in truths]
    done = DONE_EVENT if with_done else b""
    tail_lens = np.array([len(t) for t in tails], dtype=np.int64)
    delta_bytes = n_events * DELTA_EVENT_BYTES
    seg_bytes = delta_bytes + tail_lens + len(done)
    seg_start = np.concatenate([[0], np.cumsum(seg_bytes)])
    data = np.empty(int(seg_start[-1]), dtype=np.uint8)
    chunks_per_seg = n_events // events_per_chunk + (1 if with_usage else 0) + (1 if with_done else 0)
    chunk_off = np.empty(n_streams * chunks_per_seg + 1, dtype=np.int64)
    delta_chunk = DELTA_EVENT_BYTES * events_per_chunk
    n_delta_chunks = n_events // events_per_chunk
    for s in range(n_streams):
        b = int(seg_start[s])
        ev = data[b:b + delta_bytes].reshape(n_events, DELTA_EVENT_BYTES)
        ev[:, :len(head)] = head
        ev[:, len(head):len(head) + DELTA_CONTENT] = _ALPHABET[rng.integers(0, len(_ALPHABET), (n_events, DELTA_CONTENT))]
        ev[:, len(head) + DELTA_CONTENT:] = tail
        pos = b + delta_bytes
        co = s * chunks_per_seg
        chunk_off[co:co + n_delta_chunks] = b + delta_chunk * np.arange(n_delta_chunks)
        k = co + n_delta_chunks
        if with_usage:
            t = tails[s]
            data[pos:pos + len(t)] = np.frombuffer(t, dtype=np.uint8)
            chunk_off[k] = pos; k += 1; pos += len(t)
        if with_done:
            data[pos:pos + len(done)] = np.frombuffer(done, dtype=np.uint8)
            chunk_off[k] = pos; k += 1; pos += len(done)
    chunk_off[-1] = seg_start[-1]
    assert chunk_off[-1] < 2**32
    seg_chunk = (np.arange(n_streams + 1, dtype=np.int64) * chunks_per_seg).astype(np.uint32)
    return PackedBatch(data, chunk_off.astype(np.uint32), seg_chunk,
                       (slot_base + np.arange(n_streams)).astype(np.uint32),
                       n_streams * n_events, truths)


def pack_streams(streams: list[list[bytes]], slots: list[int] | None = None) -> PackedBatch:
    """Pack arbitrary per-stream chunk lists (tests, adversarial cases) into a step."""
    offs, segc, blobs = [0], [0], []
    for chunks in streams:
        for c in chunks:
            blobs.append(c)
            offs.append(offs[-1] + len(c))
        segc.append(segc[-1] + len(chunks))
    data = np.frombuffer(b"".join(blobs), dtype=np.uint8).copy() if blobs else np.zeros(0, np.uint8)
    slots = list(range(len(streams))) if slots is None else slots
    return PackedBatch(data, np.array(offs, dtype=np.uint32), np.array(segc, dtype=np.uint32),
                       np.array(slots, dtype=np.uint32), 0, [])


# ---- config 2: client request bodies (BASELINE.json configs[1]: 1024 non-streaming requests, 4 KiB JSON bodies) ----
_WORDS = ("the of and to in is that for it as was with be by on not he this are or his from at which but have an had they you were "
          "their one all we can her has there been if more when will would who so no out up said what its about than into them only "
          "model token stream gateway provider fallback latency batch kernel memory request response usage prompt completion").split()
_NON_ASCII = ["é", "ü", "中文", "日本語", "—", "“quoted”", "\U0001F600", "naïve", "Ω"]


def chat_bodies(n: int = 1024, target_bytes: int = 4096, seed: int = 2, model: str = "gw/chain", non_ascii: float = 0.02,
                stream: bool = False) -> list[bytes]:
    """Synthetic OpenAI chat-completions request bodies of about `target_bytes` each, written the way
    client SDKs write them (json.dumps defaults or compact separators, a few escapes, some UTF-8)."""
    import json
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        msgs = [{"role": "system", "content": "You are a helpful assistant.\nAnswer briefly; cite \"sources\" when asked."}]
        body = {"model": model, "messages": msgs, "temperature": float(rng.choice([0.0, 0.2, 0.7, 1.0, 1.5])),
                "top_p": float(rng.choice([1.0, 0.9, 0.95])), "max_tokens": int(rng.choice([256, 512, 1024, 4096])), "stream": stream,
                "user": "user-%06d" % int(rng.integers(0, 10 ** 6))}
        if rng.random() < 0.3:
            body["stop"] = ["\n\n", "###"]
        if rng.random() < 0.2:
            body["tools"] = [{"type": "function", "function": {"name": "lookup", "description": "Look something up",
                                                                "parameters": {"type": "object", "properties": {"q": {"type": "string"}}, "required": ["q"]}}}]
        if rng.random() < 0.15:
            body["usage"] = {"include": True}
        compact = rng.random() < 0.5
        ascii_out = rng.random() < 0.5
        turn = 0
        while True:
            words = []
            for _ in range(int(rng.integers(20, 90))):
                words.append(_NON_ASCII[int(rng.integers(0, len(_NON_ASCII)))] if rng.random() < non_ascii else _WORDS[int(rng.integers(0, len(_WORDS)))])
                if rng.random() < 0.03:
                    words.append("\n")
            msgs.append({"role": "user" if turn % 2 == 0 else "assistant", "content": " ".join(words)})
            turn += 1
            text = json.dumps(body, ensure_ascii=ascii_out, separators=(",", ":") if compact else None)
            if len(text.encode("utf-8")) >= target_bytes - 200:
                break
        out.append(text.encode("utf-8"))
    return out


# ---- a realistic OpenAI-style stream (not BASELINE's 64-byte deltas): id / created / model on every chunk, a role chunk,
#      content pieces of varying length with the usual escapes, a finish chunk, the usage chunk, [DONE] -------------------
_PIECES = [" the", " of", " and", " model", ".", ",", "\\n", "\\n\\n", ' \\"quoted\\"', " stream", " token", " B200", " gateway", " caf\\u00e9",
           " 中文", " x", " 42", " -", " a/b", " long-ish-piece-of-text", " ok", "!", " \\\\path", " \U0001F600"]


def openai_stream(rng, n_deltas: int, truth: UsageTruth, chunk_id: str, events_per_chunk=(1, 1)) -> list[bytes]:
    """One upstream response as network chunks (each chunk = 1..k whole events here; use `recut` for arbitrary cuts)."""
    head = '{"id":"%s","object":"chat.completion.chunk","created":%d,"model":"%s",' % (chunk_id, 1700000000 + int(rng.integers(0, 10**7)), truth.model)
    ev = ['data: ' + head + '"choices":[{"index":0,"delta":{"role":"assistant","content":""},"logprobs":null,"finish_reason":null}]}\n\n']
    for _ in range(n_deltas):
        piece = _PIECES[int(rng.integers(0, len(_PIECES)))]
        ev.append('data: ' + head + '"choices":[{"index":0,"delta":{"content":"%s"},"logprobs":null,"finish_reason":null}]}\n\n' % piece)
    ev.append('data: ' + head + '"choices":[{"index":0,"delta":{},"logprobs":null,"finish_reason":"stop"}]}\n\n')
    ev.append(usage_event(truth).decode())
    ev.append(DONE_EVENT.decode())
    chunks, i = [], 0
    while i < len(ev):
        k = int(rng.integers(events_per_chunk[0], events_per_chunk[1] + 1))
        chunks.append("".join(ev[i:i + k]).encode("utf-8"))
        i += k
    return chunks


def openai_batch(n_streams: int = 1024, n_deltas: int = 256, seed: int = 9, events_per_chunk=(1, 1)) -> PackedBatch:
    rng = np.random.default_rng([seed, n_streams, n_deltas])
    truths = _usage_truths(n_streams, seed)
    streams = [openai_stream(rng, n_deltas, truths[s], "chatcmpl-%08x" % int(rng.integers(0, 2**32)), events_per_chunk) for s in range(n_streams)]
    b = pack_streams(streams)
    b.truths = truths
    return b


# ---- config 4: a 3-deep fallback chain with injected upstream failures --------------------------------------
@dataclass
class Provider:
    """Shape of loader.py:15-17 ProviderDetails as chat.py reads it (:92-94)."""
    baseUrl: str
    apikey: str


C4_MODEL = "gw/chain3"


def chain_world():
    """(providers_config, fallback_rules, fallback_provider): a 3-rule chain per gateway model, plus the rule shapes the
    walker has to get right -- rotation, retries with the log scrub, sub-provider ordering walked as a fallback list or sent
    as a hint, custom body params (one overriding `model`), custom headers, and the unknown-model fallback provider."""
    providers = {
        "alpha": Provider("http://alpha.test/v1", "ALPHA_KEY_ENV"),
        "beta": Provider("http://beta.test/api/v1/", "sk-beta-literal"),
        "gamma": Provider("http://gamma.test/v1", ""),
        "openrouter": Provider("http://openrouter.test/api/v1", "sk-or-literal"),
    }

    def rule(provider, model, **kw):
        r = {"provider": provider, "model": model, "use_provider_order_as_fallback": False, "custom_body_params": {}, "custom_headers": {}}
        r.update(kw)
        return r

    rules = {
        C4_MODEL: {"rotate_models": False, "fallback_models": [
            rule("alpha", "alpha/large-1"),
            rule("beta", "beta-chat-2", custom_body_params={"temperature": 0.25, "model": "ignored/by-the-walker"}, custom_headers={"X-Route": "beta"}),
            rule("gamma", "gamma-3-instruct")]},
        "gw/rotating": {"rotate_models": True, "fallback_models": [
            rule("alpha", "alpha/rot-a"), rule("beta", "beta-rot-b"), rule("gamma", "gamma-rot-c")]},
        "gw/retrying": {"rotate_models": False, "fallback_models": [
            rule("alpha", "alpha/flaky", retry_count=2, retry_delay=0),
            rule("gamma", "gamma-steady")]},
        "gw/or-fallback": {"rotate_models": False, "fallback_models": [
            rule("openrouter", "deepseek/deepseek-chat-v3-0324:free", providers_order=["Chutes", "Targon"], use_provider_order_as_fallback=True, retry_count=1, retry_delay=0),
            rule("beta", "beta-chat-2")]},
        "gw/or-hint": {"rotate_models": False, "fallback_models": [
            rule("openrouter", "google/gemini-2.5-pro", providers_order=["DeepInfra", "Lambda"], retry_count=1, retry_delay=0),
            rule("gamma", "gamma-3-instruct")]},
    }
    return providers, rules, "openrouter"


def chain_request_bodies(n: int, seed: int = 4, model: str = C4_MODEL, pad_to: int = 256) -> list[bytes]:
    """Streaming chat requests of exactly `pad_to` bytes (the C1 body shape, `stream: true`)."""
    rng = np.random.default_rng([seed, n, 7])
    out = []
    for i in range(n):
        head = ('{"model":"%s","stream":true,"user":"u%05d","messages":[{"role":"user","content":"' % (model, i)).encode()
        tail = b'"}]}'
        fill = max(0, pad_to - len(head) - len(tail))
        out.append(head + _ALPHABET[rng.integers(0, len(_ALPHABET), fill)].tobytes() + tail)
    return out


FAIL_NONE, FAIL_HTTP500, FAIL_ERROR_EVENT, FAIL_DETAIL_EVENT = range(4)


class ChainUpstream:
    """The provider side of config 4.  Attempt `a` of stream `sid` fails with probability `p_fail`, independently, the failure
    kind uniform over {HTTP 500 + text body, first event {"error":{"message":..}}, first event {"detail":..}}; an attempt that
    does not fail streams the stream's C3 payload (n_events 64-byte deltas, the usage event, [DONE]; one event per chunk)."""
    wants_urls = False

    def __init__(self, n_streams: int = 8192, n_events: int = 512, seed: int = 4, p_fail: float = 0.2, max_attempts: int = 8, lazy: bool = False):
        self.n, self.seed, self.p, self.n_events = n_streams, seed, p_fail, n_events
        # lazy: no packed batch; a stream's payload is generated when asked for (same shape, own seed) -- for drivers that only
        # ever touch a sample of the streams one at a time (the reference arm)
        self.batch = None if lazy else sse_batch(n_streams, n_events, seed)
        rng = np.random.default_rng([seed, n_streams, 1234])
        u = rng.random((max_attempts, n_streams))
        k = rng.integers(1, 4, (max_attempts, n_streams))
        self.kind = np.where(u < p_fail, k, 0).astype(np.uint8)              # [attempt, stream]
        self.calls = []

    def failure_chunks(self, sid: int, attempt: int, kind: int):
        if kind == FAIL_HTTP500:
            return (500, ("upstream exploded: stream %d attempt %d {\"trace\":\"x\"}" % (sid, attempt)).encode())
        if kind == FAIL_ERROR_EVENT:
            return [('data: {"error":{"message":"provider overloaded (stream %d, attempt %d)","code":503}}\n\n' % (sid, attempt)).encode()]
        return [('data: {"detail":"rate limited: stream %d attempt %d"}\n\n' % (sid, attempt)).encode(), b'data: {"choices":[]}\n\n']

    def stream_chunks(self, sid: int, attempt: int):
        kind = int(self.kind[attempt, sid])
        if kind:
            return self.failure_chunks(sid, attempt, kind)
        if self.batch is None:
            return sse_batch(1, self.n_events, self.seed * 100003 + sid).stream_chunks(0)
        return self.batch.stream_chunks(sid)

    def prepare(self, ids, max_rounds: int = 3, alloc=None):
        """Pre-compute the answers of every round for the requests `ids` (the upstream knows which of its attempts fail, so it
        knows who comes back); `alloc(nbytes) -> uint8 array` places the response bytes (e.g. in pinned memory, where a
        gateway's receive buffers live).  __call__ then hands the prepared round out when asked for exactly these requests."""
        self.prepared = {}
        ids = np.asarray(ids, dtype=np.int64)
        for a in range(max_rounds):
            if not ids.size:
                break
            ans = self._answer(a, ids)
            if alloc is not None and ans.data.size:
                buf = alloc(int(ans.data.size)); buf[:] = ans.data; ans.data = buf
            self.prepared[a] = (ids, ans)
            ids = ids[self.kind[a, ids] != 0]

    def __call__(self, attempt: int, ids, urls, payload_buf, payload_off):
        ids = np.asarray(ids, dtype=np.int64)
        self.calls.append((attempt, ids.copy()))
        hit = getattr(self, "prepared", {}).get(attempt)
        if hit is not None and np.array_equal(hit[0], ids):
            return hit[1]
        return self._answer(attempt, ids)

    def _answer(self, attempt: int, ids):
        from .chat import Answers
        kind = self.kind[attempt, ids]
        status = np.where(kind == FAIL_HTTP500, 500, 200).astype(np.int32)
        errors = [self.failure_chunks(int(s), attempt, FAIL_HTTP500)[1] for s in ids[kind == FAIL_HTTP500]]
        b = self.batch
        sids = ids[kind != FAIL_HTTP500]
        skind = kind[kind != FAIL_HTTP500]
        # packed response streams, in request order: whole C3 segments for the good attempts, short failing streams else
        fail_blobs = {int(s): self.failure_chunks(int(s), attempt, int(kd)) for s, kd in zip(sids[skind != 0], skind[skind != 0])}
        c0 = b.seg_chunk[sids].astype(np.int64); c1 = b.seg_chunk[sids + 1].astype(np.int64)
        n_chunks = np.where(skind == 0, c1 - c0, 0)
        for j in np.nonzero(skind != 0)[0]:
            n_chunks[j] = len(fail_blobs[int(sids[j])])
        seg_chunk = np.zeros(len(sids) + 1, np.uint32); np.cumsum(n_chunks, out=seg_chunk[1:])
        co = b.chunk_off.astype(np.int64)
        seg_bytes = np.where(skind == 0, co[c1] - co[c0], 0)
        for j in np.nonzero(skind != 0)[0]:
            seg_bytes[j] = sum(len(c) for c in fail_blobs[int(sids[j])])
        seg_start = np.zeros(len(sids) + 1, np.int64); np.cumsum(seg_bytes, out=seg_start[1:])
        data = np.empty(int(seg_start[-1]), np.uint8)
        chunk_off = np.empty(int(seg_chunk[-1]) + 1, np.int64)
        for j in range(len(sids)):
            lo = int(seg_start[j])
            if skind[j] == 0:
                s0 = int(co[c0[j]])
                data[lo:lo + int(seg_bytes[j])] = b.data[s0:s0 + int(seg_bytes[j])]
                chunk_off[int(seg_chunk[j]):int(seg_chunk[j + 1])] = co[c0[j]:c1[j]] - s0 + lo
            else:
                pos = lo
                for q, c in enumerate(fail_blobs[int(sids[j])]):
                    chunk_off[int(seg_chunk[j]) + q] = pos
                    data[pos:pos + len(c)] = np.frombuffer(c, np.uint8)
                    pos += len(c)
        chunk_off[-1] = seg_start[-1]
        return Answers(status, errors, data, chunk_off.astype(np.uint32), seg_chunk)
