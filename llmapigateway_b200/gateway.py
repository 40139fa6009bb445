"""Host-side mirror of the reference's streaming seam, backed by the engine.

  make_llm_request(...)   <- llm_gateway_core/services/request_handler.py:8   (same signature, same
                             (response, error_detail) return convention, never raises)
  StreamBatcher           <- the per-chunk bodies of request_handler.py:34-142 and the response tap of
                             chat_logging.py:165-231, turned into batched engine steps (SURVEY 8(b) threading)
  shard_of(stream_id, n)  <- SURVEY 8(e): stream -> GPU by hash, no cross-GPU dependency

The upstream HTTP I/O stays httpx exactly as in the reference (request_handler.py:15,23); only the
byte work moves to the GPU.  Non-streaming requests (rows a4/a12) are not accelerated in this round
and are left to the reference's own code.
"""
from __future__ import annotations

import asyncio
import concurrent.futures
import zlib
from dataclasses import dataclass, field

import numpy as np

from . import _abi


def shard_of(stream_id, n_gpus: int) -> int:
    """Stable stream -> GPU map (crc32, not Python's salted hash)."""
    return zlib.crc32(str(stream_id).encode("utf-8")) % max(1, n_gpus)


@dataclass
class FeedResult:
    emitted: bytes | None        # the chunk to relay now (original bytes), or None when dropped
    phase: int
    verdict: int


@dataclass
class _Pending:
    slot: int
    chunk: bytes
    fut: asyncio.Future


class StreamBatcher:
    """One per engine/GPU.  Coroutines call feed(); a pump task packs everything that arrived during
    `window_s` into one engine step (run in a worker thread: ctypes releases the GIL)."""

    def __init__(self, engine, window_s: float = 0.002, usage_sink=None):
        self.eng = engine
        self.window_s = window_s
        self.usage_sink = usage_sink
        self._free = list(range(engine.limits.max_streams - 1, -1, -1))
        self._pending: list[_Pending] = []
        self._wake = asyncio.Event()
        self._task: asyncio.Task | None = None
        self._closed = False
        self.steps = 0
        # the engine handle is not thread-safe: every call into it goes through this one worker thread
        self._worker = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="lgw-engine")

    # -- slots ---------------------------------------------------------------------------------------
    async def open_stream(self, http_status: int = 200) -> int:
        if not self._free:
            raise RuntimeError("no free stream slot on this engine")
        slot = self._free.pop()
        await asyncio.get_running_loop().run_in_executor(self._worker, self.eng.open, [slot], [http_status])
        return slot

    async def close_stream(self, slot: int):
        """End of upstream: returns (final StreamState, usage dict or None).  The usage dict is the
        last DB row of chat_logging.py:150 and goes to the usage sink (TokensUsageDB.insert_usage seam)."""
        st = (await asyncio.get_running_loop().run_in_executor(self._worker, self.eng.close, [slot]))[0]
        self._free.append(slot)
        usage = None
        if st.flags & _abi.SF_EMITTED_ANY:
            usage = _abi.usage_rec_to_dict(st.rec)
            if self.usage_sink is not None:
                self.usage_sink.insert_usage(usage)
        return st, usage

    # -- request bodies / non-streaming responses (rows a1-a4, a12): same worker thread, same engine ---------
    def load_rules(self, plans) -> None:
        """Compiled plan table (rewrite.RulePlans); call at config load / hot reload, before serving."""
        self._worker.submit(self.eng.load_rules, plans).result()
        self.plans = plans

    async def rewrite_bodies(self, bodies, plan_idx):
        return await asyncio.get_running_loop().run_in_executor(self._worker, self.eng.rewrite_bodies, list(bodies), list(plan_idx))

    async def scan_bodies(self, bodies):
        return await asyncio.get_running_loop().run_in_executor(self._worker, self.eng.scan_bodies, list(bodies))

    async def normalise_responses(self, contents, http_status, target_url: str):
        from .responses import normalise_responses
        return await asyncio.get_running_loop().run_in_executor(
            self._worker, lambda: normalise_responses(self.eng, self.plans, list(contents), list(http_status), target_url))

    async def detail(self, slot: int) -> str:
        raw = await asyncio.get_running_loop().run_in_executor(self._worker, self.eng.detail, slot)
        return raw.decode("utf-8", errors="replace")

    # -- data ------------------------------------------------------------------------------------------
    async def feed(self, slot: int, chunk: bytes) -> FeedResult:
        loop = asyncio.get_running_loop()
        fut = loop.create_future()
        self._pending.append(_Pending(slot, bytes(chunk), fut))
        if self._task is None or self._task.done():
            self._task = loop.create_task(self._pump())
        self._wake.set()
        return await fut

    async def _pump(self):
        loop = asyncio.get_running_loop()
        while self._pending:
            await asyncio.sleep(self.window_s)
            batch, self._pending = self._pending, []
            by_slot: dict[int, list[_Pending]] = {}
            for p in batch:
                by_slot.setdefault(p.slot, []).append(p)
            slots = list(by_slot)
            blobs, offs, segc = [], [0], [0]
            for s in slots:
                for p in by_slot[s]:
                    blobs.append(p.chunk); offs.append(offs[-1] + len(p.chunk))
                segc.append(segc[-1] + len(by_slot[s]))
            data = np.frombuffer(b"".join(blobs), dtype=np.uint8) if blobs else np.zeros(0, np.uint8)
            try:
                res = await loop.run_in_executor(self._worker, self.eng.step, data, np.array(offs, np.uint32), np.array(segc, np.uint32), np.array(slots, np.uint32))
            except Exception as exc:                      # engine failure: the endpoint answers 500/503 like chat.py:26,198
                for p in batch:
                    if not p.fut.done():
                        p.fut.set_exception(exc)
                continue
            self.steps += 1
            if self.usage_sink is not None:
                for ev in sorted(res.rows, key=lambda r: (r.slot, r.seq)):      # mid-stream rows, chat_logging.py:139
                    self.usage_sink.insert_usage(_abi.usage_rec_to_dict(ev.rec))
            for k, s in enumerate(slots):
                eb = int(res.segs["emit_chunk_begin"][k])
                for j, p in enumerate(by_slot[s]):
                    c = segc[k] + j
                    out = None
                    if c >= eb and len(p.chunk):
                        out = res.out[offs[c]:offs[c + 1]].tobytes()          # the re-emitted bytes
                    if not p.fut.done():
                        p.fut.set_result(FeedResult(out, int(res.segs["phase"][k]), int(res.segs["verdict"][k])))


async def make_llm_request(target_url: str, headers: dict, payload: dict, is_streaming: bool, *, batcher: StreamBatcher,
                           client_factory=None):
    """Drop-in for request_handler.py:8.  Returns (response, None) on success and (None, error_detail) on
    failure; never raises (request_handler.py:178-187).  `payload` is the attempt's body: the bytes
    `StreamBatcher.rewrite_bodies` produced (rows a3/a4) -- a dict is still accepted on the streaming branch
    and encoded by httpx as in the reference.  Non-streaming success is a ready `Response` whose body is
    byte-identical to what FastAPI renders from the reference's returned dict (row a12)."""
    import httpx
    from fastapi.responses import Response, StreamingResponse
    client = (client_factory or (lambda **kw: httpx.AsyncClient(**kw)))(timeout=httpx.Timeout(300.0, connect=60.0))
    slot = None
    body_kw = {"content": bytes(payload)} if isinstance(payload, (bytes, bytearray, memoryview)) else {"json": payload}
    if not is_streaming:                                                  # request_handler.py:152-176
        try:
            if "json" in body_kw:
                raise TypeError("non-streaming payloads must be the bytes produced by rewrite_bodies (no CPU serialiser in this package)")
            response = await client.post(target_url, headers=headers, timeout=None, **body_kw)
            body, detail = (await batcher.normalise_responses([response.content], [response.status_code], target_url))[0]
            if body is None:
                return None, detail
            return Response(content=body, status_code=200, media_type="application/json"), None
        except httpx.RequestError as e:                                   # request_handler.py:178-182
            return None, f"RequestError connecting to {target_url}: {str(e)}"
        except Exception as e:                                            # request_handler.py:183-187
            return None, f"Unexpected error during request to {target_url}: {str(e)}"
    try:
        ctx = client.stream("POST", target_url, headers=headers, timeout=None, **body_kw)
        response = await ctx.__aenter__()
        if response.status_code >= 400:                                   # request_handler.py:25-30
            body = await response.aread()
            await ctx.__aexit__(None, None, None)
            return None, body.decode("utf-8")
        slot = await batcher.open_stream(response.status_code)
        chunks = response.aiter_bytes()
        first_kept: list[bytes] = []
        committed = False
        async for chunk in chunks:                                        # priming: request_handler.py:69-95
            r = await batcher.feed(slot, chunk)
            if r.phase == _abi.PHASE_FAILED:
                detail = await batcher.detail(slot)
                if r.verdict == _abi.VERDICT_FAIL_PARSE:                  # :183-187 (message tail is the JSON library's text: unpinned)
                    detail = f"Unexpected error during request to {target_url}: first event is not valid JSON: {detail[:200]}"
                await batcher.close_stream(slot)
                await ctx.__aexit__(None, None, None)
                return None, detail
            if r.emitted is not None:
                first_kept.append(r.emitted)
            if r.phase == _abi.PHASE_COMMITTED:
                committed = True
                break

        async def relay():                                                # combined_generator, request_handler.py:100-144
            try:
                for c in first_kept:
                    yield c
                if committed:
                    async for chunk in chunks:
                        r = await batcher.feed(slot, chunk)
                        if r.emitted is not None:
                            yield r.emitted
            finally:
                await batcher.close_stream(slot)
                await ctx.__aexit__(None, None, None)

        return StreamingResponse(relay(), media_type="text/event-stream",
                                 headers={"Transfer-Encoding": "chunked", "X-Accel-Buffering": "no"}), None
    except httpx.RequestError as e:                                       # request_handler.py:178-182
        if slot is not None:
            await batcher.close_stream(slot)
        return None, f"RequestError connecting to {target_url}: {str(e)}"
    except Exception as e:                                                # request_handler.py:183-187
        if slot is not None:
            try:
                await batcher.close_stream(slot)
            except Exception:
                pass
        return None, f"Unexpected error during request to {target_url}: {str(e)}"


class SqliteUsageSink:
    """Writes GPU-extracted usage records into the reference's `tokens_usage` table (same schema,
    same ISO-text timestamp: tokens_usage_db.py:37-50,135) so the existing stats endpoints keep working.
    One transaction per batch instead of open-insert-commit-close per row (:131-153)."""

    def __init__(self, db_path):
        import sqlite3
        self.conn = sqlite3.connect(db_path, check_same_thread=False)
        self.conn.execute("""CREATE TABLE IF NOT EXISTS tokens_usage (id INTEGER PRIMARY KEY AUTOINCREMENT, timestamp DATETIME NOT NULL,
            prompt_tokens INTEGER DEFAULT 0, completion_tokens INTEGER DEFAULT 0, total_tokens INTEGER DEFAULT 0,
            reasoning_tokens INTEGER DEFAULT 0, cached_tokens INTEGER DEFAULT 0, cost REAL DEFAULT 0.0, model TEXT, provider TEXT)""")
        self.conn.execute("CREATE INDEX IF NOT EXISTS idx_tokens_usage_timestamp ON tokens_usage (timestamp)")
        self.conn.commit()

    def insert_many(self, usages, timestamps=None):
        from datetime import datetime
        rows = []
        for i, u in enumerate(usages):
            ts = (timestamps[i] if timestamps else datetime.now()).isoformat()
            g = u.get
            rows.append((ts, g("prompt_tokens", 0), g("completion_tokens", 0), g("total_tokens", 0), g("reasoning_tokens", 0),
                         g("cached_tokens", 0), g("cost", 0.0), g("model"), g("provider")))
        try:
            with self.conn:
                self.conn.executemany("INSERT INTO tokens_usage (timestamp, prompt_tokens, completion_tokens, total_tokens, reasoning_tokens,"
                                      " cached_tokens, cost, model, provider) VALUES (?,?,?,?,?,?,?,?,?)", rows)
        except Exception:        # the reference swallows insert errors (:155-159)
            pass

    def insert_usage(self, usage: dict):
        self.insert_many([usage])
