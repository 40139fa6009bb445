"""Drive the UNMODIFIED reference modules from /root/reference (dev container only).

Test infrastructure. Used by make_golden.py to produce the committed fixtures in
tests/golden/*.json; nothing here is imported by the product or by tests that run
on the GPU box (where /root/reference does not exist).

What is driven, and how:
  * llm_gateway_core/services/request_handler.py:8  make_llm_request  -- through
    httpx.MockTransport, so the real stream_generator / priming loop /
    combined_generator run on our synthetic chunk lists.
  * llm_gateway_core/middleware/chat_logging.py:69  ChunkProcessorThread and
    :233 get_token_usage -- run() is called synchronously on the emitted chunks,
    write_log is replaced by a recorder (it is the DB-row point, :54).
`json5` is absent from this image; a shim module maps loads/load to
the stdlib json (identical on strict RFC-8259 text, which is all the fixtures use).
"""
from __future__ import annotations

import asyncio
import copy
import json
import logging
import queue
import sys
import tempfile
import types
from pathlib import Path

REF = Path("/root/reference")


def _install_json5_shim():
    if "json5" in sys.modules:
        return
    shim = types.ModuleType("json5")

    def _strip_comments(text: str) -> str:
        out, i, n, in_str = [], 0, len(text), False
        while i < n:
            c = text[i]
            if in_str:
                out.append(c)
                if c == "\\" and i + 1 < n:
                    out.append(text[i + 1]); i += 1
                elif c == '"':
                    in_str = False
            elif c == '"':
                in_str = True; out.append(c)
            elif c == "/" and text[i:i + 2] == "//":
                while i < n and text[i] != "\n":
                    i += 1
                continue
            else:
                out.append(c)
            i += 1
        return "".join(out)

    shim.loads = lambda s, **kw: json.loads(s)
    shim.load = lambda fp, **kw: json.loads(_strip_comments(fp.read()))
    shim.dumps = lambda o, **kw: json.dumps(o)
    shim.JSONDecodeError = json.JSONDecodeError
    shim.__doc__ = "stdlib-json shim for the missing json5 package (strict JSON only)"
    sys.modules["json5"] = shim


_loaded = {}


def load_reference():
    """Import the reference modules read-only; returns (request_handler, chat_logging)."""
    if _loaded:
        return _loaded["rh"], _loaded["cl"]
    if not REF.exists():
        raise RuntimeError("/root/reference is not present (golden generation is dev-container only)")
    _install_json5_shim()
    sys.path.insert(0, str(REF))
    logging.disable(logging.CRITICAL)
    # tokens_usage_db.py:17-25 hard-codes <root>/db (read-only here): point it at a temp file.
    import llm_gateway_core.db.tokens_usage_db as tdb
    tmp = Path(tempfile.mkdtemp(prefix="lgw_ref_")) / "tokens_usage.db"

    def _init(self, db_filename: str = "tokens_usage.db"):
        self.db_path = tmp
        self._init_db()

    tdb.TokensUsageDB.__init__ = _init
    import llm_gateway_core.services.request_handler as rh
    import llm_gateway_core.middleware.chat_logging as cl
    _loaded.update(rh=rh, cl=cl, tdb=tdb)
    return rh, cl


def run_relay(chunks: list[bytes], http_status: int = 200, url: str = "http://upstream.test/v1/chat/completions"):
    """Real make_llm_request(..., is_streaming=True) over a MockTransport upstream.

    Returns dict(failed, error_detail, emitted=[bytes...], end_exception=str|None).
    """
    import httpx
    rh, _ = load_reference()

    class _Body(httpx.AsyncByteStream):
        async def __aiter__(self):
            for c in chunks:
                yield c

    def handler(request):
        return httpx.Response(http_status, headers={"content-type": "text/event-stream"}, stream=_Body())

    real_client = httpx.AsyncClient

    def patched(**kw):
        return real_client(transport=httpx.MockTransport(handler), **kw)

    async def go():
        rh.httpx.AsyncClient = patched
        try:
            resp, err = await rh.make_llm_request(url, {}, {"model": "m", "messages": []}, True)
        finally:
            rh.httpx.AsyncClient = real_client
        if resp is None:
            return dict(failed=True, error_detail=err, emitted=[], end_exception=None)
        out, end_exc = [], None
        try:
            async for c in resp.body_iterator:
                out.append(bytes(c))
        except Exception as e:  # request_handler.py:144 UnboundLocalError when no usage was seen
            end_exc = type(e).__name__
        return dict(failed=False, error_detail=err, emitted=out, end_exception=end_exc)

    return asyncio.run(go())


def run_nonstream(content: bytes, http_status: int = 200, url: str = "http://upstream.test/v1/chat/completions"):
    """Real make_llm_request(..., is_streaming=False) (request_handler.py:152-176) over a MockTransport upstream,
    then what chat.py:146 and FastAPI do with the result.  Returns dict(kind, detail, body)."""
    import httpx
    from starlette.responses import JSONResponse
    rh, _ = load_reference()

    def handler(request):
        return httpx.Response(http_status, headers={"content-type": "application/json"}, content=content)

    real_client = httpx.AsyncClient

    def patched(**kw):
        return real_client(transport=httpx.MockTransport(handler), **kw)

    async def go():
        rh.httpx.AsyncClient = patched
        try:
            return await rh.make_llm_request(url, {}, {"model": "m", "messages": []}, False)
        finally:
            rh.httpx.AsyncClient = real_client

    resp, err = asyncio.run(go())
    if resp and err is None:                                   # chat.py:146
        try:
            return dict(kind="ok", detail=None, body=bytes(JSONResponse(content=resp).body))
        except ValueError:
            return dict(kind="raise", detail=None, body=b"")
    return dict(kind="fail", detail=err, body=b"")


class _NoWaitQueue(queue.Queue):
    def get(self, block=True, timeout=None):  # chat_logging.py:94 waits 5 s; we know the stream is over
        return super().get(block=False)


def run_tap(emitted: list[bytes], is_real_streaming: bool = True):
    """Real ChunkProcessorThread.run() over the emitted chunks. Returns (rows, transcripts).

    rows = the tokens_usage dict at every write_log call (chat_logging.py:139,150 -> :54 insert_usage).
    No thread is created when nothing was emitted (chat_logging.py:198-203).
    """
    _, cl = load_reference()
    if not emitted:
        return [], []
    rows, texts = [], []
    real = cl.write_log

    def rec(h, b, accum, usage):
        rows.append(copy.deepcopy(usage)); texts.append(accum)

    cl.write_log = rec
    try:
        t = cl.ChunkProcessorThread({}, "", is_real_streaming)
        t.queue = _NoWaitQueue()
        for c in emitted:
            t.enqueue_chunk(c)
        t.run()
    finally:
        cl.write_log = real
    return rows, texts


def get_token_usage(d):
    _, cl = load_reference()
    return cl.get_token_usage(d)
