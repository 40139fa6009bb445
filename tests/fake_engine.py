"""CPU stand-in for llmapigateway_b200.Engine used ONLY by host-logic tests (batcher, gateway seam,
gloo sharding): it replays each stream through the host build of the device machine
(tests/support, test aid).  It keeps every stream's history and re-runs it per step."""
from __future__ import annotations

import numpy as np

import host_machine as hm
from llmapigateway_b200 import _abi
from llmapigateway_b200.engine import SEG_DTYPE, StepResult


class _Limits:
    def __init__(self, max_streams):
        self.max_streams = max_streams
        self.detail_cap = 4096
        self.rowq_cap = 64


class FakeEngine:
    def __init__(self, max_streams=64):
        self.limits = _Limits(max_streams)
        self.hist = {}

    def open(self, slots, http_status=None):
        for i, s in enumerate(slots):
            self.hist[int(s)] = dict(chunks=[], steps=[], status=200 if http_status is None else int(http_status[i]), rows_seen=0, last=None)

    def _run(self, h):
        return hm.run_stream(h["chunks"], h["steps"] or [0], h["status"])

    def step(self, data, chunk_off, seg_chunk, seg_slot, out=None, relay_from_host=False):
        data = np.asarray(data, dtype=np.uint8)
        segs = np.zeros(len(seg_slot), dtype=SEG_DTYPE)
        rows = []
        for k, slot in enumerate(seg_slot):
            h = self.hist[int(slot)]
            c0, c1 = int(seg_chunk[k]), int(seg_chunk[k + 1])
            h["steps"].append(len(h["chunks"]))
            base = len(h["chunks"])
            h["chunks"] += [data[int(chunk_off[c]):int(chunk_off[c + 1])].tobytes() for c in range(c0, c1)]
            r = self._run(h)
            seg = r["segs"][-1]
            segs[k] = (c0 + (int(seg.emit_chunk_begin) - base), seg.phase, seg.verdict, seg.flags, seg.detail_len)
            for ev in r["rows"][h["rows_seen"]:]:
                ev.slot = int(slot)
                rows.append(ev)
            h["rows_seen"] = len(r["rows"])
            h["last"] = r
        return StepResult(data.copy(), segs, rows)

    def detail(self, slot):
        return self.hist[int(slot)]["last"]["detail"]

    def state(self, slots):
        return [self._run(self.hist[int(s)])["state"] for s in slots]

    def close(self, slots):
        out = self.state(slots)
        for s in slots:
            self.hist.pop(int(s), None)
        return out


    # ---- request bodies / responses: host build of the body machine (fake_body_engine.HostBodyEngine) ----
    def load_rules(self, plans):
        from fake_body_engine import HostBodyEngine
        self._body = HostBodyEngine(plans, fast=True)

    def rewrite_bodies(self, bodies, plan_idx, slot_cap=None, with_matched=False):
        return self._body.rewrite_bodies(bodies, plan_idx, slot_cap, with_matched)

    def scan_bodies(self, bodies, model_cap: int = 256):
        from llmapigateway_b200.rewrite import SCAN_DTYPE
        scans = np.zeros(len(bodies), dtype=SCAN_DTYPE)
        texts = []
        for i, b in enumerate(bodies):
            sc, model = hm.scan_body(bytes(b), model_cap)
            scans[i] = sc
            texts.append(model)
        return scans, texts

    def rewrite_packed(self, buf, off, plan_idx, slot_cap, out=None):
        from llmapigateway_b200.rewrite import RESULT_DTYPE
        bodies = [bytes(buf[int(off[i]):int(off[i + 1])]) for i in range(len(off) - 1)]
        rows = self._body.rewrite_bodies(bodies, plan_idx, slot_cap, True)
        res = np.zeros(len(rows), dtype=RESULT_DTYPE)
        out_off = np.zeros(len(rows) + 1, dtype=np.uint64)
        blobs = []
        for i, (st, b, matched, root) in enumerate(rows):
            res[i] = (st, len(b), matched, root)
            blobs.append(b)
            out_off[i + 1] = out_off[i] + len(b)
        return np.frombuffer(b"".join(blobs) or b"\0", dtype=np.uint8), out_off, res

    def details(self, slots, stride=None):
        return [self.detail(s) for s in slots]

    def scan_packed(self, buf, off, model_cap: int = 256):
        bodies = [bytes(buf[int(off[i]):int(off[i + 1])]) for i in range(len(off) - 1)]
        scans, texts = self.scan_bodies(bodies, model_cap)
        models = np.zeros((len(bodies), model_cap), dtype=np.uint8)
        for i, t in enumerate(texts):
            models[i, :len(t)] = np.frombuffer(t, dtype=np.uint8)
        return scans, models

    def documents_usage(self, docs):
        """same result shape as Engine.documents_usage, from the host build of the document tap"""
        res = []
        for d in docs:
            o = hm.doc_usage(bytes(d))
            row = _abi.usage_rec_to_dict(o.rec)
            res.append(([dict(row), row] if o.error_row else [row], bool(o.exotic)))
        return res

    def documents_error_detail(self, docs, text_stride: int = 4096):
        return [hm.error_detail(bytes(d), text_stride) for d in docs]
