"""TEST AID: the bulk kernels of csrc/relay2.cuh run on the CPU through tests/support/host_relay2.cpp (SIMT emulator).
`HostBulkEngine` offers the subset of `llmapigateway_b200.Engine` the SSE parity tests use, so the same test bodies run
against the real engine on the GPU box and against the emulated kernels here.  Not a product path."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from llmapigateway_b200 import _abi
from llmapigateway_b200.engine import SEG_DTYPE, StepResult

SUP = Path(__file__).resolve().parent / "support"
SRC = SUP / "host_relay2.cpp"
LIB = SUP / "_host_relay2.so"
CSRC = Path(__file__).resolve().parent.parent / "llmapigateway_b200" / "csrc"

_lib = None


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [SRC, SUP / "simt_emu.h"] + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.inc")) + [CSRC.parent.parent / "include" / "llmgw_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def lib():
    global _lib
    if _lib is None:
        if _stale():
            tmp = LIB.with_suffix(".%d.tmp" % os.getpid())      # (built aside and moved into place: other processes may be loading the library)
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-g", "-shared", "-fPIC", "-o", str(tmp), str(SRC)])
            os.replace(tmp, LIB)
        _lib = C.CDLL(str(LIB))
        _lib.lgwt_bulk_new.restype = C.c_void_p
        _lib.lgwt_bulk_new.argtypes = [C.c_uint32] * 5
        _lib.lgwt_bulk_free.argtypes = [C.c_void_p]
        _lib.lgwt_bulk_set_mode.argtypes = [C.c_void_p, C.c_int]
        _lib.lgwt_bulk_reset_templates.argtypes = [C.c_void_p]
        _lib.lgwt_bulk_open.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        _lib.lgwt_bulk_step.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32]
        _lib.lgwt_bulk_state.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
        _lib.lgwt_bulk_detail.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        _lib.lgwt_bulk_detail.restype = C.c_uint32
        _lib.lgwt_bulk_templates.argtypes = [C.c_void_p, C.c_void_p]
        _lib.lgwt_bulk_counters.argtypes = [C.c_void_p, C.c_void_p]
        _lib.lgwt_bulk_transcripts_enable.argtypes = [C.c_void_p]
        _lib.lgwt_bulk_transcript.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        _lib.lgwt_bulk_transcript.restype = C.c_uint32
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class HostBulkEngine:
    def __init__(self, max_streams: int = 2048, carry_cap: int = 4096, detail_cap: int = 4096, rowq_cap: int = 4096,
                 n_blocks: int = 2, tiles_per_warp: int = 0):
        self._lib = lib()
        self.limits = _abi.Limits(max_streams, carry_cap, detail_cap, rowq_cap, 1 << 22, 1 << 28)
        self._h = C.c_void_p(self._lib.lgwt_bulk_new(max_streams, carry_cap, detail_cap, rowq_cap, n_blocks))
        self.tiles_per_warp = tiles_per_warp

    def close_engine(self):
        if self._h:
            self._lib.lgwt_bulk_free(self._h)
            self._h = None

    def set_mode(self, mode: int):
        self._lib.lgwt_bulk_set_mode(self._h, mode)

    def reset_templates(self):
        self._lib.lgwt_bulk_reset_templates(self._h)

    def open(self, slots, http_status=None):
        slots = np.ascontiguousarray(slots, dtype=np.uint32)
        st = np.full(len(slots), 200, dtype=np.int32) if http_status is None else np.ascontiguousarray(http_status, dtype=np.int32)
        self._lib.lgwt_bulk_open(self._h, _ptr(slots), _ptr(st), len(slots))

    def _states(self, slots, free_after):
        slots = np.ascontiguousarray(slots, dtype=np.uint32)
        out = (_abi.StreamState * max(1, len(slots)))()
        self._lib.lgwt_bulk_state(self._h, _ptr(slots), len(slots), out, free_after)
        return [out[i] for i in range(len(slots))]

    def state(self, slots):
        return self._states(slots, 0)

    def close(self, slots):
        return self._states(slots, 1)

    def detail(self, slot: int) -> bytes:
        buf = C.create_string_buffer(self.limits.detail_cap)
        n = self._lib.lgwt_bulk_detail(self._h, slot, buf, self.limits.detail_cap)
        return buf.raw[:n]

    def step(self, data, chunk_off, seg_chunk, seg_slot, out=None, relay_from_host=False) -> StepResult:
        data = np.ascontiguousarray(data, dtype=np.uint8)
        chunk_off = np.ascontiguousarray(chunk_off, dtype=np.uint32)
        seg_chunk = np.ascontiguousarray(seg_chunk, dtype=np.uint32)
        seg_slot = np.ascontiguousarray(seg_slot, dtype=np.uint32)
        n_bytes, n_chunks, n_segs = data.size, chunk_off.size - 1, seg_slot.size
        # the buffers are sized exactly: the emulated kernels must not touch a byte beyond them (run under ASan/valgrind to check)
        out = np.zeros(max(n_bytes, 1), dtype=np.uint8) if out is None else out
        segs = np.zeros(max(n_segs, 1), dtype=SEG_DTYPE)
        cap = self.limits.rowq_cap
        rows = (_abi.RowEvent * max(cap, 1))()
        n_rows = C.c_uint32(0)
        self._lib.lgwt_bulk_step(self._h, _ptr(data), n_bytes, _ptr(chunk_off), n_chunks, _ptr(seg_chunk), _ptr(seg_slot), n_segs,
                                 _ptr(out), _ptr(segs), rows, cap, C.byref(n_rows), self.tiles_per_warp)
        self._last = (data, chunk_off, seg_chunk, seg_slot, segs)
        return StepResult((data if relay_from_host else out)[:n_bytes], segs[:n_segs], [rows[i] for i in range(n_rows.value)])

    def enable_transcripts(self):
        self._lib.lgwt_bulk_transcripts_enable(self._h)

    def step_transcript(self):
        from llmapigateway_b200.engine import StepText
        data, chunk_off, seg_chunk, seg_slot, segs = self._last
        n_bytes, n_chunks, n_segs = data.size, chunk_off.size - 1, seg_slot.size
        text = np.zeros(2 * (n_bytes + n_segs * self.limits.carry_cap) + 1, dtype=np.uint8)
        seg_off = np.zeros(n_segs + 1, dtype=np.uint64)
        flags = np.zeros(max(n_segs, 1), dtype=np.uint32)
        cap = self.limits.rowq_cap
        marks = (_abi.TextMark * max(cap, 1))()
        n = self._lib.lgwt_bulk_transcript(self._h, _ptr(data), n_bytes, _ptr(chunk_off), n_chunks, _ptr(seg_chunk), _ptr(seg_slot), n_segs, _ptr(segs),
                                           _ptr(text), _ptr(seg_off), _ptr(flags), marks, cap)
        return StepText(text[:int(seg_off[n_segs])], seg_off, flags[:n_segs], [(marks[i].slot, marks[i].seq, marks[i].text_pos) for i in range(n)])

    def counters(self):
        out = (C.c_uint32 * 4)()
        self._lib.lgwt_bulk_counters(self._h, out)
        return dict(sequential=out[0], bulk=out[1], from_template=out[2], stashed=out[3])

    def templates(self):
        out = (C.c_uint32 * 16)()
        self._lib.lgwt_bulk_templates(self._h, out)
        v = list(out)
        return [dict(state=v[i], len=v[4 + i], flags=v[8 + i] & 0x7FFFFFFF, usage_ok=v[8 + i] >> 31, hits=v[12 + i]) for i in range(4)]
