"""Host side of the request-body rewrite.

The reference builds one payload per upstream attempt (llm_gateway_core/api/v1/chat.py:112-168) from the
client's body and the rule dicts of models_fallback_rules.json (config/loader.py:150-154).  Here the
rule dicts are compiled ONCE (at load / hot-reload time) into *plans* -- the ordered key assignments of
one attempt with key and value already rendered for one encoder -- and uploaded to the GPU
(`lgw_rules_load`); per request the engine only needs the plan index (`lgw_bodies_rewrite`).

Rule lookup and rotation (row a2: chat.py:48-78) stay in Python: a dict get and a list rotate.
"""
from __future__ import annotations

import json
import re
from dataclasses import dataclass, field

import numpy as np

MODES = {"httpx028": 0, "httpx027": 1, "json5": 2}

BODY_OK, BODY_PARSE_ERROR, BODY_NO_MODEL, BODY_OVERFLOW, BODY_EXOTIC, BODY_ENCODE_ERROR = range(6)
STATUS_NAMES = ["ok", "parse_error", "no_model", "overflow", "exotic", "encode_error"]

OP_DTYPE = np.dtype([("key_off", "<u4"), ("key_len", "<u4"), ("rkey_off", "<u4"), ("rkey_len", "<u4"),
                     ("rval_off", "<u4"), ("rval_len", "<u4"), ("flags", "<u4"), ("_pad", "<u4")])
PLAN_DTYPE = np.dtype([("op_begin", "<u4"), ("op_end", "<u4"), ("mode", "<u4"), ("_pad", "<u4")])
RESULT_DTYPE = np.dtype([("status", "<u4"), ("out_len", "<u4"), ("matched", "<u4"), ("root_kind", "<u4")])
KIND_STR, KIND_OBJ, KIND_ARR = 6, 8, 9       # lgw_kind values of container / string roots (include/llmgw_b200.h)
SCAN_DTYPE = np.dtype([("status", "<u4"), ("model_len", "<u4"), ("model_kind", "u1"), ("model_truthy", "u1"),
                       ("stream_kind", "u1"), ("stream_truthy", "u1"), ("_pad", "<u4")])
MAX_OPS_PER_PLAN = 32
OP_IF_ABSENT, OP_PRESENCE = 1, 2
PLAN_RESPONSE = 0x100          # or-ed into a plan's mode: the document is an upstream response (row a12)
RESPONSE_PROBES = ("error", "detail")      # request_handler.py:167

_ES5_WORDS = frozenset(
    "break case catch continue debugger default delete do else finally for function if in instanceof new return "
    "switch this throw try typeof var void while with class const enum export extends import super null true false "
    "implements interface let package private protected public static yield".split())
_IDENT_RE = re.compile(r"[A-Za-z_$][A-Za-z0-9_$]*\Z")
_J5_SHORT = {0x5C: "\\\\", 0x22: '\\"', 0x0A: "\\n", 0x0D: "\\r", 0x08: "\\b", 0x0C: "\\f", 0x09: "\\t", 0x0B: "\\v", 0x00: "\\0"}


def _json5_str(text: str) -> str:
    """json5.dumps string form: double quotes, short escapes, everything
    outside printable ASCII as \\uXXXX (UTF-16 units)."""
    buf = []
    units = text.encode("utf-16-le", "surrogatepass")
    for i in range(0, len(units), 2):
        cu = units[i] | (units[i + 1] << 8)
        if cu in _J5_SHORT:
            buf.append(_J5_SHORT[cu])
        elif 0x20 <= cu <= 0x7E:
            buf.append(chr(cu))
        else:
            buf.append("\\u%04x" % cu)
    return '"' + "".join(buf) + '"'


def _json5_text(v) -> str:
    if v is None:
        return "null"
    if v is True:
        return "true"
    if v is False:
        return "false"
    if isinstance(v, str):
        return _json5_str(v)
    if isinstance(v, int):
        return str(v)
    if isinstance(v, float):
        if v != v:
            return "NaN"
        if v in (float("inf"), float("-inf")):
            return "Infinity" if v > 0 else "-Infinity"
        return repr(v)
    if isinstance(v, dict):
        return "{" + ", ".join(render_key(k, 2) + ": " + _json5_text(x) for k, x in v.items()) + "}"
    if isinstance(v, (list, tuple)):
        return "[" + ", ".join(_json5_text(x) for x in v) + "]"
    raise TypeError("value of type %s in a rule's body parameters" % type(v).__name__)


def render_value(v, mode: int) -> bytes:
    """The bytes the reference's encoder for `mode` produces for a rule-side value."""
    if mode == 0:
        return json.dumps(v, ensure_ascii=False, separators=(",", ":"), allow_nan=False).encode("utf-8")
    if mode == 1:
        return json.dumps(v).encode("utf-8")
    return _json5_text(v).encode("utf-8")


def render_key(k: str, mode: int):
    if mode == 0:
        return json.dumps(k, ensure_ascii=False)
    if mode == 1:
        return json.dumps(k)
    return k if (_IDENT_RE.match(k) and k not in _ES5_WORDS) else _json5_str(k)


def attempt_assignments(rule: dict, provider_name: str, sub_provider=None, retry: bool = False):
    """`payload[key] = value` statements chat.py executes before one upstream attempt, in order, as
    (key, value, only_if_absent).  Later assignments to the same key win but keep the first position,
    exactly like the dict they are applied to."""
    out = [("model", rule.get("model"), False)]                          # chat.py:113
    if provider_name == "openrouter":                                     # chat.py:114-115
        out.append(("usage", {"include": True}, True))
    for k, v in (rule.get("custom_body_params") or {}).items():           # chat.py:116-119
        out.append((k, v, False))
    out.append(("model", rule.get("model"), False))                       # chat.py:135 / :160 re-assign inside the retry loop: the rule's model always wins
    order = rule.get("providers_order")
    routed = [sub_provider] if sub_provider is not None else (list(order) if order else None)
    if routed is not None:                                                # chat.py:137-139 / :164-168
        out.append(("provider", {"order": routed}, False))
        out.append(("allow_fallbacks", False, False))
    if retry:                                                             # chat.py:150: the log scrub hits the live payload
        out.append(("messages", "<REMOVED>", False))
    return out


def _merge(assignments):
    """Collapse repeated keys the way successive dict assignments do: first position, last value.
    An if-absent assignment followed by a plain one to the same key becomes a plain one."""
    merged: dict = {}
    for k, v, absent_only in assignments:
        if k not in merged:
            merged[k] = (v, absent_only)
        elif not absent_only:
            merged[k] = (v, False)            # the key exists by now whatever the client sent: plain overwrite
        # an if-absent assignment to a key assigned earlier never fires
    return [(k, v, a) for k, (v, a) in merged.items()]


@dataclass
class RulePlans:
    """Plan table compiled from `fallback_rules` (loader.py:150-154 shape: {gateway_model: {"fallback_models": [rule, ...]}})."""
    fallback_rules: dict
    fallback_provider: str = ""
    stream_mode: str = "httpx028"
    index: dict = field(default_factory=dict)
    _plans: list = field(default_factory=list)
    _ops: list = field(default_factory=list)
    _blob: bytearray = field(default_factory=bytearray)

    def __post_init__(self):
        modes = [MODES[self.stream_mode], MODES["json5"]]
        for gw_model, entry in self.fallback_rules.items():
            for ri, rule in enumerate(entry.get("fallback_models", [])):
                prov = rule.get("provider")
                subs = rule.get("providers_order") or []
                for mode in modes:
                    self._add((gw_model, ri, -1, False, mode), attempt_assignments(rule, prov), mode)
                    if rule.get("retry_count", 0):
                        self._add((gw_model, ri, -1, True, mode), attempt_assignments(rule, prov, retry=True), mode)
                    if rule.get("use_provider_order_as_fallback") and subs:
                        for si, sp in enumerate(subs):
                            self._add((gw_model, ri, si, False, mode), attempt_assignments(rule, prov, sub_provider=sp), mode)
        # unknown model -> {"provider": FALLBACK_PROVIDER, "model": requested} (chat.py:52): the model assignment
        # rewrites the value with itself, so the plan only carries what the provider name adds
        for mode in modes:
            extra = [("usage", {"include": True}, True)] if self.fallback_provider == "openrouter" else []
            self._add((None, 0, -1, False, mode), extra, mode)
        # row a12: an upstream non-streaming response is re-rendered the way Starlette's JSONResponse does
        # (same encoder settings as httpx 0.28) and probed for the keys request_handler.py:167 looks for
        begin = len(self._ops)
        for k in RESPONSE_PROBES:
            ko, kl = self._put(k.encode("utf-8"))
            self._ops.append((ko, kl, 0, 0, 0, 0, OP_PRESENCE, 0))
        self.index[("response",)] = len(self._plans)
        self._plans.append((begin, len(self._ops), MODES["httpx028"] | PLAN_RESPONSE, 0))

    def _put(self, b: bytes):
        off = len(self._blob)
        self._blob += b
        return off, len(b)

    def _add(self, key, assignments, mode):
        ops = _merge(assignments)
        if len(ops) > MAX_OPS_PER_PLAN:
            raise ValueError("rule %r assigns more than %d body keys" % (key, MAX_OPS_PER_PLAN))
        begin = len(self._ops)
        for k, v, absent_only in ops:
            ko, kl = self._put(k.encode("utf-8"))
            ro, rl = self._put(render_key(k, mode).encode("utf-8"))
            vo, vl = self._put(render_value(v, mode))
            self._ops.append((ko, kl, ro, rl, vo, vl, 1 if absent_only else 0, 0))
        self.index[key] = len(self._plans)
        self._plans.append((begin, len(self._ops), mode, 0))

    def plan_index(self, gw_model, rule_idx: int = 0, sub_idx: int = -1, retry: bool = False, stream: bool = True) -> int:
        """Plan of one attempt.  `gw_model=None` (or a model without rules) selects the fallback-provider plan."""
        mode = MODES[self.stream_mode] if stream else MODES["json5"]
        if gw_model not in self.fallback_rules:
            return self.index[(None, 0, -1, False, mode)]        # (the synthetic fallback rule has no retry_count: chat.py:52)
        if sub_idx >= 0:
            retry = False          # chat.py:158-181: the sub-provider branch never scrubs `messages`, a retry re-sends the same body
        return self.index[(gw_model, rule_idx, sub_idx, retry, mode)]

    def response_plan(self) -> int:
        return self.index[("response",)]

    def packed(self):
        plans = np.array(self._plans, dtype=PLAN_DTYPE) if self._plans else np.zeros(0, PLAN_DTYPE)
        ops = np.array(self._ops, dtype=OP_DTYPE) if self._ops else np.zeros(0, OP_DTYPE)
        blob = np.frombuffer(bytes(self._blob) or b"\0", dtype=np.uint8).copy()
        return plans, ops, blob

    def max_growth(self) -> int:
        """Upper bound on how many bytes a plan can add to a body (for sizing output slots)."""
        plans, ops, _ = self.packed()
        worst = 0
        for p in plans:
            o = ops[p["op_begin"]:p["op_end"]]
            worst = max(worst, int((o["rkey_len"] + o["rval_len"] + 4).sum()))
        return worst


def pack_bodies(bodies):
    """list[bytes] -> (uint8 buffer, uint64 offsets[n+1])"""
    off = np.zeros(len(bodies) + 1, dtype=np.uint64)
    if bodies:
        off[1:] = np.cumsum([len(b) for b in bodies], dtype=np.uint64)
    buf = np.frombuffer(b"".join(bodies) or b"\0", dtype=np.uint8).copy()
    return buf, off
