"""Generates tests/golden/write_log_cases.json: the log files the UNMODIFIED reference `write_log`
(llm_gateway_core/middleware/chat_logging.py:22-67) writes for a handful of transcripts.  Dev-container only
(reads /root/reference); the committed JSON is what the tests use.

    python tests/golden/make_write_log_golden.py
"""
import json
import os
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).parent))
import ref_driver

U = {"prompt_tokens": 10, "completion_tokens": 5, "total_tokens": 17, "reasoning_tokens": 2, "cached_tokens": 4, "cost": 0.00123, "model": "m-ok", "provider": "P"}
U0 = {"prompt_tokens": 0, "completion_tokens": 0, "total_tokens": 0, "reasoning_tokens": 0, "cached_tokens": 0, "cost": 0}
CASES = [
    ("plain", {"host": "gw", "content-type": "application/json"}, '{"model":"m","messages":[{"role":"user","content":"hi"}]}', "hello world", U),
    ("defaults_no_model", {}, "{}", "", U0),
    ("escaped_newlines_in_body_and_text", {"a": "b"}, '{"messages":[{"content":"line1\\nline2\\n\\nline3"}]}', "text with \\n literal and \\n\\n double and a real\nnewline", U),
    ("unicode", {"x-title": "é"}, '{"k":"中"}', "é 中 😀 \x00 \t", U),
    ("error_event_text", {}, "{}", 'partial{"error":{"message":"boom \\n x"},"code":500}', U0),
    ("many_headers", {("h%02d" % i): "v" * 30 for i in range(8)}, "{}", "ok", U),
    ("lone_surrogate", {}, "{}", "bad \ud83d tail", U),
    ("model_only", {}, "{}", "t", {**U0, "model": "only-model"}),
    ("float_cost_int_tokens", {}, "{}", "t", {**U0, "cost": 1, "prompt_tokens": 3.5}),
]


def main():
    _, cl = ref_driver.load_reference()
    out = []
    cwd = os.getcwd()
    for name, headers, body, accum, usage in CASES:
        rows = []
        with tempfile.TemporaryDirectory() as d:
            os.chdir(d)
            real = cl.tokens_usage_db.insert_usage
            cl.tokens_usage_db.insert_usage = lambda u: rows.append(dict(u))
            try:
                cl.write_log(headers, body, accum, usage)
            finally:
                cl.tokens_usage_db.insert_usage = real
                os.chdir(cwd)
            files = sorted(Path(d, "logs").glob("*.txt")) if Path(d, "logs").exists() else []
            content = None
            if rows:                                   # the row is inserted only after a successful write (:47-56)
                assert len(files) == 1
                content = files[0].read_bytes().decode("utf-8")
        out.append(dict(name=name, headers=headers, body=body, accum=accum, usage=usage, file=content, rows=rows))
    doc = dict(generator="tests/golden/make_write_log_golden.py", reference="llm_gateway_core/middleware/chat_logging.py:22-67 (unmodified)", cases=out)
    Path(__file__).with_name("write_log_cases.json").write_text(json.dumps(doc, indent=1, ensure_ascii=True))
    print(len(out), "cases;", sum(c["file"] is None for c in out), "without a file")


if __name__ == "__main__":
    main()
