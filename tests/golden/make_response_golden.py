"""Generate tests/golden/response_cases.json from the UNMODIFIED reference (dev container only):
    python tests/golden/make_response_golden.py
Each case: upstream status + body bytes -> what make_llm_request(is_streaming=False) + chat.py:146 +
Starlette's JSONResponse produce."""
import base64
import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE)); sys.path.insert(0, str(HERE.parent.parent))
import ref_driver  # noqa: E402

OK_DOC = {"id": "chatcmpl-1", "object": "chat.completion", "created": 1700000000, "model": "m-1",
          "choices": [{"index": 0, "message": {"role": "assistant", "content": "Hé \"there\"\n\t中 \U0001F600 \x7f /"}, "finish_reason": "stop", "logprobs": None}],
          "usage": {"prompt_tokens": 12, "completion_tokens": 7, "total_tokens": 19, "cost": 1.50e-4, "completion_tokens_details": {"reasoning_tokens": 0}},
          "system_fingerprint": None, "x": [1e16, 1e15, 0.1, -0.0, 2.5e-7, 100.0, 12345678901234567890, True]}

CASES = [
    (200, json.dumps(OK_DOC).encode()),
    (200, json.dumps(OK_DOC, ensure_ascii=False, indent=2).encode()),
    (200, json.dumps(OK_DOC, separators=(",", ":")).encode()),
    (201, b'{"a":1}'),
    (200, b'{ "cost" : 1.50e-4 , "k" : "\\u00e9\\ud83d\\ude00\\/" }'),
    (200, b'{}'),
    (200, b'{"error":{"message":"boom","code":429}}'),
    (200, b'{"error":{"code":429},"detail":"fallback detail"}'),
    (200, b'{"error":"a string"}'),
    (200, b'{"detail":"Not found"}'),
    (200, b'{"detail":null,"a":1}'),
    (200, b'{"choices":[],"error":null}'),
    (200, b'{"nested":{"error":1,"detail":2},"ok":true}'),
    (400, b'{"error":{"message":"bad request"}}'),
    (500, b'upstream exploded \xff'),
    (404, b''),
    (200, b'not json'),
    (200, b''),
    (200, b'{"a":1,}'),
    (200, b'{"a":NaN}'),
    (200, b'{"a":"\\ud800"}'),
    (200, b'{"detail":""}'),
    (200, b'{"detail":0,"error":{}}'),
    (200, b'{"error":{"message":""},"detail":"second choice"}'),
    (200, b'{"error":[1,2]}'),
    (202, b' {"id" :"r1", "choices":[{"message":{"role":"assistant","content":null,"tool_calls":[{"id":"c","type":"function","function":{"name":"f","arguments":"{\\"a\\": 1}"}}]},"finish_reason":"tool_calls"}],"usage":{"prompt_tokens":5,"completion_tokens":0,"total_tokens":5,"prompt_tokens_details":{"cached_tokens":0}}} '),
    (200, b'[{"error":1}]'),
    (200, b'["error"]'),
    (200, b'"an error string"'),
    (200, b'12'),
    (200, b'null'),
    (200, b'{"big":123456789012345678901234567890,"neg":-0,"f":-0.0,"e":1E+2,"tiny":5e-324}'),
]


# texts handed straight to the response tap in its non-streaming mode (chat_logging.py:98-150 with is_real_streaming=False):
# what the client of a non-streaming request receives (the compact body above), error bodies of the endpoint, odd shapes
TAP_TEXTS = [
    b'{"id":"r","choices":[{"index":0,"message":{"role":"assistant","content":"hi"}}],"usage":{"prompt_tokens":10,"completion_tokens":7,"total_tokens":17,"cost":0.00123,'
    b'"completion_tokens_details":{"reasoning_tokens":2},"prompt_tokens_details":{"cached_tokens":4}},"model":"m-ok","provider":"P"}',
    b'{"choices":[{"message":{"content":"x"}}],"usage":{"prompt_tokens":3,"completion_tokens":2,"total_tokens":5}}',
    b'{"choices":[{"message":{"content":"x"}}]}',
    b'{"detail":"All configured providers failed for model \'m\'. Last error: boom"}',
    b'{"detail":"Missing \'model\' in request body."}',
    b'{"usage":null,"model":"m"}',
    b'{"usage":{"prompt_tokens":1,"completion_tokens_details":null},"model":"m"}',
    b'{"usage":{"prompt_tokens":1.5,"completion_tokens":9,"completion_tokens_details":{"reasoning_tokens":3}},"provider":"Q"}',
    b'{"choices":[{"delta":null}],"usage":{"prompt_tokens":1}}',
    b'{"choices":"abc","usage":{"prompt_tokens":1}}',
    b'{"error":{"message":"late"},"usage":{"prompt_tokens":4}}',
    b'{"usage":{"prompt_tokens":4},"error":null}',
    b'data: {"usage":{"prompt_tokens":8},"model":"via-data-prefix"}  ',
    b' {"usage":{"prompt_tokens":8}}',
    b'[{"usage":{"prompt_tokens":8}}]',
    b'{"usage":{"prompt_tokens":8}}\n\n{"usage":{"prompt_tokens":9}}',
    b'not json at all',
    b'',
    b'{"usage":{"prompt_tokens":2,"cost":1e-7},"model":"caf\xc3\xa9","provider":"\\u00e9"}',
]


def main():
    out = []
    for status, content in CASES:
        r = ref_driver.run_nonstream(content, status)
        case = {"status": status, "content": base64.b64encode(content).decode(), "kind": r["kind"],
                "detail": r["detail"], "body": base64.b64encode(r["body"]).decode()}
        if r["kind"] == "ok":                                       # the tap sees what the client receives
            case["tap_rows"] = ref_driver.run_tap([r["body"]], is_real_streaming=False)[0]
        out.append(case)
    tap = []
    for text in TAP_TEXTS:
        rows = ref_driver.run_tap([text], is_real_streaming=False)[0] if text else []
        tap.append({"text": base64.b64encode(text).decode(), "rows": rows})
    # split deliveries: the tap concatenates the chunks first (:98-103)
    whole = TAP_TEXTS[0]
    tap.append({"text": base64.b64encode(whole).decode(), "chunks": [base64.b64encode(whole[:37]).decode(), base64.b64encode(whole[37:]).decode()],
                "rows": ref_driver.run_tap([whole[:37], whole[37:]], is_real_streaming=False)[0]})
    import httpx
    import starlette
    doc = {"generator": "tests/golden/make_response_golden.py", "httpx": httpx.__version__, "starlette": starlette.__version__, "cases": out, "tap_cases": tap}
    (HERE / "response_cases.json").write_text(json.dumps(doc, indent=0))
    print("wrote", len(out), "cases:", [c["kind"] for c in out])


if __name__ == "__main__":
    main()
