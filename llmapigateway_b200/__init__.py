"""llmapigateway_b200 -- B200-native streaming chat-completions transform engine.

The hot path of fabiojbg/LLMApiGateway's /v1/chat/completions (SSE split -> first-event sniff ->
usage extraction -> byte-exact relay), batched across streams and run as sm_100a CUDA kernels
behind a C ABI (include/llmgw_b200.h).  No CPU implementation exists in this package."""
from .engine import Engine, EngineError, StepResult  # noqa: F401
