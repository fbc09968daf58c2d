"""AMC_TC5_TRACE=1 python tools/probes/attn_tc5_trace.py B T h dh [--bwd]: one launch, CTA-0 phase stamps on stderr."""
import sys
sys.argv, args = sys.argv[:1], sys.argv[1:]
from attn_tc5_check import timeit  # noqa: E402
B, T, h, dh = map(int, args[:4])
print(timeit(B, T, h, dh, "--bwd" in args, iters=1))
