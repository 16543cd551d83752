"""Phase cycle counters of the CTA-pair GEMM (needs a -DTCD_GEMM_DEBUG build: python -m tcdiff_b200.build --define TCD_GEMM_DEBUG=1 --out X.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tcdiff_b200 import ops, _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[sys.argv.index("--lib") + 1])
dev = torch.device("cuda:0")
h = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_ulonglong * 8)()
for (M, N, K, outbf, act, bias) in [(96000, 512, 512, True, 0, False), (96000, 1024, 512, True, 0, False), (96000, 1024, 512, True, 2, True),
                                    (96000, 512, 1024, True, 0, True), (96000, 512, 512, False, 0, True)]:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if outbf else torch.float32)
    for _ in range(3):
        ops.gemm(a, w, b, act, out)
    h.tcd_gemm_debug_read(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm(a, w, b, act, out); e1.record()
    h.tcd_gemm_debug_read(buf, 1)
    tiles = buf[5]
    print(f"M{M} N{N} K{K} {'bf16' if outbf else 'f32'} act{act}: {e0.elapsed_time(e1) * 1e3:.1f} us (instrumented); per tile, cycles: "
          f"MMA warp waits accumulator {buf[0] / tiles:.0f}, waits operands {buf[1] / tiles:.0f}, loop {buf[2] / tiles:.0f}; "
          f"epilogue warp waits MMAs {buf[3] / tiles:.0f}, drains {buf[4] / tiles:.0f}; tiles {tiles}")
