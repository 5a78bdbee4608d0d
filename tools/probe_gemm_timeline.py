"""In-kernel clock64 timeline of CTA 0 of cb_gemm + back-to-back launch time, for a few small shapes."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clipbert_b200 import _lib as L, ops  # noqa: E402

lib = L.lib()
dev = "cuda"
NAMES = ["entry", "setup_done", "tma_first_issued", "tma_all_issued", "mma_first_full", "mma_tile0_committed", "mma_all_issued",
         "epi_acc_ready", "epi_done", "epi_store_drained", "all_synced", "exit"]


def run(label, **kw):
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    lib.cb_debug_gemm_timeline(ctypes.c_void_p(buf.data_ptr()))
    ops.gemm(**kw)
    torch.cuda.synchronize()
    lib.cb_debug_gemm_timeline(None)
    t = buf.cpu().tolist()
    base = t[0]
    stamps = " ".join("%s=%d" % (n, t[i] - base) for i, n in enumerate(NAMES) if t[i])
    for _ in range(5):
        ops.gemm(**kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record()
    for _ in range(n):
        ops.gemm(**kw)
    e1.record()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ops.gemm(**kw)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(50):
                ops.gemm(**kw)
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    g.replay()
    g1.record()
    torch.cuda.synchronize()
    print("%-44s stream %.2f us/launch | graph %.2f us/launch | cycles: %s" % (label, 1e3 * e0.elapsed_time(e1) / n, 1e3 * g0.elapsed_time(g1) / 50, stamps), flush=True)


def rnd(*shape):
    return torch.randn(*shape, device=dev).to(torch.bfloat16)


import sys as _sys
for (M, N, K, taps) in [(1312, 3072, 768, 1), (12544, 1024, 256, 1), (16384, 256, 256, 9), (5184, 2048, 768, 9), (8192, 8192, 2048, 1)]:
    A, B = rnd(M + 64, K), rnd(N, K * taps)
    C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    base = dict(mode=0, m=M, n=N, k=K, a=A, a_rows=M, a_ld=K, b=B, b_rows=N, b_ld=K * taps, out=C, out_ld=N, ntaps=taps, tap_w=30, tap_sign=1)
    for bn, knob, lab in [(128, 2, "single bn128"), (256, 2, "single bn256"), (128, 4, "pair bn128"), (256, 4, "pair bn256")]:
        run("TN %dx%dx%d t%d %s" % (M, N, K, taps, lab), block_n=bn, reserved=knob, **base)
