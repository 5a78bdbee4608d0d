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


for (M, N, K) in [(32, 768, 768), (1312, 768, 768), (1312, 3072, 768), (1312, 768, 3072), (12544, 1024, 256)]:
    A, B = rnd(M, K), rnd(N, K)
    C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    base = dict(mode=0, m=M, n=N, k=K, a=A, a_rows=M, a_ld=K, b=B, b_rows=N, b_ld=K, out=C, out_ld=N)
    run("TN %dx%dx%d tma-epi" % (M, N, K), **base)
    run("TN %dx%dx%d staged-epi" % (M, N, K), reserved=1, **base)
    R = rnd(M, N)
    run("TN %dx%dx%d tma-epi +res+relu" % (M, N, K), residual=R, res_ld=N, act=1, **base)
P, Mo, No = 1312, 768, 3072
dY, X = rnd(P, Mo), rnd(P, No)
dW = torch.zeros(Mo, No, device=dev)
for sk in (1, 3):
    run("WGRAD %dx%dx%d split %d" % (Mo, No, P, sk), mode=1, m=Mo, n=No, k=P, a=dY, a_rows=P, a_ld=Mo, b=X, b_rows=P, b_ld=No, split_k=sk,
        out=dW, out_ld=No, out_fp32=1)
# reference point: an empty-ish kernel of our own (LayerNorm over 32 rows)
x = rnd(32, 768)
y = torch.empty_like(x)
st = torch.empty(32, 2, device=dev)
g = torch.ones(768, device=dev)
for _ in range(5):
    ops.layernorm_fwd(x, g, g, y, st, 1e-12)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    ops.layernorm_fwd(x, g, g, y, st, 1e-12)
e1.record()
torch.cuda.synchronize()
print("layernorm 32 rows: stream %.2f us/launch" % (1e3 * e0.elapsed_time(e1) / 200))
