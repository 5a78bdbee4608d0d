"""Timeline probe of the staged-epilogue (row re-map) GEMM shapes of the ResNet stages."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clipbert_b200 import _lib as L, ops
lib = L.lib(); dev = "cuda"
NAMES = ["entry", "setup_done", "tma_first", "tma_all", "mma_first_full", "mma_tile0", "mma_all", "epi_acc_ready", "epi_done", "store_drained", "all_synced", "exit"]
def rnd(*s): return torch.randn(*s, device=dev).to(torch.bfloat16)
def run(label, **kw):
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    lib.cb_debug_gemm_timeline(ctypes.c_void_p(buf.data_ptr())); ops.gemm(**kw); torch.cuda.synchronize(); lib.cb_debug_gemm_timeline(None)
    t = buf.cpu().tolist(); base = t[0]
    stamps = " ".join("%s=%d" % (n, t[i] - base) for i, n in enumerate(NAMES) if t[i] and i in (3, 5, 6, 7, 8, 11))
    for _ in range(3): ops.gemm(**kw)
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(20): ops.gemm(**kw)
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    fl = 2.0 * kw["m"] * kw["n"] * kw["k"] * kw.get("ntaps", 1)
    us = 1e3 * e0.elapsed_time(e1) / 20
    print("%-52s %7.2f us %6.0f TF/s | %s" % (label, us, fl / us / 1e6, stamps), flush=True)
NB = 64
for (H, C, N, taps, mode, rowmap, lab) in [(28, 128, 128, 9, 0, 2, "res3 conv2 fwd"), (28, 128, 128, 9, 2, 2, "res3 conv2 dgrad"),
                                             (28, 512, 128, 1, 2, 1, "res3 conv3 dgrad (PAD)"), (28, 512, 128, 1, 0, 1, "res3 conv1 fwd (PAD)"),
                                             (14, 256, 256, 9, 0, 2, "res4 conv2 fwd"), (14, 256, 256, 9, 2, 2, "res4 conv2 dgrad"),
                                             (7, 2048, 768, 9, 0, 2, "grid_encoder fwd"), (7, 768, 2048, 9, 2, 2, "grid_encoder dgrad")]:
    P = NB * (H + 2) * (H + 2); R = NB * H * H
    M = P if taps == 9 or rowmap == 2 else R
    K = C
    A = rnd(M + 128, K)
    Bm = rnd(N, K * taps) if mode == 0 else rnd(K, N * taps)
    out_rows = R if rowmap == 2 else P
    out = torch.zeros(out_rows, N, device=dev, dtype=torch.bfloat16)
    aux = rnd(M, N)
    shift = torch.randn(N, device=dev)
    kw = dict(mode=mode, m=M, n=N, k=K, a=A, a_rows=M, a_ld=K, b=Bm, b_rows=Bm.shape[0], b_ld=Bm.shape[1], ntaps=taps, tap_w=H + 2,
              tap_sign=1 if mode == 0 else -1, out=out, out_ld=N, rowmap=rowmap, map_h=H, map_w=H)
    if mode == 0: kw.update(shift=shift, act=1)
    else: kw.update(aux=aux, aux_ld=N, aux_mode=1)
    for bn in (64, 128, 256):
        if bn > N: continue
        for kch in (1, 2):
            lib.cb_debug_gemm_kch(kch)
            try: run("%s bn%d kch%d" % (lab, bn, kch), block_n=bn, **kw)
            except Exception as e: print(lab, bn, kch, "ERR", str(e)[:80])
    lib.cb_debug_gemm_kch(0)
