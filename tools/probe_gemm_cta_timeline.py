"""Per-CTA in-kernel timeline of cb_gemm on the shapes of the training step (every CTA stamps clock64 at its phase boundaries and
%globaltimer at entry / exit; see dbg_stamp in csrc/gemm.cu). Prints, per launch configuration: device time per launch (50
launches replayed from a CUDA graph), the kernel's span on the global timer, and the median / max over CTAs of every phase.
usage: python tools/probe_gemm_cta_timeline.py [filter-substring]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clipbert_b200 import _lib as L, ops  # noqa: E402

lib = L.lib()
dev = "cuda"
PHASES = [("setup", 0, 1), ("to_first_tma", 1, 2), ("tma_latency", 2, 4), ("tile0_mainloop", 4, 5), ("all_mma", 4, 6), ("tile0_epilogue", 7, 14),
          ("epi_tail_after_last_mma", 6, 8), ("store_drain", 8, 9), ("teardown", 9, 11), ("whole_cta", 0, 11),
          # inside the first tile of the first epilogue warp (warp-private TMA epilogue)
          ("e:inputs+tmem_ld", 7, 16), ("e:pass0_math", 16, 17), ("e:passes1-3", 17, 18), ("e:store_issue", 18, 14)]


def med(v):
    v = sorted(v)
    return v[len(v) // 2] if v else 0


def run(label, **kw):
    buf = torch.zeros(32 * 600, dtype=torch.int64, device=dev)
    for _ in range(3):
        ops.gemm(**kw)
    torch.cuda.synchronize()
    lib.cb_debug_gemm_timeline(ctypes.c_void_p(buf.data_ptr()))
    ops.gemm(**kw)
    torch.cuda.synchronize()
    lib.cb_debug_gemm_timeline(None)
    t = buf.view(-1, 32).cpu()
    t = t[t[:, 0] != 0]
    ncta = t.shape[0]
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
    us = 1e3 * g0.elapsed_time(g1) / 50
    span = (int(t[:, 13].max()) - int(t[:, 12].min())) / 1e3
    skew = (int(t[:, 12].max()) - int(t[:, 12].min())) / 1e3
    cyc = (t[:, 11] - t[:, 0]).double()
    ns = (t[:, 13] - t[:, 12]).double()
    ghz = float((cyc / ns.clamp(min=1)).median())
    parts = []
    for name, a, b in PHASES:
        ok = (t[:, a] != 0) & (t[:, b] != 0)
        d = (t[ok, b] - t[ok, a]).tolist()
        if d:
            parts.append("%s %d/%d" % (name, med(d), max(d)))
    flop = 2.0 * kw["m"] * kw["n"] * kw["k"] * kw.get("ntaps", 1)
    print("%-58s %7.2f us/launch %6.0f TF/s | ctas %3d span %.2f us entry-skew %.2f us clk %.2f GHz | med/max cycles: %s"
          % (label, us, flop / us / 1e6, ncta, span, skew, ghz, ", ".join(parts)), flush=True)


def rnd(*shape):
    return (torch.randn(*shape, device=dev) * 0.5).to(torch.bfloat16)


def cases():
    M = 2624
    x768, x3072 = rnd(M, 768), rnd(M, 3072)
    w_qkv, w_ao, w_in, w_out = rnd(2304, 768), rnd(768, 768), rnd(3072, 768), rnd(768, 3072)
    b768, b2304, b3072 = torch.zeros(768, device=dev), torch.zeros(2304, device=dev), torch.zeros(3072, device=dev)
    o768, o2304, o3072, o3072b = (torch.zeros(M, n, device=dev, dtype=torch.bfloat16) for n in (768, 2304, 3072, 3072))
    yield "qkv_fwd      TN m2624 n2304 k768  bias", dict(mode=0, m=M, n=2304, k=768, a=x768, a_rows=M, a_ld=768, b=w_qkv, b_rows=2304, b_ld=768, shift=b2304, out=o2304, out_ld=2304)
    yield "ffn1_fwd     TN m2624 n3072 k768  gelu+stash", dict(mode=0, m=M, n=3072, k=768, a=x768, a_rows=M, a_ld=768, b=w_in, b_rows=3072, b_ld=768, shift=b3072, out=o3072, out_ld=3072,
                                                        act=ops.ACT_GELU_STASH_GRAD, out2=o3072b, out2_ld=3072)
    yield "attnout_fwd  TN m2624 n768  k768  res+drop", dict(mode=0, m=M, n=768, k=768, a=x768, a_rows=M, a_ld=768, b=w_ao, b_rows=768, b_ld=768, shift=b768, out=o768, out_ld=768,
                                                        residual=x768, res_ld=768, dropout_p=0.1, dropout_seed=5)
    yield "ffn2_fwd     TN m2624 n768  k3072 res+drop", dict(mode=0, m=M, n=768, k=3072, a=x3072, a_rows=M, a_ld=3072, b=w_out, b_rows=768, b_ld=3072, shift=b768, out=o768, out_ld=768,
                                                        residual=x768, res_ld=768, dropout_p=0.1, dropout_seed=5)
    yield "ffn2_dgrad   NN m2624 n3072 k768  aux mul", dict(mode=2, m=M, n=3072, k=768, a=x768, a_rows=M, a_ld=768, b=w_out, b_rows=768, b_ld=3072, out=o3072, out_ld=3072,
                                                       aux=x3072, aux_ld=3072, aux_mode=ops.AUX_MUL)
    gw = torch.zeros(768, 3072, device=dev)
    yield "ffn_wgrad    WG m768 n3072 k2624", dict(mode=1, m=768, n=3072, k=M, a=x768, a_rows=M, a_ld=768, b=x3072, b_rows=M, b_ld=3072, out=gw, out_ld=3072, out_fp32=1)
    Mc = 401408
    xc, wc, rc, oc = rnd(Mc, 64), rnd(256, 64), rnd(Mc, 256), torch.zeros(Mc, 256, device=dev, dtype=torch.bfloat16)
    s256 = torch.zeros(256, device=dev)
    yield "res2_conv3   TN m401408 n256 k64 res+relu", dict(mode=0, m=Mc, n=256, k=64, a=xc, a_rows=Mc, a_ld=64, b=wc, b_rows=256, b_ld=64, shift=s256, out=oc, out_ld=256, residual=rc, res_ld=256,
                                                       act=ops.ACT_RELU)
    M3 = 32768
    x3, w3, o3 = rnd(M3 + 128, 256), rnd(256, 9 * 256), torch.zeros(M3, 256, device=dev, dtype=torch.bfloat16)
    yield "res4_conv2   TN m32768 n256 k256 t9 relu", dict(mode=0, m=M3, n=256, k=256, a=x3, a_rows=M3, a_ld=256, b=w3, b_rows=256, b_ld=9 * 256, shift=s256, out=o3, out_ld=256, ntaps=9, tap_w=18,
                                                      tap_sign=1, act=ops.ACT_RELU)
    yield "res4_conv2dg NN m32768 n256 k256 t9 relumask", dict(mode=2, m=M3, n=256, k=256, a=x3, a_rows=M3, a_ld=256, b=w3, b_rows=256, b_ld=9 * 256, out=o3, out_ld=256, ntaps=9, tap_w=18,
                                                          tap_sign=-1, aux=x3, aux_ld=256, aux_mode=ops.AUX_RELU_MASK)


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    ops.set_occ2(1)
    for name, kw in cases():
        if flt not in name:
            continue
        wg = kw["mode"] == 1
        variants = [("auto", {}), ("bn256 occ1", dict(block_n=256, reserved=2 | 64)), ("bn128 occ1", dict(block_n=128, reserved=2 | 64)),
                    ("bn128 occ2", dict(block_n=128, reserved=2 | 32)), ("bn64 occ2", dict(block_n=64, reserved=2 | 32))]
        for vn, extra in variants:
            try:
                run("%s | %s" % (name, vn), **dict(kw, **extra))
            except Exception as e:
                print("%s | %s: %s" % (name, vn, e), flush=True)


if __name__ == "__main__":
    main()
