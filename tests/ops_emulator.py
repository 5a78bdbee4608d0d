"""CPU emulation of the C ABI's *contract* (include/clipbert_b200.h) in plain torch - TEST INFRASTRUCTURE ONLY.

Purpose: run the Python orchestration of clipbert_b200/modeling.py (buffer planning, stash, forward/backward call order,
epilogue flags, index bookkeeping) on a machine without a GPU and compare it with the oracle. It swaps the functions of
clipbert_b200.ops for torch code that does what the header says each entry point does (bf16 buffers in, fp32 arithmetic,
bf16 / fp32 out); nothing in the product imports this file, and the kernels themselves are only ever checked on a B200
(tests/test_gpu_*.py). Dropout must be off (the counter RNG of the kernels is not restated here).
"""
import contextlib
import math

import torch
import torch.nn.functional as F

F32 = torch.float32


def _mat(t, rows, cols, ld):
    """rows x cols window with row pitch ld starting at the first element of tensor t (a raw device pointer in the ABI)."""
    return torch.as_strided(t, (rows, cols), (ld, 1), t.storage_offset())


def _gelu(v):
    return 0.5 * v * (1.0 + torch.erf(v / math.sqrt(2.0)))


def _gelu_grad(v):
    return 0.5 * (1.0 + torch.erf(v / math.sqrt(2.0))) + v * torch.exp(-0.5 * v * v) / math.sqrt(2.0 * math.pi)


def gemm(**kw):
    mode, m, n, k = kw.get("mode", 0), kw["m"], kw["n"], kw["k"]
    assert kw.get("ntaps", 1) == 1 and kw.get("rowmap", 0) == 0, "emulator covers the transformer-side contractions only"
    assert not kw.get("dropout_p"), "emulator: dropout must be off"
    a, b, out = kw["a"], kw["b"], kw["out"]
    if mode == 1:                                       # WGRAD: out[m, n] += sum_p A[p, m] B[p, n]
        A = _mat(a, k, m, kw["a_ld"]).to(F32)
        B = _mat(b, k, n, kw["b_ld"]).to(F32)
        assert out.dtype == F32 and kw.get("out_fp32") == 1
        _mat(out, m, n, kw["out_ld"]).add_(A.t() @ B)
        return
    A = _mat(a, m, k, kw["a_ld"]).to(F32)
    if mode == 0:                                       # TN: B [n, k]
        v = A @ _mat(b, n, k, kw["b_ld"]).to(F32).t()
    else:                                               # NN: B [k, n] (the forward weight read MN-major)
        v = A @ _mat(b, k, n, kw["b_ld"]).to(F32)
    if kw.get("scale") is not None:
        v = v * kw["scale"][:n].to(F32)
    if kw.get("shift") is not None:
        v = v + kw["shift"][:n].to(F32)
    if kw.get("residual") is not None:
        v = v + _mat(kw["residual"], m, n, kw["res_ld"]).to(F32)
    act = kw.get("act", 0)
    if kw.get("out2") is not None:
        _mat(kw["out2"], m, n, kw["out2_ld"]).copy_(_gelu_grad(v) if act == 4 else v)
    if act == 1:
        v = torch.relu(v)
    elif act in (2, 4):
        v = _gelu(v)
    elif act == 3:
        v = torch.tanh(v)
    if kw.get("aux") is not None:
        x = _mat(kw["aux"], m, n, kw["aux_ld"]).to(F32)
        am = kw.get("aux_mode", 0)
        if am == 1:
            v = v * (x > 0).to(F32)
        elif am == 2:
            v = v * _gelu_grad(x)
        elif am == 3:
            v = v * (1.0 - x * x)
        elif am == 4:
            v = v * x
    assert (out.dtype == F32) == bool(kw.get("out_fp32", 0))
    _mat(out, m, n, kw["out_ld"]).copy_(v)


def _ln_fwd(v, gamma, beta, eps):
    mean = v.mean(-1, keepdim=True)
    var = ((v - mean) ** 2).mean(-1, keepdim=True)
    rstd = torch.rsqrt(var + eps)
    return (v - mean) * rstd * gamma + beta, mean.squeeze(-1), rstd.squeeze(-1)


def _ln_bwd(dy, x, mean, rstd, gamma):
    xhat = (x - mean[:, None]) * rstd[:, None]
    g = dy * gamma
    dx = rstd[:, None] * (g - g.mean(-1, keepdim=True) - xhat * (g * xhat).mean(-1, keepdim=True))
    return dx, (dy * xhat).sum(0), dy.sum(0)


def layernorm_fwd(x, gamma, beta, y, stats, eps):
    o, mean, rstd = _ln_fwd(x.to(F32), gamma, beta, eps)
    y.copy_(o)
    stats[:, 0], stats[:, 1] = mean, rstd


def layernorm_bwd(dy, x, stats, gamma, dx, dx_drop, dgamma, dbeta, dbias_drop, p, seed):
    assert not p
    d, dg, db = _ln_bwd(dy.to(F32), x.to(F32), stats[:, 0], stats[:, 1], gamma)
    dx.copy_(d)
    if dx_drop is not None:
        dx_drop.copy_(d)
    dgamma.add_(dg)
    dbeta.add_(db)
    if dbias_drop is not None:
        dbias_drop[: d.shape[1]].add_(dx.to(F32).sum(0))     # column sums of the bf16 tensor the dense's dgrad consumes


def embed_text_fwd(ids, word, pos, typ, gamma, beta, out, stats, nseq, lt, l, eps, p, seed):
    assert not p
    v = word[ids] + pos[:lt][None] + typ[0][None, None]
    o, mean, rstd = _ln_fwd(v.reshape(nseq * lt, -1), gamma, beta, eps)
    out.view(nseq, l, -1)[:, :lt].copy_(o.view(nseq, lt, -1))
    stats[:, 0], stats[:, 1] = mean, rstd


def embed_text_bwd(dh, ids, word, pos, typ, gamma, stats, dword, dpos, dtyp, dgamma, dbeta, nseq, lt, l, p, seed):
    assert not p
    h = word.shape[1]
    v = (word[ids] + pos[:lt][None] + typ[0][None, None]).reshape(nseq * lt, h)
    dy = dh.view(nseq, l, h)[:, :lt].reshape(nseq * lt, h).to(F32)
    d, dg, db = _ln_bwd(dy, v, stats[:, 0], stats[:, 1], gamma)
    dgamma.add_(dg)
    dbeta.add_(db)
    dword.index_add_(0, ids.reshape(-1), d)
    dpos[:lt].add_(d.view(nseq, lt, h).sum(0))
    dtyp[0].add_(d.sum(0))


def _visual_pre(grid, seq2vid, n_ex, rowemb, colemb, typ, nseq, t, gh, gw):
    h = grid.shape[-1]
    vid = seq2vid.long() if seq2vid is not None else torch.arange(nseq) // n_ex
    g = grid.reshape(-1, t, gh * gw, h).to(F32).sum(1) * (1.0 / t)            # [nvid, Lv, h]
    j = torch.arange(gh * gw)
    v = g[vid] + rowemb[j // gw][None] + colemb[j % gw][None] + typ[0][None, None]
    return v.reshape(nseq * gh * gw, h), vid, j


def embed_visual_fwd(grid, seq2vid, n_ex, rowemb, colemb, typ, gamma, beta, out, stats, nseq, t, gh, gw, lt, l, eps, p, seed):
    assert not p
    v, _, _ = _visual_pre(grid, seq2vid, n_ex, rowemb, colemb, typ, nseq, t, gh, gw)
    o, mean, rstd = _ln_fwd(v, gamma, beta, eps)
    out.view(nseq, l, -1)[:, lt:].copy_(o.view(nseq, gh * gw, -1))
    stats[:, 0], stats[:, 1] = mean, rstd


def embed_visual_bwd(dh, grid, seq2vid, vid_start, n_ex, rowemb, colemb, typ, gamma, stats, dv_tmp, dgrid, drow, dcol, dtyp,
                     dgamma, dbeta, nseq, nvid, t, gh, gw, lt, l, p, seed):
    assert not p
    h = grid.shape[-1]
    lv = gh * gw
    v, vid, j = _visual_pre(grid, seq2vid, n_ex, rowemb, colemb, typ, nseq, t, gh, gw)
    dy = dh.view(nseq, l, h)[:, lt:].reshape(nseq * lv, h).to(F32)
    d, dg, db = _ln_bwd(dy, v, stats[:, 0], stats[:, 1], gamma)
    dgamma.add_(dg)
    dbeta.add_(db)
    dv_tmp.copy_(d)
    d3 = d.view(nseq, lv, h)
    drow.index_add_(0, j // gw, d3.sum(0))
    dcol.index_add_(0, j % gw, d3.sum(0))
    dtyp[0].add_(d.sum(0))
    if dgrid is not None:
        per_vid = torch.zeros(nvid, lv, h).index_add_(0, vid, d3) * (1.0 / t)
        dgrid.view(nvid, t, lv, h).copy_(per_vid[:, None].expand(nvid, t, lv, h))


def _attention(qkv, text_mask, nseq, l, lt, heads):
    hd = qkv.shape[1] // (3 * heads)
    q, k, v = (x.reshape(nseq, l, heads, hd).permute(0, 2, 1, 3) for x in qkv.view(nseq, l, 3, heads * hd).unbind(2))
    mask = torch.cat([text_mask.to(F32), torch.ones(nseq, l - lt)], dim=1)
    s = q @ k.transpose(-1, -2) / math.sqrt(hd) + ((1.0 - mask) * -10000.0)[:, None, None, :]
    pr = torch.softmax(s, dim=-1)
    return (pr @ v).permute(0, 2, 1, 3).reshape(nseq * l, heads * hd), torch.logsumexp(s, dim=-1)


def attention_fwd(qkv, text_mask, ctx, lse, nseq, l, lt, heads, p, seed):
    assert not p
    o, ls = _attention(qkv.to(F32), text_mask, nseq, l, lt, heads)
    ctx.copy_(o)
    if lse is not None:
        lse.copy_(ls)


def attention_bwd(qkv, text_mask, ctx, dctx, lse, dqkv, nseq, l, lt, heads, p, seed):
    assert not p
    x = qkv.to(F32).clone().requires_grad_(True)
    with torch.enable_grad():
        o, _ = _attention(x, text_mask, nseq, l, lt, heads)
        o.backward(dctx.to(F32))
    dqkv.copy_(x.grad)


def colsum(x, out, m, n, ld=None):
    out[:n].add_(_mat(x, m, n, n if ld is None else ld).to(F32).sum(0))


def dropout(x, y, p, seed):
    assert not p
    y.copy_(x)


def gelu_bwd(dy, u, dx):
    dx.copy_(dy.to(F32) * _gelu_grad(u.to(F32)))


def pad_cast(src, dst):
    dst.zero_()
    dst[:, : src.shape[1]].copy_(src)


def cast_scale(src, dst, rowscale=None, row_len=1):
    v = src if rowscale is None else (src.view(-1, row_len) * rowscale[:, None]).reshape(-1)
    dst.copy_(v)


_NAMES = ("gemm", "layernorm_fwd", "layernorm_bwd", "embed_text_fwd", "embed_text_bwd", "embed_visual_fwd", "embed_visual_bwd",
          "attention_fwd", "attention_bwd", "colsum", "dropout", "gelu_bwd", "pad_cast", "cast_scale")


@contextlib.contextmanager
def emulated_transformer_ops():
    """Swap the transformer-side wrappers of clipbert_b200.ops (and the device check of modeling.py) for the torch code above."""
    from clipbert_b200 import modeling, ops
    saved = {n: getattr(ops, n) for n in _NAMES}
    saved_overlap, saved_req = ops.overlap_wgrad, modeling._require_cuda
    calls = {n: 0 for n in _NAMES}

    def counted(name, fn):
        def f(*a, **k):
            calls[name] += 1
            return fn(*a, **k)
        return f
    try:
        for n in _NAMES:
            setattr(ops, n, counted(n, globals()[n]))
        ops.overlap_wgrad = False
        modeling._require_cuda = lambda t: None
        yield calls
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
        ops.overlap_wgrad, modeling._require_cuda = saved_overlap, saved_req
