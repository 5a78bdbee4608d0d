"""Thin Python wrappers over the C ABI (include/clipbert_b200.h). Each function takes CUDA torch
tensors (used only as device buffers), validates the few things the C side cannot see (dtype,
contiguity, device) and enqueues the kernel on the current torch stream. No arithmetic happens in
Python or in torch on this path.
"""
import ctypes

import torch

from . import _lib as L
from ._lib import (ACT_GELU, ACT_GELU_STASH_GRAD, ACT_NONE, ACT_RELU, ACT_TANH, AUX_GELU_GRAD, AUX_MUL, AUX_NONE,  # noqa: F401
                   AUX_RELU_MASK, AUX_TANH_GRAD, CB_GEMM_TN, CB_GEMM_WGRAD, ROWMAP_NONE, ROWMAP_PAD, ROWMAP_UNPAD)

CB_GEMM_NN = 2
_c = ctypes
_vp, _i, _i64, _f, _u64 = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float, _c.c_uint64

_SIGS = {
    "cb_layernorm_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp],
    "cb_layernorm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _u64, _vp],
    "cb_embed_text_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _u64, _vp],
    "cb_embed_text_bwd": [_vp] * 12 + [_i, _i, _i, _i, _i, _f, _u64, _vp],
    "cb_embed_visual_fwd": [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _f, _u64, _vp],
    "cb_embed_visual_bwd": [_vp, _vp, _vp, _vp, _i] + [_vp] * 12 + [_i, _i, _i, _i, _i, _i, _i, _i, _f, _u64, _vp],
    "cb_colsum": [_vp, _i64, _vp, _i, _i, _vp],
    "cb_dropout": [_vp, _vp, _i64, _f, _u64, _vp],
    "cb_dropout_offset_advance": [_vp, _vp, _vp],
    "cb_gelu_bwd": [_vp, _vp, _vp, _i64, _vp],
    "cb_pad_cast": [_vp, _i64, _vp, _i, _i, _i, _vp],
    "cb_cast_scale": [_vp, _vp, _i64, _vp, _i64, _vp],
    "cb_cast_bf16_f32": [_vp, _vp, _i64, _vp],
    "cb_cast_scale_segments": [_vp, _vp, _vp, _i, _vp, _vp],
    "cb_nvls_allreduce_f32": [_vp, _i64, _i, _i, _f, _i, _vp],
    "cb_clip_lse_loss": [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp],
    "cb_clip_pool_ce_loss": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp],
    "cb_cross_entropy_fwd": [_vp, _i64, _vp, _vp, _vp, _i64, _i, _i64, _vp],
    "cb_cross_entropy_bwd": [_vp, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _i, _i64, _vp],
    "cb_attention_fwd": [_vp, _i64, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _i, _f, _u64, _vp],
    "cb_attention_bwd": [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _f, _u64, _vp],
    "cb_stem_im2col": [_vp, _i, _vp, _i, _i, _i, _i, _f, _f, _f, _vp],
    "cb_maxpool3x3s2": [_vp, _vp, _i, _i, _i, _i, _vp],
    "cb_maxpool3x3s2_strided": [_vp, _vp, _i, _i, _i, _i, _i64, _i64, _vp],
    "cb_stem_s2d": [_vp, _i, _vp, _i, _i, _i, _i, _f, _f, _f, _vp],
    "cb_resize_pad": [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "cb_subsample2": [_vp, _vp, _i, _i, _i, _i, _vp],
    "cb_unsubsample2_mask": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "cb_maxpool2x2_relu_fwd": [_vp, _vp, _i, _i, _i, _i, _vp],
    "cb_maxpool2x2_relu_bwd": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "cb_relu_mask": [_vp, _vp, _vp, _i64, _vp],
}
_bound = {}


def _fn(name):
    f = _bound.get(name)
    if f is None:
        f = getattr(L.lib(), name)
        f.argtypes = _SIGS[name]
        f.restype = _c.c_int
        _bound[name] = f
    return f


def _p(t):
    return None if t is None else t.data_ptr()


def _s():
    return torch.cuda.current_stream().cuda_stream


_op_timing = None


def set_op_timing(events):
    """Profiling hook: when ``events`` is a list, EVERY kernel launch made through this module is bracketed by
    CUDA events on the launch stream and (label, start, end) is appended. None disables."""
    global _op_timing
    _op_timing = events


def _call(name, *args):
    if _op_timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = _fn(name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, L.lib().cb_last_error().decode()))
    if _op_timing is not None:
        e1.record()
        _op_timing.append((name, e0, e1))


def launch_count():
    return int(L.lib().cb_launch_count())


def set_attention_flash(on):
    """Forward attention of sequences longer than 64 tokens: 1 (default) = tensor-core online-softmax kernel (2.1x on config 5,
    profiles/r02_ab_runs.txt), 0 = the CUDA-core kernel."""
    L.lib().cb_debug_attention_flash(int(bool(on)))


def set_attention_flash_pipe(on):
    """Long-sequence attention forward: 1 (default) = key / value tiles double-buffered through cp.async, 0 = synchronous loads (A/B)."""
    L.lib().cb_debug_attention_flash_pipe(int(bool(on)))


def set_attention_rows48(on):
    """Attention of sequences of up to 48 tokens (L = 41: every 224-px configuration): 1 (default) = 48-row tiles, three warps per
    (sequence, head); 0 = the 64-row / four-warp kernels (A/B)."""
    L.lib().cb_debug_attention_rows48(int(bool(on)))


def set_mn3d(on):
    """MN-major GEMM operands (B of the dgrad mode, both operands of the wgrad mode) through ONE 3-D TMA box per k-chunk
    (default) instead of BN/64 2-D boxes (the single producer lane issues 2 instead of 5-6 TMA instructions per chunk)."""
    L.lib().cb_debug_gemm_mn3d(int(bool(on)))


def set_occ2(mode, max_gflop=0.0):
    """Two GEMM CTAs per SM (128 x <=128 tiles, 8 epilogue warps, <= 113 KB smem each): 0 = never, 1 = only launches that ask
    for it (tuning table / reserved bit 5), 2 = every eligible launch of at most ``max_gflop`` GFLOP (0 = no limit)."""
    f = L.lib().cb_debug_gemm_occ2
    f.argtypes = [ctypes.c_int, ctypes.c_double]
    f.restype = None
    f(int(mode), float(max_gflop))


# ------------------------------------------------------------------------------------------------
# dropout stream position on the device (fresh masks under CUDA-graph replay)
# ------------------------------------------------------------------------------------------------
def dropout_offset_bind(word):
    """Launches made after this call add the uint64 device word ``word`` (a 1-element int64 tensor; None unbinds), read when
    they RUN, into their dropout seed - see cb_dropout_offset_bind."""
    f = L.lib().cb_dropout_offset_bind
    f.argtypes = [_vp]
    f.restype = _c.c_int
    f(None if word is None else word.data_ptr())


def dropout_offset_advance(counter, snapshot=None):
    """++counter on the device (stream-ordered, graph-capturable); snapshot <- the new value."""
    _call("cb_dropout_offset_advance", _p(counter), _p(snapshot), _s())


def set_sm_limit(n):
    """Tuning hook: cap the persistent GEMM grid at ``n`` CTAs (0 = every SM): leaves SMs to a co-resident NCCL kernel so
    that the static tile schedule does not spill into a second wave while a gradient all-reduce overlaps the backward."""
    L.lib().cb_debug_gemm_sm_limit(int(n))


def set_pdl(enable):
    """Programmatic dependent launch between the library's kernels: 0 off (default), 1 every kernel, 2 every kernel except the
    persistent GEMMs. Returns the previous setting."""
    return int(L.lib().cb_set_pdl(2 if int(enable) == 2 else int(bool(enable))))


# ------------------------------------------------------------------------------------------------
# side queue: weight-gradient work off the backward critical path
# ------------------------------------------------------------------------------------------------
class SideQueue:
    """Runs launches that nothing downstream in the backward pass waits for (wgrad GEMMs, bias column sums) on a second
    CUDA stream, forked from / joined back into the current stream with events, so that they fill SMs the dgrad chain
    leaves idle (grids below 148 CTAs, tails, the LayerNorm / attention kernels between GEMMs). Works under CUDA-graph
    capture (fork/join become graph edges). Operands are kept referenced until ``join()`` so the caching allocator cannot
    hand their memory to a later launch on the main stream while the side stream still reads them."""
    _streams = {}

    def __init__(self, enabled=True):
        self.enabled = bool(enabled) and overlap_wgrad
        self.keep = []
        self.side = None
        self.forked = False

    def _side_stream(self):
        dev = torch.cuda.current_device()
        st = SideQueue._streams.get(dev)
        if st is None:
            st = SideQueue._streams[dev] = torch.cuda.Stream(device=dev)
        return st

    def run(self, fn, *keep):
        if not self.enabled:
            fn()
            return
        if self.side is None:
            self.side = self._side_stream()
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            fn()
        self.keep.extend(keep)
        self.forked = True

    def join(self):
        if self.forked:
            torch.cuda.current_stream().wait_stream(self.side)
            self.forked = False
        self.keep.clear()


overlap_wgrad = True      # module switch (bench --overlap_wgrad 0 / tests flip it)


# ------------------------------------------------------------------------------------------------
# tensor-core contraction
# ------------------------------------------------------------------------------------------------
_GEMM_PTR_FIELDS = ("a", "b", "scale", "shift", "residual", "aux", "out", "out2")
_gemm_timing = None


def set_gemm_timing(events):
    """Profiling hook: when ``events`` is a list, every cb_gemm launch is bracketed by CUDA events
    recorded on the launch stream and the (start, end) pair is appended to it. None disables."""
    global _gemm_timing
    _gemm_timing = events


# Measured launch configurations (tools/autotune_gemm.py on a B200 -> clipbert_b200/gemm_tuning.json): for each GEMM
# shape of the workload the fastest (tile width, wgrad K-split, k-chunks per stage). Shapes that are not in the table use
# the library's analytic model (choose_config in csrc/gemm.cu).
_tuning = None
_gemm_record = None


def gemm_key(kw):
    return "m%d n%d k%d mode%d t%d r%d a%d o%d f%d rm%d act%d" % (
        kw["m"], kw["n"], kw["k"], kw.get("mode", 0), kw.get("ntaps", 1), kw.get("residual") is not None, kw.get("aux") is not None,
        kw.get("out2") is not None, kw.get("out_fp32", 0), kw.get("rowmap", 0), kw.get("act", 0))


def _load_tuning():
    global _tuning
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_tuning.json")
    _tuning = {}
    if os.path.exists(path) and not os.environ.get("CB_NO_TUNING"):
        try:
            _tuning = json.load(open(path)).get("configs", {})
        except Exception:
            _tuning = {}


def gemm(**kw):
    """cb_gemm with keyword fields of cb_gemm_desc; tensor-valued fields are converted to pointers."""
    if _tuning is None:
        _load_tuning()
    if _gemm_record is not None:
        _gemm_record.append(dict(kw))
    if _tuning and "block_n" not in kw and "reserved" not in kw:
        t = _tuning.get(gemm_key(kw))
        if t is not None:
            kw = dict(kw, block_n=t[0], split_k=t[1], reserved=(t[2] << 8) | (32 if (len(t) > 3 and t[3]) else 0))
    d = L.GemmDesc()
    d.ntaps = 1
    d.tap_sign = 1
    d.split_k = 0
    for k, v in kw.items():
        if k in _GEMM_PTR_FIELDS:
            v = _p(v) if isinstance(v, torch.Tensor) else v
        setattr(d, k, v)
    timing = _gemm_timing is not None or _op_timing is not None
    if timing:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = L.lib().cb_gemm(ctypes.byref(d), _s())
    if rc != 0:
        raise RuntimeError("cb_gemm failed (%d): %s" % (rc, L.lib().cb_last_error().decode()))
    if timing:
        e1.record()
        if _gemm_timing is not None:
            _gemm_timing.append((e0, e1))
        if _op_timing is not None:
            _op_timing.append(("gemm mode=%d m=%d n=%d k=%d taps=%d res=%d aux=%d o2=%d f32=%d rm=%d" % (
                d.mode, d.m, d.n, d.k, d.ntaps, bool(d.residual), bool(d.aux), bool(d.out2), d.out_fp32, d.rowmap), e0, e1))


# module switch (bench --group_wgrad): which weight-gradient GEMMs go out as grouped launches. 0 none; 1 BertLayer (4 in one) +
# bottleneck block; 2 bottleneck blocks only; 3 BertLayer only; 4 BertLayer as two pairs (FFN pair as soon as du exists, attention
# pair after dqkv) + bottleneck blocks
group_wgrad = 3


def gemm_wgrad_group(kws):
    """cb_gemm_wgrad_group: the CB_GEMM_WGRAD problems ``kws`` (keyword dicts as for ``gemm``) in one persistent launch."""
    if not kws:
        return
    if len(kws) == 1:
        for kw in kws:
            gemm(**kw)
        return
    if _gemm_record is not None:
        _gemm_record.append(dict(group=[dict(kw) for kw in kws]))
    arr = (L.GemmDesc * len(kws))()
    for d, kw in zip(arr, kws):
        d.ntaps = 1
        d.tap_sign = 1
        d.split_k = 0
        for k, v in kw.items():
            if k in _GEMM_PTR_FIELDS:
                v = _p(v) if isinstance(v, torch.Tensor) else v
            setattr(d, k, v)
    timing = _gemm_timing is not None or _op_timing is not None
    if timing:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = L.lib().cb_gemm_wgrad_group(arr, len(kws), _s())
    if rc != 0:
        raise RuntimeError("cb_gemm_wgrad_group failed (%d): %s" % (rc, L.lib().cb_last_error().decode()))
    if timing:
        e1.record()
        if _gemm_timing is not None:
            _gemm_timing.append((e0, e1))
        if _op_timing is not None:
            _op_timing.append(("gemm wgrad group x%d" % len(kws), e0, e1))


def wgrad_split(m, n, k, ntaps=1, block_n=128):
    """0 = the library's launch-configuration model picks tile width, CTA pairing and the K-split."""
    return 0


# ------------------------------------------------------------------------------------------------
# BERT-side memory-bound ops
# ------------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, y, stats, eps):
    _call("cb_layernorm_fwd", _p(x), _p(gamma), _p(beta), _p(y), _p(stats), x.shape[0], x.shape[1], eps, _s())


def layernorm_bwd(dy, x, stats, gamma, dx, dx_drop, dgamma, dbeta, dbias_drop, p, seed):
    _call("cb_layernorm_bwd", _p(dy), _p(x), _p(stats), _p(gamma), _p(dx), _p(dx_drop), _p(dgamma), _p(dbeta),
          _p(dbias_drop), x.shape[0], x.shape[1], p, seed, _s())


def embed_text_fwd(ids, word, pos, type0, gamma, beta, out, stats, nseq, lt, l, eps, p, seed):
    _call("cb_embed_text_fwd", _p(ids), _p(word), _p(pos), _p(type0), _p(gamma), _p(beta), _p(out), _p(stats), nseq, lt, l,
          word.shape[0], word.shape[1], eps, p, seed, _s())


def embed_text_bwd(dh, ids, word, pos, type0, gamma, stats, dword, dpos, dtype0, dgamma, dbeta, nseq, lt, l, p, seed):
    _call("cb_embed_text_bwd", _p(dh), _p(ids), _p(word), _p(pos), _p(type0), _p(gamma), _p(stats), _p(dword), _p(dpos),
          _p(dtype0), _p(dgamma), _p(dbeta), nseq, lt, l, word.shape[0], word.shape[1], p, seed, _s())


def embed_visual_fwd(grid, seq2vid, n_ex, rowemb, colemb, type0, gamma, beta, out, stats, nseq, t, gh, gw, lt, l, eps, p,
                     seed):
    _call("cb_embed_visual_fwd", _p(grid), _p(seq2vid), n_ex, _p(rowemb), _p(colemb), _p(type0), _p(gamma), _p(beta), _p(out),
          _p(stats), nseq, t, gh, gw, lt, l, rowemb.shape[1], eps, p, seed, _s())


def embed_visual_bwd(dh, grid, seq2vid, vid_start, n_ex, rowemb, colemb, type0, gamma, stats, dv_tmp, dgrid, drow, dcol,
                     dtype0, dgamma, dbeta, nseq, nvid, t, gh, gw, lt, l, p, seed):
    _call("cb_embed_visual_bwd", _p(dh), _p(grid), _p(seq2vid), _p(vid_start), n_ex, _p(rowemb), _p(colemb), _p(type0),
          _p(gamma), _p(stats), _p(dv_tmp), _p(dgrid), _p(drow), _p(dcol), _p(dtype0), _p(dgamma), _p(dbeta), nseq, nvid, t,
          gh, gw, lt, l, rowemb.shape[1], p, seed, _s())


def nvls_allreduce(multicast_ptr, n, rank, world, scale, max_ctas=0):
    """All-reduce n fp32 elements at a multicast (symmetric-memory) address through the NVSwitch; see cb_nvls_allreduce_f32."""
    _call("cb_nvls_allreduce_f32", multicast_ptr, n, rank, world, scale, max_ctas, _s())


def clip_lse_loss(logits, labels, loss, dlogits, n_clips, nseq, ncls, grad_scale=1.0):
    _call("cb_clip_lse_loss", _p(logits), _p(labels), _p(loss), _p(dlogits), n_clips, nseq, ncls, grad_scale, _s())


def clip_pool_ce_loss(logits, labels, loss, dlogits, n_clips, nseq, ncls, pool, grad_scale=1.0):
    """pool: 1 = mean, 2 = max over the clips, then cross entropy (cb_clip_pool_ce_loss)."""
    _call("cb_clip_pool_ce_loss", _p(logits), _p(labels), _p(loss), _p(dlogits), n_clips, nseq, ncls, pool, grad_scale, _s())


def cross_entropy_fwd(logits, labels, loss, lse, ignore_index=-100):
    _call("cb_cross_entropy_fwd", _p(logits), logits.stride(0), _p(labels), _p(loss), _p(lse), logits.shape[0], logits.shape[1], ignore_index, _s())


def cross_entropy_bwd(logits, labels, lse, grad_loss, dlogits, ignore_index=-100):
    _call("cb_cross_entropy_bwd", _p(logits), logits.stride(0), _p(labels), _p(lse), _p(grad_loss), _p(dlogits), dlogits.stride(0), logits.shape[0],
          logits.shape[1], ignore_index, _s())


def colsum(x, out, m, n, ld=None):
    _call("cb_colsum", _p(x), n if ld is None else ld, _p(out), m, n, _s())


def dropout(x, y, p, seed):
    _call("cb_dropout", _p(x), _p(y), x.numel(), p, seed, _s())


def gelu_bwd(dy, u, dx):
    _call("cb_gelu_bwd", _p(dy), _p(u), _p(dx), dy.numel(), _s())


def pad_cast(src, dst):
    _call("cb_pad_cast", _p(src), src.stride(0), _p(dst), src.shape[0], src.shape[1], dst.shape[1], _s())


def cast_scale(src, dst, rowscale=None, row_len=1):
    _call("cb_cast_scale", _p(src), _p(rowscale), row_len, _p(dst), src.numel(), _s())


def cast_bf16_f32(src, dst):
    _call("cb_cast_bf16_f32", _p(src), _p(dst), src.numel(), _s())


def cast_scale_segments(master, packed, segments, scales):
    _call("cb_cast_scale_segments", _p(master), _p(packed), _p(segments), segments.shape[0], _p(scales), _s())


def attention_fwd(qkv, text_mask, ctx, lse, nseq, l, lt, heads, p, seed):
    _call("cb_attention_fwd", _p(qkv), qkv.shape[1], _p(text_mask), _p(ctx), ctx.shape[1], _p(lse), nseq, l, lt, heads, 64, p,
          seed, _s())


def attention_bwd(qkv, text_mask, ctx, dctx, lse, dqkv, nseq, l, lt, heads, p, seed):
    _call("cb_attention_bwd", _p(qkv), qkv.shape[1], _p(text_mask), _p(ctx), _p(dctx), ctx.shape[1], _p(lse), _p(dqkv),
          dqkv.shape[1], nseq, l, lt, heads, 64, p, seed, _s())


# ------------------------------------------------------------------------------------------------
# CNN-side memory-bound ops (NHWC bf16)
# ------------------------------------------------------------------------------------------------
def stem_im2col(x, out, n, h, w, kp, mean=(0.0, 0.0, 0.0)):
    dt = 0 if x.dtype == torch.float32 else 1
    _call("cb_stem_im2col", _p(x), dt, _p(out), n, h, w, kp, mean[0], mean[1], mean[2], _s())


def stem_s2d(x, out, n, h, w, ld, mean=(0.0, 0.0, 0.0)):
    dt = 0 if x.dtype == torch.float32 else 1
    _call("cb_stem_s2d", _p(x), dt, _p(out), n, h, w, ld, mean[0], mean[1], mean[2], _s())


def resize_pad(x, y, new_h, new_w):
    """x: (..., h, w) uint8 / fp32 planes; y: (..., S, S) fp32 - bilinear (align_corners=False) resize to new_h x new_w, zero pad to S."""
    dt = 0 if x.dtype == torch.float32 else 1
    _call("cb_resize_pad", _p(x), dt, _p(y), x.numel() // (x.shape[-1] * x.shape[-2]), x.shape[-2], x.shape[-1], new_h, new_w, y.shape[-1], _s())


def maxpool3x3s2(x, y, n, h, w, c, row_pitch=None, img_pitch=None):
    if row_pitch is None:
        _call("cb_maxpool3x3s2", _p(x), _p(y), n, h, w, c, _s())
    else:
        _call("cb_maxpool3x3s2_strided", _p(x), _p(y), n, h, w, c, row_pitch, img_pitch, _s())


def subsample2(x, y, n, h, w, c):
    _call("cb_subsample2", _p(x), _p(y), n, h, w, c, _s())


def unsubsample2_mask(dsub, act, dx, n, h, w, c):
    _call("cb_unsubsample2_mask", _p(dsub), _p(act), _p(dx), n, h, w, c, _s())


def maxpool2x2_relu_fwd(x, y, n, h, w, c):
    _call("cb_maxpool2x2_relu_fwd", _p(x), _p(y), n, h, w, c, _s())


def maxpool2x2_relu_bwd(dy, x, dx_pad, n, h, w, c):
    _call("cb_maxpool2x2_relu_bwd", _p(dy), _p(x), _p(dx_pad), n, h, w, c, _s())


def relu_mask(dy, act, dx):
    _call("cb_relu_mask", _p(dy), _p(act), _p(dx), dy.numel(), _s())
