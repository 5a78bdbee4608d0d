"""Flat parameter / gradient / packed-weight storage for one half of the model (CNN or transformer).

HBM layout (all per group):
  master  fp32 [total]   the nn.Parameters are views into this buffer (state_dict-compatible; conv
                         weights are stored KRSC = torch channels_last, so the GEMM "B" operand and the
                         wgrad output share the parameter's own physical layout)
  grad    fp32 [total]   .grad of every parameter is a view into this buffer; the wgrad kernels
                         red.global.add straight into it, and data-parallel training all-reduces this
                         single buffer (replaces Horovod's per-parameter fusion buffer,
                         src/tasks/run_video_retrieval.py:299-301,432)
  packed  bf16 [total]   tensor-core operand copy, refreshed only when a master parameter changes
                         (the role of apex amp O2's model-weight copy, run_video_retrieval.py:307-309)
"""
import torch

ALIGN = 64  # elements; keeps every tensor 128 B (bf16) / 256 B (fp32) aligned for TMA


def _round_up(n, a=ALIGN):
    return (n + a - 1) // a * a


class FlatGroup:
    # Optional allocator of the gradient buffer, ``f(numel, device) -> zero-filled fp32 tensor``: the NVLS exchange
    # (ClipBert.enable_overlapped_allreduce(exchange="nvls")) needs it in symmetric memory. None = torch.zeros.
    grad_factory = None

    def __init__(self, device):
        self.device = device
        self.entries = []      # dict(name, param, offset, numel, slot, kind, meta)
        self.total = 0
        self.master = None
        self.grad = None
        self.packed = None
        self._version = None
        self.packed_prefix = 0  # only [0, packed_prefix) is cast to bf16 (embeddings live after it)

    # -- registration ---------------------------------------------------------------------------
    def add(self, name, param, slot_numel=None, kind="plain", **meta):
        """Reserve a slot. kind: 'plain' (contiguous), 'conv' (KCRS param stored KRSC)."""
        numel = param.numel()
        slot = _round_up(slot_numel if slot_numel is not None else numel)
        self.entries.append(dict(name=name, param=param, offset=self.total, numel=numel, slot=slot, kind=kind, meta=meta))
        self.total += slot
        return self.entries[-1]

    def mark_packed_prefix(self):
        self.packed_prefix = self.total

    # -- materialisation ------------------------------------------------------------------------
    def _view(self, buf, e):
        p = e["param"]
        chunk = buf[e["offset"]: e["offset"] + e["numel"]]
        if e["kind"] == "conv":
            co, ci, kh, kw = p.shape
            return chunk.view(co, kh, kw, ci).permute(0, 3, 1, 2)
        return chunk.view(p.shape)

    def materialize(self):
        dev = self.device
        self.master = torch.zeros(self.total, dtype=torch.float32, device=dev)
        make = FlatGroup.grad_factory
        self.grad = torch.zeros(self.total, dtype=torch.float32, device=dev) if make is None else make(self.total, dev)
        self.packed = torch.zeros(self.total, dtype=torch.bfloat16, device=dev)
        for e in self.entries:
            p = e["param"]
            v = self._view(self.master, e)
            v.copy_(p.data.to(dev, torch.float32))
            p.data = v
            e["ptr"] = v.data_ptr()
        self.attach_grads(force=True)
        self._version = None

    def is_current(self):
        """False if a parameter was re-allocated behind our back (e.g. model.to(), .half())."""
        if self.master is None:
            return False
        for e in (self.entries[0], self.entries[-1]):
            if e["param"].data_ptr() != e["ptr"]:
                return False
        return True

    def attach_grads(self, force=False):
        """Make sure p.grad views the flat buffer; zero it when grads were dropped (set_to_none)."""
        probe = next((e["param"] for e in self.entries if e["param"].requires_grad), None)
        if probe is None:
            return
        if not force and probe.grad is not None and probe.grad.data_ptr() == self._grad_ptr0:
            return
        self.grad.zero_()
        for e in self.entries:
            p = e["param"]
            p.grad = self._view(self.grad, e) if p.requires_grad else None
        self._grad_ptr0 = probe.grad.data_ptr()

    def zero_grad(self):
        self.grad.zero_()

    # -- views ------------------------------------------------------------------------------------
    def master_flat(self, e, rows=None, cols=None):
        n = e["slot"] if rows is None else rows * cols
        t = self.master[e["offset"]: e["offset"] + n]
        return t if rows is None else t.view(rows, cols)

    def grad_flat(self, e, rows=None, cols=None):
        n = e["slot"] if rows is None else rows * cols
        t = self.grad[e["offset"]: e["offset"] + n]
        return t if rows is None else t.view(rows, cols)

    def packed_flat(self, e, rows=None, cols=None):
        n = e["slot"] if rows is None else rows * cols
        t = self.packed[e["offset"]: e["offset"] + n]
        return t if rows is None else t.view(rows, cols)

    # -- change tracking --------------------------------------------------------------------------
    def version(self):
        return sum(e["param"]._version for e in self.entries)

    def needs_repack(self):
        v = self.version()
        if v != self._version:
            self._version = v
            return True
        return False
