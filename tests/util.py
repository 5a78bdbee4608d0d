import types

import torch


def relerr(a, b):
    """Normwise relative error ||a-b|| / ||b|| in float64."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cosine(a, b):
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float32)


def make_cfg(**extra):
    from oracle.clipbert_ref import BERT_CFG
    d = dict(BERT_CFG)
    d.update(num_labels=2, loss_type="ce", margin=0.2, classifier="mlp", cls_hidden_scale=2)
    d.update(extra)
    return types.SimpleNamespace(**d)


# Tolerances (normwise relative error), stated once:
#   TOL_BF16_OP : one fused op whose output is stored in bf16 (rounding 2^-9 ~ 2e-3 per element)
#   TOL_FP32_OP : one op with fp32 output from bf16 operands (accumulation-order noise only)
#   TOL_MATCHED : end-to-end forward against the bf16-rounding-matched oracle (north_star: 1e-3)
#   TOL_FP32_E2E: end-to-end forward against the plain fp32 oracle (bf16 storage through ~70 layers)
#   TOL_GRAD    : parameter gradients (bf16 activations + bf16 upstream grads) against fp32 autograd
TOL_BF16_OP = 4e-3
TOL_FP32_OP = 2e-5
TOL_MATCHED = 1e-3
TOL_FP32_E2E = 2e-2
TOL_GRAD = 5e-2
# TOL_MATCHED_DEEP: a deep stack (50 convs / 12 encoder layers) against the rounding-matched oracle. The two
# pipelines round the same quantities to bf16, but fp32 accumulation order differs, so values that sit on a
# bf16 rounding boundary flip by one ulp (2^-8 relative) and the flips compound with depth.
TOL_MATCHED_DEEP = 1e-2
# TOL_LOGITS: final logits (|logit| ~ 0.1-0.2 after 12 bf16 encoder layers) against either oracle
TOL_LOGITS = 2.5e-2
