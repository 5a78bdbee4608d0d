"""Deterministic synthetic workload for the ClipBERT hot path: random-init weights in the reference's state-dict
layout and synthetic inputs of the benchmark shapes. Shapes and value ranges only - no arithmetic of the path lives here.
Used by bench.py / tools (there is no network for datasets or checkpoints) and, through ``oracle/synth.py``, by the tests.

Weights follow the reference initialisers:
  Linear/Embedding ~ N(0, 0.02), LayerNorm (1, 0), biases 0      src/modeling/transformers.py:559-570
  convs: Kaiming-normal fan_out (d2 c2_msra_fill), FrozenBN buffers randomised
``perturb=True`` additionally randomises biases / LayerNorm affine so that parity tests exercise
every term (an all-zero bias hides bias bugs).
"""
import math

import torch

import types

from .grid_feat import RESNET50_STAGES

# src/configs/base_model.json (BertConfig of the reference) + the ClipBERT additions (modeling.py:40-59)
BERT_CFG = dict(
    hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
    vocab_size=30522, max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12,
    hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, initializer_range=0.02,
    max_grid_row_position_embeddings=100, max_grid_col_position_embeddings=100,
    backbone_channel_in_size=2048, pad_token_id=0)


def make_cfg(**extra):
    """The model config the task scripts assemble (run_video_retrieval.py:184-192): base_model.json + head settings."""
    d = dict(BERT_CFG)
    d.update(num_labels=2, loss_type="ce", margin=0.2, classifier="mlp", cls_hidden_scale=2)
    d.update(extra)
    return types.SimpleNamespace(**d)


def _n(g, *shape, std=1.0):
    return torch.randn(*shape, generator=g) * std


def cnn_state_dict(seed=42, perturb=True):
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, cout, cin, k, gain=1.0):
        std = gain * math.sqrt(2.0 / (cout * k * k))       # kaiming_normal_(mode="fan_out", relu)
        sd[name + ".weight"] = _n(g, cout, cin, k, k, std=std)

    def bn(name, c, last=False):
        # keep the residual branch small (scale ~0.2..0.5 on conv3) so 16 blocks do not blow up
        lo, hi = (0.2, 0.5) if last else (0.5, 1.5)
        sd[name + ".weight"] = torch.rand(c, generator=g) * (hi - lo) + lo
        sd[name + ".bias"] = _n(g, c, std=0.1)
        sd[name + ".running_mean"] = _n(g, c, std=0.1)
        sd[name + ".running_var"] = torch.rand(c, generator=g) + 0.5

    p = "cnn.feature.backbone."
    conv(p + "stem.conv1", 64, 3, 7)
    bn(p + "stem.conv1.norm", 64)
    cin = 64
    for name, nblocks, mid, cout, _ in RESNET50_STAGES:
        for b in range(nblocks):
            q = "%s%s.%d." % (p, name, b)
            if b == 0:
                conv(q + "shortcut", cout, cin, 1)
                bn(q + "shortcut.norm", cout)
            conv(q + "conv1", mid, cin, 1)
            bn(q + "conv1.norm", mid)
            conv(q + "conv2", mid, mid, 3)
            bn(q + "conv2.norm", mid)
            conv(q + "conv3", cout, mid, 1)
            bn(q + "conv3.norm", cout, last=True)
            cin = cout
    conv("cnn.grid_encoder.0", BERT_CFG["hidden_size"], 2048, 3)
    return sd


def transformer_state_dict(seed=43, head="retrieval", num_labels=2, perturb=True, cfg=BERT_CFG, cls_hidden_scale=2):
    g = torch.Generator().manual_seed(seed)
    d, ff = cfg["hidden_size"], cfg["intermediate_size"]
    std = cfg["initializer_range"]
    sd = {}

    def lin(name, out_f, in_f):
        sd[name + ".weight"] = _n(g, out_f, in_f, std=std)
        sd[name + ".bias"] = _n(g, out_f, std=std) if perturb else torch.zeros(out_f)

    def ln(name):
        sd[name + ".weight"] = 1.0 + (_n(g, d, std=0.1) if perturb else torch.zeros(d))
        sd[name + ".bias"] = _n(g, d, std=0.1) if perturb else torch.zeros(d)

    def emb(name, n):
        sd[name + ".weight"] = _n(g, n, d, std=std)

    b = "transformer.bert."
    emb(b + "embeddings.word_embeddings", cfg["vocab_size"])
    emb(b + "embeddings.position_embeddings", cfg["max_position_embeddings"])
    emb(b + "embeddings.token_type_embeddings", cfg["type_vocab_size"])
    ln(b + "embeddings.LayerNorm")
    emb(b + "visual_embeddings.position_embeddings", cfg["max_position_embeddings"])   # allocated, unused (modeling.py:97)
    emb(b + "visual_embeddings.row_position_embeddings", cfg["max_grid_row_position_embeddings"])
    emb(b + "visual_embeddings.col_position_embeddings", cfg["max_grid_col_position_embeddings"])
    emb(b + "visual_embeddings.token_type_embeddings", 1)
    ln(b + "visual_embeddings.LayerNorm")
    for i in range(cfg["num_hidden_layers"]):
        q = "%sencoder.layer.%d." % (b, i)
        lin(q + "attention.self.query", d, d)
        lin(q + "attention.self.key", d, d)
        lin(q + "attention.self.value", d, d)
        lin(q + "attention.output.dense", d, d)
        ln(q + "attention.output.LayerNorm")
        lin(q + "intermediate.dense", ff, d)
        lin(q + "output.dense", d, ff)
        ln(q + "output.LayerNorm")
    lin(b + "pooler.dense", d, d)
    if head == "pretraining":
        lin("transformer.cls.predictions.transform.dense", d, d)
        ln("transformer.cls.predictions.transform.LayerNorm")
        sd["transformer.cls.predictions.bias"] = _n(g, cfg["vocab_size"], std=std) if perturb else torch.zeros(cfg["vocab_size"])
        lin("transformer.cls.seq_relationship", 2, d)
    else:
        lin("transformer.classifier.0", d * cls_hidden_scale, d)
        lin("transformer.classifier.2", num_labels, d * cls_hidden_scale)
    return sd


def full_state_dict(seed=42, **kw):
    sd = cnn_state_dict(seed)
    sd.update(transformer_state_dict(seed + 1, **kw))
    return sd


IMAGE_MEAN = (123.675, 116.28, 103.53)     # src/configs/msrvtt_ret_base_resnet50.json:18-19 (std = 1)


def synth_images(n_videos, n_frames, size=224, seed=42, as_uint8=False):
    """uint8 U[0,255] frames; float path = minus mean as ImageNorm (src/datasets/data_utils.py:256-276)."""
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (n_videos, n_frames, 3, size, size), generator=g, dtype=torch.uint8)
    if as_uint8:
        return u8
    return u8.float() - torch.tensor(IMAGE_MEAN).view(1, 1, 3, 1, 1)


def synth_text(n_seq, max_len=32, seed=42, vocab=30522):
    """[CLS] body [SEP] pad, valid length U[8, max_len] (min(8,max_len) if shorter)."""
    g = torch.Generator().manual_seed(seed + 7)
    ids = torch.zeros(n_seq, max_len, dtype=torch.long)
    mask = torch.zeros(n_seq, max_len, dtype=torch.long)
    lo = min(8, max_len)
    lens = torch.randint(lo, max_len + 1, (n_seq,), generator=g)
    for i in range(n_seq):
        n = int(lens[i])
        ids[i, :n] = torch.randint(1000, vocab, (n,), generator=g)
        ids[i, 0] = 101
        ids[i, n - 1] = 102
        mask[i, :n] = 1
    return ids, mask


def synth_batch(n_videos, n_frames, n_ex=1, size=224, max_len=32, num_classes=2, seed=42):
    ids, mask = synth_text(n_videos * n_ex, max_len, seed)
    g = torch.Generator().manual_seed(seed + 11)
    labels = torch.randint(0, num_classes, (n_videos * n_ex,), generator=g)
    return dict(visual_inputs=synth_images(n_videos, n_frames, size, seed), text_input_ids=ids,
                text_input_mask=mask, labels=labels, n_examples_list=[n_ex] * n_videos)
