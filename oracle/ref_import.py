"""Import the reference's OWN transformer code (read-only, from /root/reference) through import shims.

TEST INFRASTRUCTURE. Used to pin oracle/clipbert_ref.py and to generate tests/golden/*. The reference
targets transformers==2.11 + apex; this container has transformers 5.x and no apex, so a handful of
names are aliased. Nothing here is copied from the reference: its files are
imported where they lie. Unavailable on the GPU box (no /root/reference there).
"""
import json
import os
import sys
import types

import torch
import torch.nn.functional as F
from torch import nn

REFERENCE_ROOT = os.environ.get("CLIPBERT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "modeling"))


_mod = None


def load():
    """Returns the imported module src.modeling.modeling of the reference."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    # apex FusedLayerNorm == F.layer_norm (apex itself falls back to it on CPU)
    apex = types.ModuleType("apex")
    apex_n = types.ModuleType("apex.normalization")
    apex_f = types.ModuleType("apex.normalization.fused_layer_norm")
    apex_f.FusedLayerNorm = nn.LayerNorm
    sys.modules.setdefault("apex", apex)
    sys.modules.setdefault("apex.normalization", apex_n)
    sys.modules.setdefault("apex.normalization.fused_layer_norm", apex_f)

    import transformers
    import transformers.activations as act
    import transformers.file_utils as fu
    import transformers.modeling_utils as mu
    if not hasattr(act, "gelu"):
        act.gelu = F.gelu
    if not hasattr(act, "gelu_new"):
        act.gelu_new = lambda x: F.gelu(x, approximate="tanh")
    if not hasattr(act, "swish"):
        act.swish = F.silu
    cfgmod = types.ModuleType("transformers.configuration_bert")
    cfgmod.BertConfig = transformers.BertConfig
    sys.modules.setdefault("transformers.configuration_bert", cfgmod)
    if not hasattr(fu, "add_start_docstrings"):
        fu.add_start_docstrings = lambda *a, **k: (lambda f: f)
    if not hasattr(fu, "add_start_docstrings_to_callable"):
        fu.add_start_docstrings_to_callable = lambda *a, **k: (lambda f: f)
    if not hasattr(mu, "prune_linear_layer"):
        from transformers.pytorch_utils import prune_linear_layer
        mu.prune_linear_layer = prune_linear_layer

    class PreTrainedModel(nn.Module):
        """hf 2.11 semantics needed by the reference: init_weights, extended mask, head mask."""
        config_class = None
        base_model_prefix = ""

        def __init__(self, config, *a, **k):
            super().__init__()
            self.config = config

        @property
        def dtype(self):
            return next(self.parameters()).dtype

        def init_weights(self):
            self.apply(self._init_weights)

        def get_extended_attention_mask(self, attention_mask, input_shape, device):
            ext = attention_mask[:, None, None, :].to(dtype=self.dtype)
            return (1.0 - ext) * -10000.0

        def get_head_mask(self, head_mask, num_hidden_layers, is_attention_chunked=False):
            return [None] * num_hidden_layers

    mu.PreTrainedModel = PreTrainedModel
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import src.modeling.modeling as m
    _mod = m
    return m


def bert_config(**extra):
    import transformers
    with open(os.path.join(REFERENCE_ROOT, "src", "configs", "base_model.json")) as f:
        base = json.load(f)
    base.update(extra)
    cfg = transformers.BertConfig(**base)
    for k, v in base.items():
        setattr(cfg, k, v)
    if not hasattr(cfg, "output_attentions") or cfg.output_attentions is None:
        cfg.output_attentions = False
    if not hasattr(cfg, "output_hidden_states") or cfg.output_hidden_states is None:
        cfg.output_hidden_states = False
    cfg.is_decoder = False
    return cfg


def build_reference_transformer(sd, cls_name="ClipBertForVideoTextRetrieval", **cfg_extra):
    """Instantiate the reference class and load the 'transformer.*' entries of a flat state dict."""
    m = load()
    defaults = dict(num_labels=2, classifier="mlp", cls_hidden_scale=2, loss_type="ce", margin=0.2)
    defaults.update(cfg_extra)
    cfg = bert_config(**defaults)
    model = getattr(m, cls_name)(cfg)
    sub = {k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")}
    if cls_name == "ClipBertForPreTraining":
        sub["cls.predictions.decoder.weight"] = sub["bert.embeddings.word_embeddings.weight"]
    missing, unexpected = model.load_state_dict(sub, strict=False)
    missing = [k for k in missing if "position_ids" not in k and k != "cls.predictions.decoder.bias"]
    assert not missing and not unexpected, (missing, unexpected)
    return model.eval()
