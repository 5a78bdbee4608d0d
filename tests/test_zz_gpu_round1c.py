"""GPU parity tests for the paths added after the round's GPU budget was spent (written against the CPU oracle, first run on
a B200 by the round-end driver). They sort last so that a surprise here cannot hide the result of the established suites.
Every kernel they launch is one the earlier suites already cover; what is new is the host-side index bookkeeping.
"""
import numpy as np
import pytest
import torch

from util import TOL_GRAD, TOL_LOGITS, cosine, make_cfg, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def weights():
    from oracle import synth
    return synth.full_state_dict(42)


def _pretraining_model(weights, cuda, **cfg_extra):
    import clipbert_b200 as cb
    from oracle import synth
    sd = {k: v for k, v in weights.items() if not k.startswith("transformer.classifier.")}
    sd.update({k: v for k, v in synth.transformer_state_dict(60, head="pretraining").items() if k.startswith("transformer.cls.")})
    cfg = make_cfg(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **cfg_extra)
    model = cb.ClipBertForPreTraining(cfg)
    res = model.load_state_dict({k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")}, strict=False)
    assert set(res.missing_keys) <= {"cls.predictions.decoder.weight", "cls.predictions.decoder.bias"} and not res.unexpected_keys
    return model.to(cuda), sd


def test_pretraining_random_visual_token_sampling(cuda, weights):
    """VisualInputEmbedding's train-mode sampling (src/modeling/modeling.py:15-34,80-88): with numpy seeded like the oracle's
    draw the B200 path keeps the same sorted subset; forward, the scattered grid gradient and the row / column position
    table gradients (index_add of the virtual-grid gradient) against fp32 autograd on the oracle."""
    from oracle import clipbert_ref as R, synth
    model, sd = _pretraining_model(weights, cuda, pixel_random_sampling_size=7)
    model.train()
    g = torch.Generator().manual_seed(5)
    grid0 = torch.randn(2, 2, 4, 5, 768, generator=g).abs().bfloat16()          # 20 visual tokens, keep 7
    ids, mask = synth.synth_text(4, 12, seed=9)                                 # 2 captions per video
    mlm = torch.full((4, 12), -100, dtype=torch.long)
    mlm[:, 3] = ids[:, 3]
    mlm[:, 7] = ids[:, 7]
    itm = torch.tensor([1, 0, 1, 0])
    grid = grid0.to(cuda).requires_grad_(True)
    np.random.seed(77)
    out = model(ids.to(cuda), grid, mask.to(cuda), mlm_labels=mlm.to(cuda), itm_labels=itm.to(cuda), _repeat_counts=[2, 2])
    assert out["mlm_scores"].shape == (4, 12, 30522)
    sdr = {k: (v.clone().requires_grad_(True) if k.startswith("transformer.") else v) for k, v in sd.items()}
    gr = grid0.float().requires_grad_(True)
    np.random.seed(77)
    ref = R.pretraining(ids, R.repeat_tensor_rows(gr, [2, 2]), mask, sdr, mlm, itm, pixel_random_sampling_size=7)
    assert relerr(out["itm_scores"], ref["itm_scores"]) < TOL_LOGITS
    assert relerr(out["mlm_scores"][:, :, :256], ref["mlm_scores"][:, :, :256]) < TOL_LOGITS
    (out["mlm_loss"].sum() / 8 + out["itm_loss"].mean()).backward()
    (ref["mlm_loss"].sum() / 8 + ref["itm_loss"].mean()).backward()
    np.random.seed(77)
    kept = R.random_sample_indices(20, 7)
    dropped = torch.tensor([j for j in range(20) if j not in set(kept.tolist())])
    gg = grid.grad.float().cpu().view(2, 2, 20, 768)
    assert float(gg[:, :, dropped].abs().max()) == 0.0, "dropped visual tokens must receive no gradient"
    assert cosine(grid.grad, gr.grad) > 0.999 and relerr(grid.grad, gr.grad) < TOL_GRAD
    named = dict(model.named_parameters())
    for name in ("bert.visual_embeddings.row_position_embeddings.weight", "bert.visual_embeddings.col_position_embeddings.weight",
                 "bert.visual_embeddings.LayerNorm.weight", "bert.visual_embeddings.token_type_embeddings.weight",
                 "bert.encoder.layer.0.attention.self.query.weight", "bert.encoder.layer.11.output.dense.weight"):
        r = sdr["transformer." + name].grad
        assert relerr(named[name].grad, r) < TOL_GRAD and cosine(named[name].grad, r) > 0.999, name
    # eval mode and sample sizes >= the token count leave the sequence whole (reference: self.training / num_samples >= seq_len)
    model.eval()
    with torch.no_grad():
        e1 = model(ids.to(cuda), grid0.to(cuda), mask.to(cuda), _repeat_counts=[2, 2])
    model.config.pixel_random_sampling_size = 0
    with torch.no_grad():
        e0 = model(ids.to(cuda), grid0.to(cuda), mask.to(cuda), _repeat_counts=[2, 2])
    assert torch.equal(e1["itm_scores"], e0["itm_scores"])
