"""Generate tests/golden/*.pt from the REFERENCE's own code (run in the authoring container only).

Transformer half: /root/reference/src/modeling/{modeling,transformers}.py imported read-only through
oracle/ref_import.py (fp32, eval mode). CNN half: detectron2 is not installable offline, so those
vectors come from the restated oracle cross-checked against torchvision's ResNet-50 (see
tests/test_oracle.py) and are labelled "restated" inside the file.

Weights are regenerated from seeds (oracle/synth.py); weight checksums are stored so that an RNG
drift between torch versions is detected instead of producing silent mismatches.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import clipbert_ref as R, ref_import, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def checksum(sd, keys):
    return {k: float(sd[k].double().sum()) for k in keys}


def main():
    assert ref_import.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    sd = synth.full_state_dict(42)
    wkeys = ["transformer.bert.encoder.layer.0.attention.self.query.weight", "transformer.bert.embeddings.word_embeddings.weight",
             "transformer.classifier.2.weight", "cnn.feature.backbone.res5.2.conv3.weight", "cnn.grid_encoder.0.weight"]

    # ---------------- retrieval head, reference classes ----------------
    g = torch.Generator().manual_seed(7)
    nvid, T, n_ex, lt = 2, 2, 2, 32
    grid = (torch.randn(nvid, T, 3, 3, 768, generator=g).abs() * 2).to(torch.bfloat16).float()
    ids, mask = synth.synth_text(nvid * n_ex, lt, seed=11)
    labels = torch.tensor([1, 0, 1, 0])
    rep = R.repeat_tensor_rows(grid, [n_ex] * nvid)
    model = ref_import.build_reference_transformer(sd, "ClipBertForVideoTextRetrieval")
    for p in model.parameters():
        p.requires_grad_(True)
    gr = rep.clone().requires_grad_(True)
    out = model(ids, gr, mask, labels=labels)
    seq, pooled = model.bert(ids, gr, mask)[:2]
    out["loss"].mean().backward()
    named = dict(model.named_parameters())
    gsel = ["bert.encoder.layer.0.attention.self.query.weight", "bert.encoder.layer.11.output.dense.weight", "bert.pooler.dense.bias",
            "classifier.2.weight", "bert.embeddings.LayerNorm.weight", "bert.visual_embeddings.row_position_embeddings.weight"]
    torch.save(dict(
        source="reference: src/modeling/modeling.py ClipBertForVideoTextRetrieval (imported via oracle/ref_import.py), fp32, eval",
        weights_seed=42, weight_checksums=checksum(sd, wkeys), grid=grid.to(torch.bfloat16), ids=ids, mask=mask, labels=labels,
        n_ex=n_ex, logits=out["logits"].detach(), loss=out["loss"].detach(), pooled=pooled.detach(),
        seq_first_rows=seq.detach()[:, :2, :32].clone(), seq_mean=float(seq.mean()), seq_std=float(seq.std()),
        dgrid=gr.grad.detach().view(nvid, n_ex, T, 3, 3, 768).sum(1).to(torch.bfloat16),
        grad_norms={k: float(named[k].grad.norm()) for k in gsel},
        grad_slices={k: named[k].grad.detach().flatten()[:64].clone() for k in gsel},
    ), os.path.join(OUT, "transformer_retrieval.pt"))

    # ---------------- multiple choice (TGIF-QA style), reference class ----------------
    sd_mc = dict(sd)
    sd_mc.update(synth.transformer_state_dict(50, num_labels=1))
    ids5, mask5 = synth.synth_text(10, 25, seed=5)
    grid1 = (torch.randn(2, 1, 3, 3, 768, generator=g).abs()).to(torch.bfloat16).float()
    model = ref_import.build_reference_transformer(sd_mc, "ClipBertForMultipleChoice", num_labels=5)
    with torch.no_grad():
        o = model(ids5, R.repeat_tensor_rows(grid1, [5, 5]), mask5, labels=torch.tensor([1, 4]))
    torch.save(dict(source="reference: ClipBertForMultipleChoice, num_labels=5, loss_type ce", weights_seed=(42, 50), ids=ids5, mask=mask5,
                    grid=grid1.to(torch.bfloat16), labels=torch.tensor([1, 4]), logits=o["logits"], loss=o["loss"]),
               os.path.join(OUT, "transformer_multiple_choice.pt"))

    # ---------------- pre-training heads (MLM on text positions + ITM), reference class ----------------
    sd_pt = {k: v for k, v in sd.items() if not k.startswith("transformer.classifier.")}
    sd_pt.update({k: v for k, v in synth.transformer_state_dict(60, head="pretraining").items() if k.startswith("transformer.cls.")})
    model = ref_import.build_reference_transformer(sd_pt, "ClipBertForPreTraining")
    mlm_labels = torch.full((4, lt), -100)
    mlm_labels[0, 3], mlm_labels[1, 5], mlm_labels[3, 2] = 2003, 1996, 7592
    itm = torch.tensor([1, 0, 1, 1])
    with torch.no_grad():
        o = model(ids, rep, mask, mlm_labels=mlm_labels, itm_labels=itm)
    torch.save(dict(source="reference: ClipBertForPreTraining (MLM + ITM)", weights_seed=(42, 60), ids=ids, mask=mask, grid=grid.to(torch.bfloat16),
                    n_ex=n_ex, mlm_labels=mlm_labels, itm_labels=itm, itm_scores=o["itm_scores"], itm_loss=o["itm_loss"],
                    mlm_loss=o["mlm_loss"], mlm_scores_slice=o["mlm_scores"][:, :4, :64].clone(),
                    mlm_argmax=o["mlm_scores"].argmax(-1)), os.path.join(OUT, "transformer_pretraining.pt"))

    # ---------------- CNN (restated oracle), 2 frames at 96x96 and 1 frame at 224 ----------------
    x = synth.synth_images(1, 2, size=96, seed=21)
    with torch.no_grad():
        grid96, st = R.grid_feat_backbone(x, sd, return_stages=True)
        x224 = synth.synth_images(1, 1, size=224, seed=22)
        grid224 = R.grid_feat_backbone(x224, sd)
    torch.save(dict(source="restated detectron2 MSRA R-50 (oracle/clipbert_ref.py); not reference-executed: detectron2 unavailable",
                    weights_seed=42, weight_checksums=checksum(sd, wkeys), image_seed=(21, 22),
                    stage_stats={k: (float(v.mean()), float(v.std())) for k, v in st.items()},
                    res5_slice=st["res5"][:, :32, :, :].clone(), grid96=grid96.clone(), grid224=grid224.clone()),
               os.path.join(OUT, "cnn_grid.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
