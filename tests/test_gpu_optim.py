"""FusedAdamW (clipbert_b200/optim.py + csrc/optim.cu) against the CPU oracle of the reference optimizer step
(oracle/adamw_ref.py, pinned to src/optimization/adamw.py): parameter trajectories over several clipped steps for all
eight reference parameter groups, the emitted bf16 tensor-core operands, gradient zeroing and state_dict layout.
Tolerance: fp32 arithmetic with approximate division / sqrt on the device -> 2e-6 relative on parameters, 1e-3 on the
per-step update."""
import pytest
import torch

from util import make_cfg, relerr

pytestmark = pytest.mark.gpu


def e2e_param_groups(model, lr=5e-5, wd=1e-3, cnn_lr=5e-5, cnn_wd=1e-3, transformer_lr_mul=1.0, transformer_lr_mul_prefix="",
                     cnn_lr_mul=5.0, cnn_lr_mul_prefix="grid_encoder"):
    """The eight groups of setup_e2e_optimizer / build_e2e_optimizer_w_lr_mul (src/optimization/utils.py:96-161)."""
    no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight"]

    def build(named, base_lr, base_wd, mul, prefix):
        top = [(n, p) for n, p in named if prefix and prefix in n]
        rest = [(n, p) for n, p in named if not (prefix and prefix in n)]
        return [dict(params=[p for n, p in top if not any(nd in n for nd in no_decay)], lr=mul * base_lr, weight_decay=base_wd),
                dict(params=[p for n, p in top if any(nd in n for nd in no_decay)], lr=mul * base_lr, weight_decay=0.0),
                dict(params=[p for n, p in rest if not any(nd in n for nd in no_decay)], lr=base_lr, weight_decay=base_wd),
                dict(params=[p for n, p in rest if any(nd in n for nd in no_decay)], lr=base_lr, weight_decay=0.0)]
    tr = [(n, p) for n, p in model.named_parameters() if "transformer" in n and p.requires_grad]
    cnn = [(n, p) for n, p in model.named_parameters() if "cnn" in n and p.requires_grad]
    return build(tr, lr, wd, transformer_lr_mul, transformer_lr_mul_prefix) + build(cnn, cnn_lr, cnn_wd, cnn_lr_mul, cnn_lr_mul_prefix)


def test_fused_adamw_matches_reference_optimizer(cuda):
    import clipbert_b200 as cb
    from clipbert_b200.optim import FusedAdamW
    from oracle import adamw_ref as A, synth
    sd = synth.full_state_dict(42)
    cfg = make_cfg(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = cb.ClipBert(cfg, detectron2_model_cfg="R-50-grid.yaml")
    model.load_state_dict(sd)
    model = model.to(cuda).train()
    groups = e2e_param_groups(model)
    assert len(groups) == 8          # the hard-coded group count of run_video_retrieval.py:455
    groups = [g for g in groups if g["params"]]          # torch rejects nothing here, but the reference's empty "top" groups carry no work
    opt = FusedAdamW(groups, lr=5e-5, betas=(0.9, 0.98), model=model)
    # one real forward so that both halves own flat buffers and packed operands
    batch = synth.synth_batch(1, 2, n_ex=1, size=96, seed=3)
    mb = {k: (v.to(cuda) if torch.is_tensor(v) else list(v)) for k, v in batch.items()}
    model(mb)["loss"].mean().backward()
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    gidx = {id(p): gi for gi, g in enumerate(opt.param_groups) for p in g["params"]}
    # sample of parameters followed on the CPU oracle: every 9th + one of each special kind
    special = ("grid_encoder.0.weight", "res5.2.conv2.weight", "word_embeddings.weight", "LayerNorm.weight", "classifier.2.bias",
               "attention.self.key.weight", "res3.0.shortcut.weight")
    sel = [i for i, (n, p) in enumerate(named) if i % 9 == 0 or any(n.endswith(s) for s in special)]
    p0 = {i: named[i][1].detach().float().cpu().clone() for i in sel}
    steps, max_norm = 3, 1.0
    grads_cpu, norms = [], []
    gen = torch.Generator(device=cuda).manual_seed(17)
    for t in range(steps):
        for m, scale in ((model.transformer, 3e-3 if t == 1 else 3e-5), (model.cnn, 1e-3 if t == 1 else 1e-5)):   # step 1 clips, 0 and 2 do not
            m._flat.attach_grads()
            m._flat.grad.zero_()
            for e in m._flat.entries:                   # only real parameter elements carry gradient (as after a backward)
                if e["param"].requires_grad:
                    m._flat.grad[e["offset"]: e["offset"] + e["numel"]].copy_(torch.randn(e["numel"], generator=gen, device=cuda) * scale)
        grads_cpu.append({i: named[i][1].grad.detach().float().cpu().clone() for i in sel})
        total = torch.sqrt(sum((p.grad.double() ** 2).sum() for _, p in named))
        norms.append(float(total))
        n = opt.clip_grad_norm(max_norm)
        assert abs(float(n) - float(total)) < 1e-4 * float(total)
        opt.step(zero_grad=(t == steps - 1))
    if cuda.type == "cuda":
        torch.cuda.synchronize()
    assert norms[1] > max_norm > norms[0], norms           # the test really exercises both branches of the clip
    assert float(model.transformer._flat.grad.abs().max()) == 0.0 and float(model.cnn._flat.grad.abs().max()) == 0.0
    # ---- oracle trajectories ----
    worst = 0.0
    for i in sel:
        n_, p = named[i]
        g = opt.param_groups[gidx[id(p)]]
        ref, m, v = p0[i], torch.zeros_like(p0[i]), torch.zeros_like(p0[i])
        for t in range(steps):
            coef = max_norm / (norms[t] + 1e-6)
            gr = grads_cpu[t][i] * (coef if coef < 1.0 else 1.0)
            prev = ref
            ref, m, v = A.adamw_step(ref, gr, m, v, t + 1, g["lr"], g["betas"], g["eps"], g["weight_decay"], g["correct_bias"])
        got = p.detach().float().cpu()
        assert relerr(got, ref) < 2e-6, (n_, relerr(got, ref))
        upd = relerr(got - p0[i], ref - p0[i])
        worst = max(worst, upd)
        assert upd < 1e-3, (n_, upd)
        st = opt.state[p]
        assert st["step"] == steps and st["exp_avg"].shape == p.shape and relerr(st["exp_avg"].float().cpu(), m) < 1e-5
        assert relerr(st["exp_avg_sq"].float().cpu(), v) < 1e-5
    # ---- emitted bf16 operands == a fresh cast of the updated masters (FrozenBN scale folded for convs) ----
    tf = model.transformer._flat
    assert torch.equal(tf.packed[: tf.packed_prefix], tf.master[: tf.packed_prefix].to(torch.bfloat16))
    cnn = model.cnn
    for name, mconv in cnn._convs():
        if not mconv.weight.requires_grad:
            continue
        e = mconv._e
        w = cnn._flat.master[e["offset"]: e["offset"] + e["numel"]].view(mconv.cout, -1)
        want = (w * mconv._scale[:, None]) if mconv._scale is not None else w
        assert torch.equal(mconv._w.view(mconv.cout, -1), want.to(torch.bfloat16)), name
    assert not cnn._dirty and not model.transformer._dirty
    # ---- the next forward uses the new weights without a re-cast: same bits as a freshly loaded copy of the updated model ----
    model.eval()
    fresh = cb.ClipBert(cfg, detectron2_model_cfg="R-50-grid.yaml")
    fresh.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()})
    fresh = fresh.to(cuda).eval()
    with torch.no_grad():
        mb1 = {k: (v.to(cuda) if torch.is_tensor(v) else list(v)) for k, v in batch.items()}
        mb2 = {k: (v.to(cuda) if torch.is_tensor(v) else list(v)) for k, v in batch.items()}
        from clipbert_b200 import ops
        n0 = ops.launch_count()
        a = model(mb1)["logits"]
        n1 = ops.launch_count()
        b = fresh(mb2)["logits"]
        n2 = ops.launch_count()
    assert torch.equal(a, b)
    if cuda.type == "cuda":      # (the CPU replay of this test through the ABI emulator launches no kernels)
        assert (n2 - n1) - (n1 - n0) >= 2, "the freshly loaded model re-casts its weights, the optimized one must not"
    # ---- state_dict carries the reference's keys ----
    osd = opt.state_dict()
    some = next(iter(osd["state"].values()))
    assert set(some) >= {"step", "exp_avg", "exp_avg_sq"}


def test_fused_adamw_state_dict_round_trip_after_steps(cuda):
    """TrainingRestorer (src/utils/load_save.py:245-300) restores the optimizer in the middle of training. torch's
    load_state_dict replaces the state tensors, so FusedAdamW must re-adopt the restored moments / step counts into its flat
    buffers: run 2 steps, save, run a 3rd step, restore, re-run the 3rd step with the same gradients -> identical parameters,
    and both equal the oracle's third step from the saved state."""
    import copy

    import clipbert_b200 as cb
    from clipbert_b200.optim import FusedAdamW
    from oracle import adamw_ref as A, synth
    sd = synth.full_state_dict(42)
    cfg = make_cfg(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = cb.ClipBert(cfg, detectron2_model_cfg="R-50-grid.yaml")
    model.load_state_dict(sd)
    model = model.to(cuda).train()
    groups = [g for g in e2e_param_groups(model) if g["params"]]
    opt = FusedAdamW(groups, lr=5e-5, betas=(0.9, 0.98), model=model)
    batch = synth.synth_batch(1, 2, n_ex=1, size=96, seed=3)
    model({k: (v.to(cuda) if torch.is_tensor(v) else list(v)) for k, v in batch.items()})["loss"].mean().backward()
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    gidx = {id(p): gi for gi, g in enumerate(opt.param_groups) for p in g["params"]}
    gen = torch.Generator(device=cuda).manual_seed(23)

    def fill_grads(seed_scale):
        for m in (model.transformer, model.cnn):
            m._flat.attach_grads()
            m._flat.grad.zero_()
            for e in m._flat.entries:
                if e["param"].requires_grad:
                    m._flat.grad[e["offset"]: e["offset"] + e["numel"]].copy_(torch.randn(e["numel"], generator=gen, device=cuda) * seed_scale)

    for _ in range(2):
        fill_grads(1e-5)
        opt.step(zero_grad=True)
    saved_opt = copy.deepcopy(opt.state_dict())
    saved_params = {n: p.detach().clone() for n, p in named}
    sel = [i for i, (n, p) in enumerate(named) if i % 17 == 0 or n.endswith(("grid_encoder.0.weight", "word_embeddings.weight", "classifier.2.bias"))]
    mom = {i: (opt.state[named[i][1]]["exp_avg"].detach().float().cpu().clone(), opt.state[named[i][1]]["exp_avg_sq"].detach().float().cpu().clone())
           for i in sel}
    fill_grads(2e-5)
    g3 = {n: p.grad.detach().clone() for n, p in named}
    opt.step(zero_grad=True)
    after_a = {n: p.detach().clone() for n, p in named}
    # ---- restore parameters + optimizer, replay step 3 ----
    with torch.no_grad():
        for n, p in named:
            p.copy_(saved_params[n])
    model.transformer.mark_weights_updated()
    model.cnn.mark_weights_updated()
    opt.load_state_dict(saved_opt)
    for m in (model.transformer, model.cnn):
        m._flat.attach_grads()
    for n, p in named:
        p.grad.copy_(g3[n])
    opt.step(zero_grad=True)
    for n, p in named:
        assert torch.equal(p.detach(), after_a[n]), n        # the restored run is bit-identical to the uninterrupted one
    assert all(opt.state[p]["step"] == 3 for _, p in named)
    # ---- and it is the reference's third step from the saved state ----
    for i in sel:
        n_, p = named[i]
        g = opt.param_groups[gidx[id(p)]]
        ref, _, _ = A.adamw_step(saved_params[n_].float().cpu(), g3[n_].float().cpu(), mom[i][0], mom[i][1], 3, g["lr"], g["betas"], g["eps"],
                                 g["weight_decay"], g["correct_bias"])
        assert relerr(p.detach().float().cpu(), ref) < 2e-6, n_
