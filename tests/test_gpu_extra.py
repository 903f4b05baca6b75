"""Round-2 parity cases (VERDICT r1 "What's missing" 1-3) against anchors produced by the REAL reference
(oracle/make_golden.py --extra -> tests/golden/extra_anchors.json, extra_samples.npz):

  bench64    the bench workload itself: shipped JSON dims, B = 64, un-padded           (fp32 and bf16/tcgen05)
  cfg5       BASELINE cfg 5 dims: 6 + 6 layers, d 768, T = 32 (M = 33), B = 16, padded  (fp32 and bf16/tcgen05)
  evalfast   eval() + no_grad + padded frames: the nested-tensor fast path semantics
  decode256  BASELINE cfg 3: greedy decode B = 256, max_len 30

Tolerances: fp32 -- loss rel 2e-5, gradient norms rel 2e-3, sampled per-element gradients rel-L2 2e-3;
bf16 (bf16 operands/activations, fp32 accumulate) -- loss rel 2e-3, gradient norms rel 5e-2, sampled per-element
gradients rel-L2 6e-2 per tensor (measured values are printed with -s)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import load_extra, sampled, synth_inputs  # noqa: E402

DEV = torch.device("cuda")
MODES = [("fp32", "simt"), ("bf16", "tcgen05"), ("bf16x6", None)]
# bf16x6 = the reference-precision tensor-core mode (fp32 storage, every GEMM as 6 bf16 cross terms on tcgen05,
# csrc/gemm_split.cu): held to the fp32 tolerances
FP32_CLASS = ("fp32", "bf16x6")


def make_model(tokenizer_dir, enc_layers, dec_layers, precision, gemm, dropout=0.0):
    from model.MMT4Caption import MMT4Caption
    from vct.synthetic import shipped_model_config
    torch.manual_seed(666)
    m = MMT4Caption(shipped_model_config(tokenizer_dir, enc_layers=enc_layers, dec_layers=dec_layers, dropout=dropout),
                    device=DEV).to(DEV)
    m.vct_precision, m.vct_gemm = precision, gemm
    m.mode("caption")
    return m


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("precision,gemm", MODES)
@pytest.mark.parametrize("tag,Le,Ld,B,T,padded", [("bench64", 1, 3, 64, 12, False), ("cfg5", 6, 6, 16, 32, True)])
def test_extra_anchor_forward_backward(tokenizer_dir, tag, Le, Ld, B, T, padded, precision, gemm):
    anchors, smp = load_extra()
    a = anchors[tag]
    model = make_model(tokenizer_dir, Le, Ld, precision, gemm)
    model.train()                                      # dropout p = 0: train-mode plans, eval-mode numbers
    x, vm, tok = synth_inputs(B, T, 512, 21, 30522, 1234, padded=padded)
    xd, vd, td = x.to(DEV), vm.to(DEV), tok.to(DEV)
    mem, _, _ = model.video_encoder([xd], [vd])
    logits, loss = model.cap_decoder(mem, td, td == 0)
    loss.backward()
    torch.cuda.synchronize()
    f32 = precision in FP32_CLASS
    assert abs(float(loss) - a["loss"]) <= (2e-5 if f32 else 2e-3) * a["loss"], (float(loss), a["loss"])
    step = max(1, B // 4)
    mem_s = mem.detach().float().cpu()[::step, ::4, ::16]
    lg_s = logits.float().cpu()[::step, ::5, ::509]
    torch.testing.assert_close(mem_s, torch.from_numpy(smp[f"{tag}/memory_rows"]), rtol=0, atol=2e-4 if f32 else 6e-2)
    torch.testing.assert_close(lg_s, torch.from_numpy(smp[f"{tag}/logits_rows"]), rtol=0, atol=3e-4 if f32 else 6e-2)
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    worst = ("", 0.0)
    for k, want in a["grad_norms"].items():
        e = abs(float(grads[k].double().norm()) - want) / (want + 1e-30)
        if e > worst[1]:
            worst = (k, e)
    assert worst[1] <= (2e-3 if f32 else 5e-2), worst
    n, worst_s = 0, ("", 0.0)
    for key in smp.files:
        if key.startswith(tag + "/grad/"):
            k = key[len(tag + "/grad/"):]
            e = rel_l2(sampled(grads[k].cpu(), smp[f"{tag}/stride/{k}"]), torch.from_numpy(smp[key]))
            print(f"{tag} {precision} sampled grad rel-L2 {k}: {e:.3e}")
            if e > worst_s[1]:
                worst_s = (k, e)
            n += 1
    assert n == 8 and worst_s[1] <= (2e-3 if f32 else 6e-2), worst_s
    if tag == "cfg5" and f32:
        model.eval()
        with torch.no_grad():
            ys = model.greedy_decode_ids([xd], [vd], max_len=6)
        assert ys.cpu().tolist() == a["greedy_ys"]


@pytest.mark.parametrize("precision,gemm", MODES)
def test_eval_fastpath_padded_frames(tokenizer_dir, precision, gemm):
    """val_epoch / eval.py semantics (eval + no_grad + masks): padded memory rows = norm.bias; loss and greedy ids equal
    the reference's fast-path run.  With gradients enabled (train path) the slow-path numbers apply."""
    anchors, smp = load_extra()
    model = make_model(tokenizer_dir, 1, 3, precision, gemm)
    x, vm, tok = synth_inputs(8, 12, 512, 21, 30522, 1234, padded=True, vid_padded=True)
    xd, vd, td = x.to(DEV), vm.to(DEV), tok.to(DEV)
    f32 = precision in FP32_CLASS
    model.eval()
    with torch.no_grad():
        mem, gm, _ = model.video_encoder([xd], [vd])
        loss_fast = float(model([xd], [vd], td))
        ys = model.greedy_decode_ids([xd], [vd], max_len=6)
    torch.testing.assert_close(mem.float().cpu()[:, :, ::16], torch.from_numpy(smp["evalfast/fast/memory"]), rtol=0,
                               atol=2e-4 if f32 else 6e-2)
    full = torch.cat([torch.zeros(8, 1, dtype=torch.bool), vm], 1)
    bias = model.video_encoder.transformer_encoder.norm.bias.detach().float().cpu()
    torch.testing.assert_close(mem.float().cpu()[full], bias.expand(int(full.sum()), -1), rtol=0, atol=1e-6 if f32 else 1e-2)
    assert abs(loss_fast - anchors["evalfast"]["fast"]["loss"]) <= (2e-5 if f32 else 2e-3) * loss_fast
    if f32:
        assert ys.cpu().tolist() == anchors["evalfast"]["fast"]["greedy_ys"]
    model.train()                                      # p = 0; gradients enabled -> the layers' own values in padded rows
    loss_slow = float(model([xd], [vd], td))
    assert abs(loss_slow - anchors["evalfast"]["slow"]["loss"]) <= (2e-5 if f32 else 2e-3) * loss_slow
    assert abs(loss_slow - loss_fast) > 1e-3


@pytest.mark.parametrize("precision,gemm", MODES)
def test_greedy_decode_cfg3_b256(tokenizer_dir, precision, gemm):
    """BASELINE cfg 3 (B = 256, max_len 30).  fp32: ids equal the reference's for every token whose reference
    top-1/top-2 margin is >= 1e-4 (rows are compared up to their first ambiguous step: the smallest margin of the
    7424 argmaxes is 5.7e-6, below what fp32 summation order preserves).  bf16: agreement is REPORTED against the
    margin distribution (SURVEY "Precision vs argmax exact") and must be exact wherever the margin exceeds the bf16
    logit error bound of 6e-2."""
    anchors, smp = load_extra()
    model = make_model(tokenizer_dir, 1, 3, precision, gemm)
    model.eval()
    x, vm, _ = synth_inputs(256, 12, 512, 21, 30522, 1234, padded=False)
    want, margins = torch.from_numpy(smp["decode256/ys"]).long(), torch.from_numpy(smp["decode256/margins"])
    with torch.no_grad():
        ys = model.greedy_decode_ids([x.to(DEV)], [vm.to(DEV)], max_len=30, sync_every=29).cpu()
    assert ys.shape == (256, 30)
    thr = 1e-4 if precision in FP32_CLASS else 6e-2
    exact_rows, tokens_ok, tokens_cmp = 0, 0, 0
    for b in range(256):
        amb = (margins[b] < thr).nonzero()
        upto = int(amb[0]) + 1 if len(amb) else 30
        assert ys[b, :upto].tolist() == want[b, :upto].tolist(), (b, upto)
        exact_rows += int(ys[b].tolist() == want[b].tolist())
        same = (ys[b] == want[b])
        first_diff = int((~same).nonzero()[0]) if not bool(same.all()) else 30
        tokens_ok += first_diff
        tokens_cmp += 30
    print(f"decode256 {precision}: {exact_rows}/256 rows identical to the reference, {tokens_ok}/{tokens_cmp} tokens before the "
          f"first divergence; reference margins: min {float(margins.min()):.2e}, median {float(margins.median()):.2e}, "
          f"share < 1e-4: {float((margins < 1e-4).float().mean()):.4f}, share < 6e-2: {float((margins < 6e-2).float().mean()):.4f}")
    if precision in FP32_CLASS:
        assert exact_rows >= 240
