"""The op program (layout + tap algebra + weight folding) against the oracle, on the CPU emulator.
This is the check that let the layout design be validated before any GPU time was spent."""
import sys

import pytest
import torch

from conftest import REPO

sys.path.insert(0, REPO)
import emulator as E  # noqa: E402
from oracle import vqa_oracle as O  # noqa: E402
from vqa_b200 import program as P  # noqa: E402
from vqa_b200.model import VQAModel  # noqa: E402
from vqa_b200.synth import randomise_state, synth_batch  # noqa: E402

ABL = dict(use_se_attention=False, use_spatial_attention=False, use_gating=False, num_transformer_layers=1,
           num_cross_layers=1, max_question_length=12, vocab_size=500, num_answers=37)


@pytest.mark.parametrize("ctor,B,L,fmt,mk,window", [
    ({}, 2, 20, "nchw_f32", "i64", True), ({}, 1, 20, "hwc_u8", "f32", False), (ABL, 2, 12, "nchw_f32", "none", True)])
def test_program_matches_oracle(ctor, B, L, fmt, mk, window):
    torch.manual_seed(0)
    model = VQAModel(**ctor).eval()
    sd = randomise_state(model.state_dict(), 1)
    u8, img, ids, mask = synth_batch(B, 1234, max_len=L, vocab=model.config["vocab_size"])
    W = P.build_weights(sd, model.config, "cpu")
    code = {"i64": P.MASK_I64, "f32": P.MASK_F32, "none": P.MASK_NONE}[mk]
    m = {"i64": mask, "f32": mask.float(), "none": None}[mk]
    prog = P.Program(W, model.config, B, L, fmt, code, want_aux=True, top_k=5, device="cpu", window=window)
    logits, ext = E.run_program(prog, u8 if fmt == "hwc_u8" else img, ids, m, top_k=5)
    want, aux = O.vqa_forward(sd, img, ids, m, return_aux=True)
    assert float((logits - want).abs().max() / want.abs().max()) < 1e-2
    feat = prog.tensor("aux.image_features")
    assert float((feat - aux["image_features"]).abs().max() / aux["image_features"].abs().max()) < 3e-2
    assert torch.equal(ext[P.EXT["top_idx"]][:, 0], want.argmax(1))
    kinds = [op.kind for op in prog.ops]
    assert kinds.count("gemm") + kinds.count("stem_pool") + 6 * kinds.count("mlp_chain") >= 30 and kinds[0] == "ingest"
    assert all((op.lane & ~P.LANE_JOIN) in (0, 1) for op in prog.ops) and any(op.lane & 1 for op in prog.ops)
    assert any(op.lane & P.LANE_JOIN for op in prog.ops)


@pytest.mark.parametrize("ctor,B,L", [({}, 2, 20), (ABL, 1, 12)])
def test_tf32_precision_mode_within_1e_3(ctor, B, L):
    """precision="tf32" (BASELINE.json's tolerance mode): fp32 activations, tf32 operands, fp32 accumulation in the
    backbone too; logits within 1e-3 (max-abs relative) of the fp32 oracle."""
    torch.manual_seed(0)
    model = VQAModel(**ctor).eval()
    sd = randomise_state(model.state_dict(), 1)
    u8, img, ids, mask = synth_batch(B, 1234, max_len=L, vocab=model.config["vocab_size"])
    W = P.build_weights(sd, model.config, "cpu", precision="tf32")
    prog = P.Program(W, model.config, B, L, "nchw_f32", P.MASK_I64, want_aux=True, top_k=0, device="cpu")
    assert prog.tf32 and not prog.pair and all(op.i["dtype"] == P.DT_TF32 for op in prog.ops if op.kind == "gemm")
    logits, _ = E.run_program(prog, img, ids, mask)
    want, aux = O.vqa_forward(sd, img, ids, mask, return_aux=True)
    err = float((logits - want).abs().max() / want.abs().max())
    print("tf32 mode max-abs relative logit error", err)
    assert err < 1e-3
    feat = prog.tensor("aux.image_features")
    assert float((feat - aux["image_features"]).abs().max() / aux["image_features"].abs().max()) < 2e-3


def test_one_image_many_questions_program():
    """n_images < B (BASELINE config "one image, many questions"): the image side of the program runs once per image
    and its K/V are shared by B / n_images questions; same logits as the oracle on the repeated image."""
    torch.manual_seed(0)
    ctor = dict(max_question_length=24, num_transformer_layers=1, vocab_size=300, num_answers=40)
    model = VQAModel(**ctor).eval()
    sd = randomise_state(model.state_dict(), 1)
    u8, img, ids, mask = synth_batch(6, 99, max_len=24, vocab=300)
    img2 = img[:2]                                            # 2 images x 3 questions each
    W = P.build_weights(sd, model.config, "cpu")
    prog = P.Program(W, model.config, 6, 24, "nchw_f32", P.MASK_I64, want_aux=False, top_k=0, device="cpu", n_images=2)
    assert prog.Bi == 2 and [op.i["q_per_kv"] for op in prog.ops if op.kind == "cross_attn"] == [3, 3]
    assert next(op for op in prog.ops if op.kind == "ingest").i["B"] == 2
    logits, _ = E.run_program(prog, img2, ids, mask)
    want, _ = O.vqa_forward(sd, img2.repeat_interleave(3, dim=0), ids, mask, return_aux=True)
    assert float((logits - want).abs().max() / want.abs().max()) < 1e-2
    assert torch.equal(logits.argmax(1), want.argmax(1))


def test_weight_folding_is_exact_in_fp32():
    """BN fold + shortcut-as-extra-K reproduce conv->BN (+downsample->BN) exactly up to fp32 rounding."""
    torch.manual_seed(3)
    model = VQAModel().eval()
    sd = randomise_state(model.state_dict(), 2)
    q = "image_encoder.stage2.blocks.0"
    w2, b2 = P._fold_bn(sd, q + ".conv2.weight", q + ".bn2")
    x = torch.randn(1, 128, 6, 6)
    want = O._bn(sd, q + ".bn2", torch.nn.functional.conv2d(x, sd[q + ".conv2.weight"], padding=1))
    got = torch.nn.functional.conv2d(x, w2, b2, padding=1)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
    m = P._stem_matrix(torch.arange(64 * 3 * 49, dtype=torch.float32).view(64, 3, 7, 7))
    assert m.shape == (64, 256) and int((m != 0).sum()) == 64 * 147 - 1   # every tap placed once (one weight is 0)


def test_image_side_and_question_side_programs_compose():
    """side="image" + side="question" (the cache that outlives a call): the image-side program's K/V
    buffers, handed to the question-side program as external slots, give exactly the logits of the whole forward; a
    cached image reused for several questions and a per-question gather of cache rows do too."""
    torch.manual_seed(0)
    ctor = dict(max_question_length=16, num_transformer_layers=1, vocab_size=300, num_answers=40)
    model = VQAModel(**ctor).eval()
    sd = randomise_state(model.state_dict(), 1)
    u8, img, ids, mask = synth_batch(4, 77, max_len=16, vocab=300)
    W = P.build_weights(sd, model.config, "cpu")
    whole = P.Program(W, model.config, 4, 16, "nchw_f32", P.MASK_I64, top_k=0, device="cpu", n_images=2)
    want, _ = E.run_program(whole, img[:2], ids, mask)

    image_side = P.Program(W, model.config, 2, 16, "nchw_f32", P.MASK_NONE, top_k=0, device="cpu", n_images=2, side="image")
    kinds = [op.kind for op in image_side.ops]
    assert kinds[0] == "ingest" and "embed" not in kinds and "cross_attn" not in kinds and "self_attn" not in kinds
    assert all(op.lane == 0 for op in image_side.ops)
    E.Emulator(image_side).run([img[:2]] + [None] * (len(P.EXT) - 1))
    n_layers = image_side.n_cross_layers
    assert n_layers == 2
    kv = [image_side.tensor(f"x.{l}.kv").clone() for l in range(n_layers)]          # [2 * 49, 512] fp32 per layer
    assert kv[0].shape == (2 * 49, 512)

    question_side = P.Program(W, model.config, 4, 16, "nchw_f32", P.MASK_I64, top_k=0, device="cpu", n_images=2,
                              side="question")
    kinds = [op.kind for op in question_side.ops]
    assert "ingest" not in kinds and "stage_tail" not in kinds and kinds.count("cross_attn") == 2
    assert all(op.lane == 0 for op in question_side.ops)
    assert all(isinstance(op.p["kv"], P.ExtRef) for op in question_side.ops if op.kind == "cross_attn")
    got, _ = E.run_program(question_side, None, ids, mask, extra={f"kv{l}": kv[l].view(-1) for l in range(n_layers)})
    assert torch.equal(got, want)

    # per-question gather: questions 0..3 ask about images 1, 0, 0, 1
    pick = torch.tensor([1, 0, 0, 1])
    gathered = {f"kv{l}": kv[l].view(2, 49 * 512)[pick].reshape(-1) for l in range(n_layers)}
    q4 = P.Program(W, model.config, 4, 16, "nchw_f32", P.MASK_I64, top_k=0, device="cpu", n_images=4, side="question")
    got4, _ = E.run_program(q4, None, ids, mask, extra=gathered)
    ref4 = P.Program(W, model.config, 4, 16, "nchw_f32", P.MASK_I64, top_k=0, device="cpu")
    want4, _ = E.run_program(ref4, img[:2][pick], ids, mask)
    # the image side ran at batch 2 here and at batch 4 in ref4: the CPU BLAS behind the emulator blocks its fp32 sums by
    # matrix shape, so a bf16 rounding may flip (the CUDA kernels are batch-invariant: tests/test_gpu_cache_metrics.py
    # checks this composition bit for bit)
    torch.testing.assert_close(got4, want4, atol=5e-3, rtol=0)


def test_fused_and_unfused_program_forms_agree(monkeypatch):
    """Launch fusions change the op list, never the arithmetic: the LayerNorms that share a launch (embedding + first encoder
    norm + mask, final text norm + first query norm, projector norm + key/value norms) and the softmax / top-k in the head
    GEMM's epilogue give bit-identical logits and winners on the emulator, with fewer ops."""
    torch.manual_seed(0)
    model = VQAModel().eval()
    sd = randomise_state(model.state_dict(), 1)
    u8, img, ids, mask = synth_batch(2, 99, max_len=20, vocab=model.config["vocab_size"])
    W = P.build_weights(sd, model.config, "cpu")

    def run(fuse_ln, fuse_topk):
        monkeypatch.setenv("VQA_FUSE_LN", fuse_ln)
        monkeypatch.setenv("VQA_FUSED_TOPK", fuse_topk)
        prog = P.Program(W, model.config, 2, 20, "nchw_f32", P.MASK_I64, want_aux=False, top_k=5, device="cpu")
        logits, ext = E.run_program(prog, img, ids, mask, top_k=5)
        return logits.clone(), ext[P.EXT["top_idx"]].clone(), ext[P.EXT["top_probs"]].clone(), [op.kind for op in prog.ops]

    base = run("0", "0")
    ln = run("1", "0")
    tk = run("1", "1")
    assert torch.equal(base[0], ln[0]) and torch.equal(base[0], tk[0])
    assert torch.equal(base[1], ln[1]) and torch.equal(base[1], tk[1])
    torch.testing.assert_close(base[2], tk[2], rtol=1e-6, atol=0)
    assert len(ln[3]) == len(base[3]) - 5 and "mask_prep" not in ln[3]          # 3 + 2 + 3 launches became 1 + 1 + 1
    assert len(tk[3]) == len(ln[3]) - 1 and "softmax_topk" not in tk[3]
