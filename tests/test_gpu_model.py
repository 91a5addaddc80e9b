"""End-to-end parity of the sm_100a VQAModel with the oracle / golden vectors of the reference."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import GOLDEN, REPO  # noqa: E402

sys.path.insert(0, REPO)
from oracle import vqa_oracle as O  # noqa: E402
from vqa_b200.model import VQAModel, create_vqa_model, load_vqa_model  # noqa: E402
from vqa_b200.synth import randomise_state, synth_batch  # noqa: E402

# BASELINE.json north_star: bf16 mode, logits within 2e-2 max-abs relative error
LOGIT_REL_TOL = 2e-2


def rel_err(got, want):
    return float((got - want).abs().max() / want.abs().max())


def make(meta, precision="bf16"):
    torch.manual_seed(meta["seed_weights"])
    model = VQAModel(**meta["ctor"], precision=precision).eval()
    sd = model.state_dict()
    if meta["randomise"]:
        sd = randomise_state(sd, 1)
        model.load_state_dict(sd, strict=True)
    u8, img, ids, mask = synth_batch(meta["batch"], meta["seed_inputs"], max_len=meta["max_len"], vocab=meta["vocab"])
    return model.cuda(), sd, u8, img, ids, mask


@pytest.mark.parametrize("case", ["default_b4", "plain_b2", "ablate_b3", "nospatial_b2"])
def test_forward_matches_reference_golden(golden_meta, case):
    meta = golden_meta[case]
    model, sd, u8, img, ids, mask = make(meta)
    g = np.load(os.path.join(GOLDEN, f"{case}.npz"))
    with torch.no_grad():
        logits, aux = model(img.cuda(), ids.cuda(), mask.cuda(), return_aux=True)
    torch.cuda.synchronize()
    assert logits.dtype == torch.float32 and tuple(logits.shape) == g["logits"].shape
    errs = {"logits": rel_err(logits.cpu(), torch.from_numpy(g["logits"]))}
    for k in ("image_features", "text_features", "fused", "image_projected", "attended_pooled", "text_pooled"):
        errs[k] = rel_err(aux[k].cpu(), torch.from_numpy(g[k]))
    for i, w in enumerate(aux["cross_attention_weights"]):
        errs[f"xattn{i}"] = rel_err(w.cpu(), torch.from_numpy(g[f"cross_attention_weights_{i}"]))
    print(case, {k: f"{v:.2e}" for k, v in errs.items()})
    assert errs["logits"] <= LOGIT_REL_TOL, errs
    assert errs["image_features"] <= 5e-2, errs      # bf16 backbone, 17 conv layers deep
    assert errs["text_features"] <= 1e-2 and errs["fused"] <= 5e-2, errs
    top_idx, top_p = model.predict(img.cuda(), ids.cuda(), mask.cuda(), top_k=5)
    assert top_idx.dtype == torch.int64 and tuple(top_idx.shape) == (meta["batch"], 5)
    assert np.array_equal(top_idx[:, 0].cpu().numpy(), g["top_indices"][:, 0])
    np.testing.assert_allclose(top_p.cpu().numpy(), g["top_probs"], rtol=2e-2, atol=1e-4)


@pytest.mark.parametrize("case", ["default_b4", "plain_b2", "ablate_b3", "nospatial_b2"])
def test_tf32_tolerance_mode_within_1e_3(golden_meta, case):
    """BASELINE.json: "in fp32-accumulate TF32 mode, logits must be within 1e-3" (max-abs relative) of the fp32
    reference.  precision="tf32": fp32 activations and tf32 operands in the backbone, 3xTF32 Linears in the tail."""
    meta = golden_meta[case]
    model, sd, u8, img, ids, mask = make(meta, precision="tf32")
    g = np.load(os.path.join(GOLDEN, f"{case}.npz"))
    with torch.no_grad():
        logits, aux = model(img.cuda(), ids.cuda(), mask.cuda(), return_aux=True)
        logits_u8, _ = model(u8.cuda(), ids.cuda(), mask.cuda())
    errs = {"logits": rel_err(logits.cpu(), torch.from_numpy(g["logits"])),
            "image_features": rel_err(aux["image_features"].cpu(), torch.from_numpy(g["image_features"])),
            "text_features": rel_err(aux["text_features"].cpu(), torch.from_numpy(g["text_features"])),
            "fused": rel_err(aux["fused"].cpu(), torch.from_numpy(g["fused"]))}
    print("tf32", case, {k: f"{v:.2e}" for k, v in errs.items()})
    assert errs["logits"] <= 1e-3, errs
    assert errs["image_features"] <= 2e-3 and errs["text_features"] <= 1e-4 and errs["fused"] <= 1e-3, errs
    assert torch.equal(logits_u8, logits)
    assert np.array_equal(logits.argmax(1).cpu().numpy(), g["top_indices"][:, 0])


def test_large_answer_and_word_vocabularies():
    """VQA-v2-sized head (3129 answers: wider than the epilogue's 2048-entry bias table) and a 20k-word embedding."""
    torch.manual_seed(0)
    model = VQAModel(vocab_size=20000, num_answers=3129, num_transformer_layers=1, num_cross_layers=3, se_reduction=4).eval()
    sd = randomise_state(model.state_dict(), 1)
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    u8, img, ids, mask = synth_batch(3, 5, vocab=20000)
    with torch.no_grad():
        logits, _ = model(img.cuda(), ids.cuda(), mask.cuda())
    want, _ = O.vqa_forward(sd, img, ids, mask)
    assert tuple(logits.shape) == (3, 3129) and rel_err(logits.cpu(), want) <= LOGIT_REL_TOL
    assert torch.equal(logits.argmax(1).cpu(), want.argmax(1))


def test_one_image_many_questions():
    """BASELINE configs[4]: one image, 16 questions of 64 tokens on VQAModel(max_question_length=64); the backbone,
    projector and K/V projections run once per image.  Must equal the plain forward on the repeated image bit for
    bit, and match the fp32 oracle within the bf16 gate."""
    torch.manual_seed(0)
    model = VQAModel(max_question_length=64).eval()
    sd = randomise_state(model.state_dict(), 1)
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    u8, img, ids, mask = synth_batch(16, 77, max_len=64)
    one = img[:1]
    with torch.no_grad():
        shared, aux = model(one.cuda(), ids.cuda(), mask.cuda(), return_aux=True)
        plain, _ = model(one.repeat(16, 1, 1, 1).cuda(), ids.cuda(), mask.cuda())
        two, _ = model(img[:2].cuda(), ids.cuda(), mask.cuda())          # 2 images x 8 questions
        plain2, _ = model(img[:2].repeat_interleave(8, dim=0).cuda(), ids.cuda(), mask.cuda())
    assert tuple(shared.shape) == (16, 1000) and tuple(aux["image_features"].shape) == (1, 512, 7, 7)
    assert torch.equal(shared, plain) and torch.equal(two, plain2)
    want, _ = O.vqa_forward(sd, one.repeat(16, 1, 1, 1), ids, mask)
    assert rel_err(shared.cpu(), want) <= LOGIT_REL_TOL
    idx, probs = model.predict(one.cuda(), ids.cuda(), mask.cuda(), top_k=3)
    assert tuple(idx.shape) == (16, 3)
    with pytest.raises(ValueError):
        model(img[:3].cuda(), ids.cuda(), mask.cuda())                    # 16 questions over 3 images


def test_uint8_input_equals_normalised_input(golden_meta):
    model, sd, u8, img, ids, mask = make(golden_meta["plain_b2"])
    with torch.no_grad():
        a, _ = model(img.cuda(), ids.cuda(), mask.cuda())
        b, _ = model(u8.cuda(), ids.cuda(), mask.cuda())
    assert torch.equal(a, b)  # same fp32 normalisation arithmetic, same bf16 rounding


def test_mask_variants_and_padding_invariance(golden_meta):
    model, sd, u8, img, ids, mask = make(golden_meta["plain_b2"])
    x, t, m = img.cuda(), ids.cuda(), mask.cuda()
    with torch.no_grad():
        a, _ = model(x, t, m)
        b, _ = model(x, t, m.float())
        c, _ = model(x, t, m.int())
        d, _ = model(x, torch.where(m == 0, torch.full_like(t, 17), t), m)
        full, _ = model(x, t, None)
        want_full, _ = O.vqa_forward(sd, img, ids, None)
    assert torch.equal(a, b) and torch.equal(a, c) and torch.equal(a, d)
    assert rel_err(full.cpu(), want_full) <= LOGIT_REL_TOL


def test_top1_agreement_2000_pairs():
    """BASELINE.json: >= 99 % top-1 agreement with the fp32 reference (plain seeded random init).
    2000 pairs here to keep the suite short; bench.py --agreement runs the full 10 000."""
    torch.manual_seed(0)
    model = VQAModel().eval()
    sd = model.state_dict()
    model = model.cuda()
    agree = n = 0
    worst = 0.0
    for b in range(8):
        _, img, ids, mask = synth_batch(250, 1234 + b)
        with torch.no_grad():
            got, _ = model(img.cuda(), ids.cuda(), mask.cuda())
        want, _ = O.vqa_forward(sd, img, ids, mask)
        got = got.cpu()
        agree += int((got.argmax(1) == want.argmax(1)).sum())
        n += 250
        worst = max(worst, rel_err(got, want))
    print(f"top-1 agreement {agree}/{n} = {100.0 * agree / n:.2f}%  worst logit rel err {worst:.2e}")
    assert worst <= LOGIT_REL_TOL
    assert agree / n >= 0.99


def test_batch_sizes_and_lengths():
    torch.manual_seed(0)
    model = VQAModel().eval()
    sd = model.state_dict()
    model = model.cuda()
    for B, L in ((1, 20), (5, 7), (33, 20), (130, 1)):
        _, img, ids, mask = synth_batch(B, 99 + B, max_len=L)
        with torch.no_grad():
            got, _ = model(img.cuda(), ids.cuda(), mask.cuda())
        want, _ = O.vqa_forward(sd, img, ids, mask)
        assert rel_err(got.cpu(), want) <= LOGIT_REL_TOL, (B, L)
    with pytest.raises(RuntimeError):
        _, img, ids, mask = synth_batch(1, 1, max_len=21)
        model(img.cuda(), ids.cuda(), mask.cuda())  # L > max_question_length, like the reference's PE buffer


def test_api_surface_and_errors(tmp_path):
    torch.manual_seed(0)
    model = create_vqa_model(vocab_size=300, num_answers=20, use_attention=False)
    assert model.config["use_se_attention"] is False and model.num_answers == 20
    with pytest.raises(Exception):
        model.eval()(torch.zeros(1, 3, 224, 224), torch.zeros(1, 20, dtype=torch.long))  # CPU tensors: no fallback
    model = model.cuda()
    with pytest.raises(NotImplementedError):
        model.train()(torch.zeros(1, 3, 224, 224).cuda(), torch.zeros(1, 20, dtype=torch.long).cuda())
    model.eval()
    path = str(tmp_path / "ckpt.pt")
    torch.save({"config": model.config, "model_state_dict": model.state_dict(), "epoch": 3}, path)
    again = load_vqa_model(path, "cuda")
    _, img, ids, mask = synth_batch(2, 5, vocab=300)
    with torch.no_grad():
        a, _ = model(img.cuda(), ids.cuda(), mask.cuda())
        b, _ = again.eval()(img.cuda(), ids.cuda(), mask.cuda())
    assert torch.equal(a, b)
    maps = model.get_attention_maps(img.cuda(), ids.cuda(), mask.cuda())
    assert tuple(maps["cross_attention_spatial"].shape) == (2, 20, 7, 7)
    with pytest.raises(NotImplementedError):
        VQAModel(embed_dim=32, num_attention_heads=4).cuda().eval()(img.cuda(), ids.cuda())
