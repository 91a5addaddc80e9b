"""Generate the golden vectors under tests/golden/ by running the REAL reference.

Run in the build container only (needs /root/reference):

    cd /tmp && python /root/repo/tests/golden/make_golden.py

The reference ships no numerical golden vectors of its own, so these
files pin the oracle (oracle/vqa_oracle.py) and, through it, the CUDA path.  Weights are not
stored: they are a pure function of ``torch.manual_seed`` + the module construction order,
which ``vqa_b200.modules`` reproduces; ``state_fingerprint`` values are stored so a drift in
that reproduction is caught on machines that do not have the reference.
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VQA_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from vqa_b200.synth import (randomise_state, state_fingerprint, synth_batch)  # noqa: E402


def model_case(name, ctor_kwargs, batch, seed_w, seed_in, randomise, max_len, vocab):
    from models.vqa_model import VQAModel  # the reference
    torch.manual_seed(seed_w)
    ref = VQAModel(**ctor_kwargs).eval()
    sd = ref.state_dict()
    if randomise:
        sd = randomise_state(sd, seed=1)
        ref.load_state_dict(sd, strict=True)
    _, images, ids, mask = synth_batch(batch, seed_in, max_len=max_len, vocab=vocab)
    with torch.no_grad():
        logits, aux = ref(images, ids, mask, return_aux=True)
        # per-stage taps through the reference's own modules
        x = ref.image_encoder.stem(images)
        taps = {"stem": x}
        for s in (1, 2, 3, 4):
            stage = getattr(ref.image_encoder, f"stage{s}")
            y = stage.blocks(x)
            taps[f"stage{s}.blocks"] = y
            x = stage.attention(y)
            taps[f"stage{s}"] = x
    out = {
        "logits": logits.numpy(),
        "image_features": aux["image_features"].numpy(),
        "text_features": aux["text_features"].numpy(),
        "text_pooled_encoder": ref.text_encoder(ids, mask)[1].detach().numpy(),
        "fused": aux["fused"].numpy(),
        "image_projected": aux["image_projected"].numpy(),
        "attended_pooled": aux["attended_pooled"].numpy(),
        "text_pooled": aux["text_pooled"].numpy(),
    }
    for i, w in enumerate(aux["cross_attention_weights"]):
        out[f"cross_attention_weights_{i}"] = w.numpy()
    for k, v in taps.items():  # big tensors: keep moments + a strided sample
        v64 = v.double()
        out[f"tap.{k}.moments"] = np.array([float(v64.sum()), float(v64.abs().sum()),
                                            float((v64 * v64).sum()), float(v64.max())])
        out[f"tap.{k}.sample"] = v.flatten()[::997].numpy().copy()
    top_idx, top_p = ref.predict(images, ids, mask, top_k=5)
    out["top_indices"] = top_idx.numpy()
    out["top_probs"] = top_p.numpy()
    np.savez(os.path.join(HERE, f"{name}.npz"), **out)
    meta = {"ctor": ctor_kwargs, "batch": batch, "seed_weights": seed_w, "seed_inputs": seed_in,
            "randomise": randomise, "max_len": max_len, "vocab": vocab,
            "fingerprint": state_fingerprint(sd),
            "num_parameters": ref.get_num_parameters(), "state_keys": len(sd)}
    return meta


def text_utils_case():
    from utils.tokenizer import Tokenizer
    from data.build_vocab import AnswerVocabulary
    corpus = ["What color is the dog?", "How many people are there?", "Is this a cat?",
              "What is the man doing?", "Where is the ball?"]
    tok = Tokenizer(max_length=15)
    tok.build_vocab(corpus, min_freq=1)
    probes = ["What color is the dog?", "", "   ", "What's that -- over THERE?!", "don't  know\tTABS\nnewline",
              "a " * 40, "naïve café ünïcode ok", "under_score 123 4.5", "what what what color color is",
              "?!?", "Is   this,a;cat", "the quick brown fox jumps over the lazy dog again and again and again ok"]
    enc = {p: tok.encode(p) for p in probes}
    enc_nospecial = {p: tok.encode(p, add_special_tokens=False) for p in probes}
    enc_nopad = {p: tok.encode(p, padding=False, truncation=False) for p in probes}
    # default tokenizer of api/inference.py:114-119
    dtok = Tokenizer(max_length=20)
    dtok.build_vocab(["what is this", "what color", "how many", "is there", "where is", "what type"], min_freq=1)
    dprobes = ["What COLOR is this, really?!  don't know", "how many are there", "where is what type"]
    # frequency-ordered vocab with ties and min_freq filtering
    tok2 = Tokenizer(max_length=8, vocab_size=9)
    tok2.build_vocab(["b a a c", "c b a d", "e e d d d", "f"], min_freq=2)
    av = AnswerVocabulary(num_answers=5)
    av.build_from_qa_pairs([{"answer": a} for a in
                            ["Yes", "yes", "The dog", "a dog", "no", "No.", "two", "2", "yes", "red!", "RED"]])
    answers_probe = ["yes", "YES", "unknown", "the dog", "A Dog!", "no", "red", "an apple"]
    return {
        "corpus": corpus, "word2idx": tok.word2idx, "probes": probes,
        "encode": {p: [list(v[0]), list(v[1])] for p, v in enc.items()},
        "encode_nospecial": {p: [list(v[0]), list(v[1])] for p, v in enc_nospecial.items()},
        "encode_nopad": {p: [list(v[0]), list(v[1])] for p, v in enc_nopad.items()},
        "decode": {p: tok.decode(enc[p][0]) for p in probes},
        "default_word2idx": dtok.word2idx,
        "default_encode": {p: [list(dtok.encode(p)[0]), list(dtok.encode(p)[1])] for p in dprobes},
        "tok2_word2idx": tok2.word2idx,
        "answer2idx": av.answer2idx,
        "answer_encode": {a: av.encode(a) for a in answers_probe},
        "answer_decode": {str(i): av.decode(i) for i in range(-1, 8)},
        "answer_preprocess": {a: av.preprocess_answer(a) for a in
                              ["The  quick, brown fox!", "An apple a day", "theatre", " a ", "it's"]},
    }


def preprocess_case():
    """The reference transform on 224x224 and on non-224 PIL images (PIL antialiased bilinear)."""
    from PIL import Image
    from data.preprocess import get_inference_transforms
    tf = get_inference_transforms(224)
    g = torch.Generator().manual_seed(77)
    out = {}
    for name, (h, w) in {"id224": (224, 224), "down": (300, 400), "up": (100, 160), "odd": (333, 211)}.items():
        u8 = torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8)
        # smooth the noise a little so resampling differences are meaningful, keep uint8
        img = Image.fromarray(u8.numpy(), "RGB")
        t = tf(img)
        out[f"{name}.u8"] = u8.numpy()
        if name == "id224":
            out[f"{name}.out"] = t.numpy()
        else:  # the float output is a pure function of the resized uint8 image; keep a checksum
            t64 = t.double()
            out[f"{name}.out_moments"] = np.array([float(t64.sum()), float(t64.abs().sum())])
        resized = np.asarray(img.resize((224, 224), Image.BILINEAR))
        out[f"{name}.resized_u8"] = resized
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)


def main():
    os.chdir(tempfile.mkdtemp())  # utils.config makes directories relative to CWD
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    meta = {"torch": torch.__version__}
    meta["default_b4"] = model_case("default_b4", {}, 4, 0, 1234, True, 20, 10000)
    meta["plain_b2"] = model_case("plain_b2", {}, 2, 0, 4321, False, 20, 10000)
    meta["ablate_b3"] = model_case(
        "ablate_b3", dict(use_se_attention=False, use_spatial_attention=False, use_gating=False,
                          num_transformer_layers=2, num_cross_layers=1, max_question_length=12,
                          vocab_size=500, num_answers=37), 3, 5, 99, True, 12, 500)
    meta["nospatial_b2"] = model_case(
        "nospatial_b2", dict(use_spatial_attention=False, max_question_length=64, num_cross_layers=3),
        2, 3, 7, True, 64, 10000)
    meta["text_utils"] = text_utils_case()
    preprocess_case()
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True, ensure_ascii=False)
    print("golden written to", HERE)


if __name__ == "__main__":
    main()
