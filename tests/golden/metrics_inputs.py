"""Seeded inputs of the accuracy-counter golden vectors (shared by make_metrics_golden.py and the tests)."""
import torch

CASES = [(1, 1000, 11), (7, 37, 12), (256, 1000, 13), (1000, 3129, 14), (33, 5, 15)]
QTYPES = ["yes/no", "number", "other"]


def make(B: int, N: int, seed: int):
    """logits [B, N] fp32, targets [B] int64 (half of them the argmax, a quarter inside the top 5, some unknown = -1,
    like the reference's AnswerVocabulary.encode for an unseen answer), question types [B]."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, N, generator=g)
    top = logits.topk(min(5, N), -1).indices
    targets = torch.randint(0, N, (B,), generator=g)
    sel = torch.rand(B, generator=g)
    targets = torch.where(sel < 0.5, top[:, 0], targets)
    targets = torch.where((sel >= 0.5) & (sel < 0.75), top[:, min(3, top.shape[1] - 1)], targets)
    targets = torch.where(sel > 0.95, torch.full_like(targets, -1), targets)
    qtypes = [QTYPES[int(i)] for i in torch.randint(0, 3, (B,), generator=g)]
    return logits, targets, qtypes
