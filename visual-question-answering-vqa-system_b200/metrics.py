"""Evaluation on the fused forward with device-resident accuracy counters.

``VQAAccuracy`` keeps the reference class's interface (utils/metrics.py:25-136: ``reset`` / ``update(predictions,
targets, question_types=None)`` / ``compute`` / ``__str__``, same result keys) but its counters live in HBM and are
updated by one kernel per batch (``vqa_accuracy_update``); nothing is copied to the host and nothing synchronises until
``compute()``.  ``evaluate`` is the loop of ``Evaluator.evaluate`` (training/evaluate.py:77-139) and
``Trainer.validate`` (training/train.py:229-264) on top of it.  CUDA only: there is no CPU path.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch

from .runtime import accuracy_update


class VQAAccuracy:
    """Top-1 / top-5 (``k``) accuracy; a target outside [0, num_classes) is never correct but counts in the total
    (``AnswerVocabulary.encode`` returns -1 for unknown answers).  Ties go to the lower index."""

    def __init__(self, k: int = 5, keep_predictions: bool = False):
        self.k = int(k)
        self.keep_predictions = keep_predictions
        self.reset()

    def reset(self):
        self._counters: Optional[torch.Tensor] = None      # int64 [3] on the device: top-1, top-k, total
        self._had_logits = False
        self._typed: List[tuple] = []                      # (question types, device rank tensor) per typed batch
        self._preds: List[torch.Tensor] = []
        self._targets: List[torch.Tensor] = []

    def update(self, predictions: torch.Tensor, targets: torch.Tensor, question_types: Optional[List[str]] = None):
        if self._counters is None:
            self._counters = torch.zeros(3, dtype=torch.int64, device=predictions.device)
        B = int(targets.shape[0])
        targets = targets.to(predictions.device).long()
        rank = torch.empty(B, dtype=torch.int32, device=predictions.device) if question_types is not None else None
        pred = torch.empty(B, dtype=torch.int64, device=predictions.device) if self.keep_predictions else None
        accuracy_update(predictions, targets, self._counters, k=self.k, pred_out=pred, rank_out=rank)
        self._had_logits = self._had_logits or predictions.dim() == 2
        if question_types is not None:
            if len(question_types) != B:
                raise ValueError("question_types must have one entry per target")
            self._typed.append((list(question_types), rank))
        if pred is not None:
            self._preds.append(pred)
            self._targets.append(targets)

    # -- the only host synchronisation
    def compute(self) -> Dict:
        if self._counters is None:
            c1 = ck = total = 0
        else:
            c1, ck, total = (int(v) for v in self._counters.cpu().tolist())
        out = {"accuracy": c1 / max(total, 1), "accuracy_top5": (ck if self._had_logits else 0) / max(total, 1),
               "correct": c1, "total": total}
        if self._typed:
            right: Dict[str, int] = {}
            seen: Dict[str, int] = {}
            for types, rank in self._typed:
                for t, r in zip(types, rank.cpu().tolist()):
                    seen[t] = seen.get(t, 0) + 1
                    right[t] = right.get(t, 0) + (1 if r == 0 else 0)
            out["per_type"] = {t: right[t] / max(seen[t], 1) for t in seen}
        return out

    def predictions(self):
        """(predicted indices, targets) of every update so far, on the host (``keep_predictions=True``)."""
        if not self._preds:
            return torch.empty(0, dtype=torch.long), torch.empty(0, dtype=torch.long)
        return torch.cat(self._preds).cpu(), torch.cat(self._targets).cpu()

    def __str__(self) -> str:
        m = self.compute()
        return f"Accuracy: {m['accuracy']:.4f} | Top-5: {m['accuracy_top5']:.4f}"


def compute_confusion_matrix(predictions: torch.Tensor, targets: torch.Tensor, num_classes: int) -> torch.Tensor:
    """[num_classes, num_classes] counts, rows = target, columns = prediction (utils/metrics.py:213-235); pairs with an
    index outside the range are skipped."""
    ok = (targets >= 0) & (targets < num_classes) & (predictions >= 0) & (predictions < num_classes)
    flat = targets[ok] * num_classes + predictions[ok]
    return torch.bincount(flat, minlength=num_classes * num_classes).view(num_classes, num_classes)


def get_per_class_accuracy(conf_matrix: torch.Tensor) -> torch.Tensor:
    """diagonal / row sum, 0 for classes that never occur (utils/metrics.py:237-253)."""
    rows = conf_matrix.sum(dim=1).clamp(min=1).float()
    return conf_matrix.diag().float() / rows


def evaluate(model, batches: Iterable[Dict[str, torch.Tensor]], device=None, answer_vocab=None, top_errors: int = 10) -> Dict:
    """``Evaluator.evaluate`` on the fused forward.  ``batches`` yields the reference loader's dicts (``images``,
    ``token_ids``, ``attention_mask``, ``answers``; data/dataset.py collate).  Per batch: H2D of the inputs, the forward,
    one accuracy kernel; predictions come back in ONE copy at the end (the reference does three ``.cpu()`` per batch)."""
    device = device or next(model.parameters()).device
    model.eval()
    acc = VQAAccuracy(keep_predictions=True)
    with torch.no_grad():
        for batch in batches:
            logits, _ = model(batch["images"].to(device, non_blocking=True), batch["token_ids"].to(device, non_blocking=True),
                              batch["attention_mask"].to(device, non_blocking=True))
            acc.update(logits, batch["answers"].to(device, non_blocking=True))
    m = acc.compute()
    preds, targets = acc.predictions()
    num_answers = int(getattr(model, "num_answers", int(preds.max()) + 1 if preds.numel() else 1))
    conf = compute_confusion_matrix(preds, targets, num_answers)
    pairs: Dict[tuple, int] = {}
    for p, t in zip(preds.tolist(), targets.tolist()):
        if p != t:
            pairs[(p, t)] = pairs.get((p, t), 0) + 1
    errors = []
    for (p, t), n in sorted(pairs.items(), key=lambda kv: kv[1], reverse=True)[:top_errors]:
        e = {"predicted_idx": p, "target_idx": t, "count": n}
        if answer_vocab is not None:
            e["predicted"], e["target"] = answer_vocab.decode(p), answer_vocab.decode(t)
        errors.append(e)
    return {"accuracy": m["accuracy"], "accuracy_top5": m["accuracy_top5"], "total_samples": m["total"],
            "correct": m["correct"], "per_class_accuracy": get_per_class_accuracy(conf)[: min(100, num_answers)].tolist(),
            "common_errors": errors}
