"""Golden vectors of SURVEY 8f row f4 from the REAL reference: utils/metrics.py::VQAAccuracy (update :56-105, compute
:107-134) run on the seeded inputs of metrics_inputs.py.  Build container only:

    cd /tmp && python /root/repo/tests/golden/make_metrics_golden.py
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VQA_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(0, REF)

import metrics_inputs as MI  # noqa: E402


def main():
    os.chdir(tempfile.mkdtemp())           # utils.config makes directories relative to CWD
    from utils.metrics import VQAAccuracy  # the reference
    out = []
    for B, N, seed in MI.CASES:
        logits, targets, qtypes = MI.make(B, N, seed)
        acc = VQAAccuracy()
        acc.update(logits, targets, qtypes)
        acc.update(logits.flip(0), targets.flip(0), list(reversed(qtypes)))      # counters accumulate over batches
        res = acc.compute()
        idx = VQAAccuracy()
        idx.update(logits.argmax(-1), targets)                                    # the 1-D (indices) form: top-1 only
        out.append({"B": B, "N": N, "seed": seed, "correct": acc.correct, "correct_top5": acc.correct_top5,
                    "total": acc.total, "accuracy": res["accuracy"], "accuracy_top5": res["accuracy_top5"],
                    "per_type": res.get("per_type", {}), "index_form": [idx.correct, idx.correct_top5, idx.total]})
    with open(os.path.join(HERE, "metrics_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("written", len(out), "cases")


if __name__ == "__main__":
    main()
