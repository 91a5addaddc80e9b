"""Achieved HBM bandwidth of the accuracy kernel (csrc/metrics.cu) against the measured copy bandwidth.

    python tools/accuracy_probe.py        # logits larger than the 126 MB L2, CUDA events, 20 launches after 5 warm-ups
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_b200.runtime import accuracy_update  # noqa: E402

peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
out = {}
for B, N in ((65536, 1000), (32768, 3129), (256, 1000)):
    logits = torch.randn(B, N, device="cuda")
    targets = torch.randint(0, N, (B,), device="cuda")
    counters = torch.zeros(3, dtype=torch.int64, device="cuda")
    for _ in range(5):
        accuracy_update(logits, targets, counters)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        accuracy_update(logits, targets, counters)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / 20 * 1e3
    nbytes = B * N * 4 + B * 8
    out[f"{B}x{N}"] = {"us_per_launch": us, "algorithmic_bytes": nbytes, "achieved_gbps": nbytes / us / 1e3,
                       "rows_per_sec": B / us * 1e6}
    want = int((logits.argmax(-1) == targets).sum()) * 25
    assert int(counters[0]) == want and int(counters[2]) == 25 * B
print(json.dumps({"kernel": "accuracy_kernel", "bound": "hbm", "peaks": peaks, "shapes": out}))
