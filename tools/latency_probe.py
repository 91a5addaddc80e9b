"""Where the batch-1 `VQAInference.predict(PIL, str)` latency goes (BASELINE configs[3]): host pieces vs the GPU graph.

    python tools/latency_probe.py
"""
import os
import statistics
import sys
import time

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_b200.inference import VQAInference  # noqa: E402

inf = VQAInference(device="cuda:0")
inf.load()
rng = np.random.default_rng(0)
pil = Image.fromarray(rng.integers(0, 256, (224, 224, 3), dtype=np.uint8))
q = "what color is the car on the left"
for _ in range(20):
    inf.predict(pil, q)


def med(fn, n=300):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts)


print(f"predict(PIL, str)            p50 {med(lambda: inf.predict(pil, q)):.3f} ms")
print(f"  preprocess_image_u8        p50 {med(lambda: inf.preprocess_image_u8(pil)):.3f} ms")
print(f"  preprocess_question        p50 {med(lambda: inf.preprocess_question(q)):.3f} ms")
u8 = inf.preprocess_image_u8(pil).unsqueeze(0)
ids, mask = inf.preprocess_question(q)
print(f"  _run (copies+replay+sync)  p50 {med(lambda: inf._run(u8, ids, mask, 5)):.3f} ms")
g = next(iter(inf._graphs.values()))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for _ in range(200):
    g["graph"].replay()
b.record()
torch.cuda.synchronize()
print(f"  graph replay, GPU time     {a.elapsed_time(b) / 200:.3f} ms per replay (back to back)")
idx, probs = inf._run(u8, ids, mask, 5)
print(f"  _format                    p50 {med(lambda: inf._format(q, idx[0], probs[0])):.3f} ms")
