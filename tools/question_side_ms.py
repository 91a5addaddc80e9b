"""Question side only against cached images (encode_images once, then answer as one CUDA-graph replay): ms per 256 questions.

    python tools/question_side_ms.py [batch]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_b200.model import VQAModel  # noqa: E402
from vqa_b200.synth import synth_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
model = VQAModel().eval().cuda()
_, img, ids, mask = synth_batch(B, 1234, full_length=True)
img, ids, mask = img.cuda(), ids.cuda(), mask.cuda()
with torch.no_grad():
    eng = model.engine()
    cache = eng.encode_images(img)
    for k in (5, 0):
        for _ in range(2):
            eng.answer(cache, ids, mask, top_k=k)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = eng.answer(cache, ids, mask, top_k=k)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        print(f"top_k={k}: {a.elapsed_time(b) / 50:.4f} ms per {B} questions, {len(eng.last_plan.prog.ops) if hasattr(eng.last_plan, 'prog') else '?'} ops")
