"""Host launch time vs GPU time of one whole-plan run (debugging aid)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_b200.model import VQAModel
from vqa_b200.synth import synth_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
model = VQAModel().eval().cuda()
_, img, ids, mask = synth_batch(B, 1234, full_length=True)
img, ids, mask = img.cuda(), ids.cuda(), mask.cuda()
with torch.no_grad():
    for _ in range(3):
        model(img, ids, mask)
    torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            model(img, ids, mask)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"rep {rep}: host launch {1e2 * (t1 - t0):.3f} ms/step, wall {1e2 * (t2 - t0):.3f} ms/step, gpu {e0.elapsed_time(e1) / 10:.3f} ms/step")
    eng = model.engine()
    prog, plan = eng.plan_for(B, 20, "nchw_f32", 1, False, 0) if False else (None, None)
