"""Throughput of the device-resident forward at several batch sizes (uint8 HWC input, GPU normalise included)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_b200.model import VQAModel
from vqa_b200.synth import synth_batch

torch.manual_seed(0)
model = VQAModel().eval().cuda()
for B in [int(a) for a in sys.argv[1:]] or [256, 512, 1024]:
    u8, _, ids, mask = synth_batch(B, 1234, full_length=True)
    u8, ids, mask = u8.cuda(), ids.cuda(), mask.cuda()
    with torch.no_grad():
        for _ in range(2):
            model(u8, ids, mask)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            logits, _ = model(u8, ids, mask)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            g.replay()
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"batch {B}: {ms:.3f} ms/step, {B / ms * 1e3:.0f} pairs/s, finite={bool(torch.isfinite(logits).all())}", flush=True)
    del g
    model._engine = None
    torch.cuda.empty_cache()
