"""Host-side cost of the pieces of one pipelined step (graph launch, H2D enqueue, events): is the e2e loop host-bound?

    python tools/host_overhead_probe.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_b200.model import VQAModel  # noqa: E402
from vqa_b200.synth import synth_batch  # noqa: E402

torch.manual_seed(0)
model = VQAModel().eval().cuda()
u8, _, ids, mask = synth_batch(256, 1234, full_length=True)
h_u8, h_ids, h_mask = u8.pin_memory(), ids.pin_memory(), mask.pin_memory()
d_u8, d_ids, d_mask = u8.cuda(), ids.cuda(), mask.cuda()
eng = model.engine()
with torch.no_grad():
    for _ in range(2):
        eng.predict(d_u8, d_ids, d_mask, 5)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = eng.predict(d_u8, d_ids, d_mask, 5)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()


def host_us(fn, n=50):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6


print("graph.replay(): host %.1f us per call, %.1f us per call incl. GPU drain" % host_us(g.replay))
print("H2D 38.5 MB pinned copy_ enqueue: host %.1f us, %.1f us incl. drain" % host_us(lambda: d_u8.copy_(h_u8, non_blocking=True)))
print("small H2D copy_ enqueue: host %.1f us" % host_us(lambda: d_ids.copy_(h_ids, non_blocking=True))[0])
ev = torch.cuda.Event()
print("event record + stream wait: host %.1f us" % host_us(lambda: (ev.record(), torch.cuda.current_stream().wait_event(ev)))[0])
h_idx = torch.empty(256, 5, dtype=torch.long).pin_memory()
print("small D2H copy_ enqueue: host %.1f us" % host_us(lambda: h_idx.copy_(out[0], non_blocking=True))[0])
with torch.no_grad():
    print("eager engine.predict (87 launches through ctypes): host %.1f us, %.1f incl. drain"
          % host_us(lambda: eng.predict(d_u8, d_ids, d_mask, 5), 20))
