"""Two / three compute lanes replaying different captured forwards (fp32 NCHW input, uint8 HWC input, uint8 + top-5):
which part of the end-to-end graph costs what against the bench's device-resident figure.

    python tools/lane_probe.py [lanes]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_b200.model import VQAModel  # noqa: E402
from vqa_b200.synth import synth_batch  # noqa: E402

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
model = VQAModel().eval().cuda()
eng = model.engine()
u8, img, ids, mask = synth_batch(256, 1234, full_length=True)
u8, img, ids, mask = u8.cuda(), img.cuda(), ids.cuda(), mask.cuda()
streams = [torch.cuda.Stream() for _ in range(lanes)]
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40


def measure(name, fn):
    graphs = []
    with torch.no_grad():
        for l in range(lanes):
            with torch.cuda.stream(streams[l]):
                for _ in range(2):
                    fn(l)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=streams[l]):
                out = fn(l)
            graphs.append((g, out))
    cur = torch.cuda.current_stream()
    import time
    for rep in range(2):
        torch.cuda.synchronize()
        time.sleep(0.5)                       # let the power-averaging window drain: every repetition is a burst measurement
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(cur)
        for st in streams:
            st.wait_stream(cur)
        for i in range(K):
            with torch.cuda.stream(streams[i % lanes]):
                graphs[i % lanes][0].replay()
        for st in streams:
            cur.wait_stream(st)
        b.record(cur)
        torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / K:.4f} ms/step, {256 * K / a.elapsed_time(b) * 1e3:.0f} pairs/s ({lanes} lanes)")


u8s = [u8.clone() for _ in range(lanes)]
imgs = [img.clone() for _ in range(lanes)]
measure("fp32 NCHW forward        ", lambda l: eng.run(img, ids, mask, slot=l)[0])
measure("fp32, own input per lane ", lambda l: eng.run(imgs[l], ids, mask, slot=l)[0])
measure("uint8 predict, own input ", lambda l: eng.predict(u8s[l], ids, mask, 5, slot=l))
measure("uint8 HWC forward        ", lambda l: eng.run(u8, ids, mask, slot=l)[0])
measure("uint8 HWC predict (top-5)", lambda l: eng.predict(u8, ids, mask, 5, slot=l))
