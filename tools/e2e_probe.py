"""What the end-to-end loop loses against the device-resident figure: predict_tensors_pipelined fed from pinned host
memory (the bench's e2e leg), from device memory (same loop, D2D instead of H2D copies) and without the result copies.

    python tools/e2e_probe.py [lanes]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_b200.inference import VQAInference  # noqa: E402
from vqa_b200.model import VQAModel  # noqa: E402
from vqa_b200.synth import synth_batch  # noqa: E402

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
model = VQAModel().eval().cuda()
inf = VQAInference(device="cuda:0")
inf.model, inf._is_loaded = model, True
inf.pipeline_lanes = lanes
inf.pipeline_slots = int(os.environ.get("VQA_PIPE_SLOTS", "0"))
u8, _, ids, mask = synth_batch(256, 1234, full_length=True)
host = (u8.pin_memory(), ids.pin_memory(), mask.pin_memory())
dev = (u8.cuda(), ids.cuda(), mask.cuda())
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
IDLE = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0


def run(src, name):
    with torch.no_grad():
        for _ in inf.predict_tensors_pipelined([src] * (4 * lanes + 4), 5):
            pass
        torch.cuda.synchronize()
        time.sleep(IDLE)
        t0 = time.perf_counter()
        for _ in inf.predict_tensors_pipelined([src] * K, 5):
            pass
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / K
    print(f"{name}: {ms:.4f} ms/step, {256 / ms * 1e3:.0f} pairs/s ({lanes} lanes)")


run(host, "pinned host inputs (H2D every step)")
run(dev, "device inputs (D2D every step)   ")

# the same captured graphs replayed with nothing in between (no copies, no events): the floor of this loop's GPU work
slots = next(iter(inf._pipe_slots.values()))
streams = inf._pipe_streams[1]
torch.cuda.synchronize()
for nuse in (len(slots), lanes):
    for rep in range(2):
        torch.cuda.synchronize()
        time.sleep(IDLE)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        a.record(cur)
        for st in streams:
            st.wait_stream(cur)
        for i in range(K):
            sl = slots[i % nuse]
            with torch.cuda.stream(streams[(i % nuse) % lanes]):
                sl["graph"].replay()
        for st in streams:
            cur.wait_stream(st)
        b.record(cur)
        torch.cuda.synchronize()
    print(f"graphs only ({nuse} of {len(slots)} slots), no copies / events: {a.elapsed_time(b) / K:.4f} ms/step, {256 * K / a.elapsed_time(b) * 1e3:.0f} pairs/s")
