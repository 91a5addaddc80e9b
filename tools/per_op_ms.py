"""Per-op CUDA-event timings of one forward at the headline shape: every op of the plan is launched alone
(vqa_plan_run_range k..k+1), `reps` times back to back after a warm-up of the whole plan; the median is reported.
Launched alone an op neither overlaps its neighbours (PDL) nor inherits their L2 contents, so the sum exceeds the
live step; use it for the ranking and for before/after comparisons of one kernel.

    python tools/per_op_ms.py [batch] > profiles/rNN_x_per_op_ms.json
"""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_b200 import program as P  # noqa: E402
from vqa_b200.model import VQAModel  # noqa: E402
from vqa_b200.runtime import Plan  # noqa: E402
from vqa_b200.synth import synth_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = 7
torch.manual_seed(0)
model = VQAModel().eval().cuda()
W = P.build_weights(model.state_dict(), model.config, "cuda")
prog = P.Program(W, model.config, B, 20, "nchw_f32", P.MASK_I64, want_aux=False, top_k=5, device="cuda")
plan = Plan(prog.ops, 0)
_, img, ids, mask = synth_batch(B, 1234, full_length=True)
img, ids, mask = img.cuda(), ids.cuda(), mask.cuda()
logits = torch.empty(B, 1000, device="cuda")
idx = torch.empty(B, 5, dtype=torch.int64, device="cuda")
probs = torch.empty(B, 5, device="cuda")
ext = [img.data_ptr(), ids.data_ptr(), mask.data_ptr(), logits.data_ptr(), idx.data_ptr(), probs.data_ptr()]
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    plan.run(ext, st)
torch.cuda.synchronize()
ops = []
for k, op in enumerate(prog.ops):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run(ext, st, k, k + 1)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ops.append({"op": k, "name": op.name, "kernel": plan.kernel_name(k), "ms": statistics.median(ts)})
groups = {}
for o in ops:
    n = o["name"]
    key = ("stem" if n.startswith("stem") else n[:2] + ".convs" if n[:2] in ("s1", "s2", "s3", "s4") and "conv" in n
           else "stage tails" if n.endswith(".tail") else "ingest" if n == "ingest" else "text/fusion/head")
    groups[key] = groups.get(key, 0.0) + o["ms"]
print(json.dumps({"batch": B, "sum_ms": sum(o["ms"] for o in ops), "groups_ms": groups, "ops": ops}, indent=1))
