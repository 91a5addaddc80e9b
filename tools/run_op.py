"""Run selected ops of the headline plan alone (after one whole forward filled the buffers) -- the target for
`ncu --set full -k regex:<kernel>`:   python tools/run_op.py s1.tail [reps] [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_b200 import program as P  # noqa: E402
from vqa_b200.model import VQAModel  # noqa: E402
from vqa_b200.runtime import Plan  # noqa: E402
from vqa_b200.synth import synth_batch  # noqa: E402

pat = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
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
plan.run(ext, st)
torch.cuda.synchronize()
sel = [k for k, op in enumerate(prog.ops) if pat in op.name]
torch.cuda.cudart().cudaProfilerStart()
for _ in range(reps):
    for k in sel:
        plan.run(ext, st, k, k + 1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ran", [prog.ops[k].name for k in sel], "x", reps)
