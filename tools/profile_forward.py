"""One profiled forward at the headline shape (for ncu: --profile-from-start off).

    python tools/profile_forward.py [--batch 256]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_b200.model import VQAModel  # noqa: E402
from vqa_b200.synth import synth_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
args = ap.parse_args()
torch.manual_seed(0)
model = VQAModel().eval().cuda()
_, img, ids, mask = synth_batch(args.batch, 1234, full_length=True)
img, ids, mask = img.cuda(), ids.cuda(), mask.cuda()
with torch.no_grad():
    for _ in range(3):
        model(img, ids, mask)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    logits, _ = model(img, ids, mask)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("ok", float(logits.abs().max()))
