"""Per-role clock64() timeline of CTA 0 for every GEMM launch of one forward (profiling aid)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_b200 import program as P  # noqa: E402
from vqa_b200.model import VQAModel  # noqa: E402
from vqa_b200.runtime import Plan  # noqa: E402
from vqa_b200.synth import synth_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
model = VQAModel().eval().cuda()
W = P.build_weights(model.state_dict(), model.config, "cuda")
prog = P.Program.__new__(P.Program)
P.OpList.__init__(prog, W, "cuda", True)
prog.cfg, prog.B, prog.L, prog.in_fmt, prog.mask_dtype, prog.want_aux, prog.top_k = model.config, B, 20, "nchw_f32", P.MASK_I64, False, 0
prog.Bi, prog.side = B, "both"
prog._build()
gemms = [k for k, op in enumerate(prog.ops) if op.kind in ("gemm", "stem_pool")]
for k in gemms:
    prog.ops[k].p["dbg"] = prog._buf(f"dbg{k}", torch.int64, 32)
prog.commit()
plan = Plan(prog.ops, 0)
_, img, ids, mask = synth_batch(B, 1234, full_length=True)
img, ids, mask = img.cuda(), ids.cuda(), mask.cuda()
logits = torch.empty(B, 1000, device="cuda")
ext = [img.data_ptr(), ids.data_ptr(), mask.data_ptr(), logits.data_ptr(), 0, 0]
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    plan.run(ext, st)
torch.cuda.synchronize()
if os.environ.get("VQA_TIMELINE_ISOLATED"):   # every GEMM launched alone: no PDL overlap with its neighbours in the numbers
    for k in gemms:
        plan.run(ext, st, k, k + 1)
        torch.cuda.synchronize()
print("op name: setup | first_tma | first_a_full | tile0_mma_issued | tile0_acc_full_seen | tile0_epi_done | last_mma_issued | last_epi_done | exit   (cycles from kernel entry)")
only = os.environ.get("VQA_TIMELINE_ONLY", "")
for k in gemms:
    if only and not prog.ops[k].name.startswith(only):
        continue
    t = prog.tensor(f"dbg{k}").cpu().tolist()
    d = [x - t[0] for x in t]
    oi = prog.ops[k].i
    if prog.ops[k].kind == "stem_pool":
        tiles = oi["B"] * (oi["Ho"] + oi["Ho"] // oi["run_len"] - 1)
    else:
        tiles = oi["tiles_per_img"] * oi["n_imgs"] if oi.get("tiles_per_img") else (oi["M"] + 128 * oi["MT"] - 1) // (128 * oi["MT"])
    print(f"{k:3d} {prog.ops[k].name:14s} setup {d[1]:5d} tma0 {d[2]:5d} a_full0 {d[3]:6d} mma0_issued {d[4]:6d} acc_seen0 {d[5]:6d} "
          f"epi0_done {d[6]:6d} last_mma {d[9]:7d} last_epi {d[7]:7d} exit {d[8]:7d} | waits: prod a_empty {t[16]:7d} b_empty {t[17]:7d} "
          f"| mma acc_empty {t[18]:7d} a_full {t[19]:7d} b_full {t[20]:7d} | epi(w2) acc_full {t[21]:7d} | m_tiles {tiles} "
          f"| wall {(t[23] - t[22]) / 1e3:7.1f} us sm_clk {d[8] / max(t[23] - t[22], 1) * 1e3:6.0f} MHz "
          f"| epi(w2) phases: tmem_ld {t[10]} xchg {t[11]} combine {t[12]} finish {t[14]} store_wait {t[15]}")
