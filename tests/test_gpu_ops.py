"""Every op of the real forward program, one at a time, against the CPU emulator.

Before each op the GPU workspace is overwritten with the emulator's state, so each kernel sees
bit-identical inputs and an error is attributed to exactly one launch (op index + name are in
the assertion message).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

import emulator as E  # noqa: E402
from vqa_b200 import program as P  # noqa: E402
from vqa_b200.model import VQAModel  # noqa: E402
from vqa_b200.runtime import Plan  # noqa: E402
from vqa_b200.synth import randomise_state, synth_batch  # noqa: E402

OUTPUTS = {
    "ingest": [("dst", torch.bfloat16)], "gemm": [("out", None), ("sums", torch.float32)], "stem_pool": [("out", torch.bfloat16)], "mlp_chain": [("xout", torch.float32), ("y", torch.float32)], "maxpool": [("dst", torch.bfloat16)],
    "se_squeeze": [("sums", torch.float32)], "se_excite": [("scale", torch.float32)],
    "spatial_map": [("att", torch.float32)], "scale_relayout": [("dst", torch.bfloat16)],
    "embed": [("dst", torch.float32), ("ln_dst", torch.float32), ("mask_dst", torch.int32)],
    "layernorm": [("dst", torch.float32), ("dst2", torch.float32), ("dst3", torch.float32)],
    "self_attn": [("out", torch.float32)], "cross_attn": [("out", torch.float32), ("weights", torch.float32)],
    "pool_gate_ln": [("fused", torch.float32), ("att_pooled", torch.float32), ("txt_pooled", torch.float32),
                     ("cat", torch.float32)],
    "softmax_topk": [("idx", torch.int64), ("probs", torch.float32)], "mask_prep": [("dst", torch.int32)],
    "grid_to_nchw": [("dst", torch.float32)], "copy_rows": [("dst", torch.float32)], "split_tf32": [("dst", torch.float32)],
    "stage_tail": [("dst", torch.bfloat16), ("scale", torch.float32), ("att", torch.float32)],
}


def _view(ref, dtype, ext):
    if isinstance(ref, P.ExtRef):
        return ext[ref.slot].reshape(-1)
    return ref.arena.tensor[ref.offset: ref.offset + ref.nbytes].view(dtype)


def _case(ctor, B, L, in_fmt, mask_kind, window=True, seed=0, precision="bf16"):
    torch.manual_seed(seed)
    model = VQAModel(**ctor).eval()
    sd = randomise_state(model.state_dict(), 1)
    u8, img, ids, mask = synth_batch(B, 1234, max_len=L, vocab=model.config["vocab_size"])
    images = u8 if in_fmt == "hwc_u8" else img
    code = {"i64": P.MASK_I64, "f32": P.MASK_F32, "none": P.MASK_NONE}[mask_kind]
    m = {"i64": mask, "f32": mask.float(), "none": None}[mask_kind]
    progs = {}
    for dev in ("cpu", "cuda"):
        W = P.build_weights(sd, model.config, dev, precision=precision)
        progs[dev] = P.Program(W, model.config, B, L, in_fmt, code, want_aux=True, top_k=5, device=dev, window=window)
    NA = model.config["num_answers"]
    ext = [images, ids, m, torch.zeros(B, NA), torch.zeros(B, 5, dtype=torch.int64), torch.zeros(B, 5)]
    return progs["cpu"], progs["cuda"], ext


@pytest.mark.parametrize("ctor,B,L,in_fmt,mask_kind,window,precision", [
    ({}, 2, 20, "nchw_f32", "i64", True, "bf16"),
    ({}, 2, 20, "hwc_u8", "i64", True, "tf32"),
    ({}, 3, 20, "hwc_u8", "f32", False, "bf16"),
    (dict(use_se_attention=False, use_spatial_attention=False, use_gating=False, num_transformer_layers=1,
          num_cross_layers=1, max_question_length=12, vocab_size=500, num_answers=37), 2, 12, "nchw_f32", "none", True, "bf16"),
    (dict(use_se_attention=False, use_spatial_attention=False, use_gating=False, num_transformer_layers=1,
          num_cross_layers=1, max_question_length=12, vocab_size=500, num_answers=37), 2, 12, "nchw_f32", "none", True, "tf32"),
    (dict(use_spatial_attention=False, max_question_length=64, num_cross_layers=1, num_transformer_layers=1),
     1, 64, "nchw_f32", "i64", True, "bf16"),
])
def test_each_op_against_emulator(ctor, B, L, in_fmt, mask_kind, window, precision):
    cpu, gpu, ext = _case(ctor, B, L, in_fmt, mask_kind, window, precision=precision)
    ext_gpu = [None if t is None else t.cuda() for t in ext]
    plan = Plan(gpu.ops, 0)
    emu = E.Emulator(cpu)
    stream = torch.cuda.current_stream().cuda_stream
    failures = []
    for k, op in enumerate(cpu.ops):
        gpu.ws.tensor.copy_(cpu.ws.tensor)
        for a, b in zip(ext_gpu, ext):
            if a is not None:
                a.copy_(b)
        emu.run(ext, k, k + 1)
        plan.run([0 if t is None else t.data_ptr() for t in ext_gpu], stream, k, k + 1)
        torch.cuda.synchronize()
        gop = gpu.ops[k]
        for field, dtype in OUTPUTS[op.kind]:
            ref_c, ref_g = op.p.get(field), gop.p.get(field)
            if ref_c is None:
                continue
            if dtype is None:
                dtype = {P.OUT_BF16: torch.bfloat16, P.OUT_F32: torch.float32, P.OUT_F16: torch.float16}[op.i["out_dtype"]]
            if dtype == torch.bfloat16 and op.i.get("f32"):
                dtype = torch.float32          # tf32 precision mode: fp32 activation grids
            # operands of the fp16 tail Linears are stored as fp16 by their producers
            if (op.kind in ("self_attn", "cross_attn") and field == "out" and op.i.get("no_round") == 2) or \
               (op.kind == "layernorm" and op.i.get({"dst": "round_tf32", "dst2": "rnd2", "dst3": "rnd3"}[field]) == 2) or \
               (op.kind == "embed" and field == "ln_dst" and op.i.get("round_tf32") == 2) or \
               (op.kind == "pool_gate_ln" and field == "cat" and op.i.get("no_round") == 2):
                dtype = torch.float16
            want = _view(ref_c, dtype, ext)
            got = _view(ref_g, dtype, ext_gpu).cpu()
            if op.kind == "split_tf32":        # hi may differ by one tf32 ulp on rounding ties; hi + lo may not
                M, K = op.i["M"], op.i["K"]
                want = want[: M * 2 * K].view(M, 2, K).sum(dim=1)
                got = got[: M * 2 * K].view(M, 2, K).sum(dim=1)
            if dtype in (torch.int32, torch.int64):
                ok = torch.equal(got, want)
                msg = f"op {k} {op.name} ({plan.kernel_name(k)}) {field}: integer mismatch"
            else:
                g32, w32 = got.float(), want.float()
                atol, rtol = (1e-2, 1.6e-2) if dtype == torch.bfloat16 else (2e-3, 2e-3)   # fp16: one ulp = 1e-3 relative
                if op.kind in ("se_squeeze",):
                    atol = 1e-2
                err = (g32 - w32).abs()
                finite = torch.isfinite(g32).all()
                ok = bool(finite) and bool((err <= atol + rtol * w32.abs()).all())
                msg = (f"op {k} {op.name} ({plan.kernel_name(k)}) {field}: max_abs_err={err.max().item():.3e} "
                       f"ref_absmax={w32.abs().max().item():.3e} bad={(err > atol + rtol * w32.abs()).sum().item()}"
                       f"/{err.numel()} finite={bool(finite)}")
            if not ok:
                failures.append(msg)
                print("FAIL", msg)
    assert not failures, "\n".join(failures[:20])


@pytest.mark.parametrize("B,H,C,CS", [(40, 56, 64, 4), (77, 28, 128, 2), (150, 56, 64, 4), (3, 12, 256, 1)])
def test_stage_tail_se_only_many_images(B, H, C, CS):
    """SE-only stage tail (two-pass kernel: sums streamed through registers, rows re-read from L2) with more clusters than
    fit the GPU at once, against the emulator."""
    import gpu_util as G
    g, gn = P.Grid(B, H, H), P.Grid(B, H // 2, H // 2)
    R = C // 16

    def build(device):
        gen = torch.Generator().manual_seed(5)
        W = P.Weights(device)
        W.add("w1", torch.randn(R, C, generator=gen) * 0.2, torch.float32)
        W.add("w2", torch.randn(R, C, generator=gen) * 0.5, torch.float32)
        W.finalize()
        ol = P.OpList(W, device)
        x = ol._buf("x", torch.bfloat16, g.rows, C)
        dst = ol._buf("dst", torch.bfloat16, 4 * gn.rows, C)
        sc = ol._buf("scale", torch.float32, B, C)
        ol._op("stage_tail", "tail",
               dict(B=B, C=C, H=H, W=H, P=g.P, RPI=g.rpi, R=R, ks=0, mode=1, Po=gn.P, RPIo=gn.rpi, phase_rows=gn.rows,
                    CS=CS, f32=0, split=0),
               dict(src=x, w1=W.buf("w1"), w2=W.buf("w2"), wconv=None, dst=dst, scale=sc, att=None, sums=None))
        ol.commit()
        xv = torch.randn(g.rows, C, generator=gen).clamp_min(0)
        r = torch.arange(g.rows) % g.rpi
        xv[~(((r // g.P) < g.H) & ((r % g.P) < g.W))] = 0
        G.named(ol, "x").copy_(xv.to(torch.bfloat16))
        G.named(ol, "dst").fill_(3.0)           # pads must be rewritten with zeros
        return ol

    cpu, gpu, _, _ = G.run_pair(build)
    G.report("SE scale", G.named(gpu, "scale"), G.named(cpu, "scale"), atol=1e-5, rtol=1e-4)
    G.report("tail dst", G.named(gpu, "dst"), G.named(cpu, "dst"), atol=1e-2, rtol=1.6e-2)
