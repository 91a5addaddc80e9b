"""Helpers for the -m gpu tests: build the same op list on CPU (emulator) and CUDA (library)."""
from __future__ import annotations

import os

import torch

import emulator as E
from vqa_b200 import program as P
from vqa_b200.runtime import Plan


def run_pair(build, ext_cpu=None, n_ext=6):
    """``build(device) -> OpList`` (deterministic).  Runs it through the CPU emulator and through
    libvqa_b200 on cuda:0; returns (cpu_oplist, gpu_oplist, ext_cpu, ext_gpu)."""
    cpu = build("cpu")
    if os.environ.get("VQA_DRY"):  # CPU-only rehearsal of the test's Python side (no library call)
        ext_cpu = list(ext_cpu) if ext_cpu is not None else [None] * n_ext
        E.Emulator(cpu).run(ext_cpu)
        return cpu, cpu, ext_cpu, ext_cpu
    gpu = build("cuda")
    assert len(cpu.ops) == len(gpu.ops)
    # identical workspace contents (inputs are written into named buffers by build())
    gpu.ws.tensor.copy_(cpu.ws.tensor)
    ext_cpu = list(ext_cpu) if ext_cpu is not None else [None] * n_ext
    ext_gpu = [None if t is None else t.cuda() for t in ext_cpu]
    E.Emulator(cpu).run(ext_cpu)
    plan = Plan(gpu.ops, 0)
    ptrs = [0 if t is None else t.data_ptr() for t in ext_gpu]
    plan.run(ptrs, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return cpu, gpu, ext_cpu, ext_gpu


def named(oplist, name):
    b, dtype, shape = oplist.named[name]
    return b.view(dtype, *shape)


def alloc(ol, name, dtype, *shape):
    return ol._buf(name, dtype, *shape)


def report(tag, got, want, atol, rtol=0.0):
    got, want = got.float().cpu(), want.float().cpu()
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = err > tol
    msg = (f"{tag}: max_abs_err={err.max().item():.3e} ref_absmax={want.abs().max().item():.3e} "
           f"bad={int(bad.sum())}/{bad.numel()}")
    if bad.any():
        idx = bad.nonzero()
        rows = sorted(set(int(r[0]) for r in idx[:2000]))
        msg += f" first_bad={idx[0].tolist()} bad_rows(sample)={rows[:24]}"
    print(msg)
    assert not bad.any(), msg
