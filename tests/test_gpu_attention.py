"""Attention core (VQA_OP_SELF_ATTN / VQA_OP_CROSS_ATTN, attn_mma_kernel) through the C ABI against the CPU
emulator: every key-tile template (3, 4, 7, 8 tiles of 8 keys), more than 32 queries (two query blocks),
key masks incl. a fully masked row (NaN like the reference) and the aux attention weights."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from vqa_b200 import program as P  # noqa: E402
import gpu_util as G  # noqa: E402


def _empty_weights(device):
    W = P.Weights(device)
    W.add("dummy", torch.zeros(8), torch.float32)
    return W.finalize()


@pytest.mark.parametrize("B,L", [(3, 20), (2, 24), (2, 32), (3, 40), (2, 64), (5, 7)])
def test_self_attention_with_key_mask(B, L):
    H, hd = 8, 32
    D = H * hd

    def build(device):
        ol = P.OpList(_empty_weights(device), device)
        qkv = ol._buf("qkv", torch.float32, B * L, 3 * D)
        mask = ol._buf("mask", torch.int32, B, L)
        out = ol._buf("out", torch.float32, B * L, D)
        ol._op("self_attn", "attn", dict(B=B, L=L, H=H, hd=hd, ld_qkv=3 * D), dict(qkv=qkv, mask=mask, out=out))
        ol.commit()
        g = torch.Generator().manual_seed(L)
        G.named(ol, "qkv").copy_(torch.randn(B * L, 3 * D, generator=g))
        m = torch.zeros(B, L, dtype=torch.int32)
        for b in range(B):
            m[b, : max(1, (b + 1) * L // B)] = 1
        if B >= 3:
            m[1].zero_()                      # a fully masked sequence: every row of it is NaN in the reference
        G.named(ol, "mask").copy_(m)
        return ol
    cpu, gpu, _, _ = G.run_pair(build)
    want, got = G.named(cpu, "out"), G.named(gpu, "out").cpu()
    nan_w, nan_g = torch.isnan(want), torch.isnan(got)
    assert torch.equal(nan_w, nan_g), "NaN rows (fully masked sequences) differ"
    G.report(f"self-attn B{B} L{L}", torch.nan_to_num(got), torch.nan_to_num(want), atol=2e-3, rtol=2e-3)


@pytest.mark.parametrize("B,L,T", [(3, 20, 49), (2, 40, 49), (2, 5, 64), (2, 20, 33), (2, 12, 20)])
def test_cross_attention_and_weights(B, L, T):
    H, hd = 8, 32
    D = H * hd

    def build(device):
        ol = P.OpList(_empty_weights(device), device)
        q = ol._buf("q", torch.float32, B * L, D)
        kv = ol._buf("kv", torch.float32, B * T, 2 * D)
        out = ol._buf("out", torch.float32, B * L, D)
        w = ol._buf("w", torch.float32, B, H, L, T)
        ol._op("cross_attn", "attn", dict(B=B, L=L, H=H, hd=hd, T=T, ld_q=D, ld_kv=2 * D, k_off=0, v_off=D),
               dict(q=q, kv=kv, out=out, weights=w))
        ol.commit()
        g = torch.Generator().manual_seed(T)
        G.named(ol, "q").copy_(torch.randn(B * L, D, generator=g))
        G.named(ol, "kv").copy_(torch.randn(B * T, 2 * D, generator=g))
        return ol
    cpu, gpu, _, _ = G.run_pair(build)
    G.report(f"cross-attn B{B} L{L} T{T}", G.named(gpu, "out"), G.named(cpu, "out"), atol=2e-3, rtol=2e-3)
    G.report(f"cross-attn weights B{B} L{L} T{T}", G.named(gpu, "w"), G.named(cpu, "w"), atol=1e-5, rtol=1e-4)
    rows = G.named(gpu, "w").cpu().sum(dim=-1)
    assert torch.allclose(rows, torch.ones_like(rows), atol=1e-5)
