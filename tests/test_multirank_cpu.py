"""Multi-rank host logic on CPU (world size 2, gloo): the batch shards by pairs with no data-path
collective; the only collective is the load-time broadcast of the packed weight arena."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO


def _worker(rank, world, port, out):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emulator as E
    from oracle import vqa_oracle as O
    from vqa_b200 import program as P
    from vqa_b200.model import VQAModel
    from vqa_b200.synth import synth_batch
    torch.manual_seed(0 if rank == 0 else 12345)          # rank 1 starts with DIFFERENT weights on purpose
    model = VQAModel().eval()
    W = P.build_weights(model.state_dict(), model.config, "cpu")
    if rank != 0:
        W.arena.tensor.zero_()
    dist.broadcast(W.arena.tensor, src=0)                  # the one collective of the path (NCCL on GPUs)
    # every rank now holds rank 0's packed weights; each processes its own shard of the global batch
    B = 2
    _, img, ids, mask = synth_batch(B * world, 777)
    sl = slice(rank * B, (rank + 1) * B)
    prog = P.Program(W, model.config, B, 20, "nchw_f32", P.MASK_I64, device="cpu")
    logits, _ = E.run_program(prog, img[sl], ids[sl], mask[sl])
    torch.manual_seed(0)
    ref_sd = VQAModel().eval().state_dict()
    want, _ = O.vqa_forward(ref_sd, img[sl], ids[sl], mask[sl])
    err = float((logits - want).abs().max() / want.abs().max())
    gathered = [torch.zeros_like(logits) for _ in range(world)]
    dist.all_gather(gathered, logits)                      # optional final gather of results
    if rank == 0:
        torch.save({"err": err, "shape": tuple(torch.cat(gathered).shape)}, out)
    assert err < 2e-2, err
    dist.destroy_process_group()


def test_weight_broadcast_and_batch_sharding_world2(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["shape"] == (4, 1000) and res["err"] < 2e-2
