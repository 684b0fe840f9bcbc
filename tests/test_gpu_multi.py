"""2-GPU test (-m gpu; skipped with < 2 devices): sequence-sharded forward over NCCL == single-GPU forward."""
import os
import socket

import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tag, q):
    import torch.distributed as dist
    import caduceus
    from caduceus_b200 import seqshard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        fx = golden(f"model_{tag}.pt")
        cfg = caduceus.CaduceusConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in fx["config"].items()})
        model = caduceus.CaduceusForMaskedLM(cfg)
        model.load_state_dict(fx["state_dict"])
        model = model.to(dev).eval()
        g = torch.Generator().manual_seed(0)
        L = 1024 * world
        ids = torch.randint(7, 11, (2, L), generator=g)
        Ls = L // world
        with torch.no_grad():
            with seqshard.sequence_parallel():
                local = model(ids[:, rank * Ls:(rank + 1) * Ls].to(dev)).logits
            parts = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(parts, local)
            if rank == 0:
                full = model(ids.to(dev)).logits
                q.put((torch.cat(parts, dim=1).cpu(), full.cpu()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("tag", ["ps_small", "ph_small"])
def test_sequence_sharded_forward_matches_single_gpu(tag):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, tag, q)) for r in range(world)]
    for p in procs:
        p.start()
    sharded, full = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    err = (sharded - full).abs()
    assert torch.all(err <= 2e-3 + 6e-4 * full.abs()), err.max()
