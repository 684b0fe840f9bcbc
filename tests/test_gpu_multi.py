"""Sequence-sharded forward (-m gpu): ONE sequence cut over several ranks == the single-GPU forward.

Three data planes, none of which skips on a 1-GPU box except the one that needs real peers:
  * the peer-memory exchange kernels (csrc/peer_exchange.cu) with several VIRTUAL ranks inside this process on one GPU — one
    stream / thread per rank, "peer" workspaces = separate device buffers: kernel level and model level;
  * the collective exchange (seqshard.gather_halo / gather_carry) over NCCL when the box has the GPUs, otherwise over gloo ranks
    that share cuda:0 (host-staged) — the fallback tests/test_gpu_shard_train.py uses;
  * real peers: NCCL ranks + PeerExchange.create (torch symmetric memory over NVLink), >= 2 GPUs only."""
import os
import socket
import threading

import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(tag, dev):
    import caduceus
    fx = golden(f"model_{tag}.pt")
    cfg = caduceus.CaduceusConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in fx["config"].items()})
    model = caduceus.CaduceusForMaskedLM(cfg)
    model.load_state_dict(fx["state_dict"])
    return model.to(dev).eval(), cfg


def _close(a, b, what):
    err = (a.float() - b.float()).abs()
    assert torch.all(err <= 2e-3 + 6e-4 * b.float().abs()), (what, err.max().item())


# ---- virtual ranks on one GPU: kernel level ---------------------------------------------------------------------------------
def _virtual_peers(world, nseq, njobs, E, N):
    from caduceus_b200 import functional as CF, seqshard
    nbytes = CF.peer_ws_bytes(world, nseq, njobs, E, N)
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=DEV) for _ in range(world)]
    return [seqshard.PeerExchange.from_buffers(r, bufs, nseq_max=nseq, njobs_max=njobs, E=E, N=N) for r in range(world)]


def _subprocess_case(*argv):
    """Virtual ranks run in a process of their own: kernels of one rank WAIT for kernels of another, so anything that orders the
    device between them ends in the kernels' 10 s watchdog trap — a lazily loaded module (CUDA_MODULE_LOADING=EAGER here, and only
    here: eager loading of everything a test process imports, vLLM included, takes minutes), a cuBLAS handle created late, a
    cudaMalloc (an implicit synchronisation point between streams) — and a trapped context would take every later test of the
    process with it.  Production has one rank per process and none of these couplings."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.abspath(__file__), *[str(a) for a in argv]], capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, CUDA_MODULE_LOADING="EAGER", CUDA_DEVICE_MAX_CONNECTIONS="32"))
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("dtype", ["bfloat16", "float32"])
def test_peer_exchange_kernels_virtual_ranks(world, dtype):
    """halo + boundary-state exchange of csrc/peer_exchange.cu against the collective formulation of seqshard.py, three times in a
    row (both parities of the double-buffered workspace, epoch counters advancing)."""
    _subprocess_case("--virtual-kernels", world, dtype)


def _virtual_kernel_case(world, dtype):
    from caduceus_b200 import functional as CF, seqshard
    dtype = getattr(torch, dtype)
    E, N, Ls, B, nstrand = 64, 16, 40, 2, 2
    jobs = CF.job_tables(B, nstrand, 2, False, torch.device(DEV))
    nseq, njobs = B * nstrand, jobs[0].numel()
    peers = _virtual_peers(world, nseq, njobs, E, N)
    streams = [torch.cuda.Stream() for _ in range(world)]
    g = torch.Generator().manual_seed(world)
    A2 = (-torch.rand(2, E, N, generator=g) * 3 - 0.1).to(DEV)
    seq_l, _, rev_l = CF.JOB_HOST[jobs[0].data_ptr()]
    for it in range(3):
        xz = torch.randn(nseq, 2 * E, world * Ls, generator=g).to(DEV).to(dtype)
        hl = torch.randn(world, njobs, E, N, generator=g).to(DEV)
        ds = (torch.rand(world, njobs, E, generator=g) * 4).to(DEV)
        torch.cuda.synchronize()
        halos, h0s, dts = [None] * world, [None] * world, [None] * world
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                shard = torch.zeros(nseq, 2 * E, 48, device=DEV, dtype=dtype)          # row pitch 48 >= Ls
                shard[..., :Ls] = xz[..., r * Ls:(r + 1) * Ls]
                halos[r] = CF.peer_halo_exchange(peers[r].ctx, shard, Ls, jobs)
                h0s[r], dts[r] = CF.peer_carry_exchange(peers[r].ctx, hl[r].contiguous(), ds[r].contiguous(), A2, jobs,
                                                        want_dtsum_all=True)
        torch.cuda.synchronize()
        A2_job = A2.index_select(0, jobs[1].long())
        for r in range(world):
            want_h0 = seqshard.compose_carry(hl, ds, A2_job, jobs[2], r)
            assert torch.allclose(h0s[r], want_h0, rtol=2e-5, atol=1e-5), (it, r, (h0s[r] - want_h0).abs().max())
            assert torch.equal(dts[r], ds), (it, r)
            x = xz[:, :E]
            for j in range(njobs):
                s = seq_l[j]
                if not rev_l[j]:
                    want = x[s, :, r * Ls - 3:r * Ls] if r > 0 else torch.zeros(E, 3, device=DEV, dtype=dtype)
                else:
                    want = x[s, :, (r + 1) * Ls:(r + 1) * Ls + 3].flip(-1) if r + 1 < world else torch.zeros(E, 3, device=DEV, dtype=dtype)
                assert torch.equal(halos[r][j], want), (it, r, j)


# ---- virtual ranks on one GPU: model level ----------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,world", [("ps_small", 2), ("ph_small", 3)])
def test_sequence_sharded_forward_virtual_ranks(tag, world):
    """The whole model, `world` shards of one batch, one THREAD + stream per virtual rank, halo / boundary states through the
    peer-exchange kernels: concatenated logits == unsharded forward.  Three forwards in a row (epochs, both buffer parities)."""
    _subprocess_case("--virtual-model", tag, world)


def _virtual_model_case(tag, world):
    from caduceus_b200 import seqshard
    model, cfg = _model(tag, DEV)
    E, N = 2 * cfg.d_model, 16
    B, Ls = 2, 384
    nstrand = 2 if cfg.rcps else 1
    peers = _virtual_peers(world, B * nstrand, B * nstrand * 2, E, N)
    streams = [torch.cuda.Stream() for _ in range(world)]
    for st in streams:       # give every rank's stream a cached large block and small-pool blocks: no cudaMalloc while ranks wait
        with torch.cuda.stream(st):
            warm = [torch.empty(256 << 20, dtype=torch.uint8, device=DEV)] + [torch.empty(1 << 18, device=DEV) for _ in range(16)]
            del warm
    torch.cuda.synchronize()
    g = torch.Generator().manual_seed(1)
    for it in range(3):
        ids = torch.randint(7, 11, (B, world * Ls), generator=g).to(DEV)
        with torch.no_grad():
            full = model(ids).logits
        torch.cuda.synchronize()
        outs, errs = [None] * world, []
        ready = threading.Barrier(world)

        def run(r):
            try:
                with torch.cuda.stream(streams[r]), torch.no_grad(), seqshard.sequence_parallel(peer=peers[r]):
                    # a NEW thread may have to create its cuBLAS handle (cublasCreate synchronises the device): do that, and anything
                    # else a first GEMM on this thread / stream sets up, BEFORE any rank launches a kernel that waits for another rank
                    a = torch.ones(64, 64, device=DEV, dtype=torch.bfloat16)
                    (a @ a).sum().item()
                    ready.wait(60)
                    outs[r] = model(ids[:, r * Ls:(r + 1) * Ls].contiguous()).logits
                    torch.cuda.current_stream().synchronize()
            except Exception as ex:      # noqa: BLE001
                errs.append((r, repr(ex)))

        threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(120)
        assert not errs, errs
        assert all(o is not None for o in outs)
        _close(torch.cat(outs, dim=1), full, f"{tag} world {world} pass {it}")


# ---- separate processes -----------------------------------------------------------------------------------------------------
def _worker(rank, world, port, tag, backend, use_peer, q):
    import torch.distributed as dist
    from caduceus_b200 import seqshard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model, cfg = _model(tag, dev)
        g = torch.Generator().manual_seed(0)
        B, Ls = 2, 1000
        L = Ls * world
        ids = torch.randint(7, 11, (B, L), generator=g)
        peer = None
        if use_peer:
            nstrand = 2 if cfg.rcps else 1
            peer = seqshard.PeerExchange.create(nseq_max=B * nstrand, njobs_max=B * nstrand * 2, E=2 * cfg.d_model, N=16, device=dev)
        with torch.no_grad():
            for _ in range(2):
                with seqshard.sequence_parallel(peer=peer):
                    local = model(ids[:, rank * Ls:(rank + 1) * Ls].to(dev)).logits
            payload = local.float().cpu()
            parts = [torch.empty_like(payload) for _ in range(world)] if rank == 0 else None
            if backend == "nccl":
                dev_parts = [torch.empty_like(local) for _ in range(world)]
                dist.all_gather(dev_parts, local)
                parts = [p.float().cpu() for p in dev_parts]
            else:
                dist.gather(payload, parts, dst=0)
            if rank == 0:
                full = model(ids.to(dev)).logits
                q.put((torch.cat(parts, dim=1).numpy(), full.float().cpu().numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run_procs(world, tag, backend, use_peer):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, tag, backend, use_peer, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        sharded, full = q.get(timeout=300)
    finally:
        for p in procs:
            p.join(120)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    _close(torch.from_numpy(sharded), torch.from_numpy(full), f"{tag} {backend} peer={use_peer}")


@pytest.mark.parametrize("tag,world", [("ps_small", 2), ("ph_small", 3)])
def test_sequence_sharded_forward_collective_exchange(tag, world):
    """all_gather data plane: NCCL with one GPU per rank when the box has them, else gloo ranks sharing cuda:0."""
    _run_procs(world, tag, "nccl" if torch.cuda.device_count() >= world else "gloo", False)


@pytest.mark.parametrize("tag", ["ps_small", "ph_small"])
def test_sequence_sharded_forward_real_peers(tag):
    """NVLink peer memory between two processes (torch symmetric memory): needs 2 GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_procs(2, tag, "nccl", True)


if __name__ == "__main__":          # subprocess entry of test_sequence_sharded_forward_virtual_ranks
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    if len(sys.argv) == 4 and sys.argv[1] == "--virtual-model":
        _virtual_model_case(sys.argv[2], int(sys.argv[3]))
        print("ok")
    elif len(sys.argv) == 4 and sys.argv[1] == "--virtual-kernels":
        _virtual_kernel_case(int(sys.argv[2]), sys.argv[3])
        print("ok")
