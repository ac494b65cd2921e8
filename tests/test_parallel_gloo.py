"""N>1 path on CPU: world_size-2 gloo processes shard clips with funcodec_b200.parallel.ShardedCodec (scatter ->
per-rank hot path -> gather in the reference layouts).  The per-rank compute is the CPU oracle on the tiny config
(the CUDA path needs a GPU); what is under test is the sharding / placement logic."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_clips, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from funcodec_b200 import get_config, init_state_dict
    from funcodec_b200.parallel import ShardedCodec, shard_bounds
    from oracle.encodec_oracle import OracleEncodec
    cfg = get_config("tiny_ds40")
    sd = init_state_dict(cfg, 3)
    oracle = OracleEncodec(sd, cfg.ratios, cfg.sample_rate, cfg.lstm_layers)

    def run(w):
        r = oracle.inference(w, need_recon=True)
        return r["code_indices"][0], r["recon_speech"]

    L = 40 * 9
    g = torch.Generator().manual_seed(42)
    wav = 0.1 * torch.randn(n_clips, L, generator=g)
    sharded = ShardedCodec(run)
    out = sharded(wav if rank == 0 else None, n_clips, L, torch.device("cpu"))
    if rank == 0:
        codes, recon = out
        ref_codes, ref_recon = run(wav)
        ok = torch.equal(codes, ref_codes) and torch.allclose(recon, ref_recon, atol=1e-6)
        ok = ok and shard_bounds(n_clips, world)[-1][1] == n_clips
        ret.put(bool(ok))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [5, 4, 2, 1])
def test_sharded_codec_world2(n_clips):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() + n_clips) % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True


def test_shard_bounds():
    from funcodec_b200.parallel import shard_bounds
    assert shard_bounds(512, 8) == [(64 * i, 64 * i + 64) for i in range(8)]
    assert shard_bounds(5, 2) == [(0, 3), (3, 5)]
    assert shard_bounds(1, 2) == [(0, 1), (1, 1)]


def _shared_host_worker(rank, world, port, n_clips, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from funcodec_b200 import get_config, init_state_dict
    from funcodec_b200.parallel import SharedHostBatch
    from oracle.encodec_oracle import OracleEncodec
    cfg = get_config("tiny_ds40")
    oracle = OracleEncodec(init_state_dict(cfg, 3), cfg.ratios, cfg.sample_rate, cfg.lstm_layers)
    L = 40 * 9
    tf = cfg.frames(L)
    shb = SharedHostBatch(f"fcb_test_{port}", n_clips, L, cfg.num_quantizers, tf, rank, world, create=(rank == 0))
    dist.barrier()
    shb.map()
    if rank == 0:
        g = torch.Generator().manual_seed(42)
        shb.wav.copy_(0.1 * torch.randn(n_clips, L, generator=g))
    dist.barrier()
    lo, hi = shb.shard()
    r = oracle.inference(shb.wav[lo:hi].clone(), need_recon=True)
    shb.codes[rank].copy_(r["code_indices"][0])
    shb.recon[lo:hi].copy_(r["recon_speech"])
    dist.barrier()
    if rank == 0:
        ref = oracle.inference(shb.wav.clone(), need_recon=True)
        codes = shb.codes.permute(1, 0, 2, 3).reshape(cfg.num_quantizers, n_clips, tf)
        ret.put(bool(torch.equal(codes, ref["code_indices"][0]) and torch.allclose(shb.recon, ref["recon_speech"], atol=1e-6)))
    dist.barrier()
    shb.close()
    dist.destroy_process_group()


def test_shared_host_batch_world2():
    """SharedHostBatch: one /dev/shm batch mapped by both ranks, each rank reads its shard and writes its results in place
    (the cudaHostRegister step only happens when CUDA is available)."""
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 30500 + os.getpid() % 1000
    procs = [ctx.Process(target=_shared_host_worker, args=(r, 2, port, 4, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True
