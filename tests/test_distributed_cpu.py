"""CPU, world_size 2, gloo: the data-parallel plumbing (flat gradient bucket + one all-reduce, frame-balanced
sharding).  Gradients here come from the CPU oracle -- the kernels themselves are covered by the -m gpu tests."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from asr_b200.distributed import FlatGradBucket, frame_balanced_shards


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import torch_path
        from oracle.make_golden import synth_batch

        torch.set_num_threads(2)
        p = torch_path.init_params("gru", 16, 1, 29)
        params = {k: torch.nn.Parameter(v.clone()) for k, v in p.items() if k in torch_path.trainable(p)}
        bucket = FlatGradBucket(params.values())
        full = synth_batch(5, 4, 41, [3, 3, 2, 2], 29, [41, 41, 30, 30])
        shards = frame_balanced_shards([41, 41, 30, 30], world)
        idx = shards[rank]
        tsz = full[3][idx]
        offs = [0] + torch.cumsum(full[3], 0).tolist()
        tgt = torch.cat([full[1][offs[i]:offs[i + 1]] for i in idx])
        local = (full[0][idx], tgt, full[2][idx], tsz)
        bucket.zero()
        q = {**p, **params}
        loss, _ = torch_path.fit_loss(q, *local, rnn_type="gru")
        loss.backward()
        local_flat = bucket.flat.clone()
        bucket.all_reduce_mean()
        gathered = [torch.zeros_like(local_flat) for _ in range(world)]
        dist.all_gather(gathered, local_flat)
        want = sum(gathered) / world
        ret[rank] = (torch.allclose(bucket.flat, want, atol=1e-7), bool(bucket.flat.abs().sum() > 0),
                     all(prm.grad.data_ptr() == v.data_ptr() for prm, v in zip(bucket.params, bucket.views)))
    finally:
        dist.destroy_process_group()


def test_flat_bucket_allreduce_two_ranks():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 1000, ret), nprocs=2, join=True)
    assert dict(ret) == {0: (True, True, True), 1: (True, True, True)}


def test_frame_balanced_shards_beats_whole_bin_dealing():
    g = torch.Generator().manual_seed(0)
    frames = sorted((torch.randint(500, 2001, (512,), generator=g)).tolist(), reverse=True)
    shards = frame_balanced_shards(frames, 8)
    assert sorted(i for s in shards for i in s) == list(range(512)) and all(len(s) == 64 for s in shards)
    loads = [sum(frames[i] for i in s) for s in shards]
    assert max(loads) / min(loads) < 1.02
    whole_bins = [sum(frames[r * 64:(r + 1) * 64]) for r in range(8)]       # the reference's dealing rule
    assert max(whole_bins) / min(whole_bins) > 2.0
    for s in shards:
        ls = [frames[i] for i in s]
        assert ls == sorted(ls, reverse=True)


def _worker_split(rank, world, port, ret):
    """OverlappedGradSync: the bucket reduced in two pieces (head = the conv stack's gradients, tail = the rest) equals the
    one-piece mean.  On the CPU there is no communication stream, so the hook is a no-op and finish() does both pieces."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace

        from asr_b200.distributed import OverlappedGradSync

        g = torch.Generator().manual_seed(10 + rank)
        conv = torch.nn.Conv2d(1, 2, 3)
        rest = torch.nn.Linear(4, 3)
        model = SimpleNamespace(conv=conv)
        params = list(conv.parameters()) + list(rest.parameters())
        bucket = FlatGradBucket(params)
        sync = OverlappedGradSync(bucket, model)
        assert sync.head == sum(p.numel() for p in conv.parameters()) and callable(model.after_rnn_backward)
        bucket.zero()
        for p in params:
            p.grad.copy_(torch.randn(p.shape, generator=g))
        local = bucket.flat.clone()
        model.after_rnn_backward()          # CPU: nothing to overlap
        sync.finish()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        ret[rank] = bool(torch.allclose(bucket.flat, sum(gathered) / world, atol=1e-7))
        # second form: the tail already reduced by the hook (emulated), finish() completes the head and waits
        bucket.zero()
        for p in params:
            p.grad.copy_(torch.randn(p.shape, generator=g))
        local = bucket.flat.clone()
        sync.work = dist.all_reduce(bucket.flat[sync.head:], op=dist.ReduceOp.SUM, async_op=True)
        sync.work.wait()
        bucket.flat[sync.head:].div_(world)
        sync.finish()
        dist.all_gather(gathered, local)
        ret[rank] = ret[rank] and bool(torch.allclose(bucket.flat, sum(gathered) / world, atol=1e-7))
    finally:
        dist.destroy_process_group()


def test_overlapped_grad_sync_two_ranks():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_split, args=(2, 30500 + os.getpid() % 1000, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}
