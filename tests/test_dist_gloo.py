"""world_size-2 gloo test of the multi-GPU host logic (sharding, checkpoint broadcast, gather) on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sin3dm_b200.dist import all_reduce_gradients, broadcast_parameters, sample_sharded, shard_range


def test_shard_range_partitions_exactly():
    for n in (1, 7, 8, 64, 65):
        for w in (1, 2, 4, 8):
            got = [shard_range(n, w, r) for r in range(w)]
            assert sum(c for _, c in got) == n
            pos = 0
            for s, c in got:
                assert s == pos
                pos += c


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                      # different weights on every rank before the broadcast
        m = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.GroupNorm(1, 7))
        n = broadcast_parameters(m, src=0)
        torch.manual_seed(100)
        ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.GroupNorm(1, 7))
        same = all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), ref.state_dict().values()))

        # a "sampler" whose output encodes the global sample index: checks sample_base bookkeeping + gather order
        def fake(shape, base):
            return torch.stack([torch.full(shape[1:], float(base + i)) for i in range(shape[0])])

        out = sample_sharded(fake, 5, (2, 3), batch_size=2, gather=True)
        ok = out.shape == (5, 2, 3) and all(float(out[i, 0, 0]) == i for i in range(5))
        # data-parallel training: one all-reduce of the flat gradient buffer, averaged over the ranks
        flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
        all_reduce_gradients(flat)
        ok = ok and torch.allclose(flat, torch.arange(10, dtype=torch.float32) * (sum(range(1, world + 1)) / world))
        q.put((rank, n, same, ok))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, n, same, ok in res:
        assert n == 5 * 7 + 7 + 7 + 7 and same and ok, (rank, n, same, ok)


def _sampler_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sin3dm_b200.resample import LossSecondMomentResampler

        class D:
            num_timesteps = 4
        r = LossSecondMomentResampler(D(), history_per_term=2, uniform_prob=0.0)
        # ragged per-rank batches: rank 0 reports steps {0,1,2}, rank 1 reports {3}; twice -> every history is full
        for rep in range(2):
            ts = torch.tensor([0, 1, 2]) if rank == 0 else torch.tensor([3])
            ls = torch.tensor([1.0, 2.0, 3.0]) * (rep + 1) if rank == 0 else torch.tensor([4.0]) * (rep + 1)
            r.update_with_local_losses(ts, ls)
        q.put((rank, r.weights().tolist()))
    finally:
        dist.destroy_process_group()


def test_loss_aware_sampler_synchronises_ranks_world2():
    """LossAwareSampler.update_with_local_losses (resample.py:70-103 in the reference): every rank ends with the same weights."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sampler_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == res[1]
    rms = [(1 * 1 + 2 * 2) ** 0.5 * k for k in (1.0, 2.0, 3.0, 4.0)]
    want = [v / sum(rms) for v in rms]
    assert all(abs(a - b) < 1e-12 for a, b in zip(res[0], want))
