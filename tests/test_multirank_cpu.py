"""N>1 host logic on CPU: two gloo ranks agree on max-over-ranks timing and whole-job unit counts, and
the batch sharding covers every unit exactly once (no data-path collective)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stereo_toolbox_b200.distrib import shard_range, reduce_stats, FlatGradAllReduce


def test_shard_range_partitions():
    for total in (1, 7, 8, 64, 100):
        for world in (1, 2, 3, 8):
            seen = [i for r in range(world) for i in shard_range(total, r, world)]
            assert seen == list(range(total))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    times, counts = reduce_stats([10.0 + 5 * rank, 3.0 - rank], [8.0, float(rank)])
    dist.barrier()
    q.put((rank, times, counts))
    dist.destroy_process_group()


def test_two_rank_gloo_reduction():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, times, counts in res:
        assert times == [15.0, 3.0]        # slowest rank per region
        assert counts == [16.0, 1.0]       # whole-job units


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                   # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    bucket = FlatGradAllReduce(net.parameters())
    g = torch.Generator().manual_seed(100 + rank)          # each rank has its own shard of the batch
    x = torch.randn(4, 6, generator=g)
    for _ in range(2):                                     # second step: zero_() + in-place accumulation keep the views alive
        bucket.zero_()
        net(x).square().mean().backward()
        local = [p.grad.clone() for p in net.parameters()]
        bucket.allreduce_()
    views_ok = all(p.grad.data_ptr() >= bucket.flat.data_ptr() and
                   p.grad.data_ptr() < bucket.flat.data_ptr() + bucket.nbytes for p in net.parameters())
    q.put((rank, [t.tolist() for t in local], [p.grad.tolist() for p in net.parameters()], views_ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce():
    """The training exchange step: after FlatGradAllReduce every rank holds the MEAN of the per-rank gradients
    (DistributedDataParallel semantics), and the parameters' .grad are still views of the flat buffer."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, r0, v0), (_, l1, r1, v1) = res
    assert v0 and v1
    for a, b, m0, m1 in zip(l0, l1, r0, r1):
        want = (torch.tensor(a) + torch.tensor(b)) / 2
        torch.testing.assert_close(torch.tensor(m0), want)
        torch.testing.assert_close(torch.tensor(m1), want)


def _grad_worker_zero_grad(rank, world, port, q):
    """The reference trainer's loop (trainer/trainer_torchrun.py:272-301): optimizer.zero_grad() with its default
    set_to_none=True drops the flat views; allreduce_() must notice and still exchange the real gradients."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    for p in net[3].parameters():
        p.requires_grad_(True)
    opt = torch.optim.SGD(net.parameters(), lr=0.0)
    bucket = FlatGradAllReduce(net.parameters())
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(4, 6, generator=g)
    for _ in range(2):
        opt.zero_grad()                                    # set_to_none=True: p.grad = None, views dropped
        net[:3](x).square().mean().backward()              # net[3] receives no gradient this step (stays None)
        local = [None if p.grad is None else p.grad.clone() for p in net.parameters()]
        stale = sum(1 for p in net.parameters() if p.grad is None or p.grad.data_ptr() < bucket.flat.data_ptr()
                    or p.grad.data_ptr() >= bucket.flat.data_ptr() + bucket.nbytes)
        bucket.allreduce_()
    views_ok = all(p.grad is not None and bucket.flat.data_ptr() <= p.grad.data_ptr() < bucket.flat.data_ptr() + bucket.nbytes
                   for p in net.parameters())
    q.put((rank, [None if t is None else t.tolist() for t in local], [p.grad.tolist() for p in net.parameters()], views_ok, stale))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_survives_zero_grad_set_to_none():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker_zero_grad, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, r0, v0, s0), (_, l1, r1, v1, s1) = res
    assert v0 and v1 and s0 == 6 and s1 == 6               # every gradient had left the flat buffer before the exchange
    for a, b, m0, m1 in zip(l0, l1, r0, r1):
        if a is None:                                      # no gradient this step: contributes zeros, like DDP
            assert torch.tensor(m0).abs().max() == 0 and torch.tensor(m1).abs().max() == 0
            continue
        want = (torch.tensor(a) + torch.tensor(b)) / 2
        assert want.abs().max() > 0
        torch.testing.assert_close(torch.tensor(m0), want)
        torch.testing.assert_close(torch.tensor(m1), want)


def _ddp_worker(rank, world, port, q):
    """A drop-in model under torch's DistributedDataParallel ('DDP(...) must keep working around the
    patched model', trainer/trainer_torchrun.py:116-121).  CPU: the 3-D path is answered by the oracle's TrainBackend
    stand-in; what is exercised is the model's train-mode host code under the DDP wrapper and its gradient hooks."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import torch.nn.functional as F
    import stereo_toolbox_b200 as S
    import stereo_toolbox_b200.aggregation as agg
    from oracle_backend import OracleTrainBackend, oracle_hot_path
    from stereo_toolbox_b200.synth import synth_gt, synth_pair, synth_state_dict
    net = S.GwcNet_G(32)
    net.load_state_dict(synth_state_dict(net.state_dict(), 0))             # identical replicas
    keys = list(net.state_dict())
    sync_keys = list(torch.nn.SyncBatchNorm.convert_sync_batchnorm(S.GwcNet_G(32)).state_dict())   # trainer_torchrun.py:112-113
    ddp = torch.nn.parallel.DistributedDataParallel(net.train())
    left, right = synth_pair(1, 64, 128, seed=10 + rank, shift=4)          # each rank its own shard
    gt = synth_gt(1, 64, 128)
    mask = (gt > 0) & (gt < 32)
    with oracle_hot_path():
        old, agg.TrainBackend = agg.TrainBackend, OracleTrainBackend
        try:
            preds = ddp(left, right)
            loss = sum(F.smooth_l1_loss(p[mask], gt[mask]) for p in preds)
            loss.backward()
        finally:
            agg.TrainBackend = old
    g = net.dres0[0][0].weight.grad
    q.put((rank, len(preds), loss.item(), g.abs().sum().item(), g.flatten()[:64].tolist(), keys == sync_keys))
    dist.barrier()
    dist.destroy_process_group()


def test_dropin_model_under_ddp_two_ranks():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, n0, loss0, s0, g0, k0), (_, n1, loss1, s1, g1, k1) = res
    assert n0 == n1 == 4 and k0 and k1                  # four training predictions; SyncBatchNorm conversion keeps the state-dict layout
    assert loss0 != loss1                               # different shards ...
    assert s0 > 0 and g0 == g1                          # ... identical (averaged) gradients after DDP's all-reduce
