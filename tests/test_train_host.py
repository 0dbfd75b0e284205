"""CPU suite for the training-side host logic: the segmenter's length arithmetic
(qpnet_train.py:268-284) and the data-parallel gradient bucket (world_size 2, gloo)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_segment_geometry_matches_reference_arithmetic():
    from qpnet_b200.train import segment_geometry
    # qpnet_train.py:181-199, 268-284 with the SI defaults: rf 1 / 45 / 15, U 110, batch_length 20000, max 30000
    R, bl, h_bs, x_bs = segment_geometry(61.25, 20000, 110, 1, 45, 15, 30000)
    assert (R, bl, h_bs, x_bs) == (976, 19924, 190, 20901)
    assert (R + bl) % 110 == 0
    # x0.5 F0 -> M = 123: R = 1 + 45 + 15 * 123 = 1891
    R, bl, h_bs, x_bs = segment_geometry(122.5, 20000, 110, 1, 45, 15, 30000)
    assert R == 1891 and (R + bl) % 110 == 0 and bl <= 20000 and x_bs == h_bs * 110 + 1
    # max_length clamp (batch_mod1)
    R, bl, _, _ = segment_geometry(61.25, 29900, 110, 1, 45, 15, 30000)
    assert R + bl <= 30000 and (R + bl) % 110 == 0
    with pytest.raises(ValueError):
        segment_geometry(61.25, 100, 110, 1, 45, 15, 900)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qpnet_b200.train import GradBucket
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 2))]
    params[0].grad = torch.full((3, 4), float(rank + 1))
    params[1].grad = torch.arange(5, dtype=torch.float32) * (rank + 1)
    # params[2] never receives a gradient (the dead resA_1x1 of the last block): must reduce as zeros
    b = GradBucket(params)
    assert b.numel == 12 + 5 + 4 and b.world == world
    flat = b.allreduce_mean()
    mean = (1 + world) / 2.0
    ok = (torch.allclose(params[0].grad, torch.full((3, 4), mean))
          and torch.allclose(params[1].grad, torch.arange(5, dtype=torch.float32) * mean)
          and torch.equal(params[2].grad, torch.zeros(2, 2))
          and flat.numel() == 21)
    # gradients that already live in one flat buffer (what the hand-written backward returns): reduced in place
    from qpnet_b200.qpnet import flat_grad_views
    q = [torch.nn.Parameter(torch.zeros(3, 3)), torch.nn.Parameter(torch.zeros(1)), torch.nn.Parameter(torch.zeros(6))]
    fl, views = flat_grad_views(q)
    assert fl.numel() == 12 + 4 + 8 and all(v.data_ptr() % 16 == 0 for v in views)
    for i, (pp, v) in enumerate(zip(q, views)):
        v.fill_(float((rank + 1) * (i + 1)))
        pp.grad = v
    b2 = GradBucket(q)
    assert b2.flat_in_place(fl)
    ret = b2.allreduce_mean(fl)
    ok = ok and ret is fl and all(torch.allclose(pp.grad, torch.full_like(pp, mean * (i + 1))) for i, pp in enumerate(q))
    q[1].grad = torch.ones(1)                     # a foreign gradient tensor: falls back to the packed bucket
    ok = ok and not b2.flat_in_place(fl)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_grad_bucket_allreduce_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_grad_bucket_single_process_is_identity():
    from qpnet_b200.train import GradBucket
    p = [torch.nn.Parameter(torch.zeros(4))]
    p[0].grad = torch.tensor([1.0, 2.0, 3.0, 4.0])
    b = GradBucket(p)
    assert b.world == 1
    b.allreduce_mean()
    assert torch.equal(p[0].grad, torch.tensor([1.0, 2.0, 3.0, 4.0]))


def test_segmenter_length_arithmetic_matches_oracle():
    """The host half of the device segmenter (qpnet_b200/segmenter.py): _validate_length as pure length arithmetic and
    the per-file segment geometry, against the oracle's array version / the training-step geometry, over ragged lengths."""
    import numpy as np
    from oracle import qpnet_oracle as orc
    from qpnet_b200.segmenter import segment_lengths, validated_lengths
    from qpnet_b200.train import segment_geometry
    rs = np.random.RandomState(0)
    for _ in range(300):
        U = int(rs.choice([80, 110, 120]))
        n_h = int(rs.randint(1, 60))
        n_x = max(1, n_h * U + int(rs.randint(-3 * U, 3 * U)))
        x, h = orc.validate_length(np.zeros(n_x, np.float32), np.zeros((n_h, 2)), U)
        assert validated_lengths(n_x, n_h, U) == (len(x), len(h))
        assert len(x) == len(h) * U
    # same clamps as Trainer's segment_geometry (qpnet_train.py:268-284) for a given receptive field
    for M in (20, 62, 123):
        R, bl, h_bs, x_bs = segment_geometry(M - 0.5, 20000, 110, 1, 45, 15, 30000)
        assert segment_lengths(R, 20000, 30000, 110) == (bl, h_bs, x_bs)
    bl, h_bs, x_bs = segment_lengths(976, 29900, 30000, 110)
    assert 976 + bl <= 30000 and (976 + bl) % 110 == 0 and x_bs == h_bs * 110 + 1



def _reducer_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qpnet_b200.qpnet import QPNet, flat_grad_views
    from qpnet_b200.train import OverlappedReducer
    m = QPNet(n_resch=16, n_skipch=8)          # parameters on the CPU: only the host-side layout / bucket logic runs here
    params = list(m.parameters())
    red = OverlappedReducer(m, limits_mb=(0.02, 0.01, 0.005))
    ranges = red.stage_ranges()
    ok = ranges[0][0] == 0 and ranges[-1][1] == m.n_backward_stages and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    ok = ok and len(ranges) >= 3 and red.buckets[0][2] == 0 and red.buckets[-1][3] == red.total
    ok = ok and all(a[3] == b[2] for a, b in zip(red.buckets, red.buckets[1:]))          # contiguous pieces of the flat buffer
    # emulate the staged backward: every stage writes the gradients of ITS parameters, then its bucket is reduced
    flat, views = flat_grad_views(params, m.flat_order)
    for k, (s0, s1) in enumerate(ranges):
        for i, v in enumerate(views):
            if s0 <= m._stage_of_param[i] < s1:
                v.fill_(float((rank + 1) * (i + 1)))
        lo, hi = red.buckets[k][2], red.buckets[k][3]
        inside = [i for i, v in enumerate(views) if lo <= (v.data_ptr() - flat.data_ptr()) // 4 < hi]
        ok = ok and sorted(inside) == sorted(i for i in range(len(views)) if s0 <= m._stage_of_param[i] < s1)
        red.range_done(k, flat)
    red.finish()
    mean = (1 + world) / 2.0
    ok = ok and all(torch.allclose(v, torch.full_like(v, mean * (i + 1))) for i, v in enumerate(views))
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_reducer_world2_gloo():
    """The bucketed gradient all-reduce that overlaps the staged backward (train.OverlappedReducer): stage ranges cover the
    backward, every bucket is one contiguous piece of the flat gradient buffer holding exactly the parameters its stages
    finish, and after the last bucket every rank holds the mean."""
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_reducer_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
