"""World-size-2 gloo run of the multi-GPU host logic (range all-reduce, unit sharding) on CPU --
the way the reference tests its only distributed piece (tests/test_quantized_tensor_fsdp.py:57-119)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from fastforward_b200 import distributed as D

    g = torch.Generator().manual_seed(100 + rank)
    # each rank saw different data: ranges of three quantizers (per-tensor, per-channel fp32, per-channel bf16)
    ranges = []
    for n, dt in ((1, torch.float32), (16, torch.float32), (8, torch.bfloat16)):
        data = torch.randn(n, 64, generator=g).to(dt)
        ranges.append((data.min(1).values.clone(), data.max(1).values.clone()))
    local = [(a.clone(), b.clone()) for a, b in ranges]
    flags = [torch.zeros(1, dtype=torch.int32), torch.full((1,), rank, dtype=torch.int32), None]
    D.all_reduce_ranges(ranges, flags)
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    for i, (mn, mx) in enumerate(ranges):
        want_min = torch.stack([gathered[r][i][0] for r in range(world)]).min(0).values
        want_max = torch.stack([gathered[r][i][1] for r in range(world)]).max(0).values
        assert torch.equal(mn, want_min) and torch.equal(mx, want_max), (rank, i)
    assert int(flags[0]) == 0 and int(flags[1]) == world - 1          # "+-inf seen" flag ORs across ranks
    units = D.shard_units(10)
    assert units == list(range(rank, 10, world))
    allu = [None] * world
    dist.all_gather_object(allu, units)
    assert sorted(u for part in allu for u in part) == list(range(10))
    out.put((rank, "ok"))
    dist.destroy_process_group()


def test_range_allreduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(out.get(timeout=5) for _ in range(2)) == [(0, "ok"), (1, "ok")]


def test_single_process_is_a_noop():
    from fastforward_b200 import distributed as D

    r = [(torch.tensor([1.0]), torch.tensor([2.0]))]
    D.all_reduce_ranges(r)
    assert float(r[0][0]) == 1.0 and D.shard_units(5, rank=1, world_size=2) == [1, 3]
    packed, sizes = D.pack_ranges(r)
    assert packed.tolist() == [1.0, -2.0] and sizes == [1]


# ---- the estimator's block-exit exchange when the ranks did not see the same quantizers -------------------------
class _DummyQuantizer:
    """A RangeSettable that is not a LinearQuantizer: the exchange goes through its own setter."""
    num_bits, symmetric, allow_one_sided, has_uninitialized_params, granularity = 8, False, True, False, None

    def __init__(self):
        self._range = (None, None)

    @property
    def quantization_range(self):
        return self._range

    @quantization_range.setter
    def quantization_range(self, value):
        self._range = (value[0].clone(), value[1].clone())


def _sync_worker(rank, world, port, out):
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from fastforward_b200.range_setting import minmax as M

    est = M.RunningMinMaxRangeEstimator(sync_ranges=True)
    quantizers = [_DummyQuantizer() for _ in range(3)]
    steps = [(q, M.RunningMinMaxEstimator(q, state=est._state)) for q in quantizers]
    g = torch.Generator().manual_seed(7 + rank)
    # quantizer 0: seen by both ranks; quantizer 1: seen by rank 0 only (an expert without tokens on rank 1);
    # quantizer 2: seen by nobody
    local = {}
    for i, (q, s) in enumerate(steps):
        if i == 0 or (i == 1 and rank == 0):
            data = torch.randn(5, 32, generator=g)
            s.min, s.max = data.min(1).values.clone(), data.max(1).values.clone()
            local[i] = (s.min.clone(), s.max.clone())
    est._sync(steps, est._state)
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    want0 = (torch.stack([gathered[r][0][0] for r in range(world)]).min(0).values,
             torch.stack([gathered[r][0][1] for r in range(world)]).max(0).values)
    got0 = quantizers[0].quantization_range
    assert torch.equal(got0[0], want0[0]) and torch.equal(got0[1], want0[1]), rank
    got1 = quantizers[1].quantization_range           # rank 1 adopts what rank 0 measured
    assert torch.equal(got1[0], gathered[0][1][0]) and torch.equal(got1[1], gathered[0][1][1]), rank
    assert quantizers[2].quantization_range == (None, None)
    out.put((rank, "ok"))
    dist.destroy_process_group()


def test_estimator_exchange_with_rank_dependent_quantizer_sets_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(out.get(timeout=5) for _ in range(2)) == [(0, "ok"), (1, "ok")]
