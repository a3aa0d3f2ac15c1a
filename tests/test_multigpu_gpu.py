"""Data-parallel calibration over NCCL (needs >= 2 GPUs; skipped on a 1-GPU box): the ranges every
rank ends up with after ``estimate_ranges(..., running_minmax(sync_ranges=True))`` over its shard of
the batches are bit-identical to a single-GPU calibration over the union of the batches
(``disable_quantization=True``: SURVEY.md section 7, order-independence caveat)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(dev):
    import bench_workloads as bw
    import fastforward_b200 as ff

    model = bw.DecoderStack(bw.TINY, dtype=torch.bfloat16, device=dev)
    bw.init_weights_(model, seed=3)
    bw.quantize_for_w8a8(ff, model)
    model.to(dev)
    return ff, model


def _batches():
    g = torch.Generator().manual_seed(77)
    return [torch.randint(0, 1024, (1, 64), generator=g) for _ in range(4)]


def _ranges(ff, model):
    out = {}
    for name, q in ff.nn.named_quantizers(model):
        out[name] = (q.scale.detach().cpu().clone(), None if q.offset is None else q.offset.detach().cpu().clone())
    return out


def _worker(rank, world, port, out, fused=False):
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    ff, model = _build(dev)
    if fused:
        from fastforward_b200.nn import qlinear
        qlinear.install()
    est = ff.range_setting.running_minmax(disable_quantization=not fused, sync_ranges=True)
    with torch.no_grad(), ff.strict_quantization(False), ff.estimate_ranges(model, est):
        for b in _batches()[rank::world]:
            model(b.to(dev))
    out.put((rank, _ranges(ff, model)))
    dist.destroy_process_group()


@pytest.mark.skipif(NGPU < 2, reason="needs 2 GPUs")
def test_dp_calibration_matches_single_gpu():
    import torch.multiprocessing as mp

    ff, model = _build(torch.device("cuda", 0))
    with torch.no_grad(), ff.strict_quantization(False), ff.estimate_ranges(
            model, ff.range_setting.running_minmax(disable_quantization=True)):
        for b in _batches():
            model(b.to("cuda:0"))
    want = _ranges(ff, model)

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank in (0, 1):
        assert got[rank].keys() == want.keys()
        for name, (s, o) in want.items():
            gs, go = got[rank][name]
            assert torch.equal(gs, s), f"rank {rank} {name}: scale differs"
            assert (o is None and go is None) or torch.equal(go, o), f"rank {rank} {name}: offset differs"


@pytest.mark.skipif(NGPU < 2, reason="needs 2 GPUs")
def test_dp_fused_calibration_ranks_agree():
    """The production schedule (fused steps, dedupe, memoisation, W8A8 linears, activations-only exchange): after the
    block every rank holds the same parameters, and the weight parameters equal a single-GPU calibration."""
    import torch.multiprocessing as mp

    ff, model = _build(torch.device("cuda", 0))
    from fastforward_b200.nn import qlinear
    qlinear.install()
    with torch.no_grad(), ff.estimate_ranges(model, ff.range_setting.running_minmax()):
        for b in _batches()[0::2]:
            model(b.to("cuda:0"))
    single = _ranges(ff, model)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, True)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got[0].keys() == got[1].keys() == single.keys()
    for name in got[0]:
        assert torch.equal(got[0][name][0], got[1][name][0]), f"{name}: scale differs between ranks"
        if got[0][name][1] is not None:
            assert torch.equal(got[0][name][1], got[1][name][1]), f"{name}: offset differs between ranks"
        if name.endswith("weight_quantizer"):
            assert torch.equal(got[0][name][0], single[name][0]), f"{name}: weight scale differs from single GPU"
