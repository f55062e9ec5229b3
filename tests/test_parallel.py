"""Data-parallel gradient sync host logic on CPU: bucket layout + world_size-2 gloo all-reduce(mean)."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeLora(nn.Module):
    """Has the attribute surface GradSync needs (lora_A/lora_B ModuleDicts, _grad_sinks)."""

    def __init__(self, n, k, r, name="a"):
        super().__init__()
        self.lora_A = nn.ModuleDict({name: nn.Linear(k, r, bias=False)})
        self.lora_B = nn.ModuleDict({name: nn.Linear(r, n, bias=False)})
        self._grad_sinks = {}


def _mods():
    torch.manual_seed(0)
    return [FakeLora(64, 32, 8), FakeLora(32, 64, 8), FakeLora(128, 32, 8)]


def test_bucket_layout_reverse_order_and_views():
    par = importlib.import_module("causal-unified-language-vision_b200.parallel")
    mods = _mods()
    gs = par.GradSync(mods, "a", bucket_bytes=3000, grad_dtype=torch.float32)
    assert len(gs.buckets) > 1
    assert gs.grad_bytes() == sum(p.numel() * 4 for m in mods for p in m.parameters())
    # first bucket starts with the LAST module's B (what backward produces first)
    first = gs.buckets[0].flat
    assert mods[-1].lora_B["a"].weight.grad.data_ptr() == first.data_ptr()
    for m in mods:
        assert m.lora_A["a"].weight.grad.shape == m.lora_A["a"].weight.shape
        assert m._grad_sinks["a"] is gs
    # writing through a sink lands in param.grad; the accumulate flag flips after the first write of a step
    s = gs.sink_for(mods[0])
    assert not s.accumulate()
    s.dA.fill_(2.0)
    s.ready()
    assert gs.sink_for(mods[0]).accumulate()
    assert float(mods[0].lora_A["a"].weight.grad.mean()) == 2.0
    gs.begin_step()
    assert not gs.sink_for(mods[0]).accumulate()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = importlib.import_module("causal-unified-language-vision_b200.parallel")
    mods = _mods()
    gs = par.GradSync(mods, "a", bucket_bytes=3000, grad_dtype=torch.float32)
    # replicas that start from different weights are made identical by the parameter broadcast
    with torch.no_grad():
        for m in mods:
            m.lora_A["a"].weight.add_(float(rank))
    gs.broadcast_parameters(0)
    w0 = float(mods[0].lora_A["a"].weight.sum())
    gs.begin_step()
    for i, m in enumerate(reversed(mods)):  # backward order
        s = gs.sink_for(m)
        s.dA.fill_(float(rank + 1) * (i + 1))
        s.dB.fill_(float(rank + 1) * 10 * (i + 1))
        s.ready()
    gs.finish()
    out = [(float(m.lora_A["a"].weight.grad.mean()), float(m.lora_B["a"].weight.grad.mean())) for m in reversed(mods)]
    q.put((rank, (out, w0)))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_allreduce_mean_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    # mean over ranks of (rank+1)*c = 1.5*c
    assert res[0][1] == res[1][1]  # broadcast_parameters: both ranks hold rank 0's weights
    for r in (0, 1):
        for i, (a, b) in enumerate(res[r][0]):
            assert abs(a - 1.5 * (i + 1)) < 1e-6 and abs(b - 15.0 * (i + 1)) < 1e-6


def test_parameter_dtype_must_match_the_buckets():
    """PEFT creates LoRA weights in fp32; with bf16 buckets such a parameter would never see a gradient (the kernels write
    into the bucket views and autograd gets None) -- GradSync must refuse instead of training nothing."""
    par = importlib.import_module("causal-unified-language-vision_b200.parallel")
    with pytest.raises(TypeError, match="fp32 -> bf16"):
        par.GradSync(_mods(), "a", grad_dtype=torch.bfloat16)


def test_write_after_launch_is_refused_and_accumulation_defers():
    par = importlib.import_module("causal-unified-language-vision_b200.parallel")
    mods = _mods()
    gs = par.GradSync(mods, "a", bucket_bytes=3000, grad_dtype=torch.float32, overlap=True)
    gs.world = 2                       # pretend: exercise the launch bookkeeping without a process group
    launched = []
    gs._launch_real = gs._launch

    def fake_launch(b):                # record instead of calling a collective
        assert b is gs.buckets[gs._next_bucket]
        b.launched = True
        gs._next_bucket += 1
        launched.append([id(x) for x in gs.buckets].index(id(b)))
    gs._launch = fake_launch
    gs._reduce_extra = lambda: None
    # micro-batch 1 of 2: nothing may be reduced
    gs.begin_step()
    gs.defer = True
    for m in reversed(mods):
        s = gs.sink_for(m)
        assert not s.accumulate()
        s.ready()
    gs.finish()
    assert launched == []
    # micro-batch 2: kernels accumulate, still no early launch (a bucket is complete only after ALL its modules re-ran)
    gs.defer = False
    for m in reversed(mods):
        s = gs.sink_for(m)
        assert s.accumulate()
        s.ready()
        assert launched == []
    gs.finish()
    assert launched == list(range(len(gs.buckets)))
    # a gradient that arrives after its bucket went out is an error, not a silent overwrite of a buffer in flight
    with pytest.raises(RuntimeError, match="begin_step"):
        gs.sink_for(mods[0]).accumulate()
    # next step, one module skipped: early launches stop at the first incomplete bucket, finish() sends the rest in order
    launched.clear()
    gs.begin_step()
    for m in reversed(mods[1:]):
        s = gs.sink_for(m)
        s.accumulate()
        s.ready()
    early = list(launched)
    gs.finish()
    assert launched == list(range(len(gs.buckets))) and len(early) < len(gs.buckets)
    assert early == list(range(len(early)))


def _worker_skip_and_extra(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = importlib.import_module("causal-unified-language-vision_b200.parallel")
    mods = _mods()
    torch.manual_seed(1)
    proj = nn.Linear(4, 3)             # a non-LoRA trainable module (the reference's multi_modal_projector)
    frozen = nn.Linear(2, 2).requires_grad_(False)
    gs = par.GradSync(mods, "a", bucket_bytes=3000, grad_dtype=torch.float32, overlap=True,
                      extra_params=list(proj.parameters()) + list(frozen.parameters()))
    assert len(gs.extra_params) == 2
    gs.zero_grad()
    for i, m in enumerate(reversed(mods)):
        if rank == 1 and i == 1:
            continue                    # rank 1 never runs this module (data-dependent branch): must neither hang nor skip
        s = gs.sink_for(m)
        s.accumulate()
        s.dA.fill_(float(rank + 1))
        s.dB.fill_(float(rank + 1))
        s.ready()
    proj.weight.grad = torch.full_like(proj.weight, float(rank + 1))   # autograd's job; bias has no grad on rank 1
    if rank == 0:
        proj.bias.grad = torch.full_like(proj.bias, 4.0)
    gs.finish()
    out = [float(m.lora_A["a"].weight.grad.mean()) for m in reversed(mods)]
    q.put((rank, (out, float(proj.weight.grad.mean()), float(proj.bias.grad.mean()))))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_skipped_module_and_extra_params_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_skip_and_extra, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert res[0] == res[1]
    lora, w, b = res[0]
    assert lora == [1.5, 0.5, 1.5]      # skipped on rank 1: mean of (1, 0)
    assert w == 1.5 and b == 2.0        # extra parameters: mean over ranks, a missing gradient counts as zero
