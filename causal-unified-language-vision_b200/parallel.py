"""Data-parallel LoRA-gradient sync: flat buckets + NCCL all-reduce(mean), optionally overlapped with backward.

What it replaces: the DDP reducer that ``accel.prepare`` installs
(/root/reference/trainer/utils_trainer.py:32-37) and ``accel.backward`` is meant to drive
(/root/reference/trainer/default_trainer.py:83-84).  The reference unwraps ``.module`` right
after ``prepare`` (utils_trainer.py:37), which disarms DDP's reducer, and adds a full barrier
between forward and backward (/root/reference/pipeline/CuLLaVOPipeline.py:87); this module
implements the INTENDED semantics -- mean-reduced gradients every step -- and no barrier
(deviation recorded in DESIGN.md).

Design (one process per GPU, weights replicated, batch sharded):
  * every trainable LoRA ``A [r,K]`` / ``B [N,r]`` gets a view into a flat bucket, assigned in
    REVERSE forward order (the order backward produces them); ``param.grad`` is that view;
  * the backward kernels (``b2q_lora_grads``) write dA / dB straight into the views -- no
    autograd accumulation pass, no flatten / unflatten copies;
  * ``overlap=True`` (default): when the last gradient of a bucket has been produced an event is recorded on the
    compute stream, the comm stream waits on it and issues ONE ``all_reduce`` for the bucket
    (NCCL over NVLink 5 / NVSwitch; ``ReduceOp.AVG``), overlapping the remaining backward, and
    ``finish()`` makes the compute stream wait for the comm stream before the optimizer step;
  * ``overlap=False`` (``B2Q_GRAD_OVERLAP=0``): ``finish()`` issues the all-reduces of all buckets on the compute
    stream -- 360 MB per step, < 1 ms on NVLink;
  * every rank issues exactly one all-reduce per bucket per step, in bucket order, whatever subset of modules ran:
    a bucket is launched early only when all buckets before it have been launched, and ``finish()`` launches the
    rest (a module skipped on one rank -- data-dependent branch, text-only batch -- can therefore neither leave a
    bucket unreduced nor reorder the collectives between ranks);
  * trainable parameters that are NOT LoRA weights (the reference also trains ``multi_modal_projector``,
    /root/reference/cullavo/load_cullavo.py:128-130) are passed as ``extra_params``: autograd produces their ``.grad``
    as usual and ``finish()`` mean-reduces them in one flat fp32 all-reduce.
Gradient accumulation (``accel.accumulate``, /root/reference/trainer/default_trainer.py:167): call ``begin_step()``
ONCE per optimizer step, set ``sync.defer = True`` for all but the last micro-batch (the kernels then accumulate into
the buckets and nothing is reduced) and ``False`` for the last one.
The same class runs on CPU tensors with the ``gloo`` backend (SUM then divide) for tests.

``backend="b2q"`` (or ``B2Q_COMM_BACKEND=b2q``) issues the bucket all-reduces through the library's own C-ABI
communicator (``b2q_comm_*`` in include/b2q.h: NCCL resolved with dlopen, the same entry points a host without a torch
process group would bind) instead of ``torch.distributed``; the unique id travels over the existing process group.
Default off.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


class GradSink:
    """Where one module's LoRA gradients go; handed to ``autograd.QLoRALinear`` and consulted in backward."""

    def __init__(self, dA: torch.Tensor, dB: torch.Tensor, accumulate: Callable[[], bool],
                 on_ready: Optional[Callable[[], None]] = None):
        self.dA, self.dB, self._accumulate, self._on_ready = dA, dB, accumulate, on_ready

    def accumulate(self) -> bool:
        return bool(self._accumulate())

    def ready(self) -> None:
        if self._on_ready is not None:
            self._on_ready()


@dataclass(eq=False)
class _Bucket:
    flat: torch.Tensor
    members: List[int] = field(default_factory=list)  # slot ids
    offsets: List[int] = field(default_factory=list)  # element offset of every member inside `flat`
    pending: int = 0
    launched: bool = False


class GradSync:
    def __init__(self, modules: Sequence, adapter_name: str, process_group=None, bucket_bytes: int = 64 << 20,
                 grad_dtype: torch.dtype = torch.bfloat16, install: bool = True, overlap: Optional[bool] = None,
                 backend: Optional[str] = None, extra_params: Sequence = ()):
        """``modules``: the LoRA-wrapped linears in FORWARD order (``LoraLinear4bit`` or anything with
        ``lora_A[adapter].weight`` / ``lora_B[adapter].weight``).  ``extra_params``: other trainable parameters whose
        autograd-produced ``.grad`` must be mean-reduced with the LoRA gradients."""
        # overlap=True (default, B2Q_GRAD_OVERLAP=0 flips it): a bucket's all-reduce is issued on a side stream as soon as
        # its last gradient kernel is enqueued and runs concurrently with the rest of backward -- the schedule of the DDP
        # reducer this class stands in for.  overlap=False: all buckets are reduced on the compute stream at finish().
        if overlap is None:
            import os
            overlap = os.environ.get("B2Q_GRAD_OVERLAP", "1") == "1"
        self.overlap = bool(overlap)
        self.extra_params = [p for p in extra_params if p.requires_grad]
        self._extra_flat = None
        self.adapter = adapter_name
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.modules = list(modules)
        self.grad_dtype = grad_dtype
        params = []
        for m in reversed(self.modules):  # backward order
            params.append((m, m.lora_B[adapter_name].weight, "B"))
            params.append((m, m.lora_A[adapter_name].weight, "A"))
        if not params:
            raise ValueError("no LoRA modules")
        device = params[0][1].device
        self.device = device
        self.slots = params  # (module, parameter, "A" | "B") in backward order; bucket members index into it
        elem = torch.empty(0, dtype=grad_dtype).element_size()
        # bucket assignment
        plan, cur, cur_bytes = [], [], 0
        for i, (_, prm, _) in enumerate(params):
            nbytes = prm.numel() * elem
            if cur and cur_bytes + nbytes > bucket_bytes:
                plan.append(cur)
                cur, cur_bytes = [], 0
            cur.append(i)
            cur_bytes += nbytes
        if cur:
            plan.append(cur)
        self.buckets: List[_Bucket] = []
        self._views = {}
        self._slot_bucket = {}
        for members in plan:
            total = sum(params[i][1].numel() for i in members)
            flat = torch.zeros(total, dtype=grad_dtype, device=device)
            b = _Bucket(flat=flat, members=list(members))
            off = 0
            for i in members:
                mod, prm, which = params[i]
                # keep 16-byte alignment of every view (kernels use 128-bit accesses)
                assert (off * elem) % 16 == 0
                view = flat[off:off + prm.numel()].view_as(prm)
                b.offsets.append(off)
                off += prm.numel()
                self._views[(id(mod), which)] = view
                self._slot_bucket[(id(mod), which)] = len(self.buckets)
                if prm.dtype != grad_dtype:
                    # the backward kernels write into the bucket views and hand autograd None for A / B: a parameter of
                    # another dtype would never see a gradient -- PEFT creates LoRA weights in fp32, the reference sweeps
                    # them to bf16 before it builds the optimizer (/root/reference/cullavo/load_cullavo.py:124-126)
                    raise TypeError(f"LoRA parameter {which} of {type(mod).__name__} is {prm.dtype}, the gradient buckets "
                                    f"are {grad_dtype}: convert the adapter weights first (the reference's fp32 -> bf16 "
                                    "sweep) or pass grad_dtype=")
                prm.grad = view
            self.buckets.append(b)
        self._written = set()
        self._use_cuda = device.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=device) if self._use_cuda else None
        self._done_events = []
        if backend is None:
            import os
            backend = os.environ.get("B2Q_COMM_BACKEND", "torch")
        if backend not in ("torch", "b2q"):
            raise ValueError(f"unknown gradient-sync backend {backend!r}")
        self.backend = backend
        self._comm = None
        if backend == "b2q" and self.world > 1:
            if not self._use_cuda:
                raise ValueError("backend='b2q' is the NCCL communicator of libb2q.so: CUDA buckets only")
            self._init_b2q_comm()
        if install:
            for m in self.modules:
                m._grad_sinks[adapter_name] = self
        self.begin_step()

    def _init_b2q_comm(self) -> None:
        import ctypes as ct

        from . import _lib
        lib = _lib.load()
        rank = dist.get_rank(self.group)
        ids = [None]
        if rank == 0:
            buf = ct.create_string_buffer(128)
            _lib.check(lib.b2q_comm_unique_id(buf, 128), "b2q_comm_unique_id")
            ids[0] = buf.raw
        dist.broadcast_object_list(ids, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                                   group=self.group)
        comm = ct.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.b2q_comm_init(ct.byref(comm), ids[0], 128, self.world, rank), "b2q_comm_init")
        self._comm = comm

    def close(self) -> None:
        """Destroy the library communicator (backend='b2q'); the torch backend owns nothing."""
        if self._comm is not None:
            from . import _lib
            _lib.check(_lib.load().b2q_comm_destroy(self._comm), "b2q_comm_destroy")
            self._comm = None

    # ---- per step -----------------------------------------------------------------------
    def begin_step(self) -> None:
        """Call after ``zero_grad`` / before the first backward of an optimizer step."""
        self._written.clear()
        self._next_bucket = 0   # buckets [0, _next_bucket) have been launched this step
        for b in self.buckets:
            b.pending = len(b.members)
            b.launched = False
        self._extra_reduced = False
        self._deferred_in_step = False   # a micro-batch was accumulated without reducing: no early launches in this step
        self._done_events = []

    def sink_for(self, mod) -> GradSink:
        """Gradient destination for ``mod`` in the coming backward (used by ``LoraLinear4bit.forward``)."""
        key = id(mod)
        return GradSink(self._views[(key, "A")], self._views[(key, "B")], lambda: self._accumulate(key),
                        lambda: self._ready(mod))

    def _accumulate(self, key) -> bool:
        """Called by backward just before the gradient kernels of a module write: accumulate (second visit) or overwrite?"""
        for which in ("A", "B"):
            if self.buckets[self._slot_bucket[(key, which)]].launched:
                raise RuntimeError(
                    "a LoRA gradient arrived for a bucket whose all-reduce has already been issued in this step: call "
                    "GradSync.begin_step() once per optimizer step, and for gradient accumulation set sync.defer = True "
                    "on all but the last micro-batch (parallel.py docstring)")
        return key in self._written

    def _ready(self, mod) -> None:
        key = id(mod)
        first = key not in self._written
        self._written.add(key)
        if self.defer:
            self._deferred_in_step = True
        if not first:
            return
        for which in ("A", "B"):
            bi = self._slot_bucket[(key, which)]
            b = self.buckets[bi]
            b.pending -= 1
        if self.world > 1 and not self.defer and not self._deferred_in_step and self.overlap:
            # in bucket order only: the collective sequence must be the same on every rank
            while self._next_bucket < len(self.buckets) and self.buckets[self._next_bucket].pending <= 0:
                self._launch(self.buckets[self._next_bucket])

    defer = False  # True while accumulating micro-batches: reduce only on the last one

    def _launch(self, b: _Bucket) -> None:
        assert not b.launched and b is self.buckets[self._next_bucket]
        b.launched = True
        self._next_bucket += 1
        if self._comm is not None:
            from . import _lib
            dtype = {torch.bfloat16: 0, torch.float32: 1}[b.flat.dtype]
            stream = torch.cuda.current_stream(self.device).cuda_stream
            with torch.cuda.device(self.device):
                _lib.check(_lib.load().b2q_comm_allreduce_bucket(self._comm, b.flat.data_ptr(), b.flat.numel(), dtype,
                                                                 1 if self.overlap else 0, stream),
                           "b2q_comm_allreduce_bucket")
        elif self._use_cuda and not self.overlap:
            dist.all_reduce(b.flat, op=dist.ReduceOp.AVG, group=self.group)   # on the compute stream, in order
        elif self._use_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(b.flat, op=dist.ReduceOp.AVG, group=self.group)
                done = torch.cuda.Event()
                done.record(self.comm_stream)
            self._done_events.append(done)
        else:
            dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group)
            b.flat.div_(self.world)

    def reduce_all(self) -> None:
        """Reduce every bucket that has not been reduced in this step, in bucket order (idempotent)."""
        if self.world > 1:
            while self._next_bucket < len(self.buckets):
                self._launch(self.buckets[self._next_bucket])
            self._reduce_extra()

    def _reduce_extra(self) -> None:
        """Mean-reduce the autograd-produced gradients of ``extra_params`` in one flat fp32 all-reduce.  A parameter
        without a gradient on this rank contributes zeros (and receives the mean), so every rank issues the same call."""
        if self._extra_reduced or not self.extra_params:
            return
        self._extra_reduced = True
        n = sum(p.numel() for p in self.extra_params)
        if self._extra_flat is None or self._extra_flat.numel() != n:
            self._extra_flat = torch.zeros(n, dtype=torch.float32, device=self.extra_params[0].device)
        flat, off = self._extra_flat, 0
        for p in self.extra_params:
            seg = flat[off:off + p.numel()]
            if p.grad is None:
                seg.zero_()
            else:
                seg.copy_(p.grad.reshape(-1))
            off += p.numel()
        if self._use_cuda:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)
        off = 0
        for p in self.extra_params:
            seg = flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = seg.to(p.dtype)
            else:
                p.grad.copy_(seg)
            off += p.numel()

    def finish(self) -> None:
        """Issue every all-reduce that has not been issued in this step and order the optimizer step after all of them.
        No-op while ``defer`` is set (accumulating micro-batches)."""
        if self.world > 1 and not self.defer:
            self.reduce_all()
        if self._use_cuda:
            cur = torch.cuda.current_stream(self.device)
            for ev in self._done_events:
                cur.wait_event(ev)
            if self._comm is not None:
                from . import _lib
                _lib.check(_lib.load().b2q_comm_wait(self._comm, cur.cuda_stream), "b2q_comm_wait")
        self._done_events = []

    # ---- helpers ------------------------------------------------------------------------
    def broadcast_parameters(self, src: int = 0) -> None:
        """Make every rank start from rank ``src``'s LoRA weights (what DDP does at construction;
        /root/reference/trainer/utils_trainer.py:35-36 relies on it through ``accel.prepare``)."""
        if self.world <= 1:
            return
        for m in self.modules:
            for which in ("lora_A", "lora_B"):
                dist.broadcast(getattr(m, which)[self.adapter].weight.data, src=src, group=self.group)

    def zero_grad(self) -> None:
        for b in self.buckets:
            b.flat.zero_()
        self.begin_step()

    def grad_bytes(self) -> int:
        return sum(b.flat.numel() * b.flat.element_size() for b in self.buckets)

    def flat_grads(self) -> List[torch.Tensor]:
        return [b.flat for b in self.buckets]
