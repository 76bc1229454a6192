"""Data-parallel exchange of the head's parameter gradients (one image per GPU, one exchange per step;
the reference wraps the model in DistributedDataParallel, uwsod/projects/WSL/tools/train_net_multi.py:75-78, which
all-reduces 483.8 MB of fp32 gradients per rank and step and then runs the full optimizer on every rank).

Three modes behind one object, all giving every rank the SAME fp32 result as DDP's average:

  "allreduce"  what DDP does: an averaging all-reduce per gradient, started from the engine's gradient hook as soon as
               the producing kernel is queued.
  "sharded"    for the two big matrices (fc1.weight 411 MB, fc2.weight 67 MB = 99 % of the bytes): REDUCE-SCATTER the
               fp32 gradient (rank r receives the averaged rows it owns, in place), update only those rows of the fp32
               master / momentum with the fused SGD pass -- which writes the rows of the bf16 GEMM operand -- and
               ALL-GATHER the bf16 operand rows.  Bytes on the wire per rank: (n-1)/n x (484 + 242) MB instead of
               2 (n-1)/n x 484 MB, optimizer work 1/n, and the all-gather runs behind the next step's ROI pooling
               (the engine waits for it right before the first GEMM, engine.operand_gate).  The fp32 masters of rows a
               rank does not own are refreshed on demand (`sync_master()`: state_dict / checkpoint time); the forward
               and backward only ever read the bf16 operands.  Small tensors (biases, the fused head block) are
               all-reduced.

  "nvls"       the sharded scheme with NO collective library on the big tensors: the weight gradients of fc1 / fc2 and
               their bf16 operands live in symmetric memory mapped for NVSwitch multicast
               (torch.distributed._symmetric_memory); ONE kernel per rank (soswsod_sgd_nvls) reads the sum over the
               ranks of its rows' gradient through the switch (multimem.ld_reduce), applies the update and stores the
               refreshed operand rows to every rank (multimem.st).  Two cross-rank barriers bracket it.  The backward
               runs without any NCCL kernel next to it (no row panels needed); the small tensors are all-reduced.

Every collective is issued with async_op=True from the stream that produced its input; `Work.wait()` orders the
consumer's stream behind it -- no host synchronisation anywhere.  The consumer is the exchange's UPDATE STREAM
(`update_stream()`): solver.B200SGD queues "wait for the collectives -> fused SGD on the owned rows -> operand
all-gather" there, so the compute stream goes straight on to the next step's ROI pooling (which reads no weights) and
only waits at `operand_gate()`, right before its first GEMM.  The engine produces fc1.weight's gradient LAST in the
backward, which puts the tail of the exchange + update + all-gather under that pooling instead of in front of it.  Backends without reduce_scatter_tensor / AVG (gloo,
used by the CPU tests of this host logic) take an all-reduce based path with the same result."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


class GradientExchange:
    SHARDED_KEYS = ("fc1_w", "fc2_w")

    def __init__(self, master: Dict[str, torch.Tensor], group=None, mode: str = "sharded", min_shard_elems: int = 1 << 20):
        """master: HeadOperands.master (key -> fp32 parameter).  mode: "sharded" | "allreduce"."""
        if not dist.is_initialized():
            raise RuntimeError("GradientExchange needs an initialised torch.distributed process group")
        assert mode in ("sharded", "allreduce", "nvls")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.mode = mode
        self.backend = dist.get_backend(group)
        self.native = self.backend == "nccl"
        self.master = master
        self.sharded = set()
        self.symm = None                 # nvls: handle of the symmetric arena
        self.symm_tensors: Dict[str, torch.Tensor] = {}     # nvls: "g:<key>" fp32 gradient, "w0:<key>" / "w1:<key>" bf16 operands
        self.operand_slot = 0            # nvls: which operand copy the engine computes with in the current step
        self._grad_ready: Dict[str, torch.cuda.Event] = {}   # nvls: recorded behind the kernel that produced a big gradient
        self._symm_offsets: Dict[str, int] = {}
        if mode in ("sharded", "nvls") and self.world > 1:
            for k in self.SHARDED_KEYS:
                t = master[k]
                if t.numel() >= min_shard_elems and t.size(0) % self.world == 0:
                    self.sharded.add(k)
        self._works: List = []                         # all-reduces of this step
        self._rs: List[Tuple[str, torch.Tensor, int, object]] = []     # (key, panel, row0, work) reduce-scatters of this step
        self._gather: List = []                        # operand all-gathers still in flight
        self.master_stale = False
        self._update_stream = None
        self._update_done = None                       # event on the update stream: this step's parameter update is queued
        # True: `backward()` returns without ordering the compute stream behind the collectives -- the attached optimizer
        # consumes the gradients on the update stream.  False: DDP's contract (gradients are final when backward returns).
        self.lazy_wait = False
        self._expected_ptrs: Dict[str, int] = {}
        self._layout: Dict[str, List[Tuple[int, int]]] = {}
        self.bytes_last_step = {"reduce_scatter": 0, "all_reduce": 0, "all_gather": 0}

    # ------------------------------------------------------------------ gradient side
    def begin_step(self) -> None:
        """Start of a training step: nothing of a previous step may be left pending (a step whose gradients were never
        consumed by an optimizer is dropped after ordering the stream behind its collectives)."""
        self.wait_gradients()
        self._rs.clear()
        self.bytes_last_step = {"reduce_scatter": 0, "all_reduce": 0, "all_gather": 0}

    # ------------------------------------------------------------------ nvls: symmetric multicast memory
    def setup_nvls(self) -> None:
        """Allocates ONE symmetric arena holding, for every sharded matrix, its fp32 gradient and its bf16 GEMM operand,
        and maps it for multicast.  Raises RuntimeError when the platform has no NVLS multicast."""
        import torch.distributed._symmetric_memory as symm_mem

        dev = next(iter(self.master.values())).device
        off, plan = 0, []
        for key in sorted(self.sharded):
            shape = tuple(self.master[key].shape)
            n = self.master[key].numel()
            # the gradient, and TWO operand copies: the fused update of step i writes the copy step i + 1 computes with
            # while step i's own input-gradient GEMM is still reading the other one
            for tag, dt, nbytes in (("g", torch.float32, 4 * n), ("w0", torch.bfloat16, 2 * n), ("w1", torch.bfloat16, 2 * n)):
                plan.append((f"{tag}:{key}", dt, shape, off, nbytes))
                off += (nbytes + 255) // 256 * 256
        arena = symm_mem.empty(off, dtype=torch.uint8, device=dev)
        group = self.group if self.group is not None else dist.group.WORLD
        self.symm = symm_mem.rendezvous(arena, group=group)
        if not getattr(self.symm, "has_multicast_support", False) or int(self.symm.multicast_ptr) == 0:
            raise RuntimeError("GradientExchange(mode='nvls'): the symmetric-memory rendezvous gave no multicast address "
                               "(no NVSwitch multicast on this platform); use mode='sharded'")
        self._arena = arena
        for name, dt, shape, o, nbytes in plan:
            self.symm_tensors[name] = arena[o:o + nbytes].view(dt).view(shape)
            self._symm_offsets[name] = o
        self._layout = {key: [(0, self.master[key].size(0))] for key in self.sharded}

    def multicast_address(self, name: str, row0: int = 0) -> int:
        """Multicast virtual address of row `row0` of symmetric tensor `name` ("g:fc1_w", "w:fc2_w", ...)."""
        t = self.symm_tensors[name]
        return int(self.symm.multicast_ptr) + self._symm_offsets[name] + row0 * t.stride(0) * t.element_size()

    def operand(self, key: str, slot: Optional[int] = None) -> torch.Tensor:
        return self.symm_tensors[f"w{self.operand_slot if slot is None else slot}:{key}"]

    def wait_big_gradients(self) -> None:
        """Orders the current stream behind the kernels that produced the big weight gradients of this step."""
        for ev in self._grad_ready.values():
            torch.cuda.current_stream().wait_event(ev)
        self._grad_ready.clear()

    def barrier(self) -> None:
        """Cross-rank barrier on the current stream (signal pads of the symmetric arena; no host synchronisation)."""
        self.symm.barrier(channel=0)

    def hook(self, key: str, grad: torch.Tensor, row0: int) -> None:
        """The engine's grad_hook: starts the collective of one gradient (or one row panel of fc1.weight)."""
        if self.world == 1:
            return
        if self.mode == "nvls" and key in self.sharded:
            # nothing to start: the engine wrote the gradient straight into the symmetric buffer; the fused update pulls
            # the ranks' sum through the switch
            if grad.data_ptr() != self.symm_tensors[f"g:{key}"].data_ptr() + row0 * grad.stride(0) * 4:
                raise RuntimeError(f"GradientExchange(nvls): the gradient of {key} was not produced in the symmetric buffer")
            ev = torch.cuda.Event()
            ev.record()          # the update stream waits for THIS, not for the end of the backward
            self._grad_ready[key] = ev
            self.bytes_last_step["nvls_ld_reduce"] = self.bytes_last_step.get("nvls_ld_reduce", 0) + grad.numel() * 4 // self.world
            self.bytes_last_step["nvls_multicast_store"] = self.bytes_last_step.get("nvls_multicast_store", 0) + grad.numel() * 2 // self.world
            return
        # Work on an independent tensor over the same storage: a reference to `grad` itself (or to a view of it, which
        # keeps its base alive) held by this object or by the process group would make autograd's AccumulateGrad CLONE the
        # gradient instead of adopting the buffer -- and the clone would be taken before the collective has run.
        grad = torch.empty(0, dtype=grad.dtype, device=grad.device).set_(grad.untyped_storage(), grad.storage_offset(),
                                                                        grad.shape, grad.stride())
        if key in self.sharded:
            rows = grad.size(0)
            if rows % self.world != 0:
                raise RuntimeError(f"GradientExchange: a {rows}-row panel of {key} does not split over {self.world} ranks")
            self.bytes_last_step["reduce_scatter"] += grad.numel() * 4
            if self.native:
                mine = grad.chunk(self.world, 0)[self.rank]
                w = dist.reduce_scatter_tensor(mine, grad, op=dist.ReduceOp.AVG, group=self.group, async_op=True)   # in place
            else:
                grad.div_(self.world)
                w = dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._rs.append((key, grad, row0, w))
        else:
            self.bytes_last_step["all_reduce"] += grad.numel() * 4
            if self.native:
                w = dist.all_reduce(grad, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            else:
                grad.div_(self.world)
                w = dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._works.append(w)

    def wait_allreduces(self) -> None:
        """Orders the current stream behind the all-reduces of this step (the small tensors; in "allreduce" mode all)."""
        for w in self._works:
            w.wait()
        self._works.clear()

    def wait_gradients(self) -> None:
        """Orders the current stream behind every gradient collective of this step."""
        self.wait_allreduces()
        for _, _, _, w in self._rs:
            w.wait()

    def owned_rows_nvls(self, key: str) -> Tuple[int, int]:
        rows = self.master[key].size(0)
        per = rows // self.world
        return self.rank * per, (self.rank + 1) * per

    def owned_rows(self, key: str) -> List[Tuple[int, int]]:
        """Row ranges [lo, hi) of parameter `key` this rank updates (after wait_gradients): one per exchanged panel for
        a sharded tensor -- the averaged gradient of exactly those rows sits in the gradient tensor -- else all rows."""
        if key not in self.sharded:
            return [(0, self.master[key].size(0))]
        if self.mode == "nvls":
            return [self.owned_rows_nvls(key)]
        out = []
        for k, panel, row0, _ in self._rs:
            if k == key:
                per = panel.size(0) // self.world
                out.append((row0 + self.rank * per, row0 + (self.rank + 1) * per))
        return out

    # ------------------------------------------------------------------ operand side
    def gather_operands(self, operands: Dict[str, torch.Tensor]) -> None:
        """After the sharded update: all-gather the bf16 operand rows (in place) of every panel exchanged this step.
        operands: key -> bf16 [rows, cols] GEMM operand of master[key]."""
        for key, panel, row0, _ in self._rs:
            opnd = operands[key][row0:row0 + panel.size(0)]
            self.bytes_last_step["all_gather"] += opnd.numel() * 2
            if self.native:
                mine = opnd.chunk(self.world, 0)[self.rank]
                self._gather.append(dist.all_gather_into_tensor(opnd, mine, group=self.group, async_op=True))
            else:
                per = panel.size(0) // self.world
                parts = [torch.empty_like(opnd[:per]) for _ in range(self.world)]
                dist.all_gather(parts, opnd[self.rank * per:(self.rank + 1) * per].contiguous(), group=self.group)
                for r, t in enumerate(parts):
                    opnd[r * per:(r + 1) * per].copy_(t)
        if self._rs:
            self.master_stale = True
            lay: Dict[str, List[Tuple[int, int]]] = {}
            for key, panel, row0, _ in self._rs:
                lay.setdefault(key, []).append((row0, row0 + panel.size(0)))
            self._layout = lay
        self._rs.clear()

    def update_stream(self):
        if self._update_stream is None:
            self._update_stream = torch.cuda.Stream(device=next(iter(self.master.values())).device)
        return self._update_stream

    def mark_update_done(self) -> None:
        """Called on the update stream after the optimizer's launch (and the operand all-gathers) are queued."""
        self._update_done = torch.cuda.Event()
        self._update_done.record()

    def expect_gradient_buffers(self, ptrs: Dict[str, int]) -> None:
        """The storage addresses of the gradient tensors the collectives of this step work on (in place)."""
        self._expected_ptrs = dict(ptrs)

    def check_gradient_buffer(self, key: str, grad: torch.Tensor) -> None:
        """The optimizer's `.grad` must BE the buffer the collective reduced in place; if autograd had to clone it (a
        second reference was alive) the clone holds this rank's un-averaged values."""
        want = self._expected_ptrs.get(key)
        if want is not None and grad.data_ptr() != want:
            raise RuntimeError(f"GradientExchange: the gradient of {key} is not the buffer the collective reduced (autograd "
                               "cloned it: something else held a reference to the step's gradient tensors)")

    def operand_gate(self) -> None:
        """engine.operand_gate: the current stream waits for the parameter update and the operand all-gathers before
        the first GEMM reads the weights."""
        if self._update_done is not None:
            torch.cuda.current_stream().wait_event(self._update_done)
            self._update_done = None
        for w in self._gather:
            w.wait()
        self._gather.clear()

    def gather_rows(self, t: torch.Tensor, key: str) -> None:
        """All-gathers (in place, blocking the stream) the rows of a [rows, cols] tensor laid out like parameter `key`:
        every rank contributes the rows it owns."""
        for lo, hi in self._panel_ranges(key):
            blk = t[lo:hi]
            if self.native:
                dist.all_gather_into_tensor(blk, blk.chunk(self.world, 0)[self.rank], group=self.group)
            else:
                per = (hi - lo) // self.world
                parts = [torch.empty_like(blk[:per]) for _ in range(self.world)]
                dist.all_gather(parts, blk[self.rank * per:(self.rank + 1) * per].contiguous(), group=self.group)
                for r, part in enumerate(parts):
                    blk[r * per:(r + 1) * per].copy_(part)

    def sync_master(self) -> None:
        """Brings the fp32 master rows owned by other ranks up to date (checkpoint / state_dict time, or before the
        operands are re-cast from the masters).  Collective: every rank must call it."""
        if not self.master_stale or self.world == 1:
            return
        self.operand_gate()
        for key in sorted(self.sharded):      # same order on every rank (set iteration order is per-process)
            self.gather_rows(self.master[key].detach(), key)
        self.master_stale = False

    def _panel_ranges(self, key: str) -> List[Tuple[int, int]]:
        """Panel layout of a sharded parameter as last exchanged (kept so that sync_master can run between steps)."""
        return self._layout.get(key, [(0, self.master[key].size(0))])
