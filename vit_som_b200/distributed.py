"""Multi-GPU execution of the SOM hot path: one process per GPU, ``torch.distributed`` (NCCL over NVLink 5 /
NVSwitch) for the exchange steps.  The reference has no distributed code of its own; what it gets implicitly is
Lightning's DDP (``experiments/benchmarking/train_vit_som.py:45,86-87``), i.e. a batch-sharded data-parallel
step with an all-reduce of every gradient including ``som_layer.prototypes``.

Two partitionings:

* :class:`DataParallelSOM` - batch rows sharded, prototypes replicated.  The only exchange is the prototype
  gradient ``dW[K, D]`` (mean over ranks, DDP semantics).  The all-reduce is issued on a side stream as soon as the
  dW GEMM is enqueued, so it runs under the dx GEMM (and, in a full model, under the ViT backward).
* :class:`PrototypeShardedSOM` - prototypes sharded in contiguous row blocks of the map (rank r owns map cells
  ``[r*K/G, (r+1)*K/G)``), latents replicated.  Exchanges per step: (1) the per-row packed ``(key, global index)``
  minima, ``all_reduce(MIN)`` on int64 - B * 8 bytes; NCCL has no MINLOC, the packing gives the first-index
  tie-break of ``torch.argmin``; (2) the scalar loss, ``all_reduce(SUM)``; (3) ``dx[B, D]``, ``all_reduce(SUM)``
  of the per-shard partial gradients.  ``dW`` of the local shard needs no exchange.

The helpers at the top are device-agnostic (they run on CPU tensors over gloo in the unit tests); the numeric
work stays in libsom_b200.so.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import ops
from ._lib import SomError
from .som_layer import NeighbourhoodWeights, SOMLayer

INT64_MAX = 0x7FFFFFFFFFFFFFFF


# ------------------------------------------------------------------------------------------------------------
# exchange helpers (device-agnostic)
# ------------------------------------------------------------------------------------------------------------
def shard_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [begin, end) of ``n_items`` owned by ``rank``; the first ``n_items % world`` ranks get one
    extra item, so any map size shards over any world size."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def reduce_packed_min(packed: torch.Tensor, group=None) -> torch.Tensor:
    """In-place elementwise signed-int64 MIN over ranks of the packed (ordered key << 32 | global index) minima.
    The smaller key wins; on equal keys the smaller global index wins (= torch.argmin over the full map)."""
    if packed.dtype != torch.int64:
        raise TypeError("packed minima must be int64")
    dist.all_reduce(packed, op=dist.ReduceOp.MIN, group=group)
    return packed


def unpack_bmu(packed: torch.Tensor, k_total: int) -> torch.Tensor:
    """Global BMU index = low 32 bits of the packed minimum (rows that never saw a finite key fall back to 0,
    like som_bmu_decode)."""
    idx = packed & 0xFFFFFFFF
    return torch.where(idx < k_total, idx, torch.zeros_like(idx))


def all_reduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_reduce_mean(t: torch.Tensor, group=None) -> torch.Tensor:
    """DDP semantics for a replicated parameter's gradient: sum over ranks / world size.  On NCCL the division is
    part of the collective (ReduceOp.AVG: one kernel, no extra pass over the gradient); gloo has no AVG."""
    if t.is_cuda and dist.get_backend(group) == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(dist.get_world_size(group))
    return t


# ------------------------------------------------------------------------------------------------------------
# batch-sharded data parallel
# ------------------------------------------------------------------------------------------------------------
class DataParallelSOM:
    """Attach DDP-equivalent gradient averaging to a :class:`SOMLayer` whose batch is sharded over ranks.

    ``DataParallelSOM(layer)`` broadcasts the prototypes from rank 0 and installs itself as the layer's dW hook.  The
    exchange of dW (+ 1/world scaling) runs on a communication stream beside the rest of the backward - and, in a full
    model, under the ViT backward; the compute stream joins before backward returns dW to autograd.  The local loss is
    the mean over the local rows, as under DDP.  ``overlap`` selects how the exchange is started:

    * ``"split"`` (default; fastest in the measurements of round 2): the dW GEMM is a launch of its own, the exchange
      starts when it is done and runs beside the dx GEMM, which leaves ``148 - gemm_sm_limit`` SMs to it;
    * ``"counter"``: the backward stays ONE launch for both gradient GEMMs with a two-phase stream-K schedule (every CTA
      pair works off its share of the dW tiles first); the dW epilogues raise a counter and the exchange sits behind a
      stream-ordered wait on it (``som_stream_wait_value``: executed by the GPU front end, no SM is held while waiting),
      so it runs under the dx half of the same launch;
    * ``"after"``: one fused launch, exchange after it (no overlap; the baseline of the other two).

    Measured, config 2 (dW = 20 MB), ms per step at 1 / 2 / 8 GPUs: split 0.199 / 0.240 / 0.240, counter 0.199 / 0.248 /
    0.245, after - / 0.289 / -.  The exchange itself takes 75-80 us at any GPU count (two-shot NVLS all-reduce at the
    NVSwitch's rate, 274 GB/s algorithm bandwidth) and cannot start before dW is complete (~65-75 us into the backward),
    which bounds the step of this layer-only benchmark at ~0.23 ms; in ViT-SOM training it hides under the ViT backward."""

    def __init__(self, layer: SOMLayer, group=None, broadcast: bool = True, gemm_sm_limit: int | None = None,
                 nvls: bool | None = None, overlap: str = "split"):
        if not dist.is_initialized():
            raise SomError("DataParallelSOM needs an initialised torch.distributed process group")
        if overlap not in ("counter", "split", "after"):
            raise ValueError("overlap must be 'counter' (fused backward, exchange started by the dW-complete counter), "
                             "'split' (dW launch -> exchange beside a separate dx launch) or 'after' (fused backward, "
                             "exchange after it)")
        self.overlap = overlap
        self.layer, self.group = layer, group
        self.nvls = None
        dev = layer.prototypes.device
        self.world = dist.get_world_size(group)
        # SMs the gradient GEMMs may occupy while the exchange runs beside them (0 = all)
        self.gemm_sm_limit = int(gemm_sm_limit) if gemm_sm_limit is not None and dev.type == "cuda" else 0
        # lowest priority: when the exchange and GEMM tiles become runnable together the GEMM's CTA pairs are placed first
        self.comm_stream = torch.cuda.Stream(dev, priority=0) if dev.type == "cuda" else None
        self._counter = torch.zeros(4, device=dev, dtype=torch.int32) if dev.type == "cuda" else None
        # grid of the NVLS exchange kernel: two blocks per SM the gradient GEMMs leave free (it is bound by the bytes it
        # keeps in flight; blocks that do not fit beside the GEMM would only start after it)
        self.nvls_blocks = 32
        if dev.type == "cuda" and self.gemm_sm_limit:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            self.nvls_blocks = max(8, min(64, 2 * (sms - self.gemm_sm_limit)))
        if broadcast:
            # in place on the parameter itself (not on .data): the version counter moves, so a prototype staging that
            # was cached before wrapping is not reused with the broadcast values
            with torch.no_grad():
                dist.broadcast(layer.prototypes, src=dist.get_global_rank(group, 0) if group is not None else 0,
                               group=group)
        layer.invalidate_staging()
        layer._dw_hook = self
        want_nvls = nvls if nvls is not None else os.environ.get("SOM_DP_NVLS", "1") != "0"
        if want_nvls and dev.type == "cuda" and self.world > 1:
            self.nvls = self._setup_nvls(layer, dev)
            if nvls and self.nvls is None:
                raise SomError("NVLS all-reduce requested but symmetric multicast memory is not available")

    # ---- NVLS (NVLink SHARP) exchange: our own two-shot multimem kernel over torch symmetric memory -------------
    def _setup_nvls(self, layer, dev):
        """Allocate dW and the barrier flags in symmetric memory and rendezvous; None when multicast is unavailable
        (then the NCCL all-reduce is used).  Every rank takes the same decision (it is all-reduced)."""
        state = None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            from . import _lib
            grp = self.group if self.group is not None else dist.group.WORLD
            K, D = layer.prototypes.shape
            n = K * D
            if n % 4 == 0:
                dw = symm_mem.empty((K, D), dtype=torch.float32, device=dev)
                hdl = symm_mem.rendezvous(dw, grp)
                words = int(_lib.lib().som_nvls_flag_words(self.world))
                flags = symm_mem.empty((max(words, 1024),), dtype=torch.int32, device=dev)
                flags.zero_()
                fh = symm_mem.rendezvous(flags, grp)
                if int(getattr(hdl, "multicast_ptr", 0) or 0) != 0:
                    state = {"dw": dw, "hdl": hdl, "flags": flags, "fh": fh, "n": n,
                             "mc": int(hdl.multicast_ptr), "flag_ptrs": int(fh.buffer_ptrs_dev)}
        except Exception as exc:  # noqa: BLE001
            self.nvls_error = repr(exc)
            state = None
        ok = torch.tensor([1 if state is not None else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        torch.cuda.synchronize(dev)
        if int(ok.item()) == 0:
            return None
        dist.barrier(group=self.group)                       # flags are zero everywhere before the first kernel
        layer._dw_out = state["dw"]
        return state

    # ---- hook protocol (called from ops.FusedLossFn.backward) ----------------------------------------------------
    def counter_ptr(self):
        """Device word the dW epilogues raise; also forks the communication stream off the compute stream (what is
        enqueued on it afterwards is ordered behind everything that precedes the backward launch - and belongs to the
        same CUDA-graph capture)."""
        if self.comm_stream is None or self.overlap != "counter":
            return None
        self.comm_stream.wait_stream(torch.cuda.current_stream(self._counter.device))
        return self._counter.data_ptr()

    def _reduce(self, dw: torch.Tensor):
        nv = self.nvls
        if nv is not None and dw.data_ptr() == nv["dw"].data_ptr():
            from . import _lib
            _lib.check(_lib.lib().som_allreduce_mean_nvls(nv["mc"], nv["flag_ptrs"], nv["n"], dist.get_rank(self.group),
                                                          self.world, self.nvls_blocks, _lib.stream_ptr(dw.device)),
                       "som_allreduce_mean_nvls")
        else:
            all_reduce_mean(dw, self.group)

    def exchange_counted(self, dw: torch.Tensor, expected: int):
        """The fused backward launch is enqueued and will raise the counter to ``expected`` when its last dW tile is
        written: wait for that on the communication stream, reset the word, run the exchange.  Returns the join."""
        from . import _lib
        L = _lib.lib()
        cur = torch.cuda.current_stream(dw.device)
        with torch.cuda.stream(self.comm_stream):
            sp = _lib.stream_ptr(dw.device)
            _lib.check(L.som_stream_wait_value(self._counter.data_ptr(), expected, sp), "som_stream_wait_value")
            self._reduce(dw)
            # reset for the next backward (ordered before it through the join), off the exchange's critical path
            _lib.check(L.som_stream_write_value(self._counter.data_ptr(), 0, sp), "som_stream_write_value")
        dw.record_stream(self.comm_stream)
        return lambda: cur.wait_stream(self.comm_stream)

    def exchange_after(self, dw: torch.Tensor):
        """dW has been enqueued by a kernel of its own on the current stream: exchange it once that kernel is done
        (beside whatever the compute stream runs next).  Returns the join callable (None on CPU tensors)."""
        if self.comm_stream is None:                       # CPU tensors (gloo unit tests)
            all_reduce_mean(dw, self.group)
            return None
        cur = torch.cuda.current_stream(dw.device)
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            self._reduce(dw)
        dw.record_stream(self.comm_stream)
        return lambda: cur.wait_stream(self.comm_stream)

    _on_dw = exchange_after

    def reduce_accumulator(self):
        """Row-chunked batches accumulate dW in ``layer.grad_accumulator`` with the hook detached (see ``detach``);
        call this once after the last chunk to average the accumulated gradient over the ranks."""
        acc = self.layer.grad_accumulator
        if acc is None:
            raise SomError("no grad_accumulator is set on the layer")
        all_reduce_mean(acc, self.group)
        return acc

    def detach(self):
        self.layer._dw_hook = None
        self.layer._dw_out = None


# ------------------------------------------------------------------------------------------------------------
# prototype-sharded
# ------------------------------------------------------------------------------------------------------------
class _DxExchange:
    """Hook object of a prototype shard for ``ops.FusedLossFn.backward`` (asynchronous mode, ``dx_overlap="kernel"``):
    the fused backward computes the dx tiles FIRST and raises a counter; the exchange of the partial dx sits behind a
    stream-ordered wait on it and runs on the SMs the second phase (the dW tiles) leaves free - so even the LAST row
    chunk's exchange is hidden.  ``exchange_*`` return the tensor handed to autograd (complete after ``wait_dx``)."""

    def __init__(self, layer, nv):
        self.layer, self.nv = layer, nv
        self.gemm_sm_limit = layer.gemm_sm_limit
        sms = torch.cuda.get_device_properties(nv["buf"].device).multi_processor_count
        self.blocks = max(8, min(64, 2 * (sms - self.gemm_sm_limit))) if self.gemm_sm_limit else layer.async_blocks

    def counter_ptr(self):
        lay = self.layer
        dev = self.nv["buf"].device
        if lay._dx_counter is None:
            lay._dx_counter = torch.zeros(4, device=dev, dtype=torch.int32)
        lay._comm_stream(dev).wait_stream(torch.cuda.current_stream(dev))       # fork (also joins a graph capture)
        return lay._dx_counter.data_ptr()

    def _run(self, dx, expected):
        from . import _lib
        L, lay, nv = _lib.lib(), self.layer, self.nv
        dev = dx.device
        comm = lay._comm_stream(dev)
        out = torch.empty_like(dx)
        if expected is None:
            comm.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(comm):
            sp = _lib.stream_ptr(dev)
            if expected is not None:
                _lib.check(L.som_stream_wait_value(lay._dx_counter.data_ptr(), expected, sp), "som_stream_wait_value")
            _lib.check(L.som_allreduce_nvls(nv["mc"], nv["flag_ptrs"], nv["n"], nv["rank"], nv["world"], 1.0,
                                            self.blocks, sp), "som_allreduce_nvls")
            out.copy_(dx)
            if expected is not None:
                _lib.check(L.som_stream_write_value(lay._dx_counter.data_ptr(), 0, sp), "som_stream_write_value")
        done = torch.cuda.Event()
        done.record(comm)
        out.record_stream(comm)
        nv["pending"] = done
        lay._dx_events.append(done)
        return out

    def exchange_counted(self, dx, expected):
        return self._run(dx, expected)

    def exchange_after(self, dx):
        return self._run(dx, None)


class _ShardedLossFn(torch.autograd.Function):
    """Global loss of a prototype-sharded map as one autograd node over (x, W_shard): local fused loss kernel,
    scalar all-reduce; backward = local gradient GEMMs + all-reduce of the partial dx."""

    @staticmethod
    def forward(ctx, x, W, state, bmu, grid_pos, T_dev, inv_count, k_offset, want_grad, group, grid_dims):
        loss = ops.FusedLossFn.forward(ctx, x, W, state, bmu, grid_pos, T_dev, inv_count, k_offset, want_grad, None,
                                       grid_dims)
        ctx.group = group
        ctx.nvls = getattr(state, "nvls_dx", None)
        layer = getattr(state, "layer", None)
        if layer is not None and layer.async_loss and loss.is_cuda:
            # the scalar all-reduce leaves the compute stream's critical path: the VALUE of the loss is complete after
            # wait_dx() (backward does not need it - its upstream gradient is independent of the loss value)
            cur = torch.cuda.current_stream(loss.device)
            comm = layer._comm_stream(loss.device)
            comm.wait_stream(cur)
            with torch.cuda.stream(comm):
                all_reduce_sum(loss, group)
            done = torch.cuda.Event()
            done.record(comm)
            loss.record_stream(comm)
            layer._dx_events.append(done)
            return loss
        return all_reduce_sum(loss, group)

    @staticmethod
    def backward(ctx, g_out):
        nv = ctx.nvls
        layer = getattr(ctx.state, "layer", None)
        if nv is not None and nv.get("pending") is not None:
            # the symmetric buffer still carries an exchange of an earlier call (asynchronous mode): the gradient GEMM
            # of this call may only overwrite it once that exchange and its copy-out are done
            torch.cuda.current_stream(nv["buf"].device).wait_event(nv["pending"])
            nv["pending"] = None
        if (nv is not None and layer is not None and layer.async_dx and layer.dx_overlap == "kernel"
                and ctx.needs_input_grad[0] and ctx.x_dtype == torch.float32):
            ctx.state.dx_hook = _DxExchange(layer, nv)
        grads = ops.FusedLossFn.backward(ctx, g_out)
        dx = grads[0]
        if ctx.state.dx_exchanged:
            return grads[:11]                       # exchanged (asynchronously) from inside the fused backward
        if dx is not None:
            if nv is not None and dx.data_ptr() == nv["buf"].data_ptr() and dx.dtype == torch.float32:
                # partial dx of all shards summed in the NVSwitch (our two-shot multimem kernel), in place in the
                # symmetric buffer; autograd gets a private copy because the buffer is reused by a later call
                from . import _lib
                L = _lib.lib()
                if layer is not None and layer.async_dx:
                    # asynchronous: exchange and copy-out run on the communication stream, under the next row chunk's
                    # kernels; the caller joins with layer.wait_dx() before it reads the latent gradients
                    cur = torch.cuda.current_stream(dx.device)
                    comm = layer._comm_stream(dx.device)
                    out = torch.empty_like(dx)
                    comm.wait_stream(cur)
                    with torch.cuda.stream(comm):
                        _lib.check(L.som_allreduce_nvls(nv["mc"], nv["flag_ptrs"], nv["n"], nv["rank"], nv["world"], 1.0,
                                                        layer.async_blocks, _lib.stream_ptr(dx.device)), "som_allreduce_nvls")
                        out.copy_(dx)
                    done = torch.cuda.Event()
                    done.record(comm)
                    out.record_stream(comm)
                    nv["pending"] = done
                    layer._dx_events.append(done)
                    grads = (out,) + tuple(grads[1:])
                else:
                    _lib.check(L.som_allreduce_nvls(nv["mc"], nv["flag_ptrs"], nv["n"], nv["rank"], nv["world"], 1.0, 64,
                                                    _lib.stream_ptr(dx.device)), "som_allreduce_nvls")
                    grads = (dx.clone(),) + tuple(grads[1:])
            else:
                all_reduce_sum(dx, ctx.group)
        return grads[:11]


class PrototypeShardedSOM(SOMLayer):
    """A SOM whose prototypes are sharded over the ranks of ``group`` (large maps, BASELINE config 5).

    Same constructor dict and call protocol as :class:`SOMLayer`; ``prototypes`` holds only the local block
    ``[k_begin, k_end)`` of the map (drawn from the same RNG stream as the full map so that a seeded construction
    matches the unsharded layer), ``grid_positions`` holds the full grid.  ``forward`` returns the *local* slice
    of the distance matrix ``[B, K_local]`` and the *global* BMU indices; ``som_loss`` returns the global loss."""

    def __init__(self, config, group=None):
        if not dist.is_initialized():
            raise SomError("PrototypeShardedSOM needs an initialised torch.distributed process group")
        super().__init__(config)
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.k_total = self.n_prototypes
        self.k_begin, self.k_end = shard_range(self.k_total, self.world, self.rank)
        if self.k_end <= self.k_begin:
            raise ValueError(f"map of {self.k_total} cells cannot be sharded over {self.world} ranks")
        full = self.prototypes.data
        self.prototypes = torch.nn.Parameter(full[self.k_begin:self.k_end].clone())
        self._nvls_dx = {}                    # (B, D, slot) -> symmetric dx buffer + multicast mapping (None: NCCL)
        self.use_nvls = os.environ.get("SOM_DP_NVLS", "1") != "0"
        # Asynchronous exchange of the latent gradients (opt-in): backward returns as soon as the exchange of the partial
        # dx is ENQUEUED on a communication stream, so it runs under the kernels of the next row chunk (two symmetric
        # buffers alternate).  The returned gradient is complete only after wait_dx(): use it when the latents are
        # leaves of the graph (row-chunked scoring of a large map); leave it off when dx flows on into an encoder.
        self.async_dx = False
        self.async_blocks = 32                # grid of the exchange kernel in asynchronous mode (two blocks per SM)
        # how the asynchronous exchange overlaps: "stream" - enqueued behind the backward launch, runs under whatever
        # the compute stream does next (the next row chunk); "kernel" - the fused backward computes the dx tiles first
        # and its dW half leaves gemm_sm_limit .. 148 SMs to the exchange, which a counter starts (hides the exchange
        # of a chunk that has no successor, at the price of a second phase on fewer SMs)
        self.dx_overlap = "stream"
        # asynchronous sum of the scalar loss over the shards (opt-in): som_loss returns at once, the returned tensor
        # holds the global loss only after wait_dx() - do not compute with it on the device before that
        self.async_loss = False
        self.gemm_sm_limit = 136
        self._dx_counter = None
        self._dx_events = []
        self._dx_turn = 0
        self._comm = None

    def _comm_stream(self, dev):
        if self._comm is None:
            self._comm = torch.cuda.Stream(dev)
        return self._comm

    def wait_dx(self):
        """Join the asynchronous dx exchanges enqueued so far: afterwards the current stream sees complete latent
        gradients (no host synchronisation)."""
        if self._dx_events:
            cur = torch.cuda.current_stream(self.prototypes.device)
            for ev in self._dx_events:
                cur.wait_event(ev)
            self._dx_events = []
            for nv in self._nvls_dx.values():             # the current stream is now behind every exchange
                if nv is not None:
                    nv["pending"] = None

    # ---- checkpoints: the state dict holds the FULL map under the reference's key ---------------------------------
    def gather_prototypes(self) -> torch.Tensor:
        """The full [K, D] map assembled from the shards of all ranks (a collective: every rank must call it)."""
        local = self.prototypes.detach()
        sizes = [shard_range(self.k_total, self.world, r) for r in range(self.world)]
        longest = max(e - b for b, e in sizes)
        padded = local.new_zeros((longest, local.shape[1]))
        padded[:local.shape[0]] = local
        parts = [torch.empty_like(padded) for _ in range(self.world)]
        dist.all_gather(parts, padded, group=self.group)
        return torch.cat([p[:e - b] for p, (b, e) in zip(parts, sizes)], dim=0)

    def state_dict(self, *args, **kwargs):
        """Same keys and shapes as the unsharded layer (``prototypes`` [K, D], ``grid_positions``): a checkpoint written
        by a sharded run loads into ``SOMLayer`` / the reference and vice versa.  Collective (all ranks call it)."""
        sd = super().state_dict(*args, **kwargs)
        prefix = kwargs.get("prefix", args[1] if len(args) > 1 else "")
        key = prefix + "prototypes"
        if key in sd:
            sd[key] = self.gather_prototypes()
        return sd

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        key = prefix + "prototypes"
        full = state_dict.get(key)
        if full is not None and full.shape[0] == self.k_total and self.k_total != self.prototypes.shape[0]:
            state_dict = dict(state_dict)
            state_dict[key] = full[self.k_begin:self.k_end]       # this rank's block of the full map
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self.invalidate_staging()

    def _nvls_dx_buffer(self, B: int, D: int, dev, slot: int = 0):
        """Symmetric [B, D] buffer for the partial dx of this rank plus what som_allreduce_nvls needs; allocated and
        rendezvoused once per shape and slot (a collective: every rank reaches it with the same shape at the same call)."""
        key = (B, D, slot)
        if key in self._nvls_dx:
            return self._nvls_dx[key]
        state = None
        if self.use_nvls and dev.type == "cuda" and self.world > 1 and (B * D) % 4 == 0:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                from . import _lib
                grp = self.group if self.group is not None else dist.group.WORLD
                buf = symm_mem.empty((B, D), dtype=torch.float32, device=dev)
                hdl = symm_mem.rendezvous(buf, grp)
                if not hasattr(self, "_nvls_flags"):
                    words = int(_lib.lib().som_nvls_flag_words(self.world))
                    flags = symm_mem.empty((max(words, 1024),), dtype=torch.int32, device=dev)
                    flags.zero_()
                    self._nvls_flags = (flags, symm_mem.rendezvous(flags, grp))
                if int(getattr(hdl, "multicast_ptr", 0) or 0) != 0:
                    state = {"buf": buf, "hdl": hdl, "n": B * D, "mc": int(hdl.multicast_ptr), "pending": None,
                             "flag_ptrs": int(self._nvls_flags[1].buffer_ptrs_dev), "rank": self.rank, "world": self.world}
            except Exception as exc:  # noqa: BLE001
                self.nvls_error = repr(exc)
                state = None
            ok = torch.tensor([1 if state is not None else 0], device=dev, dtype=torch.int32)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                state = None
                self.use_nvls = False
            else:
                torch.cuda.synchronize(dev)
                dist.barrier(group=self.group)           # flags are zero everywhere before the first kernel
        self._nvls_dx[key] = state
        return state

    def _forward_impl(self, x, want_dist=True):
        if x.dim() > 2:
            x = x.flatten(start_dim=1)
        if not x.is_cuda or not self.prototypes.is_cuda:
            raise SomError("PrototypeShardedSOM runs on B200s only: move the module and its input to cuda")
        if x.dim() != 2 or x.shape[1] != self.latent_dim:
            raise ValueError(f"latent shape {tuple(x.shape)} does not end in latent_dim {self.latent_dim}")
        mode = self._mode()
        ws, refill = self._staged_prototypes(mode)
        try:
            state, _ = ops.forward(x, self.prototypes, mode, ws, refill, want_dist=want_dist, want_bmu=False,
                                   idx_offset=self.k_begin, k_total=self.k_total)
        except Exception:
            self.invalidate_staging()
            raise
        reduce_packed_min(state.packed, self.group)          # exchange step 1: B x 8 bytes over NVLink
        bmu = ops.bmu_decode(state.packed, self.k_total, state=state)
        if not want_dist:
            return None, bmu
        state.x_in, state.W_in = x, self.prototypes
        state.grad_accum = self.grad_accumulator
        state.nvls_dx = None
        state.layer = self
        slot = 0
        if self.async_dx and torch.is_grad_enabled() and x.requires_grad:
            slot = self._dx_turn                         # two buffers alternate under the asynchronous exchange
            self._dx_turn ^= 1
        if torch.is_grad_enabled() and x.requires_grad and not torch.cuda.is_current_stream_capturing():
            nv = self._nvls_dx_buffer(state.B, state.D, x.device, slot)
        else:
            nv = self._nvls_dx.get((state.B, state.D, slot))   # capture / no_grad: only what already exists
        if nv is not None:
            state.dx_out, state.nvls_dx = nv["buf"], nv
        dist_local = ops.DistanceFn.apply(x, self.prototypes, state)
        dist_local._som_state = state
        return dist_local, bmu

    def som_loss(self, weights, distances):
        state = getattr(distances, "_som_state", None)
        if not (isinstance(weights, NeighbourhoodWeights) and weights._dense is None and state is not None):
            raise SomError("PrototypeShardedSOM.som_loss needs the lazy weights of compute_weights() and the "
                           "distances returned by this layer's forward")
        B = distances.shape[0] if self.batch_rows is None else int(self.batch_rows)
        want_grad = torch.is_grad_enabled() and (state.x_in.requires_grad or state.W_in.requires_grad)
        return _ShardedLossFn.apply(state.x_in, state.W_in, state, weights.bmu, self.grid_positions, weights.T_dev,
                                    1.0 / (B * self.k_total), self.k_begin, want_grad, self.group,
                                    self._square_grid_dims())
