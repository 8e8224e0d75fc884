"""Data parallelism for the hot path.  The reference is single-process / single-GPU; this is new.

One process per GPU (torchrun).  A training mini-batch of B samples is sharded by rows (rank r takes rows
[r*B/W, (r+1)*B/W)); every rank computes d(mean-CE over the GLOBAL batch)/d(theta) on its shard (the loss head is
given denom = B), so ONE sum-allreduce of the flat gradient buffer yields the exact global-batch gradient and all
replicas then run the identical fused update.  Importance passes are sharded by WHOLE batches (the reference squares
/ abs-es the batch-summed gradient, main_EWC.py:151-156, train_MAS.py:163-177) followed by one allreduce of omega.

`torch.distributed` is the rendezvous plumbing; the data-path collective is ncclAllReduce on a communicator owned by
libclb (clb_nccl_*), enqueued on the compute stream.  With the `gloo` backend (CPU tests) the collective falls back
to torch.distributed.all_reduce on host tensors -- that path exists for host-logic tests only.
"""
import ctypes
import os

import torch

from . import _capi

_state = {"world": 1, "rank": 0, "comm": None, "backend": None}


def world_size():
    return _state["world"]


def rank():
    return _state["rank"]


def is_distributed():
    return _state["world"] > 1


class local_only:
    """Context manager: inside it this rank behaves like a single-GPU job (no collectives are issued).  bench.py uses it to
    time a rank's step WITHOUT its all-reduces, which gives the exposed communication time of the data-parallel step."""

    def __enter__(self):
        self._saved = _state["world"]
        _state["world"] = 1
        return self

    def __exit__(self, *exc):
        _state["world"] = self._saved
        return False


def init(backend=None):
    """Initialise from torchrun's env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).  No-op for a single process."""
    import torch.distributed as td
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return
    rk = int(os.environ["RANK"])
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rk)))
    if not td.is_initialized():
        td.init_process_group(backend=backend, rank=rk, world_size=world)
    _state.update(world=world, rank=rk, backend=backend)
    if torch.cuda.is_available():
        # own NCCL communicator behind the C ABI; the 128-byte unique id travels through torch.distributed
        ident = (ctypes.c_char * 128)()
        if rk == 0:
            _capi.call("clb_nccl_unique_id", ctypes.addressof(ident))
        box = [bytes(ident)]
        td.broadcast_object_list(box, src=0)
        ident = (ctypes.c_char * 128).from_buffer_copy(box[0])
        comm = ctypes.c_void_p()
        _capi.call("clb_nccl_init", ctypes.addressof(ident), rk, world, ctypes.addressof(comm))
        _state["comm"] = comm
        _comm_stream()


def shutdown():
    import torch.distributed as td
    if _state["comm"] is not None:
        # NCCL does not release a communicator while a captured CUDA graph still references it
        from . import engine as _engine
        for e in set(_engine._ENGINES.values()):
            e.drop_graphs()
        import gc
        gc.collect()
        torch.cuda.synchronize()
        _capi.call("clb_nccl_destroy", _state["comm"])
        _state["comm"] = None
    if td.is_available() and td.is_initialized():
        td.destroy_process_group()
    _state.update(world=1, rank=0, backend=None, stream=None)


def shard_rows(n, world=None, rk=None):
    """Row range [lo, hi) of a mini-batch of n samples owned by this rank (equal shards; remainder to low ranks)."""
    world = _state["world"] if world is None else world
    rk = _state["rank"] if rk is None else rk
    base, rem = divmod(n, world)
    lo = rk * base + min(rk, rem)
    return lo, lo + base + (1 if rk < rem else 0)


def shard_batches(n_batches, world=None, rk=None):
    """Indices of the importance-pass batches this rank processes (round robin: r, r+W, ...)."""
    world = _state["world"] if world is None else world
    rk = _state["rank"] if rk is None else rk
    return list(range(rk, n_batches, world))


def allreduce_flat(t):
    """In-place sum-allreduce of a flat fp32 tensor on the current stream."""
    if _state["world"] <= 1:
        return
    if t.is_cuda:
        _capi.call("clb_nccl_allreduce_f32", _state["comm"], t.data_ptr(), t.numel(),
                   torch.cuda.current_stream().cuda_stream)
    else:
        import torch.distributed as td
        td.all_reduce(t)


def _comm_stream():
    if _state.get("stream") is None:
        _state["stream"] = _capi.private_stream("comm")
    return _state["stream"]


def start_tail_allreduce(engine, cut, end=None):
    """Called by Engine.backward once every gradient at flat offset >= cut is final: all-reduce grad[cut:end] (one bucket of
    Engine._dp_buckets) on the communication stream while the compute stream goes on with the backward pass of the layers
    in front.  All collectives of a step are issued on the communication stream, in the same order on every rank."""
    main, side = torch.cuda.current_stream(), _comm_stream()
    ev = torch.cuda.Event()
    ev.record(main)
    side.wait_event(ev)
    tail = engine.grad[cut:end]
    _capi.call("clb_nccl_allreduce_f32", _state["comm"], tail.data_ptr(), tail.numel(), side.cuda_stream)
    engine.n_launch += 1
    return cut


def allreduce_grads(engine):
    """Sum the flat gradient over ranks.  If backward() already started the tail (dp_overlap), only the head is left:
    it goes on the communication stream behind the tail, and the compute stream waits for both."""
    if _state["world"] <= 1:
        return
    cut = getattr(engine, "_dp_pending", None)
    engine._dp_pending = None
    if not cut:
        allreduce_flat(engine.grad)
        engine.n_launch += 1
        return
    main, side = torch.cuda.current_stream(), _comm_stream()
    ev = torch.cuda.Event()
    ev.record(main)
    side.wait_event(ev)
    head = engine.grad[:cut]
    _capi.call("clb_nccl_allreduce_f32", _state["comm"], head.data_ptr(), head.numel(), side.cuda_stream)
    done = torch.cuda.Event()
    done.record(side)
    main.wait_event(done)
    engine.n_launch += 1


def broadcast_flat(t, src=0):
    """Rank `src`'s copy of a flat tensor to every rank (replica initialisation; torch.distributed plumbing)."""
    if _state["world"] <= 1:
        return
    import torch.distributed as td
    td.broadcast(t, src=src)


def sync_replicas(engine):
    """Data parallelism assumes bit-identical replicas: broadcast rank 0's parameters and optimiser / importance state
    (fresh heads come from each rank's host generator; nothing else guarantees that they agree).  Called by the trainers
    at the start of every train_model and by the importance passes."""
    if _state["world"] <= 1:
        return
    for name in ("theta", "omega", "theta_star", "momentum", "w"):
        t = getattr(engine, name, None)
        present = torch.tensor([0 if t is None else 1])
        if _state["backend"] == "nccl":
            present = present.cuda()
        import torch.distributed as td
        td.all_reduce(present, op=td.ReduceOp.MIN)
        if int(present.item()) == 1:                       # a buffer that some rank lacks cannot be diverged state
            broadcast_flat(t)


def shared_seed():
    """One seed for all ranks (drawn by rank 0 from its host generator): the trainers re-seed the host generator with it
    at the start of every epoch, so that DataLoader shuffles and host-drawn dropout masks agree across ranks."""
    import torch.distributed as td
    dev = "cuda" if _state["backend"] == "nccl" else "cpu"
    t = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).to(dev)
    td.broadcast(t, src=0)
    return int(t.item())


def allreduce_scalars(vals):
    """Sum python numbers over ranks (epoch statistics: running loss / corrects)."""
    if _state["world"] <= 1:
        return list(vals)
    import torch.distributed as td
    dev = "cuda" if _state["backend"] == "nccl" else "cpu"
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    td.all_reduce(t)
    return t.tolist()
