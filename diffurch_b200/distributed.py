"""Multi-GPU host logic: the ensemble shards by contiguous trajectory range, one
process per GPU, no data-path collective (trajectories never interact; the
reference has no coupling concept).  The only communication is the final gather
or reduction of per-trajectory results over ``torch.distributed`` (NCCL over
NVLink on GPUs; gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(n_total, rank, world):
    """Contiguous range [lo, hi) of rank `rank`; remainders go to the first ranks."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_rows(array, rank, world):
    """This rank's rows of a per-trajectory array (None passes through)."""
    if array is None:
        return None
    lo, hi = shard_bounds(array.shape[0], rank, world)
    return array[lo:hi]


def all_gather_rows(local, group=None, device=None):
    """Concatenate per-trajectory rows of every rank in rank order (ragged shards allowed).

    `local` is a numpy array [n_local, ...].  Uses all_gather on padded tensors, so it
    works with NCCL (pass device="cuda") and gloo alike."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.array(local, copy=True)
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        t = t.to(device)
    n_local = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts)
    pad = torch.zeros((n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)], axis=0)


def all_reduce_sum(values, group=None, device=None):
    """Sum a small vector of ensemble statistics (event counts, exponent sums) over ranks."""
    import torch
    import torch.distributed as dist

    v = np.atleast_1d(np.asarray(values, dtype=np.float64))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return v.copy()
    t = torch.from_numpy(v.copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def run_sharded(make_solver, y0, params, rank, world, runner=None, group=None, device=None):
    """Integrate this rank's shard and gather every rank's final states.

    make_solver(y0_shard, params_shard) -> diffurch_b200.Solver.  `runner(solver)` defaults
    to ``solver.run()`` (the CUDA path); tests pass a CPU checker instead.  Returns a dict
    with the gathered ``t, p, aux, steps, status`` in global trajectory order."""
    y0s, prms = shard_rows(y0, rank, world), shard_rows(params, rank, world)
    solver = make_solver(y0s, prms)
    res = solver.run() if runner is None else runner(solver)
    out = {}
    for name in ("t", "p", "aux", "steps", "events", "status"):
        a = getattr(res, name, None)
        if a is None:
            continue
        a = np.asarray(a)
        if a.ndim == 2 and a.shape[1] == 0:
            continue
        if a.dtype == np.uint64:
            a = a.astype(np.int64)
        if a.dtype == np.uint32:
            a = a.astype(np.int64)
        out[name] = all_gather_rows(a, group=group, device=device)
    return out


class NcclComm:
    """A raw ``ncclComm_t`` for `dfb_gather_final`, created with the NCCL that is already loaded in
    the process (torch's bundled libnccl): rank 0 makes the unique id, torch.distributed broadcasts
    it, every rank calls ``ncclCommInitRank``.  The C ABI takes the handle as ``void*``."""

    def __init__(self, rank, world, group=None):
        import ctypes as C
        import glob
        import os

        import torch
        import torch.distributed as dist

        lib = None
        for cand in (["libnccl.so.2"] +
                     glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so*"))):
            try:
                lib = C.CDLL(cand, mode=C.RTLD_GLOBAL)
                break
            except OSError:
                continue
        if lib is None:
            raise RuntimeError("libnccl not found")

        class UniqueId(C.Structure):
            _fields_ = [("internal", C.c_char * 128)]  # NCCL_UNIQUE_ID_BYTES

        self._lib = lib
        uid = UniqueId()
        if rank == 0:
            rc = lib.ncclGetUniqueId(C.byref(uid))
            if rc != 0:
                raise RuntimeError(f"ncclGetUniqueId failed: {rc}")
        raw = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone()
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else None
        t = raw.to(dev) if dev is not None else raw
        dist.broadcast(t, src=0, group=group)
        C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
        comm = C.c_void_p()
        lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
        rc = lib.ncclCommInitRank(C.byref(comm), int(world), uid, int(rank))
        if rc != 0:
            raise RuntimeError(f"ncclCommInitRank failed: {rc}")
        self.handle = comm
        self.world = int(world)

    def destroy(self):
        if self.handle:
            self._lib.ncclCommDestroy.argtypes = [__import__("ctypes").c_void_p]
            self._lib.ncclCommDestroy(self.handle)
            self.handle = None
