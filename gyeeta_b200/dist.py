"""torch.distributed plumbing of the multi-GPU merge step: one all-reduce per reduction kind over the
engine's merge arena (u64 SUM, i64 MAX, u8 MAX) and one all-gather of the fixed t-digest slab, NCCL over NVLink/NVSwitch.
The sketches themselves never leave the device; torch only wraps the device pointers."""
import numpy as np

RED_SUM_U64, RED_MAX_U8, RED_MAX_I64 = 0, 1, 2


def shard_of_host(host_idx, world):
    """the partition rule of the ingest path: a host (partha) belongs to rank host_idx % world"""
    return np.asarray(host_idx) % world


class _DevBuf:
    """device pointer -> object torch.as_tensor() understands"""

    def __init__(self, ptr, nbytes, typestr, itemsize):
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr, "data": (ptr, False), "version": 3}


def wrap(torch, ptr, nbytes, redop, device):
    if redop == RED_MAX_U8:
        return torch.as_tensor(_DevBuf(ptr, nbytes, "|u1", 1), device=device)
    return torch.as_tensor(_DevBuf(ptr, nbytes, "<i8", 8), device=device)     # u64 sums == i64 sums bit for bit


_wrap_cache = {}


def _cached(torch, key, ptr, nbytes, redop, device):
    t = _wrap_cache.get(key)
    if t is None or t.data_ptr() != ptr or t.numel() * t.element_size() != nbytes:
        t = wrap(torch, ptr, nbytes, redop, device)
        _wrap_cache[key] = t
    return t


def merge_global(eng, torch, dist, device=None):
    """fold -> all-reduce / all-gather -> finish, all in stream order on the engine's own CUDA stream (no host sync: NCCL is
    enqueued behind the fold kernels, the finish kernel behind NCCL). Returns the device time of the collectives in ms."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    world = dist.get_world_size() if dist.is_initialized() else 1
    eng.merge_prepare()
    if world == 1:
        eng.merge_finish(None, 1)
        return 0.0
    stream = torch.cuda.ExternalStream(eng.stream(), device=device)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        t0.record()
        for k, (_name, ptr, nbytes, redop) in enumerate(eng.merge_buffers()):
            t = _cached(torch, (id(eng), k), ptr, nbytes, redop, device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM if redop == RED_SUM_U64 else dist.ReduceOp.MAX)
        sptr, snbytes = eng.merge_tdigest_slab()
        slab = _cached(torch, (id(eng), "slab"), sptr, snbytes, RED_MAX_U8, device)
        gathered = _wrap_cache.get((id(eng), "gathered"))
        if gathered is None or gathered.numel() != world * snbytes:
            gathered = torch.empty(world * snbytes, dtype=torch.uint8, device=device)
            _wrap_cache[(id(eng), "gathered")] = gathered
        dist.all_gather_into_tensor(gathered, slab)
        t1.record()
    eng.merge_finish(gathered.data_ptr(), world)
    merge_global.last_events = (t0, t1)
    return 0.0


def nccl_comm_init(eng, dist):
    """create the library-side NCCL communicator (gysk_nccl_comm_init): rank 0's unique id travels through torch.distributed
    once; from then on the merge step is ONE C-ABI call, gysk_merge_global, with NCCL issued inside libgysketch.so"""
    world, rank = dist.get_world_size(), dist.get_rank()
    box = [eng.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    eng.nccl_comm_init(box[0], world, rank)
