"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

* Frames / cloud pairs / BA problems / sequences are independent: `shard_range` splits them across
  ranks with no data-path collective.
* One LocalInertialBA problem can also be partitioned by landmark: rank r owns the landmarks p with
  p % world == r (`landmark_owner`), every rank keeps all keyframe states, and the Schur-reduced pose
  system is summed over ranks once per LM trial (`ba_allreduce_callback`, handed to
  gfs_ba_set_partition).  The reference has no counterpart; the correctness statement is "the sum of
  the per-shard reduced systems equals the single-shard system".
"""
import ctypes as C


def shard_range(n, rank, world):
    """Contiguous, balanced [start, stop) of n items for `rank`."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def landmark_owner(p, world):
    return p % world


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)


class _DevArray:
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def ba_allreduce_callback(group=None):
    """ctypes callback for gfs_ba_set_partition: in-place SUM all-reduce of a device fp64 buffer over
    NCCL via torch.distributed (the buffer is wrapped through the CUDA array interface, no copy)."""
    import torch
    import torch.distributed as dist

    def _cb(ptr, count, _user):
        try:
            t = torch.as_tensor(_DevArray(ptr, count), device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            torch.cuda.current_stream().synchronize()
            return 0
        except Exception as e:  # never let an exception cross the C boundary
            print("all-reduce callback failed:", e)
            return 1

    return ALLREDUCE_FN(_cb)
