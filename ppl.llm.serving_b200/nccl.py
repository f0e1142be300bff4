"""A raw ncclComm_t for the C ABI's tensor-parallel engines when the host is one process per GPU (torchrun).

The reference creates its communicators in ONE process (ppl::common::InitNccl -> one comm per GPU,
src/backends/cuda/resource_manager.cc:393; the C++ host side here does the same with ncclCommInitAll).  Under
``torch.distributed`` each rank is a process, so the unique id is made on rank 0, broadcast through the process
group (plumbing only) and every rank joins with ncclCommInitRank.  The library is the libnccl.so.2 already loaded
into the process by torch, which is also the one libb2llm.so resolves ``ncclAllReduce`` from.
"""
from __future__ import annotations

import ctypes as C

_NCCL_UNIQUE_ID_BYTES = 128
_lib = None


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * _NCCL_UNIQUE_ID_BYTES)]


def _nccl():
    global _lib
    if _lib is None:
        import torch  # noqa: F401  (loads the bundled libnccl.so.2 into the process)
        _lib = C.CDLL("libnccl.so.2", mode=C.RTLD_GLOBAL)
        _lib.ncclGetUniqueId.argtypes = [C.POINTER(_UniqueId)]
        _lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        _lib.ncclCommDestroy.argtypes = [C.c_void_p]
        _lib.ncclGetErrorString.restype = C.c_char_p
    return _lib


def create_comm(world_size: int, rank: int) -> C.c_void_p:
    """collective over the default torch.distributed group; the CUDA device must already be set"""
    import torch
    import torch.distributed as dist
    lib = _nccl()
    uid = _UniqueId()
    if rank == 0:
        rc = lib.ncclGetUniqueId(C.byref(uid))
        if rc != 0:
            raise RuntimeError(f"ncclGetUniqueId: {lib.ncclGetErrorString(rc).decode()}")
    # NB: reading a c_char array field stops at the first NUL -- copy the raw 128 bytes instead
    raw = C.string_at(C.byref(uid), _NCCL_UNIQUE_ID_BYTES)
    buf = torch.tensor(list(raw) if rank == 0 else [0] * _NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        buf = buf.cuda()
    dist.broadcast(buf, src=0)
    C.memmove(C.byref(uid), bytes(buf.cpu().tolist()), _NCCL_UNIQUE_ID_BYTES)
    comm = C.c_void_p()
    rc = lib.ncclCommInitRank(C.byref(comm), world_size, uid, rank)
    if rc != 0:
        raise RuntimeError(f"ncclCommInitRank: {lib.ncclGetErrorString(rc).decode()}")
    return comm


def destroy_comm(comm) -> None:
    if comm:
        _nccl().ncclCommDestroy(comm)
