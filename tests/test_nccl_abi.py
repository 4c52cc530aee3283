"""The torch-free data-parallel entry points of the C ABI (sefd_nccl_*): a single-rank communicator reduces a buffer in place
(world 1: the sum is the buffer itself); the multi-rank case is covered by tests/test_dist_gpu.py through torch.distributed."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_single_rank_allreduce_through_the_c_abi():
    from sefd import _lib
    from sefd.ops import ptr, stream
    lib = _lib.load()
    assert lib.sefd_nccl_unique_id_bytes() == 128
    uid = (C.c_char * 128)()
    torch.cuda.init()
    try:
        import torch.cuda.nccl as _nccl     # makes sure torch's bundled libnccl is in the process for dlopen
        _nccl.version()
    except Exception:
        pass
    rc = lib.sefd_nccl_unique_id(uid)
    if rc != 0:
        pytest.skip("libnccl.so.2 not loadable here: " + lib.sefd_last_error().decode())
    comm = lib.sefd_nccl_init(0, 1, uid)
    assert comm, lib.sefd_last_error().decode()
    x = torch.arange(1000, device="cuda", dtype=torch.float32)
    ref = x.clone()
    _lib.check(lib.sefd_nccl_allreduce(comm, ptr(x), x.numel(), stream()), "nccl_allreduce")
    torch.cuda.synchronize()
    assert torch.equal(x, ref) and lib.sefd_nccl_world(comm) == 1
    lib.sefd_nccl_destroy(comm)
