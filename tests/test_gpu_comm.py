"""The single-process form of the path's one collective (include/zuko_b200.h: zk_comm_*, zk_allreduce_sum):
all-reduce(sum) of {sum log p, count} over NCCL.  Needs >= 2 GPUs in one process (skipped on a 1-GPU box;
the one-process-per-GPU form is covered by tests/test_distributed.py and bench.py under torchrun)."""

import ctypes

import pytest
import torch

from zuko_b200 import _engine as E

pytestmark = pytest.mark.gpu


def test_allreduce_sum_of_nll_terms_across_devices():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs in this process")
    comm = ctypes.c_void_p()
    E.check(E.lib().zk_comm_init_all(n, ctypes.byref(comm)))
    try:
        assert E.lib().zk_comm_size(comm) == n
        bufs = [torch.tensor([-(d + 1.0) * 1000.0, 256.0 + d], dtype=torch.float64, device=f"cuda:{d}") for d in range(n)]
        E.check(E.lib().zk_comm_group_begin(comm))
        for d in range(n):
            with torch.cuda.device(d):
                E.check(E.lib().zk_allreduce_sum(comm, d, bufs[d].data_ptr(), 2, ctypes.c_void_p(torch.cuda.current_stream(d).cuda_stream)))
        E.check(E.lib().zk_comm_group_end(comm))
        for d in range(n):
            torch.cuda.synchronize(d)
        want = [-1000.0 * n * (n + 1) / 2, 256.0 * n + n * (n - 1) / 2]
        for d in range(n):
            assert bufs[d].tolist() == want
        mean_nll = -bufs[0][0].item() / bufs[0][1].item()
        assert mean_nll > 0
    finally:
        E.lib().zk_comm_destroy(comm)
