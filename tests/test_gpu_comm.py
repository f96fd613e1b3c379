"""ds_comm_* / ds_allgather (the sharded job's one collective) on the library's run-time bound NCCL: world size 1 on any
box, world size 2 when two GPUs are visible (one process per GPU, id broadcast over gloo)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_allgather_world1():
    from diffusynth_b200.engine import Comm
    c = Comm(0, 1)
    x = torch.arange(1000, dtype=torch.float32, device="cuda").reshape(10, 100)
    y = c.all_gather(x)
    torch.cuda.synchronize()
    assert torch.equal(x, y)
    i = torch.arange(77, dtype=torch.int64, device="cuda")
    assert torch.equal(c.all_gather(i), i)


def _worker(rank, world, port):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffusynth_b200.engine import Comm

    def bootstrap(buf):
        dist.broadcast(buf, src=0)
        return buf

    c = Comm(rank, world, bootstrap)
    x = torch.full((64, 65280), float(rank + 1), device="cuda")
    x[:, 0] = torch.arange(64, device="cuda") + 1000 * rank
    y = c.all_gather(x)
    torch.cuda.synchronize()
    for r in range(world):
        blk = y[r * 64:(r + 1) * 64]
        assert float(blk[0, 1]) == r + 1 and float(blk[5, 0]) == 5 + 1000 * r, (rank, r)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_allgather_world2():
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, 29533), nprocs=2, join=True)
