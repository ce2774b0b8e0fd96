"""world_size-2 gloo tests of the multi-rank host logic: flat-gradient all-reduce == gradient of the concatenated batch,
and SetCriterion's num_boxes normalisation across ranks (reference models/tubedetr.py:411-413)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tubedetr_b200.model import SetCriterion
    from tubedetr_b200.parallel import FlatGradBuffer
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    fb = FlatGradBuffer(net.parameters())
    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(2 * 6, 8, generator=g), torch.randn(2 * 6, 4, generator=g)
    xs, ys = x[rank * 6:(rank + 1) * 6], y[rank * 6:(rank + 1) * 6]
    fb.zero()
    ((net(xs) - ys) ** 2).mean().backward()
    assert all(p.grad.data_ptr() >= fb.flat.data_ptr() for p in net.parameters())  # grads were written in place
    flat = fb.all_reduce().clone()
    # criterion: rank 0 has 3 boxes, rank 1 has 5 -> num_boxes = 4 on both
    crit = SetCriterion(["boxes"])
    nb = 3 if rank == 0 else 5
    prep = crit.prepare([{"boxes": torch.rand(1, 4)} for _ in range(nb)])
    q.put((rank, flat, float(prep["num_boxes"])))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_and_num_boxes_gloo():
    port = 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(12, 8, generator=g), torch.randn(12, 4, generator=g)
    ((net(x) - y) ** 2).mean().backward()
    ref = torch.cat([p.grad.flatten() for p in net.parameters()])
    for rank, flat, nb in res:
        torch.testing.assert_close(flat, ref, atol=1e-6, rtol=1e-5)
        assert nb == 4.0
