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
    ref = torch.cat([torch.cat([p.grad.flatten(), torch.zeros(-p.numel() % 8)]) for p in net.parameters()])
    for rank, flat, nb in res:
        torch.testing.assert_close(flat, ref, atol=1e-6, rtol=1e-5)
        assert nb == 4.0


class _TwoTrunk(torch.nn.Module):
    """toy with the step's shape: a text trunk and a backbone trunk feeding a shared head (names carry the LR-group substrings
    of reference main.py:381-405)"""

    def __init__(self):
        super().__init__()
        self.text_encoder = torch.nn.Sequential(torch.nn.Linear(6, 12), torch.nn.Tanh(), torch.nn.Linear(12, 8))
        self.backbone = torch.nn.Sequential(torch.nn.Linear(5, 8), torch.nn.ReLU(), torch.nn.Linear(8, 8))
        self.head = torch.nn.Linear(8, 3)
        self.unused = torch.nn.Linear(2, 2)          # like RoBERTa's pooler: never reached by the loss -> zero gradient

    def forward(self, xt, xb):
        self.cut = (self.text_encoder(xt), self.backbone(xb))
        return self.head(self.cut[0] * self.cut[1] + self.cut[1])


def _worker_overlap(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tubedetr_b200.parallel import (GROUP_BACKBONE, GROUP_REST, GROUP_TEXT, FlatGradBuffer, backward_overlapped,
                                        default_group_of)
    torch.manual_seed(0)
    net = _TwoTrunk()
    fb = FlatGradBuffer(net.named_parameters(), groups=default_group_of)
    # group-major layout: text | rest | backbone, each one contiguous slice
    assert fb.group_ids == sorted(fb.group_ids)
    assert fb.segment(GROUP_TEXT).numel() == sum((p.numel() + 7) // 8 * 8 for p in net.text_encoder.parameters())
    assert all(o % 8 == 0 for o in fb.offsets)
    assert fb.segment(GROUP_REST, GROUP_BACKBONE).data_ptr() == fb.segment(GROUP_REST).data_ptr()
    g = torch.Generator().manual_seed(7)
    xt, xb, y = torch.randn(8, 6, generator=g), torch.randn(8, 5, generator=g), torch.randn(8, 3, generator=g)
    sl = slice(rank * 4, rank * 4 + 4)
    for p in fb.params:
        p.grad = None
    loss = ((net(xt[sl], xb[sl]) - y[sl]) ** 2).mean()
    flat = backward_overlapped(loss, fb, net.cut[0], net.cut[1]).clone()
    assert all(p.grad.data_ptr() >= fb.flat.data_ptr() for p in fb.params)
    # frozen text trunk: its slice must come out zero and nothing may break
    for p in net.text_encoder.parameters():
        p.requires_grad_(False)
    fb2 = FlatGradBuffer(net.named_parameters(), groups=default_group_of)
    loss = ((net(xt[sl], xb[sl]) - y[sl]) ** 2).mean()
    flat2 = backward_overlapped(loss, fb2, net.cut[0], net.cut[1]).clone()
    q.put((rank, flat, flat2, [n for n, p in net.named_parameters() if p.requires_grad]))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_backward_equals_plain_backward_gloo():
    port = 31500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_overlap, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from tubedetr_b200.parallel import default_group_of
    torch.manual_seed(0)
    net = _TwoTrunk()
    g = torch.Generator().manual_seed(7)
    xt, xb, y = torch.randn(8, 6, generator=g), torch.randn(8, 5, generator=g), torch.randn(8, 3, generator=g)
    ((net(xt, xb) - y) ** 2).mean().backward()
    named = sorted(enumerate(net.named_parameters()), key=lambda t: (default_group_of(t[1][0]), t[0]))
    def padded(p):
        g = (p.grad if p.grad is not None else torch.zeros_like(p)).flatten()
        return torch.cat([g, torch.zeros(-g.numel() % 8)])
    ref = torch.cat([padded(p) for _, (n, p) in named])
    ref2 = torch.cat([padded(p) for _, (n, p) in named if "text_encoder" not in n])
    for rank, flat, flat2, _ in res:
        torch.testing.assert_close(flat, ref, atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(flat2, ref2, atol=1e-6, rtol=1e-5)
