"""2-rank check of the overlapped gradient all-reduce (run under torchrun on 2 GPUs):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dp_check.py
eval-mode step (deterministic kernels), three ways: (a) plain backward + pack + ONE all-reduce, (b) backward_overlapped
eagerly, (c) backward_overlapped captured in a CUDA graph (NCCL inside the capture) and replayed.  The flat gradient
buffers must agree; (a) is also checked against the mean of the two ranks' local gradients."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
model, crit, wd = bench.build_everything(dev)
model.eval()
os.environ["TDB_OVERLAP"] = "0"
st = bench.Step(model, crit, wd, dev, rank, world, use_graph=False)
st.body()
local_flat = st.flat.clone()
st.fgb.all_reduce()
torch.cuda.synchronize()
ref = st.flat.clone()
gathered = [torch.empty_like(local_flat) for _ in range(world)]
dist.all_gather(gathered, local_flat)
mean = sum(gathered) / world


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


r0 = rel(ref, mean)
st.overlap = True
if os.environ.get("TDB_STAGED_AR", "1") != "0":
    st.setup_staged()       # per-stage all-reduce on a second communicator, gradients written in place (the bench default at N > 1)
st.body()
torch.cuda.synchronize()
eager = st.flat.clone()
r1 = rel(eager, ref)
if rank == 0 and r1 >= 1e-4:        # which parameters differ
    names = {id(p): n for n, p in model.named_parameters()}
    worst = []
    for p, off in zip(st.fgb.params, st.fgb.offsets):
        a, b = eager[off:off + p.numel()], ref[off:off + p.numel()]
        worst.append((((a - b).norm() / (b.norm() + 1e-30)).item(), names.get(id(p), "?"), b.norm().item()))
    worst.sort(reverse=True)
    for w in worst[:8]:
        print(f"  mismatch {w[1]}: rel {w[0]:.3e} (ref norm {w[2]:.3e})")
ok_graph, r2 = True, float("nan")
try:
    st.use_graph = True
    st._capture()
    st.flat.zero_()
    st.run()
    st.run()
    torch.cuda.synchronize()
    r2 = rel(st.flat, ref)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    e0.record()
    for _ in range(10):
        st.run()
    e1.record()
    torch.cuda.synchronize()
    ms_overlap = e0.elapsed_time(e1) / 10
except Exception as e:
    ok_graph = False
    ms_overlap = float("nan")
    print(f"[rank {rank}] graph capture with NCCL failed: {type(e).__name__}: {e}", flush=True)
if rank == 0:
    print(f"allreduce vs mean of local grads: rel {r0:.3e}")
    print(f"overlapped eager vs serialised:   rel {r1:.3e}")
    print(f"overlapped in-graph vs serialised: rel {r2:.3e} (graph ok: {ok_graph}, graph object: {st.graph is not None}, staged: {st.comm_stream is not None}), "
          f"{ms_overlap:.3f} ms/step eval mode")
    good = r0 < 1e-5 and r1 < 1e-4 and (not ok_graph or r2 < 1e-4)
    print("dp_check ok" if good else f"dp_check FAILED: {(r0, r1, r2)}")
sys.stdout.flush()
st.graph = None
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
