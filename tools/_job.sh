date +%s > /tmp/t0
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench2_err.log | tail -1 > gpurun_out/bench_n2_overlap.json
echo "exit code $? after $(( $(date +%s) - $(cat /tmp/t0) )) s"
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_overlap.json')); print('N=2 ms', d['ms_per_step'], 'value', d['value'], d['config']['allreduce'], 'graph', d['config']['cuda_graph'], 'e2e', d['e2e']['value'])"
grep -v "Warn\|warn" gpurun_out/bench2_err.log | tail -5
