# peer-memory all-reduce of the statistics vector against NCCL inside bench.py's timed window (run under gpurun --gpus N)
N=${1:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $N --steps 20 --warmup 5 --no-configs --no-extra $3 2>gpurun_out/sx_$1.err | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$1', d['n_gpus'], 'ms/step', round(d['ms_per_step'], 5), 'value', round(d['value'] / 1e9, 2), 'G', d['per_rank_ms_per_step']['device_step'], (d.get('stats_exchange') or {}).get('how', '')[:40], (d.get('stats_exchange') or {}).get('why_not_peer_memory', ''))"; }
run peer 29521 ""
run nccl 29522 "--nccl-stats"
run peer2 29523 ""
run nccl2 29524 "--nccl-stats"
