# multi-GPU evidence run (one box, N = 8/4/2): weak scaling of config 2, and BASELINE config[3]
# (10 000 servers over 8 GPUs, capacity caps, gather + replicated greedy)
run() { # n out args...
  n=$1; shift; out=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err || tail -c 400 gpurun_out/$out.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/$out.json') if l.startswith('{')][-1]); print('$out', d['n_gpus'], '%.4e'%d['value'], '%.3f ms'%d['ms_per_step'], {k:round(v,3) for k,v in d.get('phases_ms',{}).items()}, 'e2e %.3e'%d['e2e']['value'])"
}
run 8 bench_r01_n8 --steps 10 --warmup 3 --no-cpu-baseline
run 4 bench_r01_n4 --steps 10 --warmup 3 --no-cpu-baseline
run 2 bench_r01_n2 --steps 10 --warmup 3 --no-cpu-baseline
run 8 bench_r01_n8_cfg4_limited --config 4 --servers-per-rank 1250 --limited --steps 3 --warmup 3 --no-cpu-baseline
timeout 200 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -1
