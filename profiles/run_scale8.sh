run() { n=$1; shift; if [ $n -eq 1 ]; then timeout 300 python bench.py --gpus 1 "$@"; else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n "$@"; fi; }
run 8 --workload big --steps 20 --warmup 5 > gpurun_out/big8.log 2>&1; grep metric gpurun_out/big8.log | cut -c1-700
for n in 8 4 2 1; do run $n --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/yelp_n$n.log 2>&1; grep metric gpurun_out/yelp_n$n.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('yelp n=%d value=%.0f ms=%.4f e2e=%.0f'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value']))"; done
run 8 --workload yelp100 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/yelp100_n8.log 2>&1; grep metric gpurun_out/yelp100_n8.log | cut -c1-260
