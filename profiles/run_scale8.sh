run() { n=$1; shift; if [ $n -eq 1 ]; then timeout 300 python bench.py --gpus 1 "$@"; else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n "$@"; fi; }
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tests/dist_fused_check.py > gpurun_out/dfc8.log 2>&1; grep -E "Error|assert|ok" gpurun_out/dfc8.log | cut -c1-200 | head -4
for n in 8 4 2 1; do run $n --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/yelp_n$n.log 2>&1; grep metric gpurun_out/yelp_n$n.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('yelp fused n=%d value=%.0f ms=%.4f e2e=%.0f'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value']))"; done
run 8 --steps 30 --warmup 5 --no-cpu-baseline --torch-adam > gpurun_out/yelp_n8_nccl.log 2>&1; grep metric gpurun_out/yelp_n8_nccl.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('yelp nccl+torch-adam n=%d value=%.0f ms=%.4f e2e=%.0f'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value']))"
run 8 --workload big --steps 20 --warmup 5 > gpurun_out/big8.log 2>&1; grep metric gpurun_out/big8.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('big n=%d value=%.0f ms=%.4f e2e=%.0f'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value']), d['kernels'])"
run 8 --workload yelp100 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/yelp100_n8.log 2>&1; grep metric gpurun_out/yelp100_n8.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('yelp100 n=%d value=%.0f ms=%.4f e2e=%.0f'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value']))"
