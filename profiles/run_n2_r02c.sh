# 2-GPU sanity of the host-batch recording (bench dp self-check + e2e) on C2 weak and the row-partitioned C5
N=2
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29601 --steps 30 --warmup 5 > gpurun_out/r02c_yelp_weak_n$N.json 2> gpurun_out/r02c_yelp_weak_n$N.err
run 29604 --steps 20 --warmup 5 --workload big > gpurun_out/r02c_big_n$N.json 2> gpurun_out/r02c_big_n$N.err
for f in yelp_weak big; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02c_${f}_n$N.json").read().strip().splitlines()[-1])
    print("$f n=$N", round(d["ms_per_step"], 4), "ms", round(d["value"] / 1e6, 2), "M/s", "e2e", round(d["e2e"]["value"] / 1e6, 2), d.get("dp_self_check"))
except Exception as e:
    print("$f n=$N FAILED", e); print(open("gpurun_out/r02c_${f}_n$N.err").read()[-1500:])
PY
done
