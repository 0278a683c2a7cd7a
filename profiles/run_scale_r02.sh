# Round-2 multi-GPU runs on ONE box: bash profiles/run_scale_r02.sh N [all]   (N = 2, 4 or 8)
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29601 --steps 30 --warmup 5 > gpurun_out/r02_yelp_weak_n$N.json 2> gpurun_out/r02_yelp_weak_n$N.err
run 29602 --steps 30 --warmup 5 --workload yelp100 --scaling strong > gpurun_out/r02_yelp100_strong_n$N.json 2> gpurun_out/r02_yelp100_strong_n$N.err
run 29604 --steps 20 --warmup 5 --workload big > gpurun_out/r02_big_n$N.json 2> gpurun_out/r02_big_n$N.err
LIST="yelp_weak yelp100_strong big"
if [ "$2" = "all" ]; then
  run 29603 --steps 30 --warmup 5 --workload yelp100 > gpurun_out/r02_yelp100_weak_n$N.json 2> gpurun_out/r02_yelp100_weak_n$N.err
  LIST="$LIST yelp100_weak"
fi
for f in $LIST; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_${f}_n$N.json").read().strip().splitlines()[-1])
    print("$f n=$N", round(d["ms_per_step"], 4), "ms", round(d["value"] / 1e6, 2), "M/s", "e2e", round(d["e2e"]["value"] / 1e6, 2), d.get("dp_self_check", {}).get("replicas_bit_identical"))
except Exception as e:
    print("$f n=$N FAILED", e); print(open("gpurun_out/r02_${f}_n$N.err").read()[-1500:])
PY
done
