# Round-2 evidence, one gpurun call on 1 GPU: timelines, traces, bench lines of every workload.
make -C pc-gnn_b200/csrc trace > /dev/null 2>&1
for w in yelp amazon yelp100; do python profiles/anatomy.py $w > gpurun_out/r02_anatomy_$w.txt 2>&1; done
python profiles/big_step_kernels.py > gpurun_out/r02_anatomy_big.txt 2>&1
PCG_LIB_VARIANT=trace python profiles/trace_tile.py > gpurun_out/r02_tile_trace.txt 2>&1
PCG_LIB_VARIANT=trace python profiles/trace_choose.py > gpurun_out/r02_choose_trace.txt 2>&1
python profiles/replay_cost.py > gpurun_out/r02_replay_cost.txt 2>&1
for m in 0 2 10 6 15; do echo "PCG_PDL_MASK=$m"; PCG_PDL_MASK=$m python profiles/anatomy.py 2>&1 | grep "replay"; done > gpurun_out/r02_pdl_mask_sweep.txt
for w in yelp amazon yelp100 amazon_gcn big; do
  python bench.py --workload $w --steps 50 --warmup 5 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_bench_$w.json"))
    print("$w", round(d["ms_per_step"], 4), "ms", round(d["value"] / 1e6, 2), "M/s  e2e", round(d["e2e"]["value"] / 1e6, 2),
          "ref-loop", round(d.get("e2e_reference_loop", {}).get("value", 0) / 1e6, 2), "cpu", round(d["cpu_baseline"]["value"], 1),
          {k: round(v["ms"] * 1e3, 1) for k, v in d["kernels"].items()}, "roof", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("$w FAILED", e)
PY
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; tail -c 400 gpurun_out/r02_bench_reference_arm.json
