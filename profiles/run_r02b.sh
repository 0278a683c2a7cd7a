# Round-2 (second session) check, one gpurun call on 1 GPU: GPU tests, default bench line, step timeline.
# usage: bash profiles/run_r02b.sh [tag]
tag=${1:-r02b}
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/${tag}_gputests.txt
tail -3 gpurun_out/${tag}_gputests.txt
python bench.py --steps 50 --warmup 5 > gpurun_out/${tag}_bench_yelp.json 2> gpurun_out/${tag}_bench_yelp.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_yelp.json"))
    print("yelp", round(d["ms_per_step"], 4), "ms", round(d["value"] / 1e6, 2), "M/s  e2e", round(d["e2e"]["value"] / 1e6, 2), round(d["e2e"]["ms_per_step"], 4),
          "ref-loop", round(d.get("e2e_reference_loop", {}).get("value", 0) / 1e6, 2),
          {k: round(v["ms"] * 1e3, 1) for k, v in d["kernels"].items()}, "roof", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("yelp FAILED", e)
PY
python profiles/anatomy.py yelp > gpurun_out/${tag}_anatomy_yelp.txt 2>&1
grep -A12 "replay 2" gpurun_out/${tag}_anatomy_yelp.txt
grep -A14 "host-batch" gpurun_out/${tag}_anatomy_yelp.txt
if [ -d baseline/_ref/reference/src ]; then
  # one-off: the reference's own callers on the CUDA modules (needs the reference tree next to a GPU)
  PCGNN_REFERENCE_ROOT=$PWD/baseline/_ref/reference timeout 300 python -m pytest tests/test_gpu_reference_callers.py -m gpu -q -rA 2>&1 | tail -40 > gpurun_out/${tag}_reference_callers.txt
  tail -15 gpurun_out/${tag}_reference_callers.txt
fi
if [ -f pc-gnn_b200/libpcgnn_b200_trace.so ]; then
  PCG_LIB_VARIANT=trace python profiles/trace_choose.py > gpurun_out/${tag}_choose_trace.txt 2>&1
  grep -A8 "wide tier" gpurun_out/${tag}_choose_trace.txt | tail -30
fi
