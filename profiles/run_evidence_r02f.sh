# Final round-2 evidence, one gpurun call on 1 GPU: bash profiles/run_evidence_r02f.sh
T=r02f
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${T}_gputests.txt; tail -2 gpurun_out/${T}_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in yelp yelp100; do python profiles/anatomy.py $w > gpurun_out/${T}_anatomy_$w.txt 2>&1; done
python profiles/big_step_kernels.py > gpurun_out/${T}_anatomy_big.txt 2>&1
for w in yelp amazon yelp100 amazon_gcn big; do
  python bench.py --workload $w --steps 50 --warmup 5 > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_$w.json"))
    print("$w", round(d["ms_per_step"], 4), "ms", round(d["value"] / 1e6, 2), "M/s  e2e", round(d["e2e"]["value"] / 1e6, 2),
          "ref-loop", round(d.get("e2e_reference_loop", {}).get("value", 0) / 1e6, 2), "fast", round(d.get("e2e_reference_loop_fast", {}).get("value", 0) / 1e6, 2),
          "cpu", round(d["cpu_baseline"]["value"], 1),
          {k: round(v["ms"] * 1e3, 1) for k, v in d["kernels"].items()}, "roof", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("$w FAILED", e)
PY
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2>/dev/null; tail -c 300 gpurun_out/${T}_bench_reference_arm.json
python profiles/replay_cost.py > gpurun_out/${T}_replay_cost.txt 2>&1
# ncu: launch list of the bench command, then one full capture of the C2 step (each after the same command ran plain)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
rm -f gpurun_out/${T}_traffic.json
python profiles/prof_step.py yelp > gpurun_out/${T}_plain_step_yelp.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/${T}_step_yelp -f \
    python profiles/prof_step.py yelp > gpurun_out/${T}_ncu_step_yelp.log 2>&1
python profiles/ncu_step_report.py gpurun_out/${T}_step_yelp.ncu-rep gpurun_out/${T}_traffic.json yelp > gpurun_out/${T}_ncu_full_step_yelp.txt
cat gpurun_out/${T}_ncu_full_step_yelp.txt
