# Round-2 captures (one gpurun call, 1 GPU). Every ncu run follows a plain run of the same command.
set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
tail -c 300 gpurun_out/r02_ncu_bench.log
for wl in yelp yelp100; do
python profiles/prof_step.py $wl > gpurun_out/r02_plain_step_$wl.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_step_$wl -f \
    python profiles/prof_step.py $wl > gpurun_out/r02_ncu_step_$wl.log 2>&1
tail -2 gpurun_out/r02_ncu_step_$wl.log
done
ls -la gpurun_out/*.ncu-rep
