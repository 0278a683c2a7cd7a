# Round-2 captures (one gpurun call, 1 GPU). Every ncu run follows a plain run of the same command. The reports are
# reduced to text on the box (profiles/ncu_step_report.py): gpurun brings back at most 64 MiB.
set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
tail -c 200 gpurun_out/r02_ncu_bench.log
rm -f gpurun_out/r02_traffic.json
for wl in yelp yelp100; do
python profiles/prof_step.py $wl > gpurun_out/r02_plain_step_$wl.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_step_$wl -f \
    python profiles/prof_step.py $wl > gpurun_out/r02_ncu_step_$wl.log 2>&1
tail -1 gpurun_out/r02_ncu_step_$wl.log
python profiles/ncu_step_report.py gpurun_out/r02_step_$wl.ncu-rep gpurun_out/r02_traffic.json $wl > gpurun_out/r02_ncu_full_step_$wl.txt
done
PCG_NCU=1 python profiles/big_step_kernels.py > gpurun_out/r02_plain_step_big.log 2>&1 && \
PCG_NCU=1 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/r02_step_big -f \
    python profiles/big_step_kernels.py > gpurun_out/r02_ncu_step_big.log 2>&1
tail -1 gpurun_out/r02_ncu_step_big.log
python profiles/ncu_step_report.py gpurun_out/r02_step_big.ncu-rep gpurun_out/r02_traffic.json big > gpurun_out/r02_ncu_full_step_big.txt
rm -f gpurun_out/r02_step_yelp100.ncu-rep gpurun_out/r02_step_big.ncu-rep      # keep the C2 report (22 MB) only
ls -la gpurun_out/
