set -x
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_r1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1.log 2>&1
tail -2 gpurun_out/ncu_r1.log | cut -c1-300
python profiles/prof_kernels.py > gpurun_out/plain_pk.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_choose_small|k_choose_wide|k_aggregate|k_choose_prep" -s 10 -c 10 -o gpurun_out/prof_r1_final python profiles/prof_kernels.py > gpurun_out/ncu_pk.log 2>&1
tail -3 gpurun_out/ncu_pk.log
ls -la gpurun_out/*.ncu-rep
