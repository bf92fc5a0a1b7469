# One GPU call: GPU tests, the C2 bench line, the ncu launch list and the --set full capture of the three hot kernels.
set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r02c_gpu_tests.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_bench_C2.json 2> gpurun_out/r02c_bench_C2.err
python bench.py --steps 1 --warmup 1 --niter 4 --no-cpu-baseline > gpurun_out/r02c_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02c_launches.csv \
    python bench.py --steps 1 --warmup 1 --niter 4 --no-cpu-baseline > gpurun_out/r02c_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:k_estep_tc|k_w_v2|k_cols_v1' -c 3 -f -o gpurun_out/r02c_full \
    python bench.py --steps 1 --warmup 1 --niter 4 --no-cpu-baseline > gpurun_out/r02c_ncu_f.log 2>&1
cat gpurun_out/r02c_gpu_tests.txt
