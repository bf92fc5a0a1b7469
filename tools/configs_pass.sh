# bench lines of the other BASELINE configurations on one B200 (C2 is the default run)
for c in C1 C3 C4; do python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_$c.json 2> gpurun_out/r02c_bench_$c.err; done
python bench.py --config C5 --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/r02c_bench_C5.json 2> gpurun_out/r02c_bench_C5.err
python bench.py --config C5 --waves 4 --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/r02c_bench_C5_w4.json 2> gpurun_out/r02c_bench_C5_w4.err
for f in gpurun_out/r02c_bench_C*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['frac'],4), round(d['roofline_nmf']['frac'],4))"; done
