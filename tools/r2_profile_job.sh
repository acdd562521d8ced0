set -x
python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 600 gpurun_out/bench_r2c.json
# launch list of the default bench command (primary only keeps it short)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2.csv python bench.py --steps 3 --warmup 3 --only-primary --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
# full capture of the chain kernel at cfg2
ncu --set full --clock-control none --import-source on -k regex:btk_chain_ws_kernel -s 4 -c 1 -o gpurun_out/r2_chain_ws_cfg2 -f python bench.py --steps 3 --warmup 3 --only-primary --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
# cfg3 and cfg4 (kbench shapes: 64 x 10 s x 16 ch; 32 x 10 s x 64 ch)
ls -la gpurun_out/*.ncu-rep | tail -5
