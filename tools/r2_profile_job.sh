# Round-2 measurement job (run on the B200 box through gpurun): the bench line, the launch list of the same command, and one
# full ncu capture per chain configuration plus the covariance kernel.  Outputs under gpurun_out/, summaries go to profiles/.
set -x
python bench.py > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err; tail -c 300 gpurun_out/bench_r2k.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches_cfg2.csv python bench.py --steps 3 --warmup 3 --only-primary --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:btk_chain_ws_kernel -s 4 -c 1 -o gpurun_out/r2d_chain_ws_cfg2 -f python bench.py --steps 3 --warmup 3 --only-primary --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:btk_chain_ws_kernel -s 4 -c 1 -o gpurun_out/r2d_chain_ws_cfg3 -f python tools/kbench.py cfg3 --steps 3 > gpurun_out/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:btk_chain_ws_kernel -s 4 -c 1 -o gpurun_out/r2d_chain_ws_cfg4 -f python tools/kbench.py cfg4 --steps 3 > gpurun_out/ncu_f4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:btk_covariance_tc_kernel -c 1 -o gpurun_out/r2_cov_tc_cfg4 -f python tools/staged_run.py cfg4 > gpurun_out/ncu_cov.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_staged.csv python tools/staged_run.py > gpurun_out/ncu_ls.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -6
