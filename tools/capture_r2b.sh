set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err || exit 1
B200SIM_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_l_r2b.log 2>&1
B200SIM_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:k_smooth -s 60 -c 1 -o gpurun_out/prof_k_smooth_r2b -f python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_ks_r2b.log 2>&1
ls -la gpurun_out/*r2b*
