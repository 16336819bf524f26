# round 2, final build: smoke, GPU tests, both bench arms, launch list and one full ncu capture of the dominant kernel
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2h.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r2h.log
python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2h.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests_r2h.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2h.json 2> gpurun_out/bench_ref_r2h.err
B200SIM_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2h.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l_r2h.log 2>&1
B200SIM_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 60 -c 1 -o gpurun_out/prof_k_solve_r2h -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ks_r2h.log 2>&1
tail -3 gpurun_out/smoke_r2h.log gpurun_out/gpu_tests_r2h.log
ls -la gpurun_out/*r2h*
