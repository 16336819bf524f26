set -x
python bench.py --steps 2 --warmup 1 > gpurun_out/b_r2c.json 2> gpurun_out/b_r2c.err || exit 1
B200SIM_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:k_post -s 60 -c 1 -o gpurun_out/prof_k_post_r2c -f python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_kp_r2c.log 2>&1
B200SIM_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 60 -c 1 -o gpurun_out/prof_k_solve_r2c -f python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_ks_r2c.log 2>&1
ls -la gpurun_out/*r2c*
