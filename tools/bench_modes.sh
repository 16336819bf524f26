# the non-headline bench lines of DESIGN.md section 6 (one B200)
set -x
python bench.py --task rough --steps 20 --warmup 5 --no-roofline-pass > gpurun_out/r02b_rough.json 2> gpurun_out/r02b_rough.err
python bench.py --task rough --model go1_rough --steps 20 --warmup 5 --no-roofline-pass > gpurun_out/r02b_rough_go1.json 2>> gpurun_out/r02b_rough.err
python bench.py --task tracking --steps 20 --warmup 5 --no-roofline-pass > gpurun_out/r02b_tracking.json 2> gpurun_out/r02b_tracking.err
python bench.py --mode physics --model go1_flat --steps 20 --warmup 5 --no-roofline-pass > gpurun_out/r02b_physics_go1.json 2> gpurun_out/r02b_physics.err
python bench.py --mode physics --steps 20 --warmup 5 --no-roofline-pass > gpurun_out/r02b_physics_g1.json 2>> gpurun_out/r02b_physics.err
python bench.py --model go1_flat --steps 20 --warmup 5 --no-roofline-pass > gpurun_out/r02b_go1_env.json 2>> gpurun_out/r02b_physics.err
python bench.py --task train --steps 5 --warmup 3 > gpurun_out/r02b_train_1gpu.json 2> gpurun_out/r02b_train.err
python bench.py --task train --steps 5 --warmup 3 --matmul-precision high > gpurun_out/r02b_train_1gpu_tf32.json 2>> gpurun_out/r02b_train.err
for f in gpurun_out/r02b_*.json; do python -c "
import sys,json
t=open('$f').read().strip().splitlines()
j=json.loads(t[-1]) if t else {}
print('$f', j.get('value'), j.get('ms_per_step'), (j.get('config') or {}).get('workload','')[:60])
"; done
tail -3 gpurun_out/r02b_*.err
