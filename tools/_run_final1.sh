set -x
mkdir -p gpurun_out
python tools/tune.py --workload C4 --reps 8 --variants "base;git:9824c59" 2>&1 | tee gpurun_out/tune41.log | cut -c1-200
python bench.py > gpurun_out/r02bb_bench_c4.json 2> gpurun_out/r02bb_bench_c4.err; tail -2 gpurun_out/r02bb_bench_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02bb_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02bb_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_brick_emit|k_brick_raster|k_brick_ranks|k_brick_flat|k_brick_pairs|k_brick_heads' -c 8 -o gpurun_out/r02bb_brick python tools/profile_step.py C4 > gpurun_out/r02bb_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
