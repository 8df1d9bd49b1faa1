set -x
mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02bk_bench_c4_n$N.json 2> gpurun_out/r02bk_n$N.err; tail -3 gpurun_out/r02bk_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02bk_bench_c4_n$N.json").read().strip().splitlines()[-1])
print("N=$N", d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("stitch_check"), d.get("phases_ms"), d["config"]["parallelism"])
PY
