set -x
mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
SVO_SLAB_TRACE=1 timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02be_bench_c4_n$N.json 2> gpurun_out/r02be_n$N.err; grep "slab trace" gpurun_out/r02be_n$N.err | sort | uniq | head -40
python - <<PY
import json
d=json.loads(open("gpurun_out/r02be_bench_c4_n$N.json").read().strip().splitlines()[-1])
print("N=$N", d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("stitch_check"), d.get("phases_ms"))
PY
