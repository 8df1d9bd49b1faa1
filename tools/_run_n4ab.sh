set -x
mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
for D in 1 0; do
SVO_DEFER_REST=$D SVO_SLAB_TRACE=1 timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-stitch-check > gpurun_out/r02bi_c4_n${N}_defer$D.json 2> gpurun_out/r02bi_n${N}_defer$D.err
grep "slab trace" gpurun_out/r02bi_n${N}_defer$D.err | sed 's/\[slab trace\]/\n[slab trace]/g' | grep "rank 0" | sort -u | head -2
python - <<PY
import json
d=json.loads(open("gpurun_out/r02bi_c4_n${N}_defer$D.json").read().strip().splitlines()[-1])
print("N=$N defer=$D", d["ms_per_step"], d["e2e"]["ms_per_step"])
PY
done
