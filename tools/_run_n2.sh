set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "compact_gather or two_rank" 2>&1 | tail -15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02ba_bench_c4_n2.json 2> gpurun_out/r02ba_n2.err; tail -3 gpurun_out/r02ba_n2.err
SVO_COMPACT=0 timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-stitch-check --no-cpu-baseline > gpurun_out/r02ba_bench_c4_n2_full.json 2> gpurun_out/r02ba_n2_full.err
python - <<'PY'
import json
for f in ["r02ba_bench_c4_n2.json","r02ba_bench_c4_n2_full.json"]:
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("stitch_check"), d.get("phases_ms"))
    except Exception as e: print(f, "ERR", e)
PY
