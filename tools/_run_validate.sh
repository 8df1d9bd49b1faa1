set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
python bench.py > gpurun_out/r02bj_bench_c4.json 2> gpurun_out/r02bj_bench_c4.err; tail -2 gpurun_out/r02bj_bench_c4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02bj_bench_c4.json").read().strip().splitlines()[-1])
print("C4", d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["parity_check"]["result"], d["gpu_launches"], d["clocks"])
PY
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
