set -x
for w in C1 C5 C4; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r02bh_bench_$w.json 2> gpurun_out/r02bh_bench_$w.err; tail -2 gpurun_out/r02bh_bench_$w.err
done
python - <<'PY'
import json
for w in ["C1","C5","C4"]:
    d=json.loads(open(f"gpurun_out/r02bh_bench_{w}.json").read().strip().splitlines()[-1])
    print(w, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["e2e"]["steps"])
PY
