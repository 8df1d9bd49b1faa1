set -x
mkdir -p gpurun_out
for w in C1 C2 C3 C5; do
  timeout 600 python bench.py --workload $w > gpurun_out/r02bf_bench_$w.json 2> gpurun_out/r02bf_bench_$w.err; tail -2 gpurun_out/r02bf_bench_$w.err
done
python - <<'PY'
import json
for w in ["C1","C2","C3","C5"]:
    try:
        d=json.loads(open(f"gpurun_out/r02bf_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["build_path"], (d.get("parity_check") or {}).get("result"), d["config"]["fragments"], d["config"]["leaf_voxels"], d.get("phases_ms"), (d.get("roofline") or {}).get("frac"))
    except Exception as e: print(w, "ERR", e)
PY
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
