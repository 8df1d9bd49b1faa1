set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "compact_gather or two_rank" 2>&1 | tail -5
bash tools/_run_nX_trace.sh 2
