set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "occupancy or other_networks or at_size or other_value" 2>&1 | tail -5
timeout 600 python scripts/other_nets_timing.py 8192 5 2>&1 | tail -4 | tee gpurun_out/r02s_other_nets_timing.txt
