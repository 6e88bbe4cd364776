set -x
timeout 900 python -m pytest tests/test_wire_formats.py tests/test_gpu_prove.py -m gpu -x -q 2>&1 | tail -25
