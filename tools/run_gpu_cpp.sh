set -x
timeout 900 python -m pytest tests/test_cpp_mirror.py tests/test_wire_formats.py -m gpu -x -q 2>&1 | tail -30
