python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_cpp_mirror.py tests/test_gpu_prove.py -m gpu -x -q 2>&1 | tail -2
