python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -m gpu -x -q 2>&1 | tail -5
echo "== new (prefetch + batch16)"
python tools/quick_bench.py f16 4096 256 1000
python tools/quick_bench.py f16 4096 256 1000
python tools/quick_bench.py f16 1024 4096 1000
python tools/quick_bench.py f16 4096 4096 100
echo "== batch16, no prefetch"
DPI_B200_LIB=$PWD/deeppicarditeration_b200/variants/lib_nopf.so python tools/quick_bench.py f16 4096 256 1000
DPI_B200_LIB=$PWD/deeppicarditeration_b200/variants/lib_nopf.so python tools/quick_bench.py f16 1024 4096 1000
