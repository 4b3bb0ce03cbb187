python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or oracle_fp32 or bf16 or fuzz" 2>&1 | tail -2
for dt in bf16 f32; do
python tools/run_case.py down 256 128 64 64 $dt auto 8
python tools/run_case.py down 1024 64 32 32 $dt auto 8
python tools/run_case.py down 16 64 256 256 $dt auto 8
python tools/run_case.py up_bwd 256 128 64 64 $dt auto 8
done
