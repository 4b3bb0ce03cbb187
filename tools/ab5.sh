python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -6
for n in 6 4; do
AFR_CASE_N=$n python tools/run_case.py up 256 128 64 64 f32 auto 4
AFR_CASE_N=$n python tools/run_case.py down 256 128 64 64 f32 auto 4
AFR_CASE_N=$n python tools/run_case.py up_bwd 256 128 64 64 f32 auto 4
AFR_CASE_N=$n python tools/run_case.py down_bwd 256 128 64 64 f32 auto 4
done
AFR_CASE_N=6 python tools/run_case.py up 256 128 64 64 bf16 auto 4
AFR_CASE_N=6 python tools/run_case.py down 256 128 64 64 bf16 auto 4
