set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
for p in tma tma_general; do
python tools/run_case.py fgelu_fwd 256 128 64 64 f32 $p 8
python tools/run_case.py fgelu_bwd 256 128 64 64 f32 $p 8
python tools/run_case.py fgelu_fwd 256 128 64 64 bf16 $p 8
python tools/run_case.py fgelu_fwd_res 256 128 64 64 f32 $p 8
done
python tools/run_case.py fgelu_fwd 64 64 256 256 f32 tma 8
python tools/run_case.py fgelu_fwd 4096 32 32 32 f32 tma 8
python tools/run_case.py fgelu_fwd 4096 256 4 4 f32 direct 8
python tools/run_case.py fgelu_fwd 4096 256 4 4 f32 direct_general 8
