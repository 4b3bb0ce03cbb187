python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tools/run_case.py down 256 128 64 64 bf16 auto 5
python tools/run_case.py down 1024 64 32 32 bf16 auto 5
python tools/run_case.py down 256 128 64 64 f32 auto 5
python tools/run_case.py up_bwd 256 128 64 64 bf16 auto 5
python tools/run_case.py up 256 128 64 64 bf16 auto 5
