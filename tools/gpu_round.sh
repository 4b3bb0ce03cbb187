# full GPU pass: tests, default bench line, launch list of the headline command, sweeps, ncu capture of the headline kernel
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
python bench.py --steps 20 --warmup 5 --no-sweep --no-ddpm --no-train --no-cpu > gpurun_out/bench_headline.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_headline.csv \
    python bench.py --steps 20 --warmup 5 --no-sweep --no-ddpm --no-train --no-cpu > gpurun_out/ncu_headline.log 2>&1
python bench.py --steps 20 --warmup 5 --grid-sweep --no-ddpm --no-train --no-cpu > gpurun_out/bench_grid.json 2> gpurun_out/bench_grid.err
python bench.py --steps 20 --warmup 5 --full-sweep --no-ddpm --no-train --no-cpu > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python tools/run_case.py fgelu_fwd 256 128 64 64 f32 tma 4 && \
ncu --set full --clock-control none --import-source on -k regex:fgelu3_tma -s 2 -c 1 -o gpurun_out/prof_sym_fwd -f \
    python tools/run_case.py fgelu_fwd 256 128 64 64 f32 tma 4 > gpurun_out/ncu_sym_fwd.log 2>&1
python tools/run_case.py fgelu_bwd 256 128 64 64 f32 tma 4 && \
ncu --set full --clock-control none --import-source on -k regex:fgelu3_tma -s 2 -c 1 -o gpurun_out/prof_sym_bwd -f \
    python tools/run_case.py fgelu_bwd 256 128 64 64 f32 tma 4 > gpurun_out/ncu_sym_bwd.log 2>&1
