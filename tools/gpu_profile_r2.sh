# round-2 profiling pass (one GPU): launch lists + ncu --set full captures of the dominant kernels.
# Each ncu command runs only after the same command has exited 0 without ncu.  Reports are summarised ON THE BOX
# (tools/ncu_summarize.py) and deleted: gpurun copies back at most 64 MiB.
set -x
mkdir -p gpurun_out/prof
H="python bench.py --steps 20 --warmup 5 --no-sweep --no-ddpm --no-train --no-cpu"
$H > gpurun_out/prof/bench_headline.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof/launches_headline.csv $H > gpurun_out/prof/ncu_headline.log 2>&1
cap () {  # name kernel-regex op B C H W dtype path title
  python tools/run_case.py $3 $4 $5 $6 $7 $8 $9 4 && \
  ncu --set full --clock-control none -k regex:$2 -s 2 -c 1 -o /tmp/prof_$1 -f \
      python tools/run_case.py $3 $4 $5 $6 $7 $8 $9 4 > gpurun_out/prof/ncu_$1.log 2>&1 && \
  python tools/ncu_summarize.py rep /tmp/prof_$1.ncu-rep gpurun_out/prof/r02_ncu_$1.md "r02 ncu --set full: $3 on [$4,$5,$6,$7] $8 (kernel regex $2)"
  rm -f /tmp/prof_$1.ncu-rep
}
cap fgelu3_sym_fwd fgelu3_tma fgelu_fwd 256 128 64 64 f32 auto
cap fgelu3_sym_bwd fgelu3_tma fgelu_bwd 256 128 64 64 f32 auto
cap fgelu3_sym_fwd_bf16 fgelu3_tma fgelu_fwd 256 128 64 64 bf16 auto
cap fgelu3_sym_bwd_bf16 fgelu3_tma fgelu_bwd 256 128 64 64 bf16 auto
cap fgelu3_plane4_fwd fgelu3_plane fgelu_fwd 4096 256 4 4 f32 auto
cap fgelu3_plane8_fwd fgelu3_plane fgelu_fwd 4096 128 8 8 f32 auto
cap up3_warp_4x4 up3_warp up 4096 256 4 4 f32 auto
cap down3_warp_8x8 down3_warp down 4096 128 8 8 f32 auto
cap down3_warp_16x16 down3_warp down 4096 64 16 16 f32 auto
cap gelu_down3 gelu_down3 gelu_down 256 64 64 64 f32 auto
python tools/train_step_case.py 32 3 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prof/launches_train32.csv python tools/train_step_case.py 32 3 > gpurun_out/prof/ncu_train32.log 2>&1
python tools/ncu_summarize.py list gpurun_out/prof/launches_train32.csv gpurun_out/prof/r02_ncu_launches_train_step_b32.md "r02 launch list: three eager Config-D training steps at batch 32 (the per-rank batch of the 8-GPU configs[1] run)"
python tools/ddpm_step_case.py 4096 2 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prof/launches_ddpm4096.csv python tools/ddpm_step_case.py 4096 2 > gpurun_out/prof/ncu_ddpm4096.log 2>&1
python tools/ncu_summarize.py list gpurun_out/prof/launches_ddpm4096.csv gpurun_out/prof/r02_ncu_launches_ddpm_step_b4096.md "r02 launch list: two eager Config-D reverse steps at batch 4096 (configs[2] on one GPU)"
python tools/ncu_summarize.py list gpurun_out/prof/launches_headline.csv gpurun_out/prof/r02_ncu_launches_headline.md "r02 launch list of the headline bench command (python bench.py --steps 20 --warmup 5 --no-sweep --no-ddpm --no-train --no-cpu)"
rm -f gpurun_out/prof/launches_train32.csv gpurun_out/prof/launches_ddpm4096.csv
du -sh gpurun_out/prof; ls -la gpurun_out/prof
