# round-2 last pass: ncu --set full of the fp32 up2x kernel after the store fix
set -x
mkdir -p gpurun_out/prof
python tools/run_case.py up 256 128 64 64 f32 auto 4 && \
ncu --set full --clock-control none -k regex:up3_flat -s 2 -c 1 -o /tmp/prof_up3_flat -f python tools/run_case.py up 256 128 64 64 f32 auto 4 > gpurun_out/prof/ncu_up3_flat.log 2>&1 && \
python tools/ncu_summarize.py rep /tmp/prof_up3_flat.ncu-rep gpurun_out/prof/r02_ncu_up3_flat_f32.md "r02 ncu --set full: up2x on [256,128,64,64] f32 (up3_flat_kernel, two input columns per thread)"
rm -f /tmp/prof_up3_flat.ncu-rep
