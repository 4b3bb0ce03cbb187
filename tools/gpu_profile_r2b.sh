# round-2 follow-up pass: ncu --set full of the channels-last resamplers (fp32 and bf16)
set -x
mkdir -p gpurun_out/prof
cap () {  # name kernel-regex op B C H W dtype path
  AFR_CASE_CL=1 python tools/run_case.py $3 $4 $5 $6 $7 $8 $9 4 && \
  AFR_CASE_CL=1 ncu --set full --clock-control none -k regex:$2 -s 2 -c 1 -o /tmp/prof_$1 -f \
      python tools/run_case.py $3 $4 $5 $6 $7 $8 $9 4 > gpurun_out/prof/ncu_$1.log 2>&1 && \
  python tools/ncu_summarize.py rep /tmp/prof_$1.ncu-rep gpurun_out/prof/r02_ncu_$1.md "r02 ncu --set full: $3 on channels-last [$4,$5,$6,$7] $8 (kernel regex $2)"
  rm -f /tmp/prof_$1.ncu-rep
}
cap up3_nhwc up3_nhwc up 256 128 64 64 f32 auto
cap down3_nhwc down3_nhwc down 256 128 64 64 f32 auto
cap up3_nhwc_bf16 up3_nhwc up 256 128 64 64 bf16 auto
cap down3_nhwc_bf16 down3_nhwc down 256 128 64 64 bf16 auto
