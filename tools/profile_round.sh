#!/bin/bash
# gpurun calls that capture the round's ncu evidence into gpurun_out/ (summarised into profiles/ by tools/summarize_profiles.py).
# gpurun brings back at most 64 MiB per call, so the captures are split in two parts.
# Usage (on the GPU box): tools/profile_round.sh r02 1|2
r=${1:-r02}
part=${2:-1}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
if [ "$part" = "1" ]; then
  $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/${r}_launches.csv \
      python tools/profile_step.py 2 > gpurun_out/${r}_launches.log 2>&1
  $NCU --set full --import-source on -k regex:attn_ --launch-skip 4 -c 2 -f -o gpurun_out/${r}_attn python tools/profile_attn.py > gpurun_out/${r}_attn.log 2>&1
  $NCU --set full --import-source on -k regex:ln_ --launch-skip 6 -c 3 -f -o gpurun_out/${r}_ln python tools/profile_ln.py > gpurun_out/${r}_ln.log 2>&1
else
  $NCU --set full -k regex:gemm_bf16 --launch-skip 8 -c 4 -f -o gpurun_out/${r}_fc1 python tools/profile_fc1.py > gpurun_out/${r}_fc1.log 2>&1
  $NCU --set full -k regex:gemm_bf16 --launch-skip 6 -c 3 -f -o gpurun_out/${r}_wgrad python tools/profile_wgrad.py > gpurun_out/${r}_wgrad.log 2>&1
  $NCU --set full -k regex:gemm_bf16 --launch-skip 6 -c 3 -f -o gpurun_out/${r}_head python tools/profile_head.py > gpurun_out/${r}_head.log 2>&1
fi
ls -la gpurun_out/${r}_*
