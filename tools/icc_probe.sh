#!/bin/bash
# I-cache probe of the limb rollout kernel: the same launch with parts of the step switched off (smaller executed
# code footprint per horizon iteration). usage: tools/icc_probe.sh > gpurun_out/icc_probe.txt
M=sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gcc__average_cache_request_hit_rate.pct,smsp__inst_executed.sum,smsp__cycles_active.avg,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
for v in "0 0" "256 0" "1 0" "0 1" "0 2" "16384 0"; do
  set -- $v
  echo "== ABR_PROF_DISABLE=$1 ABR_PROF_LS=$2 (worlds ${W:-4096})"
  ABR_PROF_DISABLE=$1 ABR_PROF_LS=${2/#0/} ncu --metrics $M --clock-control none -k regex:k_limb_rollout -c 1 python tools/prof_rollout.py 0 20 ${W:-4096} 1 2>&1 | grep -E "icc|gcc|inst_executed|cycles_active|time_duration|no_instruction|issue_active|world-steps"
done
