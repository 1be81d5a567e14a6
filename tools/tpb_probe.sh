#!/bin/bash
# CTA-size probe of the limb kernels: 1, 2 or 4 phase-synchronised warps per CTA (ABR_LIMB_TPB = 32 / 64 / 128)
M=sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,smsp__inst_executed.sum,smsp__cycles_active.avg,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
for tpb in 32 64 128; do
  for cfg in "4096 400 barkour" "65536 100 barkour" "16384 200 biped"; do
    set -- $cfg
    echo "== ABR_LIMB_TPB=$tpb $3 $1 worlds x $2"
    ABR_LIMB_TPB=$tpb python tools/prof_rollout.py 0 $2 $1 3 $3 2>&1 | tail -1
  done
  echo "== ncu ABR_LIMB_TPB=$tpb barkour 4096x20"
  ABR_LIMB_TPB=$tpb ncu --metrics $M --clock-control none -k regex:k_limb_rollout -c 1 python tools/prof_rollout.py 0 20 4096 1 2>&1 | grep -E "icc|gcc|inst_executed|cycles_active|time_duration|no_instruction|barrier|issue_active"
done
