set -x
mkdir -p gpurun_out
prof() { # name, extra ncu args
  w=$1; shift
  ncu --set full --clock-control none "$@" -f -o /tmp/prof_final_$w python profiles/profile_kernels.py --which $w > gpurun_out/prof_final_$w.log 2>&1
  python profiles/ncu_summary.py /tmp/prof_final_$w.ncu-rep > gpurun_out/final_${w}_ncu_summary.json 2>> gpurun_out/prof_final_$w.log
  rm -f /tmp/prof_final_$w.ncu-rep
}
prof k5 -k "regex:fdo_step_encode|encode_pi" -s 31 -c 3
prof k3 -k regex:fdo_determinize -s 1 -c 1
prof k4 -k regex:fdo_leaf_rollouts -s 1 -c 1
prof pimc -k regex:fdo_pimc -s 1 -c 1
prof sp -k "regex:sp_begin|sp_apply" -s 2 -c 2
du -sh gpurun_out
