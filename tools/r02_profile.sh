#!/bin/bash
# Round-2 ncu evidence (run under gpurun on ONE B200): launch list of one complete wrap + one `--set full` capture per kernel.
# The raw metric pages are exported to CSV on the box (gpurun_out/ may carry 64 MiB back); only one report is kept whole.
# usage: tools/r02_profile.sh [prefix]   (default r02f = the final state of round 2)
P=${1:-r02f}
NCU="ncu --profile-from-start off --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${P}_launches.csv python tools/profile_one_wrap.py > gpurun_out/${P}_launches.log 2>&1
# name:kernel regex:launches of that kernel to skip inside the profiled wrap (accumulate<Fp> #7 = the Z MSM, 8.4 M full-width scalars)
for spec in "tape_staged:k_tape_staged:0" "tape_poseidon4:k_tape_poseidon4:0" "digits1:k_msm_digits<.*1>:0" "window_partial:k_msm_window_partial<.*FpParams:0" \
            "accumulate_fp:k_msm_accumulate<.*FpParams:7" "accumulate_fp2:k_msm_accumulate<.*Fp2:0" "tape_wide_inv:k_tape_wide_inv:0" \
            "ntt_pass:k_ntt_pass:0" "fixup:k_msm_fixup<.*FpParams:0"; do
  name=${spec%%:*}; rest=${spec#*:}; rx=${rest%:*}; skip=${rest##*:}
  $NCU --set full --import-source on --kernel-name-base demangled -k "regex:$rx" --launch-skip $skip -c 1 -o gpurun_out/${P}_$name python tools/profile_one_wrap.py > gpurun_out/${P}_$name.log 2>&1
  tail -1 gpurun_out/${P}_$name.log
  ncu -i gpurun_out/${P}_$name.ncu-rep --page raw --csv > gpurun_out/${P}_$name.raw.csv 2>/dev/null
  case $name in accumulate_fp) ;; *) rm -f gpurun_out/${P}_$name.ncu-rep ;; esac
done
ls -la gpurun_out/ | head -40
