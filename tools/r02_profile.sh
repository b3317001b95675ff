#!/bin/bash
# Round-2 ncu evidence (run under gpurun on ONE B200): launch list of one complete wrap + one `--set full` capture per kernel.
# The raw metric pages are exported to CSV on the box (gpurun_out/ may carry 64 MiB back); only two reports are kept whole.
NCU="ncu --profile-from-start off --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches.csv python tools/profile_one_wrap.py > gpurun_out/r02_launches.log 2>&1
for spec in "tape_staged:k_tape_staged" "digits1:k_msm_digits<.*1>" "window_partial:k_msm_window_partial<.*FpParams" "accumulate_fp:k_msm_accumulate<.*FpParams" \
            "accumulate_fp2:k_msm_accumulate<.*Fp2" "tape_wide_inv:k_tape_wide_inv" "ntt_pass:k_ntt_pass" "fixup:k_msm_fixup<.*FpParams"; do
  name=${spec%%:*}; rx=${spec#*:}
  $NCU --set full --import-source on --kernel-name-base demangled -k "regex:$rx" -c 1 -o gpurun_out/r02_$name python tools/profile_one_wrap.py > gpurun_out/r02_$name.log 2>&1
  tail -1 gpurun_out/r02_$name.log
  ncu -i gpurun_out/r02_$name.ncu-rep --page raw --csv > gpurun_out/r02_$name.raw.csv 2>/dev/null
  case $name in accumulate_fp|tape_staged) ;; *) rm -f gpurun_out/r02_$name.ncu-rep ;; esac
done
ls -la gpurun_out/ | head -40
