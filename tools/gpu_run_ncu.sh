set -x
timeout 60 python tools/phase_times.py C3 > gpurun_out/r02c_phase_times.log 2>&1; echo "plain rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_align|k_split|k_dt_tree|k_compose" -s 18 -c 6 -f -o gpurun_out/r02c_full python tools/phase_times.py C3 > gpurun_out/r02c_ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r02c_full.ncu-rep; cat gpurun_out/r02c_phase_times.log
