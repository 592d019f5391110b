cd $GRAFT_REPO_ROOT
rm -f gpurun_out/parity_errors.jsonl
(time timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -6) 2>&1 | tail -10
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lj2_force_kernel -s 12 -c 1 -o gpurun_out/r2h_lj2 python tools/profile_step.py --steps 16 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cq_force_kernel -s 3 -c 1 -o gpurun_out/r2h_cq python tools/profile_step.py --workload spce --lattice 32 --steps 6 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r2h_launches_lj_1M.csv python tools/profile_step.py --steps 4 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2h_launches_spce_98k.csv python tools/profile_step.py --workload spce --lattice 32 --steps 3 2>&1 | tail -1
timeout 900 bash tools/sanitize.sh > gpurun_out/sanitize_all.log 2>&1; tail -12 gpurun_out/sanitize_all.log
