set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/r1i_launches.csv python tools/profile_step.py --steps 60 2>&1 | tail -2
