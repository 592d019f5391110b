set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lj_force_kernel -s 2 -c 1 -o gpurun_out/r1f_lj_force python tools/profile_step.py --steps 4 2>&1 | tail -3
