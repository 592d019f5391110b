timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiled_kspace or structure_factor or water" 2>&1 | tail -5
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ewald -c 4 --csv --log-file gpurun_out/r1n_ewald_launches.csv python tools/profile_step.py --workload spce --lattice 32 --steps 1 2>&1 | tail -1
