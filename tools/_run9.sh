cd $GRAFT_REPO_ROOT
N="ncu --set full --clock-control none --import-source on"
timeout 400 $N -k regex:ewald_rho_tiled_kernel -s 1 -c 1 -o gpurun_out/r2i_rho python tools/profile_step.py --workload spce --lattice 32 --steps 3 2>&1 | tail -1
timeout 400 $N -k regex:ewald_force_tiled_kernel -s 1 -c 1 -o gpurun_out/r2i_kforce python tools/profile_step.py --workload spce --lattice 32 --steps 3 2>&1 | tail -1
timeout 400 $N -k regex:rebuild2_kernel -c 1 -o gpurun_out/r2i_rebuild python tools/profile_step.py --steps 2 2>&1 | tail -1
timeout 400 $N -k regex:list_reorder2_kernel -c 1 -o gpurun_out/r2i_reorder python tools/profile_step.py --steps 2 2>&1 | tail -1
timeout 600 $N -k regex:lj2_force_kernel -s 4 -c 1 -o gpurun_out/r2i_lj2_8M python tools/profile_step.py --lattice 256x256x128 --steps 8 2>&1 | tail -1
ls -la gpurun_out/*.ncu-rep
