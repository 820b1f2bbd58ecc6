timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_me_ctu' -c 1 -f -o gpurun_out/prof_me_walk python tools/profile_prepass.py 1 > gpurun_out/q3.log 2>&1
ls -la gpurun_out/prof_me_walk.ncu-rep
