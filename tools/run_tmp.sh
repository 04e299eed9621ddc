ncu --set full --clock-control none --import-source on -k regex:skel_palette -s 3 -c 1 -o gpurun_out/prof_skel2 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo rc=$?
