ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_r01k.csv python tools/bench_configs.py --config c5 > /dev/null 2>&1; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4_r01k.csv python tools/bench_configs.py --config c4 --views 32 > /dev/null 2>&1; echo rc=$?
