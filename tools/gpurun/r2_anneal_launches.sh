mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/anneal_launches.csv python tools/anneal_launches.py > gpurun_out/anneal_launches.log 2>&1
tail -2 gpurun_out/anneal_launches.log
