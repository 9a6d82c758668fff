set -x
timeout 600 ncu --set full --import-source on --clock-control none -k regex:walk_kernel -c 1 -o gpurun_out/r2h_walk_s48 -f python scripts/profile_run.py 48 48 > gpurun_out/r2h_ncu_walk.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:push_kernel -s 3 -c 1 -o gpurun_out/r2h_push_s48 -f python scripts/profile_run.py 48 48 > gpurun_out/r2h_ncu_push.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches_bench_s48.csv python scripts/profile_run.py 48 48 > gpurun_out/r2h_launches.log 2>&1
timeout 600 python bench.py > gpurun_out/r2h_bench_1gpu.json 2> gpurun_out/r2h_bench_1gpu.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2h_bench_reference_arm.json 2> gpurun_out/r2h_bench_reference_arm.err
tail -c 400 gpurun_out/r2h_bench_1gpu.json; tail -c 300 gpurun_out/r2h_bench_reference_arm.json
