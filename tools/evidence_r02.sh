set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/gputest.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -c 600 gpurun_out/bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ns_f64.csv python probes/one_eval.py 4096 16 65 > gpurun_out/ncu_ns.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3_f64.csv python probes/one_eval.py 1024 8 33 > gpurun_out/ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_node128_v2 --launch-skip 100 -c 1 -o gpurun_out/node_v2_full python probes/one_eval.py 1024 8 33 > gpurun_out/ncu_node.log 2>&1
ls -la gpurun_out | tail -12
