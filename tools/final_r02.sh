set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/gputest_final.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python bench.py --dtype f32 --no-cpu-baseline --no-sub-benchmarks > gpurun_out/bench_ns_f32.json 2> gpurun_out/bench_ns_f32.err
python probes/c3_phases.py > gpurun_out/c3_phases_final.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
cat gpurun_out/gputest_final.log gpurun_out/smoke.log
