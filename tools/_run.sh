python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/tests_r02c.log 2>&1; tail -3 gpurun_out/tests_r02c.log
python bench.py --no-cpu --no-sweep --no-strong > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; head -c 400 gpurun_out/bench_r02c.json; echo
python tools/exp_geometry.py 4 2 > gpurun_out/geom_r02c.log 2>&1; cat gpurun_out/geom_r02c.log
