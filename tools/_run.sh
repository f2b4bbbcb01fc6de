python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "not fuzz and not checked" > gpurun_out/tests_r02h.log 2>&1; tail -3 gpurun_out/tests_r02h.log
python bench.py --no-cpu --no-strong > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02h.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms']); [print(x) for x in d['sweep']['cfg4']]; [print(x['cfg'], x['device'], x['e2e_pinned']) for x in d['sweep']['configs']]"
