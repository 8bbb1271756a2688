# r02ao: single-copy scene upload: full GPU suite + the bench's e2e phases
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02ao_tests_all.log 2>&1; tail -3 gpurun_out/r02ao_tests_all.log
timeout 600 python bench.py --no-ref-cuda --no-cpu-baseline --steps 5 > gpurun_out/r02ao_bench_c3.json 2> gpurun_out/r02ao_bench_c3.err; python -c "
import json; d=json.load(open('gpurun_out/r02ao_bench_c3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['phases_ms_rank0'])"
