set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_final.log
timeout 400 python bench.py > gpurun_out/r2_bench_line.json 2> gpurun_out/r2_bench_line.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_win.csv python bench.py --steps 2 --warmup 1 --iters 100 --no-cpu-baseline --no-configs --no-parity --no-e2e > /dev/null 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_win_kernel -s 250 -c 1 -f -o /tmp/prof_win python bench.py --steps 2 --warmup 1 --iters 100 --no-cpu-baseline --no-configs --no-parity --no-e2e > /dev/null 2>&1; echo "ncu win rc=$?"
python profiles/summarize.py /tmp/prof_win.ncu-rep gpurun_out/r2_step_win_bench_full.txt > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_win_kernel -s 80 -c 1 -f -o /tmp/prof_hhq python tools/bench_hh.py 2048 quiet > /dev/null 2>&1; echo "ncu hh rc=$?"
python profiles/summarize.py /tmp/prof_hhq.ncu-rep gpurun_out/r2_step_win_hh_finite_full.txt > /dev/null 2>&1
timeout 200 python tools/bench_configs.py > gpurun_out/r2_configs.json 2>&1
timeout 120 python tools/bench_hh.py 2048 quiet > gpurun_out/r2_bench_hh.txt 2>&1
timeout 120 python tools/bench_reward.py > gpurun_out/r2_bench_reward.json 2>&1
head -12 gpurun_out/r2_step_win_bench_full.txt; head -12 gpurun_out/r2_step_win_hh_finite_full.txt; cat gpurun_out/r2_configs.json | tr -d "\n "; echo; grep HH gpurun_out/r2_bench_hh.txt; cat gpurun_out/r2_bench_reward.json
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_line.json'))
print({k:d[k] for k in ('value','us_per_timestep','clocks','gpu_launches')}); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline']); print(d['parity']['ok'], d['parity']['checked'])
r=json.load(open('gpurun_out/r2_bench_reference.json')); print('REF', r['value'], r['ms_per_step'], r['config']['timesteps_per_step'], r['cpu_baseline']['sample'])
"
