T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $T tests/mgpu_parity.py --rows 403 --cols 97 --steps 120 2>&1 | tail -2
timeout 200 $T tests/mgpu_parity.py --rows 160 --cols 64 --steps 100 --graph random --graph-radius 7 2>&1 | grep -E "PARITY|MISMATCH|Error" | tail -3
timeout 200 $T tests/mgpu_parity.py --rows 403 --cols 97 --steps 80 --reward 2>&1 | tail -1
timeout 500 $T bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
print(d['value'], d['us_per_timestep'], d['roofline']['frac'], d.get('e2e',{}).get('value'), d['clocks'])
print(' strong:', {k:d['strong'][k] for k in ('value','us_per_timestep','neurons_per_gpu')})
print(' parity:', d['parity']['ok'], d['parity']['checked'], d['parity']['boundaries_covered'], 'strong parity:', d['strong']['parity']['ok'], d['strong']['parity']['boundaries_covered'])
"
