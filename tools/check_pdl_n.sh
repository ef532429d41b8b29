# usage: bash tools/check_pdl_n.sh N   — multi-GPU parity drivers, then bench.py with and without programmatic dependent launch
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
timeout 200 $T tests/mgpu_parity.py --rows 403 --cols 97 --steps 120 2>&1 | tail -2
timeout 200 $T tests/mgpu_parity.py --rows 160 --cols 64 --steps 100 --graph random --graph-radius 7 2>&1 | grep -E "PARITY|MISMATCH|Error" | tail -3
timeout 200 $T tests/mgpu_parity.py --rows 403 --cols 97 --steps 80 --reward 2>&1 | tail -1
for v in 3 0; do
  SNN_B200_PDL=$v timeout 300 $T bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/r2_bench_n${N}_pdl$v.json 2> gpurun_out/r2_bench_n${N}_pdl$v.err
  python -c "
import sys,json
d=json.loads(open('gpurun_out/r2_bench_n${N}_pdl$v.json').readlines()[-1])
print('pdl mask $v:', d['value'], d['us_per_timestep'], 'strong', d['strong']['value'], d['strong']['us_per_timestep'], d['parity']['ok'], d['strong']['parity']['ok'])"
done
