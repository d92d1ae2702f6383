#!/bin/bash
# End-of-round evidence pass on one B200 (run under gpurun from the repo root): full GPU test suite, smoke(), the bench lines of every
# workload, the launch list of the default bench and the ncu --set full capture of one full-size encoder layer.
TAG=${1:-final4}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
python bench.py > gpurun_out/r02_bench_default_$TAG.jsonl 2> gpurun_out/bench_default_$TAG.err; echo "bench rc=$?"
for w in cfg1 cfg2 cfg3 cfg4; do
  python bench.py --workload $w --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/r02_bench_${w}_$TAG.jsonl 2> gpurun_out/bench_${w}_$TAG.err; echo "$w rc=$?"
done
python bench.py --precision fp32 --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/r02_bench_fp32_$TAG.jsonl 2> gpurun_out/bench_fp32_$TAG.err; echo "fp32 rc=$?"
for f in gpurun_out/r02_bench_*_$TAG.jsonl; do python -c "
import json,sys
j=json.loads(open('$f').read().strip().splitlines()[-1]); r=j['roofline']
print('$f', round(j['value']), round(j['ms_per_step'],3), r.get('kernel'), round(r.get('frac',0),4), j['clocks']['sm_mhz'], j['clocks']['reasons'])"; done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
python tools/agg_launches.py gpurun_out/r02_launches_$TAG.csv > gpurun_out/r02_launches_${TAG}_summary.txt 2>&1; head -16 gpurun_out/r02_launches_${TAG}_summary.txt
bash tools/ncu_capture_r02.sh $TAG > gpurun_out/ncu_capture_$TAG.log 2>&1; tail -14 gpurun_out/ncu_capture_$TAG.log
