set -x
T=r1n
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
python bench.py --mode model --steps 5 --warmup 3 > gpurun_out/bench_${T}_model.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${T}_ref.json 2>/dev/null
# launch list of the bench command itself (after it exited 0 without ncu)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches_bench_py.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_${T}_bench.log 2>&1
# full capture of the two dominant kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_icp -s 9 -c 2 -f -o gpurun_out/${T}_icp python tools/profile_step.py --batch 300 --groups 1 > gpurun_out/ncu_${T}_icp.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ingest -c 1 -f -o gpurun_out/${T}_ingest python tools/profile_step.py --batch 300 --groups 1 > gpurun_out/ncu_${T}_ingest.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r1n*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j['value']), round(j['e2e']['value']), {k:v['ms'] for k,v in (j.get('per_kernel_ms_per_step') or {}).items()})
    except Exception as e: print(f, 'ERR', e)
PY
ls -la gpurun_out | tail -12
