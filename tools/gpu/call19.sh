mkdir -p gpurun_out
# 1) concurrent instances per GPU: 24 / 32 / 40 with the final build (timed region only)
for b in 24 32 40; do
  TN_BENCH_LITE=1 timeout 400 python bench.py --steps 2 --warmup 1 --batch $b --no-extra > gpurun_out/bench19_b$b.json 2> gpurun_out/bench19_b$b.err
  tail -c 200 gpurun_out/bench19_b$b.err
done
BEST=$(python - <<'PY'
import json
best = None
for b in (24, 32, 40):
    try:
        v = json.loads(open('gpurun_out/bench19_b%d.json' % b).read().strip().splitlines()[-1])['value']
    except Exception:
        continue
    print('batch', b, v, file=__import__('sys').stderr)
    # a larger batch must win by more than the run-to-run noise (3 %)
    if best is None or v < best[1] * 0.97:
        best = (b, v)
print(best[0] if best else 24)
PY
)
echo "best batch $BEST"
# 2) the default bench line at that batch
timeout 900 python bench.py --steps 3 --warmup 3 --batch $BEST > gpurun_out/bench19_default.json 2> gpurun_out/bench19_default.err; tail -c 300 gpurun_out/bench19_default.json; tail -3 gpurun_out/bench19_default.err
# 3) steady-state launch list: 1500 launches from the middle of one instance
TN_BENCH_LITE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 30000 -c 1500 --csv --log-file gpurun_out/launches_bench19_mid.csv python bench.py --steps 1 --warmup 0 --batch 1 --no-extra > gpurun_out/ncu_bench19.log 2>&1
python tools/summarise_launches.py gpurun_out/launches_bench19_mid.csv "ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 30000 -c 1500: TN_BENCH_LITE=1 python bench.py --steps 1 --warmup 0 --batch 1 --no-extra (1500 launches from the middle of the boundary-MPS build of one instance, round 2 final build)" > gpurun_out/launches_bench19_mid_summary.csv; head -40 gpurun_out/launches_bench19_mid_summary.csv
