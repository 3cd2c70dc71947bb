cd $GRAFT_REPO_ROOT
SECONDS=0; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; echo rc=$?
echo wall_s=$SECONDS
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_full.json').read().strip().splitlines()[-1])
print("value %.3e ms %.2f e2e %.3e link %.1f d2h %.1f" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['link_ceiling_gbs'], d['e2e']['d2h_gbs_per_gpu']))
print("pipe", d['e2e']['device_resident_pipeline'])
print("roof", d['roofline'])
print("cpu", d.get('cpu_baseline'))
for k,v in d.get('secondary',{}).items():
    print(k, {kk: v.get(kk) for kk in ('value','ms_per_step','us_per_frame','wall_s','note','survivor_fraction')}, v.get('roofline',{}).get('frac'), v.get('cpu_baseline'))
PY
tail -5 gpurun_out/r2_bench_full.err
