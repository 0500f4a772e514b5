# repeat the inference legs of bench.py (resident, e2e, uint8) and report every exit code
mkdir -p gpurun_out
for rep in 1 2 3 4 5 6; do
  timeout 100 python bench.py --no-train > gpurun_out/bench_rep$rep.log 2>&1; echo "rep=$rep rc=$?"
done
grep -h '^{' gpurun_out/bench_rep*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('e2e_uint8', {}).get('value'), d['clocks'])
"
