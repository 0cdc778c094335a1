# per-task bench table (profiles/rNN_bench_tasks.jsonl): every task shape, GAUSS / JAC (fp32 SIMT) / JAC (opt-in tcgen05 kernel)
out=gpurun_out/bench_tasks.jsonl; : > $out
for task in slcp lotka_volterra sir; do
  python bench.py --task $task --steps 3 --no-cpu-baseline 2>/dev/null | tail -1 >> $out
  python bench.py --task $task --steps 3 --cov-mode JAC --no-cpu-baseline 2>/dev/null | tail -1 >> $out
  python bench.py --task $task --steps 3 --cov-mode JAC --jac-tc --no-cpu-baseline 2>/dev/null | tail -1 >> $out
done
python bench.py --task jrnnm --steps 3 --no-cpu-baseline 2>/dev/null | tail -1 >> $out
python bench.py --task jrnnm --steps 3 --cov-mode JAC --no-cpu-baseline 2>/dev/null | tail -1 >> $out
python bench.py --task pf_slcp --steps 3 --no-cpu-baseline 2>/dev/null | tail -1 >> $out
python bench.py --task stress --n-obs 10000 --samples 2048 --ddim-steps 10 --steps 2 --no-cpu-baseline 2>/dev/null | tail -1 >> $out
python -c "
import json
for l in open('$out'):
    d=json.loads(l); print(d['config']['workload'].split(' shape')[0], d['config']['cov_mode'], d['config']['jac_kernel'], round(d['value']), round(d['e2e']['value']))
"
