# dev tool: bench.py under a few environment settings (one line each: value, e2e, gpu busy)
for cfg in "$@"; do
  echo "== $cfg"; env $cfg python bench.py --steps 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['gpu_busy_estimate'],3), d['roofline']['kernel'], round(d['roofline']['frac'],3))"
done
