for cfg in "BPPP_LANES=8" "BPPP_LANES=12" "BPPP_LANES=16" "BPPP_LANES=24" "BPPP_LANES=4" "BPPP_LANES=16 BPPP_LANE_THREADS=4" "BPPP_LANES=8 BPPP_LANE_THREADS=8"; do
  echo "== $cfg"; env $cfg python bench.py --steps 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['roofline']['frac'])"
done
