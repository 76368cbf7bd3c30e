for args in "" "--lines-per-pass 512" "--in-flight 16" "--lines-per-pass 512 --in-flight 16" ""; do
  python bench.py --steps 12 --warmup 3 --no-side-c2 --no-api --cpu-seconds 0 $args 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$args', '|', round(d['value']), round(d['e2e']['value']), d['token_identity']['identical_lines'], d['clocks']['sm_mhz'])"
done
