# generic A/B: each line of VARIANTS is a set of env assignments
run() {
  env $1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('$1', 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v,4) for k,v in s.items()})"
}
while read -r v; do [ -n "$v" ] && run "$v"; done < profiles/ab_variants.txt
