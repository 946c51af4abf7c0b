# generic A/B: each line of profiles/ab_variants.txt is a set of env assignments
run() {
  env $1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; e=d['e2e']
print('$1', 'fps', round(d['value'],1), 'e2e', round(e['value'],1), {k: round(v,4) for k,v in s.items()})
print('    e2e ms', round(e.get('ms_per_step',0),4), {k: round(v,4) for k,v in e.get('device_stage_ms_per_step',{}).items()})"
}
while read -r v; do [ -n "$v" ] && run "$v"; done < profiles/ab_variants.txt
