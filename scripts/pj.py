import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']),
      'stage_ms', {k: round(v, 4) for k, v in d['stage_ms'].items()})
