"""Generates the committed fixtures under tests/golden/ from the read-only reference checkout.
Run here (where /root/reference exists):  python tests/golden/make_fixtures.py

  sampler_configs.json  the `training.sampler` / `model` / `data` blocks of the reference's MCLMC and NUTS experiment
                        YAMLs (experiments/**/mclmc.yaml, nuts.yaml, covertype.yaml ...): configuration DATA the host
                        mirror must parse unchanged.  (The reference holds no golden vectors for the
                        arithmetic itself -- SURVEY.md section 4.)
"""
import json
from pathlib import Path

import yaml

REF = Path('/root/reference/experiments')
OUT = Path(__file__).resolve().parent
cfgs = {}
for p in sorted(REF.rglob('*.yaml')):
    try:
        y = yaml.safe_load(p.read_text())
    except Exception:
        continue
    if not isinstance(y, dict) or 'training' not in y or 'sampler' not in (y.get('training') or {}):
        continue
    s = y['training']['sampler']
    if s.get('name') not in ('mclmc', 'nuts'):
        continue
    cfgs[str(p.relative_to(REF))] = {'sampler': s, 'model': y.get('model'),
                                     'data': {k: y['data'].get(k) for k in ('path', 'task', 'train_split', 'valid_split', 'test_split')}}
(OUT / 'sampler_configs.json').write_text(json.dumps(cfgs, indent=1, sort_keys=True))
print(len(cfgs), 'configs written')
