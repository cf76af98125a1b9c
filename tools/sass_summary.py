"""Per-kernel counts of the SASS mnemonics that identify the hardware paths (profiles/sass_summary.txt):
UTCHMMA / UTC*MMA (tcgen05.mma), UTMALDG (TMA loads), LDTM / STTM (tcgen05.ld / st), HMMA (mma.sync, the register-operand
tensor path of the narrow-MLP evaluator), FFMA, LDS / STS, BAR, SHFL, MUFU.  Run after mile_b200.build."""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parent.parent / 'mile_b200' / '_lib' / 'libmile_b200.so'
sass = subprocess.run(['cuobjdump', '-sass', str(lib)], capture_output=True, text=True, check=True).stdout
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDTM', 'STTM', 'UTCBAR', 'HMMA', 'FFMA', 'LDS', 'STS', 'LDG', 'STG',
        'BAR', 'SHFL', 'MUFU', 'SYNCS', 'UCGABAR']
cur, rows = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        rows[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        op = m.group(1)
        rows[cur]['total'] += 1
        for k in KEYS:
            if op.startswith(k):
                rows[cur][k] += 1
out = [f'# cuobjdump -sass {lib.name}: instruction counts per kernel (static code, sm_100a)', '']
for name, c in rows.items():
    short = re.sub(r'\(.*', '', name)[:110]
    out.append(f'{short}\n    total {c["total"]:6d}  ' + '  '.join(f'{k} {c[k]}' for k in KEYS if c[k]))
Path(sys.argv[1] if len(sys.argv) > 1 else 'profiles/sass_summary.txt').write_text('\n'.join(out) + '\n')
print('\n'.join(out[:40]))
