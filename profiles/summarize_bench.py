"""Print the key fields of bench.py JSON lines (one short row per line)."""
import json
import sys

for path in sys.argv[1:]:
    for ln in open(path):
        ln = ln.strip()
        if not ln.startswith('{'):
            continue
        d = json.loads(ln)
        r = d.get('roofline') or {}
        e = d.get('e2e') or {}
        c = d.get('cpu_baseline') or {}
        print('%s | %s | value %.3e | ms/step %.2f | frac %.3f (%.1f us/launch) | e2e %.3e | cpu %.3e (%s cores) | clocks %s' % (
            d.get('impl', 'ours'), d['config']['workload'][:60], d['value'], d['ms_per_step'], r.get('frac', 0), r.get('avg_launch_us', 0),
            e.get('value', 0), c.get('value', 0), c.get('cores'), d.get('clocks')))
        py = d.get('cpu_baseline_python') or {}
        if py:
            print('    reference Python vector env: %.3e env-steps/s (%s)' % (py.get('value', 0), py.get('sample', '')[:90]))
        for v in d.get('variants') or []:
            vr = v.get('roofline') or {}
            print('    variant | %s | value %.3e | frac %.3f (%.1f us/launch)' % (v['workload'][:70], v['value'], vr.get('frac', 0), vr.get('avg_launch_us', 0)))
