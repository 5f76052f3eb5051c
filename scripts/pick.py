"""Prints selected fields of the last JSON line on stdin: python scripts/pick.py key1 key2 ..."""
import json, sys
line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]
d = json.loads(line)
out = {}
for k in sys.argv[1:]:
    v = d
    for part in k.split("."):
        v = v[part] if isinstance(v, dict) else None
    out[k] = v
print(json.dumps(out))
