"""Print the steady-state part of a TANTE_TAIL_TRACE dump (tools/tail_trace.py) as a merged MMA / epilogue timeline."""
import struct, sys
def load(path):
    raw = open(path, 'rb').read()
    vals = struct.unpack(f"{len(raw)//8}q", raw)
    roles = [[(vals[(r*1024+i)*2], vals[(r*1024+i)*2+1]) for i in range(1024) if vals[(r*1024+i)*2+1] > 0] for r in range(3)]
    t0 = min(e[0][1] for e in roles if e)
    return sorted((t - t0, r, e) for r, ev in enumerate(roles) for e, t in ev)
ev = load(sys.argv[1])
print('total cycles', ev[-1][0])
idx = [i for i, (t, r, e) in enumerate(ev) if r == 1 and e == 10]
a, b = idx[int(sys.argv[2]) if len(sys.argv) > 2 else 5], idx[int(sys.argv[3]) if len(sys.argv) > 3 else 7]
prev = ev[a][0]
for t, r, e in ev[a:b]:
    print(f"  {t:8d} (+{t-prev:6d}) {('MMA', 'EPI', 'EP1')[r]:4s} {e}")
    prev = t
