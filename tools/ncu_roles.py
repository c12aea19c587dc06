#!/usr/bin/env python
"""Per-role stall profile of the fused pass from an ncu report with sources (test/profiling harness).
Usage: tools/ncu_roles.py gpurun_out/<tag>_fused_outer_kernel.ncu-rep
Splits the SASS of fused_outer_kernel into its warp roles at the USETMAXREG instructions / role loops (by execution count) and
prints, per role: warp-samples, share of the time the role waits on an mbarrier (long scoreboard at the try-wait branch), the
top stall reasons and the instructions executed per tile."""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(f(r, '# Samples') for r in data)
src = [r[ix['Source']] for r in data]
# role boundaries: producer ... U ... C ... A ... epilogue, found from the USETMAXREG instructions
marks = [i for i, s in enumerate(src) if 'USETMAXREG' in s]
print('instructions', len(data), 'samples', int(tot), 'USETMAXREG at', marks)
bounds = [0] + marks + [len(data)]
ntile = max(f(r, 'Instructions Executed') for r in data if 'UBLKCP' in r[ix['Source']]) if any('UBLKCP' in s for s in src) else 1.0
print('tiles (bulk-copy executions / copies per tile):', ntile)
for a, b in zip(bounds[:-1], bounds[1:]):
    seg = data[a:b]
    s = sum(f(r, '# Samples') for r in seg)
    if s == 0: continue
    st = {k: sum(f(r, k) for r in seg) for k in stalls}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:7]
    ex = sum(f(r, 'Instructions Executed') for r in seg)
    ops = {}
    for r in seg:
        t = [o for o in r[ix['Source']].split() if not o.startswith('@')]
        if not t: continue
        op = t[0].split('.')[0]
        ops[op] = ops.get(op, 0) + f(r, 'Instructions Executed')
    topo = sorted(ops.items(), key=lambda kv: -kv[1])[:10]
    print(f"[{a:5d},{b:5d}) samples {s:8.0f} {100 * s / tot:5.1f}%  inst/tile {ex / ntile:8.1f}  " + ' '.join(f"{k[6:]}={100 * v / s:.0f}%" for k, v in top))
    print("      ops/tile: " + ' '.join(f"{k}={v / ntile:.0f}" for k, v in topo))
