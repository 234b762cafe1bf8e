"""Summarise gpurun_out/conv_trace.txt (BVG_CONV_TRACE): per role, mean cycles between consecutive events."""
import sys, collections
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/conv_trace.txt"
rows = collections.defaultdict(list)
for ln in open(path):
    if ln.startswith("#"):
        print(ln.strip()); continue
    role, ev, tile, clk = map(int, ln.split())
    rows[role].append((ev, tile, clk))
names = {0: "producer", 1: "mma", 2: "epilogue", 3: "act"}
for role, evs in sorted(rows.items()):
    d = collections.defaultdict(list)
    for (e0, t0, c0), (e1, t1, c1) in zip(evs, evs[1:]):
        d[(e0, e1)].append((c1 - c0) & 0xffffffffff)
    tiles = sorted({t for _, t, _ in evs})
    span = (evs[-1][2] - evs[0][2]) & 0xffffffffff
    print(f"role {role} {names.get(role)}: {len(evs)} events, {len(tiles)} tiles, span {span} cyc, {span / max(len(tiles), 1):.0f} cyc/tile")
    for k, v in sorted(d.items()):
        v2 = v[2:] if len(v) > 4 else v
        print(f"   ev{k[0]}->ev{k[1]}: n={len(v)} mean {sum(v2) / len(v2):.0f} min {min(v2)} max {max(v2)}")
