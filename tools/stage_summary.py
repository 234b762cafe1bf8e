"""Summarise a BVG_PROF_DUMP file (written by bvg_profile_read) per generator stage (last step only)."""
import sys
recs = [l.split() for l in open(sys.argv[1])]
recs = [(int(c), float(ms), float(fl), float(by)) for c, ms, fl, by in recs]
convs = [r for r in recs if r[0] in (1, 2)]
acts = [r for r in recs if r[0] == 0]
nstep = len(convs) // 115
# lockstep AMP blocks (BVG_ACT_GROUP, the default in the 16-bit modes): 6 grouped Activation1d launches per stage + act_post,
# convolutions ordered (m, c1/c2, j); sequential: 18 launches per stage + act_post, convolutions ordered (j, m, c1/c2)
nact = len(acts) // max(nstep, 1)
lock = nact == 37
convs, acts = convs[-115:], acts[-nact:]
apb = 6 if lock else 18
tot = sum(r[1] for r in convs)
print(f"steps {nstep} conv {tot:.2f} ms  act {sum(r[1] for r in acts):.2f} ms")
print(f"pre {convs[0][1]:.3f} ms {convs[0][2] / convs[0][1] / 1e9:.0f} TF/s")
i = 1
for st in range(6):
    up = convs[i]; i += 1
    blk = convs[i:i + 18]; i += 18
    ms = sum(r[1] for r in blk); fl = sum(r[2] for r in blk)
    ks = []
    for j in range(3):
        c = [blk[m * 6 + h * 3 + j] for m in range(3) for h in range(2)] if lock else blk[j * 6:(j + 1) * 6]
        ks.append(f"k{(3, 7, 11)[j]} c1 {sum(x[1] for x in c[0::2]) / 3:.3f} c2 {sum(x[1] for x in c[1::2]) / 3:.3f}")
    a = acts[st * apb:(st + 1) * apb]
    print(f"stage {st}: up {up[1]:.3f} | convs {ms:.2f} ms {fl / ms / 1e9:.0f} TF/s | " + " | ".join(ks) +
          f" | act {sum(x[1] for x in a):.2f} ms {sum(x[3] for x in a) / sum(x[1] for x in a) / 1e6:.0f} GB/s")
