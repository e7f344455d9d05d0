"""Pretty-print an AVT_ATTN_TRACE log of the attention backward (btrace <item> <id> <cycles>): per chunk, control / worker events."""
import sys
ev = {}
for line in open(sys.argv[1]):
    if line.startswith("btrace"):
        _, n, i, t = line.split()
        ev[(int(n), int(i))] = int(t)
names = [(130, "ctl:top"), (90, "ctl:mma1 issued"), (100, "ctl:bar_p seen"), (110, "ctl:mma2 issued"), (200, "wrk:bar_s seen"),
         (210, "wrk:bar_m2(g-2) seen"), (220, "wrk:arrived bar_p"), (230, "wrk:readout m2 seen"), (240, "wrk:chunk end")]
for n in (0, 1):
    print(f"item {n}: tiles landed {ev.get((n, 1))}  tiles free {ev.get((n, 120))}")
    for lc in range(8):
        row = sorted((ev[(n, b + lc)], nm) for b, nm in names if (n, b + lc) in ev)
        print(f"  chunk {lc}: " + "  ".join(f"{nm}={t}" for t, nm in row))
