"""Pretty-print an AVT_ATTN_TRACE log of the attention forward: one row per tile, columns = roles/phases (cycles)."""
import sys
ev = {}
for line in open(sys.argv[1]):
    if line.startswith("ftrace"):
        _, i, t = line.split()
        i, t = int(i), int(t)
        ev[(i // 96, (i % 96) // 16, i % 16)] = t
names = {(4, 0): "S.iss", (4, 1): "S.done", (1, 0): "aux:S", (1, 2): "aux:max", (0, 2): "exp0:go", (0, 3): "exp0:end", (2, 2): "exp1:go",
         (2, 3): "exp1:end", (3, 0): "pv0", (3, 4): "pv0e", (3, 1): "pv1", (3, 5): "pv1e", (3, 2): "pv2", (3, 6): "pv2e", (3, 3): "pv3", (3, 7): "pv3e",
         (1, 5): "aux:O", (1, 6): "aux:drained"}
for u in range(6):
    row = sorted((t, names.get((r, ph), f"{r}.{ph}")) for (r, uu, ph), t in ev.items() if uu == u)
    print(f"tile {u}: " + "  ".join(f"{n}={t}" for t, n in row))
