"""Prints the constant tables of cc4_device.cuh that are derived from pair lists (same derivation as
oracle/cc4_oracle.c init_tables(); sources: EnterpriseScenarioGenerator.py:281-306, BlueFlatWrapper.py:267-302)."""
RZA, OZA, RZB, OZB, CON, PAZ, ADM, OFF, INT = range(9)


def pairs_to_masks(pairs, own=False):
    t = [0] * 9
    for a, b in pairs:
        t[a] |= 1 << b
        t[b] |= 1 << a
    if own:
        for s in range(9):
            t[s] |= 1 << s
    return t


hq = [PAZ, ADM, OFF]
base = [(h, x) for h in hq for x in (CON, RZA, RZB)]
green = [
    base + [(RZA, CON), (OZA, RZA), (RZB, CON), (RZB, RZA), (OZB, RZB)],
    base + [(RZB, CON), (OZB, RZB)],
    base + [(RZA, CON), (OZA, RZA)],
]
core = [INT, ADM, OFF, PAZ, CON, RZA, RZB]
complete = [(core[i], core[j]) for i in range(7) for j in range(i + 1, 7)] + [(RZA, OZA), (RZB, OZB)]
rm = [[], [(RZA, OZA), (RZA, CON), (RZA, RZB), (RZA, INT)], [(RZB, OZB), (RZB, CON), (RZB, RZA), (RZB, INT)]]


def fmt(rows):
    return ",\n".join("    { " + ", ".join(f"0x{v:03X}" for v in r) + " }" for r in rows)


print("c_green_allowed:\n" + fmt([pairs_to_masks(p, own=True) for p in green]))
comms = []
for ph in range(3):
    gone = {frozenset(x) for x in rm[ph]}
    comms.append(pairs_to_masks([p for p in complete if frozenset(p) not in gone]))
print("c_comms_adj:\n" + fmt(comms))
