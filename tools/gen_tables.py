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

# c_route[src subnet][dst subnet]: subnets whose routers lie on the tree path (Action.py:101-117, topology
# EnterpriseScenarioGenerator.py:373-416), 4 bits each from bit 0, count in bits 28-31; the internet "router" is subnet 8
parent = {RZA: INT, OZA: RZA, RZB: INT, OZB: RZB, CON: INT, PAZ: INT, ADM: PAZ, OFF: PAZ, INT: None}


def chain(v):
    out = []
    while v is not None:
        out.append(v)
        v = parent[v]
    return out


def route(a, b):
    ca, cb = chain(a), chain(b)
    ia, ib = len(ca) - 1, len(cb) - 1
    while ia > 0 and ib > 0 and ca[ia - 1] == cb[ib - 1]:
        ia -= 1
        ib -= 1
    return ca[:ia + 1] + cb[:ib][::-1]


rows = []
for a in range(9):
    row = []
    for b in range(9):
        p = route(a, b)
        assert len(p) <= 5
        row.append(sum(v << (4 * i) for i, v in enumerate(p)) | len(p) << 28)
    rows.append(row)
print("c_route:\n" + ",\n".join("    { " + ", ".join(f"0x{v:08X}u" for v in r) + " }" for r in rows))
