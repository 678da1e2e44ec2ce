import os, sys, random
sys.path.insert(0, ".")
import quantumcomputer_b200 as q
from quantumcomputer_b200 import schedule_describe
lib = q.lib()
random.seed(7)
ok = bad = 0
for it in range(400):
    n = random.randint(1, 36)
    world = random.choice([1, 1, 2, 4, 8])
    rank = random.randrange(world)
    gates = []
    for _ in range(random.randint(0, 120)):
        if random.random() < 0.45:
            gates.append(("h", random.randrange(n)))
        else:
            gates.append(("cp", random.randrange(n), random.randrange(n), 0.3))
    try:
        schedule_describe(n, gates, world, rank); ok += 1
    except q.QcsError:
        bad += 1
print("schedule_describe ok", ok, "rejected", bad)
# bad arguments on purpose
for n, g in ((0, [("h", 0)]), (5, [("h", 7)]), (70, [("h", 3)]), (5, [("cp", 1, 9, 0.1)])):
    try:
        schedule_describe(n, g); print("accepted?!", n, g)
    except q.QcsError as e:
        pass
for _ in range(300):
    n = random.randint(10, 40); lo = random.randint(0, 12); tb = random.choice([0, 11, 12, 13])
    r = lib.qcs_pair_selfcheck(n, lo, n, tb)
print("pair_selfcheck done")
