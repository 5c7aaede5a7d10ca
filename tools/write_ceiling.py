"""Steady-state bandwidth of a plain streaming write (pk_fill, 256 MiB) -- the ceiling a
write-only kernel such as pk_expand_blocks can be compared with."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pockit_b200.lobatto as lob
from pockit_b200 import problems
from pockit_b200.engine import Engine
S = problems.lqr(lob, 4, 4)
e = Engine(S.lowering)
for _ in range(10): e.flush_l2()
e.sync()
n = 200
t0 = time.perf_counter()
for _ in range(n): e.flush_l2()
e.sync()
dt = (time.perf_counter() - t0) / n
print(f"pk_fill 256 MiB: {dt*1e6:.1f} us per launch = {268435456/dt/1e9:.0f} GB/s (wall clock over {n} back-to-back launches)")
