"""Runs the blind-rotation kernel back to back for a few seconds while sampling clocks/power every 50 ms."""
import subprocess, sys, threading, time; sys.path.insert(0, '.')
import numpy as np
import peba1_b200 as P
from peba1_b200.cuda import make_gates
eng = P.Engine()
rng = np.random.default_rng(1)
p = eng.params
eng.load_keys(rng.integers(-2**31, 2**31, size=p.n*6*2*1024, dtype=np.int64).astype(np.int32),
              rng.integers(-2**31, 2**31, size=1024*8*4*631, dtype=np.int64).astype(np.int32))
n = 16576
eng.reserve(3*n); eng.upload(0, rng.integers(-2**31, 2**31, size=(2*n, 631), dtype=np.int64).astype(np.int32))
g = make_gates("NAND", np.arange(n), n+np.arange(n), 2*n+np.arange(n))
eng.enable_timing(True)
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu",
                         "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l.strip()) for l in proc.stdout], daemon=True).start()
time.sleep(0.5)
for rep in range(20):
    eng.gates(g)
    s = eng.stats()
    if rep % 4 == 0:
        print(f"rep {rep}: blind rotate {s.last_blind_rotate_ms:.2f} ms ({n/s.last_blind_rotate_ms*1e3:.0f} PBS/s)", flush=True)
proc.terminate()
clk = [float(r.split(',')[0]) for r in rows if r]
pw = [float(r.split(',')[1]) for r in rows if r]
print("samples", len(clk), "clock min/median/max", min(clk), sorted(clk)[len(clk)//2], max(clk), "power max", max(pw), "median", sorted(pw)[len(pw)//2])
print(rows[10:70:4])
