"""Whole-plant comparison: teacher-forced (each step starts from the reference state) and
free-running (C trajectory never re-synchronised)."""
import sys, os, argparse
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import *
import ctypes
ap = argparse.ArgumentParser()
ap.add_argument("--dt", type=float, default=5.0)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--action", default="oil_top_off")
ap.add_argument("--heat", default="constant")
ap.add_argument("--noise", type=int, default=1)
ap.add_argument("--tol", type=float, default=1e-12)
ap.add_argument("--policy", default="none")
a = ap.parse_args()
cfg = R.compose_config(a.action)
rp = R.make_reference_plant(cfg, dt=a.dt, heat_source=a.heat, noise_enabled=bool(a.noise))
sim = rp.sim
p = R.extract_params(sim)
names = R._layout().field_names()
rng = np.random.RandomState(7)
free = R.extract_state(sim)
worst_tf = 0.0; worst_free = 0.0
def cstep(c, act, mag, noise):
    LIB.nps_oracle_step(ptr(c), ptr(p), ptr(np.array([act], dtype=np.int8)), ptr(np.array([mag])), ptr(noise), ctypes.c_int64(1), 1)
for k in range(a.steps):
    s0 = R.extract_state(sim)
    noise = np.array([rng.standard_normal(), rng.standard_normal(), rng.random_sample(), rng.random_sample(), rng.random_sample()])
    if a.policy == "rods":
        act = 1 if np.sin(k / 100.0) > 0.5 else (0 if np.sin(k / 100.0) < -0.5 else 8)
        if k % 7 == 3: act = [2, 3, 9, 10, 4, 5][(k // 7) % 6]
    else:
        act = 8
    mag = 0.6
    rp.step(act, mag, noise)
    s1 = R.extract_state(sim)
    c = s0.copy(); cstep(c, act, mag, noise)
    cstep(free, act, mag, noise)
    bad, mx = compare(c, s1, top=15)
    worst_tf = max(worst_tf, mx)
    if bad and mx > a.tol:
        print("TEACHER-FORCED step", k, "max", mx)
        for b in bad: print("   ", b)
        break
    badf, mxf = compare(free, s1, top=8)
    worst_free = max(worst_free, mxf)
ix = R._layout().field_index()
print(f"steps {k+1} worst teacher-forced {worst_tf:.3e}  worst free-running {worst_free:.3e}  P_el {s1[ix['sec.electrical_power_output']]:.3f} power {s1[ix['pri.power_level']]:.4f}")
for b in badf[:5]: print("   free:", b)
