import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import *
import ctypes
dt = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
cfg = R.compose_config(sys.argv[2] if len(sys.argv) > 2 else 'oil_top_off')
rp = R.make_reference_plant(cfg, dt=dt)
sim = rp.sim
cond = sim.secondary_physics.condenser
p = R.extract_params(sim, strict=False)
rng = np.random.RandomState(0)
worst = 0.0
mk = {'tds': 300.0, 'hardness': 100.0, 'chloride': 30.0, 'ph': 7.2, 'dissolved_oxygen': 8.0}
cd = {'chlorine': 1.0, 'antiscalant': 5.0, 'corrosion_inhibitor': 10.0, 'biocide': 0.0}
keys = ['heat_rejection_rate', 'condenser_pressure', 'cooling_water_temp_rise', 'cooling_water_outlet_temp', 'thermal_performance_factor', 'vacuum_system_efficiency', 'condensate_temperature']
for k in range(400):
    s0 = R.extract_state(sim, strict=False)
    pw = 1.0 if k < 100 else (0.3 if k < 200 else 0.8)
    inp = [0.007, 10.0, 1300.0 * pw + rng.uniform(-20, 20), np.float64(0.88 + rng.uniform(-0.05, 0.05)), 45000.0 * (1.0 if k % 40 else 3.2),
           25.0 + rng.uniform(-3, 3), 1.2 if (k % 70) else 0.85, 185.0]
    if 300 < k < 320: cond.vacuum_system.condenser_pressure = 0.0085
    s0 = R.extract_state(sim, strict=False)
    with R.quiet():
        res = cond.update_state(steam_pressure=inp[0], steam_temperature=inp[1], steam_flow=inp[2], steam_quality=inp[3],
                                cooling_water_flow=inp[4], cooling_water_temp_in=inp[5], motive_steam_pressure=inp[6],
                                motive_steam_temperature=inp[7], makeup_water_quality=mk, chemical_doses=cd, dt=dt / 60.0)
    s1 = R.extract_state(sim, strict=False)
    c = s0.copy(); out = np.zeros(7)
    LIB.nps_oracle_condenser(ptr(c), ptr(p), ptr(dvec(inp)), ctypes.c_double(dt / 60.0), ptr(out))
    bad, mx = compare(c, s1, prefix="cond.")
    worst = max(worst, mx)
    ref_out = np.array([res[q] for q in keys])
    oerr = np.max(np.abs(out - ref_out) / np.maximum(np.abs(ref_out), 1e-300))
    if (bad and mx > float(os.environ.get("TOL", "1e-12"))) or oerr > 1e-12:
        print("step", k, "max", mx, "out err", oerr)
        for b in bad: print("   ", b)
        print(out, ref_out)
        break
print("worst rel err", worst, "Q", res['heat_rejection_rate'], 'P', res['condenser_pressure'], [e.is_operating for e in cond.vacuum_system.ejectors.values()])
