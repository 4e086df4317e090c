import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import *
import ctypes
dt = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
cfg = R.compose_config(sys.argv[2] if len(sys.argv) > 2 else 'oil_top_off')
rp = R.make_reference_plant(cfg, dt=5.0)
sim = rp.sim
sgs = sim.secondary_physics.steam_generator_system
p = R.extract_params(sim, strict=False)
rng = np.random.RandomState(0)
worst = 0.0
for k in range(400):
    s0 = R.extract_state(sim, strict=False)
    pw = 1.0 if k < 100 else (0.6 if k < 200 else (0.02 if k < 300 else 1.0))
    tin = list(327 * (0.9 + 0.1 * pw) + rng.uniform(-1, 1, 3)); tout = list(293 + rng.uniform(-1, 1, 3))
    if 250 < k < 300: tout = [t + 30 for t in tout]
    flows = list(5700 * max(0.3, pw) + rng.uniform(-50, 50, 3))
    fwf = list(500 * pw + rng.uniform(-20, 20, 3) * (pw > 0.1))
    if 200 < k < 230: fwf = [0.0] * 3
    ldf = max(0.2, pw)
    with R.quiet():
        res = sgs.update_system(primary_conditions={'inlet_temps': tin, 'outlet_temps': tout, 'flow_rates': flows},
                                steam_demands={'load_demand_fraction': ldf, 'steam_pressure': 6.895},
                                system_conditions={'feedwater_temperature': 227.0, 'load_demand': pw, 'actual_feedwater_flows': fwf},
                                control_inputs={}, dt=dt)
    s1 = R.extract_state(sim, strict=False)
    c = s0.copy()
    LIB.nps_oracle_sg_system(ptr(c), ptr(p), ptr(dvec(tin)), ptr(dvec(tout)), ptr(dvec(flows)), ctypes.c_double(ldf),
                             ctypes.c_double(pw), ctypes.c_double(227.0), ptr(dvec(fwf)), ctypes.c_double(dt))
    bad, mx = compare(c, s1, prefix="sgs.")
    worst = max(worst, mx)
    if bad and mx > float(os.environ.get("TOL", "1e-12")):
        print("step", k, "max", mx)
        for b in bad: print("   ", b)
        break
print("worst rel err", worst, "Q", res['total_thermal_power'], 'P', res['average_steam_pressure'], 'lvl', res['sg_levels'])
