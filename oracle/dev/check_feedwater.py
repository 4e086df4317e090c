import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import *
import ctypes

rp = R.make_reference_plant(dt=5.0)
sim = rp.sim
sec = sim.secondary_physics
fw = sec.feedwater_system
p = R.extract_params(sim, strict=False)
rng = np.random.RandomState(0)
worst = 0.0
dt = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
for k in range(300):
    s0 = R.extract_state(sim, strict=False)
    levels = list(12.5 + rng.uniform(-0.5, 0.5, 3) + (k > 200) * 0.0)
    flows = list(500 + rng.uniform(-50, 50, 3) - (k > 150) * 100)
    quals = list(0.99 + rng.uniform(-0.02, 0.005, 3))
    sgc = {'levels': levels, 'pressures': [6.9] * 3, 'steam_flows': flows, 'steam_qualities': quals}
    sysc = {'sg_pressure': 6.895, 'feedwater_temperature': 40.0, 'suction_pressure': 0.5, 'discharge_pressure': 7.4}
    with R.quiet():
        res = fw.update_state(sg_conditions=sgc, steam_generator_demands={'total_flow': sum(flows)},
                              system_conditions=sysc, control_inputs={'load_demand': 100.0}, dt=dt)
    s1 = R.extract_state(sim, strict=False)
    c = s0.copy()
    out = np.zeros(5)
    LIB.nps_oracle_feedwater(ptr(c), ptr(p), ptr(dvec(levels)), ptr(dvec(flows)), ptr(dvec(quals)),
                             ctypes.c_double(sum(flows)), ctypes.c_double(40.0), ctypes.c_double(0.5),
                             ctypes.c_double(7.4), ctypes.c_double(dt), ptr(out))
    bad, mx = compare(c, s1)
    worst = max(worst, mx)
    if bad and mx > float(os.environ.get("TOL","1e-12")):
        print("step", k, "max", mx)
        for b in bad: print("   ", b)
        break
    ref_out = [res['total_flow_rate'], res['total_power_consumption'], res['num_running_pumps'], float(res['system_availability'])]
    if not np.allclose(out[:4], ref_out, rtol=1e-13):
        print("out mismatch", out, ref_out); break
print("worst rel err", worst, "total flow", res['total_flow_rate'], 'n run', res['num_running_pumps'], 'oil', s1[R._layout().field_index()['fw.pump[0].lub.oil_level']])
