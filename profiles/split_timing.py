import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
from nuclear_sim_b200 import BatchedNuclearPlantSimulator, load_snapshot
from nuclear_sim_b200 import scenarios as sc
for n in (4096, 16384):
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    pid = np.arange(n)
    sim = BatchedNuclearPlantSimulator(n, sc.randomized_states(s0, pid), params)
    acts, mags = sc.load_following_inputs(pid, 0, 64)
    for _ in range(2):
        sim.step(actions=torch.from_numpy(acts), magnitudes=torch.from_numpy(mags), K=64)
    torch.cuda.synchronize()
    print("n", n, flush=True)
