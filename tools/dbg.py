import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nbody_cosmological_simulation_b200 as nb
from oracle import reference_port as ora
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_edges import inputs, rel_rows
for n in (511, 513, 1025):
  for masses in ("ones", "random"):
    pos, vel, m = inputs(n, 2, torch.float32, seed=n, masses=masses)
    sim = nb.GalaxySimulation(pos.cuda(), vel.cuda(), m.cuda(), precision_mode=nb.PrecisionMode.FLOAT32)
    ref = ora.State(pos, vel, m, mode="float32")
    exact = ora.accelerations_presnap(pos.double(), m.double(), "float64", 0.001, 0.1)
    a = sim.accelerations.cpu().double().numpy(); b = ref.acc.double().numpy(); e = exact.numpy()
    r1 = np.linalg.norm(a-b,axis=1)/np.linalg.norm(b,axis=1); r2 = np.linalg.norm(a-e,axis=1)/np.linalg.norm(e,axis=1); r3 = np.linalg.norm(b-e,axis=1)/np.linalg.norm(e,axis=1)
    i = r1.argmax()
    print(n, masses, "mine-vs-ref %.3e at %d | mine-vs-exact %.3e | ref-vs-exact %.3e | |a|=%.3e" % (r1.max(), i, r2[i], r3[i], np.linalg.norm(e[i])), "median |a| %.3e" % np.median(np.linalg.norm(e,axis=1)))
