import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle
from spnngp_b200 import Engine
from tests.util import make_problem
n, m = 200000, 15
port = pyoracle.Port(); port.set_threads(os.cpu_count())
prob = make_problem(n, m, seed=1, port=port, domain=np.sqrt(n / 5000.0))
eng = Engine(0)
eng.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], m)
th = [5, 1, 6, 0.75]
for nb in (1, 2, 1, 2):
    out = eng.bf_loglik(2, [th] * nb, [prob["beta"]] * nb, slots=[0, 1][:nb])
print(out)
