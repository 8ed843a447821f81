import sys
sys.path.insert(0, ".")
import numpy as np
import hamilton_b200 as hb
from tests.test_gpu_parity import _random_system
rng = np.random.default_rng(105)
w, f, u, _ = _random_system(rng, 5, 5)
g = hb.mkSystem(w, f, u, n=5)
y = np.c_[rng.uniform(-0.8, 0.8, size=(97, 5)), rng.uniform(-1, 1, size=(97, 5))]
for name, fn in (("ham_eqs", lambda: g.batch_ham_eqs(y)), ("rk4", lambda: g.batch_step(y, 0.01, 3)), ("energies", lambda: g.batch_energies(y))):
    try:
        fn(); print(name, "ok")
    except Exception as ex:
        print(name, "FAILED", ex)
