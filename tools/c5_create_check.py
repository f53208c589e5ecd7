import sys, os, time, ctypes
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from sopht_b200 import _lib
from sopht_b200.numeric.eulerian_grid_ops.poisson_solvers import _reflected_axis
lib = _lib.load()
torch.cuda.set_device(0)
n = 1024
dx = np.float32(1.0 / n)
m = _reflected_axis(1.0, dx, n, np.float32)
h = ctypes.c_void_p()
torch.cuda.synchronize(); t0 = time.perf_counter()
_lib.check(lib.sopht_poisson_slab_create(ctypes.byref(h), 3, n, n, n, 8, 3, float(dx), _lib.double_array(m), _lib.double_array(m), _lib.double_array(m), float(np.float32(1 / (4 * np.pi * dx))), _lib.current_stream()))
torch.cuda.synchronize(); t1 = time.perf_counter()
free, total = torch.cuda.mem_get_info()
print(f"1024^3 rank 3 of 8: slab Poisson handle created in {t1 - t0:.1f} s, device memory in use after {(total - free) / 1e9:.1f} GB")
lib.sopht_poisson_slab_destroy(h)
