"""sopht_b200: B200 (sm_100a) implementation of SophT's Eulerian flow time step.

Hand-written CUDA kernels behind the reference's own factory API; see DESIGN.md.
"""

__version__ = "0.1.0"
