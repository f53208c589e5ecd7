"""z-slab decomposition of the 3-D unbounded flow step over the GPUs of one box:
one process per GPU, torch.distributed (NCCL over NVLink) for the halo exchanges and the two all-to-all
transposes of the distributed FFT Poisson solve; the compute phases are the same CUDA kernels the
single-GPU path uses."""

from .peer import PeerArena
from .slab import HaloExchanger, SlabPartition, exchange_halos
from .slab_flow import SlabPeriodicNavierStokesFlowSimulator3D, SlabUnboundedNavierStokesFlowSimulator3D
from .slab_ib import SlabVirtualBoundaryForcing
from .slab_poisson import SlabPeriodicPoissonSolver3D, SlabTransposePlan, SlabUnboundedPoissonSolver3D

__all__ = [
    "HaloExchanger",
    "PeerArena",
    "SlabPartition",
    "SlabPeriodicNavierStokesFlowSimulator3D",
    "SlabPeriodicPoissonSolver3D",
    "SlabTransposePlan",
    "SlabUnboundedNavierStokesFlowSimulator3D",
    "SlabUnboundedPoissonSolver3D",
    "SlabVirtualBoundaryForcing",
    "exchange_halos",
]
