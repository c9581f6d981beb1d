"""fdapde-core_b200: B200 (sm_100a) implementation of fdaPDE-core's FE assembly + sparse solve hot path.

The product is the C-ABI shared library built from csrc/ (include/fdapde_b200.h).  This Python package is the thin
host-side mirror of the reference's interface for that path (Triangulation, LagrangianBasis, Assembler<FEM,...>,
operator expressions, PDE) used by the tests and the benchmark harness; it only marshals arrays into the C ABI.
There is no CPU fallback: every compute call fails loudly without the CUDA library / a CUDA device.
"""
from .api import (FdbError, lib, lib_path, Triangulation, LagrangianBasis, Assembler, Space, Matrix, Vector, PDE,  # noqa
                  laplacian, diffusion, advection, reaction, dt, SolverOptions, Comm, solve_parabolic, mesh_topology)
from . import api  # noqa
from . import meshes  # noqa
from . import partition  # noqa
from . import meshio  # noqa
