"""21cmfast_b200 -- B200-native (sm_100a) implementation of 21cmFAST's 3-D grid hot path.

InitialConditions -> PerturbedField -> IonizedBox behind the reference's own C entry points
(``ComputeInitialConditions`` / ``ComputePerturbedField`` / ``ComputeIonizedBox``).  The product
is the C-ABI shared library ``csrc/lib21cmfast_b200.so`` (hand-written CUDA + C host glue, see
``include/py21cmfast_b200.h``); this package is the thin Python mirror of the reference's
wrapper/driver layer for that path.  The directory name starts with a digit, so import it with
``importlib.import_module("21cmfast_b200")``.
"""
from .drivers import (brightness_temperature, compute_halo_grid, compute_halobox,  # noqa: F401
                      compute_initial_conditions,
                      compute_ionization_field, get_logspaced_redshifts, perturb_field,
                      run_coeval, run_coeval_parallel)
from .inputs import (AstroOptions, AstroParams, CosmoParams, InputParameters,  # noqa: F401
                     MatterOptions, SimulationOptions)
from .outputs import (BrightnessTemp, HaloBox, InitialConditions, IonizedBox,  # noqa: F401
                      PerturbedField, TsBox)
from ._lib import Backend, BackendError, get_backend  # noqa: F401
from .distributed import SlabGroup, ionize_radius_parallel, perturb_slab_parallel  # noqa: F401

__version__ = "0.1.0"
